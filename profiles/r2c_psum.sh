#!/bin/bash
# round 2: first run of k_em_psum: sanitizer on a small case, parity suites, then where the workloads stand
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "test_solve_matches_oracle and (se_small or hubs)" > gpurun_out/r2c_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2c_sanitizer.log
tail -n 30 gpurun_out/r2c_sanitizer.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_gpu.py -x -q > gpurun_out/r2c_pytest_small.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest_small.log
tail -n 30 gpurun_out/r2c_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for w in config2_human_se config2_r1 config2_shuffled config5_full config2_100m; do
  timeout 600 $B --workload $w > gpurun_out/r2c_bench_$w.log 2>&1
  tail -c 900 gpurun_out/r2c_bench_$w.log; echo
done
EMSAR_EM_MODE=legacy timeout 600 $B --workload config2_human_se > gpurun_out/r2c_bench_config2_human_se_legacy.log 2>&1
timeout 1500 python -m pytest tests/test_full_size_gpu.py -x -q > gpurun_out/r2c_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest_full.log
tail -n 20 gpurun_out/r2c_pytest_full.log
