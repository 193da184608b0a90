#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_gpu.py tests/test_integration_gpu.py -x -q > gpurun_out/r2k_pytest_small.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2k_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for w in config2_human_se config5_full config2_r1 config2_shuffled config2_100m config3_pe_100m; do
  timeout 900 $B --workload $w > gpurun_out/r2k_bench_$w.log 2>&1; echo "$w: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2k_bench_$w.log) $(grep -a -o '"frac": [0-9.]*' gpurun_out/r2k_bench_$w.log | head -1)"
done
timeout 600 python profiles/trace_psum.py config5_full > gpurun_out/r2k_trace_config5.log 2>&1; head -8 gpurun_out/r2k_trace_config5.log
timeout 1200 python -m pytest tests/test_full_size_gpu.py -x -q > gpurun_out/r2k_pytest_full.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2k_pytest_full.log
