#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_gpu.py -x -q > gpurun_out/r2g_pytest_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_small.log; tail -n 5 gpurun_out/r2g_pytest_small.log
for w in config2_human_se config5_full; do timeout 600 python profiles/trace_psum.py $w > gpurun_out/r2g_trace_$w.log 2>&1; head -8 gpurun_out/r2g_trace_$w.log; done
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for w in config2_human_se config2_r1 config2_shuffled config2_100m config5_full; do
  timeout 900 $B --workload $w > gpurun_out/r2g_bench_$w.log 2>&1; echo "$w: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2g_bench_$w.log) $(grep -a -o '"frac": [0-9.]*' gpurun_out/r2g_bench_$w.log | head -1)"
done
for kb in 160 192; do
  EMSAR_EM_SMEM_KB=$kb timeout 600 $B --workload config2_human_se > gpurun_out/r2g_bench_config2_smem$kb.log 2>&1; echo "smem $kb: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2g_bench_config2_smem$kb.log)"
done
timeout 1200 python -m pytest tests/test_full_size_gpu.py -x -q > gpurun_out/r2g_pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_full.log; tail -n 6 gpurun_out/r2g_pytest_full.log
