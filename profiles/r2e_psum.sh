#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2e_pytest_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_small.log; tail -n 5 gpurun_out/r2e_pytest_small.log
timeout 600 python profiles/trace_psum.py config2_human_se > gpurun_out/r2e_trace_config2.log 2>&1; cat gpurun_out/r2e_trace_config2.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
EMSAR_VERBOSE=1 timeout 900 $B --workload config5_full > gpurun_out/r2e_bench_config5_full.log 2>&1; grep -a "not used" gpurun_out/r2e_bench_config5_full.log | head -3; tail -c 600 gpurun_out/r2e_bench_config5_full.log; echo
timeout 600 $B --workload config2_human_se > gpurun_out/r2e_bench_config2_base.log 2>&1; tail -c 500 gpurun_out/r2e_bench_config2_base.log; echo
for cc in 0 16 40; do for cr in 6 30; do
  EMSAR_PS_COST_CLASS=$cc EMSAR_PS_COST_ROW=$cr timeout 600 $B --workload config2_human_se > gpurun_out/r2e_bench_config2_cc${cc}_cr${cr}.log 2>&1
  echo "cc=$cc cr=$cr: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2e_bench_config2_cc${cc}_cr${cr}.log)"
done; done
timeout 600 python profiles/trace_psum.py config5_full > gpurun_out/r2e_trace_config5.log 2>&1; cat gpurun_out/r2e_trace_config5.log
