#!/bin/bash
# r2p: convergence poll by the first warps (no registers held across M items); per-CTA feature / time tables for the cost model
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2p_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/r2p_pytest_parity.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-converge"
for m in 1 0; do
  EMSAR_PS_MORDER=$m timeout 300 $B --workload config2_human_se > gpurun_out/r2p_c2_m$m.log 2>&1
  echo "config2 morder=$m: $(grep -o '"us_per_iter": [0-9.]*' gpurun_out/r2p_c2_m$m.log | head -1)"
  EMSAR_PS_MORDER=$m timeout 600 $B --workload config5_full > gpurun_out/r2p_c5_m$m.log 2>&1
  echo "config5 morder=$m: $(grep -o '"us_per_iter": [0-9.]*' gpurun_out/r2p_c5_m$m.log | head -1)"
done
for w in config2_human_se config2_shuffled config5_full; do
  timeout 300 python profiles/trace_psum.py $w > gpurun_out/r2p_trace_$w.log 2>&1; tail -n 11 gpurun_out/r2p_trace_$w.log | head -6
done
