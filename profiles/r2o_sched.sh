#!/bin/bash
# A/B of the k_em_psum changes of r2o: E tiles dealt statically (EMSAR_PS_SCHED), remote-owner rows first in the M-phase (EMSAR_PS_MORDER);
# the convergence read behind the first M item and the halo slots in registers are in both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2o_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/r2o_pytest_parity.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-converge"
for combo in 00 10 01 11; do
  s=${combo:0:1}; m=${combo:1:1}
  EMSAR_PS_SCHED=$s EMSAR_PS_MORDER=$m timeout 300 $B --workload config2_human_se > gpurun_out/r2o_c2_$combo.log 2>&1
  echo "config2 sched=$s morder=$m: $(grep -o '"us_per_iter": [0-9.]*' gpurun_out/r2o_c2_$combo.log | head -1)"
done
for combo in 00 11; do
  s=${combo:0:1}; m=${combo:1:1}
  EMSAR_PS_SCHED=$s EMSAR_PS_MORDER=$m timeout 600 $B --workload config5_full > gpurun_out/r2o_c5_$combo.log 2>&1
  echo "config5 sched=$s morder=$m: $(grep -o '"us_per_iter": [0-9.]*' gpurun_out/r2o_c5_$combo.log | head -1)"
done
timeout 300 python profiles/trace_psum.py config2_human_se > gpurun_out/r2o_trace_config2.log 2>&1; tail -n 12 gpurun_out/r2o_trace_config2.log
