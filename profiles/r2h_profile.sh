#!/bin/bash
mkdir -p gpurun_out
for st in 0 3; do EMSAR_PS_STAGE=$st timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "solve_matches or iterates" > gpurun_out/r2h_pytest_stage$st.log 2>&1; echo "stage $st rc=$?"; tail -n 2 gpurun_out/r2h_pytest_stage$st.log; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2h_pytest_small.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2h_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for st in 0 1 2 3; do
  EMSAR_PS_STAGE=$st timeout 600 $B --workload config2_human_se > gpurun_out/r2h_bench_config2_stage$st.log 2>&1; echo "config2 stage $st: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2h_bench_config2_stage$st.log)"
done
for st in 0 2; do
  EMSAR_VERBOSE=1 EMSAR_PS_STAGE=$st timeout 900 $B --workload config5_full > gpurun_out/r2h_bench_config5_stage$st.log 2>&1; echo "config5 stage $st: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2h_bench_config5_stage$st.log) $(grep -a 'not used' gpurun_out/r2h_bench_config5_stage$st.log | head -2)"
done
P="python bench.py --steps 1 --warmup 1 --em-iters 20 --no-e2e --no-cpu-baseline --no-converge --no-extras"
$P > gpurun_out/r2h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2h.csv $P > gpurun_out/r2h_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2h $P > gpurun_out/r2h_ncu_full.log 2>&1
EMSAR_PS_STAGE=0 ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2h_stage0 $P > gpurun_out/r2h_ncu_full0.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
