#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_gpu.py -x -q > gpurun_out/r2i_pytest_small.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2i_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for w in config2_human_se config5_full; do
  timeout 900 $B --workload $w > gpurun_out/r2i_bench_$w.log 2>&1; echo "$w: $(grep -a -o '"us_per_iter": [0-9.]*' gpurun_out/r2i_bench_$w.log) $(grep -a -o '"frac": [0-9.]*' gpurun_out/r2i_bench_$w.log | head -1)"
done
P="python bench.py --steps 1 --warmup 1 --em-iters 20 --no-e2e --no-cpu-baseline --no-converge --no-extras"
$P > gpurun_out/r2i_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i.csv $P > gpurun_out/r2i_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2i $P > gpurun_out/r2i_ncu_full.log 2>&1
P5="python bench.py --workload config5_full --steps 1 --warmup 1 --em-iters 5 --no-e2e --no-cpu-baseline --no-converge --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2i_config5 $P5 > gpurun_out/r2i_ncu_full5.log 2>&1
ls -la gpurun_out/*r2i*.ncu-rep
