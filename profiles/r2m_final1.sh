#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_gpu.py tests/test_integration_gpu.py -x -q > gpurun_out/r2m_pytest_small.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2m_pytest_small.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_n1.log 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2m_bench_n1.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --workload config2_r1 > gpurun_out/r2m_bench_config2_r1.log 2>&1
P="python bench.py --steps 1 --warmup 1 --em-iters 20 --no-e2e --no-cpu-baseline --no-converge --no-extras"
$P > gpurun_out/r2m_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2m.csv $P > gpurun_out/r2m_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2m $P > gpurun_out/r2m_ncu_full.log 2>&1
P5="python bench.py --workload config5_full --steps 1 --warmup 1 --em-iters 5 --no-e2e --no-cpu-baseline --no-converge --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_em_psum -c 1 -o gpurun_out/prof_em_r2m_config5 $P5 > gpurun_out/r2m_ncu_full5.log 2>&1
ls -la gpurun_out/*r2m*.ncu-rep
tail -c 1500 gpurun_out/r2m_bench_n1.log
