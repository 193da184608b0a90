#!/bin/bash
# round 2: full-size parity tests + the bench line with every leg (N = 1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_full_size_gpu.py -x -q --durations=10 > gpurun_out/r2b_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_full.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench_n1.log 2> gpurun_out/r2b_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2b_bench_n1.err
tail -n 15 gpurun_out/r2b_pytest_full.log; tail -n 5 gpurun_out/r2b_bench_n1.err; tail -c 3000 gpurun_out/r2b_bench_n1.log
