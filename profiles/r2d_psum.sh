#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2d_pytest_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_small.log; tail -n 5 gpurun_out/r2d_pytest_small.log
for w in config2_human_se config5_full; do timeout 600 python profiles/trace_psum.py $w > gpurun_out/r2d_trace_$w.log 2>&1; cat gpurun_out/r2d_trace_$w.log; done
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
for w in config2_human_se config5_full; do
  timeout 900 $B --workload $w > gpurun_out/r2d_bench_$w.log 2>&1; tail -c 700 gpurun_out/r2d_bench_$w.log; echo
done
timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q -k "config5 or shuffled" > gpurun_out/r2d_pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_full.log; tail -n 8 gpurun_out/r2d_pytest_full.log
