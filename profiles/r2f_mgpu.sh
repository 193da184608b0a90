#!/bin/bash
# 2 GPUs: multi-GPU parity tests, then the bench line with the class_sharded legs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2f_pytest_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_mgpu.log; tail -n 25 gpurun_out/r2f_pytest_mgpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n2.log 2> gpurun_out/r2f_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2f_bench_n2.err; tail -n 8 gpurun_out/r2f_bench_n2.err; tail -c 4000 gpurun_out/r2f_bench_n2.log
