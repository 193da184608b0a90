#!/bin/bash
# 8 GPUs: the bench line exactly as the driver launches it (weak -M scaling, m64 = BASELINE configs[3], class_sharded legs)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2l_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n8.log 2> gpurun_out/r2l_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r2l_bench_n8.err; tail -n 5 gpurun_out/r2l_bench_n8.err; tail -c 3000 gpurun_out/r2l_bench_n8.log
