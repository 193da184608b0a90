#!/bin/bash
# round 2, first GPU visit: the existing GPU suite + where every BASELINE workload stands before any kernel work.
# Run with: gpurun --timeout 1500 -- 'bash profiles/r2a_probe.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
free -g >> gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-converge"
for w in config2_r1 config2_human_se config2_shuffled config2_100m config3_pe_100m config5_full; do
  timeout 600 $B --workload $w > gpurun_out/r2a_bench_$w.log 2>&1
done
for w in config2_human_se config2_shuffled; do
  EMSAR_ORDER=tid timeout 600 $B --workload $w > gpurun_out/r2a_bench_${w}_tidorder.log 2>&1
done
EMSAR_EM_MODE=barrier timeout 600 $B --workload config5_full > gpurun_out/r2a_bench_config5_full_barrier.log 2>&1
EMSAR_EM_MODE=pipe timeout 600 $B --workload config5_full > gpurun_out/r2a_bench_config5_full_pipe.log 2>&1
tail -n 3 gpurun_out/r2a_pytest.log
for f in gpurun_out/r2a_bench_*.log; do echo "== $f"; tail -c 1500 $f; echo; done
