set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final1.log 2>&1; tail -3 gpurun_out/pytest_final1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final1.log 2>&1; tail -1 gpurun_out/smoke_final1.log
timeout 600 python bench.py > gpurun_out/bench_final1.log 2>&1; tail -1 gpurun_out/bench_final1.log | cut -c1-300
timeout 900 python bench.py --impl reference --ref-binary > gpurun_out/bench_ref_final1.log 2>&1; tail -1 gpurun_out/bench_ref_final1.log | cut -c1-1500
timeout 300 python bench.py --workload small --steps 3 --warmup 3 --em-iters 2000 --no-cpu-baseline > gpurun_out/bench_small_final1.log 2>&1; tail -1 gpurun_out/bench_small_final1.log | cut -c1-200
timeout 400 python bench.py --workload config5_stress --steps 3 --warmup 3 --em-iters 500 --no-cpu-baseline > gpurun_out/bench_c5_final1.log 2>&1; tail -1 gpurun_out/bench_c5_final1.log | cut -c1-200
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-converge"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_count -c 2 --csv --log-file gpurun_out/count_r1i.csv $B --em-iters 20 --no-e2e > gpurun_out/ncu_count.log 2>&1
grep k_count gpurun_out/count_r1i.csv | awk -F'","' '{print $13, $14, $15}'
