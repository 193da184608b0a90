set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1i.log 2>&1; tail -5 gpurun_out/pytest_s2g.log
timeout 600 python bench.py > gpurun_out/bench_r1i.log 2>&1; echo "rc=$?" >> gpurun_out/bench_r1i.log; tail -2 gpurun_out/bench_r1i.log | cut -c1-600
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-converge"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1i.csv $B --em-iters 20 > gpurun_out/ncu_launch_r1i.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_em_persistent -s 1 -c 1 -f -o gpurun_out/prof_em_r1i $B --em-iters 20 --no-e2e > gpurun_out/ncu_full_r1i.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_em_persistent -s 1 -c 1 --csv --log-file gpurun_out/traffic_r1i_200.csv $B --em-iters 200 --no-e2e > gpurun_out/ncu_t200.log 2>&1
tail -3 gpurun_out/traffic_r1i_200.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_count -c 2 --csv --log-file gpurun_out/count_r1i.csv $B --em-iters 20 --no-e2e > gpurun_out/ncu_count.log 2>&1
tail -8 gpurun_out/count_r1i.csv | cut -c1-300
