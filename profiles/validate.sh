set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final2.log 2>&1; tail -3 gpurun_out/pytest_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; tail -1 gpurun_out/smoke_final2.log
timeout 600 python bench.py > gpurun_out/bench_final2.log 2>&1; tail -1 gpurun_out/bench_final2.log | cut -c1-300
