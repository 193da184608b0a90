#!/bin/bash
# 1 GPU: the whole -m gpu suite on the code after r2m (complete index image, set decomposition on the device, restart rounds), smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size_gpu.py > gpurun_out/r2n_pytest_small.log 2>&1; echo "small rc=$?"; tail -n 8 gpurun_out/r2n_pytest_small.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2n_smoke.log
timeout 1500 python -m pytest tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest_full.log 2>&1; echo "full rc=$?"; tail -n 8 gpurun_out/r2n_pytest_full.log
