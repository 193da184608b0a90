#!/bin/bash
# r2q: index construction on the device (csrc/build.cu) - parity tests, then device vs host builder vs reference on a larger transcriptome
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_build_index_gpu.py -x -q > gpurun_out/r2q_pytest_build.log 2>&1; echo "build tests rc=$?"; tail -n 15 gpurun_out/r2q_pytest_build.log
timeout 900 python profiles/build_bench.py > gpurun_out/r2q_build_bench.log 2>&1; echo "bench rc=$?"; tail -n 12 gpurun_out/r2q_build_bench.log
