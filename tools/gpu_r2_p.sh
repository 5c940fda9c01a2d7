#!/bin/bash
# 2 GPUs: the 2-rank NCCL equivalence test, smoke(), and the N=2 bench line of the current build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_dispatch.py -q -m gpu -k "nccl or rank" > gpurun_out/pytest_n2.log 2>&1; echo "pytest 2-rank rc=$?"; tail -2 gpurun_out/pytest_n2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_r2_n2.sh 2 ""
