#!/bin/bash
# 2 GPUs: the 2-rank NCCL equivalence test and the N=2 bench line (its self-check runs the early-exchange path too)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_dispatch.py -q -m gpu -k "nccl or rank" > gpurun_out/pytest_n2.log 2>&1; echo "pytest 2-rank rc=$?"; tail -2 gpurun_out/pytest_n2.log
bash tools/gpu_r2_n2.sh 2 ""
tail -5 gpurun_out/bench_n2_default.err
