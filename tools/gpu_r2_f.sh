#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -q -x -s > gpurun_out/pytest_chain.log 2>&1; echo "pytest chain rc=$?"; grep -v Warning gpurun_out/pytest_chain.log | tail -25
