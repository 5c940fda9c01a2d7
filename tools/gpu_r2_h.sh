#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest chain rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|AssertionError" gpurun_out/pytest_chain.log | head -30
timeout 600 python tools/chain_time.py > gpurun_out/chain_time.log 2>&1; echo "chain_time rc=$?"; grep -v Warn gpurun_out/chain_time.log | tail -20
timeout 600 python tools/chain_time.py --trace > gpurun_out/chain_trace.log 2>&1; echo "chain_trace rc=$?"; grep -v Warn gpurun_out/chain_trace.log | tail -20
