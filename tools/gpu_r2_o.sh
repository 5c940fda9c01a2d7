#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_o.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_o.log | head
