#!/bin/bash
# 8-warp BPTT kernel: LSTM tests, kernel timing, co-residency probe (BPTT beside the 128x128 TF32 GEMM)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest lstm rc=$?"; tail -4 gpurun_out/pytest_lstm.log
for B in 256 1024; do timeout 200 python tools/lstm_probe.py --B $B --time 2>&1 | grep -v Warn | grep "time:" ; done
timeout 200 python tools/coresidency_probe.py 2>&1 | grep -v Warn | tail -4
