#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest lstm rc=$?"; tail -4 gpurun_out/pytest_lstm.log
for hs in 1 3; do for B in 256 100 1024; do echo "halfsplit=$hs"; timeout 200 python tools/lstm_probe.py --B $B --time --halfsplit $hs 2>&1 | grep -v Warn | grep "time:\|bwd: dgates" ; done; done
timeout 200 python tools/lstm_probe.py --B 256 --prof --halfsplit 1 2>&1 | grep -v Warn | grep "prof bwd"
