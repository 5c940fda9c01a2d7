#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/x3_probe.py > gpurun_out/x3_probe.log 2>&1; echo "x3 probe rc=$?"; cat gpurun_out/x3_probe.log | tail
timeout 900 python tools/parity_diag.py 256 > gpurun_out/diag256.log 2>&1; echo "diag256 rc=$?"; grep -v Warning gpurun_out/diag256.log | tail -120
timeout 600 python tools/parity_diag.py 4 default,exact_engine_simt,all_simt > gpurun_out/diag4.log 2>&1; echo "diag4 rc=$?"; grep -v Warning gpurun_out/diag4.log | tail -60
