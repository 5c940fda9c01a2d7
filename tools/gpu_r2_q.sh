#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_passes.py -q -k "input_projection" > gpurun_out/pytest_xin.log 2>&1; echo "pytest xin rc=$?"; tail -5 gpurun_out/pytest_xin.log
run() { # label, flags
  timeout 300 python bench.py --steps 40 --warmup 5 --train-only --no-loss-check $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.4f ms  launches %d' % ('$1', d['ms_per_step'], d['launches_per_step']))"
}
timeout 300 python tools/xin_probe.py 2>&1 | grep "B="
run "xin" ""
run "no xin" "--lstm-xin 0"
run "xin" ""
run "no xin" "--lstm-xin 0"
