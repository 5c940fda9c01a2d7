#!/bin/bash
mkdir -p gpurun_out
OLD=$PWD/tools/probes/_bin/libdeer_b200_v20.so
timeout 600 python -m pytest tests/test_gpu_fused_passes.py -q -x > gpurun_out/pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -15 gpurun_out/pytest_fused.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_n.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_n.log | head
run() { # label, env, flags
  DEER_B200_LIB=$2 timeout 300 python bench.py --steps 40 --warmup 5 --train-only --no-loss-check $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.4f ms  launches %d' % ('$1', d['ms_per_step'], d['launches_per_step']))"
}
run "new" "" ""
run "new" "" ""
