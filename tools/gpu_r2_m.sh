#!/bin/bash
mkdir -p gpurun_out
# (the other build: `git worktree add /tmp/old <commit>`, build() there, copy its libdeer_b200.so here; not tracked)
OLD=$PWD/tools/probes/_bin/libdeer_b200_v20.so
timeout 600 python -m pytest tests/test_gpu_lstm.py -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest lstm rc=$?"; tail -1 gpurun_out/pytest_lstm.log
DEER_B200_LIB=$OLD timeout 200 python tools/lstm_probe.py --B 256 --time 2>&1 | grep "time:"
timeout 200 python tools/lstm_probe.py --B 256 --time 2>&1 | grep "time:"
timeout 200 python tools/coresidency_probe.py 2>&1 | grep -v Warn | tail -2
run() { # label, env, flags
  DEER_B200_LIB=$2 timeout 300 python bench.py --steps 40 --warmup 5 --train-only --no-loss-check $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.4f ms  launches %d' % ('$1', d['ms_per_step'], d['launches_per_step']))"
}
run "v20" $OLD ""
run "8-warp BPTT" "" ""
run "v20" $OLD ""
run "8-warp BPTT" "" ""
