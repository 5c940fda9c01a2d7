#!/bin/bash
# same-box A/B: v20 library (9-warp BPTT) vs the 8-warp BPTT build (train step only, 40 graph replays each)
mkdir -p gpurun_out
# (the other build: `git worktree add /tmp/old <commit>`, build() there, copy its libdeer_b200.so here; not tracked)
OLD=$PWD/tools/probes/_bin/libdeer_b200_v20.so
run() { # label, env, flags
  DEER_B200_LIB=$2 timeout 300 python bench.py --steps 40 --warmup 5 --train-only --no-loss-check $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.4f ms  launches %d' % ('$1', d['ms_per_step'], d['launches_per_step']))"
}
run "v20" $OLD ""
run "8-warp BPTT" "" ""
run "v20" $OLD ""
run "8-warp BPTT" "" ""
run "8-warp BPTT --tf32-pair 0" "" "--tf32-pair 0"
run "v20 --tf32-pair 0" $OLD "--tf32-pair 0"
