#!/bin/bash
# A/B of round-2 switches on ONE box (train step only, 40 graph replays each)
mkdir -p gpurun_out
for cfg in "" "--lstm-stasync 0" "--lstm-halfsplit 1" "--lstm-halfsplit 1 --lstm-stasync 0" "--chain" ""; do
  timeout 300 python bench.py --steps 40 --warmup 5 --train-only --no-loss-check $cfg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.4f ms  launches %d' % ('$cfg' or 'default', d['ms_per_step'], d['launches_per_step']))"
done
for sa in 1 0; do timeout 200 python tools/lstm_probe.py --B 256 --time --stasync $sa --halfsplit 0 2>&1 | grep "time:"; done
for sa in 1 0; do timeout 200 python tools/lstm_probe.py --B 1024 --time --stasync $sa --halfsplit 0 2>&1 | grep "time:"; done
