#!/bin/bash
mkdir -p gpurun_out
for x in "" "--no-text-stream" "" "--no-text-stream"; do
timeout 400 python bench.py --steps 30 --warmup 3 --no-pooled --no-strong --no-cpu-baseline --no-loss-check $x 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$x]: train %.4f ms  inference %.4f ms (%d launches)' % (d['ms_per_step'], d['inference']['ms_per_step'], d['inference']['launches_per_step']))"
done
timeout 300 python tools/step_timeline.py 2>&1 | tail -19
