#!/bin/bash
mkdir -p gpurun_out
for x in "" "--no-pool-rowterm" "" "--no-pool-rowterm"; do
timeout 400 python bench.py --steps 40 --warmup 5 --train-only $x 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$x]: train %.4f ms (%d launches)' % (d['ms_per_step'], d['launches_per_step']))"
done
