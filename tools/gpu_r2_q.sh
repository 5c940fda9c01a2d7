#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_q.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_q.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 40 --warmup 5 --train-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train %.4f ms (%d launches)' % (d['ms_per_step'], d['launches_per_step']))"
