#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_q.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_q.log | head
for x in "" ""; do
timeout 400 python bench.py --steps 30 --warmup 3 --no-pooled --no-strong --no-cpu-baseline $x 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$x]: train %.4f ms e2e %.4f inference %.4f ms (%d launches) loss check %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['inference']['ms_per_step'], d['inference']['launches_per_step'], d['config'].get('loss_check_vs_cpu_port')))"
done
