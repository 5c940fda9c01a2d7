#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_q.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_q.log | head
for B in 1024; do timeout 200 python tools/lstm_probe.py --B $B --time 2>&1 | grep -v Warn | grep "time:" ; done
for x in 1 1; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-pooled --no-strong --no-cpu-baseline --no-loss-check --lstm-xin $x 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lstm-xin $x: train %.4f ms  inference %.4f ms (%d launches)' % (d['ms_per_step'], d['inference']['ms_per_step'], d['inference']['launches_per_step']))"
done
