#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest lstm rc=$?"; tail -4 gpurun_out/pytest_lstm.log
for B in 256 100 1024; do timeout 200 python tools/lstm_probe.py --B $B --time 2>&1 | grep -v Warn | grep "time:" ; done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_k.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-pooled --no-strong --no-cpu-baseline > gpurun_out/bench_k.log 2> gpurun_out/bench_k.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_k.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_k.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'])
    r=d['roofline']; print('lstm', r['lstm_recurrence']['fwd_us_per_step'], r['lstm_recurrence']['bwd_us_per_step'], r['lstm_recurrence']['fwd_hbm']['frac'], r['lstm_recurrence']['bwd_hbm']['frac'])
except Exception as e:
    print("bench parse failed", e)
PY
