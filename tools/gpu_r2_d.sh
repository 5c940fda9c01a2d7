#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_lstm.py -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest x3/lstm rc=$?"; tail -15 gpurun_out/pytest_d.log
timeout 900 python tools/parity_diag.py 256 default,exact_scorers+conv > gpurun_out/diag256.log 2>&1; echo "diag256 rc=$?"; grep -v Warning gpurun_out/diag256.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_all.log | head -30
timeout 600 python bench.py --steps 20 --warmup 3 --no-pooled --no-strong > gpurun_out/bench_d.log 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_d.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'], d['inference']['launches_per_step'])
    r=d['roofline']; print(r['frac'], r['us_per_launch'], r['nig_head_loss']['frac'], r['lstm_recurrence']['fwd_us_per_step'], r['lstm_recurrence']['bwd_us_per_step'], r['lstm_recurrence']['fwd_hbm']['frac'], r['attn_pool']['frac'])
except Exception as e:
    print("bench parse failed", e)
PY
