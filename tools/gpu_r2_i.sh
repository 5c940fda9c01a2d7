#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_i.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-pooled --no-strong --no-cpu-baseline > gpurun_out/bench_i.log 2> gpurun_out/bench_i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_i.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_i.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'], d['inference']['launches_per_step'])
    r=d['roofline']; print(r['frac'], r['us_per_launch'], 'nig', r['nig_head_loss']['frac'], r['nig_head_loss']['us_per_call'], 'nig L2', r['nig_head_loss_l2_resident_operands']['frac'], r['nig_head_loss_l2_resident_operands']['us_per_call'], 'lstm', r['lstm_recurrence']['fwd_us_per_step'], r['lstm_recurrence']['bwd_us_per_step'], 'pool', r['attn_pool']['frac'])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python tools/step_timeline.py > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"; tail -20 gpurun_out/timeline.log
