#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_j.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-pooled --no-strong --no-cpu-baseline > gpurun_out/bench_j.log 2> gpurun_out/bench_j.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_j.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_j.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'], d['inference']['launches_per_step'])
    r=d['roofline']; print(r['frac'], r['us_per_launch'], 'nig', r['nig_head_loss']['frac'], r['nig_head_loss']['us_per_call'], r['nig_head_loss'].get('l2_window_mb_sweep_us'), 'lstm', r['lstm_recurrence']['fwd_us_per_step'], r['lstm_recurrence']['bwd_us_per_step'], 'pool', r['attn_pool']['frac'])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python tools/step_timeline.py > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"; tail -19 gpurun_out/timeline.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_cluster -c 1 -o gpurun_out/lstm_fwd_r2 python tools/lstm_probe.py --B 256 --skip-bwd > gpurun_out/ncu_lstm_fwd.log 2>&1; echo "ncu rc=$?"
