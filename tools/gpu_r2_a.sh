#!/bin/bash
# round-2 GPU call A: the whole -m gpu suite (incl. the bench-dispatch parity tests), a bench line, ncu of the LSTM kernels in the real step
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_a.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_a.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','launches_per_step')}); print(d['inference']); print(d.get('strong_scaling')); print(d.get('pooled_model_sweep'))
    print(d.get('cpu_baseline')); r=d['roofline']; print(r['frac'], r['us_per_launch'], r['nig_head_loss']['frac'], r['lstm_recurrence'], r['attn_pool']); print(d['config'].get('loss_check_vs_cpu_port'))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 ncu --set full --clock-control none -k regex:lstm_ -s 4 -c 4 -o gpurun_out/r2_lstm_step python tools/profile_step.py train 256 2 > gpurun_out/ncu_lstm.log 2>&1; echo "ncu rc=$?"
