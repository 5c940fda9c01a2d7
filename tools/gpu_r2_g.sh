#!/bin/bash
# round-2 GPU call G (re-entry): whole -m gpu suite, bench line, launch lists train/infer, step timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_g.log | head -40
timeout 700 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_g.log 2> gpurun_out/bench_g.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_g.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_g.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'], d['inference']['launches_per_step']); print(d.get('strong_scaling'))
    for r in d.get('pooled_model_sweep') or []: print(r)
    r=d['roofline']; print(r['frac'], r['us_per_launch'], r['nig_head_loss']['frac'], r['lstm_recurrence'], r['attn_pool']['frac']); print(d['config'].get('loss_check_vs_cpu_port'))
except Exception as e:
    print("bench parse failed", e)
PY
for mode in train infer; do
  B=256; [ $mode = infer ] && B=1024
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python tools/profile_step.py $mode $B 3 > gpurun_out/ncu_$mode.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_$mode.csv > gpurun_out/launches_${mode}_summary.txt; head -60 gpurun_out/launches_${mode}_summary.txt
done
timeout 300 python tools/step_timeline.py > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"; tail -25 gpurun_out/timeline.log
