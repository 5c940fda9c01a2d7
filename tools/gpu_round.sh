#!/bin/bash
# One GPU round: full bench line (driver contract), eager launch lists under ncu for the train / inference step.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','launches_per_step')}); print(d['inference']); print(d.get('pooled_model_sweep'))
print(d.get('cpu_baseline')); r=d['roofline']; print(r['frac'], r['us_per_launch'], r['nig_head_loss']['frac'], r['lstm_recurrence'])
PY
for mode in train infer; do
  B=256; [ $mode = infer ] && B=1024
  python tools/profile_step.py $mode $B 3 > gpurun_out/plain_$mode.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python tools/profile_step.py $mode $B 3 > gpurun_out/ncu_$mode.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_$mode.csv > gpurun_out/launches_${mode}_summary.txt; head -30 gpurun_out/launches_${mode}_summary.txt
done
