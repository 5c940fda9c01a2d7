#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','launches_per_step')}); print(d['inference']); print(d.get('pooled_model_sweep'))
PY
python tools/profile_step.py infer 1024 3 > gpurun_out/plain_infer.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python tools/profile_step.py infer 1024 3 > gpurun_out/ncu_infer.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infer.csv > gpurun_out/launches_infer_summary.txt; head -24 gpurun_out/launches_infer_summary.txt
