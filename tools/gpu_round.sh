#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','launches_per_step')}); print(d['inference'])
PY
tail -3 gpurun_out/bench.err
python tools/profile_step.py train 256 3 > gpurun_out/plain_train.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py train 256 3 > gpurun_out/ncu_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.txt; head -30 gpurun_out/launches_train_summary.txt
