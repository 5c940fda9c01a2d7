#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
out=gpurun_out/lstm_probe.log; : > $out
run() { echo "=== $*" >> $out; timeout 180 python tools/lstm_probe.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run --ts 1 --tile 16 --B 256 --T 300 --time --prof
run --ts 1 --tile 32 --B 1024 --T 300 --time
grep -E "prof|time" $out
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','launches_per_step')}); print(d['inference'])
PY
