#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/gemm_h16_probe.py 2>&1 | grep -v block0
out=gpurun_out/lstm_probe.log; : > $out
run() { echo "=== $*" >> $out; timeout 180 python tools/lstm_probe.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run --ts 1 --tile 16 --B 256 --T 300 --time --prof
run --ts 1 --tile 32 --B 1024 --T 300 --time
grep -E "prof|time" $out
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['inference']); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('achieved','frac','us_per_call','fwd_us_per_step','bwd_us_per_step')}) for k,v in d['roofline'].items() if k in ('achieved','frac','frac_of_tf32_peak','us_per_launch','nig_head_loss','lstm_recurrence')})
PY
tail -3 gpurun_out/bench.err
