#!/bin/bash
# Full GPU check of a build: the whole -m gpu suite, the NIG passes alone (timing + instruction counts), a short bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/nig_probe.py 20 22 2>&1 | tail -12
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum -k regex:nig_loss -s 46 -c 4 python tools/nig_probe.py 22 2>&1 | grep -E "nig_loss|inst_executed|duration|issue_active|warps_active|dram__" | cut -c1-120 | head -40
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-pooled > gpurun_out/bench_chk.log 2> gpurun_out/bench_chk.err; tail -c 400 gpurun_out/bench_chk.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_chk.log").read().strip().splitlines()[-1])
print("train/e2e/infer ms", d["ms_per_step"], d["e2e"]["ms_per_step"], d["inference"]["ms_per_step"])
r=d["roofline"]; print("roof", r["us_per_launch"], r["frac"], "nig", r["nig_head_loss"]["us_per_call"], r["nig_head_loss"]["frac"], "lstm", r["lstm_recurrence"]["fwd_us_per_step"], r["lstm_recurrence"]["bwd_us_per_step"])
PY
