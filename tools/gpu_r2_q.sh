#!/bin/bash
mkdir -p gpurun_out
timeout 140 python bench.py --no-pooled --no-strong --no-cpu-baseline > gpurun_out/bench_v24.log 2> gpurun_out/bench_v24.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v24.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step'], d['inference']['ms_per_step'], d['inference']['value'], d['inference']['e2e']['ms_per_step'])
r=d['roofline']; print(r['frac'], r['nig_head_loss']['frac'], r['attn_pool']['frac'], r['lstm_recurrence']['fwd_us_per_step'], r['lstm_recurrence']['bwd_us_per_step'], d['config'].get('loss_check_vs_cpu_port'))
PY
