#!/bin/bash
# N = 8 bench lines: default exchange vs early (overlapped) exchange
mkdir -p gpurun_out
for mode in default overlap; do
  extra=""; [ $mode = overlap ] && extra="--overlap-exchange"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-strong $extra > gpurun_out/bench_n8_$mode.log 2> gpurun_out/bench_n8_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8_$mode.log') if l.startswith('{')][-1])
    print('$mode', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], 'overlap', d['exchange_overlap'], d['exchange_overlap_check'])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n8_$mode.err').read()[-1500:])
PY
done
