#!/bin/bash
# N = 8: the driver's command line (weak scaling headline + strong-scaling block in the same line)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8.log') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'], 'strong', d.get('strong_scaling'), 'inference', d['inference']['value'])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n8.err').read()[-1500:])
PY
