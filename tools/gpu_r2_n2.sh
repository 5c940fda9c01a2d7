#!/bin/bash
# N-rank bench lines: default exchange vs bucketed (overlapped) exchange; usage: gpu_r2_n2.sh NGPUS "ctas list"
N=${1:-2}; CT=${2:-8}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-strong --no-pooled "$@" > gpurun_out/bench_n${N}_$tag.log 2> gpurun_out/bench_n${N}_$tag.err; echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n${N}_$tag.log') if l.startswith('{')][-1])
    print('$tag', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['ms_per_step'], 'overlap', d['exchange_overlap'], d['exchange_overlap_check'])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_n${N}_$tag.err').read()[-2500:])
PY
}
run default
for c in $CT; do run overlap_c$c --overlap-exchange --exchange-ctas $c; done
