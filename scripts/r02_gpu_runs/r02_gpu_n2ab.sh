#!/bin/bash
# 2 GPUs, same box: L1-first probe of the barrier-free peer sweep on / off / on
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
for v in 1 0 1 0; do
  MMG_TMAFLOW_L1=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus $N --steps 10 --warmup 3 --skip-lex --skip-cpu --skip-solve 2>/dev/null | grep "^{" > gpurun_out/ab_$v.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('L1=$v', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'frac', round(d['roofline']['frac'],3), d['roofline'].get('class_shares'))
PY
done
