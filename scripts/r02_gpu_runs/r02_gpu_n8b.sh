#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=$1
for thr in 300000 1200000; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --skip-solve --partition-threshold $thr > gpurun_out/r02_bench_4M_n${N}_thr$thr.json 2> gpurun_out/thr.err; echo "thr $thr rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_4M_n${N}_thr$thr.json').read().strip().splitlines()[-1])
print('thr $thr', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['roofline'].get('class_shares'), d.get('comm',{}).get('partitioned_levels'))
PY
done
