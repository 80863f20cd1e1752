#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_vcycle_check.py 700 4 2>&1 | grep "world" | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'][:40], d.get('solve'), d.get('check'))
PY
tail -2 gpurun_out/r02_bench_n2b.err
