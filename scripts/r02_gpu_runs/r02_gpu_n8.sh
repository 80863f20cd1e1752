#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_4M_n$N.json 2> gpurun_out/r02_bench_4M_n$N.err; echo "bench 4M n$N rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --side 4000 --steps 10 --warmup 3 > gpurun_out/r02_bench_16M_n$N.json 2> gpurun_out/r02_bench_16M_n$N.err; echo "bench 16M n$N rc=$?"
python - <<PY
import json
for f in ('gpurun_out/r02_bench_4M_n$N.json','gpurun_out/r02_bench_16M_n$N.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:36], d.get('solve'), 'setup', round(d['setup_s'],1), d['roofline'].get('class_shares'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -2 gpurun_out/r02_bench_16M_n$N.err
