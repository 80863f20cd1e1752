#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nproc
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('impl','value','ms_per_step','steps','warmup')}); print(d['solve']); print(d['cpu_baseline']['sample'][:500])
PY
tail -2 gpurun_out/r02_bench_reference.err
