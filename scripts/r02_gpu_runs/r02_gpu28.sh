#!/bin/bash
# the polyDeg-6 bench line with the in-run check against tests/golden/bench_2000_p6.json
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python bench.py --fine-poly 6 --steps 20 --warmup 5 > gpurun_out/r02_bench_4M_p6.json 2> gpurun_out/r02_bench_4M_p6.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_4M_p6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic')}, d['roofline']['kernel'][:30], d['roofline']['per_class_GBps'])
print('solve', d.get('solve')); print('check', d.get('check')); print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('lexicographic'))
PY
