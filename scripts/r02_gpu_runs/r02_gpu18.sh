#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | tail -4; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','scaling')}, 'e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic','kernel')})
print('solve', d.get('solve')); print('lex', d.get('lexicographic')); print('cpu', d.get('cpu_baseline')); print('check', d.get('check')); print('clocks', d.get('clocks'))
PY
tail -2 gpurun_out/r02_bench_default.err
