#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_final.log 2>&1; echo "gputests rc=$?"; tail -4 gpurun_out/r02_gputests_final.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','scaling')}, 'e2e', d['e2e'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic')}, d['roofline']['kernel'][:30], d['roofline']['per_class_GBps'], d['roofline']['class_shares'])
print('solve', d.get('solve')); print('lex', d.get('lexicographic')); print('cpu', d.get('cpu_baseline',{}).get('value')); print('check', d.get('check')); print('clocks', d.get('clocks'))
PY
