#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_vcycle.py -x -q -k "neumann or mixed or ppe" > gpurun_out/r02_lexneu3.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_lexneu3.log | cut -c1-200
for pipe in 1 0; do
MMG_LEX_NEUMANN_PIPE=$pipe timeout 900 python scripts/bench_configs.py config3 900 4 0 12 2> gpurun_out/c3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PIPE=$pipe', {k:(round(v['s_per_cycle'],4), v['cycles'], v['kernel']) for k,v in d.items() if isinstance(v,dict) and 's_per_cycle' in v}, 'mc ms/cycle', round(d['multicolour_fused']['ms_per_cycle'],2))"
done
tail -3 gpurun_out/c3.err
