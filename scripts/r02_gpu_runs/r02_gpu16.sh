#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vcycle.py tests/test_gpu_at_size.py -x -q -k "variants or fast_multicolour or multicolour_history_at_1m" > gpurun_out/r02_tmaflow_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_tmaflow_tests.log | cut -c1-250
timeout 900 python scripts/sweep_kernels.py 2000 4 5 'MMG_MC_TMAFLOW=0;;MMG_MC_TMAFLOW_MIN_ROWS=10000;MMG_TMAFLOW_ROWS=1,MMG_TMAFLOW_CTAS=4;MMG_TMAFLOW_CTAS=2,MMG_TMAFLOW_STAGES=3;MMG_MC_FLOW_MAX_ROWS=5000000' > gpurun_out/r02_sweep7.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import re
for l in open('gpurun_out/r02_sweep7.log'):
    if not l.startswith('CFG'): print(l[:200]); continue
    cfg=l.split('|')[0]
    ms=re.search(r'([\d.]+) ms/cycle', l).group(1)
    lv=re.findall(r'(L\d) sor (\d+)GB/s ([\d.]+)ms', l)
    ct=re.search(r'coarse-tail sor ([\d.]+)', l).group(1)
    print(cfg, l.split('|')[1].strip()[:40], 'same', 'True' in l.split('|')[2], ms, 'ms/cycle tail', ct, lv)
PY
