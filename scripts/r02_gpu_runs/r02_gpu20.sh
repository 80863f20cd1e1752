#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python scripts/sweep_kernels.py 2000 4 5 ';MMG_MC_TMAFLOW_MIN_ROWS=2000;MMG_MC_TMAFLOW_MIN_ROWS=2000,MMG_TMAFLOW_CTAS=2;MMG_MC_TMAFLOW_MIN_ROWS=2000,MMG_TMAFLOW_CTAS=1,MMG_TMAFLOW_SMEM_KB=60;MMG_MC_TMAFLOW_MIN_ROWS=10000,MMG_TMAFLOW_CTAS=2' > gpurun_out/r02_sweep9.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import re
for l in open('gpurun_out/r02_sweep9.log'):
    if not l.startswith('CFG'): print(l[:200]); continue
    cfg=l.split('|')[0]
    ms=re.search(r'([\d.]+) ms/cycle', l).group(1)
    lv=re.findall(r'(L\d) sor (\d+)GB/s ([\d.]+)ms', l)
    ct=re.search(r'coarse-tail sor ([\d.]+)', l).group(1)
    print(cfg, 'same', 'True' in l.split('|')[2], ms, 'ms/cycle tail', ct, lv)
PY
