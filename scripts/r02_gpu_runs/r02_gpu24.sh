#!/bin/bash
# repeatability of the static-share gain: defaults vs 8/8, twice each, n=37 and n=70
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 8 ';MMG_TMA_STATIC_8THS=8;MMG_TMA_STATIC_8THS=0;MMG_TMA_STATIC_8THS=8;MMG_TMA_STATIC_8THS=5' > gpurun_out/r02_static_share2.txt 2>&1
echo rc=$?
python scripts/sweep_kernels.py 2000 6 6 ';MMG_TMA_STATIC_8THS=8;MMG_TMA_STATIC_8THS=0;MMG_TMA_STATIC_8THS=8' > gpurun_out/r02_static_share2_p6.txt 2>&1
echo rc=$?
cut -c1-300 gpurun_out/r02_static_share2.txt gpurun_out/r02_static_share2_p6.txt
