#!/bin/bash
# co-resident CTAs take neighbouring shares (MMG_TMA_SM_SHARES) on top of the static share
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 8 ';MMG_TMA_SM_SHARES=0;MMG_TMA_SM_SHARES=1;MMG_TMA_SM_SHARES=0;MMG_TMA_STATIC_8THS=0' > gpurun_out/r02_static_share3.txt 2>&1
echo rc=$?
cut -c1-300 gpurun_out/r02_static_share3.txt
