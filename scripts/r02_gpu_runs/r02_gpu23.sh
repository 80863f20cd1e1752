#!/bin/bash
# contiguous static share of the tiles of a phase (in 1/8 of the even share) before the ticket-dealt rest, colour-barrier TMA sweep, 4M rows
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 6 ';MMG_TMA_STATIC_8THS=4;MMG_TMA_STATIC_8THS=6;MMG_TMA_STATIC_8THS=7;MMG_TMA_STATIC_8THS=8;MMG_TMA_STATIC_8THS=8,MMG_TMA_ROWS=1,MMG_TMA_CTAS=4' > gpurun_out/r02_static_share.txt 2>&1
echo rc=$?
cut -c1-330 gpurun_out/r02_static_share.txt
