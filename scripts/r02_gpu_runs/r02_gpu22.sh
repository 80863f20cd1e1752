#!/bin/bash
# L2-ahead prefetch distance of the colour-barrier TMA sweep (tiles of 64 rows), 4M rows n=37
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 6 ';MMG_TMA_L2AHEAD=222;MMG_TMA_L2AHEAD=444;MMG_TMA_L2AHEAD=888;MMG_TMA_L2AHEAD=1776;MMG_TMA_L2AHEAD=3552' > gpurun_out/r02_l2ahead.txt 2>&1
echo rc=$?
cut -c1-420 gpurun_out/r02_l2ahead.txt
