#!/bin/bash
# where the colour-barrier sweep takes over from the barrier-free one (MMG_MC_FLOW_MAX_ROWS), with the static tile shares and the L1-first probe in place
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 8 ';MMG_MC_FLOW_MAX_ROWS=500000;MMG_MC_FLOW_MAX_ROWS=200000;MMG_MC_FLOW_MAX_ROWS=1500000' > gpurun_out/r02_flow_threshold2.txt 2>&1
echo rc=$?
cut -c1-640 gpurun_out/r02_flow_threshold2.txt
