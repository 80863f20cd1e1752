#!/bin/bash
# first probe of the barrier-free TMA sweep through L1 (MMG_TMAFLOW_L1): per-level times, and the finest level forced through it for the bit check
mkdir -p gpurun_out
python scripts/sweep_kernels.py 2000 4 8 ';MMG_TMAFLOW_L1=0;MMG_TMAFLOW_L1=1;MMG_MC_FLOW_MAX_ROWS=99999999,MMG_TMAFLOW_L1=0;MMG_MC_FLOW_MAX_ROWS=99999999,MMG_TMAFLOW_L1=1' > gpurun_out/r02_flow_l1.txt 2>&1
echo rc=$?
cut -c1-640 gpurun_out/r02_flow_l1.txt
