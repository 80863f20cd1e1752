#!/bin/bash
# ncu capture of the barrier-free TMA sweep on a 1M-row level (plain run of the same command first)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=/tmp/ncu_r02; mkdir -p $T
B="python scripts/profile_cycle.py 1000 4 mc 1"
$B > gpurun_out/ncu2_plain.log 2>&1 && ncu --set full --clock-control none -k regex:'k_sor_mc_tma_flow' -c 3 -f -o $T/tflow $B > gpurun_out/ncu2.log 2>&1; echo "rc=$?"
python scripts/ncu_summary.py $T/tflow.ncu-rep > gpurun_out/r02_sor_mc_tma_flow_1M_ncu.txt
cut -c1-160 gpurun_out/r02_sor_mc_tma_flow_1M_ncu.txt | head -60
