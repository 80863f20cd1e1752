#!/bin/bash
# ncu capture of the lexicographic dependency-DAG sweep on the 4M-row level (plain run of the same command first)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=/tmp/ncu_r02; mkdir -p $T
Cc="python scripts/profile_cycle.py 2000 4 lex 1"
$Cc > gpurun_out/ncu3_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none -k regex:'k_sor_lex_chunk' -c 1 -f -o $T/lex4m $Cc > gpurun_out/ncu3.log 2>&1; echo "rc=$?"
python scripts/ncu_summary.py $T/lex4m.ncu-rep > gpurun_out/r02_sor_lex_chunk_4M_ncu.txt
cut -c1-160 gpurun_out/r02_sor_lex_chunk_4M_ncu.txt | head -30; tail -3 gpurun_out/ncu3.log | cut -c1-200
