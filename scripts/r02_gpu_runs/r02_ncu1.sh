#!/bin/bash
# ncu captures of the round-2 kernels (one gpurun call; every ncu command follows a plain run of the same command).
# Only text summaries (ncu -i ... --page raw, scripts/ncu_summary.py) and the report of the dominant kernel travel back.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=/tmp/ncu_r02; mkdir -p $T
NCU="ncu --set full --clock-control none"
A="python scripts/profile_cycle.py 2000 4 mc 1"
$A > gpurun_out/ncu_plain_a.log 2>&1 && $NCU --import-source on -k regex:'k_sor_mc_tma' -c 2 -f -o gpurun_out/r02_sor_mc_tma_4M $A > gpurun_out/ncu_a.log 2>&1; echo "A rc=$?"
python scripts/ncu_summary.py gpurun_out/r02_sor_mc_tma_4M.ncu-rep > gpurun_out/r02_sor_mc_tma_4M_ncu.txt
$NCU -k regex:'k_spmv_tma' -c 12 -f -o $T/spmv $A > gpurun_out/ncu_a2.log 2>&1; echo "A2 rc=$?"
python scripts/ncu_summary.py $T/spmv.ncu-rep > gpurun_out/r02_spmv_tma_4M_ncu.txt
B="python scripts/profile_cycle.py 1000 4 mc 1"
$B > gpurun_out/ncu_plain_b.log 2>&1 && $NCU -k regex:'k_sor_mc_flow|k_sor_mc_resident|k_sor_mc_small' -c 6 -f -o $T/flow $B > gpurun_out/ncu_b.log 2>&1; echo "B rc=$?"
python scripts/ncu_summary.py $T/flow.ncu-rep > gpurun_out/r02_sor_mc_flow_resident_1M_ncu.txt
Cc="python scripts/profile_cycle.py 1000 4 lex 1"
$Cc > gpurun_out/ncu_plain_c.log 2>&1 && $NCU -k regex:'k_sor_lex_chunk' -s 4 -c 2 -f -o $T/lex $Cc > gpurun_out/ncu_c.log 2>&1; echo "C rc=$?"
python scripts/ncu_summary.py $T/lex.ncu-rep > gpurun_out/r02_sor_lex_chunk_1M_ncu.txt
D="python scripts/profile_cycle.py 500 6 mc 1"
$D > gpurun_out/ncu_plain_d.log 2>&1 && $NCU -k regex:'k_knn|k_weights' -s 10 -c 6 -f -o $T/asm $D > gpurun_out/ncu_d.log 2>&1; echo "D rc=$?"
python scripts/ncu_summary.py $T/asm.ncu-rep > gpurun_out/r02_assembly_250k_p6_ncu.txt
E="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-lex --skip-solve"
$E > gpurun_out/ncu_plain_e.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bench.csv $E > gpurun_out/ncu_e.log 2>&1; echo "E rc=$?"
du -sh gpurun_out; ls -la gpurun_out | tail -15
