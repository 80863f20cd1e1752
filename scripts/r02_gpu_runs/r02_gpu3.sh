#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests.log 2>&1; echo "gputests rc=$?"
tail -4 gpurun_out/r02_gputests.log
timeout 900 python scripts/sweep_kernels.py 2000 6 3 'MMG_MC_TMA=0,MMG_SPMV_TMA=0;;MMG_TMA_LPR=16,MMG_SPMV_TMA_LPR=16;MMG_TMA_LPR=16,MMG_SPMV_TMA_LPR=16,MMG_TMA_ROWS=1,MMG_TMA_CTAS=4,MMG_SPMV_TMA_ROWS=1,MMG_SPMV_TMA_CTAS=4' > gpurun_out/r02_sweep3_p6.log 2>&1; echo "sweep p6 rc=$?"
cut -c1-420 gpurun_out/r02_sweep3_p6.log
timeout 900 python bench.py --steps 20 --warmup 5 --skip-lex > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02_bench_a.json; tail -3 gpurun_out/r02_bench_a.err
