#!/bin/bash
# round 2, GPU call 1: at-size parity tests + first sweep of the TMA-fed kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu1_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_at_size.py -x -q > gpurun_out/r02_atsize.log 2>&1; echo "atsize rc=$?"
tail -5 gpurun_out/r02_atsize.log
timeout 900 python scripts/sweep_kernels.py 2000 4 5 'MMG_MC_TMA=0,MMG_SPMV_TMA=0;;MMG_TMA_ROWS=2;MMG_TMA_CTAS=3,MMG_TMA_SMEM_KB=180;MMG_TMA_CTAS=1,MMG_TMA_SMEM_KB=100;MMG_TMA_CTAS=4,MMG_TMA_SMEM_KB=200,MMG_TMA_STAGES=3;MMG_TMA_DYNAMIC=0;MMG_TMA_STAGES=2;MMG_SPMV_TMA_ROWS=1;MMG_SPMV_TMA_LPR=16;MMG_MC_FLOW_MAX_ROWS=0' > gpurun_out/r02_sweep1.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r02_sweep1.log | cut -c1-900
