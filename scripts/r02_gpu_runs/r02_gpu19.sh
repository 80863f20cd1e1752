#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_operators.py -x -q -k "residual or restrict or prolong" > gpurun_out/r02_spmv_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_spmv_tests.log | cut -c1-250
timeout 900 python scripts/sweep_kernels.py 2000 4 5 ';MMG_SPMV_TMA_CTAS=4,MMG_SPMV_TMA_SMEM_KB=224' > gpurun_out/r02_sweep8.log 2>&1; echo "sweep rc=$?"
cut -c1-520 gpurun_out/r02_sweep8.log
