#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vcycle.py tests/test_gpu_at_size.py -x -q -k "variants or fast_multicolour or multicolour_history_at_1m" > gpurun_out/r02_tflow_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02_tflow_tests.log | cut -c1-250
timeout 900 python scripts/sweep_kernels.py 2000 4 5 'MMG_TMA_FLOW=0;;MMG_MC_FLOW_MAX_ROWS=0;MMG_MC_FLOW_MAX_ROWS=100000;MMG_TMA_CTAS=4,MMG_TMA_SMEM_KB=224' > gpurun_out/r02_sweep6.log 2>&1; echo "sweep rc=$?"
cut -c1-480 gpurun_out/r02_sweep6.log
