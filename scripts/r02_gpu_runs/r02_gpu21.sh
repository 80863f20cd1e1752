#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vcycle.py -x -q -k "variants or fast_arithmetic" > gpurun_out/r02_solo_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_solo_tests.log | cut -c1-250
SWEEP_ALL_LEVELS=1 timeout 900 python scripts/sweep_kernels.py 2000 4 5 ';MMG_MC_SOLO=0;MMG_MC_SOLO_MAX_ROWS=20000' 2>&1 | cut -c1-330
