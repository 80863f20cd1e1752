#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests.log 2>&1; echo "gputests rc=$?"
tail -30 gpurun_out/r02_gputests.log | cut -c1-250
timeout 900 python scripts/sweep_kernels.py 2000 4 5 ';MMG_MC_RESIDENT=0' > gpurun_out/r02_sweep5.log 2>&1; echo "sweep rc=$?"
cut -c1-330 gpurun_out/r02_sweep5.log
