#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python scripts/bench_configs.py config3 1860 4 0 70 > gpurun_out/r02_config3_4M.json 2> gpurun_out/r02_config3_4M.err; echo "config3 4M rc=$?"; cut -c1-1600 gpurun_out/r02_config3_4M.json; tail -2 gpurun_out/r02_config3_4M.err
PPE_MODE=lex_exact timeout 1500 python scripts/bench_configs.py config4 1316 4 2 70 > gpurun_out/r02_config4_2M.json 2> gpurun_out/r02_config4_2M.err; echo "config4 2M rc=$?"; cut -c1-1500 gpurun_out/r02_config4_2M.json; tail -2 gpurun_out/r02_config4_2M.err
