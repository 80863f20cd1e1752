#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_facade.py tests/test_gpu_vcycle.py tests/test_gpu_fracstep.py -x -q > gpurun_out/r02_gputests_b.log 2>&1; echo "gputests rc=$?"
tail -12 gpurun_out/r02_gputests_b.log | cut -c1-250
timeout 900 python scripts/bench_configs.py config3 900 4 0 40 > gpurun_out/r02_config3_1M.json 2> gpurun_out/r02_config3_1M.err; echo "config3 1M rc=$?"; cut -c1-1800 gpurun_out/r02_config3_1M.json; tail -3 gpurun_out/r02_config3_1M.err
timeout 900 python scripts/bench_configs.py config4 660 4 5 60 > gpurun_out/r02_config4_500k.json 2> gpurun_out/r02_config4_500k.err; echo "config4 500k rc=$?"; cut -c1-1500 gpurun_out/r02_config4_500k.json; tail -3 gpurun_out/r02_config4_500k.err
