#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_facade.py tests/test_gpu_vcycle.py tests/test_gpu_fracstep.py tests/test_gpu_operators.py -x -q > gpurun_out/r02_gputests_b.log 2>&1; echo "gputests rc=$?"
tail -5 gpurun_out/r02_gputests_b.log | cut -c1-250
timeout 900 python scripts/bench_configs.py config3 900 4 0 40 > gpurun_out/r02_config3_1M.json 2> gpurun_out/r02_config3_1M.err; echo "config3 1M rc=$?"; cut -c1-1800 gpurun_out/r02_config3_1M.json; tail -3 gpurun_out/r02_config3_1M.err
PPE_MODE=lex_exact timeout 900 python scripts/bench_configs.py config4 400 4 3 40 > gpurun_out/r02_config4_180k.json 2> gpurun_out/r02_config4_180k.err; echo "config4 rc=$?"; cut -c1-1500 gpurun_out/r02_config4_180k.json; tail -3 gpurun_out/r02_config4_180k.err
