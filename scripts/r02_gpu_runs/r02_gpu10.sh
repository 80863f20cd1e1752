#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py -x -q > gpurun_out/r02_dropin.log 2>&1; echo "dropin rc=$?"
tail -25 gpurun_out/r02_dropin.log | cut -c1-300
