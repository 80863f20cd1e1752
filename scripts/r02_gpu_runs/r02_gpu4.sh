#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_at_size.py -x -q -k "neumann or mirror" > gpurun_out/r02_neumann.log 2>&1; echo "neumann tests rc=$?"
tail -25 gpurun_out/r02_neumann.log | cut -c1-300
