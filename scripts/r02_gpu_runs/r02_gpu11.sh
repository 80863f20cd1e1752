#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_operators.py tests/test_gpu_vcycle.py tests/test_gpu_fracstep.py -x -q -k "sor or history or neumann or mixed or ppe or time_step or quirk" > gpurun_out/r02_lexneu.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02_lexneu.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_at_size.py -x -q -k "neumann" > gpurun_out/r02_lexneu2.log 2>&1; echo "at-size rc=$?"
tail -6 gpurun_out/r02_lexneu2.log | cut -c1-250
timeout 600 python scripts/probe_lex_neumann.py mixed 450 900 2>&1 | tail -8
