#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_c.log 2>&1; echo "gputests rc=$?"; tail -4 gpurun_out/r02_gputests_c.log | cut -c1-250
timeout 900 python bench.py --side 4000 --steps 10 --warmup 3 --skip-cpu --skip-lex > gpurun_out/r02_bench_16M_n1.json 2> gpurun_out/r02_bench_16M_n1.err; echo "bench 16M rc=$?"
cut -c1-1500 gpurun_out/r02_bench_16M_n1.json; tail -2 gpurun_out/r02_bench_16M_n1.err
timeout 900 python bench.py --fine-poly 6 --steps 10 --warmup 3 --skip-cpu --skip-lex > gpurun_out/r02_bench_4M_p6.json 2> gpurun_out/r02_bench_4M_p6.err; echo "bench p6 rc=$?"
cut -c1-1500 gpurun_out/r02_bench_4M_p6.json; tail -2 gpurun_out/r02_bench_4M_p6.err
