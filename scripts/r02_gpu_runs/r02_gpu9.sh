#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_vcycle_check.py 300 4 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-1200 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
