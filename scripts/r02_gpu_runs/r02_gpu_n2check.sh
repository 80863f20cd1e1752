#!/bin/bash
# 2 GPUs after the L1-first probe: parity against one GPU (bit-identical at 400^2, 1e-11 at 2000^2), then the 4M bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version"; }
run 29511 scripts/dist_vcycle_check.py 400 4 | grep "^world\|^rank\|finest" | cut -c1-330
run 29512 scripts/dist_vcycle_check.py 2000 4 3 | grep "^world\|^rank\|finest" | cut -c1-330
bash scripts/r02_gpu_runs/r02_gpu_n8.sh 2 2>&1 | grep -v "^Setting\|^\*\*\*" | cut -c1-600
