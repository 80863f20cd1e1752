timeout 300 python -m pytest tests/test_gpu_vcycle.py tests/test_gpu_operators.py -m gpu -q -x 2>&1 | tail -2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_vcycle_check.py 400 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
MMG_DIST_PEER=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_vcycle_check.py 400 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --skip-lex --skip-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-2500
