N=${1:-4}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_vcycle_check.py 400 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-lex --skip-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-3000
