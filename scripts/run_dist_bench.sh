N=${1:-2}
if [ "$N" = "2" ]; then timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_vcycle_check.py 400 4 2>&1 | grep "^world" | cut -c1-200; fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-lex --skip-cpu > gpurun_out/bench_r01_n$N.json 2> gpurun_out/bench_r01_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r01_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks')}, d['e2e']['value'], d.get('solve'), d['roofline']['frac'], d.get('comm',{}).get('messages'))
PY
