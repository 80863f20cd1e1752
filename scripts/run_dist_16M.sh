N=${1:-4}
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --side 4000 --steps 5 --warmup 3 --skip-lex --skip-cpu > gpurun_out/bench_16M_n$N.json 2> gpurun_out/bench_16M_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_16M_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_16M_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','setup_s')}, d['e2e']['value'], d.get('solve'), d['roofline']['frac'], d.get('comm',{}).get('messages'))
PY
free -g | head -2
