# Round-end validation on one B200: tests, default bench, launch list and one full ncu capture of the dominant kernel.
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_r01_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv \
  python bench.py --steps 2 --warmup 3 --skip-solve --skip-lex --skip-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sor_mc_flow -s 0 -c 1 -o gpurun_out/r01_sor_mc_flow_1M -f \
  python scripts/profile_cycle.py 2000 4 mc 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
