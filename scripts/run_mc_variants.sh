timeout 300 python -m pytest tests/test_gpu_vcycle.py -m gpu -q -x -k "variants or fast_arith" 2>&1 | tail -2
for e in "MMG_X=1" "MMG_MC_FLOW_ROWS=2"; do
  env $e timeout 200 python scripts/kernel_rates.py 2000 4 5 2>&1 | tail -1 | cut -c1-1000
done
