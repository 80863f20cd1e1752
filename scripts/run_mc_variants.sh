timeout 400 python -m pytest tests/test_gpu_vcycle.py tests/test_gpu_operators.py -m gpu -q -x 2>&1 | tail -2
for e in "MMG_X=1" "MMG_MC_SMALL=0"; do
  env $e timeout 200 python scripts/kernel_rates.py 2000 4 5 2>&1 | tail -1 | cut -c1-1400
done
