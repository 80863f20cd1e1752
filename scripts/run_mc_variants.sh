timeout 300 python -m pytest tests/test_gpu_operators.py tests/test_gpu_vcycle.py -m gpu -q -x 2>&1 | tail -3
for e in "MMG_MC_PACKED=0" "MMG_MC_ORDER=0" "MMG_MC_ROWS=2" "MMG_MC_ROWS=4" "MMG_MC_ROWS=1"; do
  env $e timeout 200 python scripts/kernel_rates.py 2000 4 5 2>&1 | tail -1 | cut -c1-420
done
