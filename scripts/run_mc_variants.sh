for e in "MMG_MC_SMALL_MAX=512" "MMG_MC_SMALL_MAX=0"; do
  env $e timeout 200 python scripts/kernel_rates.py 2000 4 5 2>&1 | tail -1 | cut -c1-80
  env $e timeout 200 python scripts/kernel_rates.py 2000 4 5 2>&1 | tail -1 | grep -o "L2(3969).*" | cut -c1-400
done
