"""Throughput (multicolour) mode: cycles to 1e-8 for per-level relaxation factors (coarsest ... finest) on the bench workload."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy, grid_props
side = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
L = len(sides)
mg = make_hierarchy(sides, "dirichlet", 4)
mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST)
def run(omegas, max_cycles=260):
    for l in range(L):
        p = grid_props(4 if l == L - 1 else 3, omega=omegas[l])
        mg.grid(l).set_props(p)
        mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    mg.sync(); t = time.perf_counter()
    n, r = mg.solve(1e-8, max_cycles); mg.sync()
    dt = time.perf_counter() - t
    print("omegas %s -> %d cycles, %.3f s, final %.2e %s" % (" ".join("%.2f" % o for o in omegas), n, dt, r, "" if r < 1e-8 else "NOT CONVERGED"), flush=True)
schedules = [[0.8] * L]
for oc in (1.0, 1.2):
    schedules.append([oc] * (L - 1) + [0.8])
    schedules.append([oc] * (L - 2) + [0.8, 0.8])
    schedules.append([oc] * (L - 3) + [1.0, 0.9, 0.8] if L >= 4 else [0.8] * L)
schedules.append([1.2] * (L - 3) + [1.1, 1.0, 0.8])
schedules.append([1.3] * (L - 3) + [1.2, 1.0, 0.8])
schedules.append([1.2] * (L - 1) + [0.9])
schedules.append([1.1] * (L - 1) + [0.85])
for s in schedules: run(s)
