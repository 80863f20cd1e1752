"""Residual histories of the device-built hierarchy for a given side / smoother / omega (convergence evidence for DESIGN.md)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
side, poly, smoother, omega, cycles = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], float(sys.argv[4]), int(sys.argv[5])
jitter = float(sys.argv[6]) if len(sys.argv) > 6 else 0.3
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
mg = make_hierarchy(sides, "dirichlet", poly, jitter=jitter)
mg.set_smoother(capi.MULTICOLOUR if smoother == "mc" else capi.LEXICOGRAPHIC)
mg.set_arithmetic(capi.ARITH_FAST if smoother == "mc" else capi.ARITH_REFERENCE_ORDER)
mg.set_omega(omega)
t0 = time.time(); mg.vCycle(cycles); dt = time.time() - t0
h = mg.residuals_
idx = [i for i in (1, 2, 5, 10, 20, 30, 40, 50, 60, 80, 100) if i < cycles]
print("side %d poly %d %s omega %.2f jitter %.2f: %.1f ms/cycle; hist %s" % (side, poly, smoother, omega, jitter, 1e3 * dt / cycles, " ".join("%d:%.2e" % (i, h[i]) for i in idx)), flush=True)
