"""Profiling driver: build the bench hierarchy on the device, then run V-cycles in the chosen mode (for ncu)."""
import os, sys
sys.path.insert(0, os.getcwd())
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
side, poly, mode, cycles = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
mg = make_hierarchy(sides[::-1], "dirichlet", poly)
if mode == "mc":
    mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
else:
    mg.set_smoother(capi.LEXICOGRAPHIC)
mg.vCycle(cycles)
print("ok", mg.residuals_[-1])
