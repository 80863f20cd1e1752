import sys, os, time; sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
side = int(sys.argv[1]); poly = int(sys.argv[2])
sides=[side]
while sides[-1] > 16: sides.append((sides[-1]+1)//2)
sides = sides[::-1]
mg = make_hierarchy(sides, "dirichlet", poly)
mg.set_smoother(capi.LEXICOGRAPHIC)
mg.vCycle(1)
ms = mg.time_vcycles(3)/3
mg.enable_timers(True); mg.reset_timers(); mg.vCycle(1)
t = {l: mg.timers(l)["sor"]["ms"] for l in range(len(sides))}
print("side", side, "poly", poly, "env", {k:v for k,v in os.environ.items() if k.startswith("MMG_")}, "ms/cycle %.1f" % ms, "sor ms by level", {k: round(v,2) for k,v in t.items()}, "hist", mg.residuals_[:3])
