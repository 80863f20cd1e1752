"""Block-lexicographic smoother: convergence and time per cycle vs block size on the device-built hierarchy."""
import os, sys, time
sys.path.insert(0, os.getcwd())
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
side, poly, cycles = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
mg = make_hierarchy(sides, "dirichlet", poly)
for B in [int(b) for b in sys.argv[4:]]:
    for l in range(len(sides)): mg.grid(l).values_ = 0 * mg.grid(l).values_
    mg.set_smoother(capi.BLOCK_LEXICOGRAPHIC); mg.set_block_size(B)
    n0 = len(mg.residuals_)
    mg.vCycle(1)
    ms = mg.time_vcycles(cycles - 1) / (cycles - 1)
    h = mg.residuals_[n0:]
    mg.enable_timers(True); mg.reset_timers(); mg.vCycle(1); t = {l: round(mg.timers(l)["sor"]["ms"], 2) for l in range(len(sides))}; mg.enable_timers(False)
    print("side %d poly %d B %d colours(fine) %d: %.1f ms/cycle; sor ms by level %s; hist %s" % (side, poly, B, mg.grid(-1).block_colouring()[0], ms, t,
          " ".join("%d:%.2e" % (i, h[i]) for i in (1, 5, 10, 20, 30, 40) if i < len(h))), flush=True)
