"""Per-kernel-class achieved GB/s (algorithmic bytes / CUDA-event time) on the finest level of the bench hierarchy."""
import os, sys
sys.path.insert(0, os.getcwd())
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
side, poly, cycles = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
mg = make_hierarchy(sides, "dirichlet", poly)
mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
mg.vCycle(2)
ms = mg.time_vcycles(cycles) / cycles
mg.enable_timers(True); mg.reset_timers(); mg.vCycle(cycles)
out = []
for l in range(len(sides) - 1, -1, -1):
    t = mg.timers(l)
    out.append("L%d(%d) " % (l, sides[l] ** 2) + " ".join("%s %.0fGB/s(%.2fms)" % (k, v["bytes"] / max(v["ms"], 1e-9) / 1e6, v["ms"] / cycles) for k, v in t.items() if v["ms"] > 0))
try:
    cc = mg.grid(-1).colour_counts()
    out.append("finest colours %d sizes max %d min %d" % (len(cc), max(cc), min(cc)))
except Exception as e:
    out.append("colour counts unavailable: %s" % e)
print("side %d poly %d env %s: %.2f ms/cycle | %s" % (side, poly, {k: v for k, v in os.environ.items() if k.startswith("MMG_")}, ms, " | ".join(out)))
