"""One hierarchy, many kernel configurations: per-level / per-class CUDA-event times of the multicolour V-cycle for each
environment in the list, plus a bit-level cross-check of the finest smoother against the first configuration.
usage: sweep_kernels.py SIDE POLY CYCLES 'K=V,K=V;K=V;...'   (configurations separated by ';', empty = defaults)"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy

side, poly, cycles = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in (sys.argv[4] if len(sys.argv) > 4 else "").split(";")]
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
mg = make_hierarchy(sides, "dirichlet", poly)
mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
fine = mg.grid(-1)
rng = np.random.default_rng(5)
v0 = 1e-3 * rng.standard_normal(fine.A_size)
known = set(k for c in configs for k in c)
ref_hash = None
for cfg in configs:
    for k in known: os.environ.pop(k, None)
    os.environ.update(cfg)
    try:
        fine.values_ = v0
        fine.sor(capi.MULTICOLOUR)
        kern = capi.last_kernel(0)
        h = hashlib.sha1(fine.values_.tobytes()).hexdigest()[:12]
        if ref_hash is None: ref_hash = h
        r = fine.residual(); spk = capi.last_kernel(1)
        for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
        mg.vCycle(2)
        ms = mg.time_vcycles(cycles) / cycles
        mg.enable_timers(True); mg.reset_timers(); mg.vCycle(cycles); mg.sync()
        parts = []
        for l in range(len(sides) - 1, max(len(sides) - 4, -1), -1):
            t = mg.timers(l)
            parts.append("L%d " % l + " ".join("%s %.0fGB/s %.3fms" % (k, v["bytes"] / max(v["ms"], 1e-9) / 1e6, v["ms"] / cycles) for k, v in t.items() if v["ms"] > 0))
        tot = mg.timers(-1)
        coarse_sor = sum(mg.timers(l)["sor"]["ms"] for l in range(len(sides) - 2)) / cycles
        if os.environ.get("SWEEP_ALL_LEVELS"):
            print("   per level sor ms/cycle: " + " ".join("L%d(%d) %.3f" % (l, sides[l] ** 2, mg.timers(l)["sor"]["ms"] / cycles) for l in range(len(sides)))
                  + " | spmv ms/cycle: " + " ".join("L%d %.3f" % (l, sum(mg.timers(l)[k]["ms"] for k in ("residual", "restrict", "prolong", "other")) / cycles) for l in range(len(sides))), flush=True)
        mg.enable_timers(False)
        print("CFG %s | %s %s | same_bits %s | %.3f ms/cycle launches/cycle %d | coarse-tail sor %.3f ms | %s" % (cfg, kern, spk, h == ref_hash, ms, sum(v["launches"] for v in tot.values()) // cycles, coarse_sor, " | ".join(parts)), flush=True)
    except Exception as e:
        print("CFG %s FAILED: %s" % (cfg, e), flush=True)
