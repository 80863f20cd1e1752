import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
for side in [int(a) for a in sys.argv[2:]]:
    kind = sys.argv[1]
    mg = make_hierarchy([side // 2, side], kind, 4, cloud="hex")
    g = mg.grid(-1)
    nl, lev = g.lex_levels()
    print(kind, side, "nodes", g.getSize(), "A", g.A_size, "dag levels", nl, flush=True)
    for arith, name in ((capi.ARITH_FAST, "fast"), (capi.ARITH_REFERENCE_ORDER, "exact")):
        mg.set_arithmetic(arith)
        try:
            g.values_ = np.zeros(g.A_size)
            t = time.perf_counter(); g.sor(capi.LEXICOGRAPHIC); mg.sync(); dt = time.perf_counter() - t
            t = time.perf_counter(); g.sor(capi.LEXICOGRAPHIC); mg.sync(); dt = time.perf_counter() - t
            print("  ", name, "5 sweeps: %.3f s" % dt, capi.last_kernel(0), flush=True)
        except Exception as e:
            print("  ", name, "ERR", e, flush=True)
