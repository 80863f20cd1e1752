"""Golden residual histories of the benchmark workload, from the CPU oracle alone (no GPU): bench.py asserts its timed
configuration against these in-run.  Takes ~10 minutes on 8 cores at 4M nodes.
usage: python tests/make_bench_golden.py [SIDE] [FINE_POLY]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
poly = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
t0 = time.time()
mg = oracle.make_hierarchy(sides, kind=oracle.KIND_DIRICHLET, fine_poly=poly)
setup_s = time.time() - t0
out = {"sides": sides, "fine_poly": poly, "setup_s": setup_s, "threads": os.cpu_count()}
t0 = time.time(); mg.vcycle(5); out["lexicographic_omega1.4"] = mg.history().tolist() + [mg.residual()]; out["lex_s_per_cycle"] = (time.time() - t0) / 5
for l in range(mg.nlevels): mg.level(l).set_vec(oracle.VEC_VALUES, np.zeros(mg.level(l).A))
mg.set_smoother(1); mg.set_omega(0.8)
n0 = len(mg.history())
t0 = time.time(); mg.vcycle(5); out["multicolour_omega0.8"] = mg.history()[n0:].tolist() + [mg.residual()]; out["mc_s_per_cycle"] = (time.time() - t0) / 5
p = os.path.join(ROOT, "tests", "golden", "bench_%d_p%d.json" % (side, poly))
json.dump(out, open(p, "w"), indent=1)
print(p, out)
