"""BASELINE configs 3 and 4 on the Gmsh-like clouds (clouds.hex_square), where the reference's scheme converges with Neumann
and mixed boundaries (it diverges on the jittered lattices beyond ~10k nodes).

  config3  mixed Dirichlet+Neumann Poisson, ~4M nodes: (a) lexicographic omega=1.4 (the only ordering that converges with a
           Neumann boundary) solve to 1e-8 with the reference's stop rule: cycles, seconds, s/cycle; (b) per-cycle throughput
           of the fused multicolour sweep (regularisation row + boundary evaluation inside the launch), which does not converge
           here and is reported as kernel throughput only.
  config4  fractional-step pressure-Poisson (all Neumann) through FractionalStepMultigrid, ~2M nodes, N time steps of the
           Kovasznay driver (FractionalStepSim.cpp:130-147): seconds/step and V-cycles/step.

usage: bench_configs.py config3|config4 SIDE [FINE_POLY] [STEPS] [MAX_CYCLES]
One JSON line per run."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.clouds import hex_square
from meshlessmultigridpoisson_b200.problems import make_hierarchy, make_ppe_grid, fracstep_time_step

which, side = sys.argv[1], int(sys.argv[2])
poly = int(sys.argv[3]) if len(sys.argv) > 3 else 4
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
max_cycles = int(sys.argv[5]) if len(sys.argv) > 5 else 60
sides = [side]
while sides[-1] > 16: sides.append((sides[-1] + 1) // 2)
sides = sides[::-1]
out = {"config": which, "cloud": "hex (Gmsh-like)", "sides": sides, "fine_poly": poly}
t0 = time.time()
if which == "config3":
    mg = make_hierarchy(sides, "mixed", poly, cloud="hex")
    mg.sync()
    fine = mg.grid(-1)
    out.update(nodes=fine.getSize(), setup_s=time.time() - t0)
    # (a) reference-faithful solve
    mg.set_smoother(capi.LEXICOGRAPHIC); mg.set_arithmetic(capi.ARITH_REFERENCE_ORDER)
    mg.vCycle(1); mg.sync()
    for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    n0 = len(mg.residuals_)
    t1 = time.perf_counter(); n, r = mg.solve(1e-8, max_cycles); mg.sync(); dt = time.perf_counter() - t1
    h = mg.residuals_[n0:]
    out["lexicographic"] = {"cycles": n, "seconds": dt, "s_per_cycle": dt / max(n, 1), "final_residual": r, "converged": bool(r < 1e-8),
                            "rate_per_cycle": float((h[-1] / h[min(3, len(h) - 1)]) ** (1.0 / max(1, len(h) - 1 - min(3, len(h) - 1)))) if len(h) > 4 else None,
                            "kernel": capi.last_kernel(0)}
    # (a') the same ordering, MMG_ARITH_FAST: only the regularisation row's in-order sum becomes a tree reduction; three cycles timed
    mg.set_arithmetic(capi.ARITH_FAST)
    for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    mg.vCycle(1); mg.sync()
    t1 = time.perf_counter(); mg.vCycle(3); mg.sync(); dt = time.perf_counter() - t1
    out["lexicographic_fast_arith"] = {"cycles_timed": 3, "s_per_cycle": dt / 3, "kernel": capi.last_kernel(0)}
    # (b) fused multicolour sweep: throughput only
    mg.set_smoother(capi.MULTICOLOUR); mg.set_omega(0.8)
    for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    mg.vCycle(2)
    for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    ms = mg.time_vcycles(3) / 3
    mg.enable_timers(True); mg.reset_timers()
    for l in range(len(sides)): mg.grid(l).values_ = np.zeros(mg.grid(l).A_size)
    mg.vCycle(2); mg.sync()
    t = mg.timers(len(sides) - 1)
    out["multicolour_fused"] = {"ms_per_cycle": ms, "vcycles_per_s": 1e3 / ms, "kernel": capi.last_kernel(0), "converges": False,
                                "finest_GBps": {k: v["bytes"] / max(v["ms"], 1e-9) / 1e6 for k, v in t.items() if v["ms"] > 0 and k != "other"}}
else:
    dt_, mu, rho, tol = 2e-4, 0.025, 1.0, 1e-10                  # FractionalStepSim.cpp:202
    mg = capi.FractionalStepMultigrid()
    for l, s in enumerate(sides):
        x, y = hex_square(s, 1000 + l)
        last = l == len(sides) - 1
        mg.addGrid(make_ppe_grid(x, y, poly if last else 3, dt_, mu, rho, fine=last))
    mg.buildMatrices(); mg.sync()
    fine = mg.grid(-1)
    out.update(nodes=fine.getSize(), setup_s=time.time() - t0, dt=dt_, mu=mu, rho=rho, ppe_tol=tol)
    mode = os.environ.get("PPE_MODE", "lex_fast")
    mg.set_smoother(capi.LEXICOGRAPHIC)
    mg.set_arithmetic(capi.ARITH_REFERENCE_ORDER if mode == "lex_exact" else capi.ARITH_FAST)
    fine.set_uv_bound()
    per_step, cyc = [], []
    for k in range(steps):
        t1 = time.perf_counter()
        n, res = fracstep_time_step(mg, tol, max_cycles)
        mg.sync()
        per_step.append(time.perf_counter() - t1); cyc.append(n)
    out["time_steps"] = {"mode": mode, "steps": steps, "s_per_step": float(np.mean(per_step[1:] or per_step)), "first_step_s": per_step[0], "vcycles_per_step": float(np.mean(cyc)),
                         "vcycles": cyc, "final_fs_residual": res, "ppe_converged_every_step": bool(max(cyc) < max_cycles), "kernel": capi.last_kernel(0)}
print(json.dumps(out), flush=True)
