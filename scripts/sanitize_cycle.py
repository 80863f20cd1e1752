"""One small V-cycle through every solve-path kernel family, meant to run under `compute-sanitizer --target-processes all`
(closed on the round-2 GPU pool, so only the plain run was made there: all five variants finite).

Each variant runs in its own child process (several knobs are read once per process):
  default     resident / small / value-flag / TMA-fed value-flag sweeps, TMA-fed SpMV on the finest level
  barrier     MMG_MC_FLOW_MAX_ROWS=0, MMG_MC_RESIDENT=0, MMG_MC_SMALL=0: the colour-barrier TMA sweep on every level
  neumann     Neumann problem on the hex cloud through the barrier TMA sweep (regularisation row, boundary evaluation, overflow tails)
  lex         lexicographic dependency-DAG sweep, Dirichlet and Neumann (auxiliary CTA)
  fracstep    one fractional-step time step (derivative operators, PPE source, corrections)

usage: sanitize_cycle.py [SIDE]            (parent: runs all variants)
       sanitize_cycle.py SIDE VARIANT      (child)
Prints one line per variant with the kernels seen and the residual history; exit code != 0 if a variant fails."""
import os, subprocess, sys
sys.path.insert(0, os.getcwd())

VARIANTS = {
    "default": {},
    "barrier": {"MMG_MC_FLOW_MAX_ROWS": "0", "MMG_MC_RESIDENT": "0", "MMG_MC_SMALL": "0"},
    "neumann": {"MMG_MC_FLOW_MAX_ROWS": "0", "MMG_MC_RESIDENT": "0", "MMG_MC_SMALL": "0"},
    "lex": {},
    "fracstep": {},
}


def sides_of(side):
    s = [side]
    while s[-1] > 16: s.append((s[-1] + 1) // 2)
    return s[::-1]


def child(side, variant):
    import numpy as np
    from meshlessmultigridpoisson_b200 import capi
    from meshlessmultigridpoisson_b200.clouds import hex_square
    from meshlessmultigridpoisson_b200.problems import make_hierarchy, make_ppe_grid, fracstep_time_step
    sides = sides_of(side)
    seen = []
    if variant in ("default", "barrier"):
        mg = make_hierarchy(sides, "dirichlet", 4)
        mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
        mg.vCycle(2); mg.sync(); seen.append(capi.last_kernel(0))
    elif variant == "neumann":
        mg = make_hierarchy(sides, "neumann", 4, cloud="hex")
        mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
        mg.vCycle(2); mg.sync(); seen.append(capi.last_kernel(0))
    elif variant == "lex":
        for kind, cloud in (("dirichlet", "jittered"), ("neumann", "hex")):
            mg = make_hierarchy(sides, kind, 4, cloud=cloud)
            mg.set_smoother(capi.LEXICOGRAPHIC); mg.set_arithmetic(capi.ARITH_REFERENCE_ORDER)
            mg.vCycle(2); mg.sync(); seen.append(capi.last_kernel(0))
    else:
        mg = capi.FractionalStepMultigrid()
        for l, s in enumerate(sides):
            x, y = hex_square(s, 1000 + l)
            last = l == len(sides) - 1
            mg.addGrid(make_ppe_grid(x, y, 4 if last else 3, 2e-4, 0.025, 1.0, fine=last))
        mg.buildMatrices(); mg.sync()
        mg.set_smoother(capi.LEXICOGRAPHIC); mg.set_arithmetic(capi.ARITH_FAST)
        mg.grid(-1).set_uv_bound()
        fracstep_time_step(mg, 1e-10, 2); mg.sync(); seen.append(capi.last_kernel(0))
    h = np.asarray(mg.residuals_, dtype=float)
    assert np.all(np.isfinite(h)), h
    print("sanitize %-8s sides=%s kernels=%s history=%s" % (variant, sides, seen, ["%.3e" % v for v in h[-3:]]), flush=True)


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 180
    if len(sys.argv) > 2:
        return child(side, sys.argv[2])
    rc = 0
    for v, env in VARIANTS.items():
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, str(side), v], env=e)
        if r.returncode != 0:
            print("sanitize %s FAILED rc=%d" % (v, r.returncode), flush=True); rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
