"""Shared test plumbing: build oracle hierarchies and mirror them into the CUDA library
through the C-ABI upload path (identical matrices on both sides)."""
import numpy as np

import oracle
from meshlessmultigridpoisson_b200 import capi

REF_ITERS, REF_OMEGA, REF_RBFEXP = 5, 1.4, 3   # gen_mg_param, testing_functions.cpp:374-377


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


def rel_l2(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def gpu_grid_from_oracle(lv, device=0):
    """One oracle level -> mmg Grid with the same points (already reordered), flags, boundaries,
    source, values and operator (upload path)."""
    x, y = lv.points()
    bnds = lv.boundaries()
    boundaries = [capi.Boundary(p, v, type=t) for (t, p, v) in bnds]
    g = capi.Grid(x, y, boundaries, lv.props, lv.source, device=device)
    g.set_implicitFlag(lv.implicit)
    for b, (t, p, v) in enumerate(bnds):
        g.setBCFlag(b, "dirichlet" if t == 1 else "neumann", v)
    (_, _), ptr, idx, val = lv.csr(oracle.MAT_A)
    (_, _), nptr, nidx, nval = lv.csr(oracle.MAT_NBC)
    g.set_laplacian_csr(ptr, idx, val, diags=lv.diags, nbc=(nptr, nidx, nval))
    g.values_ = lv.values
    return g


def gpu_solver_from_oracle(mg, device=0, fracstep=False):
    s = capi.FractionalStepMultigrid() if fracstep else capi.Multigrid()
    for l in range(mg.nlevels):
        s.addGrid(gpu_grid_from_oracle(mg.level(l), device))
    L = mg.nlevels
    for l in range(1, L):
        shape, ptr, idx, val = mg.level(l).csr(oracle.MAT_R)
        s.set_interp_csr(capi.MAT_RESTRICT, l, shape, ptr, idx, val)
    for l in range(L - 1):
        shape, ptr, idx, val = mg.level(l).csr(oracle.MAT_P)
        s.set_interp_csr(capi.MAT_PROLONG, l, shape, ptr, idx, val)
    return s


def random_values(lv, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(lv.A)


def oracle_mirror_of_gpu(mg_gpu, kind, fine_poly, coarse_poly=3, fracstep=False, **props):
    """Mirror a DEVICE-built hierarchy into the oracle without running the oracle's own set-up (kNN + one LU per node take
    minutes beyond ~250k nodes): level state and operators are downloaded through the C-ABI and pushed into raw oracle
    levels, so both sides hold identical matrices (SURVEY.md section 7 contract).  kind: 'dirichlet' | 'neumann' | 'mixed'."""
    from meshlessmultigridpoisson_b200.problems import grid_props

    L = mg_gpu.num_grids
    mg = oracle.Multigrid(fracstep=fracstep)
    nb = 2 if kind == "mixed" else 1
    for l in range(L):
        g = mg_gpu.grid(l)
        x, y = g.points_
        bnds = [g.boundary(b) for b in range(nb)]
        p = grid_props(fine_poly if l == L - 1 else coarse_poly, **props)
        mg.add_level_raw(x, y, p, bnds, g.source_, implicit=(kind != "dirichlet"))
    for l in range(L):
        g, lv = mg_gpu.grid(l), mg.level(l)
        shape, ptr, idx, val = g.csr()
        lv.set_csr(oracle.MAT_A, shape, ptr, idx, val)
        if kind != "dirichlet":
            shape, ptr, idx, val = g.csr(capi.MAT_NEUMANN_COEFFS)
            lv.set_csr(oracle.MAT_NBC, shape, ptr, idx, val)
        lv.set_vec(oracle.VEC_DIAGS, g.diags)
        lv.set_vec(oracle.VEC_VALUES, g.values_)
    mg.alloc_interp()
    for l in range(1, L):
        mg.level(l).set_csr(oracle.MAT_R, *mg_gpu.interp_csr(capi.MAT_RESTRICT, l))
    for l in range(L - 1):
        mg.level(l).set_csr(oracle.MAT_P, *mg_gpu.interp_csr(capi.MAT_PROLONG, l))
    return mg
