"""Problem factories: the callers of the hot path, mirrored from the reference drivers.

``make_grid`` follows genGmshGridDirichlet (testing_functions.cpp:68-159, square branch),
genGmshGridNeumann (:161-284) and a mixed-BC variant built from the same per-boundary calls, but
takes the point cloud as arrays instead of a Gmsh file.  Everything numerical runs in libmmg on the
device; this file only evaluates the manufactured right-hand side and detects the boundary exactly
the way the reference does (x==0 || x==1 || y==0 || y==1).
"""
import numpy as np

from . import capi
from .clouds import jittered_square, make_cloud

PI = 3.141592653589793238462643383279        # testing_functions.hpp:9


def stencil_size(poly_deg):
    return int(2.5 * (poly_deg + 1) * (poly_deg + 2) / 2)     # testing_functions.cpp:378


def grid_props(poly_deg, iters=5, omega=1.4, rbf_exp=3):
    """gen_mg_param, testing_functions.cpp:372-380"""
    return dict(rbfExp=rbf_exp, polyDeg=poly_deg, stencilSize=stencil_size(poly_deg), iters=iters, omega=omega)


def make_grid(kind, x, y, poly_deg, k1=1, k2=1, fine=True, device=0, **props):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    p = grid_props(poly_deg, **props)
    on_b = (x == 0) | (x == 1) | (y == 0) | (y == 1)
    coarse = "fine" if fine else "coarse"
    if kind == "dirichlet":
        source = -(k1 * k1 + k2 * k2) * PI * PI * np.sin(k1 * PI * x) * np.sin(k2 * PI * y)
        pts = np.nonzero(on_b)[0].astype(np.int32)
        vals = np.zeros(pts.size)
        g = capi.Grid(x, y, [capi.Boundary(pts, vals, type=capi.BC_DIRICHLET)], p, source, device=device)
        g.set_implicitFlag(False)
        g.setBCFlag(0, "dirichlet", vals)
        g.rcm_order_points()
        g.build_laplacian()
        return g
    if kind == "neumann":
        source = np.zeros(x.size + 1)
        source[:-1] = -(k1 * k1 + k2 * k2) * PI * PI * np.cos(k1 * PI * x) * np.cos(k2 * PI * y)
        pts = np.nonzero(on_b)[0].astype(np.int32)
        vals = np.zeros(pts.size)
        bnds = [capi.Boundary(pts, vals, type=capi.BC_NEUMANN)]
        flags = [(0, "neumann", vals)]
    elif kind == "mixed":
        # u = sin(k1 pi x) cos(k2 pi y): Dirichlet on x in {0,1}, homogeneous Neumann on y in {0,1};
        # the Neumann boundary goes first because build_normal_vecs only visits boundaries_[0] (grid.cpp:445)
        source = np.zeros(x.size + 1)
        source[:-1] = -(k1 * k1 + k2 * k2) * PI * PI * np.sin(k1 * PI * x) * np.cos(k2 * PI * y)
        d = (x == 0) | (x == 1)
        n = on_b & ~d
        dp, npts = np.nonzero(d)[0].astype(np.int32), np.nonzero(n)[0].astype(np.int32)
        bnds = [capi.Boundary(npts, np.zeros(npts.size), type=capi.BC_NEUMANN), capi.Boundary(dp, np.zeros(dp.size), type=capi.BC_DIRICHLET)]
        flags = [(0, "neumann", np.zeros(npts.size)), (1, "dirichlet", np.zeros(dp.size))]
    else:
        raise ValueError(kind)
    g = capi.Grid(x, y, bnds, p, source, device=device)
    g.set_implicitFlag(True)
    for b, t, v in flags:
        g.setBCFlag(b, t, v)
    g.build_normal_vecs("square")
    g.rcm_order_points()
    g.build_deriv_normal_bound()
    g.build_laplacian()
    g.modify_coeff_neumann(coarse)
    g.push_inhomog_to_rhs()
    return g


def make_hierarchy(sizes, kind="dirichlet", fine_poly=4, coarse_poly=3, seed0=1000, jitter=0.3, device=0, cloud="jittered", **kw):
    """run_mg_sim's set-up (testing_functions.cpp:328-339) on synthetic clouds: one independent jittered lattice
    per level, coarse levels polyDeg 3, finest ``fine_poly``; then buildMatrices()."""
    mg = capi.Multigrid()
    for l, s in enumerate(sizes):
        x, y = make_cloud(cloud, s, seed0 + l, jitter)
        last = l == len(sizes) - 1
        mg.addGrid(make_grid(kind, x, y, fine_poly if last else coarse_poly, fine=last, device=device, **kw))
    mg.buildMatrices()
    return mg


def make_ppe_grid(x, y, poly_deg, dt, mu, rho, fine=True, device=0, **props):
    """genFractionalStepGrid (FractionalStepSim.cpp:3-49): pressure-Poisson level of the Kovasznay fractional-step driver —
    all-Neumann square, zero source, boundary values 0.5*exp(2*lambda*x), plus the three explicit operators."""
    import math

    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    p = grid_props(poly_deg, **props)
    re = rho / mu
    lam = 0.5 * re - math.sqrt(0.25 * re * re + 4 * math.pi * math.pi)
    pts = np.nonzero((x == 0) | (x == 1) | (y == 0) | (y == 1))[0].astype(np.int32)
    vals = np.array([0.5 * math.exp(2 * lam * x[i]) for i in pts])
    g = capi.Grid(x, y, [capi.Boundary(pts, vals, type=capi.BC_NEUMANN)], p, np.zeros(x.size + 1), device=device)
    g.fs_init(dt, mu, rho)
    g.set_implicitFlag(True)
    g.setBCFlag(0, "neumann", vals)
    g.build_normal_vecs("square")
    g.rcm_order_points()
    g.build_deriv_normal_bound()
    g.build_laplacian()
    g.modify_coeff_neumann("fine" if fine else "coarse")
    g.fs_build_operators()
    g.push_inhomog_to_rhs()
    return g


def fracstep_time_step(mg, tol, max_cycles=200):
    """One pass of the loop body of run_fracstep_param (FractionalStepSim.cpp:130-147).  Returns (V-cycles used, fs_residual)."""
    fine = mg.grid(-1)
    fine.fs_set_vec(capi.FS_U_OLD, fine.fs_vec(capi.FS_U))
    fine.fs_set_vec(capi.FS_V_OLD, fine.fs_vec(capi.FS_V))
    fine.set_uv_bound()
    fine.calc_hat()
    fine.set_ppe_source()
    fine.push_inhomog_to_rhs()
    n, _ = mg.solve(tol, max_cycles, extra_bound_eval=True)     # while (mg.residual() >= tol) { mg.vCycle(); finestGrid->bound_eval_neumann(); }
    fine.correct_uv()
    fine.set_uv_bound()
    return n, fine.fs_residual()
