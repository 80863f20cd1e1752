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


def _on_circle(x, y, r2):
    return np.abs(r2 - (x - 0.5) * (x - 0.5) - (y - 0.5) * (y - 0.5)) <= 1e-10        # testing_functions.cpp:101,122,127


def _concentric_source(x, y, k1):
    """testing_functions.cpp:109-121: Laplacian of sin(pi k1 r*), r* = (r - 0.25) / 0.25, as the reference writes it"""
    x = x - 0.5
    y = y - 0.5
    r2 = x * x + y * y
    rstar = (np.sqrt(r2) - 0.25) / (0.5 - 0.25)
    out = 0.0
    for c in (x, y):
        out = out + (-PI * k1 * k1 * PI * np.sin(PI * k1 * rstar) * (4 * c * r2 ** -0.5) ** 2
                     + PI * k1 * np.cos(PI * k1 * rstar) * 4 * (r2 ** -0.5 + 2 * c * c * -0.5 * r2 ** -1.5))
    return out


def make_grid(kind, x, y, poly_deg, k1=1, k2=1, fine=True, device=0, geomtype="square", **props):
    """genGmshGridDirichlet (testing_functions.cpp:68-159) / genGmshGridNeumann (:161-284) for the three geomtypes, plus the mixed
    variant on the square.  Boundary detection and right-hand sides are evaluated on the host exactly as the drivers do."""
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    p = grid_props(poly_deg, **props)
    on_b = (x == 0) | (x == 1) | (y == 0) | (y == 1)
    coarse = "fine" if fine else "coarse"
    if geomtype != "square":
        assert kind in ("dirichlet", "neumann")
        outer = on_b if geomtype == "square_with_circle" else _on_circle(x, y, 0.25)
        inner = _on_circle(x, y, 0.0625) & ~outer
        po, pi_ = np.nonzero(outer)[0].astype(np.int32), np.nonzero(inner)[0].astype(np.int32)
        nxo, nyo = x[po] - 0.5, y[po] - 0.5
        no = np.sqrt(nxo * nxo + nyo * nyo); nxo, nyo = nxo / no, nyo / no
        nxi, nyi = x[pi_] - 0.5, y[pi_] - 0.5
        ni = np.sqrt(nxi * nxi + nyi * nyi); nxi, nyi = nxi / ni, nyi / ni
        if kind == "dirichlet":
            if geomtype == "square_with_circle":
                source = -(k1 * k1 + k2 * k2) * PI * PI * np.sin(k1 * PI * x) * np.sin(k1 * PI * y)
                vo, vi = np.zeros(po.size), np.sin(k1 * PI * x[pi_]) * np.sin(k1 * PI * y[pi_])
            else:
                source = _concentric_source(x, y, k1)
                vo, vi = np.zeros(po.size), np.zeros(pi_.size)
            g = capi.Grid(x, y, [capi.Boundary(po, vo, type=capi.BC_DIRICHLET), capi.Boundary(pi_, vi, type=capi.BC_DIRICHLET)], p, source, device=device)
            g.set_implicitFlag(False)
            g.setBCFlag(0, "dirichlet", vo)
            g.setBCFlag(1, "dirichlet", vi)
            g.rcm_order_points()
            g.build_laplacian()
            return g
        source = np.zeros(x.size + 1)
        if geomtype == "square_with_circle":
            source[:-1] = -(k1 * k1 + k2 * k2) * PI * PI * np.cos(k1 * PI * x) * np.cos(k2 * PI * y)
            vo = np.zeros(po.size)
            vi = -nxi * PI * k1 * np.sin(k1 * PI * x[pi_]) * np.cos(k2 * PI * y[pi_]) - nyi * PI * k2 * np.cos(k1 * PI * x[pi_]) * np.sin(k2 * PI * y[pi_])
        else:
            source[:-1] = _concentric_source(x, y, k1)
            def dn(px, py, nx_, ny_, sign):
                r = np.sqrt((px - 0.5) ** 2 + (py - 0.5) ** 2)
                rstar = (r - 0.25) / (0.5 - 0.25)
                return sign * (nx_ * k1 * PI * np.cos(k1 * PI * rstar) / r * 4 * (px - 0.5)) + sign * (ny_ * k1 * PI * np.cos(k1 * PI * rstar) / r * 4 * (py - 0.5))
            vo, vi = dn(x[po], y[po], nxo, nyo, -1.0), dn(x[pi_], y[pi_], nxi, nyi, 1.0)
        g = capi.Grid(x, y, [capi.Boundary(po, vo, type=capi.BC_NEUMANN), capi.Boundary(pi_, vi, type=capi.BC_NEUMANN)], p, source, device=device)
        g.set_implicitFlag(True)
        g.setBCFlag(0, "neumann", vo)
        g.setBCFlag(1, "neumann", vi)
        g.build_normal_vecs(geomtype)
        g.rcm_order_points()
        g.build_deriv_normal_bound()
        g.build_laplacian()
        g.modify_coeff_neumann(coarse)
        g.push_inhomog_to_rhs()
        return g
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


def make_hierarchy(sizes, kind="dirichlet", fine_poly=4, coarse_poly=3, seed0=1000, jitter=0.3, device=0, cloud="jittered", geomtype="square", **kw):
    """run_mg_sim's set-up (testing_functions.cpp:328-339) on synthetic clouds: one independent jittered lattice
    per level, coarse levels polyDeg 3, finest ``fine_poly``; then buildMatrices()."""
    mg = capi.Multigrid()
    for l, s in enumerate(sizes):
        x, y = make_cloud(cloud if geomtype == "square" else geomtype, s, seed0 + l, jitter)
        last = l == len(sizes) - 1
        mg.addGrid(make_grid(kind, x, y, fine_poly if last else coarse_poly, fine=last, device=device, geomtype=geomtype, **kw))
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
