"""Fractional-step path (BASELINE config 4's call pattern): FractionalStepMultigrid V-cycle as the
pressure-Poisson solver plus the explicit operators of FractionalStepGrid, against the oracle (itself pinned bit-exact
against fractionalStepGrid.cpp / FractionalStepSim.cpp compiled on the Eigen shim)."""
import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.clouds import jittered_square
from meshlessmultigridpoisson_b200.problems import make_ppe_grid
from tests import helpers as H

pytestmark = pytest.mark.gpu
DT, MU, RHO = 2e-4, 0.025, 1.0          # run_frac_step_test, FractionalStepSim.cpp:202
SIZES = [13, 25, 50]


def mirrored():
    """oracle PPE hierarchy + the same operators uploaded into the CUDA library"""
    mg = oracle.make_hierarchy(SIZES, kind=oracle.KIND_PPE, fine_poly=3, fracstep=True)
    s = H.gpu_solver_from_oracle(mg, fracstep=True)
    lv, g = mg.level(-1), s.grid(-1)
    g.fs_init(DT, MU, RHO)
    nx, ny = lv.normals()
    g.L.mmg_grid_set_normal_vecs(g.h, nx, ny)
    g.fs_init(DT, MU, RHO)                                  # refresh the device copy of the normals
    for which_o, which_g in ((oracle.MAT_DX, capi.MAT_DERIVX), (oracle.MAT_DY, capi.MAT_DERIVY), (oracle.MAT_UVLAP, capi.MAT_UVLAPLACE)):
        _, ptr, idx, val = lv.csr(which_o)
        g.fs_set_operator_csr(which_g, ptr, idx, val)
    return mg, s


def test_time_step_matches_the_oracle_bit_for_bit(libmmg):
    mg, s = mirrored()
    lv, g = mg.level(-1), s.grid(-1)
    L = mg.nlevels - 1
    for step in range(2):
        # FractionalStepSim.cpp:131-137
        mg.L.orc_fs_step_pre(mg.h, L)
        g.fs_set_vec(capi.FS_U_OLD, g.fs_vec(capi.FS_U)); g.fs_set_vec(capi.FS_V_OLD, g.fs_vec(capi.FS_V))
        g.set_uv_bound(); g.calc_hat(); g.set_ppe_source(); g.push_inhomog_to_rhs()
        for which_o, which_g in ((oracle.VEC_UHAT, capi.FS_U_HAT), (oracle.VEC_VHAT, capi.FS_V_HAT)):
            assert np.array_equal(g.fs_vec(which_g), lv.vec(which_o))
        assert np.array_equal(g.source_, lv.source)
        # :139-142 with a fixed number of cycles so both sides do identical work
        for _ in range(6):
            mg.vcycle(1); lv.bound_eval_neumann()
            s.vCycle(1); g.bound_eval_neumann()
        assert np.array_equal(g.values_, lv.values)
        # :144-147
        r_o = mg.L.orc_fs_step_post(mg.h, L)
        g.correct_uv(); g.set_uv_bound()
        assert np.array_equal(g.fs_vec(capi.FS_U), lv.vec(oracle.VEC_U)) and np.array_equal(g.fs_vec(capi.FS_V), lv.vec(oracle.VEC_V))
        assert abs(g.fs_residual() - r_o) <= 1e-13 * abs(r_o)
    assert np.allclose(s.residuals_, mg.history(), rtol=1e-10, atol=0)


def test_device_built_operators_match_the_oracle(libmmg):
    x, y = jittered_square(32, seed=1000)
    mg = oracle.Multigrid(fracstep=True)
    mg.add_level(oracle.KIND_PPE, x, y, 3, fine=True, dt=DT, mu=MU, rho=RHO)
    lv = mg.level(0)
    g = make_ppe_grid(x, y, 3, DT, MU, RHO)
    assert np.array_equal(g.perm(), lv.perm())
    for which_o, which_g in ((oracle.MAT_A, capi.MAT_LAPLACE), (oracle.MAT_DX, capi.MAT_DERIVX), (oracle.MAT_DY, capi.MAT_DERIVY), (oracle.MAT_UVLAP, capi.MAT_UVLAPLACE)):
        so, po, io, vo = lv.csr(which_o)
        sg, pg, ig, vg = g.csr(which_g)
        assert so == sg and np.array_equal(po, pg) and np.array_equal(io, ig)
        assert H.rel_err(vg, vo) < 1e-8
    # derivative operators differentiate a smooth field
    gx, gy = g.points_
    f = np.sin(2 * gx) * np.cos(3 * gy)
    g.fs_set_vec(capi.FS_U, f); g.fs_set_vec(capi.FS_V, 0 * f)
    sh, ptr, idx, val = g.csr(capi.MAT_DERIVX)
    import scipy.sparse as sp
    dfdx = sp.csr_matrix((val, idx, ptr), shape=sh) @ f
    interior = g.bcFlags_ == 0
    assert np.abs(dfdx - 2 * np.cos(2 * gx) * np.cos(3 * gy))[interior].max() < 5e-3


def test_time_step_driver_runs_on_a_device_built_hierarchy(libmmg):
    from meshlessmultigridpoisson_b200.problems import fracstep_time_step

    mg = capi.FractionalStepMultigrid()
    for l, sd in enumerate(SIZES):
        x, y = jittered_square(sd, seed=1000 + l)
        mg.addGrid(make_ppe_grid(x, y, 3, DT, MU, RHO, fine=(l == len(SIZES) - 1)))
    mg.buildMatrices()
    n1, r1 = fracstep_time_step(mg, 1e-6)
    n2, r2 = fracstep_time_step(mg, 1e-6)
    assert 0 < n1 < 200 and 0 <= n2 < 200 and np.isfinite(r1) and np.isfinite(r2)
    u = mg.grid(-1).fs_vec(capi.FS_U)
    assert np.isfinite(u).all() and np.abs(u).max() < 10
