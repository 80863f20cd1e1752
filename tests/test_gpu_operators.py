"""Per-operator parity of the CUDA path against the CPU oracle on identical matrices.

Tolerances (BASELINE.md §5 / north star): SpMV, restriction, prolongation 1e-12 relative;
integer artefacts bit-exact; scatter ops bit-exact.
"""
import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_OP = 1e-12

CASES = {
    "dirichlet_p4": dict(sizes=[13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4),
    "dirichlet_p6": dict(sizes=[13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=6),
    "neumann_p3": dict(sizes=[13, 25, 50], kind=oracle.KIND_NEUMANN, fine_poly=3),
    "mixed_p4": dict(sizes=[13, 25, 50], kind=oracle.KIND_MIXED, fine_poly=4),
}


ARITH = {"reference_order": capi.ARITH_REFERENCE_ORDER, "fast": capi.ARITH_FAST}


@pytest.fixture(scope="module", params=[(c, a) for c in sorted(CASES) for a in sorted(ARITH)], ids=lambda p: "%s-%s" % p)
def pair(request, libmmg):
    cfg = CASES[request.param[0]]
    mg = oracle.make_hierarchy(cfg["sizes"], kind=cfg["kind"], fine_poly=cfg["fine_poly"])
    s = H.gpu_solver_from_oracle(mg)
    s.set_arithmetic(ARITH[request.param[1]])
    s.exact = request.param[1] == "reference_order"
    return mg, s


def close(s, got, want, tol):
    """reference-order arithmetic must reproduce the oracle bit for bit; the fast mode to `tol` relative"""
    if s.exact:
        assert np.array_equal(got, want), "reference-order mode is not bit-identical (max rel %g)" % H.rel_err(got, want)
    else:
        assert H.rel_err(got, want) < tol


def test_native_library_is_the_one_loaded(libmmg):
    assert b"sm_100a" in libmmg.mmg_build_info()
    assert capi.device_count() >= 1


def test_csr_roundtrip_is_bit_exact(pair):
    mg, s = pair
    for l in range(mg.nlevels):
        _, ptr, idx, val = mg.level(l).csr()
        _, p2, i2, v2 = s.grid(l).csr()
        assert np.array_equal(ptr, p2) and np.array_equal(idx, i2) and np.array_equal(val, v2)
    for l in range(1, mg.nlevels):
        a, b = mg.level(l).csr(oracle.MAT_R), s.interp_csr(capi.MAT_RESTRICT, l)
        assert a[0] == b[0] and all(np.array_equal(u, v) for u, v in zip(a[1:], b[1:]))


def test_residual(pair):
    mg, s = pair
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        v = H.random_values(lv, 7 + l)
        lv.set_vec(oracle.VEC_VALUES, v)
        g.values_ = v
        close(s, g.residual(), lv.residual(), TOL_OP)


def test_residual_norm(pair):
    mg, s = pair
    lv, g = mg.level(-1), s.grid(-1)
    v = H.random_values(lv, 3)
    lv.set_vec(oracle.VEC_VALUES, v)
    g.values_ = v
    assert abs(s.residual() - mg.residual()) <= 1e-12 * abs(mg.residual())


def test_scatter_ops_bit_exact(pair):
    mg, s = pair
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        v = H.random_values(lv, 11 + l)
        for coarse in (False, True):
            lv.set_vec(oracle.VEC_VALUES, v); g.values_ = v
            lv.boundary_op(coarse); g.boundaryOp("coarse" if coarse else "fine")
            assert np.array_equal(g.values_, lv.values)
        assert np.array_equal(g.fix_vector_bound_coarse(v), lv.fix_vector_bound_coarse(v))
        if lv.neumann:
            src = lv.source
            for coarse in (True, False):
                lv.modify_coeff_neumann(coarse); g.modify_coeff_neumann("coarse" if coarse else "fine")
                assert np.array_equal(g.source_, lv.source)
            lv.set_vec(oracle.VEC_SOURCE, src); g.source_ = src


def test_bound_eval_neumann(pair):
    mg, s = pair
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        if not lv.neumann:
            continue
        v = H.random_values(lv, 5 + l)
        lv.set_vec(oracle.VEC_VALUES, v); g.values_ = v
        lv.bound_eval_neumann(); g.bound_eval_neumann()
        close(s, g.values_, lv.values, TOL_OP)


def test_push_inhomog_to_rhs(pair):
    mg, s = pair
    lv, g = mg.level(-1), s.grid(-1)
    if not lv.implicit:
        pytest.skip("explicit grid")
    src = lv.source
    rng = np.random.default_rng(5)
    t = src + rng.standard_normal(src.size)
    lv.set_vec(oracle.VEC_SOURCE, t); g.source_ = t
    lv.push_inhomog_to_rhs(); g.push_inhomog_to_rhs()
    assert H.rel_err(g.source_, lv.source) < 1e-14
    lv.set_vec(oracle.VEC_SOURCE, src); g.source_ = src


@pytest.mark.parametrize("smoother", ["lexicographic", "multicolour"])
def test_sor(pair, smoother):
    mg, s = pair
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        v = 1e-3 * H.random_values(lv, 21 + l)
        lv.set_vec(oracle.VEC_VALUES, v); g.values_ = v
        if smoother == "lexicographic":
            lv.sor(); g.sor(capi.LEXICOGRAPHIC)
        else:
            lv.sor_multicolour(); g.sor(capi.MULTICOLOUR)
        close(s, g.values_, lv.values, 1e-11)


def test_schedules_bit_exact(pair):
    mg, s = pair
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        nc, col = lv.colouring()
        nc2, col2 = g.colouring()
        assert nc == nc2 and np.array_equal(col, col2)
        lev = lv.lex_levels()
        nl2, lev2 = g.lex_levels()
        assert np.array_equal(lev, lev2) and nl2 == lev.max() + 1


def test_restrict_and_prolong(pair):
    mg, s = pair
    for l in range(1, mg.nlevels):
        fine, coarse = mg.level(l), mg.level(l - 1)
        v = H.random_values(fine, 31 + l)
        fine.set_vec(oracle.VEC_VALUES, v); s.grid(l).values_ = v
        # oracle: statement multigrid.cpp:81-86
        r = fine.residual()
        src = coarse.source
        src[: coarse.n] = fine.spmv(oracle.MAT_R, r[: fine.n])
        src = coarse.fix_vector_bound_coarse(src)
        if fine.neumann:
            src[-1] = 0
            coarse.set_vec(oracle.VEC_SOURCE, src)
            coarse.modify_coeff_neumann(True)
            src = coarse.source
        s.restrict(l)
        close(s, s.grid(l - 1).source_, src, TOL_OP)
        # prolongation + correction, multigrid.cpp:102-106
        vc = H.random_values(coarse, 41 + l)
        coarse.set_vec(oracle.VEC_VALUES, vc); s.grid(l - 1).values_ = vc
        corr = coarse.spmv(oracle.MAT_P, vc[: coarse.n])
        full = np.zeros(fine.A); full[: fine.n] = corr
        if not fine.neumann:
            full = fine.fix_vector_bound_coarse(full)
        expect = v + full
        s.prolong_correct(l)
        close(s, s.grid(l).values_, expect, TOL_OP)
