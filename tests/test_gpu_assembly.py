"""Device assembly parity: kNN lists and the reordering are integer work (bit-exact); RBF-FD weights are
checked three ways: bitwise against the oracle's full-pivot LU where the inputs round
identically, entrywise at a conditioning-aware tolerance everywhere, and through the analytic
known-answer properties (polynomial reproduction, partition of unity)."""
import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.clouds import jittered_square
from meshlessmultigridpoisson_b200.problems import make_grid
from tests import helpers as H

pytestmark = pytest.mark.gpu

KINDS = {"dirichlet": oracle.KIND_DIRICHLET, "neumann": oracle.KIND_NEUMANN, "mixed": oracle.KIND_MIXED}


def build_both(kind, s, poly, seed=1000, fine=True):
    x, y = jittered_square(s, seed=seed)
    mg = oracle.Multigrid()
    mg.add_level(KINDS[kind], x, y, poly, fine=fine)
    g = make_grid(kind, x, y, poly, fine=fine)
    return mg.level(0), g


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("poly", [3, 4, 6])
def test_knn_bit_exact(libmmg, kind, poly):
    lv, g = build_both(kind, 40, poly)
    x, y = lv.points()
    flags = lv.bcflags()
    k = oracle.stencil_size(poly)
    ids = np.arange(lv.n)
    got = g.kNearestNeighbors(x, y, k, neumann=lv.neumann, q_bcflag=(flags != 0).astype(np.int32))
    for i in ids[:: max(1, lv.n // 400)].tolist() + np.nonzero(flags)[0][:200].tolist():
        want = lv.knn(x[i], y[i], k, neumann=lv.neumann, point_bc=bool(flags[i]), cells=False)
        assert np.array_equal(got[i], want), i
    # arbitrary query points (interpolation targets), no exclusion
    rng = np.random.default_rng(3)
    qx, qy = rng.uniform(0, 1, 200), rng.uniform(0, 1, 200)
    got = g.kNearestNeighbors(qx, qy, k)
    for i in range(200):
        assert np.array_equal(got[i], lv.knn(qx[i], qy[i], k, cells=False))


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_rcm_permutation_bit_exact(libmmg, kind):
    lv, g = build_both(kind, 40, 4)
    assert np.array_equal(g.perm(), lv.perm())
    gx, gy = g.points_
    ox, oy = lv.points()
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)
    assert np.array_equal(g.bcFlags_, lv.bcflags())
    assert np.array_equal(g.source_, lv.source)
    for b, (t, pts, vals) in enumerate(lv.boundaries()):
        t2, p2, v2 = g.boundary(b)
        assert t == t2 and np.array_equal(pts, p2) and np.array_equal(vals, v2)


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("poly", [3, 4, 6])
def test_laplacian_matches_oracle(libmmg, kind, poly):
    lv, g = build_both(kind, 32, poly)
    (_, _), ptr, idx, val = lv.csr()
    (_, _), p2, i2, v2 = g.csr()
    assert np.array_equal(ptr, p2) and np.array_equal(idx, i2)          # structure: integer work, bit-exact
    rows = np.repeat(np.arange(ptr.size - 1), np.diff(ptr))
    scale = np.maximum.reduceat(np.abs(val), ptr[:-1])[rows]
    rel = np.abs(val - v2) / scale
    same = np.mean(val == v2)
    # conditioning-aware bound: the local saddle systems have cond up to ~1e10 at polyDeg 6
    assert rel.max() < {3: 1e-9, 4: 1e-8, 6: 1e-5}[poly], rel.max()
    # Bitwise agreement needs every pow() feeding the (n+m)^2 local system to round identically; glibc's pow is
    # not correctly rounded (~2e-4 of calls differ from the device's <0.5000001-ulp evaluation), so at polyDeg 6
    # (~8000 pow calls per stencil) only a fraction of stencils can match bit for bit.
    assert same > {3: 0.5, 4: 0.3, 6: 0.05}[poly], same
    assert np.allclose(g.diags[: lv.n][lv.bcflags() != 2], lv.diags[: lv.n][lv.bcflags() != 2], rtol=1e-5)


@pytest.mark.parametrize("poly", [3, 4, 6])
def test_weights_reproduce_polynomials(libmmg, poly):
    """Known-answer test the reference implies (grid.cpp:282-297,404-417): Laplacian weights are exact on
    every monomial of degree <= polyDeg; interpolation rows are a partition of unity."""
    _, g = build_both("dirichlet", 32, poly)
    n = oracle.stencil_size(poly)
    x, y = g.points_
    ids = np.arange(0, x.size, 7)
    w, nb = g.laplaceWeights(ids, n)
    h = 1.0 / 31
    for a in range(poly + 1):
        for q in range(a + 1):
            px_, py_ = a - q, q
            f = (x[nb] - x[ids, None]) ** px_ * (y[nb] - y[ids, None]) ** py_      # centred monomial
            lap = (w * f).sum(1)
            exact = 2.0 if (px_, py_) in ((2, 0), (0, 2)) else 0.0
            assert np.abs(lap - exact).max() < 1e-6 / h ** 0, (a, q, np.abs(lap - exact).max())
    rng = np.random.default_rng(0)
    qx, qy = rng.uniform(0.05, 0.95, 300), rng.uniform(0.05, 0.95, 300)
    wi, nbi = g.pointInterpWeights(qx, qy, poly)
    assert np.abs(wi.sum(1) - 1).max() < 1e-9
    fx = np.sin(2 * x) * np.cos(3 * y)
    assert np.abs((wi * fx[nbi]).sum(1) - np.sin(2 * qx) * np.cos(3 * qy)).max() < 1e-3


def test_interp_weights_match_oracle(libmmg):
    lv, g = build_both("dirichlet", 32, 4)
    rng = np.random.default_rng(1)
    qx, qy = rng.uniform(0, 1, 64), rng.uniform(0, 1, 64)
    for poly in (3, 4, 6):
        w, nb = g.pointInterpWeights(qx, qy, poly)
        bitwise = 0
        for i in range(64):
            wo, nbo = lv.interp_weights(qx[i], qy[i], poly)
            assert np.array_equal(nb[i], nbo)
            assert np.abs(w[i] - wo[: nbo.size]).max() < 1e-6 * np.abs(wo).max()
            bitwise += np.array_equal(w[i], wo[: nbo.size])
        assert bitwise >= {3: 32, 4: 20, 6: 3}[poly], bitwise


@pytest.mark.parametrize("kind,fine_poly,sizes", [("dirichlet", 4, [13, 25, 50]), ("dirichlet", 6, [13, 25, 50]), ("mixed", 4, [13, 25, 50]),
                                                  ("neumann", 3, [13, 25, 50])])
def test_device_built_hierarchy_solves_like_the_oracle(libmmg, kind, fine_poly, sizes):
    """Full device pipeline (reorder, assemble, buildMatrices, vCycle).  The operators differ from the oracle's at
    cond*eps level, so the GPU-built operators are downloaded INTO the oracle and the two V-cycles compared on
    identical matrices (SURVEY.md §7 contract)."""
    from meshlessmultigridpoisson_b200.problems import make_hierarchy

    mg = oracle.make_hierarchy(sizes, kind=KINDS[kind], fine_poly=fine_poly)
    s = make_hierarchy(sizes, kind, fine_poly)
    L = mg.nlevels
    for l in range(L):
        lv, g = mg.level(l), s.grid(l)
        assert np.array_equal(g.perm(), lv.perm())
        shape, ptr, idx, val = g.csr()
        o = lv.csr()
        assert np.array_equal(o[1], ptr) and np.array_equal(o[2], idx)
        assert H.rel_err(val, o[3]) < 1e-5
        lv.set_csr(oracle.MAT_A, shape, ptr, idx, val)
        assert H.rel_err(g.source_, lv.source) < 1e-9
        lv.set_vec(oracle.VEC_SOURCE, g.source_)
    for l in range(1, L):
        shape, ptr, idx, val = s.interp_csr(capi.MAT_RESTRICT, l)
        o = mg.level(l).csr(oracle.MAT_R)
        assert o[0] == shape and np.array_equal(o[1], ptr) and np.array_equal(o[2], idx) and H.rel_err(val, o[3]) < 1e-5
        mg.level(l).set_csr(oracle.MAT_R, shape, ptr, idx, val)
    for l in range(L - 1):
        shape, ptr, idx, val = s.interp_csr(capi.MAT_PROLONG, l)
        o = mg.level(l).csr(oracle.MAT_P)
        assert o[0] == shape and np.array_equal(o[1], ptr) and np.array_equal(o[2], idx) and H.rel_err(val, o[3]) < 1e-5
        mg.level(l).set_csr(oracle.MAT_P, shape, ptr, idx, val)
    mg.vcycle(15)
    s.vCycle(15)
    ho, hg = mg.history(), s.residuals_
    live = ho > 1e-11 * ho[0]
    assert (np.abs(hg[live] - ho[live]) / ho[live]).max() < 1e-10
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < 1e-8
