"""Parity of the kernel instantiations the benchmark actually times, at the sizes that select them.

The small hierarchies of test_gpu_operators.py / test_gpu_vcycle.py never reach the big-level dispatch: 8 lanes per row is
chosen from 200 000 rows up, the TMA-fed / packed multicolour sweep above MMG_MC_FLOW_MAX_ROWS, the chunked lexicographic
sweep's DAG is only deep on large levels.  Here a 250k-node (and a 1M-node, BASELINE config 2) hierarchy is built on the
DEVICE, mirrored into the oracle through the C-ABI (identical matrices, tests/helpers.oracle_mirror_of_gpu) and every
operator of the V-cycle is compared with the oracle's restatement of grid.cpp:104-151 / multigrid.cpp:62-110:
reference-order arithmetic bit for bit, fast arithmetic to 1e-12 relative; lexicographic history to 1e-10 per cycle.
Each test asserts the kernel instantiation that ran (mmg_debug_last_kernel), so a silent fallback fails.
"""
import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.problems import make_hierarchy
from tests import helpers as H

pytestmark = pytest.mark.gpu

SIDES_250K = [32, 63, 125, 250, 500]
SIDES_1M = [32, 63, 125, 250, 500, 1000]          # BASELINE config 2: 1M nodes, 6 levels
MC_KNOBS = ("MMG_MC_FLOW", "MMG_MC_FLOW_MAX_ROWS", "MMG_MC_TMA", "MMG_MC_PACKED", "MMG_SPMV_TMA", "MMG_MC_TMAFLOW", "MMG_MC_TMAFLOW_MIN_ROWS")


@pytest.fixture(scope="module")
def h250k(libmmg):
    gpu = make_hierarchy(SIDES_250K, "dirichlet", 4)
    return gpu, H.oracle_mirror_of_gpu(gpu, "dirichlet", 4)


@pytest.fixture(scope="module")
def h250k_p6(libmmg):
    gpu = make_hierarchy(SIDES_250K, "dirichlet", 6)
    return gpu, H.oracle_mirror_of_gpu(gpu, "dirichlet", 6)


@pytest.fixture(scope="module")
def h1m(libmmg):
    gpu = make_hierarchy(SIDES_1M, "dirichlet", 4)
    return gpu, H.oracle_mirror_of_gpu(gpu, "dirichlet", 4)


def clean_env(monkeypatch, **env):
    for k in MC_KNOBS:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)


def set_values(gpu, ref, level, seed, scale=1e-3):
    lv, g = ref.level(level), gpu.grid(level)
    v = scale * H.random_values(lv, seed)
    lv.set_vec(oracle.VEC_VALUES, v)
    g.values_ = v
    return lv, g


def test_mirror_is_faithful(h250k):
    gpu, ref = h250k
    for l in range(gpu.num_grids):
        g, lv = gpu.grid(l), ref.level(l)
        assert np.array_equal(g.bcFlags_, lv.bcflags())
        assert np.array_equal(g.source_, lv.source)
        for a, b in zip(g.csr()[1:], lv.csr()[1:]):
            assert np.array_equal(a, b)


def test_schedules_bit_exact_at_size(h250k):
    gpu, ref = h250k
    lv, g = ref.level(-1), gpu.grid(-1)
    nc, col = lv.colouring()
    nc2, col2 = g.colouring()
    assert nc == nc2 and np.array_equal(col, col2)
    lev = lv.lex_levels()
    nl2, lev2 = g.lex_levels()
    assert np.array_equal(lev, lev2) and nl2 == lev.max() + 1


# (environment, expected kernel prefix): every schedule of the fast multicolour sweep the dispatcher can pick on a big level
MC_VARIANTS = {
    "tma": (dict(MMG_MC_FLOW_MAX_ROWS="0"), "k_sor_mc_tma<8,5"),
    "packed": (dict(MMG_MC_FLOW_MAX_ROWS="0", MMG_MC_TMA="0"), "k_sor_mc_packed<8,5,2>"),
    "tma_flow": (dict(), "k_sor_mc_tma_flow<8,5,"),
    "flow": (dict(MMG_MC_TMAFLOW="0"), "k_sor_mc_flow<8,5,1>"),
}


@pytest.mark.parametrize("variant", sorted(MC_VARIANTS))
def test_fast_multicolour_sweep_at_size(h250k, monkeypatch, variant):
    """the instantiations bench.py times (finest level: TMA-fed sweep; next level: barrier-free sweep) and the register-fed fallback
    vs the oracle's Grid-level restatement of the multicolour smoother (rows of one colour updated from the state before the phase)"""
    gpu, ref = h250k
    env, kernel = MC_VARIANTS[variant]
    clean_env(monkeypatch, **env)
    gpu.set_arithmetic(capi.ARITH_FAST)
    lv, g = set_values(gpu, ref, -1, 71)
    lv.sor_multicolour(); g.sor(capi.MULTICOLOUR)
    assert capi.last_kernel(0).startswith(kernel), capi.last_kernel(0)
    assert H.rel_err(g.values_, lv.values) < 1e-12


def test_fast_multicolour_sweep_at_size_p6(h250k_p6, monkeypatch):
    gpu, ref = h250k_p6
    gpu.set_arithmetic(capi.ARITH_FAST)
    for env in (dict(MMG_MC_FLOW_MAX_ROWS="0"), dict()):
        clean_env(monkeypatch, **env)
        lv, g = set_values(gpu, ref, -1, 72)
        lv.sor_multicolour(); g.sor(capi.MULTICOLOUR)
        assert H.rel_err(g.values_, lv.values) < 1e-12, capi.last_kernel(0)


@pytest.mark.parametrize("arith,poly", [("reference_order", 4), ("fast", 4), ("fast", 6)])
def test_residual_restrict_prolong_at_size(h250k, h250k_p6, monkeypatch, arith, poly):
    """k_spmv2 / the TMA-fed SpMV on >= 200k rows (fast; polyDeg 4: 8 lanes per row, polyDeg 6: 16) and k_spmv_exact (reference order)"""
    gpu, ref = h250k if poly == 4 else h250k_p6
    clean_env(monkeypatch)
    exact = arith == "reference_order"
    gpu.set_arithmetic(capi.ARITH_REFERENCE_ORDER if exact else capi.ARITH_FAST)
    L = gpu.num_grids - 1

    def close(got, want):
        if exact:
            assert np.array_equal(got, want), H.rel_err(got, want)
        else:
            assert H.rel_err(got, want) < 1e-12

    fine, g = set_values(gpu, ref, L, 81, scale=1.0)
    close(g.residual(), fine.residual())
    assert abs(gpu.residual() - ref.residual()) <= 1e-12 * abs(ref.residual())
    if not exact:
        assert capi.last_kernel(1).startswith("k_spmv_tma<8,5" if poly == 4 else "k_spmv_tma<16,5"), capi.last_kernel(1)
    coarse = ref.level(L - 1)
    r = fine.residual()
    src = coarse.source
    src[: coarse.n] = fine.spmv(oracle.MAT_R, r[: fine.n])       # multigrid.cpp:81
    src = coarse.fix_vector_bound_coarse(src)                    # :82
    gpu.restrict(L)
    close(gpu.grid(L - 1).source_, src)
    vc = H.random_values(coarse, 83)
    coarse.set_vec(oracle.VEC_VALUES, vc); gpu.grid(L - 1).values_ = vc
    corr = np.zeros(fine.A); corr[: fine.n] = coarse.spmv(oracle.MAT_P, vc[: coarse.n])   # :102
    expect = fine.values + fine.fix_vector_bound_coarse(corr)     # :103-106
    gpu.prolong_correct(L)
    close(g.values_, expect)


def test_lexicographic_sweep_at_1m_is_bit_identical(h1m, monkeypatch):
    """k_sor_lex_chunk on a 1M-row level (DAG depth ~20k): five pipelined sweeps against Grid::sor (grid.cpp:104-146)"""
    gpu, ref = h1m
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_REFERENCE_ORDER)
    for level in (-1, -2):
        lv, g = set_values(gpu, ref, level, 91)
        lv.sor(); g.sor(capi.LEXICOGRAPHIC)
        assert capi.last_kernel(0).startswith("k_sor_lex_chunk"), capi.last_kernel(0)
        assert np.array_equal(g.values_, lv.values), H.rel_err(g.values_, lv.values)


def test_lexicographic_history_config2(h1m, monkeypatch):
    """BASELINE config 2 (1M nodes, 6 levels): five lexicographic V-cycles, residual history within 1e-10 per cycle"""
    gpu, ref = h1m
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_REFERENCE_ORDER)
    gpu.set_smoother(capi.LEXICOGRAPHIC)
    for l in range(gpu.num_grids):
        z = np.zeros(ref.level(l).A)
        ref.level(l).set_vec(oracle.VEC_VALUES, z); gpu.grid(l).values_ = z
    n0 = len(gpu.residuals_)
    ref.vcycle(5); gpu.vCycle(5)
    ho, hg = ref.history()[-5:], gpu.residuals_[n0:]
    assert np.abs(hg - ho).max() / ho.min() < 1e-6 and (np.abs(hg - ho) / ho).max() < 1e-10, (hg, ho)
    assert H.rel_l2(gpu.grid(-1).values_, ref.level(-1).values) < 1e-8
    assert ho[-1] < 0.8 * ho[0]          # six levels stop at a 32x32 cloud: ~0.92 per cycle


def test_multicolour_history_at_1m(h1m, monkeypatch):
    """the throughput mode end to end at 1M nodes: eight multicolour (omega 0.8) cycles vs the oracle's multicolour cycle,
    1e-10 of the initial residual per cycle (fast arithmetic reorders the row sums)"""
    gpu, ref = h1m
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_FAST); gpu.set_smoother(capi.MULTICOLOUR); gpu.set_omega(0.8)
    ref.set_smoother(1)
    for l in range(gpu.num_grids):
        lv = ref.level(l)
        z = np.zeros(lv.A)
        lv.set_vec(oracle.VEC_VALUES, z); gpu.grid(l).values_ = z
    ref.set_omega(0.8)
    n0 = len(gpu.residuals_)
    k0 = len(ref.history())
    ref.vcycle(8); gpu.vCycle(8)
    ho, hg = ref.history()[k0:], gpu.residuals_[n0:]
    assert np.abs(hg - ho).max() < 1e-10 * ho[0], (hg, ho)
    gpu.set_omega(1.4); ref.set_omega(1.4); ref.set_smoother(0)


# ---------------------------------------------------------------- Neumann-type grids on the fused (TMA-fed) sweep
# Gmsh-like clouds (clouds.hex_square): the reference's scheme converges on them with Neumann / mixed boundaries.
# Regularisation row as an in-kernel reduction phase, boundary evaluation in the sweep tail, overflow rows (grid.cpp:73-103,
# 566-576, 607-657) -- every sweep of a smoothing call in ONE launch, against Grid-level restatements in the oracle.
HEX_SIDES = {"small": [13, 25, 50, 100, 200], "big": [25, 50, 100, 200, 450]}


@pytest.fixture(scope="module", params=[("neumann", "small"), ("mixed", "small"), ("neumann", "big"), ("mixed", "big")], ids=lambda p: "%s-%s" % p)
def hneu(request, libmmg):
    kind, size = request.param
    gpu = make_hierarchy(HEX_SIDES[size], kind, 4, cloud="hex")
    return kind, gpu, H.oracle_mirror_of_gpu(gpu, kind, 4)


def test_neumann_fused_sweep_matches_the_oracle(hneu, monkeypatch):
    kind, gpu, ref = hneu
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_FAST)
    for level in (-1, -2):
        lv, g = set_values(gpu, ref, level, 101)
        nc, col = lv.colouring()
        nc2, col2 = g.colouring()
        assert nc == nc2 and np.array_equal(col, col2)
        lv.sor_multicolour(); g.sor(capi.MULTICOLOUR)
        assert capi.last_kernel(0).startswith("k_sor_mc_tma"), capi.last_kernel(0)
        # five compounded sweeps incl. the near-boundary rows of the implicit elimination (diagonal -2e2 against off-diagonals summing to
        # 8e4): 1e-11 as in test_gpu_operators.py::test_sor; the single-pass operators below are held to 1e-12
        assert H.rel_err(g.values_, lv.values) < 1e-11, (kind, level)


def test_neumann_operators_at_size(hneu, monkeypatch):
    kind, gpu, ref = hneu
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_FAST)
    L = gpu.num_grids - 1
    fine, g = set_values(gpu, ref, L, 111, scale=1.0)
    assert H.rel_err(g.residual(), fine.residual()) < 1e-12
    assert abs(gpu.residual() - ref.residual()) <= 1e-12 * abs(ref.residual())
    lv2, g2 = set_values(gpu, ref, L, 112, scale=1.0)
    lv2.bound_eval_neumann(); g2.bound_eval_neumann()
    assert H.rel_err(g2.values_, lv2.values) < 1e-12


def test_neumann_multicolour_history(hneu, monkeypatch):
    """six throughput-mode cycles (fused sweeps on every level) against the oracle's multicolour cycle.  The multicolour
    ordering does NOT converge on Neumann-type problems (only the reference's lexicographic order does), so
    the history grows -- the comparison is relative per cycle, which also holds while it diverges."""
    kind, gpu, ref = hneu
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_FAST); gpu.set_smoother(capi.MULTICOLOUR); gpu.set_omega(0.8)
    ref.set_smoother(1); ref.set_omega(0.8)
    for l in range(gpu.num_grids):
        lv = ref.level(l)
        z = np.zeros(lv.A)
        lv.set_vec(oracle.VEC_VALUES, z); gpu.grid(l).values_ = z
    n0, k0 = len(gpu.residuals_), len(ref.history())
    ref.vcycle(6); gpu.vCycle(6)
    ho, hg = ref.history()[k0:], gpu.residuals_[n0:]
    assert (np.abs(hg - ho) / ho).max() < 1e-9, (hg, ho)
    gpu.set_omega(1.4); ref.set_omega(1.4); ref.set_smoother(0)


def test_neumann_lexicographic_history(hneu, monkeypatch):
    """the reference-faithful mode on the Gmsh-like clouds, where the reference's scheme converges with Neumann and mixed
    boundaries: residual history within 1e-10 per cycle, solution within 1e-8 (north-star bars)"""
    kind, gpu, ref = hneu
    clean_env(monkeypatch)
    gpu.set_arithmetic(capi.ARITH_REFERENCE_ORDER); gpu.set_smoother(capi.LEXICOGRAPHIC); gpu.set_omega(1.4)
    ref.set_smoother(0); ref.set_omega(1.4)
    for l in range(gpu.num_grids):
        lv = ref.level(l)
        z = np.zeros(lv.A)
        lv.set_vec(oracle.VEC_VALUES, z); gpu.grid(l).values_ = z
    n0, k0 = len(gpu.residuals_), len(ref.history())
    ref.vcycle(6); gpu.vCycle(6)
    ho, hg = ref.history()[k0:], gpu.residuals_[n0:]
    assert (np.abs(hg - ho) / ho).max() < 1e-10, (hg, ho)
    assert ho[-1] < 0.6 * ho[0], ho                      # it converges here (it diverges on the jittered lattices beyond ~10k nodes)
    assert H.rel_l2(gpu.grid(-1).values_, ref.level(-1).values) < 1e-8
