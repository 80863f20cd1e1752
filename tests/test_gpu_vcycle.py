"""V-cycle parity: residual history within 1e-10 relative per cycle (lexicographic, reference-faithful
mode against the oracle; multicolour mode against the oracle's multicolour restatement) and final
solution within 1e-8 relative L2."""
import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_HISTORY = 1e-10     # per cycle, relative (north star)
TOL_SOLUTION = 1e-8     # relative L2 (north star)
FLOOR = 1e-11           # below this the residual is rounding noise of the fp64 solve itself


def run_pair(sizes, kind, fine_poly, cycles, multicolour=False, fracstep=False, fast=False):
    mg = oracle.make_hierarchy(sizes, kind=kind, fine_poly=fine_poly, fracstep=fracstep)
    s = H.gpu_solver_from_oracle(mg, fracstep=fracstep)
    s.set_arithmetic(capi.ARITH_FAST if fast else capi.ARITH_REFERENCE_ORDER)
    if fracstep:
        mg.L.orc_fs_step_pre(mg.h, mg.nlevels - 1)
        s.grid(-1).source_ = mg.level(-1).source
    mg.set_multicolour(multicolour)
    s.set_smoother(capi.MULTICOLOUR if multicolour else capi.LEXICOGRAPHIC)
    mg.vcycle(cycles)
    s.vCycle(cycles)
    return mg, s


def check_history(mg, s):
    """reference-order arithmetic: every cycle within 1e-10 RELATIVE of the oracle, down to the floor"""
    ho, hg = mg.history(), s.residuals_
    assert ho.size == hg.size
    live = np.isfinite(ho) & (ho > FLOOR * ho[0]) & (ho < 1e100)
    assert live.sum() >= min(5, ho.size)
    rel = np.abs(hg[live] - ho[live]) / ho[live]
    assert rel.max() < TOL_HISTORY, rel


def check_history_fast(mg, s):
    """throughput arithmetic (reordered sums): the history agrees to 1e-10 of the INITIAL residual; the relative
    deviation necessarily grows like eps*cond/residual as the residual approaches its rounding floor"""
    ho, hg = mg.history(), s.residuals_
    assert ho.size == hg.size
    assert np.abs(hg - ho).max() < TOL_HISTORY * ho[0]


@pytest.mark.parametrize("fine_poly", [4, 6])
@pytest.mark.parametrize("multicolour", [False, True])
def test_dirichlet_config1(libmmg, fine_poly, multicolour):
    # BASELINE config 1: ~10k nodes, 4 levels (13/25/50/100 lattice sides)
    mg, s = run_pair([13, 25, 50, 100], oracle.KIND_DIRICHLET, fine_poly, 25, multicolour)
    check_history(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


@pytest.mark.parametrize("multicolour", [False, True])
def test_mixed_bc(libmmg, multicolour):
    mg, s = run_pair([13, 25, 50], oracle.KIND_MIXED, 4, 20, multicolour)
    check_history(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


@pytest.mark.parametrize("multicolour", [False, True])
def test_pure_neumann(libmmg, multicolour):
    mg, s = run_pair([13, 25, 50], oracle.KIND_NEUMANN, 3, 20, multicolour)
    check_history(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


def test_fracstep_ppe(libmmg):
    mg, s = run_pair([13, 25, 50], oracle.KIND_PPE, 3, 15, fracstep=True)
    check_history(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


def test_two_level_quirk(libmmg):
    # 2-level hierarchy: boundaryOp("coarse") of multigrid.cpp:91 zeroes the FINEST grid's Dirichlet values
    mg, s = run_pair([25, 50], oracle.KIND_DIRICHLET, 4, 10)
    check_history(mg, s)


def test_single_grid_fracstep_shortcut(libmmg):
    mg, s = run_pair([25], oracle.KIND_PPE, 3, 3, fracstep=True)   # FracStepMultigrid.cpp:64-67
    assert mg.history().size == 0 and s.residuals_.size == 0
    assert H.rel_err(s.grid(0).values_, mg.level(0).values) < 1e-11


def test_solve_to_tolerance_matches_cycle_count(libmmg):
    mg = oracle.make_hierarchy([13, 25, 50, 100], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    s = H.gpu_solver_from_oracle(mg)
    n_o, _ = mg.solve(1e-8, 200)
    n_g, r = s.solve(1e-8, 200)
    assert n_o == n_g and r < 1e-8
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


def test_multicolour_and_lexicographic_reach_the_same_solution(libmmg):
    mg = oracle.make_hierarchy([13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    a, b = H.gpu_solver_from_oracle(mg), H.gpu_solver_from_oracle(mg)
    b.set_smoother(capi.MULTICOLOUR)
    a.solve(1e-12, 300); b.solve(1e-12, 300)
    assert H.rel_l2(b.grid(-1).values_, a.grid(-1).values_) < TOL_SOLUTION


@pytest.mark.parametrize("multicolour", [False, True])
def test_fast_arithmetic_history(libmmg, multicolour):
    mg, s = run_pair([13, 25, 50, 100], oracle.KIND_DIRICHLET, 4, 25, multicolour, fast=True)
    check_history_fast(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


# ---------------------------------------------------------------- block-lexicographic smoother (throughput mode)
@pytest.mark.parametrize("block", [64, 256, 4096])
def test_block_lexicographic_sor_is_bit_identical_to_the_oracle(libmmg, block):
    mg = oracle.make_hierarchy([13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    s = H.gpu_solver_from_oracle(mg)
    s.set_block_size(block)
    for l in range(mg.nlevels):
        lv, g = mg.level(l), s.grid(l)
        nc, col = lv.block_colouring(block)
        nc2, col2 = g.block_colouring()
        assert nc == nc2 and np.array_equal(col, col2)                     # integer artefact: bit-exact
        v = 1e-3 * H.random_values(lv, 51 + l)
        lv.set_vec(oracle.VEC_VALUES, v); g.values_ = v
        lv.sor_blocklex(block); g.sor(capi.BLOCK_LEXICOGRAPHIC)
        assert np.array_equal(g.values_, lv.values), (l, H.rel_err(g.values_, lv.values))


@pytest.mark.parametrize("block", [256, 4096])
def test_block_lexicographic_vcycle_history(libmmg, block):
    mg = oracle.make_hierarchy([13, 25, 50, 100], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    s = H.gpu_solver_from_oracle(mg)
    mg.set_smoother(2, block)
    s.set_smoother(capi.BLOCK_LEXICOGRAPHIC); s.set_block_size(block)
    mg.vcycle(20); s.vCycle(20)
    check_history(mg, s)
    assert H.rel_l2(s.grid(-1).values_, mg.level(-1).values) < TOL_SOLUTION


def test_block_lexicographic_with_one_block_is_the_reference_sweep(libmmg):
    mg = oracle.make_hierarchy([13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    a, b = H.gpu_solver_from_oracle(mg), H.gpu_solver_from_oracle(mg)
    b.set_smoother(capi.BLOCK_LEXICOGRAPHIC); b.set_block_size(1 << 20)
    a.vCycle(8); b.vCycle(8)
    assert np.array_equal(a.residuals_, b.residuals_)


def test_fast_multicolour_sweep_variants_are_bit_identical(libmmg, monkeypatch):
    """The throughput-mode multicolour sweep has several schedules of the same per-row arithmetic: per-colour launches over
    the natural-order operator; over the colour-major packed copy the TMA-fed ring with counter barriers (k_sor_mc_tma), the
    register-fed kernel with grid.sync (k_sor_mc_packed), the barrier-free sweep register-fed (k_sor_mc_flow) or TMA-fed (k_sor_mc_tma_flow), and on the coarsest levels
    the single-CTA kernels with the vectors (k_sor_mc_small) or the whole operator (k_sor_mc_resident) in shared memory.
    Rows of one colour are independent, so all must produce the same bits -- a stale read across a barrier shows up here."""
    from meshlessmultigridpoisson_b200.problems import make_hierarchy

    knobs = ("MMG_MC_PACKED", "MMG_MC_SMALL", "MMG_MC_FLOW", "MMG_MC_TMA", "MMG_MC_RESIDENT", "MMG_MC_FLOW_MAX_ROWS", "MMG_TMA_ROWS", "MMG_TMA_DYNAMIC", "MMG_MC_TMAFLOW_MIN_ROWS", "MMG_MC_TMAFLOW")
    results = []
    for env in ({"MMG_MC_PACKED": "0"}, {}, {"MMG_MC_TMAFLOW_MIN_ROWS": "0"}, {"MMG_MC_TMAFLOW": "0"}, {"MMG_MC_FLOW": "0"}, {"MMG_MC_FLOW": "0", "MMG_TMA_ROWS": "1", "MMG_TMA_DYNAMIC": "0"},
                {"MMG_MC_FLOW": "0", "MMG_MC_TMA": "0"}, {"MMG_MC_SMALL": "0"}, {"MMG_MC_RESIDENT": "0"}, {"MMG_MC_RESIDENT": "0", "MMG_MC_SMALL": "0"}):
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        mg = make_hierarchy([19, 38, 75, 150, 300], "dirichlet", 4)
        mg.set_smoother(capi.MULTICOLOUR); mg.set_arithmetic(capi.ARITH_FAST); mg.set_omega(0.8)
        mg.vCycle(8)
        results.append((np.array(mg.residuals_), mg.grid(-1).values_))
    for hist, vals in results[1:]:
        assert np.array_equal(hist, results[0][0]) and np.array_equal(vals, results[0][1])
    assert results[0][0][-1] < 0.6 * results[0][0][0]


def test_two_gpu_partitioned_vcycle_matches_single_gpu(libmmg):
    """Row-partitioned V-cycle over 2 GPUs (peer-memory smoother + NCCL halos) against the single-GPU V-cycle: owned entries
    bit-identical, history to 1e-12.  Needs two devices; the driver's single-GPU box skips it."""
    import os, subprocess, sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(root, "scripts", "dist_vcycle_check.py"), "300", "4"]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "owned solution identical True" in out.stdout
