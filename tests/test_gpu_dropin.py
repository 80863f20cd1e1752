"""Drop-in proof: the reference's UNMODIFIED drivers on the GPU, and the hole / annulus geomtypes.

oracle/_ref/dropin_driver is built (oracle/Makefile, target dropin, in the container that has /root/reference) from the
reference's own testing_functions.cpp, FractionalStepSim.cpp, fileReadingFunctions.cpp and general_computation_functions.cpp,
compiled where they lie against meshlessmultigridpoisson_b200/cpp/dropin/ and linked with libmmg.so.  It runs the reference's
run_mg_sim (testing_functions.cpp:328-350) and the time-loop body of run_fracstep_param (FractionalStepSim.cpp:130-147) on Gmsh
$Nodes files written here, and the results are compared with oracle/_ref/libref.so -- the same reference sources over the Eigen
shim on the CPU -- reading the same files.

Tolerance: the driver assembles its operators on the device, whose stencil weights differ from the CPU factorisation at
cond*eps (entries to 1e-9 .. 1e-8 at polyDeg 3 / 4, test_gpu_assembly.py), so histories agree to ~1e-6 here; the 1e-10-per-cycle
bar is tested on identical matrices in test_gpu_vcycle.py / test_gpu_at_size.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200 import capi
from meshlessmultigridpoisson_b200.clouds import make_cloud, write_msh_nodes
from meshlessmultigridpoisson_b200.problems import make_hierarchy
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref.so")


def _ref():
    if not (os.path.exists(DRIVER) and os.path.exists(REF_SO)):
        pytest.skip("oracle/_ref/dropin_driver not built (needs the reference checkout at build time)")
    R = C.CDLL(REF_SO)
    dp = np.ctypeslib.ndpointer(np.float64)
    R.ref_new_geom.restype = C.c_void_p
    R.ref_new_geom.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    R.ref_vcycle.argtypes = [C.c_void_p, C.c_int]
    R.ref_history.argtypes = [C.c_void_p, dp, C.c_int]
    R.ref_lv_vec.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    R.ref_lv_A.argtypes = [C.c_void_p, C.c_int]
    R.ref_lv_n.argtypes = [C.c_void_p, C.c_int]
    R.ref_residual.argtypes = [C.c_void_p]
    R.ref_residual.restype = C.c_double
    R.ref_fs_step_pre.argtypes = [C.c_void_p]
    R.ref_fs_step_post.argtypes = [C.c_void_p]
    R.ref_fs_step_post.restype = C.c_double
    R.ref_fine_bound_eval.argtypes = [C.c_void_p]
    R.ref_fs_vec.argtypes = [C.c_void_p, C.c_int, dp]
    return R


def _write_levels(tmp_path, cloud, sizes):
    names = []
    for l, s in enumerate(sizes):
        x, y = make_cloud(cloud, s, 1000 + l)
        write_msh_nodes(str(tmp_path / ("l%d.msh" % l)), x, y)
        names.append("l%d.msh" % l)
    return names


def _parse(stdout, key):
    for line in stdout.splitlines():
        if line.startswith(key + " ") or line == key:
            return np.array([float(t) for t in line.split()[1:]])
    raise AssertionError("no %s line in driver output" % key)


@pytest.mark.parametrize("geomtype,neumann,poly", [("square", 0, 4), ("square", 1, 3), ("square_with_circle", 0, 4), ("concentric_circles", 0, 4),
                                                   ("square_with_circle", 1, 3)])
def test_unmodified_run_mg_sim_on_the_gpu(libmmg, tmp_path, geomtype, neumann, poly):
    R = _ref()
    sizes, cycles = [13, 25, 40], 6
    names = _write_levels(tmp_path, "jittered" if geomtype == "square" else geomtype, sizes)
    d = str(tmp_path) + "/"
    out = subprocess.run([DRIVER, "mg", geomtype, str(neumann), d, "t", str(cycles), str(poly)] + names, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    hist, vals = _parse(out.stdout, "HISTORY"), _parse(out.stdout, "VALUES")
    files = [(d + n).encode() for n in names]
    polys = [3] * (len(sizes) - 1) + [poly]
    h = C.c_void_p(R.ref_new_geom(neumann, geomtype.encode(), len(files), (C.c_char_p * len(files))(*files), (C.c_int * len(files))(*polys), 1, 1, 2e-4, 0.025, 1.0))
    R.ref_vcycle(h, cycles)
    href = np.empty(cycles)
    R.ref_history(h, href, cycles)
    vref = np.empty(R.ref_lv_A(h, len(sizes) - 1))
    R.ref_lv_vec(h, len(sizes) - 1, 0, vref)
    assert hist.size == cycles and vals.size == vref.size
    assert (np.abs(hist - href) / href).max() < 1e-6, (hist, href)
    assert H.rel_l2(vals, vref) < 1e-6 * max(1.0, href[-1] / href[0])
    # the files run_mg_sim itself wrote (6 significant digits, fileReadingFunctions.cpp:70-79)
    resid = np.loadtxt(d + "resid_t.txt")
    assert resid.size == cycles and np.allclose(resid, href, rtol=2e-5)
    temp = np.loadtxt(d + "temp_t.txt")
    assert temp.size == vref.size - 1                                     # rows()-1 entries, testing_functions.cpp:303
    assert os.path.exists(d + "cond_error_t.txt") and os.path.exists(d + "x_t.txt")


def test_unmodified_fractional_step_loop_on_the_gpu(libmmg, tmp_path):
    R = _ref()
    sizes, steps = [13, 25], 2
    names = _write_levels(tmp_path, "jittered", sizes)
    d = str(tmp_path) + "/"
    out = subprocess.run([DRIVER, "fs", d, str(steps), "3"] + names, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    u, p = _parse(out.stdout, "U"), _parse(out.stdout, "P")
    files = [(d + n).encode() for n in names]
    h = C.c_void_p(R.ref_new_geom(2, b"square", len(files), (C.c_char_p * len(files))(*files), (C.c_int * len(files))(3, 3), 1, 1, 2e-4, 0.025, 1.0))
    for _ in range(steps):                                                # FractionalStepSim.cpp:130-147
        R.ref_fs_step_pre(h)
        guard = 0
        while R.ref_residual(h) >= 1e-10 and guard < 400:
            R.ref_vcycle(h, 1); R.ref_fine_bound_eval(h); guard += 1
        R.ref_fs_step_post(h)
    n = R.ref_lv_n(h, len(sizes) - 1)
    uref = np.empty(n)
    R.ref_fs_vec(h, 0, uref)
    assert u.size == n and H.rel_l2(u, uref) < 1e-6
    pref = np.empty(n + 1)
    R.ref_lv_vec(h, len(sizes) - 1, 0, pref)
    assert H.rel_l2(p, pref[:n]) < 1e-5


GEOM_CASES = [("square_with_circle", "dirichlet"), ("square_with_circle", "neumann"), ("concentric_circles", "dirichlet"), ("concentric_circles", "neumann")]


@pytest.mark.parametrize("geomtype,kind", GEOM_CASES)
def test_hole_and_annulus_hierarchies_match_the_oracle(libmmg, geomtype, kind):
    """genGmshGridDirichlet / genGmshGridNeumann for the two-boundary geomtypes (testing_functions.cpp:92-135, 186-251) with the
    analytic radial normals of grid.cpp:480-516: device-built level state against the oracle (pinned bit-identical to the reference
    sources for these geomtypes in test_oracle_cpu.py), then V-cycles on identical matrices at 1e-10."""
    sizes = [13, 25, 50]
    geom = oracle.GEOM_NAMES.index(geomtype)
    okind = oracle.KIND_DIRICHLET if kind == "dirichlet" else oracle.KIND_NEUMANN
    mg = oracle.make_hierarchy(sizes, kind=okind, fine_poly=4, geom=geom)
    s = make_hierarchy(sizes, kind, 4, geomtype=geomtype)
    L = mg.nlevels
    for l in range(L):
        lv, g = mg.level(l), s.grid(l)
        assert np.array_equal(g.perm(), lv.perm())                        # reordering (two boundaries)
        assert np.array_equal(g.bcFlags_, lv.bcflags())
        for b, (t, pts, vals) in enumerate(lv.boundaries()):
            t2, p2, v2 = g.boundary(b)
            assert t == t2 and np.array_equal(pts, p2) and np.array_equal(vals, v2)
        if kind == "neumann":
            nx, ny = lv.normals(); gx, gy = g.normalVecs_
            assert np.array_equal(nx, gx) and np.array_equal(ny, gy)      # analytic normals: same expression, same bits
        shape, ptr, idx, val = g.csr()
        o = lv.csr()
        assert np.array_equal(o[1], ptr) and np.array_equal(o[2], idx) and H.rel_err(val, o[3]) < 1e-5
        lv.set_csr(oracle.MAT_A, shape, ptr, idx, val)
        if kind == "neumann":
            lv.set_csr(oracle.MAT_NBC, *g.csr(capi.MAT_NEUMANN_COEFFS))
            lv.set_vec(oracle.VEC_DIAGS, g.diags)
        assert H.rel_err(g.source_, lv.source) < 1e-8
        lv.set_vec(oracle.VEC_SOURCE, g.source_)
    for l in range(1, L):
        mg.level(l).set_csr(oracle.MAT_R, *s.interp_csr(capi.MAT_RESTRICT, l))
    for l in range(L - 1):
        mg.level(l).set_csr(oracle.MAT_P, *s.interp_csr(capi.MAT_PROLONG, l))
    mg.vcycle(6); s.vCycle(6)
    ho, hg = mg.history(), s.residuals_
    assert (np.abs(hg - ho) / ho).max() < 1e-10, (hg, ho)
