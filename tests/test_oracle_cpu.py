"""CPU suite (`-m "not gpu"`): pins the oracle — against the reference's own sources compiled on the Eigen shim
(oracle/_ref, when /root/reference is present), against the committed golden vectors, and against the analytic
known-answer properties the reference implies — and checks the host-side logic around the C-ABI."""
import ctypes as C
import os
import re
import tempfile

import numpy as np
import pytest

import oracle
from meshlessmultigridpoisson_b200.clouds import jittered_square, level_sizes, write_msh_nodes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref.so")
REF_SRC = "/root/reference/MeshlessPoisson"


# ---------------------------------------------------------------- oracle vs the reference's own code
def _ref_lib():
    if os.path.isdir(REF_SRC):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref.so not built (reference checkout absent)")
    R = C.CDLL(REF_SO)
    dp, ip = np.ctypeslib.ndpointer(np.float64), np.ctypeslib.ndpointer(np.int32)
    R.ref_new.restype = C.c_void_p
    R.ref_new.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    R.ref_new_geom.restype = C.c_void_p
    R.ref_new_geom.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    R.ref_lv_nnz.restype = C.c_long
    R.ref_lv_nnz.argtypes = [C.c_void_p, C.c_int]
    R.ref_lv_csr.argtypes = [C.c_void_p, C.c_int, ip, ip, dp]
    R.ref_lv_vec.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    R.ref_lv_points.argtypes = [C.c_void_p, C.c_int, dp, dp]
    R.ref_vcycle.argtypes = [C.c_void_p, C.c_int]
    R.ref_history.argtypes = [C.c_void_p, dp, C.c_int]
    R.ref_fs_step_pre.argtypes = [C.c_void_p]
    R.ref_fs_step_post.argtypes = [C.c_void_p]
    R.ref_fs_step_post.restype = C.c_double
    R.ref_fs_vec.argtypes = [C.c_void_p, C.c_int, dp]
    R.ref_interp_nnz.restype = C.c_long
    R.ref_interp_nnz.argtypes = [C.c_void_p, C.c_int, C.c_int]
    R.ref_interp_csc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), ip, ip, dp]
    return R


@pytest.mark.parametrize("kind,fine_poly,geom", [(oracle.KIND_DIRICHLET, 4, 0), (oracle.KIND_DIRICHLET, 6, 0), (oracle.KIND_NEUMANN, 4, 0), (oracle.KIND_PPE, 3, 0),
                                                 (oracle.KIND_DIRICHLET, 4, 1), (oracle.KIND_NEUMANN, 3, 1), (oracle.KIND_DIRICHLET, 4, 2), (oracle.KIND_NEUMANN, 3, 2)])
def test_oracle_is_bit_identical_to_the_reference_sources(kind, fine_poly, geom):
    """The reference's grid.cpp / multigrid.cpp / FracStepMultigrid.cpp / fractionalStepGrid.cpp, compiled where they lie
    against the Eigen-subset shim and driven through its own Gmsh reader and factories, must agree with the oracle
    restatement bit for bit: operators, right-hand sides, residual history, solution."""
    R = _ref_lib()
    sizes = [13, 25, 40] if geom == 0 else [13, 25, 30]      # the reference's kNN is brute force: O(N^2) per pass
    tmp = tempfile.mkdtemp()
    files = []
    from meshlessmultigridpoisson_b200.clouds import make_cloud
    for l, s in enumerate(sizes):
        x, y = make_cloud("jittered" if geom == 0 else oracle.GEOM_NAMES[geom], s, 1000 + l)   # geom 1, 2: the hole and annulus geomtypes (grid.cpp:480-516)
        fn = os.path.join(tmp, "l%d.msh" % l)
        write_msh_nodes(fn, x, y)
        files.append(fn.encode())
    polys = [3] * (len(sizes) - 1) + [fine_poly]
    h = C.c_void_p(R.ref_new_geom(kind, oracle.GEOM_NAMES[geom].encode(), len(files), (C.c_char_p * len(files))(*files), (C.c_int * len(files))(*polys), 1, 1,
                                  2e-4, 0.025, 1.0))
    mg = oracle.make_hierarchy(sizes, kind=kind, fine_poly=fine_poly, cells=False, fracstep=(kind == oracle.KIND_PPE), geom=geom)
    for l in range(len(sizes)):
        lv = mg.level(l)
        (r, _), ptr, idx, val = lv.csr()
        nnz = R.ref_lv_nnz(h, l)
        assert nnz == idx.size
        p2, i2, v2 = np.empty(r + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
        R.ref_lv_csr(h, l, p2, i2, v2)
        assert np.array_equal(ptr, p2) and np.array_equal(idx, i2) and np.array_equal(val, v2)
        s2 = np.empty(r)
        R.ref_lv_vec(h, l, 1, s2)
        assert np.array_equal(s2, lv.source)
        x2, y2 = np.empty(lv.n), np.empty(lv.n)
        R.ref_lv_points(h, l, x2, y2)
        x1, y1 = lv.points()
        assert np.array_equal(x1, x2) and np.array_equal(y1, y2)          # same reordering
    for which, sel in ((0, oracle.MAT_R), (1, oracle.MAT_P)):               # interpolation matrices (column-major in the reference)
        for l in range(len(sizes)):
            nnz = R.ref_interp_nnz(h, which, l)
            if nnz < 0:
                continue
            rows, cols = C.c_int(), C.c_int()
            shape, ptr, idx, val = mg.level(l).csr(sel)
            cp, ci, cv = np.empty(shape[1] + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
            R.ref_interp_csc(h, which, l, rows, cols, cp, ci, cv)
            import scipy.sparse as sp
            ref = sp.csc_matrix((cv, ci, cp), shape=(rows.value, cols.value)).tocsr()
            ref.sort_indices()
            assert (rows.value, cols.value) == shape and np.array_equal(ref.indptr, ptr) and np.array_equal(ref.indices, idx) and np.array_equal(ref.data, val)
    if kind == oracle.KIND_PPE:
        R.ref_fs_step_pre(h)
        mg.L.orc_fs_step_pre(mg.h, mg.nlevels - 1)
    R.ref_vcycle(h, 6)
    mg.vcycle(6)
    hr = np.empty(6)
    R.ref_history(h, hr, 6)
    assert np.array_equal(hr, mg.history())
    v2 = np.empty(mg.level(-1).A)
    R.ref_lv_vec(h, len(sizes) - 1, 0, v2)
    assert np.array_equal(v2, mg.level(-1).values)
    if kind == oracle.KIND_PPE:
        assert R.ref_fs_step_post(h) == mg.L.orc_fs_step_post(mg.h, mg.nlevels - 1)
        u2 = np.empty(mg.level(-1).n)
        R.ref_fs_vec(h, 0, u2)
        assert np.array_equal(u2, mg.level(-1).vec(oracle.VEC_U))


def test_reference_objects_on_oracle_operators_cycle_bit_identically():
    """bench.py's reference arm: the reference's own Multigrid / Grid objects (libref.so) fed with the oracle's matrices
    (ReferenceHierarchy.from_oracle -- the reference's brute-force set-up is out of reach at 4M nodes) must cycle exactly like
    the oracle and like the reference built by its own factories."""
    _ref_lib()
    if not oracle.ReferenceHierarchy.available():
        pytest.skip("oracle/_ref/libref.so not built")
    for sizes, poly in (([13, 25, 40], 4), ([16, 32, 63], 6)):
        mg = oracle.make_hierarchy(sizes, kind=oracle.KIND_DIRICHLET, fine_poly=poly)
        ref = oracle.ReferenceHierarchy.from_oracle(mg)
        assert ref.residual() == mg.residual() == 1.0
        assert ref.time_vcycles(5) > 0
        mg.vcycle(5)
        assert np.array_equal(ref.history(), mg.history())
        assert np.array_equal(ref.values(), mg.level(-1).values)
        assert ref.residual() == mg.residual() < 0.2
    neu = oracle.make_hierarchy([13, 25], kind=oracle.KIND_NEUMANN, fine_poly=3)
    with pytest.raises(oracle.OracleError):
        oracle.ReferenceHierarchy.from_oracle(neu)        # Dirichlet levels only


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", ["dirichlet_p4", "dirichlet_p6", "mixed_p4", "neumann_p3"])
def test_oracle_reproduces_golden_vectors(name):
    from tests.make_golden import CASES, digest

    G = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = CASES[name]
    mg = oracle.make_hierarchy(c["sizes"], kind=c["kind"], fine_poly=c["fine_poly"], cells=True)   # cell-grid kNN == brute force
    for l in range(mg.nlevels):
        lv = mg.level(l)
        _, ptr, idx, val = lv.csr()
        assert np.array_equal(lv.perm(), G["perm%d" % l])
        assert digest(ptr, idx, val) == str(G["csr_digest%d" % l])
        assert np.array_equal(lv.source, G["source%d" % l])
        assert np.array_equal(lv.colouring()[1], G["colour%d" % l])
        assert np.array_equal(lv.lex_levels(), G["lex_levels%d" % l])
    mg.vcycle(c["cycles"])
    assert np.array_equal(mg.history(), G["history"])
    assert np.array_equal(mg.level(-1).values, G["values"])


# ---------------------------------------------------------------- pieces
def test_pow2_is_a_multiplication_in_this_build():
    """distance() is sqrt(pow(dx,2)+pow(dy,2)) (general_computation_functions.cpp:4-6); the kNN tie-break contract needs
    pow(t,2) == t*t bit for bit, which is what the device computes."""
    L = oracle.lib()
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2000, 4))
    for ax, ay, bx, by in a:
        dx, dy = ax - bx, ay - by
        assert L.orc_distance(ax, ay, bx, by) == np.sqrt(dx * dx + dy * dy)


def test_knn_cells_equals_brute_force_including_exclusions():
    x, y = jittered_square(40, seed=5)
    for kind in (oracle.KIND_DIRICHLET, oracle.KIND_NEUMANN, oracle.KIND_MIXED):
        mg = oracle.Multigrid()
        mg.add_level(kind, x, y, 4, fine=True, cells=True)
        lv = mg.level(0)
        px, py = lv.points()
        f = lv.bcflags()
        for i in list(range(0, lv.n, 37)) + np.nonzero(f)[0][:60].tolist():
            for k in (25, 37, 70):
                a = lv.knn(px[i], py[i], k, neumann=lv.neumann, point_bc=bool(f[i]), cells=False)
                b = lv.knn(px[i], py[i], k, neumann=lv.neumann, point_bc=bool(f[i]), cells=True)
                assert np.array_equal(a, b)
                assert a[0] == i                                   # neighbour 0 is the node itself
    rng = np.random.default_rng(1)
    for qx, qy in rng.uniform(-0.1, 1.1, (100, 2)):               # off-grid queries, some outside the bounding box
        assert np.array_equal(lv.knn(qx, qy, 37, cells=False), lv.knn(qx, qy, 37, cells=True))


def test_knn_tie_break_is_distance_then_index():
    # perfect lattice: many exactly equal distances; the reference orders ties by index (std::pair comparison)
    x, y = jittered_square(12, seed=0, jitter=0.0)
    mg = oracle.Multigrid()
    mg.add_level(oracle.KIND_DIRICHLET, x, y, 3, fine=True, cells=False)
    lv = mg.level(0)
    px, py = lv.points()
    i = lv.n // 2
    nn = lv.knn(px[i], py[i], 25, cells=False)
    d = np.sqrt((px - px[i]) ** 2 + (py - py[i]) ** 2)
    order = np.lexsort((np.arange(lv.n), d))[:25]
    assert np.array_equal(nn, order)
    assert np.array_equal(nn, lv.knn(px[i], py[i], 25, cells=True))


def test_fullpivlu_solves_and_picks_the_first_maximum():
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for n in (1, 2, 7, 35, 52):
        A = rng.standard_normal((n, n))
        b = rng.standard_normal(n)
        x = np.empty(n)
        L.orc_fullpivlu_solve(n, np.asfortranarray(A).ravel(order="K").copy(), b, x)
        assert np.allclose(A @ x, b, atol=1e-9 * max(1, np.abs(b).max()) * np.linalg.cond(A))
    # rank deficient: Eigen zero-fills beyond rank()
    A = np.array([[1.0, 2.0], [2.0, 4.0]])
    x = np.empty(2)
    L.orc_fullpivlu_solve(2, np.asfortranarray(A).ravel(order="K").copy(), np.array([1.0, 2.0]), x)
    assert np.allclose(A @ x, [1.0, 2.0])


def test_set_from_triplets_sums_duplicates_in_order_and_keeps_zeros():
    L = oracle.lib()
    r = np.array([0, 0, 1, 0, 1, 1], np.int32)
    c = np.array([2, 0, 1, 2, 1, 0], np.int32)
    v = np.array([1e16, 3.0, 5.0, 1.0, -5.0, 0.0])
    ptr, idx, val = np.empty(3, np.int32), np.empty(6, np.int32), np.empty(6)
    nnz = L.orc_csr_from_triplets(2, 3, 6, r, c, v, ptr, idx, val)
    assert nnz == 4
    assert ptr.tolist() == [0, 2, 4] and idx[:4].tolist() == [0, 2, 0, 1]
    assert val[:4].tolist() == [3.0, 1e16 + 1.0, 0.0, 0.0]          # (0,2): 1e16+1 in triplet order; (1,1): 5-5 kept as explicit zero


def test_bfs_order_is_plain_bfs_from_node_zero_reversed():
    adj = [[0, 2, 1], [1, 3], [2, 0], [3, 4], [4]]
    ptr = np.array([0, 3, 5, 7, 9, 10], np.int32)
    flat = np.array(sum(adj, []), np.int32)
    out, n = np.empty(5, np.int32), C.c_int()
    oracle.lib().orc_bfs_order(5, ptr, flat, out, n)
    assert n.value == 5 and out.tolist() == [4, 3, 1, 2, 0]        # visit order 0,2,1,3,4 (no degree sort), reversed


@pytest.mark.parametrize("poly", [3, 4, 6])
def test_weights_reproduce_polynomials(poly):
    """grid.cpp:282-297,404-417: Laplacian / derivative weights are exact on monomials of degree <= polyDeg; interpolation
    rows sum to one."""
    x, y = jittered_square(24, seed=3)
    mg = oracle.Multigrid()
    mg.add_level(oracle.KIND_DIRICHLET, x, y, poly, fine=True)
    lv = mg.level(0)
    px, py = lv.points()
    n = oracle.stencil_size(poly)
    for i in range(5, lv.n, 61):
        w, nb = lv.weights(0, i)
        wx, _ = lv.weights(1, i)
        for a in range(poly + 1):
            for q in range(a + 1):
                f = (px[nb] - px[i]) ** (a - q) * (py[nb] - py[i]) ** q
                lap = 2.0 if (a - q, q) in ((2, 0), (0, 2)) else 0.0
                ddx = 1.0 if (a - q, q) == (1, 0) else 0.0
                assert abs(w[:n] @ f - lap) < 2e-6
                assert abs(wx[:n] @ f - ddx) < 1e-7
        wi, _ = lv.interp_weights(px[i] + 0.003, py[i] - 0.002, poly)
        assert abs(wi[:n].sum() - 1) < 1e-10


def test_dirichlet_vcycle_converges_to_the_manufactured_solution():
    mg = oracle.make_hierarchy([13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4)
    n, _ = mg.solve(1e-9, 200)
    assert n < 60
    lv = mg.level(-1)
    x, y = lv.points()
    err = np.abs(lv.values[: lv.n] - np.sin(np.pi * x) * np.sin(np.pi * y)).mean()        # calc_l1_error, testing_functions.cpp:3-16
    assert err < 1e-5


def test_two_level_quirk_zeroes_fine_dirichlet_values():
    # multigrid.cpp:91 applies boundaryOp("coarse") to grid 1; with 2 levels that is the finest grid
    mg = oracle.make_hierarchy([13, 25], kind=oracle.KIND_DIRICHLET, fine_poly=3)
    lv = mg.level(1)
    v = np.ones(lv.A)
    lv.set_vec(oracle.VEC_VALUES, v)
    mg.vcycle(1)
    assert np.all(mg.level(1).values[lv.bcflags() == 1] == 0)


def test_level_sizes_and_msh_roundtrip(tmp_path):
    assert level_sizes(100, 4) == [13, 25, 50, 100]
    x, y = jittered_square(9, seed=1)
    assert ((x == 0) | (x == 1) | (y == 0) | (y == 1)).sum() == 4 * 9 - 4
    fn = tmp_path / "c.msh"
    write_msh_nodes(str(fn), x, y)
    rows = [l.split() for l in open(fn).read().split("$Nodes\n")[1].split("\n")[1:-2]]
    assert np.array_equal(np.array([float(r[1]) for r in rows]), x)     # %.17g round-trips fp64


# ---------------------------------------------------------------- the C-ABI library: loads, exports, fails loudly
def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmg.h")).read()
    return sorted(set(re.findall(r"\b(mmg_[a-z0-9_]+)\s*\(", src)))


def test_libmmg_exports_every_declared_symbol():
    from meshlessmultigridpoisson_b200 import build, capi

    build.build()
    L = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) > 60
    for n in names:
        assert hasattr(L, n), "libmmg.so does not export " + n
    bound = set(capi.SIGNATURES) | {"mmg_last_error", "mmg_build_info"}
    assert set(names) == bound, set(names) ^ bound                      # the ctypes mirror binds exactly the header


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/mmg.h must compile as C99 (no C++-isms, no torch / CUDA types), and a C
    program that names every declared entry point must link against libmmg.so."""
    import re
    import shutil
    import subprocess

    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "mmg.h")).read()
    names = sorted(set(re.findall(r"\b(mmg_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 60
    src = tmp_path / "abi.c"
    src.write_text('#include "mmg.h"\n#include <stdio.h>\nint main(void) {\n  const void* p[] = {' +
                   ", ".join("(const void*)%s" % n for n in names) + '};\n  printf("%d\\n", (int)(sizeof p / sizeof p[0]));\n  return 0;\n}\n')
    lib = os.path.join(root, "meshlessmultigridpoisson_b200")
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-Wno-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-L", lib, "-lmmg", "-Wl,-rpath," + lib, "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    assert int(out) == len(names)


def test_no_cpu_fallback_without_a_device():
    """Host logic only: without a GPU the product must refuse, never compute on the CPU."""
    from meshlessmultigridpoisson_b200 import capi

    L = capi.load()
    assert b"sm_100a" in L.mmg_build_info()
    n = C.c_int(-1)
    rc = L.mmg_device_count(n)
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    x, y = jittered_square(8, seed=0)
    with pytest.raises(capi.MmgError) as e:
        capi.Grid(x, y, [], dict(rbfExp=3, polyDeg=3, stencilSize=25, iters=5, omega=1.4), np.zeros(64))
    assert e.value.code == capi.ERR_CUDA


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "meshlessmultigridpoisson_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text and "mmg_oracle" not in text, f


def test_fracstep_operators_are_exact_on_polynomials_and_kovasznay_is_nearly_divergence_free():
    """Known answers for the fractional-step matrices (fractionalStepGrid.cpp:60-100): every row of derivXMat_ / derivYMat_ /
    uvLaplaceMat_ is an RBF-FD stencil with polynomial augmentation of degree polyDeg, so the assembled matrices differentiate
    polynomials of that degree exactly at every node; applied to the analytic Kovasznay field (FractionalStepSim.cpp:96-99)
    the discrete divergence is small."""
    import math
    import scipy.sparse as sp

    x, y = jittered_square(28, seed=11)
    mg = oracle.Multigrid(fracstep=True)
    mg.add_level(oracle.KIND_PPE, x, y, 4, fine=True, dt=2e-4, mu=0.025, rho=1.0)
    lv = mg.level(0)
    px, py = lv.points()
    mats = {}
    for name, which in (("dx", oracle.MAT_DX), ("dy", oracle.MAT_DY), ("lap", oracle.MAT_UVLAP)):
        shape, ptr, idx, val = lv.csr(which)
        assert shape == (lv.n, lv.n)
        mats[name] = sp.csr_matrix((val, idx, ptr), shape=shape)
    f = 1 + 2 * px - 3 * py + px * py + 0.5 * px ** 2 - py ** 2 + px ** 3 - 2 * px * py ** 2 + 0.25 * px ** 2 * py ** 2
    fx = 2 + py + px + 3 * px ** 2 - 2 * py ** 2 + 0.5 * px * py ** 2
    fy = -3 + px - 2 * py - 4 * px * py + 0.5 * px ** 2 * py
    fl = (1 + 6 * px + 0.5 * py ** 2) + (-2 - 4 * px + 0.5 * px ** 2)
    assert np.abs(mats["dx"] @ f - fx).max() < 1e-7
    assert np.abs(mats["dy"] @ f - fy).max() < 1e-7
    assert np.abs(mats["lap"] @ f - fl).max() < 1e-4
    re = 1.0 / 0.025
    lam = 0.5 * re - math.sqrt(0.25 * re * re + 4 * math.pi ** 2)
    u = 1 - np.exp(lam * px) * np.cos(2 * math.pi * py)
    v = lam / (2 * math.pi) * np.exp(lam * px) * np.sin(2 * math.pi * py)
    div = mats["dx"] @ u + mats["dy"] @ v
    assert np.abs(div).max() < 5e-2 * np.abs(mats["dx"] @ u).max()


def test_error_convention_of_the_c_abi():
    """Every entry point returns an int status and records a message (the reference has no error convention at all: void
    returns and unchecked std::vector::at).  Argument validation happens before the device is touched,
    so these checks run without a GPU."""
    import ctypes as C

    from meshlessmultigridpoisson_b200 import capi

    L = capi.load()
    assert L.mmg_grid_destroy(None) == 0 and L.mmg_solver_destroy(None) == 0          # destroying nothing is fine, like delete nullptr
    for fn, args, what in ((L.mmg_solver_vcycle, (None, 1), b"null argument s"), (L.mmg_grid_sor, (None, 0), b"null argument g"),
                           (L.mmg_solver_create, (None, 0), b"null argument out"), (L.mmg_grid_fs_calc_hat, (None, -1), b"null argument g")):
        assert fn(*args) == 1                                                         # MMG_ERR_ARG
        assert what in L.mmg_last_error()
    b = np.zeros(4, np.int32)
    assert L.mmg_partition_bounds(10, 3, b) == 0 and b.tolist() == [0, 4, 7, 10]      # sizes differ by at most one, first ranks larger
    assert L.mmg_partition_bounds(10, 0, b) == 1 and b"bad sizes" in L.mmg_last_error()
    assert L.mmg_partition_bounds(2, 3, b) == 0 and b.tolist() == [0, 1, 2, 2]        # more ranks than rows: empty trailing blocks
    info = C.c_char_p(L.mmg_build_info()).value if L.mmg_build_info.restype is not C.c_char_p else L.mmg_build_info()
    assert b"sm_100a" in info


def test_missing_nccl_is_an_error_code_not_a_crash():
    """libnccl is resolved with dlopen at first use; when it cannot be loaded the call must return MMG_ERR_NCCL with a message
    (the error path used to read dlerror() twice and build a std::string from NULL)."""
    import subprocess, sys, textwrap

    code = textwrap.dedent("""
        import ctypes, sys
        sys.path.insert(0, %r)
        from meshlessmultigridpoisson_b200 import capi
        L = capi.load()
        buf = ctypes.create_string_buffer(128)
        rc = L.mmg_comm_unique_id(buf)
        print(rc, L.mmg_last_error().decode())
    """ % ROOT)
    env = dict(os.environ, MMG_NCCL_LIB="/nonexistent/libnccl.so.2")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1000:]
    rc, msg = out.stdout.strip().split(" ", 1)
    assert int(rc) == 4 and "cannot load libnccl" in msg          # MMG_ERR_NCCL
