"""The C++ facade (reference class names over the C-ABI) driven by a run_mg_sim-style program."""
import os
import subprocess

import numpy as np
import pytest

from meshlessmultigridpoisson_b200.clouds import jittered_square, write_msh_nodes
from meshlessmultigridpoisson_b200.problems import make_hierarchy

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_driver_matches_python_driver_bit_for_bit(libmmg, tmp_path):
    cpp = os.path.join(ROOT, "meshlessmultigridpoisson_b200", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp])
    sizes, files = [13, 25, 50], []
    for l, s in enumerate(sizes):
        x, y = jittered_square(s, seed=1000 + l)
        fn = str(tmp_path / ("l%d.msh" % l))
        write_msh_nodes(fn, x, y)
        files.append(fn)
    out = subprocess.check_output([os.path.join(cpp, "run_mg_sim"), "12", "4", *files], text=True, env=dict(os.environ, MMG_OUT_DIR=str(tmp_path), MMG_PRINT_INTERP="1")).split("\n")
    hist = np.array([float(t) for t in out[:12]])
    err = float(out[12].split()[1])
    mg = make_hierarchy(sizes, "dirichlet", 4)
    mg.vCycle(12)
    assert np.array_equal(hist, mg.residuals_)          # same library, same device-built operators: identical
    # restrictionMatrices_ / prolongMatrices_ fetched through the facade: shapes of multigrid.cpp:17-48, n = 37 entries per row
    # (the finest grid's polyDeg, multigrid.cpp:22,25), interpolation weights summing to one
    interp = {(t[1], int(t[2])): (int(t[3]), int(t[4]), int(t[5]), float(t[6])) for t in (l.split() for l in out if l.startswith("interp "))}
    n2 = [s_ * s_ for s_ in sizes]
    assert set(interp) == {("R", 1), ("R", 2), ("P", 0), ("P", 1)}
    for l in (1, 2):
        assert interp[("R", l)][:3] == (n2[l - 1], n2[l], 37 * n2[l - 1]) and interp[("R", l)][3] < 1e-9
        assert interp[("P", l - 1)][:3] == (n2[l], n2[l - 1], 37 * n2[l]) and interp[("P", l - 1)][3] < 1e-9
    assert hist[-1] < 2e-3 * hist[0] and err < 1e-3
    # the reference's text writers (write_mg_resid / write_temp_contour): one value per line, 6 significant digits
    resid = np.loadtxt(tmp_path / "resid_3grid__L=4.txt")
    assert resid.shape == hist.shape and np.allclose(resid, hist, rtol=1e-5)
    temp = np.loadtxt(tmp_path / "temp_3grid__L=4.txt")
    xs = np.loadtxt(tmp_path / "x_3grid__L=4.txt")
    assert xs.size == 2500 and temp.size == 2499           # values_->rows() - 1, like the reference
    assert np.allclose(temp, mg.grid(-1).values_[:2499], rtol=1e-5, atol=1e-12)


def test_cpp_fracstep_driver_matches_python_driver_bit_for_bit(libmmg, tmp_path):
    """run_fracstep.cpp keeps the statements of run_fracstep_param (FractionalStepSim.cpp:114-148) over the facade's
    FractionalStepGrid / FractionalStepMultigrid; the Python driver makes the same C-ABI calls."""
    from meshlessmultigridpoisson_b200 import capi
    from meshlessmultigridpoisson_b200.problems import fracstep_time_step, make_ppe_grid

    cpp = os.path.join(ROOT, "meshlessmultigridpoisson_b200", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp])
    sizes, files, clouds = [13, 25, 50], [], []
    for l, s in enumerate(sizes):
        x, y = jittered_square(s, seed=1000 + l)
        fn = str(tmp_path / ("l%d.msh" % l))
        write_msh_nodes(fn, x, y)
        files.append(fn)
        clouds.append((x, y))
    out = subprocess.check_output([os.path.join(cpp, "run_fracstep"), "3", "1e-6", "3", *files], text=True, env=dict(os.environ, MMG_FS_MAX_CYCLES="200")).split("\n")
    deltas = [float(t.split()[0]) for t in out[:3]]
    cycles = [int(t.split()[1]) for t in out[:3]]
    mg = capi.FractionalStepMultigrid()
    for l, (x, y) in enumerate(clouds):
        mg.addGrid(make_ppe_grid(x, y, 3, 2e-4, 0.025, 1.0, fine=(l == len(sizes) - 1)))
    mg.buildMatrices()
    old = 1000.0
    for k in range(3):
        n, r = fracstep_time_step(mg, 1e-6, 200)
        assert n == cycles[k] and abs(r - old) == deltas[k]
        old = r
    assert 0 < cycles[0] < 200 and float(out[3].split()[1]) < 0.5


def test_ownership_rules_of_add_grid(libmmg):
    """Multigrid owns its grids (multigrid.cpp:10-16): a grid cannot be added twice, and the C-ABI refuses to destroy a grid a
    solver owns (double free); wrappers returned by Multigrid.grid() keep their solver alive."""
    import ctypes as C
    import numpy as np
    from meshlessmultigridpoisson_b200 import capi
    from meshlessmultigridpoisson_b200.clouds import jittered_square
    from meshlessmultigridpoisson_b200.problems import make_grid, make_hierarchy

    x, y = jittered_square(13, 1)
    g = make_grid("dirichlet", x, y, 3)
    mg = capi.Multigrid()
    mg.addGrid(g)
    with pytest.raises(capi.MmgError) as e:
        mg.addGrid(g)
    assert e.value.code == capi.ERR_STATE
    assert mg.L.mmg_grid_destroy(g.h) == capi.ERR_STATE
    w = make_hierarchy([13, 25], "dirichlet", 3).grid(-1)      # the solver object is unreachable except through the wrapper
    import gc; gc.collect()
    assert np.isfinite(w.values_).all() and w.getSize() == 625


def test_per_point_queries_of_the_facade_match_the_ctypes_mirror(libmmg, tmp_path):
    """Grid::kNearestNeighbors / laplaceWeights / derivx_weights / derivy_weights / pointInterpWeights / pointIDs_to_vector / diags
    of the C++ facade (grid.h:60-72) against the same C-ABI entries called through capi.py on an identically built level."""
    from meshlessmultigridpoisson_b200 import capi
    from meshlessmultigridpoisson_b200.problems import make_grid, stencil_size

    cpp = os.path.join(ROOT, "meshlessmultigridpoisson_b200", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp])
    poly, s = 4, 24
    x, y = jittered_square(s, seed=7)
    fn = str(tmp_path / "level.msh")
    write_msh_nodes(fn, x, y)
    ids = [0, 5, 100, 333, s * s - 1]                      # boundary and interior nodes (numbering after the reordering)
    out = subprocess.check_output([os.path.join(cpp, "query_stencil"), str(poly), fn, *map(str, ids)], text=True).strip().split("\n")
    rec = {}
    for line in out:
        t = line.split()
        rec[(t[0], int(t[1]))] = t[2:]
    g = make_grid("dirichlet", x, y, poly)
    n = stencil_size(poly)
    px, py = g.points_
    flags = g.bcFlags_
    for i in ids:
        knn = g.kNearestNeighbors(px[i], py[i], n, neumann=False, q_bcflag=[int(flags[i] != 0)])[0]
        assert [int(v) for v in rec[("knn", i)]] == knn.tolist()
        for tag, which in (("laplace", capi.MAT_LAPLACE), ("derivx", capi.MAT_DERIVX), ("derivy", capi.MAT_DERIVY)):
            w, nb = g.weights(which, [i], n)
            got = [t.split(":") for t in rec[(tag, i)]]
            assert [int(a) for a, _ in got] == nb[0].tolist()
            assert np.array_equal(np.array([float(b) for _, b in got]), w[0])          # 17 digits round-trip: identical
        j = int(knn[1])
        w, nb = g.pointInterpWeights(0.5 * (px[i] + px[j]), 0.5 * (py[i] + py[j]), poly)
        got = [t.split(":") for t in rec[("interp", i)]]
        assert [int(a) for a, _ in got] == nb[0].tolist()
        assert np.array_equal(np.array([float(b) for _, b in got]), w[0])
        assert abs(sum(float(b) for _, b in got) - 1.0) < 1e-9                          # interpolation weights sum to one
        assert float(rec[("diag", i)][0]) == g.diags[i]
