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
    out = subprocess.check_output([os.path.join(cpp, "run_mg_sim"), "12", "4", *files], text=True).split("\n")
    hist = np.array([float(t) for t in out[:12]])
    err = float(out[12].split()[1])
    mg = make_hierarchy(sizes, "dirichlet", 4)
    mg.vCycle(12)
    assert np.array_equal(hist, mg.residuals_)          # same library, same device-built operators: identical
    assert hist[-1] < 2e-3 * hist[0] and err < 1e-3
