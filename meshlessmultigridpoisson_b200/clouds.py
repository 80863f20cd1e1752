"""Synthetic point clouds for the measurement harness.

The reference reads Gmsh ``$Nodes`` files that were never committed
(testing_functions.cpp:355-364); its levels are independent, non-nested clouds.
Here every level is an independent jittered lattice on the unit square whose
boundary nodes lie *exactly* on x,y in {0,1}, because the reference detects the
boundary by exact equality (testing_functions.cpp:86, FractionalStepSim.cpp:16).
"""
import numpy as np


def jittered_square(s, seed, jitter=0.3):
    """s x s lattice, h = 1/(s-1); interior nodes displaced by U(-jitter*h, jitter*h) in x and y
    with ``numpy.random.default_rng(seed)``; node order = lattice row-major (y outer, x inner)."""
    h = 1.0 / (s - 1)
    idx = np.arange(s, dtype=np.float64) / (s - 1)       # idx[-1] == 1.0 exactly
    x, y = np.meshgrid(idx, idx, indexing="xy")
    rng = np.random.default_rng(seed)
    dx = rng.uniform(-jitter * h, jitter * h, size=(s, s))
    dy = rng.uniform(-jitter * h, jitter * h, size=(s, s))
    interior = np.zeros((s, s), bool)
    interior[1:-1, 1:-1] = True
    x = np.where(interior, x + dx, x)
    y = np.where(interior, y + dy, y)
    return np.ascontiguousarray(x.ravel()), np.ascontiguousarray(y.ravel())


def level_sizes(s_fine, n_levels):
    """Lattice sides for a hierarchy with ~4x node coarsening per level (coarsest first)."""
    out = [s_fine]
    for _ in range(n_levels - 1):
        out.append((out[-1] + 1) // 2)
    return out[::-1]


def write_msh_nodes(path, x, y):
    """Gmsh v2 ASCII ``$Nodes`` block as consumed by pointsFromMshFile (fileReadingFunctions.cpp:6-32)."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % x.size)
        for i in range(x.size):
            f.write("%d %.17g %.17g 0\n" % (i + 1, x[i], y[i]))
        f.write("$EndNodes\n")
