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


def hex_square(s, seed=0, jitter=0.0, margin=0.3):
    """Gmsh-frontal-like cloud on the unit square: s equispaced nodes per edge lying *exactly* on x,y in {0,1} (listed first:
    bottom, top, left, right), hexagonally packed interior nodes at the same spacing h = 1/(s-1), none closer than
    ``margin*h`` to a vertical edge.  About 1.15 s^2 nodes.  The reference's author meshed with Gmsh (files never committed,
    testing_functions.cpp:355-364); on clouds of this kind the reference's scheme converges with Neumann and mixed boundaries at
    sizes where it diverges on jittered lattices.  ``jitter`` displaces interior nodes by
    U(-jitter*h, jitter*h) with ``numpy.random.default_rng(seed)``."""
    h = 1.0 / (s - 1)
    t = np.arange(s, dtype=np.float64) / (s - 1)
    bx = np.concatenate([t, t, np.zeros(s - 2), np.ones(s - 2)])
    by = np.concatenate([np.zeros(s), np.ones(s), t[1:-1], t[1:-1]])
    ny = int(round(1.0 / (h * np.sqrt(3.0) / 2.0)))
    dy = 1.0 / ny
    rows = []
    for j in range(1, ny):
        xs = np.arange(0.5 * h if j % 2 else 0.0, 1.0 + 1e-12, h)
        xs = xs[(xs > margin * h) & (xs < 1.0 - margin * h)]
        rows.append(np.stack([xs, np.full_like(xs, j * dy)], 1))
    P = np.concatenate(rows)
    if jitter:
        P = P + np.random.default_rng(seed).uniform(-jitter * h, jitter * h, P.shape)
    return np.ascontiguousarray(np.concatenate([bx, P[:, 0]])), np.ascontiguousarray(np.concatenate([by, P[:, 1]]))


def _circle(r, h):
    n = max(8, int(round(2.0 * np.pi * r / h)))
    t = 2.0 * np.pi * np.arange(n) / n
    return 0.5 + r * np.cos(t), 0.5 + r * np.sin(t)


def hole_square(s, seed, jitter=0.3, clearance=0.5):
    """geomtype "square_with_circle" (testing_functions.cpp:92-106): jittered lattice on the unit square minus the disc of radius
    0.25 about (0.5, 0.5), plus equispaced nodes on the circle.  The reference detects the circle by
    |0.0625 - (x-0.5)^2 - (y-0.5)^2| <= 1e-10; nodes closer than ``clearance*h`` to it are dropped."""
    h = 1.0 / (s - 1)
    x, y = jittered_square(s, seed, jitter)
    r = np.hypot(x - 0.5, y - 0.5)
    keep = r > 0.25 + clearance * h
    cx, cy = _circle(0.25, h)
    return np.ascontiguousarray(np.concatenate([x[keep], cx])), np.ascontiguousarray(np.concatenate([y[keep], cy]))


def annulus(s, seed, jitter=0.3, clearance=0.5):
    """geomtype "concentric_circles" (testing_functions.cpp:107-135): nodes on the circles of radius 0.5 and 0.25 about (0.5, 0.5)
    and a jittered lattice between them."""
    h = 1.0 / (s - 1)
    idx = np.arange(s, dtype=np.float64) / (s - 1)
    x, y = np.meshgrid(idx, idx, indexing="xy")
    rng = np.random.default_rng(seed)
    x = (x + rng.uniform(-jitter * h, jitter * h, size=(s, s))).ravel()
    y = (y + rng.uniform(-jitter * h, jitter * h, size=(s, s))).ravel()
    r = np.hypot(x - 0.5, y - 0.5)
    keep = (r > 0.25 + clearance * h) & (r < 0.5 - clearance * h)
    ox, oy = _circle(0.5, h)
    ix, iy = _circle(0.25, h)
    return np.ascontiguousarray(np.concatenate([ox, ix, x[keep]])), np.ascontiguousarray(np.concatenate([oy, iy, y[keep]]))


def make_cloud(kind, s, seed, jitter=0.3):
    """'jittered': SURVEY.md section 8d lattice; 'hex': the Gmsh-like cloud"""
    if kind == "hex":
        return hex_square(s, seed)
    if kind == "square_with_circle":
        return hole_square(s, seed, jitter)
    if kind == "concentric_circles":
        return annulus(s, seed, jitter)
    return jittered_square(s, seed=seed, jitter=jitter)


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
