"""Generates tests/golden/*.npz from the CPU oracle (run here, where the oracle has been pinned bit-exact against
the reference's own sources built on the Eigen shim — tests/test_oracle_vs_reference.py).  Commit the output.

    python tests/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
    "dirichlet_p4": dict(sizes=[13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=4, cycles=12),
    "dirichlet_p6": dict(sizes=[13, 25, 50], kind=oracle.KIND_DIRICHLET, fine_poly=6, cycles=12),
    "mixed_p4": dict(sizes=[13, 25, 50], kind=oracle.KIND_MIXED, fine_poly=4, cycles=12),
    "neumann_p3": dict(sizes=[13, 25, 50], kind=oracle.KIND_NEUMANN, fine_poly=3, cycles=12),
}


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def build(name):
    c = CASES[name]
    mg = oracle.make_hierarchy(c["sizes"], kind=c["kind"], fine_poly=c["fine_poly"], cells=False)
    out = {"sizes": np.array(c["sizes"]), "kind": c["kind"], "fine_poly": c["fine_poly"]}
    for l in range(mg.nlevels):
        lv = mg.level(l)
        _, ptr, idx, val = lv.csr()
        out["perm%d" % l] = lv.perm()
        out["csr_digest%d" % l] = digest(ptr, idx, val)
        out["csr_struct_digest%d" % l] = digest(ptr, idx)
        out["source%d" % l] = lv.source
        nc, col = lv.colouring()
        out["colour%d" % l] = col
        out["lex_levels%d" % l] = lv.lex_levels()
        x, y = lv.points()
        out["knn_probe%d" % l] = np.stack([lv.knn(x[i], y[i], lv.props["stencilSize"], neumann=lv.neumann, point_bc=bool(lv.bcflags()[i]))
                                           for i in range(0, lv.n, max(1, lv.n // 16))])
    fine = mg.level(-1)
    _, ptr, idx, val = fine.csr()
    out["fine_val_probe"] = val[:: max(1, val.size // 4096)]      # a strided sample of the finest operator's entries
    mg.vcycle(c["cycles"])
    out["history"] = mg.history()
    out["values"] = fine.values
    mg.set_multicolour(True)
    mg.vcycle(3)
    out["history_mc_tail"] = mg.history()[-3:]
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **build(name))
        print("wrote", name)
