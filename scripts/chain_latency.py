"""Micro-benchmark of the lexicographic DAG kernels: a banded matrix whose every row depends on the previous
row, so a sweep is one serial chain of `rows` hops and time/rows is the per-hop latency."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from meshlessmultigridpoisson_b200 import capi

def banded(n, half):
    W = 2 * half + 1
    rows = np.arange(n)[:, None]
    cols = rows + np.arange(-half, half + 1)[None, :]
    ok = (cols >= 0) & (cols < n)
    val = np.where(cols == rows, -float(W), 0.01)
    ptr = np.concatenate([[0], np.cumsum(ok.sum(1))]).astype(np.int32)
    return ptr, cols[ok].astype(np.int32), val[ok].astype(np.float64)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
half = int(sys.argv[2]) if len(sys.argv) > 2 else 18
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
x = np.linspace(0, 1, n); y = np.zeros(n)
props = dict(rbfExp=3, polyDeg=4, stencilSize=2 * half + 1, iters=iters, omega=1.4)
g = capi.Grid(x, y, [], props, np.ones(n))
ptr, idx, val = banded(n, half)
g.set_laplacian_csr(ptr, idx, val)
g.sor(capi.LEXICOGRAPHIC)
t0 = time.perf_counter(); g.sor(capi.LEXICOGRAPHIC); dt = time.perf_counter() - t0
hops = n + (iters - 1) * (half + 1) if not os.environ.get("MMG_LEX_NO_PIPE") else n * iters
print("rows %d half-band %d iters %d env %s: %.2f ms, %.0f ns per hop" % (n, half, iters, {k: v for k, v in os.environ.items() if k.startswith("MMG_")}, dt * 1e3, dt / hops * 1e9))
if os.environ.get("MMG_LEX_TRACE"):
    import ctypes
    buf = (ctypes.c_longlong * 128)()
    g.L.mmg_debug_lex_trace(buf, 128)
    t = np.array(buf[:128]).reshape(16, 8)
    base = t[:, 0].min()
    print("warp: start loaded round1 smem-only deps-done folded stored | rounds   (cycles since the chunk barrier)")
    for w in range(16):
        print("%2d: %s | %d" % (w, " ".join("%7d" % (v - base) for v in t[w, :7]), t[w, 7]))
