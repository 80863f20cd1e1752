"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the CPU oracle (oracle/liboracle.so).

The oracle restates the reference's Grid / Multigrid / FractionalStepMultigrid
(see oracle/mmg_oracle.hpp for the citation map).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this
module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_DIRICHLET, KIND_NEUMANN, KIND_PPE, KIND_MIXED = 0, 1, 2, 3
GEOM_SQUARE, GEOM_SQUARE_WITH_CIRCLE, GEOM_CONCENTRIC_CIRCLES = 0, 1, 2
GEOM_NAMES = ["square", "square_with_circle", "concentric_circles"]
MAT_A, MAT_NBC, MAT_R, MAT_P, MAT_DX, MAT_DY, MAT_UVLAP = 0, 1, 2, 3, 4, 5, 6
VEC_VALUES, VEC_SOURCE, VEC_DIAGS, VEC_U, VEC_V, VEC_UHAT, VEC_VHAT = 0, 1, 2, 3, 4, 5, 6

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mmg_oracle.cpp", "mmg_oracle_capi.cpp", "mmg_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, so])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    v, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
    sig = {
        "orc_last_error": (C.c_char_p, []),
        "orc_mg_new": (v, [i]),
        "orc_mg_free": (None, [v]),
        "orc_mg_add_level": (i, [v, i, i, _dp, _dp, i, i, d, i, i, i, i, i, d, d, d]),
        "orc_mg_add_level_raw": (i, [v, i, i, _dp, _dp, i, i, i, d, i, i, i, _ip, _ip, _ip, _dp, _dp, i]),
        "orc_mg_alloc_interp": (None, [v]),
        "orc_mg_build": (i, [v]),
        "orc_mg_nlevels": (i, [v]),
        "orc_mg_set_multicolour": (None, [v, i]),
        "orc_mg_set_smoother": (None, [v, i, i]),
        "orc_mg_set_omega": (None, [v, d]),
        "orc_lv_sor_blocklex": (None, [v, i, i]),
        "orc_lv_block_colouring": (i, [v, i, i, _ip, i]),
        "orc_mg_vcycle": (i, [v, i]),
        "orc_mg_time_vcycles": (d, [v, i]),
        "orc_mg_solve": (i, [v, d, i, C.POINTER(d)]),
        "orc_mg_residual": (d, [v]),
        "orc_mg_history": (i, [v, _dp, i]),
        "orc_lv_n": (i, [v, i]),
        "orc_lv_A": (i, [v, i]),
        "orc_lv_neumann": (i, [v, i]),
        "orc_lv_implicit": (i, [v, i]),
        "orc_lv_props": (None, [v, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(d)]),
        "orc_lv_get_points": (None, [v, i, _dp, _dp]),
        "orc_lv_get_normals": (None, [v, i, _dp, _dp]),
        "orc_lv_get_perm": (None, [v, i, _ip]),
        "orc_lv_get_bcflags": (None, [v, i, _ip]),
        "orc_lv_nboundaries": (i, [v, i]),
        "orc_lv_boundary_size": (i, [v, i, i]),
        "orc_lv_get_boundary": (i, [v, i, i, _ip, _dp]),
        "orc_lv_csr_nnz": (l, [v, i, i]),
        "orc_lv_csr_shape": (None, [v, i, i, C.POINTER(i), C.POINTER(i)]),
        "orc_lv_get_csr": (None, [v, i, i, _ip, _ip, _dp]),
        "orc_lv_set_csr": (None, [v, i, i, i, i, _ip, _ip, _dp]),
        "orc_lv_vec_size": (i, [v, i, i]),
        "orc_lv_get_vec": (None, [v, i, i, _dp]),
        "orc_lv_set_vec": (None, [v, i, i, _dp]),
        "orc_lv_sor": (None, [v, i]),
        "orc_lv_sor_multicolour": (None, [v, i]),
        "orc_lv_residual": (None, [v, i, _dp]),
        "orc_lv_bound_eval_neumann": (None, [v, i]),
        "orc_lv_boundary_op": (None, [v, i, i]),
        "orc_lv_modify_coeff_neumann": (None, [v, i, i]),
        "orc_lv_push_inhomog_to_rhs": (None, [v, i]),
        "orc_lv_fix_vector_bound_coarse": (None, [v, i, _dp]),
        "orc_lv_spmv": (None, [v, i, i, _dp, _dp]),
        "orc_lv_knn": (i, [v, i, d, d, i, i, i, i, _ip]),
        "orc_lv_weights": (i, [v, i, i, i, _dp, _ip]),
        "orc_lv_interp_weights": (i, [v, i, d, d, i, _dp, _ip]),
        "orc_lv_coeff_matrix": (i, [v, i, d, d, i, i, i, _dp, _ip, _dp]),
        "orc_lv_colouring": (i, [v, i, _ip]),
        "orc_lv_lex_levels": (None, [v, i, _ip]),
        "orc_fs_set_params": (None, [v, i, d, d, d]),
        "orc_fs_step_pre": (None, [v, i]),
        "orc_fs_step_post": (d, [v, i]),
        "orc_distance": (d, [d, d, d, d]),
        "orc_fullpivlu_solve": (None, [i, _dp, _dp, _dp]),
        "orc_csr_from_triplets": (l, [i, i, l, _ip, _ip, _dp, _ip, _ip, _dp]),
        "orc_bfs_order": (None, [i, _ip, _ip, _ip, C.POINTER(i)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


class OracleError(RuntimeError):
    pass


def stencil_size(poly_deg):
    """(int)(2.5*(p+1)*(p+2)/2) — grid.cpp:267, testing_functions.cpp:378."""
    return int(2.5 * (poly_deg + 1) * (poly_deg + 2) / 2)


class Level:
    """View of one oracle Grid (level index after the reference's sort-by-size)."""

    def __init__(self, mg, l):
        self.mg, self.l, self.L, self.h = mg, l, mg.L, mg.h

    @property
    def n(self):
        return self.L.orc_lv_n(self.h, self.l)

    @property
    def A(self):
        return self.L.orc_lv_A(self.h, self.l)

    @property
    def neumann(self):
        return bool(self.L.orc_lv_neumann(self.h, self.l))

    @property
    def implicit(self):
        return bool(self.L.orc_lv_implicit(self.h, self.l))

    @property
    def props(self):
        p, s, it, e, om = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self.L.orc_lv_props(self.h, self.l, p, s, it, e, om)
        return dict(polyDeg=p.value, stencilSize=s.value, iters=it.value, rbfExp=e.value, omega=om.value)

    def points(self):
        x, y = np.empty(self.n), np.empty(self.n)
        self.L.orc_lv_get_points(self.h, self.l, x, y)
        return x, y

    def normals(self):
        x, y = np.empty(self.n), np.empty(self.n)
        self.L.orc_lv_get_normals(self.h, self.l, x, y)
        return x, y

    def perm(self):
        o = np.empty(self.n, np.int32)
        self.L.orc_lv_get_perm(self.h, self.l, o)
        return o

    def bcflags(self):
        f = np.empty(self.n, np.int32)
        self.L.orc_lv_get_bcflags(self.h, self.l, f)
        return f

    def boundaries(self):
        out = []
        for b in range(self.L.orc_lv_nboundaries(self.h, self.l)):
            m = self.L.orc_lv_boundary_size(self.h, self.l, b)
            pts, vals = np.empty(m, np.int32), np.empty(m)
            t = self.L.orc_lv_get_boundary(self.h, self.l, b, pts, vals)
            out.append((t, pts, vals))
        return out

    def csr(self, which=MAT_A):
        r, c = C.c_int(), C.c_int()
        self.L.orc_lv_csr_shape(self.h, self.l, which, r, c)
        nnz = self.L.orc_lv_csr_nnz(self.h, self.l, which)
        ptr, idx, val = np.empty(r.value + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
        self.L.orc_lv_get_csr(self.h, self.l, which, ptr, idx, val)
        return (r.value, c.value), ptr, idx, val

    def set_csr(self, which, shape, ptr, idx, val):
        self.L.orc_lv_set_csr(self.h, self.l, which, shape[0], shape[1], np.ascontiguousarray(ptr, np.int32),
                              np.ascontiguousarray(idx, np.int32), np.ascontiguousarray(val, np.float64))

    def vec(self, which):
        v = np.empty(self.L.orc_lv_vec_size(self.h, self.l, which))
        self.L.orc_lv_get_vec(self.h, self.l, which, v)
        return v

    def set_vec(self, which, v):
        v = np.ascontiguousarray(v, np.float64)
        assert v.size == self.L.orc_lv_vec_size(self.h, self.l, which)
        self.L.orc_lv_set_vec(self.h, self.l, which, v)

    values = property(lambda s: s.vec(VEC_VALUES))
    source = property(lambda s: s.vec(VEC_SOURCE))
    diags = property(lambda s: s.vec(VEC_DIAGS))

    def sor(self):
        self.L.orc_lv_sor(self.h, self.l)

    def sor_multicolour(self):
        self.L.orc_lv_sor_multicolour(self.h, self.l)

    def sor_blocklex(self, block_size):
        self.L.orc_lv_sor_blocklex(self.h, self.l, block_size)

    def block_colouring(self, block_size):
        nb = (self.n + block_size - 1) // block_size
        c = np.empty(nb, np.int32)
        n = self.L.orc_lv_block_colouring(self.h, self.l, block_size, c, nb)
        return n, c

    def residual(self):
        r = np.empty(self.A)
        self.L.orc_lv_residual(self.h, self.l, r)
        return r

    def bound_eval_neumann(self):
        self.L.orc_lv_bound_eval_neumann(self.h, self.l)

    def boundary_op(self, coarse):
        self.L.orc_lv_boundary_op(self.h, self.l, int(coarse))

    def modify_coeff_neumann(self, coarse):
        self.L.orc_lv_modify_coeff_neumann(self.h, self.l, int(coarse))

    def push_inhomog_to_rhs(self):
        self.L.orc_lv_push_inhomog_to_rhs(self.h, self.l)

    def fix_vector_bound_coarse(self, v):
        v = np.array(v, np.float64)
        self.L.orc_lv_fix_vector_bound_coarse(self.h, self.l, v)
        return v

    def spmv(self, which, x):
        (r, c), *_ = self.csr(which)
        x = np.ascontiguousarray(x, np.float64)
        assert x.size >= c
        y = np.empty(r)
        self.L.orc_lv_spmv(self.h, self.l, which, x, y)
        return y

    def knn(self, x, y, k, neumann=False, point_bc=False, cells=False):
        out = np.empty(k, np.int32)
        if self.L.orc_lv_knn(self.h, self.l, x, y, int(neumann), int(point_bc), k, int(cells), out):
            raise OracleError(self.L.orc_last_error().decode())
        return out

    def weights(self, which, point_id):
        p = self.props
        m = (p["polyDeg"] + 1) * (p["polyDeg"] + 2) // 2
        w, nb = np.empty(p["stencilSize"] + m), np.empty(p["stencilSize"], np.int32)
        if self.L.orc_lv_weights(self.h, self.l, which, point_id, w, nb):
            raise OracleError(self.L.orc_last_error().decode())
        return w, nb

    def interp_weights(self, x, y, poly_deg):
        m = (poly_deg + 1) * (poly_deg + 2) // 2
        n = stencil_size(poly_deg)
        w, nb = np.empty(n + m), np.empty(n, np.int32)
        if self.L.orc_lv_interp_weights(self.h, self.l, x, y, poly_deg, w, nb):
            raise OracleError(self.L.orc_last_error().decode())
        return w, nb

    def coeff_matrix(self, x, y, poly_deg, neumann=False, point_bc=False):
        m = (poly_deg + 1) * (poly_deg + 2) // 2
        n = stencil_size(poly_deg)
        M, nb, sp = np.empty((n + m) * (n + m)), np.empty(n, np.int32), np.empty(2 * (n + 2))
        if self.L.orc_lv_coeff_matrix(self.h, self.l, x, y, int(neumann), int(point_bc), poly_deg, M, nb, sp):
            raise OracleError(self.L.orc_last_error().decode())
        return M.reshape(n + m, n + m).T.copy(), nb, sp.reshape(-1, 2)  # row-major view of the column-major data

    def colouring(self):
        c = np.empty(self.A, np.int32)
        n = self.L.orc_lv_colouring(self.h, self.l, c)
        return n, c

    def lex_levels(self):
        v = np.empty(self.A, np.int32)
        self.L.orc_lv_lex_levels(self.h, self.l, v)
        return v


class Multigrid:
    """Oracle Multigrid (fracstep=False, multigrid.cpp) or FractionalStepMultigrid (fracstep=True)."""

    def __init__(self, fracstep=False):
        self.L = lib()
        self.h = self.L.orc_mg_new(int(fracstep))
        self.fracstep = fracstep

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_mg_free(self.h)
            self.h = None

    def add_level(self, kind, x, y, poly_deg, iters=5, omega=1.4, rbf_exp=3, k1=1, k2=1, fine=False, cells=True,
                  dt=2e-4, mu=0.025, rho=1.0, geom=0):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        rc = self.L.orc_mg_add_level(self.h, kind, x.size, x, y, poly_deg, iters, omega, rbf_exp, k1, k2, int(fine), int(cells) | (int(geom) << 4), dt, mu, rho)
        if rc:
            raise OracleError(self.L.orc_last_error().decode())

    def add_level_raw(self, x, y, props, boundaries, source, implicit=False, fracstep_grid=False):
        """A level whose state was built elsewhere (reordered points, boundary lists as (type, points, values), source);
        operators follow through Level.set_csr.  Used to mirror a device-built hierarchy at sizes where the oracle's own
        set-up would take minutes."""
        x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
        src = np.ascontiguousarray(source, np.float64)
        bt = np.ascontiguousarray([b[0] for b in boundaries], np.int32)
        ptr = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(b[1]) for b in boundaries])]), np.int32)
        pts = np.ascontiguousarray(np.concatenate([b[1] for b in boundaries]) if boundaries else np.zeros(0), np.int32)
        vals = np.ascontiguousarray(np.concatenate([b[2] for b in boundaries]) if boundaries else np.zeros(0), np.float64)
        rc = self.L.orc_mg_add_level_raw(self.h, int(fracstep_grid), x.size, x, y, props["polyDeg"], props["stencilSize"], props["iters"],
                                         props["omega"], props["rbfExp"], int(implicit), len(boundaries), bt, ptr, pts, vals, src, src.size)
        if rc:
            raise OracleError(self.L.orc_last_error().decode())

    def alloc_interp(self):
        self.L.orc_mg_alloc_interp(self.h)

    def build(self):
        if self.L.orc_mg_build(self.h):
            raise OracleError(self.L.orc_last_error().decode())

    @property
    def nlevels(self):
        return self.L.orc_mg_nlevels(self.h)

    def level(self, l):
        if l < 0:
            l += self.nlevels
        return Level(self, l)

    def set_multicolour(self, on):
        self.L.orc_mg_set_multicolour(self.h, int(on))

    def set_smoother(self, smoother, block_size=4096):
        """0 lexicographic (the reference), 1 multicolour, 2 block-lexicographic"""
        self.L.orc_mg_set_smoother(self.h, smoother, block_size)

    def set_omega(self, omega):
        self.L.orc_mg_set_omega(self.h, omega)

    def vcycle(self, n=1):
        if self.L.orc_mg_vcycle(self.h, n):
            raise OracleError(self.L.orc_last_error().decode())

    def time_vcycles(self, n):
        return self.L.orc_mg_time_vcycles(self.h, n)

    def solve(self, tol, max_cycles=1000):
        s = C.c_double()
        n = self.L.orc_mg_solve(self.h, tol, max_cycles, C.byref(s))
        return n, s.value

    def residual(self):
        return self.L.orc_mg_residual(self.h)

    def history(self):
        n = self.L.orc_mg_history(self.h, np.empty(1), 0)
        out = np.empty(max(n, 1))
        self.L.orc_mg_history(self.h, out, n)
        return out[:n]


REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref.so")


class ReferenceHierarchy:
    """The reference's OWN Multigrid / Grid objects (oracle/_ref/libref.so: its grid.cpp and multigrid.cpp compiled where they lie
    on the Eigen-subset shim) over operators assembled by the oracle.  The reference's set-up is a brute-force kNN -- O(N^2) per
    level -- so beyond ~100k nodes only its V-cycle can be run as it is; `from_oracle` hands it the matrices of an oracle
    hierarchy (Dirichlet levels), which are pinned bit-identical to the ones the reference would assemble itself.
    TEST INFRASTRUCTURE: used by tests/ and by bench.py's reference arm only."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        self.R = R = C.CDLL(REF_SO)
        dp, ip = np.ctypeslib.ndpointer(np.float64), np.ctypeslib.ndpointer(np.int32)
        R.ref_new_raw.restype = C.c_void_p
        R.ref_free.argtypes = [C.c_void_p]
        R.ref_add_level_raw.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, ip, dp, dp, ip, ip, dp]
        R.ref_set_interp_raw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp]
        R.ref_time_vcycles.restype = C.c_double
        R.ref_time_vcycles.argtypes = [C.c_void_p, C.c_int]
        R.ref_vcycle.argtypes = [C.c_void_p, C.c_int]
        R.ref_residual.restype = C.c_double
        R.ref_residual.argtypes = [C.c_void_p]
        R.ref_history.argtypes = [C.c_void_p, dp, C.c_int]
        R.ref_lv_vec.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        R.ref_lv_A.argtypes = [C.c_void_p, C.c_int]
        self.h = C.c_void_p(R.ref_new_raw())
        self.nlevels = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.R.ref_free(self.h)
            self.h = None

    @classmethod
    def from_oracle(cls, mg):
        import scipy.sparse as sp

        self = cls()
        n_lv = mg.nlevels
        for l in range(n_lv):                                  # ascending size, like addGrid's sort (multigrid.cpp:116-122)
            lv = mg.level(l)
            if lv.neumann:
                raise OracleError("ReferenceHierarchy.from_oracle: Dirichlet levels only")
            bnd = lv.boundaries()
            assert len(bnd) == 1 and bnd[0][0] == 1, "one Dirichlet boundary per level (the square of genGmshGridDirichlet)"
            x, y = lv.points()
            pr = lv.props
            (rows, _), ptr, idx, val = lv.csr(MAT_A)
            assert rows == lv.n
            self.R.ref_add_level_raw(self.h, lv.n, x, y, pr["polyDeg"], pr["iters"], pr["omega"], pr["rbfExp"], bnd[0][1].size,
                                     np.ascontiguousarray(bnd[0][1], np.int32), np.ascontiguousarray(bnd[0][2], np.float64), lv.source, ptr, idx, val)
        self.nlevels = n_lv
        for l in range(n_lv):                                  # restrictionMatrices_[l]: l >= 1; prolongMatrices_[l]: l < last (multigrid.cpp:34-47)
            for which, sel, ok in ((0, MAT_R, l >= 1), (1, MAT_P, l < n_lv - 1)):
                if not ok:
                    continue
                shape, ptr, idx, val = mg.level(l).csr(sel)
                m = sp.csr_matrix((val, idx, ptr), shape=shape).tocsc()     # the reference stores them column-major (multigrid.h:8-9)
                m.sort_indices()
                self.R.ref_set_interp_raw(self.h, which, l, shape[0], shape[1], np.ascontiguousarray(m.indptr, np.int32),
                                          np.ascontiguousarray(m.indices, np.int32), np.ascontiguousarray(m.data, np.float64))
        return self

    def vcycle(self, n=1):
        self.R.ref_vcycle(self.h, n)

    def time_vcycles(self, n):
        """seconds (steady_clock) around n calls of the reference's Multigrid::vCycle, the loop of testing_functions.cpp:340-344"""
        return self.R.ref_time_vcycles(self.h, n)

    def residual(self):
        return self.R.ref_residual(self.h)

    def history(self):
        n = self.R.ref_history(self.h, np.empty(1), 0)
        out = np.empty(max(n, 1))
        self.R.ref_history(self.h, out, n)
        return out[:n]

    def values(self, l=-1):
        l = l + self.nlevels if l < 0 else l
        v = np.empty(self.R.ref_lv_A(self.h, l))
        self.R.ref_lv_vec(self.h, l, 0, v)
        return v


def make_hierarchy(sizes, kind=KIND_DIRICHLET, fine_poly=4, coarse_poly=3, fracstep=False, cells=True, seed0=1000, jitter=0.3, cloud="jittered", geom=0, **kw):
    """The reference's run_mg_sim set-up (testing_functions.cpp:328-339) on synthetic jittered
    lattices: one independent cloud per level, coarse levels polyDeg 3, finest fine_poly."""
    from meshlessmultigridpoisson_b200.clouds import make_cloud

    mg = Multigrid(fracstep=fracstep)
    for l, s in enumerate(sizes):
        x, y = make_cloud(cloud if geom == 0 else GEOM_NAMES[geom], s, seed0 + l, jitter)
        last = l == len(sizes) - 1
        mg.add_level(kind, x, y, fine_poly if last else coarse_poly, fine=last, cells=cells, geom=geom, **kw)
    mg.build()
    return mg
