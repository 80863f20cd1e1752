"""ctypes binding of include/mmg.h plus thin Python mirrors of the reference classes.

``Grid``, ``Multigrid`` and ``FractionalStepMultigrid`` keep the reference's method names
(grid.h:20-79, multigrid.h:4-23, FracStepMultigrid.hpp:4-25) and forward one-to-one to the
C-ABI of libmmg.so.  Everything numerical happens in the CUDA library; this module moves host
numpy arrays across the boundary and raises ``MmgError`` on any non-zero status.  There is no
CPU path: if libmmg.so is missing, or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmg.so")

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NCCL, ERR_TIMEOUT = range(6)
BC_DIRICHLET, BC_NEUMANN = 1, 2
FINE, COARSE = 0, 1
LEXICOGRAPHIC, MULTICOLOUR, BLOCK_LEXICOGRAPHIC = 0, 1, 2
FLAVOUR_MULTIGRID, FLAVOUR_FRACSTEP = 0, 1
ARITH_REFERENCE_ORDER, ARITH_FAST = 0, 1
MAT_LAPLACE, MAT_NEUMANN_COEFFS, MAT_RESTRICT, MAT_PROLONG, MAT_DERIVX, MAT_DERIVY, MAT_UVLAPLACE = range(7)
T_SOR, T_RESIDUAL, T_RESTRICT, T_PROLONG, T_OTHER, T_COUNT = range(6)
FS_U, FS_V, FS_U_OLD, FS_V_OLD, FS_U_HAT, FS_V_HAT = range(6)
GEOMTYPES = ["square", "square_with_circle", "concentric_circles"]     # MMG_GEOM_*


class MmgProps(C.Structure):
    _fields_ = [("rbfExp", C.c_int), ("polyDeg", C.c_int), ("stencilSize", C.c_int), ("iters", C.c_int), ("omega", C.c_double)]


class MmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mmg error %d: %s" % (code, msg))
        self.code = code


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i = C.c_int
_d = C.c_double

# name -> argtypes (every function returns int status unless listed in _NON_STATUS)
SIGNATURES = {
    "mmg_device_count": [C.POINTER(_i)],
    "mmg_grid_create": [C.POINTER(_vp), _i, _i, _dp, _dp, C.POINTER(MmgProps), _dp, _i, _i, _vp, _vp, _vp, _vp],
    "mmg_grid_destroy": [_vp],
    "mmg_grid_set_implicit": [_vp, _i],
    "mmg_grid_set_bc_flag": [_vp, _i, _i, _vp, _i],
    "mmg_grid_build_normal_vecs_square": [_vp],
    "mmg_grid_build_normal_vecs": [_vp, _i],
    "mmg_grid_set_normal_vecs": [_vp, _dp, _dp],
    "mmg_grid_rcm_order_points": [_vp],
    "mmg_grid_build_deriv_normal_bound": [_vp],
    "mmg_grid_build_laplacian": [_vp],
    "mmg_grid_modify_coeff_neumann": [_vp, _i],
    "mmg_grid_push_inhomog_to_rhs": [_vp],
    "mmg_grid_boundary_op": [_vp, _i],
    "mmg_grid_bound_eval_neumann": [_vp],
    "mmg_grid_set_props": [_vp, C.POINTER(MmgProps)],
    "mmg_grid_set_arithmetic": [_vp, _i],
    "mmg_grid_sor": [_vp, _i],
    "mmg_grid_residual": [_vp, _dp],
    "mmg_grid_fix_vector_bound_coarse": [_vp, _dp],
    "mmg_grid_knn": [_vp, _i, _dp, _dp, _vp, _i, _i, _ip],
    "mmg_grid_weights": [_vp, _i, _i, _ip, _dp, _ip],
    "mmg_grid_point_interp_weights": [_vp, _i, _dp, _dp, _i, _dp, _ip],
    "mmg_grid_sizes": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "mmg_grid_get_values": [_vp, _dp],
    "mmg_grid_set_values": [_vp, _dp],
    "mmg_grid_get_source": [_vp, _dp],
    "mmg_grid_set_source": [_vp, _dp],
    "mmg_grid_get_points": [_vp, _dp, _dp],
    "mmg_grid_get_bcflags": [_vp, _ip],
    "mmg_grid_get_normals": [_vp, _dp, _dp],
    "mmg_grid_get_diags": [_vp, _dp],
    "mmg_grid_get_perm": [_vp, _ip],
    "mmg_grid_get_boundary": [_vp, _i, C.POINTER(_i), C.POINTER(_i), _vp, _vp],
    "mmg_grid_csr_nnz": [_vp, _i, C.POINTER(C.c_int64)],
    "mmg_grid_get_csr": [_vp, _i, _ip, _ip, _dp],
    "mmg_grid_apply_matrix": [_vp, _i, _dp, _dp],
    "mmg_grid_set_laplacian_csr": [_vp, _i, _ip, _ip, _dp, _vp, _vp, _vp, _vp],
    "mmg_grid_get_colouring": [_vp, C.POINTER(_i), _ip],
    "mmg_grid_get_lex_levels": [_vp, C.POINTER(_i), _ip],
    "mmg_grid_get_colour_counts": [_vp, C.POINTER(_i), _vp, _i],
    "mmg_grid_set_block_size": [_vp, _i],
    "mmg_grid_get_block_colouring": [_vp, C.POINTER(_i), C.POINTER(_i), _vp, _i],
    "mmg_grid_fs_init": [_vp, _d, _d, _d],
    "mmg_grid_fs_build_operators": [_vp],
    "mmg_grid_fs_set_operator_csr": [_vp, _i, _i, _ip, _ip, _dp],
    "mmg_grid_fs_get_vec": [_vp, _i, _dp],
    "mmg_grid_fs_set_vec": [_vp, _i, _dp],
    "mmg_grid_fs_scatter": [_vp, _i, _i, _ip, _dp],
    "mmg_grid_fs_set_uv_bound": [_vp],
    "mmg_grid_fs_calc_hat": [_vp, C.c_int],
    "mmg_grid_fs_set_ppe_source": [_vp],
    "mmg_grid_fs_correct": [_vp, C.c_int],
    "mmg_grid_fs_residual": [_vp, C.POINTER(_d)],
    "mmg_solver_create": [C.POINTER(_vp), _i],
    "mmg_solver_destroy": [_vp],
    "mmg_solver_add_grid": [_vp, _vp],
    "mmg_solver_num_grids": [_vp, C.POINTER(_i)],
    "mmg_solver_grid": [_vp, _i, C.POINTER(_vp)],
    "mmg_solver_build_matrices": [_vp],
    "mmg_solver_set_interp_csr": [_vp, _i, _i, _i, _i, _ip, _ip, _dp],
    "mmg_solver_interp_nnz": [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_int64)],
    "mmg_solver_get_interp_csr": [_vp, _i, _i, _ip, _ip, _dp],
    "mmg_solver_finish_build": [_vp],
    "mmg_solver_set_smoother": [_vp, _i],
    "mmg_solver_restrict": [_vp, _i],
    "mmg_solver_prolong_correct": [_vp, _i],
    "mmg_solver_coarse_solve": [_vp],
    "mmg_solver_set_block_size": [_vp, _i],
    "mmg_solver_set_omega": [_vp, _d],
    "mmg_solver_set_arithmetic": [_vp, _i],
    "mmg_solver_vcycle": [_vp, _i],
    "mmg_solver_residual": [_vp, C.POINTER(_d)],
    "mmg_solver_history_len": [_vp, C.POINTER(_i)],
    "mmg_solver_get_history": [_vp, _dp, _i],
    "mmg_solver_solve": [_vp, _d, _i, _i, C.POINTER(_i), C.POINTER(_d)],
    "mmg_solver_sync": [_vp],
    "mmg_solver_enable_timers": [_vp, _i],
    "mmg_solver_get_timers": [_vp, _i, _dp, _lp, _lp],
    "mmg_solver_reset_timers": [_vp],
    "mmg_solver_launch_count": [_vp, C.POINTER(C.c_int64)],
    "mmg_solver_time_vcycles": [_vp, _i, C.POINTER(_d)],
    "mmg_partition_bounds": [_i, _i, _ip],
    "mmg_comm_unique_id": [C.c_char_p],
    "mmg_solver_init_comm": [_vp, _i, _i, C.c_char_p],
    "mmg_solver_set_partition_threshold": [_vp, _i],
    "mmg_solver_comm_stats": [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(_i)],
    "mmg_solver_gather_values": [_vp],
    "mmg_solver_owned_range": [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "mmg_grid_get_values_range": [_vp, _i, _i, _dp],
    "mmg_grid_set_values_range": [_vp, _i, _i, _dp],
    "mmg_grid_set_source_range": [_vp, _i, _i, _dp],
    "mmg_debug_last_kernel": [_i, C.c_char_p, _i],
    "mmg_debug_lex_trace": [_lp, _i],
    "mmg_debug_exchange_plan": [_i, _i, _ip, _ip, C.POINTER(_i), _ip, C.POINTER(_i), _ip],
}
_NON_STATUS = {"mmg_last_error": (C.c_char_p, []), "mmg_build_info": (C.c_char_p, [])}

_LIB = None


def load(path=None):
    """dlopen libmmg.so and attach prototypes.  Fails loudly when the CUDA library is absent."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise MmgError(ERR_CUDA, "libmmg.so not found at %s — build it with `python -m meshlessmultigridpoisson_b200.build` "
                                 "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    for name, args in SIGNATURES.items():
        f = getattr(L, name)
        f.restype, f.argtypes = C.c_int, args
    for name, (res, args) in _NON_STATUS.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    if path is None:
        _LIB = L
    return L


def _ck(L, rc):
    if rc != OK:
        raise MmgError(rc, L.mmg_last_error().decode())


def device_count():
    L = load()
    n = _i()
    _ck(L, L.mmg_device_count(n))
    return n.value


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _opt(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Boundary:
    """gridclasses.hpp:15-20"""

    def __init__(self, bcPoints, values, type=0):
        self.type, self.bcPoints, self.values = type, _i32(bcPoints), _f64(values)


class Grid:
    """Mirror of the reference's Grid over the C-ABI (grid.h:20-79)."""

    def __init__(self, points_x, points_y, boundaries, properties, source, device=0, _handle=None, _owned=True):
        self.L = load()
        self._owned = _owned
        if _handle is not None:
            self.h = _handle
            return
        x, y, src = _f64(points_x), _f64(points_y), _f64(source)
        props = MmgProps(properties["rbfExp"], properties["polyDeg"], properties["stencilSize"], properties["iters"], properties["omega"])
        types = _i32([b.type for b in boundaries])
        ptr = _i32(np.concatenate([[0], np.cumsum([b.bcPoints.size for b in boundaries])]))
        pts = _i32(np.concatenate([b.bcPoints for b in boundaries])) if boundaries else np.zeros(0, np.int32)
        vals = _f64(np.concatenate([b.values for b in boundaries])) if boundaries else np.zeros(0)
        h = _vp()
        _ck(self.L, self.L.mmg_grid_create(C.byref(h), device, x.size, x, y, C.byref(props), src, src.size, len(boundaries),
                                           _opt(types), _opt(ptr), _opt(pts), _opt(vals)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and self._owned:
            self.L.mmg_grid_destroy(self.h)
            self.h = None

    # ---- sizes / members
    def _sizes(self):
        n, a, f = _i(), _i(), _i()
        _ck(self.L, self.L.mmg_grid_sizes(self.h, n, a, f))
        return n.value, a.value, bool(f.value)

    laplaceMatSize_ = property(lambda s: s._sizes()[0])
    A_size = property(lambda s: s._sizes()[1])
    neumannFlag_ = property(lambda s: s._sizes()[2])

    def getSize(self):
        return self._sizes()[0]

    def _getv(self, fn):
        out = np.empty(self.A_size)
        _ck(self.L, fn(self.h, out))
        return out

    @property
    def values_(self):
        return self._getv(self.L.mmg_grid_get_values)

    def read_values(self, out):
        """values_ straight into a caller-owned buffer (e.g. pinned host memory): one D2H copy, no intermediate array."""
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.A_size
        _ck(self.L, self.L.mmg_grid_get_values(self.h, out))
        return out

    def read_values_range(self, offset, out):
        """values_[offset, offset + out.size) into a caller-owned (pinned) buffer"""
        _ck(self.L, self.L.mmg_grid_get_values_range(self.h, offset, out.size, out))
        return out

    def write_values_range(self, offset, v):
        _ck(self.L, self.L.mmg_grid_set_values_range(self.h, offset, v.size, v))

    def write_source_range(self, offset, v):
        _ck(self.L, self.L.mmg_grid_set_source_range(self.h, offset, v.size, v))

    @values_.setter
    def values_(self, v):
        v = _f64(v)
        assert v.size == self.A_size
        _ck(self.L, self.L.mmg_grid_set_values(self.h, v))

    @property
    def source_(self):
        return self._getv(self.L.mmg_grid_get_source)

    @source_.setter
    def source_(self, v):
        v = _f64(v)
        assert v.size == self.A_size
        _ck(self.L, self.L.mmg_grid_set_source(self.h, v))

    @property
    def diags(self):
        return self._getv(self.L.mmg_grid_get_diags)

    @property
    def points_(self):
        n = self.getSize()
        x, y = np.empty(n), np.empty(n)
        _ck(self.L, self.L.mmg_grid_get_points(self.h, x, y))
        return x, y

    @property
    def normalVecs_(self):
        n = self.getSize()
        x, y = np.empty(n), np.empty(n)
        _ck(self.L, self.L.mmg_grid_get_normals(self.h, x, y))
        return x, y

    @property
    def bcFlags_(self):
        f = np.empty(self.getSize(), np.int32)
        _ck(self.L, self.L.mmg_grid_get_bcflags(self.h, f))
        return f

    def perm(self):
        o = np.empty(self.getSize(), np.int32)
        _ck(self.L, self.L.mmg_grid_get_perm(self.h, o))
        return o

    def boundary(self, b):
        t, c = _i(), _i()
        _ck(self.L, self.L.mmg_grid_get_boundary(self.h, b, t, c, None, None))
        pts, vals = np.empty(c.value, np.int32), np.empty(c.value)
        _ck(self.L, self.L.mmg_grid_get_boundary(self.h, b, t, c, _opt(pts), _opt(vals)))
        return t.value, pts, vals

    def csr(self, which=MAT_LAPLACE):
        nnz = C.c_int64()
        _ck(self.L, self.L.mmg_grid_csr_nnz(self.h, which, nnz))
        rows = self.getSize() if which in (MAT_DERIVX, MAT_DERIVY, MAT_UVLAPLACE) else self.A_size
        ptr, idx, val = np.empty(rows + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value)
        _ck(self.L, self.L.mmg_grid_get_csr(self.h, which, ptr, idx, val))
        return (rows, rows), ptr, idx, val

    # ---- reference methods
    def setBCFlag(self, boundary, type_str, bound_values):
        t = BC_DIRICHLET if type_str == "dirichlet" else BC_NEUMANN      # grid.cpp:35
        v = _f64(bound_values)
        _ck(self.L, self.L.mmg_grid_set_bc_flag(self.h, boundary, t, _opt(v), v.size))

    def set_implicitFlag(self, flag):
        _ck(self.L, self.L.mmg_grid_set_implicit(self.h, int(flag)))

    def build_normal_vecs(self, geomtype="square"):
        _ck(self.L, self.L.mmg_grid_build_normal_vecs(self.h, GEOMTYPES.index(geomtype)))

    def rcm_order_points(self):
        _ck(self.L, self.L.mmg_grid_rcm_order_points(self.h))

    def build_deriv_normal_bound(self):
        _ck(self.L, self.L.mmg_grid_build_deriv_normal_bound(self.h))

    def build_laplacian(self):
        _ck(self.L, self.L.mmg_grid_build_laplacian(self.h))

    def modify_coeff_neumann(self, coarse):
        _ck(self.L, self.L.mmg_grid_modify_coeff_neumann(self.h, COARSE if coarse == "coarse" else FINE))

    def push_inhomog_to_rhs(self):
        _ck(self.L, self.L.mmg_grid_push_inhomog_to_rhs(self.h))

    def boundaryOp(self, coarse):
        _ck(self.L, self.L.mmg_grid_boundary_op(self.h, COARSE if coarse == "coarse" else FINE))

    def bound_eval_neumann(self):
        _ck(self.L, self.L.mmg_grid_bound_eval_neumann(self.h))

    def set_props(self, properties):
        p = MmgProps(properties["rbfExp"], properties["polyDeg"], properties["stencilSize"], properties["iters"], properties["omega"])
        _ck(self.L, self.L.mmg_grid_set_props(self.h, C.byref(p)))

    def set_arithmetic(self, arithmetic):
        _ck(self.L, self.L.mmg_grid_set_arithmetic(self.h, arithmetic))

    def sor(self, smoother=LEXICOGRAPHIC):
        _ck(self.L, self.L.mmg_grid_sor(self.h, smoother))

    def residual(self):
        return self._getv(self.L.mmg_grid_residual)

    def fix_vector_bound_coarse(self, vec):
        v = np.array(vec, np.float64)
        _ck(self.L, self.L.mmg_grid_fix_vector_bound_coarse(self.h, v))
        return v

    def kNearestNeighbors(self, qx, qy, k, neumann=False, q_bcflag=None):
        qx, qy = _f64(np.atleast_1d(qx)), _f64(np.atleast_1d(qy))
        fl = None if q_bcflag is None else _i32(np.atleast_1d(q_bcflag))
        out = np.empty((qx.size, k), np.int32)
        _ck(self.L, self.L.mmg_grid_knn(self.h, qx.size, qx, qy, _opt(fl), int(neumann), k, out.reshape(-1)))
        return out

    def weights(self, which, ids, stencil):
        ids = _i32(np.atleast_1d(ids))
        w, nb = np.empty((ids.size, stencil)), np.empty((ids.size, stencil), np.int32)
        _ck(self.L, self.L.mmg_grid_weights(self.h, which, ids.size, ids, w.reshape(-1), nb.reshape(-1)))
        return w, nb

    def laplaceWeights(self, ids, stencil):
        return self.weights(MAT_LAPLACE, ids, stencil)

    def pointInterpWeights(self, px, py, polyDeg):
        px, py = _f64(np.atleast_1d(px)), _f64(np.atleast_1d(py))
        n = int(2.5 * (polyDeg + 1) * (polyDeg + 2) / 2)
        w, nb = np.empty((px.size, n)), np.empty((px.size, n), np.int32)
        _ck(self.L, self.L.mmg_grid_point_interp_weights(self.h, px.size, px, py, polyDeg, w.reshape(-1), nb.reshape(-1)))
        return w, nb

    # ---- FractionalStepGrid (fractionalStepGrid.hpp:4-30)
    def fs_init(self, dt, mu, rho):
        _ck(self.L, self.L.mmg_grid_fs_init(self.h, dt, mu, rho))

    def fs_build_operators(self):
        """build_derivX_mat(); build_derivY_mat(); build_uv_laplace_mat()"""
        _ck(self.L, self.L.mmg_grid_fs_build_operators(self.h))

    def fs_set_operator_csr(self, which, ptr, idx, val):
        ptr, idx, val = _i32(ptr), _i32(idx), _f64(val)
        _ck(self.L, self.L.mmg_grid_fs_set_operator_csr(self.h, which, ptr.size - 1, ptr, idx, val))

    def fs_vec(self, which):
        out = np.empty(self.getSize())
        _ck(self.L, self.L.mmg_grid_fs_get_vec(self.h, which, out))
        return out

    def fs_set_vec(self, which, v):
        v = _f64(v)
        assert v.size == self.getSize()
        _ck(self.L, self.L.mmg_grid_fs_set_vec(self.h, which, v))

    def set_uv_bound(self):
        """fractionalStepGrid.cpp:41-59 (kovasznay): exact velocities on every boundary node into u, v, u_old, v_old."""
        _ck(self.L, self.L.mmg_grid_fs_set_uv_bound(self.h))

    def calc_hat(self):
        """calc_u_hat(); calc_v_hat()"""
        _ck(self.L, self.L.mmg_grid_fs_calc_hat(self.h, -1))

    def set_ppe_source(self):
        _ck(self.L, self.L.mmg_grid_fs_set_ppe_source(self.h))

    def correct_uv(self):
        """correct_u(); correct_v()"""
        _ck(self.L, self.L.mmg_grid_fs_correct(self.h, -1))

    def fs_residual(self):
        r = _d()
        _ck(self.L, self.L.mmg_grid_fs_residual(self.h, r))
        return r.value

    # ---- upload path + artefacts
    def set_laplacian_csr(self, ptr, idx, val, diags=None, nbc=None):
        ptr, idx, val = _i32(ptr), _i32(idx), _f64(val)
        d = None if diags is None else _f64(diags)
        if nbc is not None:
            np_, ni, nv = _i32(nbc[0]), _i32(nbc[1]), _f64(nbc[2])
        else:
            np_ = ni = nv = None
        _ck(self.L, self.L.mmg_grid_set_laplacian_csr(self.h, ptr.size - 1, ptr, idx, val, _opt(d), _opt(np_), _opt(ni), _opt(nv)))

    def colouring(self):
        n, c = _i(), np.empty(self.A_size, np.int32)
        _ck(self.L, self.L.mmg_grid_get_colouring(self.h, n, c))
        return n.value, c

    def colour_counts(self):
        n = _i()
        _ck(self.L, self.L.mmg_grid_get_colour_counts(self.h, n, None, 0))
        c = np.empty(n.value, np.int32)
        _ck(self.L, self.L.mmg_grid_get_colour_counts(self.h, n, _opt(c), c.size))
        return c

    def set_block_size(self, rows_per_block):
        _ck(self.L, self.L.mmg_grid_set_block_size(self.h, rows_per_block))

    def block_colouring(self):
        nb, nc = _i(), _i()
        _ck(self.L, self.L.mmg_grid_get_block_colouring(self.h, nb, nc, None, 0))
        c = np.empty(nb.value, np.int32)
        _ck(self.L, self.L.mmg_grid_get_block_colouring(self.h, nb, nc, _opt(c), c.size))
        return nc.value, c

    def lex_levels(self):
        n, c = _i(), np.empty(self.A_size, np.int32)
        _ck(self.L, self.L.mmg_grid_get_lex_levels(self.h, n, c))
        return n.value, c


class Multigrid:
    """Mirror of Multigrid (multigrid.h:4-23); ``flavour`` selects the FractionalStepMultigrid twin."""

    flavour = FLAVOUR_MULTIGRID

    def __init__(self):
        self.L = load()
        h = _vp()
        _ck(self.L, self.L.mmg_solver_create(C.byref(h), self.flavour))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.mmg_solver_destroy(self.h)
            self.h = None

    def addGrid(self, grid):
        _ck(self.L, self.L.mmg_solver_add_grid(self.h, grid.h))
        grid._owned = False          # Multigrid owns its grids (multigrid.cpp:10-16)

    @property
    def num_grids(self):
        n = _i()
        _ck(self.L, self.L.mmg_solver_num_grids(self.h, n))
        return n.value

    def grid(self, level):
        """grids_[level].second, 0 = coarsest; negative indexes from the finest."""
        if level < 0:
            level += self.num_grids
        g = _vp()
        _ck(self.L, self.L.mmg_solver_grid(self.h, level, C.byref(g)))
        w = Grid(None, None, None, None, None, _handle=g, _owned=False)
        w._parent = self             # the wrapper must not outlive the solver that owns the grid
        return w

    def buildMatrices(self):
        _ck(self.L, self.L.mmg_solver_build_matrices(self.h))

    def set_interp_csr(self, which, level, shape, ptr, idx, val):
        _ck(self.L, self.L.mmg_solver_set_interp_csr(self.h, which, level, shape[0], shape[1], _i32(ptr), _i32(idx), _f64(val)))

    def interp_csr(self, which, level):
        r, c, nnz = _i(), _i(), C.c_int64()
        _ck(self.L, self.L.mmg_solver_interp_nnz(self.h, which, level, r, c, nnz))
        ptr, idx, val = np.empty(r.value + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value)
        _ck(self.L, self.L.mmg_solver_get_interp_csr(self.h, which, level, ptr, idx, val))
        return (r.value, c.value), ptr, idx, val

    def finish_build(self):
        _ck(self.L, self.L.mmg_solver_finish_build(self.h))

    def set_smoother(self, smoother):
        _ck(self.L, self.L.mmg_solver_set_smoother(self.h, smoother))

    def set_block_size(self, rows_per_block):
        _ck(self.L, self.L.mmg_solver_set_block_size(self.h, rows_per_block))

    def set_omega(self, omega):
        _ck(self.L, self.L.mmg_solver_set_omega(self.h, omega))

    def set_arithmetic(self, arithmetic):
        _ck(self.L, self.L.mmg_solver_set_arithmetic(self.h, arithmetic))

    def restrict(self, level):
        _ck(self.L, self.L.mmg_solver_restrict(self.h, level))

    def prolong_correct(self, level):
        _ck(self.L, self.L.mmg_solver_prolong_correct(self.h, level))

    def coarse_solve(self):
        _ck(self.L, self.L.mmg_solver_coarse_solve(self.h))

    def vCycle(self, n=1):
        _ck(self.L, self.L.mmg_solver_vcycle(self.h, n))

    def residual(self):
        r = _d()
        _ck(self.L, self.L.mmg_solver_residual(self.h, r))
        return r.value

    @property
    def residuals_(self):
        n = _i()
        _ck(self.L, self.L.mmg_solver_history_len(self.h, n))
        out = np.empty(max(n.value, 1))
        _ck(self.L, self.L.mmg_solver_get_history(self.h, out, n.value))
        return out[: n.value]

    def solve(self, tol, max_cycles=1000, extra_bound_eval=False):
        n, r = _i(), _d()
        _ck(self.L, self.L.mmg_solver_solve(self.h, tol, max_cycles, int(extra_bound_eval), n, r))
        return n.value, r.value

    def sync(self):
        _ck(self.L, self.L.mmg_solver_sync(self.h))

    def enable_timers(self, on=True):
        _ck(self.L, self.L.mmg_solver_enable_timers(self.h, int(on)))

    def reset_timers(self):
        _ck(self.L, self.L.mmg_solver_reset_timers(self.h))

    def timers(self, level=-1):
        """CUDA-event time, launch count and algorithmic bytes per kernel class (level -1 = all levels)."""
        ms, ln, by = np.zeros(T_COUNT), np.zeros(T_COUNT, np.int64), np.zeros(T_COUNT, np.int64)
        _ck(self.L, self.L.mmg_solver_get_timers(self.h, level, ms, ln, by))
        names = ["sor", "residual", "restrict", "prolong", "other"]
        return {k: dict(ms=float(ms[i]), launches=int(ln[i]), bytes=int(by[i])) for i, k in enumerate(names)}

    def launch_count(self):
        n = C.c_int64()
        _ck(self.L, self.L.mmg_solver_launch_count(self.h, n))
        return n.value

    def init_comm(self, rank, world, unique_id):
        _ck(self.L, self.L.mmg_solver_init_comm(self.h, rank, world, unique_id))

    def set_partition_threshold(self, rows):
        _ck(self.L, self.L.mmg_solver_set_partition_threshold(self.h, rows))

    def owned_range(self, level=-1):
        """(own_lo, own_hi, need_lo, need_hi) of this rank on `level` (whole vector when the level is not partitioned)"""
        if level < 0:
            level += self.num_grids
        a, b, c, d = _i(), _i(), _i(), _i()
        _ck(self.L, self.L.mmg_solver_owned_range(self.h, level, a, b, c, d))
        return a.value, b.value, c.value, d.value

    def gather_values(self):
        """collective: complete values_ of every partitioned level on every rank (after a partitioned vCycle / solve)"""
        _ck(self.L, self.L.mmg_solver_gather_values(self.h))

    def comm_stats(self):
        m, b, p = C.c_int64(), C.c_int64(), _i()
        _ck(self.L, self.L.mmg_solver_comm_stats(self.h, m, b, p))
        return dict(messages=m.value, bytes_sent=b.value, partitioned_levels=p.value)

    def time_vcycles(self, n):
        ms = _d()
        _ck(self.L, self.L.mmg_solver_time_vcycles(self.h, n, ms))
        return ms.value


def last_kernel(slot=0):
    """Diagnostics: the kernel instantiation the last smoother (slot 0) or SpMV-class (slot 1) call launched."""
    L = load()
    buf = C.create_string_buffer(128)
    _ck(L, L.mmg_debug_last_kernel(slot, buf, 128))
    return buf.value.decode()


def partition_bounds(n, world):
    """rank r owns rows [bounds[r], bounds[r+1]) — contiguous blocks whose sizes differ by at most one (pure host logic)."""
    L = load()
    b = np.empty(world + 1, np.int32)
    _ck(L, L.mmg_partition_bounds(n, world, b))
    return b


def comm_unique_id():
    L = load()
    buf = C.create_string_buffer(128)
    _ck(L, L.mmg_comm_unique_id(buf))
    return buf.raw


class FractionalStepMultigrid(Multigrid):
    """FracStepMultigrid.hpp:4-25 — same kernels, the twin's two differences (interp polyDeg, 1-grid shortcut)."""

    flavour = FLAVOUR_FRACSTEP
