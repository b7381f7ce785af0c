"""ctypes binding of oracle/libplo_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.dirname(_HERE)
LIB_PATH = os.path.join(ORACLE_DIR, "libplo_oracle.so")

DROP_NAMES = ["no_normal", "too_far", "invalid_normal", "normal_constraint", "mls_fail", "nan_inf_height"]
REG_STATUS = {1: "CONVERGED", 2: "MAX_ITERS", 3: "TOO_FEW_PAIRS", 4: "SOLVE_FAILED"}


class OrcParams(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("h", C.c_double), ("r", C.c_double), ("r_normal", C.c_double),
        ("is_get_normals", C.c_int32), ("search_number_normal", C.c_int32), ("search_number", C.c_int32),
        ("normal_angle_constraint", C.c_int32), ("angle_diff_threshold", C.c_double),
        ("transform_normal", C.c_int32), ("correspond_number", C.c_int32),
        ("delta_dist_threshold", C.c_double), ("delta_angle_threshold", C.c_double),
        ("solver", C.c_int32), ("weight_mode", C.c_int32),
        ("ransac_distance_threshold", C.c_double), ("huber_threshold", C.c_double),
        ("ls_threshold", C.c_double), ("ransac_max_iterations", C.c_int32),
        ("ransac_min_inliers_percentage", C.c_double), ("ransac_final", C.c_int32),
        ("drpm_threshold", C.c_double), ("drpm_stdev_points", C.c_double), ("drpm_stdev_normals", C.c_double),
        ("ransac_seed", C.c_uint64),
    ]


class OrcFrontendParams(C.Structure):
    _fields_ = [("n_scans", C.c_int32), ("min_range", C.c_float), ("max_range", C.c_float), ("scan_period", C.c_float),
                ("window_size", C.c_int32), ("iter_step", C.c_int32), ("knn_distance_threshold", C.c_float),
                ("plane_distance_threshold", C.c_float), ("valid_points_threshold", C.c_float), ("use_all_points", C.c_int32),
                ("planarity_threshold", C.c_float)]


class OrcRegStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("pairs", C.c_int64), ("rms", C.c_double),
                ("counters", C.c_int64 * 6)]


def build(force: bool = False) -> str:
    deps = [os.path.join(ORACLE_DIR, f) for f in ("plo_oracle.c", "plo_oracle_frontend.c", "plo_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "libplo_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.orc_create.restype = vp
    L.orc_destroy.argtypes = [vp]
    L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
    L.orc_set_params.argtypes = [vp, C.POINTER(OrcParams)]
    L.orc_set_threads.argtypes = [vp, C.c_int]
    L.orc_get_threads.argtypes = [vp]
    L.orc_get_threads.restype = C.c_int
    for f in (L.orc_set_target, L.orc_set_source):
        f.argtypes = [vp, vp, i64, i32]
        f.restype = i64
    for f in (L.orc_target_size, L.orc_source_size):
        f.argtypes = [vp]
        f.restype = i64
    L.orc_get_target_normals.argtypes = [vp, vp]
    for f in (L.orc_knn, L.orc_knn_brute):
        f.argtypes = [vp, vp, C.c_int, dbl, C.c_int, vp, vp]
        f.restype = C.c_int
    L.orc_compute_normal.argtypes = [vp, C.c_int, vp]
    L.orc_project.argtypes = [vp] * 13
    L.orc_project.restype = i64
    L.orc_solve_wls.argtypes = [vp, vp, vp, vp, i64, vp]
    L.orc_solve_wls.restype = C.c_int
    L.orc_solve_ls.argtypes = [vp, vp, vp, i64, dbl, vp]
    L.orc_solve_ls.restype = C.c_int
    L.orc_ransac_weights.argtypes = [vp, vp, vp, i64, vp, dbl, dbl, vp, vp]
    L.orc_ransac_weights.restype = i64
    L.orc_solve_drpm.argtypes = [vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp]
    L.orc_solve_drpm.restype = C.c_int
    L.orc_solve_ransac.argtypes = [vp, vp, vp, i64, C.POINTER(OrcParams), vp]
    L.orc_solve_ransac.restype = C.c_int
    L.orc_normal_equations.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp]
    L.orc_colpiv_qr_solve.argtypes = [vp, vp, i64, C.c_int, vp, vp]
    L.orc_colpiv_qr_solve.restype = C.c_int
    L.orc_angle_axis.argtypes = [vp, vp]
    L.orc_polar_uvt.argtypes = [vp, vp]
    L.orc_sym3_eigen.argtypes = [vp, vp, vp]
    L.orc_sym6_eigen.argtypes = [vp, vp, vp]
    L.orc_register.argtypes = [vp, vp, vp, C.POINTER(OrcRegStats), vp]
    L.orc_register.restype = C.c_int
    L.orc_last_build_seconds.argtypes = [vp]
    L.orc_last_build_seconds.restype = dbl
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**over) -> OrcParams:
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """Thin OO wrapper; mirrors the matcher/solver surface of the reference."""

    def __init__(self, params: OrcParams | None = None, threads: int = 0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.params = params or default_params()
        self.L.orc_set_params(self.h, C.byref(self.params))
        self.L.orc_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self) -> int:
        return self.L.orc_get_threads(self.h)

    def set_params(self, params: OrcParams):
        self.params = params
        self.L.orc_set_params(self.h, C.byref(params))

    def set_target(self, rec: np.ndarray) -> int:
        rec = np.ascontiguousarray(rec, dtype=np.float32)
        return self.L.orc_set_target(self.h, _p(rec), rec.shape[0], rec.strides[0])

    def set_source(self, rec: np.ndarray) -> int:
        rec = np.ascontiguousarray(rec, dtype=np.float32)
        return self.L.orc_set_source(self.h, _p(rec), rec.shape[0], rec.strides[0])

    @property
    def n_target(self) -> int:
        return self.L.orc_target_size(self.h)

    @property
    def n_source(self) -> int:
        return self.L.orc_source_size(self.h)

    @property
    def build_seconds(self) -> float:
        return self.L.orc_last_build_seconds(self.h)

    def target_normals(self) -> np.ndarray:
        out = np.empty((self.n_target, 3), np.float64)
        self.L.orc_get_target_normals(self.h, _p(out))
        return out

    def knn(self, q, k: int, r: float, allow_self: bool, brute: bool = False):
        q = _f64(q)
        idx = np.empty(k, np.int32)
        d2 = np.empty(k, np.float64)
        f = self.L.orc_knn_brute if brute else self.L.orc_knn
        cnt = f(self.h, _p(q), k, r, int(allow_self), _p(idx), _p(d2))
        return cnt, idx, d2

    def project(self, T, hooks: bool = False) -> dict:
        T = _f64(T).reshape(16)
        m = self.n_source
        k = self.params.search_number
        sx = np.empty((m, 3), np.float32)
        rx = np.empty((m, 3), np.float32)
        rn = np.empty((m, 3), np.float32)
        si = np.empty(m, np.int32)
        cnt = np.zeros(6, np.int64)
        out = {}
        if hooks:
            st = np.empty(m, np.int32)
            hgt = np.empty(m, np.float64)
            i1 = np.empty(m, np.int32)
            d1 = np.empty(m, np.float64)
            ni = np.empty((m, k), np.int32)
            nd = np.empty((m, k), np.float64)
            n = self.L.orc_project(self.h, _p(T), _p(sx), _p(rx), _p(rn), _p(si), _p(cnt),
                                   _p(st), _p(hgt), _p(i1), _p(d1), _p(ni), _p(nd))
            out.update(status=st, height=hgt, nn1_idx=i1, nn1_d2=d1, nn_idx=ni, nn_d2=nd)
        else:
            n = self.L.orc_project(self.h, _p(T), _p(sx), _p(rx), _p(rn), _p(si), _p(cnt),
                                   None, None, None, None, None, None)
        out.update(n=int(n), src_xyz=sx[:n], ref_xyz=rx[:n], ref_n=rn[:n], src_idx=si[:n], counters=cnt)
        return out

    def register(self, T0=None):
        T0 = _f64(np.eye(4) if T0 is None else T0).reshape(16)
        T = np.empty(16, np.float64)
        st = OrcRegStats()
        per = np.zeros(max(1, self.params.iterations), np.int64)
        self.L.orc_register(self.h, _p(T0), _p(T), C.byref(st), _p(per))
        return T.reshape(4, 4), dict(status=st.status, iters=st.iters, pairs=st.pairs, rms=st.rms,
                                     counters=np.array(list(st.counters), np.int64), per_iter_pairs=per)


def transform_to_end(rec: np.ndarray, T, transform_normal: bool = False) -> np.ndarray:
    """TransformToEnd (src/laser_odometry.cpp:88-114) on a copy of float32 records (n x 12)."""
    out = np.ascontiguousarray(rec, dtype=np.float32).copy()
    T = _f64(T).reshape(16)
    L = lib()
    L.orc_transform_to_end.restype = None
    L.orc_transform_to_end.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int]
    L.orc_transform_to_end(_p(out), out.shape[0], out.strides[0], _p(T), int(bool(transform_normal)))
    return out


def frontend_default_params(**over) -> OrcFrontendParams:
    p = OrcFrontendParams()
    L = lib()
    L.orc_frontend_default_params.argtypes = [C.POINTER(OrcFrontendParams)]
    L.orc_frontend_default_params.restype = None
    L.orc_frontend_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def frontend(points: np.ndarray, params: OrcFrontendParams | None = None) -> dict:
    """laserCloudHandler's normal + presample stage (src/scan_registration.cpp) on (n, >=3) float32 points."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = pts.shape[0]
    p = params or frontend_default_params()
    rec = np.zeros((max(n, 1), 12), np.float32)
    ev = np.zeros((max(n, 1), 3), np.float32)
    cand = np.zeros(max(n, 1), np.uint8)
    src = np.zeros(max(n, 1), np.int32)
    st = np.zeros(4, np.int64)
    L = lib()
    L.orc_frontend.restype = C.c_int64
    L.orc_frontend.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(OrcFrontendParams)] + [C.c_void_p] * 5
    m = L.orc_frontend(_p(pts), n, pts.strides[0] if n else 12, C.byref(p), _p(rec), _p(ev), _p(cand), _p(src), _p(st))
    return dict(n=int(m), records=rec[:m], eigenvalues=ev[:m], candidate=cand[:m].astype(bool), src_index=src[:m],
                ringed=int(st[0]), pca_failures=int(st[1]), plane_failures=int(st[2]), candidates=int(st[3]))


def solve_wls(src, ref, nrm, w=None):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    w = None if w is None else _f64(w)
    d = np.empty(16)
    lib().orc_solve_wls(_p(src), _p(ref), _p(nrm), _p(w), src.shape[0], _p(d))
    return d.reshape(4, 4)


def solve_ls(src, ref, nrm, threshold=0.02):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    d = np.empty(16)
    lib().orc_solve_ls(_p(src), _p(ref), _p(nrm), src.shape[0], threshold, _p(d))
    return d.reshape(4, 4)


def ransac_weights(src, ref, nrm, Tbest=None, distance_threshold=0.8, huber_threshold=0.648):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    Tb = _f64(np.eye(4) if Tbest is None else Tbest).reshape(16)
    n = src.shape[0]
    idx = np.empty(n, np.int32)
    w = np.empty(n, np.float64)
    cnt = lib().orc_ransac_weights(_p(src), _p(ref), _p(nrm), n, _p(Tb), distance_threshold, huber_threshold, _p(idx), _p(w))
    return idx[:cnt], w[:cnt]


def solve_drpm(src, ref, nrm, w=None, threshold=0.05, stdev_points=0.02, stdev_normals=0.05):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    w = None if w is None else _f64(w)
    d = np.empty(16)
    pr = np.empty(6)
    lib().orc_solve_drpm(_p(src), _p(ref), _p(nrm), _p(w), src.shape[0], threshold, stdev_points, stdev_normals, _p(d), _p(pr))
    return d.reshape(4, 4), pr


def solve_ransac(src, ref, nrm, params: OrcParams | None = None):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    p = params or default_params()
    d = np.empty(16)
    ok = lib().orc_solve_ransac(_p(src), _p(ref), _p(nrm), src.shape[0], C.byref(p), _p(d))
    return bool(ok), d.reshape(4, 4)


def normal_equations(src, ref, nrm, w=None):
    src, ref, nrm = _f64(src), _f64(ref), _f64(nrm)
    w = None if w is None else _f64(w)
    H = np.empty(21)
    g = np.empty(6)
    sw = C.c_double()
    sbb = C.c_double()
    lib().orc_normal_equations(_p(src), _p(ref), _p(nrm), _p(w), src.shape[0], _p(H), _p(g), C.byref(sw), C.byref(sbb))
    return H, g, sw.value, sbb.value


def colpiv_qr_solve(A, b):
    A = np.array(A, dtype=np.float64, order="C", copy=True)
    b = np.array(b, dtype=np.float64, copy=True)
    x = np.empty(A.shape[1])
    rank = C.c_int()
    lib().orc_colpiv_qr_solve(_p(A), _p(b), A.shape[0], A.shape[1], _p(x), C.byref(rank))
    return x, rank.value


def angle_axis(rot):
    rot = _f64(rot)
    R = np.empty(9)
    lib().orc_angle_axis(_p(rot), _p(R))
    return R.reshape(3, 3)


def polar_uvt(R):
    R = _f64(R).reshape(9)
    out = np.empty(9)
    lib().orc_polar_uvt(_p(R), _p(out))
    return out.reshape(3, 3)


def sym_eigen(A):
    A = _f64(A)
    n = A.shape[0]
    ev = np.empty(n)
    V = np.empty(n * n)
    (lib().orc_sym3_eigen if n == 3 else lib().orc_sym6_eigen)(_p(A.reshape(-1)), _p(ev), _p(V))
    return ev, V.reshape(n, n)


def compute_normal(pts):
    pts = _f64(pts)
    out = np.empty(3)
    lib().orc_compute_normal(_p(pts), pts.shape[0], _p(out))
    return out
