"""`Context`: one plo_ctx (one GPU, one stream) with numpy-friendly wrappers.

Inputs may be numpy arrays (host records, copied by the library) or torch CUDA tensors
(device-resident records, passed by pointer — PyTorch is only the memory/stream plumbing).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PloError, PloParams, PloProjStats, PloRegStats


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _records(x):
    """-> (pointer, n, stride_bytes, on_device, keepalive)"""
    if _is_torch_cuda(x):
        if x.dim() != 2 or x.dtype.itemsize != 4 or x.stride(1) != 1:
            raise ValueError("device records must be a 2-D float32 tensor with contiguous rows")
        return C.c_void_p(x.data_ptr()), int(x.shape[0]), int(x.stride(0) * 4) if x.shape[0] > 1 else int(x.shape[1] * 4), True, x
    if hasattr(x, "is_cuda"):   # torch CPU tensor (possibly pinned): pass its host pointer
        if x.dim() != 2 or x.dtype.itemsize != 4 or x.stride(1) != 1:
            raise ValueError("host records must be a 2-D float32 tensor with contiguous rows")
        return C.c_void_p(x.data_ptr()), int(x.shape[0]), int(x.stride(0) * 4) if x.shape[0] > 1 else int(x.shape[1] * 4), False, x
    a = np.asarray(x)
    if a.ndim != 2 or a.dtype != np.float32:
        raise ValueError("records must be a 2-D float32 array (n, >=7): xyz at floats 0..2, normal at floats 4..6")
    if a.strides[1] != 4:
        a = np.ascontiguousarray(a)
    stride = a.strides[0] if a.shape[0] > 1 else a.shape[1] * 4
    return _ptr(a), int(a.shape[0]), int(stride), False, a


class Context:
    def __init__(self, device: int = 0, params: PloParams | None = None, stream=None):
        self.L = _lib.lib()
        h = C.c_void_p()
        rc = self.L.plo_create(device, C.byref(h))
        if rc != _lib.PLO_OK:
            raise PloError(rc, self.L.plo_last_error(None).decode())
        self.h = h
        self.device = device
        self.params = params or _lib.default_params()
        self.set_params(self.params)
        self._keep = []
        if stream is not None:
            self.set_stream(stream)

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.plo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int):
        if rc != _lib.PLO_OK:
            raise PloError(rc, self.L.plo_last_error(self.h).decode())

    def set_stream(self, stream):
        """stream: torch.cuda.Stream or a raw cudaStream_t integer"""
        raw = getattr(stream, "cuda_stream", stream)
        self._ck(self.L.plo_set_stream(self.h, C.c_void_p(int(raw))))
        self._stream_keep = stream

    def _after_producer(self, x):
        """A torch CUDA tensor is read by the library on the context's own stream: order that stream after the torch
        stream that is current now (where the tensor was, or is being, produced) -- no global synchronisation."""
        if not _is_torch_cuda(x):
            return
        import torch
        cur = torch.cuda.current_stream(x.device)
        mine = getattr(getattr(self, "_stream_keep", None), "cuda_stream", getattr(self, "_stream_keep", None))
        if mine is not None and int(mine) == int(cur.cuda_stream):
            return      # same stream: already ordered
        ev = torch.cuda.Event()
        ev.record(cur)
        self._ck(self.L.plo_stream_wait_event(self.h, C.c_void_p(int(ev.cuda_event))))
        self._ev_keep = ev

    def set_tuning(self, name: str, value: int):
        """tuning / test knobs of plo_set_tuning: chunk, no_graph, force_warm"""
        self._ck(self.L.plo_set_tuning(self.h, name.encode(), int(value)))

    def synchronize(self):
        self._ck(self.L.plo_synchronize(self.h))

    # -- parameters / clouds ------------------------------------------------------------
    def set_params(self, params: PloParams):
        self._ck(self.L.plo_set_params(self.h, C.byref(params)))
        self.params = params

    def set_target(self, rec):
        p, n, stride, dev, keep = _records(rec)
        self._keep_t = keep
        self._after_producer(rec)
        self._ck((self.L.plo_set_target_device if dev else self.L.plo_set_target)(self.h, p, n, stride))

    def set_source(self, rec):
        p, n, stride, dev, keep = _records(rec)
        self._keep_s = keep
        self._after_producer(rec)
        self._ck((self.L.plo_set_source_device if dev else self.L.plo_set_source)(self.h, p, n, stride))

    # -- front-end (plo_frontend*) ----------------------------------------------------------
    def frontend(self, points, params=None, fetch: bool = True):
        """laserCloudHandler's normal + presample stage (src/scan_registration.cpp) on raw points: numpy (n, >=3)
        float32, or a CUDA tensor.  The filtered cloud stays on the device (frontend_device_records); with `fetch`
        it is also returned: records (n_out, 12), eigenvalues, candidate flags, src_index."""
        fp = params or _lib.frontend_default_params()
        st = _lib.PloFrontendStats()
        if _is_torch_cuda(points):
            t = points.contiguous()
            self._keep_fe = t
            self._after_producer(t)
            n, stride = int(t.shape[0]), int(t.stride(0) * t.element_size()) if t.shape[0] else 12
            self._ck(self.L.plo_frontend_device(self.h, C.c_void_p(t.data_ptr()), n, stride, C.byref(fp), C.byref(st)))
        else:
            a = np.ascontiguousarray(points, dtype=np.float32)
            n, stride = a.shape[0], (a.strides[0] if a.shape[0] else 12)
            self._ck(self.L.plo_frontend(self.h, _ptr(a), n, stride, C.byref(fp), C.byref(st)))
        out = dict(n=int(st.n_out), gated=int(st.gated), ringed=int(st.ringed), pca_failures=int(st.pca_failures),
                   plane_failures=int(st.plane_failures), candidates=int(st.candidates))
        if fetch:
            m = max(out["n"], 1)
            rec = np.zeros((m, 12), np.float32)
            ev = np.zeros((m, 3), np.float32)
            cand = np.zeros(m, np.uint8)
            src = np.zeros(m, np.int32)
            self._ck(self.L.plo_frontend_get(self.h, _ptr(rec), _ptr(ev), _ptr(cand), _ptr(src), m))
            k = out["n"]
            out.update(records=rec[:k], eigenvalues=ev[:k], candidate=cand[:k].astype(bool), src_index=src[:k])
        return out

    def frontend_device_records(self):
        """(device pointer, n) of the 48-byte records of the last frontend() run"""
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.L.plo_frontend_device_records(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_source_from_frontend(self):
        p, n = self.frontend_device_records()
        self._ck(self.L.plo_set_source_device(self.h, C.c_void_p(p), n, 48))

    def set_target_from_frontend(self):
        p, n = self.frontend_device_records()
        self._ck(self.L.plo_set_target_device(self.h, C.c_void_p(p), n, 48))

    def map_push_from_frontend(self, T_last_curr=None, from_last_register: bool = False, max_queue: int = 1,
                               transform_normals: bool = False):
        p, n = self.frontend_device_records()
        T = None if T_last_curr is None else np.ascontiguousarray(T_last_curr, dtype=np.float64).reshape(16)
        self._ck(self.L.plo_map_push_device(self.h, C.c_void_p(p), n, 48, _ptr(T), 1 if from_last_register else 0,
                                            int(max_queue), 1 if transform_normals else 0))

    # -- device-resident local map (plo_map_*) -------------------------------------------
    def map_reset(self):
        self._ck(self.L.plo_map_reset(self.h))

    def map_push(self, rec, T_last_curr=None, from_last_register: bool = False, max_queue: int = 1,
                 transform_normals: bool = False):
        """accumulateTargetCloud with TransformToEnd (src/laser_odometry.cpp:116-136, :88-114): the queued frames
        move into the new frame's coordinates, the new frame is appended, the index is rebuilt on the device."""
        p, n, stride, dev, keep = _records(rec)
        self._keep_t = keep
        self._after_producer(rec)
        T = None if T_last_curr is None else np.ascontiguousarray(T_last_curr, dtype=np.float64).reshape(16)
        fn = self.L.plo_map_push_device if dev else self.L.plo_map_push
        self._ck(fn(self.h, p, n, stride, _ptr(T), 1 if from_last_register else 0, int(max_queue), 1 if transform_normals else 0))

    def map_info(self):
        fr, pts = C.c_int32(), C.c_int64()
        self._ck(self.L.plo_map_info(self.h, C.byref(fr), C.byref(pts)))
        return fr.value, pts.value

    def map_records(self) -> np.ndarray:
        """(n, 8) float32: x y z 0 nx ny nz 0, oldest frame first, in the newest frame's coordinates"""
        _, n = self.map_info()
        out = np.zeros((n, 8), np.float32)
        if n:
            self._ck(self.L.plo_map_get(self.h, _ptr(out), n))
        return out

    @property
    def n_target(self) -> int:
        return int(self.L.plo_target_size(self.h))

    @property
    def n_source(self) -> int:
        return int(self.L.plo_source_size(self.h))

    # -- matcher ------------------------------------------------------------------------
    def project(self, T=None, hooks: bool = False, stats: bool = True):
        T = np.ascontiguousarray(np.eye(4) if T is None else T, dtype=np.float64).reshape(16)
        st = PloProjStats()
        self._ck(self.L.plo_project(self.h, _ptr(T), 1 if hooks else 0, C.byref(st) if stats else None))
        if not stats:
            return None
        return dict(n_source=int(st.n_source), n_pairs=int(st.n_pairs), counters=np.array(list(st.dropped), np.int64))

    def pairs(self):
        m = max(self.n_source, 1)
        sx = np.empty((m, 3), np.float32)
        rx = np.empty((m, 3), np.float32)
        rn = np.empty((m, 3), np.float32)
        si = np.empty(m, np.int32)
        n = C.c_int64()
        self._ck(self.L.plo_get_pairs(self.h, _ptr(sx), _ptr(rx), _ptr(rn), _ptr(si), m, C.byref(n)))
        n = n.value
        return dict(n=n, src_xyz=sx[:n], ref_xyz=rx[:n], ref_n=rn[:n], src_idx=si[:n])

    def neighbors(self):
        m, k = self.n_source, self.params.search_number
        ni = np.empty((m, k), np.int32)
        nd = np.empty((m, k), np.float64)
        i1 = np.empty(m, np.int32)
        d1 = np.empty(m, np.float64)
        self._ck(self.L.plo_get_neighbors(self.h, _ptr(ni), _ptr(nd), _ptr(i1), _ptr(d1)))
        return dict(nn_idx=ni, nn_d2=nd, nn1_idx=i1, nn1_d2=d1)

    def search_stats(self) -> np.ndarray:
        """(M, 3) int32: leaves scanned, internal nodes expanded, list insertions per query"""
        out = np.empty((self.n_source, 3), np.int32)
        self._ck(self.L.plo_get_search_stats(self.h, _ptr(out)))
        return out

    def query_results(self, heights: bool = True):
        m = self.n_source
        st = np.empty(m, np.int32)
        hg = np.empty(m, np.float64) if heights else None
        self._ck(self.L.plo_get_query_results(self.h, _ptr(st), _ptr(hg)))
        return dict(status=st, height=hg)

    def target_normals(self) -> np.ndarray:
        out = np.empty((max(self.n_target, 0), 3), np.float64)
        self._ck(self.L.plo_get_target_normals(self.h, _ptr(out)))
        return out

    # -- solver -------------------------------------------------------------------------
    def solve_wls(self):
        d = np.empty(16, np.float64)
        rank = C.c_int32()
        self._ck(self.L.plo_solve_wls(self.h, _ptr(d), C.byref(rank)))
        return d.reshape(4, 4), rank.value

    def solve_ls(self):
        """SolveMotionEstimationProblemLS (trimmed) on the pairs of the last projection"""
        d = np.empty(16, np.float64)
        rank = C.c_int32()
        self._ck(self.L.plo_solve_ls(self.h, _ptr(d), C.byref(rank)))
        return d.reshape(4, 4), rank.value

    def solve_ransac(self):
        """SolveMotionEstimationProblemRANSAC (+ Weighted LS / DRPM tail) on the pairs of the last projection"""
        d = np.empty(16, np.float64)
        pr = np.zeros(6, np.float64)
        inl, hyp = C.c_int64(), C.c_int32()
        self._ck(self.L.plo_solve_ransac(self.h, _ptr(d), _ptr(pr), C.byref(inl), C.byref(hyp)))
        return d.reshape(4, 4), dict(probs=pr, inliers=int(inl.value), hypotheses=int(hyp.value))

    def solve_wls_host(self, src, ref, nrm, w=None):
        src, ref, nrm = (np.ascontiguousarray(a, np.float64) for a in (src, ref, nrm))
        w = None if w is None else np.ascontiguousarray(w, np.float64)
        d = np.empty(16, np.float64)
        rank = C.c_int32()
        self._ck(self.L.plo_solve_wls_host(self.h, _ptr(src), _ptr(ref), _ptr(nrm), _ptr(w), src.shape[0], _ptr(d), C.byref(rank)))
        return d.reshape(4, 4), rank.value

    def solve_ls_host(self, src, ref, nrm, threshold: float = 0.02):
        """SolveMotionEstimationProblemLS on caller pairs (n x 3 doubles each, float32-representable)"""
        src, ref, nrm = (np.ascontiguousarray(a, np.float64) for a in (src, ref, nrm))
        d = np.empty(16, np.float64)
        rank = C.c_int32()
        self._ck(self.L.plo_solve_ls_host(self.h, _ptr(src), _ptr(ref), _ptr(nrm), src.shape[0], float(threshold), _ptr(d), C.byref(rank)))
        return d.reshape(4, 4), rank.value

    def solve_ransac_host(self, src, ref, nrm, params: PloParams | None = None):
        """SolveMotionEstimationProblemRANSAC on caller pairs; `params` carries the RANSAC / final-solver fields"""
        src, ref, nrm = (np.ascontiguousarray(a, np.float64) for a in (src, ref, nrm))
        d = np.empty(16, np.float64)
        pr = np.zeros(6, np.float64)
        inl, hyp = C.c_int64(), C.c_int32()
        p = params or self.params
        self._ck(self.L.plo_solve_ransac_host(self.h, _ptr(src), _ptr(ref), _ptr(nrm), src.shape[0], C.byref(p), _ptr(d), _ptr(pr),
                                              C.byref(inl), C.byref(hyp)))
        return d.reshape(4, 4), dict(probs=pr, inliers=int(inl.value), hypotheses=int(hyp.value))

    def solve_drpm_host(self, src, ref, nrm, w=None, threshold: float = 0.05, stdev_points: float = 0.02, stdev_normals: float = 0.05):
        """SolveMotionEstimationProblemDRPM on caller pairs and weights"""
        src, ref, nrm = (np.ascontiguousarray(a, np.float64) for a in (src, ref, nrm))
        w = None if w is None else np.ascontiguousarray(w, np.float64)
        d = np.empty(16, np.float64)
        pr = np.zeros(6, np.float64)
        self._ck(self.L.plo_solve_drpm_host(self.h, _ptr(src), _ptr(ref), _ptr(nrm), _ptr(w), src.shape[0], float(threshold),
                                            float(stdev_points), float(stdev_normals), _ptr(d), _ptr(pr)))
        return d.reshape(4, 4), pr

    def imls_height(self, xyz_normal):
        """ImplicitMLSFunction for a batch of (already transformed) points: (n, 6) float32 -> heights, ok flags"""
        a = np.ascontiguousarray(xyz_normal, np.float32).reshape(-1, 6)
        h = np.empty(a.shape[0], np.float64)
        ok = np.empty(a.shape[0], np.int32)
        self._ck(self.L.plo_imls_height(self.h, _ptr(a), a.shape[0], _ptr(h), _ptr(ok)))
        return h, ok.astype(bool)

    def compute_normal(self, pts) -> np.ndarray:
        """ComputeNormal: unit eigenvector of the smallest eigenvalue of the points' population covariance"""
        a = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.empty(3, np.float64)
        self._ck(self.L.plo_compute_normal(self.h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def normal_equations(self):
        H = np.empty(21)
        g = np.empty(6)
        sw, sbb, cnt = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.L.plo_get_normal_equations(self.h, _ptr(H), _ptr(g), C.byref(sw), C.byref(sbb), C.byref(cnt)))
        return H, g, sw.value, sbb.value, cnt.value

    # -- resident loop ------------------------------------------------------------------
    def register(self, T0=None):
        T0a = None if T0 is None else np.ascontiguousarray(T0, np.float64).reshape(16)
        T = np.empty(16, np.float64)
        st = PloRegStats()
        self._ck(self.L.plo_register(self.h, _ptr(T0a), _ptr(T), C.byref(st)))
        return T.reshape(4, 4), _reg_stats(st)

    def register_batch(self, sources, targets):
        n = len(sources)
        assert n == len(targets)
        recs_s = [_records(s) for s in sources]
        recs_t = [_records(t) for t in targets]
        if n == 0:
            return np.zeros((0, 4, 4)), []
        dev = recs_s[0][3]
        stride = recs_s[0][2]
        if dev:
            self._after_producer(sources[0])
        for r in recs_s + recs_t:
            if r[3] != dev or (r[1] > 1 and r[2] != stride):
                raise ValueError("register_batch: all clouds must live on the same side and share one stride")
        ps = (C.c_void_p * n)(*[r[0] for r in recs_s])
        pt = (C.c_void_p * n)(*[r[0] for r in recs_t])
        ns = (C.c_int64 * n)(*[r[1] for r in recs_s])
        nt = (C.c_int64 * n)(*[r[1] for r in recs_t])
        T = np.empty((n, 16), np.float64)
        st = (PloRegStats * n)()
        self._ck(self.L.plo_register_batch(self.h, n, ps, ns, pt, nt, stride, 1 if dev else 0, _ptr(T), st))
        return T.reshape(n, 4, 4), [_reg_stats(s) for s in st]

    # -- introspection ------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self.L.plo_launch_count(self.h))

    def last_timings(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.L.plo_last_timings(self.h, C.byref(a), C.byref(b)))
        return dict(ms_index_build=a.value, ms_register=b.value)

    def set_profiling(self, enabled: bool):
        self._ck(self.L.plo_set_profiling(self.h, 1 if enabled else 0))

    def last_kernel_timings(self):
        ms, n = C.c_float(), C.c_int32()
        self._ck(self.L.plo_last_kernel_timings(self.h, C.byref(ms), C.byref(n)))
        return dict(ms_project_mean=ms.value, n_project=n.value)

    def last_project_times(self) -> np.ndarray:
        """device ms of every projection launch of the last register() (profiling mode), one per ICP iteration"""
        ms = np.zeros(64, np.float32)
        n = C.c_int32()
        self._ck(self.L.plo_last_project_times(self.h, _ptr(ms), 64, C.byref(n)))
        return ms[: n.value].copy()

    def last_tile_misses(self) -> np.ndarray:
        """per projection of the last register(): queries the settled kernel handed to the tree walk (-1: it did not run)"""
        out = np.zeros(32, np.int32)
        n = C.c_int32()
        self._ck(self.L.plo_last_tile_misses(self.h, _ptr(out), 32, C.byref(n)))
        return out[: n.value].copy()

    def time_project_kernel(self, T=None, reps: int = 10) -> float:
        T = np.ascontiguousarray(np.eye(4) if T is None else T, dtype=np.float64).reshape(16)
        ms = C.c_float()
        self._ck(self.L.plo_time_project_kernel(self.h, _ptr(T), reps, C.byref(ms)))
        return ms.value


def _reg_stats(st: PloRegStats) -> dict:
    return dict(status=int(st.status), status_name=_lib.REG_STATUS.get(int(st.status), "?"), iters=int(st.iters),
                pairs=int(st.pairs), rms=float(st.rms), counters=np.array(list(st.dropped), np.int64),
                delta_dist=float(st.delta_dist), delta_angle=float(st.delta_angle), rank=int(st.rank))
