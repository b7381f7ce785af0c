"""Independent numpy/scipy restatement of the IMLS-ICP hot path — TEST INFRASTRUCTURE.

Second, structurally different oracle used only to pin oracle/plo_oracle.c
("two independent restatements must agree: neighbour sets exactly,
heights/poses <= 1e-12").  Uses scipy.spatial.cKDTree for candidate generation with
exact re-ranking, numpy.linalg.lstsq / svd / eigh instead of the hand-written
QR / Jacobi routines of the C oracle.  Pure-Python loops: small inputs only.

Follows src/imls_icp.cpp:301-483 (ImplicitMLSFunction), :496-745
(ProjSourcePtToSurface), :753-794 (ComputeNormal), src/solver.cpp:168-220
(WeightedLS), src/laser_odometry.cpp:524-647 (driver loop).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

DBL_EPS = np.finfo(np.float64).eps


def strip_nonfinite(rec: np.ndarray) -> np.ndarray:
    """RemoveNANandINFData, src/imls_icp.cpp:58-72 (pcl::isFinite tests xyz only)."""
    return rec[np.isfinite(rec[:, 0:3]).all(axis=1)]


def d2_exact(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    d = q[None, :] - p
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


class Ref:
    def __init__(self, target: np.ndarray, source: np.ndarray, h=1.0, r=3.0, k=20,
                 angle_constraint=True, angle_thr=30.0, is_get_normals=True, r_normal=1.0,
                 k_normal=10, transform_normal=False):
        self.tgt = strip_nonfinite(np.asarray(target, np.float32))
        self.src = strip_nonfinite(np.asarray(source, np.float32))
        self.tp = self.tgt[:, 0:3].astype(np.float64)
        self.tree = cKDTree(self.tp) if self.tp.shape[0] else None
        self.h, self.r, self.k = h, r, k
        self.angle_constraint, self.angle_thr = angle_constraint, angle_thr
        self.transform_normal = transform_normal
        if is_get_normals:
            self.tn = self.tgt[:, 4:7].astype(np.float64)
        else:
            self.tn = np.stack([self._pca_normal(i, r_normal, k_normal) for i in range(self.tp.shape[0])]) \
                if self.tp.shape[0] else np.zeros((0, 3))

    # libnabo knn restated: d2<=r^2, (self or d2>eps), k best by (d2, idx), pad -1/inf
    def knn(self, q, k, r, allow_self):
        idx = np.full(k, -1, np.int32)
        d2 = np.full(k, np.inf)
        if self.tree is None or not np.isfinite(q).all():
            return idx, d2
        cand = np.asarray(self.tree.query_ball_point(q, r * (1 + 1e-9) + 1e-12), dtype=np.int64)
        if cand.size == 0:
            return idx, d2
        dd = d2_exact(q, self.tp[cand])
        ok = dd <= r * r
        if not allow_self:
            ok &= dd > DBL_EPS
        cand, dd = cand[ok], dd[ok]
        order = np.lexsort((cand, dd))[:k]
        idx[:order.size] = cand[order]
        d2[:order.size] = dd[order]
        return idx, d2

    def _pca_normal(self, i, r_normal, k_normal):
        idx, d2 = self.knn(self.tp[i], k_normal, r_normal, False)
        if (idx < 0).any():                       # D1
            return np.full(3, np.inf)
        pts = self.tp[idx]
        mu = pts.mean(axis=0)
        cov = (pts - mu).T @ (pts - mu) / pts.shape[0]
        w, V = np.linalg.eigh(cov)
        n = V[:, 0] / np.linalg.norm(V[:, 0])
        return -n if n[2] < 0 else n              # D2

    @staticmethod
    def angle_deg(a, b):
        with np.errstate(invalid="ignore", divide="ignore"):
            c = ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) / (np.sqrt(a @ a) * np.sqrt(b @ b))
            return np.degrees(np.arccos(c))

    def imls(self, x, xn):
        idx, d2 = self.knn(x, self.k, self.r, True)
        keep = []
        for j in range(self.k):
            if not np.isfinite(d2[j]):
                continue
            p, n = self.tp[idx[j]], self.tn[idx[j]]
            if not np.isfinite(p).all() or not np.isfinite(n).all():
                continue
            if self.angle_constraint and self.angle_deg(xn, n) > self.angle_thr:
                continue
            keep.append(j)
        if len(keep) < 3:
            return None, idx, d2
        with np.errstate(invalid="ignore", divide="ignore"):
            hmax = np.sqrt(d2[len(keep) - 1]) / 3
            ws = ps = 0.0
            for j in keep:
                d = x - self.tp[idx[j]]
                dn = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
                w = np.exp(-dn / hmax / hmax)
                n = self.tn[idx[j]]
                ws += w
                ps += ((w * d[0]) * n[0] + (w * d[1]) * n[1]) + (w * d[2]) * n[2]
            return ps / (ws + 1e-5), idx, d2

    def project(self, T):
        T = np.asarray(T, np.float64).reshape(4, 4)
        M = self.src.shape[0]
        status = np.zeros(M, np.int32)
        height = np.full(M, np.nan)
        nn_idx = np.full((M, self.k), -1, np.int32)
        nn_d2 = np.full((M, self.k), np.inf)
        nn1 = np.full(M, -1, np.int32)
        xs, ys, ns = [], [], []
        for i in range(M):
            p = self.src[i, 0:3].astype(np.float64)
            t = ((T[:3, 0] * p[0] + T[:3, 1] * p[1]) + T[:3, 2] * p[2]) + T[:3, 3]
            xf = t.astype(np.float32)
            nf = self.src[i, 4:7]
            if self.transform_normal:
                nd = nf.astype(np.float64)
                nf = ((T[:3, 0] * nd[0] + T[:3, 1] * nd[1]) + T[:3, 2] * nd[2]).astype(np.float32)
            x, xn = xf.astype(np.float64), nf.astype(np.float64)
            i1, d1 = self.knn(x, 1, self.r, False)
            nn1[i] = i1[0]
            st = 0
            if i1[0] < 0:
                st = 1
            elif d1[0] > self.h * self.h:
                st = 2
            else:
                n0 = self.tn[i1[0]]
                if not np.isfinite(n0).all():
                    st = 3
                elif self.angle_constraint and self.angle_deg(xn, n0) > self.angle_thr:
                    st = 4
            hgt, idx, d2 = (None, None, None)
            if st == 0:
                hgt, idx, d2 = self.imls(x, xn)
                if hgt is None:
                    st = 5
                elif not np.isfinite(hgt):
                    st = 6
            if idx is None:
                idx, d2 = self.knn(x, self.k, self.r, True)
            nn_idx[i], nn_d2[i] = idx, d2
            status[i] = st
            if st == 0:
                height[i] = hgt
                xs.append(xf)
                ys.append((x - hgt * n0).astype(np.float32))
                ns.append(n0.astype(np.float32))
        z = np.zeros((0, 3), np.float32)
        return dict(status=status, height=height, nn_idx=nn_idx, nn_d2=nn_d2, nn1_idx=nn1,
                    src_xyz=np.array(xs, np.float32) if xs else z, ref_xyz=np.array(ys, np.float32) if ys else z,
                    ref_n=np.array(ns, np.float32) if ns else z, src_idx=np.nonzero(status == 0)[0].astype(np.int32),
                    counters=np.array([(status == s).sum() for s in range(1, 7)], np.int64))


def rodrigues(rot):
    a = np.linalg.norm(rot)
    if a == 0:
        return np.eye(3)
    k = rot / a
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


def solve_wls(src, ref, nrm, w=None):
    """src/solver.cpp:168-220 via numpy.linalg.lstsq."""
    s, d, n = (np.asarray(a, np.float64) for a in (src, ref, nrm))
    A = np.concatenate([np.cross(s, n), n], axis=1)
    b = np.einsum("ij,ij->i", n, d - s)
    if w is not None:
        sw = np.sqrt(np.asarray(w, np.float64))
        A, b = A * sw[:, None], b * sw
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    R = rodrigues(x[:3])
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        U[:, 2] *= -1
        R = U @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = x[3:]
    return T


def register(ref: Ref, iterations=30, correspond_number=6, dd_thr=1e-3, da_thr=1.745353e-4, T0=None):
    """src/laser_odometry.cpp:484-485,524-647."""
    rPose = np.eye(4) if T0 is None else np.array(T0, np.float64)
    iters = 0
    for _ in range(iterations):
        pr = ref.project(rPose)
        if pr["src_xyz"].shape[0] < correspond_number:
            return rPose, iters, "TOO_FEW_PAIRS"
        delta = solve_wls(pr["src_xyz"], pr["ref_xyz"], pr["ref_n"])
        rPose = delta @ rPose
        iters += 1
        dd = np.linalg.norm(delta[:3, 3])
        da = np.arccos(np.clip((np.trace(delta[:3, :3]) - 1) / 2, -1, 1))
        if dd < dd_thr and da < da_thr:
            return rPose, iters, "CONVERGED"
    return rPose, iters, "MAX_ITERS"
