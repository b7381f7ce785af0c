"""Pins oracle/plo_oracle.c (CPU oracle) — no GPU needed.

The reference has no tests or golden vectors and cannot be built here, so
the oracle is pinned by (1) brute force, (2) the independent numpy/scipy restatement in
oracle/py/imls_ref.py, (3) library routines (numpy lstsq/svd/eigh, scipy Rotation).
"""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import imls_ref


def _cloud(rng, n, scale=10.0):
    rec = np.zeros((n, 12), np.float32)
    rec[:, 0:3] = rng.uniform(-scale, scale, size=(n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3))
    v[:, 2] = np.abs(v[:, 2])
    rec[:, 4:7] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    return rec


@pytest.mark.parametrize("n,k,r,allow_self", [(1, 1, 3.0, False), (5, 20, 3.0, True), (2000, 20, 3.0, True),
                                              (2000, 1, 0.5, False), (5000, 10, 1.0, False), (300, 32, 100.0, True)])
def test_tree_knn_equals_brute_force(oracle_mod, n, k, r, allow_self):
    rng = np.random.default_rng(n * 31 + k)
    rec = _cloud(rng, n)
    o = oracle_mod.Oracle()
    assert o.set_target(rec) == n
    qs = np.concatenate([rng.uniform(-11, 11, size=(60, 3)), rec[rng.integers(0, n, 20), 0:3].astype(np.float64)])
    for q in qs:
        c1, i1, d1 = o.knn(q, k, r, allow_self)
        c2, i2, d2 = o.knn(q, k, r, allow_self, brute=True)
        assert c1 == c2
        assert np.array_equal(i1, i2)
        assert np.array_equal(d1, d2)
        # libnabo contract: ascending, padded with -1 / +inf, radius and self-match rules
        assert np.all(np.diff(d1[:c1]) >= 0)
        assert np.all(i1[c1:] == -1) and np.all(np.isinf(d1[c1:]))
        assert np.all(d1[:c1] <= r * r)
        if not allow_self:
            assert np.all(d1[:c1] > np.finfo(np.float64).eps)


def test_knn_ties_broken_by_index(oracle_mod):
    # grid-aligned points: many exact distance ties (deviation D3: (d2, index) order)
    g = np.arange(-3, 4, dtype=np.float32)
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(0)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    rec = np.zeros((xyz.shape[0], 12), np.float32)
    rec[:, 0:3] = xyz
    rec[:, 6] = 1
    o = oracle_mod.Oracle()
    o.set_target(rec)
    ref = imls_ref.Ref(rec, rec[:1])
    for q in [np.zeros(3), np.array([0.5, 0.5, 0.5]), np.array([1.0, 0.0, -2.0])]:
        c, idx, d2 = o.knn(q, 20, 3.0, True)
        ri, rd = ref.knn(q, 20, 3.0, True)
        assert np.array_equal(idx, ri) and np.array_equal(d2, rd)
        for a in range(c - 1):
            assert d2[a] < d2[a + 1] or (d2[a] == d2[a + 1] and idx[a] < idx[a + 1])


def test_knn_self_match_and_duplicates(oracle_mod):
    rec = np.zeros((6, 12), np.float32)
    rec[:, 0:3] = [[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 0], [5, 5, 5]]
    o = oracle_mod.Oracle()
    o.set_target(rec)
    c, idx, d2 = o.knn(np.zeros(3), 3, 3.0, True)
    assert list(idx) == [0, 1, 4] and np.all(d2 == 0)
    c, idx, d2 = o.knn(np.zeros(3), 3, 3.0, False)     # d2 <= DBL_EPSILON rejected
    assert list(idx) == [2, 3, -1] and c == 2
    c, idx, d2 = o.knn(np.array([np.nan, 0, 0]), 3, 3.0, True)
    assert c == 0 and list(idx) == [-1, -1, -1]


def test_strip_nonfinite_reindexes(oracle_mod):
    rng = np.random.default_rng(3)
    rec = _cloud(rng, 50)
    rec[7, 0] = np.nan
    rec[20, 2] = np.inf
    rec[30, 5] = np.nan     # a non-finite NORMAL survives the strip (pcl::isFinite tests xyz only)
    o = oracle_mod.Oracle()
    assert o.set_target(rec) == 48
    assert o.set_source(rec) == 48
    c, idx, d2 = o.knn(rec[8, 0:3].astype(np.float64), 1, 1.0, True)
    assert idx[0] == 7      # indices refer to the stripped cloud, as after the reference's erase


def test_colpiv_qr_matches_lstsq(oracle_mod):
    rng = np.random.default_rng(5)
    for m in (6, 7, 50, 5000):
        A = rng.normal(size=(m, 6)) * np.array([10, 10, 10, 1, 1, 1])
        b = rng.normal(size=m)
        x, rank = oracle_mod.colpiv_qr_solve(A, b)
        assert rank == 6
        assert np.allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], rtol=1e-10, atol=1e-12)
    # exactly-zero columns are dropped (basic solution), as Eigen's nonzeroPivots() does
    A = rng.normal(size=(100, 6))
    A[:, [2, 4]] = 0
    b = rng.normal(size=100)
    x, rank = oracle_mod.colpiv_qr_solve(A, b)
    assert rank == 4 and x[2] == 0 and x[4] == 0
    ref = np.linalg.lstsq(A[:, [0, 1, 3, 5]], b, rcond=None)[0]
    assert np.allclose(x[[0, 1, 3, 5]], ref, rtol=1e-10)
    # underdetermined 3x6 (RANSAC hypothesis, src/solver.cpp:251-273): 3 pivots, rest zero
    A = rng.normal(size=(3, 6))
    b = rng.normal(size=3)
    x, rank = oracle_mod.colpiv_qr_solve(A, b)
    assert rank == 3 and (x == 0).sum() == 3
    assert np.allclose(A @ x, b, atol=1e-12)


def test_angle_axis_and_polar(oracle_mod):
    rng = np.random.default_rng(6)
    assert np.array_equal(oracle_mod.angle_axis(np.zeros(3)), np.eye(3))
    for _ in range(50):
        rot = rng.normal(size=3) * rng.choice([1e-8, 1e-3, 0.1, 2.0])
        R = oracle_mod.angle_axis(rot)
        assert np.allclose(R, Rotation.from_rotvec(rot).as_matrix(), atol=1e-14)
        M = R + rng.normal(size=(3, 3)) * 1e-3
        U, _, Vt = np.linalg.svd(M)
        P = U @ Vt
        if np.linalg.det(P) < 0:
            U[:, 2] *= -1
            P = U @ Vt
        assert np.allclose(oracle_mod.polar_uvt(M), P, atol=1e-13)
        assert np.allclose(oracle_mod.polar_uvt(R), R, atol=1e-14)
    # reflection input: det fix of src/solver.cpp:209-213
    M = np.diag([1.0, 1.0, -1.0]) + 1e-3 * rng.normal(size=(3, 3))
    P = oracle_mod.polar_uvt(M)
    assert np.linalg.det(P) > 0.99 and np.allclose(P @ P.T, np.eye(3), atol=1e-12)


def test_sym_eigen(oracle_mod):
    rng = np.random.default_rng(7)
    for n in (3, 6):
        for _ in range(20):
            B = rng.normal(size=(n, n))
            A = B @ B.T
            ev, V = oracle_mod.sym_eigen(A)
            w = np.linalg.eigvalsh(A)
            assert np.allclose(ev, w, rtol=1e-11, atol=1e-12)
            assert np.allclose(A @ V, V * ev[None, :], atol=1e-10)
            assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)


def test_compute_normal(oracle_mod):
    rng = np.random.default_rng(8)
    n_true = np.array([0.2, -0.3, 0.93])
    n_true /= np.linalg.norm(n_true)
    basis = np.linalg.svd(n_true[None, :])[2][1:]
    pts = rng.normal(size=(10, 2)) @ basis + 5.0
    n = oracle_mod.compute_normal(pts)
    assert np.allclose(n, n_true, atol=1e-9)                       # exact plane -> its normal, +z oriented (D2)
    pts = pts + rng.normal(size=pts.shape) * 0.01
    mu = pts.mean(0)
    w, V = np.linalg.eigh((pts - mu).T @ (pts - mu) / 10)
    ref = V[:, 0] * np.sign(V[2, 0])
    assert np.allclose(oracle_mod.compute_normal(pts), ref, atol=1e-9)
