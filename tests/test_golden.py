"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py).

The reference ships no fixtures and cannot be built here: the vectors are oracle outputs
that the independent numpy/scipy restatement reproduced at generation time.  CPU tests keep the oracle
and the numpy restatement on those numbers; the GPU test holds the CUDA path (through the C ABI) to them
without any CPU code in the loop.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("urban_hdl64", "planetary_vlp16")
HEIGHT_RTOL, HEIGHT_ATOL = 1e-9, 1e-12
POSE_RAD, POSE_M = 1e-5, 1e-4


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _rot_err(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def _kw(g):
    return dict(h=float(g["h"]), r=float(g["r"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle_mod, name):
    g = _load(name)
    o = oracle_mod.Oracle(oracle_mod.default_params(**_kw(g)))
    o.set_target(g["target"])
    o.set_source(g["source"])
    pr = o.project(np.eye(4), hooks=True)
    for k in ("nn_idx", "nn_d2", "nn1_idx", "nn1_d2", "status", "counters", "src_idx", "src_xyz", "ref_n", "ref_xyz"):
        assert np.array_equal(pr[k], g[k]), k
    ok = g["status"] == 0
    assert np.allclose(pr["height"][ok], g["height"][ok], rtol=1e-13, atol=0)
    pg = o.project(g["T_gt"], hooks=True)
    assert np.array_equal(pg["nn_idx"], g["gt_nn_idx"]) and np.array_equal(pg["status"], g["gt_status"])
    assert np.array_equal(pg["counters"], g["gt_counters"])
    s, d, n = (g[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    H, gg, sw, sbb = oracle_mod.normal_equations(s, d, n)
    assert np.allclose(H, g["H"], rtol=1e-13) and np.allclose(gg, g["g"], rtol=1e-12, atol=1e-12)
    assert np.abs(oracle_mod.solve_wls(s, d, n) - g["delta_wls"]).max() < 1e-12
    assert np.abs(oracle_mod.solve_ls(s, d, n) - g["delta_ls"]).max() < 1e-12
    for final, key in ((1, "delta_ransac_wls"), (2, "delta_ransac_drpm")):
        ok_, D = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(solver=2, ransac_final=final, **_kw(g)))
        assert ok_ and np.abs(D - g[key]).max() < 1e-12
    T, st = o.register()
    assert [st["status"], st["iters"], st["pairs"]] == list(g["reg_wls_stats"]) and np.abs(T - g["reg_wls_T"]).max() < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_numpy_restatement_reproduces_golden(name):
    """The second, independent implementation (cKDTree candidates + numpy) on a slice of the queries."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "py"))
    import imls_ref
    g = _load(name)
    sel = slice(0, 200)
    ref = imls_ref.Ref(g["target"], g["source"][sel], **_kw(g))
    rp = ref.project(np.eye(4))
    assert np.array_equal(rp["nn_idx"], g["nn_idx"][sel]) and np.array_equal(rp["nn_d2"], g["nn_d2"][sel])
    assert np.array_equal(rp["status"], g["status"][sel])
    ok = g["status"][sel] == 0
    assert np.allclose(rp["height"][ok], g["height"][sel][ok], rtol=1e-12, atol=1e-15)
    s, d, n = (g[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    assert np.abs(imls_ref.solve_wls(s, d, n) - g["delta_wls"]).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_golden(name):
    """No oracle in this test: the CUDA path against the committed numbers only."""
    import plo_b200 as plo
    g = _load(name)
    ctx = plo.Context(0, plo.default_params(**_kw(g)))
    ctx.set_target(g["target"])
    ctx.set_source(g["source"])
    st = ctx.project(np.eye(4), hooks=True)
    got = {**ctx.neighbors(), **ctx.query_results(), **ctx.pairs()}
    for k in ("nn_idx", "nn_d2", "nn1_idx", "nn1_d2", "status", "src_idx", "src_xyz", "ref_n"):
        assert np.array_equal(got[k], g[k]), k
    assert np.array_equal(st["counters"], g["counters"])
    ok = g["status"] == 0
    assert (np.abs(got["height"][ok] - g["height"][ok]) <= HEIGHT_RTOL * np.abs(g["height"][ok]) + HEIGHT_ATOL).all()
    assert np.abs(got["ref_xyz"].astype(np.float64) - g["ref_xyz"]).max() <= 2e-6
    H, gg, sw, sbb, cnt = ctx.normal_equations()
    assert np.allclose(H, g["H"], rtol=1e-12) and np.allclose(gg, g["g"], rtol=1e-10, atol=1e-10)
    assert np.abs(ctx.solve_wls()[0] - g["delta_wls"]).max() < 1e-10
    ctx.set_params(plo.default_params(solver=1, **_kw(g)))
    ctx.project(np.eye(4))
    assert np.abs(ctx.solve_ls()[0] - g["delta_ls"]).max() < 1e-9
    for final, key in ((1, "delta_ransac_wls"), (2, "delta_ransac_drpm")):
        ctx.set_params(plo.default_params(solver=2, ransac_final=final, **_kw(g)))
        ctx.project(np.eye(4))
        assert np.abs(ctx.solve_ransac()[0] - g[key]).max() < 1e-9
    for label, skw in (("wls", {}), ("ls", dict(solver=1)), ("ransac_wls", dict(solver=2, ransac_final=1)),
                       ("ransac_drpm", dict(solver=2, ransac_final=2))):
        ctx.set_params(plo.default_params(**_kw(g), **skw))
        T, rs = ctx.register()
        assert [rs["status"], rs["iters"], rs["pairs"]] == list(g[f"reg_{label}_stats"]), label
        Tg = g[f"reg_{label}_T"]
        assert _rot_err(T[:3, :3], Tg[:3, :3]) < POSE_RAD and np.linalg.norm(T[:3, 3] - Tg[:3, 3]) < POSE_M, label
    pg = ctx.project(g["T_gt"], hooks=True)
    got = {**ctx.neighbors(), **ctx.query_results()}
    assert np.array_equal(got["nn_idx"], g["gt_nn_idx"]) and np.array_equal(got["status"], g["gt_status"])
    assert np.array_equal(pg["counters"], g["gt_counters"])


def _frontend_golden():
    g = _load("frontend_vlp16")
    return g, dict(n_scans=int(g["n_scans"]), plane_distance_threshold=float(g["plane_distance_threshold"]))


def test_frontend_golden_normals_against_numpy_eigh_every_point():
    """Independent of the oracle and of the CUDA path (which mirror each other statement by statement): EVERY normal
    of the golden front-end output against a float64 numpy.linalg.eigh of its 3 x 7 window (src/scan_registration.cpp
    :158-229: 7 consecutive points of the point's own ring and of the rings below / above around their nearest
    point), rebuilt here from ring ids alone.  The float32 PCA of the path agrees to 1e-6 on |cos| for all of them."""
    g, kw = _frontend_golden()
    pts, rec, ev, src = g["points"], g["records"], g["eigenvalues"], g["src_index"]
    ang = np.degrees(np.arctan(pts[:, 2] / np.hypot(pts[:, 0], pts[:, 1])))
    rid = np.floor((ang + 15) / 2 + 0.5).astype(int)          # VLP-16 ring rule, src/scan_registration.cpp:948-950
    rings = {i: pts[rid == i] for i in range(16)}
    ok = np.nonzero(ev[:, 0] > 0)[0]
    assert len(ok) > 500
    checked, worst = 0, 0.0
    for k in ok:
        q, i = rec[k, 0:3], rid[src[k]]
        own = rings[i]
        j = int(np.nonzero((own == q).all(axis=1))[0][0])
        rows = [own[j - 3:j + 4]]
        for nbr in (i - 1, i + 1):
            cl = rings[nbr]
            nn = int(np.argmin(((cl - q) ** 2).sum(axis=1)))
            rows.append(cl[nn - 3:nn + 4])
        P = np.concatenate(rows).astype(np.float64)
        assert P.shape == (21, 3), k                               # a point with a normal has three complete windows
        w, V = np.linalg.eigh(np.cov(P.T))
        worst = max(worst, abs(abs(V[:, 0] @ rec[k, 4:7]) - 1))
        assert np.allclose(w[::-1], ev[k], rtol=5e-3, atol=1e-7), k
        checked += 1
    assert checked == len(ok) and worst < 1e-6, (checked, worst)


def test_frontend_oracle_reproduces_golden(oracle_mod):
    g, kw = _frontend_golden()
    r = oracle_mod.frontend(g["points"], oracle_mod.frontend_default_params(**kw))
    assert [r["n"], r["ringed"], r["pca_failures"], r["plane_failures"], r["candidates"]] == list(g["stats"])
    assert np.array_equal(r["src_index"], g["src_index"]) and np.array_equal(r["candidate"], g["candidate"])
    assert np.array_equal(r["records"][:, 0:8], g["records"][:, 0:8]) and np.array_equal(r["eigenvalues"], g["eigenvalues"])
    assert np.abs(r["records"][:, 8] - g["records"][:, 8]).max() <= 1e-6


@pytest.mark.gpu
def test_frontend_cuda_reproduces_golden():
    import plo_b200 as plo
    g, kw = _frontend_golden()
    r = plo.Context(0).frontend(g["points"], plo.frontend_default_params(**kw))
    assert [r["n"], r["ringed"], r["pca_failures"], r["plane_failures"], r["candidates"]] == list(g["stats"])
    assert np.array_equal(r["src_index"], g["src_index"]) and np.array_equal(r["candidate"], g["candidate"])
    assert np.array_equal(r["records"][:, 0:8], g["records"][:, 0:8]) and np.array_equal(r["eigenvalues"], g["eigenvalues"])
    assert np.abs(r["records"][:, 8] - g["records"][:, 8]).max() <= 1e-6
