"""Pins the matcher / solver / driver-loop part of the CPU oracle (no GPU needed):
C oracle vs the independent numpy restatement, analytic known answers, and one test per
drop reason of src/imls_icp.cpp:506-511."""
import numpy as np
import pytest

import imls_ref
import plo_b200 as plo

W = plo.synth.workloads


def _rot_err(Ra, Rb):
    return np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))


def _compare_projection(oracle_mod, target, source, T=None, **kw):
    T = np.eye(4) if T is None else T
    o = oracle_mod.Oracle(oracle_mod.default_params(**{
        "h": kw.get("h", 1.0), "r": kw.get("r", 3.0), "search_number": kw.get("k", 20),
        "normal_angle_constraint": int(kw.get("angle_constraint", True)),
        "angle_diff_threshold": kw.get("angle_thr", 30.0),
        "is_get_normals": int(kw.get("is_get_normals", True)),
        "transform_normal": int(kw.get("transform_normal", False))}))
    o.set_target(target)
    o.set_source(source)
    a = o.project(T, hooks=True)
    ref = imls_ref.Ref(target, source, h=kw.get("h", 1.0), r=kw.get("r", 3.0), k=kw.get("k", 20),
                       angle_constraint=kw.get("angle_constraint", True), angle_thr=kw.get("angle_thr", 30.0),
                       is_get_normals=kw.get("is_get_normals", True), transform_normal=kw.get("transform_normal", False))
    b = ref.project(T)
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["counters"], b["counters"])
    assert np.array_equal(a["nn_idx"], b["nn_idx"])           # bit-exact neighbour sets
    assert np.array_equal(a["nn_d2"], b["nn_d2"])
    assert np.array_equal(a["nn1_idx"], b["nn1_idx"])
    assert np.array_equal(a["src_idx"], b["src_idx"])
    ok = a["status"] == 0
    if ok.any():
        err = np.abs(a["height"][ok] - b["height"][ok])
        assert (err <= 1e-9 * np.abs(b["height"][ok]) + 1e-12).all()
        assert np.array_equal(a["src_xyz"], b["src_xyz"])
        assert np.abs(a["ref_xyz"].astype(np.float64) - b["ref_xyz"]).max() <= 1e-6
        assert np.array_equal(a["ref_n"], b["ref_n"])
    return a, o, ref


def test_projection_matches_numpy_restatement_hdl64(oracle_mod):
    pair = W.hdl64_pair(max_source=1500, max_target=20000)
    a, _, _ = _compare_projection(oracle_mod, pair.target, pair.source)
    assert a["n"] > 1000
    # a non-identity pose exercises the double->float32 transform round trip
    _compare_projection(oracle_mod, pair.target, pair.source, T=pair.T_gt)


def test_projection_sparse_planetary_large_h(oracle_mod):
    pair = W.planetary_pair()
    src = pair.source[::12]
    for h in (1.0, 2.0):
        a, _, _ = _compare_projection(oracle_mod, pair.target, src, h=h, r=3 * h)
    assert a["counters"].sum() > 0        # neighbour-starved path really drops points


def test_projection_pca_normals_and_transform_normal(oracle_mod):
    pair = W.hdl64_pair(max_source=400, max_target=6000)
    a, o, ref = _compare_projection(oracle_mod, pair.target, pair.source, is_get_normals=False,
                                    transform_normal=True, T=pair.T_gt)
    tn = o.target_normals()
    fin = np.isfinite(tn).all(axis=1)
    assert np.array_equal(fin, np.isfinite(ref.tn).all(axis=1))
    assert np.abs(tn[fin] - ref.tn[fin]).max() < 1e-7
    assert (tn[fin][:, 2] >= 0).all()                     # D2
    assert (~fin).any() and fin.any()


def _plane_cloud(n, rng, z=0.0, normal=(0, 0, 1), spread=4.0):
    rec = np.zeros((n, 12), np.float32)
    rec[:, 0:2] = rng.uniform(-spread, spread, size=(n, 2))
    rec[:, 2] = z
    rec[:, 4:7] = normal
    return rec


def test_each_drop_reason(oracle_mod):
    rng = np.random.default_rng(11)
    tgt = _plane_cloud(3000, rng)
    tgt[0, 0:3] = [50, 50, 0]
    tgt[0, 4:7] = [np.nan, 0, 1]                           # invalid normal on an isolated point
    tgt[1, 0:3] = [80, 80, 0]                              # isolated: < 3 IMLS neighbours
    tgt[2, 0:3] = [90, 90, 0]                              # three exact duplicates: h_max = 0 -> NaN height
    tgt[3, 0:3] = [90, 90, 0]
    tgt[4, 0:3] = [90, 90, 0]
    src = np.zeros((7, 12), np.float32)
    src[:, 4:7] = [0, 0, 1]
    src[0, 0:3] = [0.1, 0.2, 0.05]                         # OK
    src[1, 0:3] = [200, 200, 0]                            # no neighbour within r      -> no_normal
    src[2, 0:3] = [0.0, 0.0, 2.0]                          # nearest within r, > h      -> too_far
    src[3, 0:3] = [50, 50, 0.1]                            # nearest has NaN normal     -> invalid_normal
    src[4, 0:3] = [0.3, 0.1, 0.05]
    src[4, 4:7] = [1, 0, 0]                                # 90 deg off                 -> normal_constraint
    src[5, 0:3] = [80, 80, 0.1]                            # single neighbour           -> mls_fail
    src[6, 0:3] = [90, 90, 0.0]                            # query ON 3 duplicates: 1-NN (no self match)
    a, _, _ = _compare_projection(oracle_mod, tgt, src)
    assert list(a["status"][:6]) == [0, 1, 2, 3, 4, 5]
    # query 6: all three neighbours have d2 == 0 <= DBL_EPSILON, so the 1-NN (no self match) finds nothing
    assert a["status"][6] == 1
    src2 = src[6:7].copy()
    src2[0, 0:3] = [90, 90, 1e-4]                          # 1-NN valid, h_max = sqrt(1e-8)/3 fine -> OK or NaN
    b, _, _ = _compare_projection(oracle_mod, tgt, src2)
    assert b["status"][0] in (0, 6)
    # zero source normal: cos = 0/0 = NaN, `angle > thr` is false => kept (src/imls_icp.cpp:681-692)
    src3 = src[0:1].copy()
    src3[0, 4:7] = 0
    c, _, _ = _compare_projection(oracle_mod, tgt, src3)
    assert c["status"][0] == 0


def test_nan_height_dropped(oracle_mod):
    # k neighbours all at the query position except the k-th: filtered count 3 with d2[2] == 0 -> h_max = 0
    tgt = np.zeros((4, 12), np.float32)
    tgt[:, 6] = 1
    tgt[0:3, 0:3] = [1, 1, 1]
    tgt[3, 0:3] = [1.5, 1, 1]
    src = np.zeros((1, 12), np.float32)
    src[0, 0:3] = [1, 1, 1]
    src[0, 6] = 1
    a, _, _ = _compare_projection(oracle_mod, tgt, src)
    # 1-NN without self match is the 4th point (d2 = .25); IMLS keeps all 4: h_max = sqrt(d2[3])/3 > 0 -> finite
    assert a["status"][0] == 0
    tgt[3, 4:7] = [np.nan, 0, 0]                           # now only the 3 coincident points survive the filter
    tgt2 = np.concatenate([tgt, tgt[3:4]])
    tgt2[4, 0:3] = [1.2, 1, 1]
    tgt2[4, 4:7] = [0, 0, 1]                               # valid 1-NN; filtered cnt = 4, d2[3] = 0.04 -> fine
    b, _, _ = _compare_projection(oracle_mod, tgt2, src)
    assert b["status"][0] == 0
    tgt3 = tgt2.copy()
    tgt3[4, 4:7] = [1, 0, 0]                               # 1-NN passes? its normal is 90deg off -> constraint drop
    c, _, _ = _compare_projection(oracle_mod, tgt3, src)
    assert c["status"][0] == 4
    c2, _, _ = _compare_projection(oracle_mod, tgt3, src, angle_constraint=False)
    assert c2["status"][0] == 0


def test_imls_height_on_exact_plane(oracle_mod):
    # SURVEY.md §8c known answer: query above an exact plane with consistent normals:
    # I(x) = signed distance * sum(w) / (sum(w) + 1e-5)
    rng = np.random.default_rng(12)
    tgt = _plane_cloud(4000, rng, z=-1.0)
    src = np.zeros((50, 12), np.float32)
    src[:, 0:2] = rng.uniform(-2, 2, size=(50, 2))
    src[:, 2] = -1.0 + rng.uniform(-0.3, 0.3, size=50).astype(np.float32)
    src[:, 6] = 1
    o = oracle_mod.Oracle()
    o.set_target(tgt)
    o.set_source(src)
    a = o.project(np.eye(4), hooks=True)
    assert (a["status"] == 0).all()
    dist = src[:, 2].astype(np.float64) + 1.0
    for i in range(50):
        d2 = a["nn_d2"][i]
        hmax = np.sqrt(d2[19]) / 3
        sw = np.exp(-d2 / hmax / hmax).sum()
        assert abs(a["height"][i] - dist[i] * sw / (sw + 1e-5)) < 1e-12
    # projected points land on the plane up to the 1e-5 regulariser and float32 storage
    assert np.abs(a["ref_xyz"][:, 2] + 1.0).max() < 1e-4


def test_wls_matches_lstsq_and_normal_equations(oracle_mod):
    pair = W.hdl64_pair(max_source=3000, max_target=30000)
    o = oracle_mod.Oracle()
    o.set_target(pair.target)
    o.set_source(pair.source)
    pr = o.project(np.eye(4))
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    D = oracle_mod.solve_wls(s, d, n)
    Dr = imls_ref.solve_wls(s, d, n)
    assert np.abs(D - Dr).max() < 1e-12
    rng = np.random.default_rng(1)
    w = rng.uniform(0.1, 1.0, size=s.shape[0])
    assert np.abs(oracle_mod.solve_wls(s, d, n, w) - imls_ref.solve_wls(s, d, n, w)).max() < 1e-12
    # the 21+6 sums the GPU reduces reproduce the same solution through H x = g
    H21, g, sw, sbb = oracle_mod.normal_equations(s, d, n, w)
    H = np.zeros((6, 6))
    H[np.triu_indices(6)] = H21
    H = H + np.triu(H, 1).T
    x = np.linalg.solve(H, g)
    A = np.concatenate([np.cross(s, n), n], axis=1) * np.sqrt(w)[:, None]
    b = np.einsum("ij,ij->i", n, d - s) * np.sqrt(w)
    assert np.allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], rtol=1e-9, atol=1e-13)
    assert np.isclose(sw, w.sum()) and np.isclose(sbb, (b * b).sum())


def test_register_recovers_rigid_transform(oracle_mod):
    # SURVEY.md §8c known answer: source = rigidly moved copy of a plane + two walls scene
    pair = W.rigid_copy_pair(n=6000)
    o = oracle_mod.Oracle()
    o.set_target(pair.target)
    o.set_source(pair.source)
    T, st = o.register()
    assert st["status"] == 1 and st["iters"] < 30
    assert _rot_err(T[:3, :3], pair.T_gt[:3, :3]) < 2e-5
    assert np.linalg.norm(T[:3, 3] - pair.T_gt[:3, 3]) < 2e-4
    # and the independent restatement follows the same trajectory
    ref = imls_ref.Ref(pair.target[::3], pair.source[::40])
    o2 = oracle_mod.Oracle()
    o2.set_target(pair.target[::3])
    o2.set_source(pair.source[::40])
    T2, st2 = o2.register()
    Tr, iters, status = imls_ref.register(ref)
    assert iters == st2["iters"] and status == oracle_mod.REG_STATUS[st2["status"]]
    assert np.abs(T2 - Tr).max() < 1e-10


def test_register_hdl64_close_to_ground_truth(oracle_mod):
    pair = W.hdl64_pair(max_source=8000, max_target=60000)
    o = oracle_mod.Oracle()
    o.set_target(pair.target)
    o.set_source(pair.source)
    T, st = o.register()
    assert st["status"] == 1
    assert _rot_err(T[:3, :3], pair.T_gt[:3, :3]) < 2e-3
    assert np.linalg.norm(T[:3, 3] - pair.T_gt[:3, 3]) < 0.02


def test_pure_plane_has_three_observable_dof(oracle_mod):
    # pure-plane scene => exactly 3 observable DoF; Eigen's ColPivHouseholderQR
    # zeroes the unobservable ones (exactly-zero columns), the loop still converges
    rng = np.random.default_rng(13)
    tgt = _plane_cloud(5000, rng, z=-1.5, spread=6)
    src = _plane_cloud(800, rng, z=-1.5, spread=3)
    T = plo.synth.scenes.pose_matrix([0.0, 0.0, 0.05], pitch_deg=0.5, roll_deg=-0.4)
    Ti = np.linalg.inv(T)
    p = src[:, 0:3].astype(np.float64)
    src[:, 0:3] = (p @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    o = oracle_mod.Oracle()
    o.set_target(tgt)
    o.set_source(src)
    pr = o.project(np.eye(4))
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    A = np.concatenate([np.cross(s, n), n], axis=1)
    x, rank = oracle_mod.colpiv_qr_solve(A, np.einsum("ij,ij->i", n, d - s))
    assert rank == 3 and x[2] == 0 and x[3] == 0 and x[4] == 0
    Tr, st = o.register()
    assert st["status"] == 1
    assert abs(Tr[2, 3] - T[2, 3]) < 2e-3
    # unobservable x/y: never solved for (only rotated by later deltas' R * t_z terms)
    assert abs(Tr[0, 3]) < 1e-3 and abs(Tr[1, 3]) < 1e-3


def test_too_few_pairs_breaks_loop(oracle_mod):
    rng = np.random.default_rng(14)
    tgt = _plane_cloud(100, rng)
    src = _plane_cloud(20, rng, z=50.0)
    o = oracle_mod.Oracle()
    o.set_target(tgt)
    o.set_source(src)
    T, st = o.register()
    assert st["status"] == 3 and st["iters"] == 0 and np.array_equal(T, np.eye(4))
    o.set_target(tgt[:0])
    T, st = o.register()
    assert st["status"] == 3 and st["counters"][0] == 20
    o.set_source(src[:0])
    assert o.project(np.eye(4))["n"] == 0


def _drpm_numpy(s, d, n, w, threshold=0.05, sp=0.02, sn=0.05):
    from scipy.stats import norm
    N = s.shape[0]
    w = np.ones(N) if w is None else w
    A = np.concatenate([np.cross(s, n), n], axis=1) * np.sqrt(w)[:, None]
    b = np.einsum("ij,ij->i", n, d - s) * np.sqrt(w)
    H = A.T @ A
    ev, U = np.linalg.eigh(H)

    def skew(v):
        z = np.zeros(v.shape[0])
        return np.stack([np.stack([z, -v[:, 2], v[:, 1]], -1), np.stack([v[:, 2], z, -v[:, 0]], -1),
                         np.stack([-v[:, 1], v[:, 0], z], -1)], -2)
    nx, px = skew(n), skew(s)
    B = np.zeros((N, 6, 6))
    B[:, 0:3, 0:3] = -nx
    B[:, 0:3, 3:6] = px @ nx
    B[:, 3:6, 3:6] = nx
    Nm = np.diag([sp * sp] * 3 + [sn * sn] * 3)
    Cm = (B @ Nm @ B.transpose(0, 2, 1)) * w[:, None, None]
    mean = Cm.sum(0)
    v = np.concatenate([np.einsum("nij,nj->ni", px, n), n], axis=1) * np.sqrt(w)[:, None]
    a = np.einsum("ik,nij,jk->nk", U, Cm, U)
    bb = v @ U
    var = (2 * a * a + 4 * a * bb * bb).sum(0)
    meas = np.einsum("ik,ij,jk->k", U, H, U)
    noise = np.einsum("ik,ij,jk->k", U, mean, U)
    probs = norm.cdf(meas / 11.0, loc=noise, scale=np.sqrt(var))
    if probs.min() < threshold:
        with np.errstate(divide="ignore", invalid="ignore"):
            dps = np.where(np.abs(ev) > 1e-10, probs / ev, 0.0)
        x = U @ (dps * (U.T @ (A.T @ b)))
    else:
        x = np.linalg.lstsq(A, b, rcond=None)[0]
    T = np.eye(4)
    T[:3, :3] = imls_ref.rodrigues(x[:3])
    T[:3, 3] = x[3:]
    return T, probs


def test_trimmed_ls_ransac_weights_drpm(oracle_mod):
    pair = W.hdl64_pair(max_source=3000, max_target=30000)
    o = oracle_mod.Oracle()
    o.set_target(pair.target)
    o.set_source(pair.source)
    pr = o.project(np.eye(4))
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    N = s.shape[0]
    # trimmed LS, src/solver.cpp:74-166
    A = np.concatenate([np.cross(s, n), n], axis=1)
    b = np.einsum("ij,ij->i", n, d - s)
    x0 = np.linalg.lstsq(A, b, rcond=None)[0]
    order = np.lexsort((np.arange(N), np.abs(A @ x0 - b)))
    lo, hi = int(0.02 * N), int(0.98 * N)
    sel = order[lo:hi + 1]
    Dref = imls_ref.solve_wls(s[sel], d[sel], n[sel])
    assert np.abs(oracle_mod.solve_ls(s, d, n, 0.02) - Dref).max() < 1e-10
    # RANSAC-final weights at T_best = I: closed form of SURVEY.md §10.2
    idx, w = oracle_mod.ransac_weights(s, d, n)
    dist = np.abs(b)
    inl = np.nonzero(dist < 0.8)[0]
    assert np.array_equal(idx, inl.astype(np.int32))
    wr = 1.0368 * np.exp(-dist[inl] / 2) - 0.26873856
    assert np.allclose(w, wr / wr.sum(), rtol=1e-12)
    # DRPM (src/solver.cpp:499-603, include/degeneracy.h:14-131) vs a vectorised numpy restatement
    for (ss, dd, nn_, ww) in ((s[inl], d[inl], n[inl], w), (s[inl][:200] * 0.05, d[inl][:200] * 0.05, n[inl][:200], None)):
        D, probs = oracle_mod.solve_drpm(ss, dd, nn_, ww)
        Dn, pn = _drpm_numpy(ss, dd, nn_, ww)
        assert np.allclose(probs, pn, rtol=1e-6, atol=1e-12)
        assert np.abs(D - Dn).max() < 1e-9
    # RANSAC wrapper (seeded FPS hypotheses) with the Weighted-LS tail: weights are evaluated at a
    # random 3-point hypothesis, so only closeness (cm-level) to the T_best = I weights is expected
    p = oracle_mod.default_params(solver=2, ransac_final=1)
    ok, Dr = oracle_mod.solve_ransac(s, d, n, p)
    assert ok and np.abs(Dr - oracle_mod.solve_wls(s[inl], d[inl], n[inl], w)).max() < 3e-2
    # RANSAC -> "LS" tail (src/solver.cpp:366-371): trimmed LS on the inliers of the best hypothesis.  With a threshold
    # no pair can fail, the inlier set is the whole list whatever hypothesis won: exactly the plain trimmed LS
    p = oracle_mod.default_params(solver=2, ransac_final=0, ransac_distance_threshold=1e6)
    ok, Dl = oracle_mod.solve_ransac(s, d, n, p)
    assert ok and np.abs(Dl - oracle_mod.solve_ls(s, d, n, 0.02)).max() < 1e-12
    # ... and at the default threshold it stays within centimetres of the trimmed LS on the T = I inliers
    ok, Dl = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(solver=2, ransac_final=0))
    assert ok and np.abs(Dl - oracle_mod.solve_ls(s[inl], d[inl], n[inl], 0.02)).max() < 3e-2
    # degenerate plane: DRPM damps the unobservable directions instead of blowing up
    rng = np.random.default_rng(2)
    sp = np.zeros((500, 3))
    sp[:, 0:2] = rng.uniform(-5, 5, size=(500, 2))
    nn = np.tile([0.0, 0.0, 1.0], (500, 1)) + rng.normal(size=(500, 3)) * 1e-4
    dp = sp + np.array([0, 0, 0.01])
    D, probs = oracle_mod.solve_drpm(sp, dp, nn, None)
    assert probs.min() < 0.05 and np.isfinite(D).all() and np.abs(D[:3, 3]).max() < 0.1


def test_huber_exp_weight_mode_and_solver_dispatch(oracle_mod):
    pair = W.hdl64_pair(max_source=2500, max_target=25000)
    res = {}
    for name, kw in {"unit": {}, "huber": {"weight_mode": 1}, "ls": {"solver": 1},
                     "ransac_ls": {"solver": 2, "ransac_final": 0}, "ransac_wls": {"solver": 2, "ransac_final": 1},
                     "ransac_drpm": {"solver": 2}}.items():
        o = oracle_mod.Oracle(oracle_mod.default_params(**kw))
        o.set_target(pair.target)
        o.set_source(pair.source)
        T, st = o.register()
        assert st["status"] == 1, name
        res[name] = T
        if name == "ransac_drpm":
            # the literal DRPM tail (config.json default) judges the along-corridor direction of this
            # scene degenerate (probability ~1e-41 with stdev_normals=0.05 and 30 m lever arms) and
            # damps it; restated faithfully, cross-checked against numpy above, not a GPU row yet
            assert np.isfinite(T).all()
            continue
        assert np.linalg.norm(T[:3, 3] - pair.T_gt[:3, 3]) < 0.03, name
    # RANSAC (first hypothesis accepted or not) + Weighted LS final ~ huber_exp WLS at T_best=I
    assert np.abs(res["huber"] - res["unit"]).max() < 5e-3


def test_transform_to_end_matches_numpy(oracle_mod):
    """orc_transform_to_end (TransformToEnd, src/laser_odometry.cpp:88-114): p' = R^T (p - t), n' = R^T n."""
    rng = np.random.default_rng(31)
    rec = np.zeros((500, 12), np.float32)
    rec[:, 0:3] = rng.uniform(-40, 40, size=(500, 3))
    nrm = rng.normal(size=(500, 3))
    rec[:, 4:7] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    rec[:, 3] = 7.0
    a, b, c = 0.03, -0.01, 0.02
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(c), -np.sin(c)], [0, np.sin(c), np.cos(c)]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [0.7, -0.05, 0.02]
    out = oracle_mod.transform_to_end(rec, T, True)
    p = (rec[:, 0:3].astype(np.float64) - T[:3, 3]) @ T[:3, :3]          # rows: R^T (p - t)
    n = rec[:, 4:7].astype(np.float64) @ T[:3, :3]
    assert np.abs(out[:, 0:3] - p).max() < 4e-6 and np.abs(out[:, 4:7] - n).max() < 1e-7     # float32 store
    assert np.array_equal(out[:, 3], rec[:, 3])                                                # other fields untouched
    keep = oracle_mod.transform_to_end(rec, T, False)
    assert np.array_equal(keep[:, 4:7], rec[:, 4:7]) and np.array_equal(keep[:, 0:3], out[:, 0:3])
    # round trip: forward with T, back with T^-1
    back = oracle_mod.transform_to_end(out, np.linalg.inv(T), True)
    assert np.abs(back[:, 0:3] - rec[:, 0:3]).max() < 1e-5
    ident = oracle_mod.transform_to_end(rec, np.eye(4), True)
    assert np.array_equal(ident, rec)


def test_frontend_oracle_properties(oracle_mod):
    """oracle/plo_oracle_frontend.c (src/scan_registration.cpp front-end): selection rules, order, and one PCA window
    re-done in numpy."""
    import plo_b200 as plo
    pair = plo.synth.workloads.planetary_pair()                        # VLP-16: small enough for the CPU suite
    pts = np.ascontiguousarray(pair.source[:, 0:3])
    r = oracle_mod.frontend(pts, oracle_mod.frontend_default_params(n_scans=16, plane_distance_threshold=0.05))
    assert 0 < r["n"] <= r["ringed"] <= pts.shape[0]
    rec = r["records"]
    assert np.array_equal(rec[:, 0:3], pts[r["src_index"]])            # points are copied, never altered
    assert np.allclose(np.linalg.norm(rec[:, 4:7], axis=1), 1.0, atol=1e-6) and (rec[:, 6] >= 0).all()   # unit, +z
    ring = np.floor(rec[:, 8]).astype(int)                             # intensity = ring + 0.1 * relTime
    assert (np.diff(ring) >= 0).all() and ring.min() >= 1 and ring.max() <= 14     # ring-major, rings 1 .. N-2 only
    assert ((rec[:, 8] - ring) >= 0).all() and ((rec[:, 8] - ring) < 0.2).all()      # relTime of a ring-major synthetic scan
    ev = r["eigenvalues"]
    okp = ev[:, 0] > 0
    assert (ev[okp, 0] >= ev[okp, 1]).all() and (ev[okp, 1] >= ev[okp, 2] - 1e-9).all()
    assert np.array_equal(ev[~okp], np.full(((~okp).sum(), 3), -1.0, np.float32)) and (~okp).sum() == r["plane_failures"]
    plan = (ev[:, 1] - ev[:, 2]) / ev[:, 0]
    assert np.array_equal(r["candidate"], okp & (plan > 0.05)) and r["candidate"].sum() == r["candidates"]
    # analytic normals of the synthetic terrain agree where the plane check held
    cosang = np.abs((pair.source[r["src_index"], 4:7] * rec[:, 4:7]).sum(axis=1))
    assert np.median(cosang[okp]) > 0.99
    # one window re-done in numpy (float64 eigh): the 21 rows are the +-3 neighbours on the ring and around the
    # nearest points of the two adjacent rings
    ang = np.degrees(np.arctan(pts[:, 2] / np.hypot(pts[:, 0], pts[:, 1])))
    rid = np.floor((ang + 15) / 2 + 0.5).astype(int)
    k = int(np.nonzero(okp)[0][len(np.nonzero(okp)[0]) // 2])
    i = ring[k]
    rows = []
    own = pts[rid == i]
    j = int(np.nonzero((own == rec[k, 0:3]).all(axis=1))[0][0])
    rows.append(own[j - 3:j + 4])
    for nbr in (i - 1, i + 1):
        cl = pts[rid == nbr]
        nn = int(np.argmin(((cl - rec[k, 0:3]) ** 2).sum(axis=1)))
        rows.append(cl[nn - 3:nn + 4])
    P = np.concatenate(rows).astype(np.float64)
    assert P.shape == (21, 3)
    C = np.cov(P.T)
    w, V = np.linalg.eigh(C)
    n = V[:, 0] if V[2, 0] >= 0 else -V[:, 0]
    assert abs(abs(n @ rec[k, 4:7]) - 1) < 1e-4 and np.allclose(w[::-1], ev[k], rtol=2e-3, atol=1e-7)
    # HDL-64: rings above 50 are dropped (:1003), NaN / out-of-range points never reach a ring
    hd = plo.synth.workloads.hdl64_pair(azimuth_steps=500)
    q = np.ascontiguousarray(hd.source[:, 0:3]).copy()
    q[::50, 0] = np.nan
    q[3::70] *= 500.0
    r64 = oracle_mod.frontend(q)
    assert r64["ringed"] < np.isfinite(q).all(axis=1).sum() and np.isfinite(r64["records"]).all()
    assert oracle_mod.frontend(q[:0])["n"] == 0
