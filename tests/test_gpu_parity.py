"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every call goes through the C ABI
(include/plo/plo_c_api.h) via ctypes; the CPU oracle (oracle/) is only the checker.

Bars (BASELINE.json north_star / SURVEY.md §8d): neighbour index sets bit-exact (ties by index),
fp64 d2 bit-exact, drop counters exact, IMLS height rel 1e-9, 21+6 sums rel 1e-12,
6-vector solve abs 1e-10, per-frame pose 1e-5 rad / 1e-4 m.
"""
import ctypes as C

import numpy as np
import pytest

import plo_b200 as plo

pytestmark = pytest.mark.gpu
W = plo.synth.workloads

# stated tolerance for the IMLS residual I(x): |dI| <= 1e-9 * |I| + 1e-12 m.  The absolute floor covers
# heights that are sums of cancelling terms (symmetric neighbourhoods); it is 5 orders of magnitude below the
# float32 resolution the projected point is stored with.
HEIGHT_RTOL, HEIGHT_ATOL = 1e-9, 1e-12
POSE_RAD, POSE_M = 1e-5, 1e-4


def _rot_err(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def _both(oracle_mod, target, source, **kw):
    """A Context and an Oracle with identical parameters and clouds."""
    okw = {k: v for k, v in kw.items()}
    ctx = plo.Context(0, plo.default_params(**kw))
    orc = oracle_mod.Oracle(oracle_mod.default_params(**okw))
    ctx.set_target(target)
    ctx.set_source(source)
    orc.set_target(target)
    orc.set_source(source)
    return ctx, orc


def _check_projection(ctx, orc, T=None, check_pairs=True):
    T = np.eye(4) if T is None else T
    st = ctx.project(T, hooks=True)
    g = {**ctx.neighbors(), **ctx.query_results(), **ctx.pairs()}
    o = orc.project(T, hooks=True)
    assert st["n_source"] == orc.n_source
    assert np.array_equal(g["nn_idx"], o["nn_idx"]), f"{(g['nn_idx'] != o['nn_idx']).sum()} neighbour index mismatches"
    assert np.array_equal(g["nn_d2"], o["nn_d2"])
    assert np.array_equal(g["nn1_idx"], o["nn1_idx"])
    assert np.array_equal(g["nn1_d2"], o["nn1_d2"])
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(st["counters"], o["counters"])
    assert st["n_pairs"] == o["n"] == g["n"]
    ok = o["status"] == 0
    if ok.any():
        err = np.abs(g["height"][ok] - o["height"][ok])
        assert (err <= HEIGHT_RTOL * np.abs(o["height"][ok]) + HEIGHT_ATOL).all()
    if check_pairs and o["n"]:
        assert np.array_equal(g["src_idx"], o["src_idx"])
        assert np.array_equal(g["src_xyz"], o["src_xyz"])            # float32 transform round trip is bit-exact
        assert np.array_equal(g["ref_n"], o["ref_n"])
        assert np.abs(g["ref_xyz"].astype(np.float64) - o["ref_xyz"]).max() <= 2e-6   # 1 ulp(float32) of |y| <= 16
    return g, o


def test_library_loads_and_reports_device():
    ctx = plo.Context(0)
    assert ctx.launch_count == 0
    ctx.close()


def test_projection_parity_hdl64(oracle_mod):
    pair = W.hdl64_pair(max_source=20000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    assert ctx.n_target == orc.n_target and ctx.n_source == orc.n_source
    g, o = _check_projection(ctx, orc)
    assert o["n"] > 15000
    _check_projection(ctx, orc, T=pair.T_gt)
    # different k / radius
    ctx2, orc2 = _both(oracle_mod, pair.target[::3], pair.source[::10], search_number=32, r=1.5, h=0.7)
    _check_projection(ctx2, orc2)
    ctx3, orc3 = _both(oracle_mod, pair.target[::3], pair.source[::10], search_number=5, r=0.4, h=0.3)
    _check_projection(ctx3, orc3)


def test_projection_parity_small_and_ragged(oracle_mod):
    rng = np.random.default_rng(21)
    for n_t, n_s in ((1, 1), (2, 5), (31, 7), (32, 32), (33, 64), (1024, 100), (1025, 100), (5000, 333)):
        tgt = np.zeros((n_t, 12), np.float32)
        tgt[:, 0:3] = rng.uniform(-2, 2, size=(n_t, 3))
        tgt[:, 4:7] = [0, 0, 1]
        src = np.zeros((n_s, 12), np.float32)
        src[:, 0:3] = rng.uniform(-2, 2, size=(n_s, 3))
        src[:, 4:7] = [0, 0, 1]
        ctx, orc = _both(oracle_mod, tgt, src)
        _check_projection(ctx, orc)


def test_empty_inputs(oracle_mod):
    rng = np.random.default_rng(22)
    tgt = np.zeros((100, 12), np.float32)
    tgt[:, 0:3] = rng.uniform(-2, 2, size=(100, 3))
    tgt[:, 6] = 1
    src = tgt[:10].copy()
    ctx, orc = _both(oracle_mod, tgt[:0], src)
    st = ctx.project(np.eye(4), hooks=True)
    assert st["n_pairs"] == 0 and st["counters"][0] == 10          # every query: no 1-NN
    T, rs = ctx.register()
    assert rs["status"] == 3 and rs["iters"] == 0 and np.array_equal(T, np.eye(4))
    ctx2, orc2 = _both(oracle_mod, tgt, src[:0])
    st = ctx2.project(np.eye(4))
    assert st["n_source"] == 0 and st["n_pairs"] == 0 and ctx2.pairs()["n"] == 0
    T, rs = ctx2.register()
    assert rs["status"] == 3
    # all-NaN clouds strip to empty
    bad = tgt.copy()
    bad[:, 0] = np.nan
    ctx3, _ = _both(oracle_mod, bad, src)
    assert ctx3.n_target == 0
    assert ctx3.project(np.eye(4))["counters"][0] == 10


def test_nonfinite_points_are_stripped_and_reindexed(oracle_mod):
    pair = W.hdl64_pair(max_source=3000, max_target=20000)
    tgt, src = pair.target.copy(), pair.source.copy()
    rng = np.random.default_rng(23)
    tgt[rng.choice(tgt.shape[0], 500, replace=False), rng.integers(0, 3, 500)] = np.nan
    tgt[rng.choice(tgt.shape[0], 100, replace=False), 1] = np.inf
    tgt[rng.choice(tgt.shape[0], 300, replace=False), 5] = np.nan      # NaN normals survive the strip
    src[rng.choice(src.shape[0], 200, replace=False), 2] = -np.inf
    ctx, orc = _both(oracle_mod, tgt, src)
    assert ctx.n_target == orc.n_target < tgt.shape[0]
    assert ctx.n_source == orc.n_source == src.shape[0] - 200
    g, o = _check_projection(ctx, orc)
    assert o["counters"][2] > 0                                        # invalid_normal drops happen


def test_ties_duplicates_and_self_match(oracle_mod):
    # grid-aligned points: massive exact-distance ties, resolved by index (D3)
    gr = np.arange(-4, 5, dtype=np.float32) * 0.5
    xyz = np.stack(np.meshgrid(gr, gr, gr, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(24)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    tgt = np.zeros((xyz.shape[0], 12), np.float32)
    tgt[:, 0:3] = xyz
    tgt[:, 6] = 1
    src = np.zeros((300, 12), np.float32)
    src[:, 0:3] = (rng.integers(-8, 9, size=(300, 3)) * 0.25).astype(np.float32)   # on / between grid points
    src[:, 6] = 1
    ctx, orc = _both(oracle_mod, tgt, src, r=1.2, h=1.0)
    _check_projection(ctx, orc)
    # coincident points: the k-list fills up with d2 == 0 entries, the 1-NN needs the second search
    dup = np.zeros((60, 12), np.float32)
    dup[:, 6] = 1
    dup[:40, 0:3] = [1, 1, 1]
    dup[40:, 0:3] = rng.uniform(1.2, 2, size=(20, 3))
    q = np.zeros((3, 12), np.float32)
    q[:, 6] = 1
    q[0, 0:3] = [1, 1, 1]
    q[1, 0:3] = [1, 1, 1.0000001]
    q[2, 0:3] = [5, 5, 5]
    ctx2, orc2 = _both(oracle_mod, dup, q)
    g, o = _check_projection(ctx2, orc2)
    assert o["nn1_idx"][0] >= 40


def test_massive_ties_fill_the_candidate_buffer(oracle_mod):
    # hundreds of coincident / equidistant points around a query overflow the per-warp candidate
    # buffer: exercises the bound-shrink and the exact tie-resolution paths of the collect phase
    rng = np.random.default_rng(1)
    tgt = np.zeros((3000, 12), np.float32)
    tgt[:, 6] = 1
    tgt[:300, 0:3] = [1, 1, 1]
    th = rng.uniform(0, 2 * np.pi, 700)
    tgt[300:1000, 0] = 1 + 0.5 * np.cos(th)
    tgt[300:1000, 1] = 1 + 0.5 * np.sin(th)
    tgt[300:1000, 2] = 1
    tgt[1000:, 0:3] = rng.uniform(-3, 5, size=(2000, 3))
    src = np.zeros((4, 12), np.float32)
    src[:, 6] = 1
    src[0, 0:3] = [1, 1, 1]
    src[1, 0:3] = [1, 1, 1.001]
    src[2, 0:3] = [1.2, 1, 1]
    for k in (20, 32, 3):
        ctx, orc = _both(oracle_mod, tgt, src, search_number=k)
        _check_projection(ctx, orc)
        assert (ctx.search_stats()[:, 2] >= 100000).any()      # at least one query shrank its buffer


def test_each_drop_reason_on_gpu(oracle_mod):
    rng = np.random.default_rng(11)
    tgt = np.zeros((3000, 12), np.float32)
    tgt[:, 0:2] = rng.uniform(-4, 4, size=(3000, 2))
    tgt[:, 6] = 1
    tgt[0, 0:3] = [50, 50, 0]
    tgt[0, 4:7] = [np.nan, 0, 1]
    tgt[1, 0:3] = [80, 80, 0]
    tgt[2:5, 0:3] = [90, 90, 0]
    src = np.zeros((8, 12), np.float32)
    src[:, 6] = 1
    src[0, 0:3] = [0.1, 0.2, 0.05]
    src[1, 0:3] = [200, 200, 0]
    src[2, 0:3] = [0, 0, 2]
    src[3, 0:3] = [50, 50, 0.1]
    src[4, 0:3] = [0.3, 0.1, 0.05]
    src[4, 4:7] = [1, 0, 0]
    src[5, 0:3] = [80, 80, 0.1]
    src[6, 0:3] = [90, 90, 0]
    src[7, 0:3] = [0.5, 0.5, 0.01]
    src[7, 4:7] = 0                                                    # zero normal: NaN angle => kept
    ctx, orc = _both(oracle_mod, tgt, src)
    g, o = _check_projection(ctx, orc)
    assert list(o["status"]) == [0, 1, 2, 3, 4, 5, 1, 0]
    ctx2, orc2 = _both(oracle_mod, tgt, src, normal_angle_constraint=0)
    _check_projection(ctx2, orc2)
    ctx3, orc3 = _both(oracle_mod, tgt, src, transform_normal=1)
    _check_projection(ctx3, orc3, T=plo.synth.scenes.pose_matrix([0.1, 0, 0], yaw_deg=40))


def test_planetary_sparse_large_h(oracle_mod):
    pair = W.planetary_pair()
    for h in (1.0, 2.0, 3.0):
        ctx, orc = _both(oracle_mod, pair.target, pair.source, h=h, r=3 * h)
        g, o = _check_projection(ctx, orc)
    assert o["counters"].sum() > 0


def test_pca_normals_mode(oracle_mod):
    pair = W.hdl64_pair(max_source=3000, max_target=30000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source, is_get_normals=0)
    tn, on = ctx.target_normals(), orc.target_normals()
    fin = np.isfinite(on).all(axis=1)
    assert np.array_equal(fin, np.isfinite(tn).all(axis=1))
    assert np.abs(tn[fin] - on[fin]).max() < 1e-7
    st = ctx.project(np.eye(4), hooks=True)
    g = {**ctx.neighbors(), **ctx.query_results()}
    o = orc.project(np.eye(4), hooks=True)
    assert np.array_equal(g["nn_idx"], o["nn_idx"])
    # PCA normals agree to ~1e-12 only, so a status may flip exactly at the 30 deg gate; none does here
    assert (g["status"] != o["status"]).sum() == 0
    ok = o["status"] == 0
    assert (np.abs(g["height"][ok] - o["height"][ok]) / np.maximum(np.abs(o["height"][ok]), 1e-12)).max() < 1e-6


def test_normal_equations_and_solve(oracle_mod):
    pair = W.hdl64_pair(max_source=30000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    ctx.project(np.eye(4))
    pr = ctx.pairs()
    delta, rank = ctx.solve_wls()
    H, g, sw, sbb, cnt = ctx.normal_equations()
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    Ho, go, swo, sbbo = oracle_mod.normal_equations(s, d, n)
    assert cnt == pr["n"] and rank == 6
    assert np.allclose(H, Ho, rtol=1e-12, atol=0) and np.allclose(g, go, rtol=1e-11, atol=1e-13)
    assert np.isclose(sw, swo, rtol=1e-15) and np.isclose(sbb, sbbo, rtol=1e-12)
    Do = oracle_mod.solve_wls(s, d, n)
    assert np.abs(delta - Do).max() < 1e-10
    # reference-shaped entry point (host vectors in, 4x4 out) incl. caller weights
    ok, D2 = plo.SolveMotionEstimationProblemWeightedLS_CUDA(s, d, n, None, ctx=ctx)
    assert ok and np.abs(D2 - Do).max() < 1e-10
    w = np.random.default_rng(1).uniform(0.1, 1.0, size=s.shape[0])
    ok, D3 = plo.SolveMotionEstimationProblemWeightedLS_CUDA(s, d, n, w, ctx=ctx)
    assert np.abs(D3 - oracle_mod.solve_wls(s, d, n, w)).max() < 1e-10
    # bitwise run-to-run reproducibility of the reduction
    ctx.project(np.eye(4))
    delta_b, _ = ctx.solve_wls()
    assert np.array_equal(delta, delta_b)


def test_solve_failed_status_when_every_row_is_zero():
    """All target normals zero: every pair is kept (finite normal, IMLS height 0) but its row of A is zero, so the
    6x6 system has no pivot.  The reference's WeightedLS would hand a zero / NaN step to its remaining iterations
    (src/laser_odometry.cpp:611-616 only breaks on `false`); the resident loop ends at once with PLO_REG_SOLVE_FAILED
    and the pose it started from -- on the loop kernel, the graph form and the enqueue-all form alike."""
    rng = np.random.default_rng(31)
    tgt = np.zeros((4000, 12), np.float32)
    tgt[:, 0:3] = rng.uniform(-3, 3, size=(4000, 3)) * [1, 1, 0.02]
    src = tgt[:600].copy()
    src[:, 0:3] += np.float32(0.01)
    src[:, 6] = 1
    for knobs in ({}, {"loop_kernel": 0}, {"no_graph": 1}):
        ctx = plo.Context(0)
        for k, v in knobs.items():
            ctx.set_tuning(k, v)
        ctx.set_target(tgt)
        ctx.set_source(src)
        st = ctx.project(np.eye(4))
        assert st["n_pairs"] >= 100
        delta, rank = ctx.solve_wls()
        assert rank == 0 and np.array_equal(delta, np.eye(4))
        T, rs = ctx.register()
        assert rs["status"] == 4 and rs["status_name"] == "SOLVE_FAILED" and rs["iters"] == 0, (knobs, rs)
        assert np.array_equal(T, np.eye(4))
        ctx.close()


def test_rank_deficient_plane(oracle_mod):
    rng = np.random.default_rng(13)
    tgt = np.zeros((5000, 12), np.float32)
    tgt[:, 0:2] = rng.uniform(-6, 6, size=(5000, 2))
    tgt[:, 2] = -1.5
    tgt[:, 6] = 1
    src = np.zeros((800, 12), np.float32)
    src[:, 0:2] = rng.uniform(-3, 3, size=(800, 2))
    src[:, 2] = -1.5
    src[:, 6] = 1
    T = plo.synth.scenes.pose_matrix([0, 0, 0.05], pitch_deg=0.5, roll_deg=-0.4)
    Ti = np.linalg.inv(T)
    src[:, 0:3] = (src[:, 0:3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    ctx, orc = _both(oracle_mod, tgt, src)
    ctx.project(np.eye(4))
    delta, rank = ctx.solve_wls()
    assert rank == 3                                     # same rank decision as Eigen's nonzeroPivots()
    pr = ctx.pairs()
    Do = oracle_mod.solve_wls(*(pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n")))
    assert np.abs(delta - Do).max() < 1e-10
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["status"] == so["status"] == 1 and sg["iters"] == so["iters"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M


@pytest.mark.parametrize("kw", [{}, {"weight_mode": 1}, {"iterations": 2}, {"search_number": 12, "h": 0.8, "r": 2.0}])
def test_register_parity(oracle_mod, kw):
    pair = W.hdl64_pair(max_source=25000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source, **kw)
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"]
    assert sg["pairs"] == so["pairs"]
    assert np.array_equal(sg["counters"], so["counters"])
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD
    assert np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    assert abs(sg["rms"] - so["rms"]) < 1e-9 or kw.get("weight_mode") == 1
    if not kw:
        assert np.linalg.norm(Tg[:3, 3] - pair.T_gt[:3, 3]) < 0.02     # and it is the right answer
        # non-identity start, resident loop vs oracle
        T0 = plo.synth.scenes.pose_matrix([0.5, 0, 0], yaw_deg=1.0)
        Tg2, sg2 = ctx.register(T0)
        To2, so2 = orc.register(T0)
        assert sg2["iters"] == so2["iters"]
        assert _rot_err(Tg2[:3, :3], To2[:3, :3]) < POSE_RAD and np.linalg.norm(Tg2[:3, 3] - To2[:3, 3]) < POSE_M


def test_known_answer_rigid_copy(oracle_mod):
    pair = W.rigid_copy_pair(n=6000)
    ctx = plo.Context(0)
    ctx.set_target(pair.target)
    ctx.set_source(pair.source)
    T, st = ctx.register()
    assert st["status"] == 1
    assert _rot_err(T[:3, :3], pair.T_gt[:3, :3]) < 2e-5
    assert np.linalg.norm(T[:3, 3] - pair.T_gt[:3, 3]) < 2e-4


def test_device_resident_input_and_batch_are_bitwise_identical(oracle_mod):
    import torch
    pair = W.hdl64_pair(max_source=8000, max_target=50000)
    pair2 = W.planetary_pair()
    ctx = plo.Context(0)
    ctx.set_target(pair.target)
    ctx.set_source(pair.source)
    T_host, s_host = ctx.register()
    tt = torch.from_numpy(pair.target).cuda()
    ts = torch.from_numpy(pair.source).cuda()
    torch.cuda.synchronize()
    ctx.set_target(tt)
    ctx.set_source(ts)
    T_dev, s_dev = ctx.register()
    assert np.array_equal(T_host, T_dev) and s_host["iters"] == s_dev["iters"]
    # batched call: same bits as one-by-one, host and device inputs
    ctx.set_target(pair2.target)
    ctx.set_source(pair2.source)
    T2, _ = ctx.register()
    Tb, sb = ctx.register_batch([pair.source, pair2.source, pair.source[:0]], [pair.target, pair2.target, pair.target])
    assert np.array_equal(Tb[0], T_host) and np.array_equal(Tb[1], T2)
    assert sb[2]["status"] == 3 and np.array_equal(Tb[2], np.eye(4))
    Tb2, _ = ctx.register_batch([ts, torch.from_numpy(pair2.source).cuda()], [tt, torch.from_numpy(pair2.target).cuda()])
    assert np.array_equal(Tb2[0], T_host) and np.array_equal(Tb2[1], T2)
    # a second context on a torch stream gives the same bits
    st = torch.cuda.Stream()
    ctx2 = plo.Context(0, stream=st)
    ctx2.set_target(tt)
    ctx2.set_source(ts)
    T3, _ = ctx2.register()
    assert np.array_equal(T3, T_host)


def test_matcher_and_driver_mirror_reference_surface(oracle_mod):
    seq = W.Sequence(seed=2001, n_frames=4, max_points=12000)
    frames = [seq.frame(k) for k in range(4)]
    odo_res = plo.LaserOdometry(resident=True)
    odo_host = plo.LaserOdometry(resident=False)
    P1 = odo_res.run(frames)
    P2 = odo_host.run(frames)
    # the host-stepped loop (one round trip per iteration, like the reference) and the resident loop
    # run the same kernels on the same inputs; only the 4x4 pose product is done by numpy instead
    assert np.abs(P1 - P2).max() < 1e-7
    orc = oracle_mod.Oracle()
    prev = np.eye(4)
    for k in range(1, 4):
        orc.set_target(frames[k - 1])
        orc.set_source(frames[k])
        To, so = orc.register()
        prev = prev @ To
        st = odo_res.frame_stats[k]
        assert st["iters"] == so["iters"] and st["status"] == so["status"]
        assert _rot_err(P1[k][:3, :3], prev[:3, :3]) < POSE_RAD * k and np.linalg.norm(P1[k][:3, 3] - prev[:3, 3]) < POSE_M * k
        gt = seq.relative_gt(k)
        assert np.linalg.norm(st["rPose"][:3, 3] - gt[:3, 3]) < 0.05
    # IMLSICPMatcher: reference method names / argument order (include/imls_icp.h:56-88)
    m = plo.IMLSICPMatcher()
    m.setSourcePointCloud(frames[1])
    m.setTargetPointCloud(frames[0])
    m.setParameters(30, 1, 3, 1, 0.8, False, True, False, 50, 0.2, 0.6, 10, 20, True, 30, "")
    out = m.ProjSourcePtToSurface(np.eye(4))
    orc.set_target(frames[0])
    orc.set_source(frames[1])
    o = orc.project(np.eye(4))
    assert np.array_equal(out["src_idx"], o["src_idx"]) and np.array_equal(out["counters"], o["counters"])
    ok, T, cov, st = m.Match()
    assert ok and np.array_equal(cov, np.eye(4))
    with pytest.raises(plo.PloError):
        m.setParameters(30, 1, 3, 1, 0.8, True, True, False, 50, 0.2, 0.6, 10, 20, True, 30, "")   # tensor voting
    with pytest.raises(plo.PloError):
        plo.Context(0, plo.default_params(search_number=40))                                       # k > 32 unsupported


def test_full_size_north_star_properties(oracle_mod):
    """BASELINE sizes (HDL-64 frame vs 1.0 M-point map): iteration-0 neighbour sets against the
    oracle for every query, then size-independent properties of the resident loop."""
    pair = W.hdl64_vs_map()
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    g, o = _check_projection(ctx, orc)
    Tg, sg = ctx.register()
    assert sg["status"] == 1
    assert _rot_err(Tg[:3, :3], pair.T_gt[:3, :3]) < 2e-3 and np.linalg.norm(Tg[:3, 3] - pair.T_gt[:3, 3]) < 0.02
    # idempotence: restarting from the converged pose stops after one solve and stays put
    Tg2, sg2 = ctx.register(Tg)
    assert sg2["iters"] == 1 and _rot_err(Tg2[:3, :3], Tg[:3, :3]) < 2e-4 and np.linalg.norm(Tg2[:3, 3] - Tg[:3, 3]) < 1e-3
    # permutation invariance of the map (index build must not depend on input order): same neighbour sets
    perm = np.random.default_rng(5).permutation(pair.target.shape[0])
    ctx.set_target(pair.target[perm])
    ctx.project(np.eye(4), hooks=True)
    nb = ctx.neighbors()
    inv = perm                                            # new index i  <-> old index perm[i]
    mapped = np.where(nb["nn_idx"] >= 0, inv[np.maximum(nb["nn_idx"], 0)], -1)
    assert np.array_equal(np.sort(mapped, axis=1), np.sort(g["nn_idx"], axis=1))
    assert np.array_equal(nb["nn_d2"], g["nn_d2"])
    To, so = orc.register()
    assert sg["iters"] == so["iters"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M


def test_cpp_host_driver_matches_python_driver(tmp_path):
    """The C++ adapter (include/plo/imls_icp_cuda.h) + driver loop restated in C++
    (host/odometry_driver.cpp, the shape of src/laser_odometry.cpp:416-683) gives the poses of
    the Python mirror, in both the resident and the reference-style stepped mode."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "planetary-lidar-odometry_b200", "host", "odometry_driver")
    assert os.path.exists(exe), "build the host driver (python -c 'import __graft_entry__ as g; g.build()')"
    seq = W.Sequence(seed=2002, n_frames=3, max_points=10000)
    frames = [seq.frame(k) for k in range(3)]
    files = []
    for k, f in enumerate(frames):
        fn = str(tmp_path / f"frame{k}.bin")
        f.astype(np.float32).tofile(fn)
        files.append(fn)
    P = plo.LaserOdometry(resident=True).run(frames)
    for mode in ("resident", "stepped"):
        out = str(tmp_path / f"poses_{mode}.txt")
        subprocess.check_call([exe, "-", mode, out] + files)
        got = np.loadtxt(out)
        assert got.shape == (3, 8)
        assert np.abs(got[:, 1:4] - P[:, :3, 3]).max() < 2e-6      # 6 decimals in the TUM file
    # the reference's own config.json layout with its default solver chain strings (RANSAC -> DRPM / Weighted LS);
    # "Ceres.max_iterations" precedes "RANSAC.max_iterations" as in the reference file
    import json
    for final in ("Weighted LS", "DRPM"):
        cfg = plo.config.load_config()
        sm = cfg["laser_odometry"]["solve_method"]
        sm_new = {"_comment": "choose method: Ceres, LS, RANSAC, ICP, Teaser", "method": "RANSAC"}
        sm_new.update({k: v for k, v in sm.items() if k not in ("method", "RANSAC", "LS")})
        sm_new["Ceres"] = {"max_iterations": 20}
        sm_new["LS"] = {"threshold": 0.02}
        sm_new["RANSAC"] = {"max_iterations": 5000, "distance_threshold": 0.8, "min_inliers_percentage": 0.95,
                            "huber_threshold": 0.648, "final_solve_method": final, "LS_threshold": 0.02,
                            "DRPM_threshold": 0.05, "DRPM_stdev_points": 0.02, "DRPM_stdev_normals": 0.05}
        cfg["laser_odometry"]["solve_method"] = sm_new
        cfn = str(tmp_path / "config.json")
        with open(cfn, "w", encoding="utf-8") as f:
            json.dump(cfg, f, indent=4)
        Pr = plo.LaserOdometry(plo.config.load_config(cfn), resident=True).run(frames)
        for mode in ("resident", "stepped"):
            out = str(tmp_path / f"poses_ransac_{mode}.txt")
            subprocess.check_call([exe, cfn, mode, out] + files)
            got = np.loadtxt(out)
            assert np.abs(got[:, 1:4] - Pr[:, :3, 3]).max() < 2e-6
        assert np.abs(Pr - P).max() > 1e-9      # a different solver really ran


def test_cfg2_vlp32c_sequence_full_size(oracle_mod):
    """BASELINE config 2 (VLP-32C-shaped sequence, IMLS matcher + weighted LS end to end), first frames at
    full size: per-frame pose parity with the oracle and closeness to the ground-truth motion."""
    seq = W.Sequence(seed=2001, n_frames=4)
    frames = [seq.frame(k) for k in range(4)]
    assert 40000 < frames[0].shape[0] < 70000
    odo = plo.LaserOdometry(resident=True)
    odo.run(frames)
    orc = oracle_mod.Oracle()
    for k in range(1, 4):
        orc.set_target(frames[k - 1])
        orc.set_source(frames[k])
        To, so = orc.register()
        st = odo.frame_stats[k]
        assert st["iters"] == so["iters"] and st["status"] == so["status"] and st["pairs"] == so["pairs"]
        assert _rot_err(st["rPose"][:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(st["rPose"][:3, 3] - To[:3, 3]) < POSE_M
        gt = seq.relative_gt(k)
        assert np.linalg.norm(st["rPose"][:3, 3] - gt[:3, 3]) < 0.05 and _rot_err(st["rPose"][:3, :3], gt[:3, :3]) < 5e-3


def test_cfg2_sequence_201_frames_ate_tum_and_oracle_parity(oracle_mod, tmp_path):
    """BASELINE config 2 at length: 201 full-size VLP-32C frames (200 registrations) through LaserOdometry on the
    resident path.  Checked: (i) pose chaining nowPose = prevLaserPose * rPose (src/laser_odometry.cpp:649-655) over
    200 products, bit for bit against a host-side product of the per-frame rPose; (ii) oracle pose / iteration /
    pair-count parity on every 20th frame; (iii) every frame converges and stays close to the generator's ground-truth
    motion, trajectory ATE (RMSE of positions, frame-0 coordinates) bounded; (iv) the TUM file of savePoseToFile
    (src/saver.cpp:46-54: `timestamp tx ty tz qx qy qz qw`, 6 decimals) round-trips the trajectory."""
    n = 201
    fs = W.generate_sequences([2001], n)[0]
    frames = [fs.frame(k) for k in range(n)]
    assert all(40000 < f.shape[0] < 70000 for f in frames)
    odo = plo.LaserOdometry(resident=True)
    P = odo.run(frames)
    assert P.shape == (n, 4, 4)
    # (i) chaining
    cur = np.eye(4)
    for k in range(1, n):
        cur = cur @ odo.frame_stats[k]["rPose"]
        assert np.array_equal(cur, P[k])
    # (ii) oracle parity on every 20th registration
    orc = oracle_mod.Oracle()
    for k in range(20, n, 20):
        orc.set_target(frames[k - 1])
        orc.set_source(frames[k])
        To, so = orc.register()
        st = odo.frame_stats[k]
        assert st["iters"] == so["iters"] and st["status"] == so["status"] and st["pairs"] == so["pairs"], k
        assert _rot_err(st["rPose"][:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(st["rPose"][:3, 3] - To[:3, 3]) < POSE_M, k
    # (iii) ground truth
    gt = np.stack([np.linalg.inv(fs.poses[0]) @ T for T in fs.poses])
    for k in range(1, n):
        st = odo.frame_stats[k]
        assert st["status"] == 1, (k, st["status_name"])
        rel = fs.relative_gt(k)
        assert np.linalg.norm(st["rPose"][:3, 3] - rel[:3, 3]) < 0.05 and _rot_err(st["rPose"][:3, :3], rel[:3, :3]) < 5e-3, k
    ate = float(np.sqrt(np.mean(np.sum((P[:, :3, 3] - gt[:, :3, 3]) ** 2, axis=1))))
    path_len = float(np.sum(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1)))
    assert path_len > 100.0
    assert ate < 0.01 * path_len, (ate, path_len)      # < 1 % of the distance travelled
    # (iv) TUM output
    out = str(tmp_path / "laser_odometry_poses.txt")
    plo.save_poses_tum(out, P)
    rows = np.loadtxt(out)
    assert rows.shape == (n, 8)
    lines = open(out, encoding="utf-8").read().splitlines()
    assert all(len(f.split(".")[1]) == 6 for f in lines[7].split(" "))      # 6 decimals, every field
    assert np.abs(rows[:, 1:4] - P[:, :3, 3]).max() <= 5.1e-7
    q = rows[:, 4:8]     # qx qy qz qw -> rotation
    assert np.abs(np.linalg.norm(q, axis=1) - 1.0).max() < 5e-6
    for k in (0, 57, 200):
        x, y, z, w = q[k]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(R - P[k, :3, :3]).max() < 5e-6      # 6-decimal quaternion components


def test_cfg2_sequence_1000_frames_full_length(oracle_mod):
    """BASELINE config 2 at its stated length: 1000 full-size VLP-32C frames (999 registrations) through LaserOdometry on
    the resident path -- every registration converges, poses chain bit for bit, oracle parity (iterations, status, pairs,
    pose) on every 100th registration, ATE below 1 % of the path.  The summary goes to gpurun_out/ when that exists."""
    import json
    import os
    import time
    n = 1000
    fs = W.generate_sequences([2001], n)[0]
    frames = [fs.frame(k) for k in range(n)]
    odo = plo.LaserOdometry(resident=True)
    odo.run(frames[:3])     # allocations, graph instantiation
    odo = plo.LaserOdometry(resident=True)
    t0 = time.perf_counter()
    P = odo.run(frames)
    run_s = time.perf_counter() - t0
    cur = np.eye(4)
    for k in range(1, n):
        st = odo.frame_stats[k]
        assert st["status"] == 1, (k, st["status_name"])
        cur = cur @ st["rPose"]
        assert np.array_equal(cur, P[k]), k
    orc = oracle_mod.Oracle()
    worst_t = worst_r = 0.0
    for k in range(100, n, 100):
        orc.set_target(frames[k - 1])
        orc.set_source(frames[k])
        To, so = orc.register()
        st = odo.frame_stats[k]
        assert st["iters"] == so["iters"] and st["status"] == so["status"] and st["pairs"] == so["pairs"], k
        worst_t = max(worst_t, float(np.linalg.norm(st["rPose"][:3, 3] - To[:3, 3])))
        worst_r = max(worst_r, _rot_err(st["rPose"][:3, :3], To[:3, :3]))
    assert worst_r < POSE_RAD and worst_t < POSE_M
    gt = np.stack([np.linalg.inv(fs.poses[0]) @ T for T in fs.poses])
    err = np.linalg.norm(P[:, :3, 3] - gt[:, :3, 3], axis=1)
    ate = float(np.sqrt(np.mean(err ** 2)))
    path_len = float(np.sum(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1)))
    assert path_len > 500.0 and ate < 0.01 * path_len, (ate, path_len)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        res = {"workload": "BASELINE config 2: 1000-frame VLP-32C sequence, synthetic, seed 2001", "registrations": n - 1,
               "points_per_frame_mean": float(np.mean([f.shape[0] for f in frames])), "wall_s_host_frames_one_at_a_time": run_s,
               "scans_per_s_wall": (n - 1) / run_s, "mean_iters": float(np.mean([odo.frame_stats[k]["iters"] for k in range(1, n)])),
               "path_length_m": path_len, "ate_rmse_m": ate, "end_point_error_m": float(err[-1]),
               "oracle_parity_every_100th": {"max_trans_diff_m": worst_t, "max_rot_diff_rad": worst_r}}
        with open(os.path.join(out_dir, "cfg2_1000_frames.json"), "w", encoding="utf-8") as f:
            f.write(json.dumps(res) + "\n")


def test_cfg4_five_million_point_map(oracle_mod):
    """BASELINE config 4 (HDL-64 frame vs 5 M-point map: index build + radius-search stress, 4 tree levels):
    iteration-0 neighbour sets against the oracle on a query subset, full-loop pose parity, index-size properties."""
    pair = W.hdl64_vs_dense_map()
    assert pair.target.shape[0] == 5_000_000
    sub = pair.source[::8]
    ctx, orc = _both(oracle_mod, pair.target, sub)
    assert ctx.n_target == 5_000_000
    g, o = _check_projection(ctx, orc)
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    # the full frame on the same index: converges to the same place as the subset (size-independent property)
    ctx.set_source(pair.source)
    Tf, sf = ctx.register()
    assert sf["status"] == 1 and np.linalg.norm(Tf[:3, 3] - Tg[:3, 3]) < 5e-3
    assert np.linalg.norm(Tf[:3, 3] - pair.T_gt[:3, 3]) < 0.02


def test_cfg5_batched_sequences_single_rank(oracle_mod):
    """BASELINE config 5 on one rank: the sharded driver (one plo_register_batch per sequence, gather, pose
    chaining) reproduces frame-by-frame registration bit for bit."""
    seqs = [W.Sequence(seed=5000 + s, n_frames=3, max_points=8000) for s in range(3)]
    ctx = plo.Context(0)
    trajs, table = plo.distributed.register_sequences_sharded(ctx, seqs)
    assert table.shape == (9, plo.distributed.RESULT_WIDTH)
    for s, seq in enumerate(seqs):
        odo = plo.LaserOdometry(resident=True)
        P = odo.run([seq.frame(k) for k in range(3)])
        assert np.array_equal(P, trajs[s])


def test_randomised_small_clouds_fuzz(oracle_mod):
    """Randomised parity sweep: cloud sizes around the leaf / level boundaries, clustered and
    grid-quantised coordinates (many exact ties and duplicates), random k / r / h / angle gate,
    sprinkled NaNs and zero normals — neighbour sets, statuses and counters must match the oracle
    exactly for every draw."""
    rng = np.random.default_rng(20240607)
    sizes = [1, 2, 3, 19, 20, 21, 31, 32, 33, 63, 64, 65, 500, 1023, 1024, 1025, 1500, 4000, 33000]
    for trial in range(48):
        n_t = int(rng.choice(sizes))
        n_s = int(rng.choice([1, 5, 33, 200, 700]))
        mode = trial % 4
        if mode == 0:      # uniform
            xyz = rng.uniform(-3, 3, size=(n_t, 3))
        elif mode == 1:    # quantised grid: exact ties and duplicates
            xyz = np.round(rng.uniform(-2, 2, size=(n_t, 3)) * 4) / 4
        elif mode == 2:    # tight clusters + far outliers
            c = rng.uniform(-20, 20, size=(max(1, n_t // 50), 3))
            xyz = c[rng.integers(0, c.shape[0], n_t)] + rng.normal(0, 0.05, size=(n_t, 3))
        else:              # thin surface (plane with noise)
            xyz = np.concatenate([rng.uniform(-5, 5, size=(n_t, 2)), rng.normal(0, 0.01, size=(n_t, 1))], axis=1)
        tgt = np.zeros((n_t, 12), np.float32)
        tgt[:, 0:3] = xyz
        nrm = rng.normal(size=(n_t, 3)) * [0.3, 0.3, 1.0]
        nrm[:, 2] = np.abs(nrm[:, 2])
        tgt[:, 4:7] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        src = np.zeros((n_s, 12), np.float32)
        pick = rng.integers(0, n_t, n_s)
        src[:, 0:3] = xyz[pick] + rng.normal(0, rng.choice([0.0, 0.02, 0.5]), size=(n_s, 3))
        src[:, 4:7] = tgt[pick, 4:7]
        if trial % 5 == 0 and n_t > 3:
            tgt[rng.integers(0, n_t, 2), rng.integers(0, 3, 2)] = np.nan
            tgt[rng.integers(0, n_t), 4:7] = 0
            src[rng.integers(0, n_s), 4:7] = 0
        kw = dict(search_number=int(rng.choice([1, 3, 10, 20, 32])), r=float(rng.choice([0.2, 1.0, 3.0, 10.0])),
                  h=float(rng.choice([0.1, 1.0, 5.0])), angle_diff_threshold=float(rng.choice([5.0, 30.0, 90.0, 200.0])),
                  normal_angle_constraint=int(rng.integers(0, 2)))
        ctx, orc = _both(oracle_mod, tgt, src, **kw)
        T = plo.synth.scenes.pose_matrix(rng.normal(0, 0.05, 3), yaw_deg=rng.normal(0, 2)) if trial % 3 else np.eye(4)
        try:
            _check_projection(ctx, orc, T=T)
            # a second projection at a nearby pose exercises the temporal bound
            T2 = plo.synth.scenes.pose_matrix(rng.normal(0, 0.01, 3), yaw_deg=rng.normal(0, 0.2)) @ T
            _check_projection(ctx, orc, T=T2)
        except AssertionError as e:
            raise AssertionError(f"trial {trial}: n_t={n_t} n_s={n_s} mode={mode} {kw}: {e}") from e
        ctx.close()


def test_carry_bounds_along_the_scan_order(oracle_mod, monkeypatch):
    """Chunks of consecutive source points per warp (PLO_CHUNK forces them on small clouds): the bound of a
    query then comes from the previous query of its chunk -- triangle inequality on its k-th distance and the
    farthest of its k neighbours.  Sorted, shuffled, duplicated and far-apart consecutive queries, quantised
    maps (ties at the bound), lists that are not full: neighbour sets, statuses and counters stay exact."""
    rng = np.random.default_rng(77)
    for chunk in (2, 8, 64):
        monkeypatch.setenv("PLO_CHUNK", str(chunk))
        for trial in range(8):
            n_t = int(rng.choice([40, 700, 5000, 33000]))
            n_s = int(rng.choice([64, 500, 3000]))
            xyz = rng.uniform(-4, 4, size=(n_t, 3)) * [1, 1, 0.02]
            if trial % 2:
                xyz = np.round(xyz * 8) / 8
            tgt = np.zeros((n_t, 12), np.float32)
            tgt[:, 0:3] = xyz
            tgt[:, 4:7] = [0, 0, 1]
            src = np.zeros((n_s, 12), np.float32)
            line = np.linspace(-4, 4, n_s)
            src[:, 0] = line                                  # a scan line: neighbours in memory are neighbours in space
            src[:, 1] = 0.3 * np.sin(line)
            if trial % 4 == 1:
                src[:, 0:3] = xyz[rng.integers(0, n_t, n_s)]      # unrelated consecutive queries, exact hits on map points
            if trial % 4 == 2:
                m3 = src[1::3].shape[0]
                src[0:3 * m3:3, 0:3] = src[1::3, 0:3]             # duplicated consecutive queries
            if trial % 4 == 3:
                src[::7, 0] += 50.0                               # every 7th query far from everything (empty list in between)
            src[:, 4:7] = [0, 0, 1]
            kw = dict(search_number=int(rng.choice([3, 20, 32])), r=float(rng.choice([0.5, 3.0])), h=1.0)
            ctx, orc = _both(oracle_mod, tgt, src, **kw)
            try:
                _check_projection(ctx, orc)
                _check_projection(ctx, orc, T=plo.synth.scenes.pose_matrix([0.3, -0.1, 0.0], yaw_deg=1.0))
            except AssertionError as e:
                raise AssertionError(f"chunk {chunk} trial {trial}: n_t={n_t} n_s={n_s} {kw}: {e}") from e
            ctx.close()


def test_per_launch_projection_times(oracle_mod):
    """plo_last_project_times: one device time per ICP iteration of the last registration (profiling mode)."""
    pair = W.hdl64_pair(max_source=20000)
    ctx = plo.Context(0)
    ctx.set_target(pair.target)
    ctx.set_source(pair.source)
    T0, s0 = ctx.register()
    ctx.set_profiling(True)
    ctx.set_target(pair.target)
    ctx.set_source(pair.source)
    T1, s1 = ctx.register()
    each = ctx.last_project_times()
    mean = ctx.last_kernel_timings()
    ctx.set_profiling(False)
    assert np.array_equal(T0, T1) and s0["iters"] == s1["iters"]          # enqueue-all path == graph path, bitwise
    assert each.shape[0] == mean["n_project"] == s1["iters"] and (each > 0).all()
    assert abs(each.mean() - mean["ms_project_mean"]) < 1e-4


def test_trimmed_ls_solver(oracle_mod):
    """"next" row of SURVEY.md §8f, rank 2: SolveMotionEstimationProblemLS (src/solver.cpp:74-166) on the
    device — first LS, |residual| rank window [thr*N, (1-thr)*N] by a stable radix sort, second LS."""
    pair = W.hdl64_pair(max_source=30000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source, solver=1)
    ctx.project(np.eye(4))
    pr = ctx.pairs()
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    for thr in (0.02, 0.1, 0.0):
        ctx.set_params(plo.default_params(solver=1, ls_threshold=thr))
        ctx.project(np.eye(4))
        delta, rank = ctx.solve_ls()
        Do = oracle_mod.solve_ls(s, d, n, thr)
        assert rank == 6 and np.abs(delta - Do).max() < 1e-9, thr
    # full loop, resident (CUDA graph) and stepped, against the oracle's LS loop
    ctx, orc = _both(oracle_mod, pair.target, pair.source, solver=1)
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    cfg = plo.config.load_config()
    cfg["laser_odometry"]["solve_method"]["method"] = "LS"
    odo = plo.LaserOdometry(cfg, resident=False)
    odo.process_frame(pair.target)
    _, st = odo.process_frame(pair.source)
    assert st["iters"] == so["iters"] and np.abs(st["rPose"] - Tg).max() < 1e-7
    # and the weighted-LS result differs (the trim really changes the estimate)
    ctx2, _ = _both(oracle_mod, pair.target, pair.source)
    Tw, _ = ctx2.register()
    assert np.abs(Tw - Tg).max() > 1e-7


@pytest.mark.parametrize("final", [0, 1, 2])
def test_ransac_front_and_drpm_tail(oracle_mod, final):
    """"next" row of SURVEY.md §8f, rank 1 — the literal config.json default chain on the device:
    RANSAC (FPS-3 hypotheses seeded by xorshift64 instead of the reference's rand(), inlier count,
    Huber/exp weights at the best hypothesis) then trimmed LS on the inliers (final=0, src/solver.cpp:366-371),
    Weighted LS (final=1) or DRPM (final=2)."""
    pair = W.hdl64_pair(max_source=20000)
    kw = dict(solver=2, ransac_final=final)
    ctx, orc = _both(oracle_mod, pair.target, pair.source, **kw)
    ctx.project(np.eye(4))
    pr = ctx.pairs()
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    delta, info = ctx.solve_ransac()
    ok, Do = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(**kw))
    assert ok and np.abs(delta - Do).max() < 1e-9
    assert 1 <= info["hypotheses"] <= 5000 and info["inliers"] > 0.95 * pr["n"]
    if final == 2:
        # probabilities against the oracle's DRPM on the same inliers / weights
        idx, w = oracle_mod.ransac_weights(s, d, n)       # T_best differs, but with > 95 % inliers the sets coincide here
        assert (info["probs"] >= 0).all() and (info["probs"] <= 1).all()
    # harder sampling: few iterations allowed, high inlier bar -> several hypotheses are evaluated
    kw2 = dict(solver=2, ransac_final=final, ransac_min_inliers_percentage=0.9999, ransac_max_iterations=7, ransac_seed=12345,
               ransac_distance_threshold=0.05)
    ctx.set_params(plo.default_params(**kw2))
    ctx.project(np.eye(4))
    delta2, info2 = ctx.solve_ransac()
    ok, Do2 = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(**kw2))
    assert info2["hypotheses"] == 7 and np.abs(delta2 - Do2).max() < 1e-9
    # the exit test never fires: 700 hypotheses, spread over the whole GPU (one per block, k_ransac_eval), must give what
    # the reference's one-after-the-other loop gives; then an inlier bar only the best of them passes: drawing stops there
    kw3 = dict(kw2, ransac_min_inliers_percentage=1.0, ransac_max_iterations=700)
    ctx.set_params(plo.default_params(**kw3))
    ctx.project(np.eye(4))
    delta3, info3 = ctx.solve_ransac()
    ok, Do3 = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(**kw3))
    assert info3["hypotheses"] == 700 and np.abs(delta3 - Do3).max() < 1e-9
    kw4 = dict(kw3, ransac_min_inliers_percentage=(info3["inliers"] - 0.5) / pr["n"])
    ctx.set_params(plo.default_params(**kw4))
    ctx.project(np.eye(4))
    delta4, info4 = ctx.solve_ransac()
    ok, Do4 = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(**kw4))
    assert 1 <= info4["hypotheses"] <= 700 and info4["inliers"] == info3["inliers"]
    assert np.abs(delta4 - Do4).max() < 1e-9 and np.array_equal(delta4, delta3)
    # full loop (resident graph) against the oracle's loop with the same solver chain
    ctx3, orc3 = _both(oracle_mod, pair.target, pair.source, **kw)
    Tg, sg = ctx3.register()
    To, so = orc3.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    # stepped driver with the reference's own config.json strings
    cfg = plo.config.load_config()
    cfg["laser_odometry"]["solve_method"]["method"] = "RANSAC"
    cfg["laser_odometry"]["solve_method"]["RANSAC"]["final_solve_method"] = ["LS", "Weighted LS", "DRPM"][final]
    odo = plo.LaserOdometry(cfg, resident=False)
    odo.process_frame(pair.target)
    _, st = odo.process_frame(pair.source)
    assert st["iters"] == so["iters"] and np.abs(st["rPose"] - Tg).max() < 1e-7


def _conditioned_pairs(eps, n=20000, seed=5):
    """Point-to-plane system whose cond(A) is ~ 6 / eps: points in a 20 m box, normals e_z + eps * (in-plane noise),
    i.e. nearly flat terrain on which x / y translation and yaw are only as observable as the normals tilt."""
    rng = np.random.default_rng(seed)
    s = rng.uniform(-10, 10, size=(n, 3)) * [1, 1, 0.05]
    nrm = np.array([0, 0, 1.0]) + eps * np.concatenate([rng.normal(size=(n, 2)), np.zeros((n, 1))], axis=1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    A = np.concatenate([np.cross(s, nrm), nrm], axis=1)
    x_true = np.array([2e-3, -1e-3, 3e-3, 0.05, -0.03, 0.02])
    b = A @ x_true + rng.normal(0, 1e-4, n)
    d = s + nrm * b[:, None]
    return s, d, nrm, A, np.einsum("ij,ij->i", nrm, d - s)


def test_conditioning_sweep_against_qr_and_svd(oracle_mod):
    """cond(A) from 1e2 to 1e8 (flatter and flatter terrain): the device solve against the oracle's column-pivoted
    Householder QR (the reference's ColPivHouseholderQR, src/solver.cpp:200) and against scipy's SVD least squares.
    Rank is 6 throughout (the reference's QR agrees: sigma_min / sigma_max stays far above machine epsilon) and the
    6-vector stays within 1e-9 * |x| of the SVD solution -- the bar a QR of A meets, not one a normal-equation solve
    does."""
    import scipy.linalg
    ctx = plo.Context(0)
    for eps in (1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7):
        s, d, nrm, A, b = _conditioned_pairs(eps)
        sv = np.linalg.svd(A, compute_uv=False)
        cond = sv[0] / sv[-1]
        assert 2 / eps < cond < 20 / eps
        x_svd = scipy.linalg.lstsq(A, b, lapack_driver="gelsd")[0]
        Dg, rank = ctx.solve_wls_host(s, d, nrm)
        Do = oracle_mod.solve_wls(s, d, nrm)
        assert rank == 6, (eps, rank)
        # 4x4 deltas: translation is x[3:6]; the rotation block is exp([x[0:3]]x)
        tol = 1e-9 * np.linalg.norm(x_svd) + 1e-13
        assert np.abs(Dg[:3, 3] - x_svd[3:6]).max() < tol, (eps, cond, np.abs(Dg[:3, 3] - x_svd[3:6]).max())
        assert np.abs(Dg - Do).max() < tol, (eps, cond, np.abs(Dg - Do).max())
        R = scipy.linalg.expm(np.array([[0, -x_svd[2], x_svd[1]], [x_svd[2], 0, -x_svd[0]], [-x_svd[1], x_svd[0], 0]]))
        assert np.abs(Dg[:3, :3] - R).max() < tol
    # exact rank deficiency: perfectly flat, parallel normals -> 3 observable directions, both sides say so
    s, d, nrm, A, b = _conditioned_pairs(0.0)
    Dg, rank = ctx.solve_wls_host(s, d, nrm)
    assert rank == 3 and np.abs(Dg - oracle_mod.solve_wls(s, d, nrm)).max() < 1e-9


def test_planetary_full_loop_pose_parity_large_h(oracle_mod):
    """BASELINE config 3 (sparse VLP-16 planetary terrain, neighbour-starved), h in {1, 2, 3} with r = 3h: the whole
    resident loop against the oracle -- status, iterations, pairs, drop counters of the last projection, pose."""
    pair = W.planetary_pair()
    for h in (1.0, 2.0, 3.0):
        ctx, orc = _both(oracle_mod, pair.target, pair.source, h=h, r=3 * h)
        Tg, sg = ctx.register()
        To, so = orc.register()
        assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"], h
        assert np.array_equal(sg["counters"], so["counters"]), h
        assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M, h
        assert sg["counters"].sum() > 0.02 * pair.source.shape[0]          # the neighbour-starved path really drops points
        ctx.close()


def test_pca_normals_full_cfg1_pair(oracle_mod):
    """a8 at a BASELINE size: the full cfg-1 HDL-64 pair with get_normals.enabled = false (ComputeNormal for every map
    point, src/imls_icp.cpp:753-794).  The device eigen-solver (cyclic Jacobi) and the oracle's are different
    algorithms, both backward stable: their unit eigenvectors differ by at most ~ eps * |C| / gap, gap = lambda_2 -
    lambda_1 of the 10-point covariance.  Asserted: that bound for every normal; every height within the stated 1e-9
    relative bar wherever the normals a query touches agree to 1e-12, and within the bound-propagated error elsewhere;
    statuses / counters exact."""
    pair = W.hdl64_pair()
    kw = dict(is_get_normals=0)
    ctx, orc = _both(oracle_mod, pair.target, pair.source, **kw)
    st = ctx.project(np.eye(4), hooks=True)
    g = {**ctx.neighbors(), **ctx.query_results()}
    o = orc.project(np.eye(4), hooks=True)
    ng, no = ctx.target_normals(), orc.target_normals()
    both = np.isfinite(ng).all(axis=1) & np.isfinite(no).all(axis=1)
    assert np.array_equal(np.isfinite(ng).all(axis=1), np.isfinite(no).all(axis=1)) and both.sum() > 100000
    nerr = np.zeros(ng.shape[0])
    nerr[both] = np.abs(ng[both] - no[both]).max(axis=1)
    # eigen-gaps of the covariances the normals come from (scipy kd-tree; points with coincident neighbours are left out)
    import scipy.spatial
    P = pair.target[np.isfinite(pair.target[:, 0:3]).all(axis=1), 0:3].astype(np.float64)
    dist, idx = scipy.spatial.cKDTree(P).query(P, k=11)
    clean = both & (dist[:, 1] > 1e-7) & (dist[:, 10] <= 1.0) & (dist[:, 10] < np.inf)
    nb = P[idx[:, 1:]]
    C = np.einsum("nki,nkj->nij", nb - nb.mean(axis=1, keepdims=True), nb - nb.mean(axis=1, keepdims=True)) / 10.0
    w = np.linalg.eigvalsh(C)
    bound = 64 * np.finfo(np.float64).eps * w[:, 2] / np.maximum(w[:, 1] - w[:, 0], 1e-300) + 1e-15
    assert (nerr[clean] <= bound[clean]).all(), float((nerr[clean] / bound[clean]).max())
    assert np.percentile(nerr[both], 99) < 1e-12
    # neighbour sets, statuses and counters do not depend on the normals' last bits except through the 30-degree gate
    assert np.array_equal(g["nn_idx"], o["nn_idx"]) and np.array_equal(g["nn_d2"], o["nn_d2"])
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(st["counters"], o["counters"])
    ok = o["status"] == 0
    touched = np.where(g["nn_idx"] >= 0, nerr[np.maximum(g["nn_idx"], 0)], 0.0).max(axis=1)   # worst normal a query's neighbours carry
    herr = np.abs(g["height"][ok] - o["height"][ok])
    hbar = HEIGHT_RTOL * np.abs(o["height"][ok]) + HEIGHT_ATOL
    tight = touched[ok] <= 1e-12
    assert tight.mean() > 0.97
    assert (herr[tight] <= hbar[tight]).all()
    # elsewhere the height moves by at most |x - p_j| * |dn| <= r * |dn| (weights sum to one)
    assert (herr[~tight] <= hbar[~tight] + 3.0 * 4 * touched[ok][~tight]).all()


def test_resident_loop_fuzz_sizes_and_long_runs(oracle_mod):
    """k_register_loop (one cooperative launch per registration) over cloud sizes from a handful of points to tens of
    thousands -- one block to the full grid, every group size of the tile path -- and over registrations forced to run all
    30 iterations (convergence thresholds at zero: the tile path answers ~27 projections in a row, misses and refreshes
    included): status, iterations, pairs, drop counters and pose against the oracle; the second run of the same context
    and the graph / enqueue-all forms of the loop give bitwise the same pose."""
    rng = np.random.default_rng(4242)
    base = W.hdl64_pair(max_source=40000, max_target=60000)
    cases = [(7, 300), (33, 2000), (500, 2000), (2000, 20000), (9000, 60000), (40000, 60000)]
    for n_s, n_t in cases:
        src = base.source[np.sort(rng.choice(base.source.shape[0], n_s, replace=False))]
        tgt = base.target[np.sort(rng.choice(base.target.shape[0], n_t, replace=False))]
        for kw in (dict(), dict(delta_dist_threshold=0.0, delta_angle_threshold=0.0)):
            ctx, orc = _both(oracle_mod, tgt, src, **kw)
            Tg, sg = ctx.register()
            To, so = orc.register()
            assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"], (n_s, n_t, kw, sg, so)
            assert np.array_equal(sg["counters"], so["counters"]), (n_s, n_t, kw)
            if so["pairs"] >= 6:
                assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M, (n_s, n_t, kw)
            if kw:
                assert sg["iters"] == 30 or sg["status"] == 3
                if sg["iters"] == 30 and n_s >= 500:
                    assert (ctx.last_tile_misses()[3:] >= 0).all()          # the tile path was in use
            T2, s2 = ctx.register()                                         # same clouds: starts from tiles and temporal bounds
            assert np.array_equal(Tg, T2) and s2["iters"] == sg["iters"]
            for knob in ("loop_kernel", "no_graph"):
                c2 = plo.Context(0, plo.default_params(**kw))
                c2.set_tuning(knob, 0 if knob == "loop_kernel" else 1)
                c2.set_target(tgt)
                c2.set_source(src)
                T3, s3 = c2.register()
                assert np.array_equal(Tg, T3) and s3["iters"] == sg["iters"], (n_s, n_t, kw, knob)
                c2.close()
            ctx.close()


def test_device_inputs_produced_on_another_stream(oracle_mod):
    """torch CUDA tensors are passed by pointer and read on the context's own stream: the Context orders that stream
    behind the torch stream that produced them (an event, no global synchronisation).  The clouds are produced on a
    side stream behind a long-running kernel, with no synchronise before the library call."""
    import torch
    pair = W.hdl64_pair(max_source=20000)
    orc = oracle_mod.Oracle()
    orc.set_target(pair.target)
    orc.set_source(pair.source)
    To, so = orc.register()
    dev = torch.device("cuda", 0)
    h_t = torch.from_numpy(pair.target).pin_memory()
    h_s = torch.from_numpy(pair.source).pin_memory()
    side = torch.cuda.Stream(device=dev)
    ctx = plo.Context(0)
    big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    for rep in range(3):
        with torch.cuda.stream(side):
            for _ in range(8):
                big.normal_()                      # ~ms of work ahead of the uploads on the side stream
            d_t = torch.empty_like(h_t, device=dev).fill_(float("nan"))
            d_s = torch.empty_like(h_s, device=dev).fill_(float("nan"))
            d_t.copy_(h_t, non_blocking=True)
            d_s.copy_(h_s, non_blocking=True)
            ctx.set_target(d_t)                    # current stream = side: the context waits for its event
            ctx.set_source(d_s)
        Tg, sg = ctx.register()
        assert sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"], rep
        assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    # pinned host buffers: the upload runs on the copy stream while the index is built; results are the same
    ctx.set_target(h_t)
    ctx.set_source(h_s)
    Th, sh = ctx.register()
    assert np.array_equal(Th, Tg) and sh["iters"] == sg["iters"]


def test_host_vector_solver_entry_points(oracle_mod):
    """SolveMotionEstimationProblem{LS,RANSAC,DRPM} with the reference's shape (include/solver.h:84-90, :100-114,
    :129-139): host vectors in, 4x4 out, against the oracle's restatements and against the same solvers run on the
    device-resident pairs; the dispatcher accepts the reference's strings; the context survives such a call."""
    pair = W.hdl64_pair(max_source=20000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    ctx.project(np.eye(4))
    pr = ctx.pairs()
    s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
    # LS (trimmed)
    ctx.set_params(plo.default_params(solver=1))
    ctx.project(np.eye(4))
    D_res, _ = ctx.solve_ls()
    for thr in (0.02, 0.1):
        Dh, rank = ctx.solve_ls_host(s, d, n, thr)
        assert rank == 6 and np.abs(Dh - oracle_mod.solve_ls(s, d, n, thr)).max() < 1e-9
        if thr == 0.02:
            assert np.abs(Dh - D_res).max() < 1e-10
    # RANSAC with each final stage; a second parameter set that needs several hypotheses
    for final in (0, 1, 2):
        for kw in (dict(ransac_final=final), dict(ransac_final=final, ransac_min_inliers_percentage=0.9999, ransac_max_iterations=7,
                                                   ransac_seed=12345, ransac_distance_threshold=0.05, ls_threshold=0.05)):
            Dh, info = ctx.solve_ransac_host(s, d, n, plo.default_params(**kw))
            ok, Do = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(solver=2, **kw))
            assert ok and np.abs(Dh - Do).max() < 1e-9, (final, kw)
            assert info["hypotheses"] == (7 if len(kw) > 1 else info["hypotheses"]) and info["inliers"] > 0
    # DRPM on caller weights: unit, the RANSAC-final weights (normalised to sum 1), and un-normalised ones
    idx, w = oracle_mod.ransac_weights(s, d, n)
    rng = np.random.default_rng(3)
    for ss, dd, nn, ww in ((s, d, n, None), (s[idx], d[idx], n[idx], w), (s, d, n, rng.uniform(0.1, 3.0, s.shape[0]))):
        Dh, probs = ctx.solve_drpm_host(ss, dd, nn, ww)
        Do, po = oracle_mod.solve_drpm(ss, dd, nn, ww)
        assert np.abs(Dh - Do).max() < 1e-9 and np.abs(probs - po).max() < 1e-9
    # the dispatcher of src/laser_odometry.cpp:173-275 with the reference's strings
    cfg = plo.config.load_config()
    sm = cfg["laser_odometry"]["solve_method"]
    sm["LS"]["threshold"], sm["RANSAC"]["LS_threshold"], sm["RANSAC"]["final_solve_method"] = 0.1, 0.05, "LS"
    ok, D1 = plo.solver.solveMotionEstimationProblem("LS", s, d, n, ctx=ctx, cfg=cfg)
    assert ok and np.abs(D1 - oracle_mod.solve_ls(s, d, n, 0.1)).max() < 1e-9
    ok, D2 = plo.solver.solveMotionEstimationProblem("RANSAC", s, d, n, ctx=ctx, cfg=cfg)
    okr, Do = oracle_mod.solve_ransac(s, d, n, oracle_mod.default_params(solver=2, ransac_final=0, ls_threshold=0.05))
    assert ok and okr and np.abs(D2 - Do).max() < 1e-9             # the trim fraction is RANSAC.LS_threshold, not LS.threshold
    ok, D3 = plo.solver.solveMotionEstimationProblem("Weighted LS", s, d, n, ctx=ctx, cfg=cfg)
    assert ok and np.abs(D3 - oracle_mod.solve_wls(s, d, n)).max() < 1e-10
    for bad in ("Ceres", "ICP", "Teaser", "nonsense"):
        with pytest.raises(ValueError):
            plo.solver.solveMotionEstimationProblem(bad, s, d, n, ctx=ctx, cfg=cfg)
    # coordinates that float32 cannot hold are refused, empty input is not an error of the call
    with pytest.raises(plo.PloError):
        ctx.solve_ls_host(s + 1e-9, d, n)
    # the context still registers (its clouds are untouched, the projection state was dropped)
    ctx.set_params(plo.default_params())
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["iters"] == so["iters"] and _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M


def test_imls_function_and_compute_normal_entry_points(oracle_mod):
    """bool ImplicitMLSFunction(PointType&, double&) and Vector3d ComputeNormal(vector<Vector3d>&)
    (include/imls_icp.h:75-76, :84) as stand-alone calls: heights bitwise equal to those the projection computes
    for the points it keeps, defined (and finite) for points the projection drops at its 1-NN gates, `false` exactly
    where fewer than 3 neighbours are usable; normals against the oracle and numpy's eigh (sign-free)."""
    pair = W.planetary_pair()                       # sparse: every drop reason occurs
    T = pair.T_gt
    ctx, orc = _both(oracle_mod, pair.target, pair.source[::4], h=0.6)
    ctx.project(T, hooks=True)
    q = ctx.query_results()
    o = orc.project(T, hooks=True)
    src = pair.source[::4]
    fin = np.isfinite(src[:, 0:3]).all(axis=1)
    src = src[fin]
    x = (src[:, 0:3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    m = plo.IMLSICPMatcher(ctx=ctx)
    ok, h = m.ImplicitMLSFunction(x, src[:, 4:7])
    kept = q["status"] == 0
    assert kept.sum() > 1000 and ok[kept].all() and np.array_equal(h[kept], q["height"][kept])
    assert np.array_equal(~ok, np.isin(q["status"], [5]) | (~ok & np.isin(q["status"], [1, 2, 3, 4])))   # MLS_FAIL <=> not ok among the gated-in
    assert (q["status"] == 5).sum() == (~ok[np.isin(q["status"], [0, 5, 6])]).sum()
    gated = np.isin(q["status"], [2, 4]) & ok       # too far / normal constraint at the 1-NN: the IMLS function itself still answers
    assert gated.any() and np.isfinite(h[gated]).all()
    ok1, h1 = m.ImplicitMLSFunction(x[7], src[7, 4:7])
    assert ok1 == bool(ok[7]) and (h1 == h[7] or not ok1)
    # heights against the oracle where it kept the point
    okk = o["status"] == 0
    assert (np.abs(h[okk] - o["height"][okk]) <= HEIGHT_RTOL * np.abs(o["height"][okk]) + HEIGHT_ATOL).all()
    # ComputeNormal
    rng = np.random.default_rng(11)
    for trial in range(20):
        k = int(rng.choice([3, 10, 10, 25, 200]))
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        basis = np.linalg.svd(np.eye(3) - np.outer(nrm, nrm))[0][:, :2]
        pts = rng.uniform(-50, 50, 3) + rng.uniform(-1, 1, size=(k, 2)) @ basis.T + rng.normal(0, 0.01, size=(k, 1)) * nrm
        g = m.ComputeNormal(pts)
        ref = oracle_mod.compute_normal(pts)
        c = np.cov(pts.T, bias=True)
        ev = np.linalg.eigh(c)[1][:, 0]
        assert abs(np.linalg.norm(g) - 1) < 1e-12
        assert min(np.abs(g - ref).max(), np.abs(g + ref).max()) < 1e-7
        assert min(np.abs(g - ev).max(), np.abs(g + ev).max()) < 1e-7
    assert m.ComputeNormal().shape == (ctx.n_target, 3)


def test_settled_kernel_parity_from_candidate_tiles(oracle_mod):
    """The candidate-tile path (tile_query, the streaming form of the settled iterations) is held to the same bars as the tree walk:
    with the `force_warm` knob a stepped projection leaves candidate tiles behind and the next ones consume them.
    Small moves (tile hits), a move too large for most tiles (misses go to the tree through the miss list), a map of
    quantised points (ties at the bound, ties by index), coincident points (second search of the 1-NN rule), k = 32,
    ragged sizes, non-finite inputs: neighbour sets, d2, statuses, counters bit-exact, heights to tolerance."""
    P = plo.synth.scenes.pose_matrix
    pair = W.hdl64_pair(max_source=20000)
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    ctx.set_tuning("force_warm", 1)
    _check_projection(ctx, orc, T=pair.T_gt)                       # tree walk, leaves the tiles
    hits = []
    for T in (P([0.002, -0.001, 0.0005], yaw_deg=0.01) @ pair.T_gt,            # mm moves: nearly every tile covers its bound
              P([0.004, 0.001, -0.001], yaw_deg=-0.02, pitch_deg=0.01) @ pair.T_gt,
              P([0.03, -0.02, 0.005], yaw_deg=0.1) @ pair.T_gt,                # cm move: dense regions miss, sparse ones hit
              P([0.5, 0.2, 0.0], yaw_deg=2.0) @ pair.T_gt,                     # a jump: (nearly) everything goes to the tree
              pair.T_gt):
        _check_projection(ctx, orc, T=T)
        hits.append(float((ctx.search_stats()[:, 2] % 100000 >= 50000).mean()))
    # the settled kernel really answered these (the first walk had no temporal reference: its chunk heads left tight tiles)
    assert hits[0] > 0.5 and hits[1] > 0.85, hits
    assert 0.02 < hits[2] < 0.98 and hits[3] < 0.2, hits           # both kernels shared the third, the tree took the fourth
    ctx.close()
    # the resident loop on the same pair: tiles come into play from the fourth projection on
    ctx, orc = _both(oracle_mod, pair.target, pair.source)
    Tg, sg = ctx.register()
    To, so = orc.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"]
    assert np.array_equal(sg["counters"], so["counters"])
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    Tg2, sg2 = ctx.register()                                     # same clouds again: starts with tiles and temporal bounds
    assert np.array_equal(Tg, Tg2) and sg["iters"] == sg2["iters"]
    ctx.close()
    # k = 32, other radii
    ctx2, orc2 = _both(oracle_mod, pair.target[::3], pair.source[::10], search_number=32, r=1.5, h=0.7)
    ctx2.set_tuning("force_warm", 1)
    _check_projection(ctx2, orc2)
    _check_projection(ctx2, orc2, T=P([0.003, 0.0, 0.001]))
    _check_projection(ctx2, orc2, T=P([0.006, 0.002, 0.0], yaw_deg=0.02))
    ctx2.close()
    rng = np.random.default_rng(77)
    for n_t, n_s in ((1, 1), (31, 7), (33, 64), (1025, 100), (5000, 333), (20000, 3000)):
        tgt = np.zeros((n_t, 12), np.float32)
        tgt[:, 0:3] = rng.uniform(-2, 2, size=(n_t, 3))
        if n_t == 20000:
            tgt[:, 0:3] = np.round(tgt[:, 0:3] * 8) / 8    # quantised map: many bit-equal distances, duplicates
        tgt[:, 4:7] = [0, 0, 1]
        src = np.zeros((n_s, 12), np.float32)
        src[:, 0:3] = rng.uniform(-2, 2, size=(n_s, 3))
        src[:, 4:7] = [0, 0, 1]
        if n_s > 50:
            src[3, 0] = np.nan                           # stripped: later indices shift
            src[7, 0:3] = tgt[5, 0:3]                    # coincident with a map point: self-match rule of the 1-NN
        if n_t > 1000:
            tgt[100:160, 0:3] = tgt[100, 0:3]            # 60 coincident map points: ties by index, 1-NN fallback
            src[11, 0:3] = tgt[100, 0:3]
        c3, o3 = _both(oracle_mod, tgt, src)
        c3.set_tuning("force_warm", 1)
        _check_projection(c3, o3)
        _check_projection(c3, o3)                        # same pose again: every valid tile hits
        _check_projection(c3, o3, T=P([0.001, -0.001, 0.0005]))
        _check_projection(c3, o3, T=P([0.0, 0.0, 0.0]))
        c3.close()


def test_local_map_transform_is_bit_exact_and_drops_oldest(oracle_mod):
    """plo_map_push against the oracle's TransformToEnd (src/laser_odometry.cpp:88-114): float32 records bit-exact,
    queue bounded by max_queue (oldest frame first out), the map becomes the target."""
    seq = W.Sequence(seed=2003, n_frames=4, max_points=6000)
    frames = [seq.frame(k).astype(np.float32) for k in range(4)]
    frames[1][5, 0] = np.nan                                        # non-finite points ride along and are stripped by the index
    rng = np.random.default_rng(5)
    ctx = plo.Context(0)
    queue = []
    for k, f in enumerate(frames):
        T = None
        if k > 0:
            T = W.scenes.pose_matrix(rng.uniform(-0.8, 0.8, 3), yaw_deg=rng.uniform(-3, 3), pitch_deg=rng.uniform(-1, 1),
                                     roll_deg=rng.uniform(-1, 1))
        for tn in (True,):
            ctx.map_push(f, T, max_queue=3, transform_normals=tn)
        if T is not None:
            queue = [oracle_mod.transform_to_end(q, T, True) for q in queue]
        queue.append(f.copy())
        queue = queue[-3:]
        want = np.concatenate(queue, axis=0)
        got = ctx.map_records()
        assert ctx.map_info() == (len(queue), want.shape[0])
        assert np.array_equal(got[:, 0:3], want[:, 0:3], equal_nan=True)
        assert np.array_equal(got[:, 4:7], want[:, 4:7], equal_nan=True)
        assert ctx.n_target == int(np.isfinite(want[:, 0:3]).all(axis=1).sum())
    # normals stay untouched without transform_normals
    ctx2 = plo.Context(0)
    ctx2.map_push(frames[0], None, max_queue=2)
    ctx2.map_push(frames[1], T, max_queue=2, transform_normals=False)
    got = ctx2.map_records()
    assert np.array_equal(got[:frames[0].shape[0], 4:7], frames[0][:, 4:7])
    assert np.array_equal(got[:frames[0].shape[0], 0:3], oracle_mod.transform_to_end(frames[0], T, False)[:, 0:3])


def test_consistent_local_map_odometry(oracle_mod):
    """a pose-consistent multi-frame local map (max_queue_size = 3) kept on the device.
    Against the same loop on the oracle (register, TransformToEnd of the queue, append), and against ground truth,
    where the reference's untransformed concatenation is visibly worse."""
    seq = W.Sequence(seed=2004, n_frames=5, max_points=12000)
    frames = [seq.frame(k) for k in range(5)]
    cfg = plo.config.load_config()
    cfg["laser_odometry"]["max_queue_size"] = 3
    odo = plo.LaserOdometry(cfg, resident=True, map_mode="consistent")
    odo.process_frame(frames[0])
    odo_h = plo.LaserOdometry(cfg, resident=False, map_mode="consistent")      # pose handed over by the host instead
    odo_h.run(frames)
    orc = oracle_mod.Oracle()
    queue = [np.asarray(frames[0], np.float32)]
    for k in range(1, 5):
        orc.set_target(np.concatenate(queue, axis=0))
        orc.set_source(frames[k])
        To, so = orc.register()
        _, st = odo.process_frame(frames[k])
        assert st["iters"] == so["iters"] and st["status"] == so["status"] and st["pairs"] == so["pairs"], k
        assert _rot_err(st["rPose"][:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(st["rPose"][:3, 3] - To[:3, 3]) < POSE_M
        assert np.abs(odo_h.frame_stats[k]["rPose"] - st["rPose"]).max() < 1e-7
        # the oracle's queue advances with the GPU's pose, so that both sides keep registering the same bytes
        queue = [oracle_mod.transform_to_end(q, st["rPose"], True) for q in queue] + [np.asarray(frames[k], np.float32)]
        queue = queue[-3:]
        got = odo.ctx.map_records()
        want = np.concatenate(queue, axis=0)
        assert np.array_equal(got[:, 0:3], want[:, 0:3]) and np.array_equal(got[:, 4:7], want[:, 4:7])
        gt = seq.relative_gt(k)
        assert np.linalg.norm(st["rPose"][:3, 3] - gt[:3, 3]) < 0.05 and _rot_err(st["rPose"][:3, :3], gt[:3, :3]) < 5e-3
    ref = plo.LaserOdometry(cfg, resident=True, map_mode="reference")          # the reference's live code, queue of 3
    ref.run(frames)
    err_c = max(np.linalg.norm(odo.frame_stats[k]["rPose"][:3, 3] - seq.relative_gt(k)[:3, 3]) for k in range(2, 5))
    err_r = max(np.linalg.norm(ref.frame_stats[k]["rPose"][:3, 3] - seq.relative_gt(k)[:3, 3]) for k in range(2, 5))
    assert err_c < err_r


def _check_frontend(oracle_mod, ctx, pts, **kw):
    g = ctx.frontend(pts, plo.frontend_default_params(**kw))
    o = oracle_mod.frontend(pts, oracle_mod.frontend_default_params(**kw))
    for key in ("n", "ringed", "pca_failures", "plane_failures", "candidates"):
        assert g[key] == o[key], (key, g[key], o[key])
    assert np.array_equal(g["src_index"], o["src_index"])
    assert np.array_equal(g["candidate"], o["candidate"])
    gr, orr = g["records"], o["records"]
    assert np.array_equal(gr[:, 0:4], orr[:, 0:4])                       # xyz + the constant 1
    assert np.array_equal(gr[:, 4:8], orr[:, 4:8])                       # normals: same float statements on both sides
    assert np.array_equal(g["eigenvalues"], o["eigenvalues"])
    assert np.abs(gr[:, 8] - orr[:, 8]).max() <= 1e-6 if g["n"] else True   # intensity goes through atan2 (libm vs CUDA)
    assert np.array_equal(gr[:, 9:12], orr[:, 9:12])
    return g, o


def test_frontend_normals_and_presample(oracle_mod):
    """SURVEY.md §8f rank 3 — the front-end stage that produces the normals the matcher consumes
    (src/scan_registration.cpp: ring assignment, windowed PCA over three rings, plane check, planarity presample)
    against the oracle's restatement: identical point selection and order, identical normals / eigenvalues /
    presample flags; intensity (scanID + scanPeriod * relTime) within 1e-6."""
    ctx = plo.Context(0)
    pair = W.hdl64_pair()
    pts = np.ascontiguousarray(pair.source[:, 0:3])
    g, o = _check_frontend(oracle_mod, ctx, pts)
    assert g["n"] > 80000 and g["candidates"] > 10000
    # the computed normals agree with the analytic surface normals of the synthetic scene where the plane check held
    ok = g["eigenvalues"][:, 0] > 0
    cosang = np.abs((pair.source[g["src_index"], 4:7] * g["records"][:, 4:7]).sum(axis=1))
    assert np.median(cosang[ok]) > 0.999
    # other ring layouts, windows, thresholds; NaNs, points out of range; use_all_points off
    bad = pts.copy()
    bad[::97, 1] = np.nan
    bad[5::131] *= 100.0
    bad[7::211] *= 0.001
    _check_frontend(oracle_mod, ctx, bad, use_all_points=0, window_size=2, plane_distance_threshold=0.05)
    _check_frontend(oracle_mod, ctx, pts, window_size=4, iter_step=2, knn_distance_threshold=0.5, planarity_threshold=0.2)
    seq = W.Sequence(seed=2001, n_frames=1)                                   # VLP-32C-shaped
    _check_frontend(oracle_mod, ctx, np.ascontiguousarray(seq.frame(0)[:, 0:3]), n_scans=32)
    pl = W.planetary_pair()                                                   # VLP-16
    _check_frontend(oracle_mod, ctx, np.ascontiguousarray(pl.source[:, 0:3]), n_scans=16, plane_distance_threshold=0.05)
    # degenerate inputs
    _check_frontend(oracle_mod, ctx, pts[:0])
    _check_frontend(oracle_mod, ctx, pts[:10])
    _check_frontend(oracle_mod, ctx, np.full((50, 3), np.nan, np.float32))


def test_frontend_feeds_the_matcher_on_the_device(oracle_mod):
    """raw scans -> plo_frontend -> plo_set_target_device / plo_set_source_device -> plo_register, the filtered
    clouds never leaving the GPU; same result as handing the oracle front-end's clouds to the oracle matcher."""
    pair = W.hdl64_pair()
    a = np.ascontiguousarray(pair.target[::2, 0:3])
    b = np.ascontiguousarray(pair.source[::2, 0:3])
    ctx = plo.Context(0)
    ctx.frontend(a, fetch=False)
    ctx.set_target_from_frontend()
    ctx.frontend(b, fetch=False)
    ctx.set_source_from_frontend()
    Tg, sg = ctx.register()
    orc = oracle_mod.Oracle()
    orc.set_target(oracle_mod.frontend(a)["records"])
    orc.set_source(oracle_mod.frontend(b)["records"])
    To, so = orc.register()
    assert sg["status"] == so["status"] and sg["iters"] == so["iters"] and sg["pairs"] == so["pairs"]
    assert _rot_err(Tg[:3, :3], To[:3, :3]) < POSE_RAD and np.linalg.norm(Tg[:3, 3] - To[:3, 3]) < POSE_M
    assert np.linalg.norm(Tg[:3, 3] - pair.T_gt[:3, 3]) < 0.1
