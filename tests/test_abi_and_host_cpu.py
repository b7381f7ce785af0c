"""CPU-only checks of the boundary and the host logic (no compute calls):
the C-ABI library loads, exports every symbol include/plo/plo_c_api.h declares, fails loudly
without a GPU; config.json flattening mirrors the reference's dispatch and rejects what is out
of scope."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import plo_b200 as plo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "plo", "plo_c_api.h")).read()
    return sorted(set(re.findall(r"PLO_API\s+[\w\s\*]+?\b(plo_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 25
    L = plo._lib.lib()
    for n in names:
        assert hasattr(L, n), f"libplo_cuda.so does not export {n}"
    assert sorted(plo._lib.EXPORTS) == names
    assert L.plo_version() >= 100


def test_struct_layouts_match_header():
    # sizes the C compiler gives the ABI structs (checked against ctypes mirrors)
    src = '#include "plo/plo_c_api.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(plo_params), sizeof(plo_proj_stats), sizeof(plo_reg_stats));return 0;}'
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert [int(x) for x in out] == [C.sizeof(plo._lib.PloParams), C.sizeof(plo._lib.PloProjStats), C.sizeof(plo._lib.PloRegStats)]


def test_defaults_match_reference_config_json():
    p = plo.default_params()
    assert (p.iterations, p.h, p.r, p.r_normal) == (30, 1.0, 3.0, 1.0)
    assert (p.is_get_normals, p.search_number_normal, p.search_number) == (1, 10, 20)
    assert (p.normal_angle_constraint, p.angle_diff_threshold, p.transform_normal, p.correspond_number) == (1, 30.0, 0, 6)
    assert (p.delta_dist_threshold, p.delta_angle_threshold) == (0.001, 0.0001745353)
    assert (p.ransac_distance_threshold, p.huber_threshold) == (0.8, 0.648)


def test_no_gpu_fails_loudly():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(plo.PloError) as e:
        plo.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_config_flattening_and_rejections(tmp_path):
    cfg = plo.config.load_config()
    p = plo.config.params_from_config(cfg)
    assert p.search_number == 20 and p.weight_mode == 0
    # the reference's own config.json layout (only the keys the path reads)
    ref_like = json.loads(json.dumps(cfg))
    ref_like["laser_odometry"]["solve_method"]["method"] = "RANSAC"
    ref_like["laser_odometry"]["solve_method"]["RANSAC"]["final_solve_method"] = "Weighted LS"
    f = tmp_path / "config.json"
    f.write_text(json.dumps(ref_like))
    p = plo.config.params_from_config(plo.config.load_config(str(f)))
    assert p.solver == 2 and p.ransac_final == 1
    # the reference's literal default chain: RANSAC -> DRPM
    ref_default = json.loads(json.dumps(ref_like))
    ref_default["laser_odometry"]["solve_method"]["RANSAC"]["final_solve_method"] = "DRPM"
    p = plo.config.params_from_config(ref_default)
    assert p.solver == 2 and p.ransac_final == 2 and p.drpm_threshold == 0.05
    ref_default["laser_odometry"]["solve_method"]["RANSAC"]["final_solve_method"] = "LS"
    assert plo.config.params_from_config(ref_default).ransac_final == 0
    for mutate, msg in [
        (lambda c: c["laser_odometry"]["matching_method"].__setitem__("method", "plane_ICP"), "plane_ICP"),
        (lambda c: c["laser_odometry"]["matching_method"].__setitem__("method", "bogus"), "Invalid MATCHING_METHOD"),
        (lambda c: c["laser_odometry"]["solve_method"].__setitem__("method", "Teaser"), "Teaser"),
        (lambda c: c["laser_odometry"]["solve_method"].__setitem__("method", "nope"), "Invalid SOLVE_METHOD"),
        (lambda c: c["laser_odometry"]["solve_method"]["RANSAC"].__setitem__("final_solve_method", "QR"), "unknown final_solve_method"),
        (lambda c: c["laser_odometry"]["matching_method"]["IMLS"]["use_tensor_voting"].__setitem__("enabled", True), "tensor"),
        (lambda c: c.__setitem__("backend", "cpu"), "no CPU fallback"),
    ]:
        c2 = json.loads(json.dumps(ref_like))
        mutate(c2)
        with pytest.raises(plo.config.ConfigError) as e:
            plo.config.params_from_config(c2)
        assert msg in str(e.value)


def test_ransac_final_ls_uses_its_own_threshold():
    """src/laser_odometry.cpp:205: the trim fraction of the final "LS" inside RANSAC is solve_method.RANSAC.LS_threshold,
    not solve_method.LS.threshold (the two keys differ in a config.json that says so)."""
    import plo_b200 as plo
    cfg = plo.config.load_config()
    sm = cfg["laser_odometry"]["solve_method"]
    sm["LS"]["threshold"], sm["RANSAC"]["LS_threshold"] = 0.1, 0.05
    sm["method"], sm["RANSAC"]["final_solve_method"] = "RANSAC", "LS"
    assert plo.config.params_from_config(cfg).ls_threshold == 0.05
    del sm["RANSAC"]["LS_threshold"]
    assert plo.config.params_from_config(cfg).ls_threshold == 0.1      # absent: fall back
    sm["RANSAC"]["LS_threshold"] = 0.05
    sm["method"] = "LS"
    assert plo.config.params_from_config(cfg).ls_threshold == 0.1


def test_transposed_butterfly_gives_the_bits_of_one_tree_per_value():
    """csrc/p2plane_device.cuh:warp_tree_sum16 sums sixteen values over the 32 lanes with 8 + 4 + 2 + 1 + 1 exchanges; the
    claim the bitwise reproducibility of the loop rests on is that every total has the bits of the plain xor butterfly
    (offsets 16, 8, 4, 2, 1) of that value alone.  Lane-by-lane restatement of both in numpy, same fp64 additions."""
    rng = np.random.default_rng(7)
    v = rng.normal(size=(32, 16)) * np.exp(rng.uniform(-30, 30, size=(32, 16)))     # [lane][value], wide dynamic range
    # plain butterfly: every lane ends with the total of every value
    plain = v.copy()
    for o in (16, 8, 4, 2, 1):
        plain = plain + plain[np.arange(32) ^ o]
    assert all(np.array_equal(plain[0], plain[l]) for l in range(32))
    # transposed: after the exchange at distance 2 * width a lane keeps `width` values
    cur = [list(v[l]) for l in range(32)]
    for width in (8, 4, 2, 1):
        nxt = []
        for l in range(32):
            up = (l & (2 * width)) != 0
            partner = cur[l ^ (2 * width)]
            p_up = ((l ^ (2 * width)) & (2 * width)) != 0
            row = []
            for i in range(width):
                keep = cur[l][i + width] if up else cur[l][i]
                recv = partner[i] if p_up else partner[i + width]       # what the partner sends: the half it does not keep
                row.append(np.float64(keep) + np.float64(recv))
            nxt.append(row)
        cur = nxt
    tot = [np.float64(cur[l][0]) + np.float64(cur[l ^ 1][0]) for l in range(32)]
    for l in range(32):
        slot = (((l >> 4) & 1) << 3) | (((l >> 3) & 1) << 2) | (((l >> 2) & 1) << 1) | ((l >> 1) & 1)
        assert tot[l] == plain[0][slot] and np.float64(tot[l]).tobytes() == np.float64(plain[0][slot]).tobytes(), (l, slot)


def _ldlt6_restatement(H, g, count):
    """csrc/p2plane_device.cuh:solve_ldlt6_warp, element by element: the 6 x 7 augmented matrix, diagonal pivoting with
    the squared ColPivHouseholderQR threshold, M[i][j] -= (M[p][hi] / dp) * M[p][lo] on the rows / columns not yet taken,
    column-wise back substitution in reverse pivot order."""
    eps = np.finfo(np.float64).eps
    M = np.concatenate([H, g[:, None]], axis=1).astype(np.float64)
    helper = (M.diagonal().max() * eps) * eps / max(count, 1.0)
    done, order = set(), []
    for k in range(6):
        p, dp = -1, 0.0
        for i in range(6):
            if i not in done and (p < 0 or M[i, i] > dp):
                p, dp = i, M[i, i]
        if not (dp > 0.0) or dp < helper * (count - k):
            break
        done.add(p)
        order.append(p)
        N = M.copy()
        for i in range(6):
            for j in range(7):
                if i in done or (j < 6 and j in done):
                    continue
                hi, lo = (i, 6) if j == 6 else (max(i, j), min(i, j))
                N[i, j] = M[i, j] - (M[p, hi] / dp) * M[p, lo]
        M = N
        live = [i for i in range(6) if i not in done]
        assert np.array_equal(M[np.ix_(live, live)], M[np.ix_(live, live)].T)       # exactly symmetric, as claimed
    x, rhs = np.zeros(6), M[:, 6].copy()
    for p in reversed(order):
        x[p] = rhs[p] / M[p, p]
        rhs = rhs - M[:, p] * x[p]
    return x, len(order)


def test_warp_solve_restatement_against_numpy():
    """The algorithm of the device's 6 x 6 solve (pivot rule, rank decision, symmetric elimination, back substitution)
    against numpy: full-rank normal equations to 1e-10 of the solution, a rank-3 system (points on a plane, normals all
    e_z: only z, roll and pitch are observable) with the dropped unknowns left at zero, and the no-pivot case."""
    rng = np.random.default_rng(11)
    for _ in range(50):
        A = rng.normal(size=(200, 6)) * rng.uniform(0.1, 10, size=6)
        b = rng.normal(size=200)
        H, g = A.T @ A, A.T @ b
        x, rank = _ldlt6_restatement(H, g, 200.0)
        ref = np.linalg.solve(H, g)
        assert rank == 6 and np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    s = rng.uniform(-5, 5, size=(300, 3)) * [1, 1, 0]
    n = np.tile([0.0, 0.0, 1.0], (300, 1))
    A = np.concatenate([np.cross(s, n), n], axis=1)          # columns 2 (yaw), 3, 4 (x, y) are exactly zero
    xt = np.array([0.01, -0.02, 0.0, 0.0, 0.0, 0.3])
    H, g = A.T @ A, A.T @ (A @ xt)
    x, rank = _ldlt6_restatement(H, g, 300.0)
    assert rank == 3 == np.linalg.matrix_rank(A)
    assert np.array_equal(x[[2, 3, 4]], np.zeros(3)) and np.abs(x - xt).max() < 1e-12
    x, rank = _ldlt6_restatement(np.zeros((6, 6)), np.zeros(6), 300.0)
    assert rank == 0 and not x.any()


def test_tum_pose_format(tmp_path):
    T = plo.synth.scenes.pose_matrix([1.5, -2.25, 0.125], yaw_deg=90)
    f = tmp_path / "poses.txt"
    plo.save_poses_tum(str(f), [np.eye(4), T])
    lines = f.read_text().strip().split("\n")
    assert lines[0] == "0.000000 0.000000 0.000000 0.000000 0.000000 0.000000 0.000000 1.000000"
    v = [float(x) for x in lines[1].split()]
    assert v[1:4] == [1.5, -2.25, 0.125] and abs(v[6] - np.sqrt(0.5)) < 1e-6 and abs(v[7] - np.sqrt(0.5)) < 1e-6
