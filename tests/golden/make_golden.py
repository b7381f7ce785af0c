"""Generates tests/golden/*.npz — committed input/output vectors for the hot path.

The reference has no golden vectors and cannot be built or imported here, so these
fixtures are NOT outputs of the reference: they are outputs of the C oracle (oracle/plo_oracle.c)
that were accepted only after the independent numpy/scipy restatement (oracle/py/imls_ref.py)
reproduced them (neighbour sets / d2 / status bit-exact, heights 1e-12, poses 1e-10) — the check runs
below, before anything is written.  They pin the oracle, the numpy restatement and the CUDA path to
one set of numbers across rounds.

    python tests/golden/make_golden.py          # rewrites the fixtures (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))

import imls_ref  # noqa: E402
import oracle_ctypes as orc  # noqa: E402
import plo_b200 as plo  # noqa: E402  (synthetic inputs only; no GPU needed)

W = plo.synth.workloads


def crop(rec, half):
    m = (np.abs(rec[:, 0]) < half) & (np.abs(rec[:, 1]) < half)
    return np.ascontiguousarray(rec[m])


def case_urban():
    """HDL-64 pair (BASELINE cfg-1 generator, seed 1001), cropped to the 16 m box around the sensor."""
    pair = W.hdl64_pair()
    tgt = crop(pair.target, 8.0)[::3]
    src = crop(pair.source, 8.0)[::40]
    return tgt, src, pair.T_gt


def case_planetary():
    """VLP-16 pair on the fractal heightfield (BASELINE cfg-3 generator, seed 3001), h = 2 / r = 6."""
    pair = W.planetary_pair()
    tgt = crop(pair.target, 12.0)[::2]
    src = crop(pair.source, 12.0)[::12]
    return tgt, src, pair.T_gt


def projection_block(tgt, src, T, **kw):
    o = orc.Oracle(orc.default_params(**kw))
    o.set_target(tgt)
    o.set_source(src)
    out = o.project(T, hooks=True)
    ref = imls_ref.Ref(tgt, src, h=kw.get("h", 1.0), r=kw.get("r", 3.0), k=kw.get("search_number", 20))
    rp = ref.project(T)
    assert np.array_equal(out["nn_idx"], rp["nn_idx"]) and np.array_equal(out["nn_d2"], rp["nn_d2"])
    assert np.array_equal(out["status"], rp["status"]) and np.array_equal(out["nn1_idx"], rp["nn1_idx"])
    ok = out["status"] == 0
    assert np.allclose(out["height"][ok], rp["height"][ok], rtol=1e-12, atol=1e-15)
    assert np.array_equal(out["src_xyz"], rp["src_xyz"]) and np.array_equal(out["ref_n"], rp["ref_n"])
    return o, out


def frontend_case():
    """Front-end golden: a 70-degree sector of the VLP-16 planetary scan (seed 3001)."""
    pair = W.planetary_pair()
    pts = np.ascontiguousarray(pair.source[:, 0:3])
    az = np.degrees(np.arctan2(pts[:, 1], pts[:, 0]))
    pts = np.ascontiguousarray(pts[(az > -35) & (az < 35)])
    kw = dict(n_scans=16, plane_distance_threshold=0.05)
    r = orc.frontend(pts, orc.frontend_default_params(**kw))
    # independent check of the normals: float64 eigh of the 21-row windows rebuilt from ring ids in numpy
    ang = np.degrees(np.arctan(pts[:, 2] / np.hypot(pts[:, 0], pts[:, 1])))
    rid = np.floor((ang + 15) / 2 + 0.5).astype(int)
    ok = np.nonzero(r["eigenvalues"][:, 0] > 0)[0]
    for k in ok[:: max(1, len(ok) // 50)]:
        q = r["records"][k, 0:3]
        i = rid[r["src_index"][k]]
        own = pts[rid == i]
        j = int(np.nonzero((own == q).all(axis=1))[0][0])
        rows = [own[j - 3:j + 4]]
        for nbr in (i - 1, i + 1):
            cl = pts[rid == nbr]
            nn = int(np.argmin(((cl - q) ** 2).sum(axis=1)))
            rows.append(cl[nn - 3:nn + 4])
        P = np.concatenate(rows).astype(np.float64)
        assert P.shape == (21, 3)
        w, V = np.linalg.eigh(np.cov(P.T))
        assert abs(abs(V[:, 0] @ r["records"][k, 4:7]) - 1) < 1e-4 and np.allclose(w[::-1], r["eigenvalues"][k], rtol=5e-3, atol=1e-7)
    out = os.path.join(HERE, "frontend_vlp16.npz")
    np.savez_compressed(out, points=pts, n_scans=16, plane_distance_threshold=0.05, records=r["records"],
                        eigenvalues=r["eigenvalues"], candidate=r["candidate"], src_index=r["src_index"],
                        stats=np.array([r["n"], r["ringed"], r["pca_failures"], r["plane_failures"], r["candidates"]], np.int64))
    print("frontend_vlp16 points", pts.shape[0], "->", r["n"], "candidates", r["candidates"], os.path.getsize(out) // 1024, "KiB")


def main():
    frontend_case()
    for name, maker, kw in (("urban_hdl64", case_urban, {}), ("planetary_vlp16", case_planetary, dict(h=2.0, r=6.0))):
        tgt, src, T_gt = maker()
        o, pr = projection_block(tgt, src, np.eye(4), **kw)
        _, pr_gt = projection_block(tgt, src, T_gt, **kw)
        s, d, n = (pr[k].astype(np.float64) for k in ("src_xyz", "ref_xyz", "ref_n"))
        d_wls = orc.solve_wls(s, d, n)
        d_ref = imls_ref.solve_wls(s, d, n)
        assert np.abs(d_wls - d_ref).max() < 1e-10
        d_ls = orc.solve_ls(s, d, n)
        _, d_rw = orc.solve_ransac(s, d, n, orc.default_params(solver=2, ransac_final=1, **kw))
        _, d_rd = orc.solve_ransac(s, d, n, orc.default_params(solver=2, ransac_final=2, **kw))
        H, g, sw, sbb = orc.normal_equations(s, d, n)
        T_reg, st = o.register()
        T_np, it_np, _ = imls_ref.register(imls_ref.Ref(tgt, src, h=kw.get("h", 1.0), r=kw.get("r", 3.0)))
        assert it_np == st["iters"] and np.abs(T_np - T_reg).max() < 1e-10
        regs = {}
        for label, skw in (("ls", dict(solver=1)), ("ransac_wls", dict(solver=2, ransac_final=1)),
                           ("ransac_drpm", dict(solver=2, ransac_final=2))):
            o2 = orc.Oracle(orc.default_params(**kw, **skw))
            o2.set_target(tgt)
            o2.set_source(src)
            T2, st2 = o2.register()
            regs[f"reg_{label}_T"] = T2
            regs[f"reg_{label}_stats"] = np.array([st2["status"], st2["iters"], st2["pairs"]], np.int64)
        out = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(
            out, target=tgt, source=src, T_gt=T_gt, h=kw.get("h", 1.0), r=kw.get("r", 3.0),
            nn_idx=pr["nn_idx"], nn_d2=pr["nn_d2"], nn1_idx=pr["nn1_idx"], nn1_d2=pr["nn1_d2"], status=pr["status"],
            height=pr["height"], counters=pr["counters"], src_idx=pr["src_idx"], src_xyz=pr["src_xyz"],
            ref_xyz=pr["ref_xyz"], ref_n=pr["ref_n"],
            gt_nn_idx=pr_gt["nn_idx"], gt_status=pr_gt["status"], gt_height=pr_gt["height"], gt_counters=pr_gt["counters"],
            H=H, g=g, sw=sw, swbb=sbb, delta_wls=d_wls, delta_ls=d_ls, delta_ransac_wls=d_rw, delta_ransac_drpm=d_rd,
            reg_wls_T=T_reg, reg_wls_stats=np.array([st["status"], st["iters"], st["pairs"]], np.int64), **regs)
        print(name, "target", tgt.shape[0], "source", src.shape[0], "pairs", pr["n"], "counters", pr["counters"],
              "iters", st["iters"], "->", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
