"""Solver surface — mirrors the free functions of include/solver.h:77-139 that are in scope
(LS, WeightedLS, RANSAC, DRPM) and the dispatcher of src/laser_odometry.cpp:173-275.
Every function has the reference's shape: n x 3 doubles in, (flag, deltaTrans 4x4) out."""
from __future__ import annotations

import numpy as np

from . import _lib
from . import config as _config
from .context import Context

_default_ctx: Context | None = None


def _ctx(ctx: Context | None) -> Context:
    global _default_ctx
    if ctx is not None:
        return ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _vecs(*arrs):
    return tuple(np.asarray(a, np.float64).reshape(-1, 3) for a in arrs)


def SolveMotionEstimationProblemWeightedLS_CUDA(source_cloud, ref_cloud, ref_normals, weights=None,
                                               timestamp: str = "", ctx: Context | None = None):
    """include/solver.h:92-98, src/solver.cpp:168-220.  The reference always returns true (:219)."""
    src, ref, nrm = _vecs(source_cloud, ref_cloud, ref_normals)
    delta, _rank = _ctx(ctx).solve_wls_host(src, ref, nrm, weights)
    return True, delta


def SolveMotionEstimationProblemLS_CUDA(source_cloud, ref_cloud, ref_normals, timestamp: str = "", threshold: float = 0.02,
                                       ctx: Context | None = None):
    """include/solver.h:84-90, src/solver.cpp:74-166 (LS, then a second LS on the [thr, 1-thr] residual rank window)."""
    src, ref, nrm = _vecs(source_cloud, ref_cloud, ref_normals)
    delta, _rank = _ctx(ctx).solve_ls_host(src, ref, nrm, threshold)
    return True, delta


def SolveMotionEstimationProblemRANSAC_CUDA(source_cloud, ref_cloud, ref_normals, timestamp: str = "", max_iterations: int = 5000,
                                           distance_threshold: float = 0.8, min_inliers_percentage: float = 0.95,
                                           huber_threshold: float = 0.648, final_solve_method: str = "DRPM",
                                           ls_threshold: float = 0.02, drpm_threshold: float = 0.05,
                                           drpm_stdev_points: float = 0.02, drpm_stdev_normals: float = 0.05,
                                           seed: int = 1, ctx: Context | None = None):
    """include/solver.h:100-114, src/solver.cpp:222-385 — argument order of the reference; `seed` replaces its
    unseeded rand().  An unknown final_solve_method raises (the reference prints and returns false, :377-380)."""
    if final_solve_method not in _config.RANSAC_FINALS:
        raise ValueError(f"Invalid final solve method! ({final_solve_method!r})")
    p = _lib.default_params(ransac_max_iterations=int(max_iterations), ransac_distance_threshold=float(distance_threshold),
                            ransac_min_inliers_percentage=float(min_inliers_percentage), huber_threshold=float(huber_threshold),
                            ransac_final=_config.RANSAC_FINALS[final_solve_method], ls_threshold=float(ls_threshold),
                            drpm_threshold=float(drpm_threshold), drpm_stdev_points=float(drpm_stdev_points),
                            drpm_stdev_normals=float(drpm_stdev_normals), ransac_seed=int(seed))
    src, ref, nrm = _vecs(source_cloud, ref_cloud, ref_normals)
    delta, _info = _ctx(ctx).solve_ransac_host(src, ref, nrm, p)
    return True, delta


def SolveMotionEstimationProblemDRPM_CUDA(source_cloud, ref_cloud, ref_normals, weights=None, timestamp: str = "",
                                         threshold: float = 0.05, stdev_points: float = 0.02, stdev_normals: float = 0.05,
                                         ctx: Context | None = None):
    """include/solver.h:129-139, src/solver.cpp:499-603."""
    src, ref, nrm = _vecs(source_cloud, ref_cloud, ref_normals)
    delta, _probs = _ctx(ctx).solve_drpm_host(src, ref, nrm, weights, threshold, stdev_points, stdev_normals)
    return True, delta


def solveMotionEstimationProblem(solve_method: str, in_cloud_vec, ref_cloud_vec, ref_normal, timestamp: str = "",
                                 ctx: Context | None = None, cfg: dict | None = None):
    """Dispatcher of src/laser_odometry.cpp:173-275 for the in-scope methods, reading the same config.json keys
    (solve_method.LS.threshold :190, solve_method.RANSAC.* :194-207); accepts the same strings as the C++ adapter and
    `LaserOdometry`.  Ceres / ICP / Teaser and unknown strings raise instead of printing (:271)."""
    sm = (cfg or _config.load_config())["laser_odometry"]["solve_method"]
    if solve_method in ("WeightedLS_CUDA", "Weighted LS"):
        return SolveMotionEstimationProblemWeightedLS_CUDA(in_cloud_vec, ref_cloud_vec, ref_normal, None, timestamp, ctx)
    if solve_method == "LS":
        return SolveMotionEstimationProblemLS_CUDA(in_cloud_vec, ref_cloud_vec, ref_normal, timestamp,
                                                   float(sm.get("LS", {}).get("threshold", 0.02)), ctx)
    if solve_method == "RANSAC":
        r = sm.get("RANSAC", {})
        return SolveMotionEstimationProblemRANSAC_CUDA(
            in_cloud_vec, ref_cloud_vec, ref_normal, timestamp, int(r.get("max_iterations", 5000)),
            float(r.get("distance_threshold", 0.8)), float(r.get("min_inliers_percentage", 0.95)),
            float(r.get("huber_threshold", 0.648)), str(r.get("final_solve_method", "DRPM")),
            float(r.get("LS_threshold", sm.get("LS", {}).get("threshold", 0.02))), float(r.get("DRPM_threshold", 0.05)),
            float(r.get("DRPM_stdev_points", 0.02)), float(r.get("DRPM_stdev_normals", 0.05)), ctx=ctx)
    if solve_method in ("Ceres", "ICP", "Teaser"):
        raise ValueError(f"solve_method {solve_method!r} is out of scope (absent third-party solvers, SURVEY.md §2.1)")
    raise ValueError(f"Invalid SOLVE_METHOD! ({solve_method!r})")
