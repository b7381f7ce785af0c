"""Solver surface — mirrors the free functions of include/solver.h:77-139 for the one solver
on the hot path: SolveMotionEstimationProblemWeightedLS (src/solver.cpp:168-220)."""
from __future__ import annotations

import numpy as np

from .context import Context

_default_ctx: Context | None = None


def _ctx(ctx: Context | None) -> Context:
    global _default_ctx
    if ctx is not None:
        return ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def SolveMotionEstimationProblemWeightedLS_CUDA(source_cloud, ref_cloud, ref_normals, weights=None,
                                               timestamp: str = "", ctx: Context | None = None):
    """Same shape as the reference: n x 3 doubles (+ n weights) in, (flag, deltaTrans 4x4) out.
    The reference always returns true (src/solver.cpp:219)."""
    src, ref, nrm = (np.asarray(a, np.float64).reshape(-1, 3) for a in (source_cloud, ref_cloud, ref_normals))
    delta, _rank = _ctx(ctx).solve_wls_host(src, ref, nrm, weights)
    return True, delta


def solveMotionEstimationProblem(solve_method: str, in_cloud_vec, ref_cloud_vec, ref_normal, timestamp: str = "",
                                 ctx: Context | None = None):
    """Dispatcher of src/laser_odometry.cpp:173-275 restricted to the in-scope method;
    unknown strings raise instead of printing (:271)."""
    if solve_method in ("WeightedLS_CUDA", "Weighted LS"):
        return SolveMotionEstimationProblemWeightedLS_CUDA(in_cloud_vec, ref_cloud_vec, ref_normal, None, timestamp, ctx)
    raise ValueError(f"Invalid SOLVE_METHOD! ({solve_method!r})")
