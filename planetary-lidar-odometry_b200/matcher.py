"""`IMLSICPMatcher` — host-side mirror of the reference class (include/imls_icp.h:45-147,
src/imls_icp.cpp) over the C ABI.  Method names and argument order follow the reference so
that the parity tests read like tests of the original; pythonic aliases are provided.

Differences that the boundary forces (SURVEY.md §8b "Ownership"): inputs are never mutated —
`ProjSourcePtToSurface` returns the surviving pairs plus `src_idx` instead of erasing
unmatched points from `in_cloud` in place; clouds are (n, 12) float32 arrays of 48-byte
PointXYZINormal records (or torch CUDA tensors of the same layout).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .context import Context


class IMLSICPMatcher:
    def __init__(self, device: int = 0, params: _lib.PloParams | None = None, ctx: Context | None = None):
        # the reference's parameterised ctor leaves `this` uninitialised (src/imls_icp.cpp:33-45);
        # here both forms initialise the same state
        self.ctx = ctx or Context(device, params)
        self.m_iterations = self.ctx.params.iterations
        self.last_counters = None

    # ---- include/imls_icp.h:56-66 --------------------------------------------------------
    def setSourcePointCloud(self, cloud):
        """src/imls_icp.cpp:74-78"""
        self.ctx.set_source(cloud)

    def setTargetPointCloud(self, cloud):
        """src/imls_icp.cpp:80-103: strips non-finite points, builds the spatial index"""
        self.ctx.set_target(cloud)

    def setTargetPointCloudDP(self, cloud):
        """src/imls_icp.cpp:105-144 (libpointmatcher DataPoints, tensor-voting only): out of scope"""
        raise _lib.PloError(-4, "setTargetPointCloudDP: tensor voting branch is out of scope (SURVEY.md §8a a13)")

    def setParameters(self, _iter, _h, _r, _r_normal, _r_proj, _useTensorVoting, _isGetNormals,
                      _useProjectedDistance, _tensor_k, _tensor_sigma, _tensor_distance_threshold,
                      _search_number_normal, _search_number, _normal_angle_constraint,
                      _angle_diff_threshold, _output_dir=""):
        """src/imls_icp.cpp:146-168 — same 16 arguments, same order"""
        if _useTensorVoting:
            raise _lib.PloError(-4, "use_tensor_voting is out of scope (SURVEY.md §8a a13)")
        if _useProjectedDistance:
            raise _lib.PloError(-4, "use_projected_distance is out of scope (SURVEY.md §8a a13)")
        p = self.ctx.params
        p.iterations = int(_iter)
        p.h, p.r, p.r_normal = float(_h), float(_r), float(_r_normal)
        p.is_get_normals = int(bool(_isGetNormals))
        p.search_number_normal = int(_search_number_normal)
        p.search_number = int(_search_number)
        p.normal_angle_constraint = int(bool(_normal_angle_constraint))
        p.angle_diff_threshold = float(_angle_diff_threshold)
        self.ctx.set_params(p)
        self.m_iterations = p.iterations

    # ---- include/imls_icp.h:79-82 --------------------------------------------------------
    def ProjSourcePtToSurface(self, rPose=None, hooks: bool = False):
        """Per-iteration transform (src/laser_odometry.cpp:527-549) + projection
        (src/imls_icp.cpp:496-745).  Returns dict(in_cloud (n,3) f32, out_cloud (n,3) f32,
        out_normal (n,3) f32, src_idx, counters[6])."""
        st = self.ctx.project(rPose, hooks=hooks)
        pr = self.ctx.pairs()
        self.last_counters = st["counters"]
        return dict(in_cloud=pr["src_xyz"], out_cloud=pr["ref_xyz"], out_normal=pr["ref_n"], src_idx=pr["src_idx"],
                    counters=st["counters"], n_source=st["n_source"])

    def ImplicitMLSFunction(self, x, normal=None):
        """bool ImplicitMLSFunction(PointType& x, double& height), include/imls_icp.h:75-76, src/imls_icp.cpp:301-483:
        returns (ok, height).  `x`: the (already transformed) point xyz with `normal`, or a 6-vector / (n, 6) batch
        (then arrays come back).  No 1-NN gates here -- those are ProjSourcePtToSurface's."""
        a = np.asarray(x, np.float32)
        if normal is not None:
            a = np.concatenate([a.reshape(-1, 3), np.asarray(normal, np.float32).reshape(-1, 3)], axis=1)
        single = a.size == 6
        h, ok = self.ctx.imls_height(a.reshape(-1, 6))
        return (bool(ok[0]), float(h[0])) if single else (ok, h)

    def ComputeNormal(self, nearPoints=None):
        """Eigen::Vector3d ComputeNormal(std::vector<Eigen::Vector3d>& nearPoints), include/imls_icp.h:84,
        src/imls_icp.cpp:753-794: unit eigenvector of the smallest eigenvalue of the population covariance (no sign
        disambiguation, like the reference).  Without an argument: the normals the matcher uses for every target
        point, (n, 3) -- the delivered ones when get_normals.enabled, else this function over each point's
        search_number_normal neighbours (oriented +z, deviation D2)."""
        if nearPoints is None:
            return self.ctx.target_normals()
        return self.ctx.compute_normal(nearPoints)

    # ---- include/imls_icp.h:86-88 --------------------------------------------------------
    def Match(self, T0=None):
        """src/imls_icp.cpp:804-919 shape: returns (ok, finalPose, covariance, stats).  The loop is
        the driver's (src/laser_odometry.cpp:524-647) and runs resident on the device; like the
        reference, covariance is only ever identity (:811)."""
        T, st = self.ctx.register(T0)
        ok = st["status"] in (1, 2)
        return ok, T, np.eye(4), st

    # pythonic aliases
    set_source = setSourcePointCloud
    set_target = setTargetPointCloud
    project = ProjSourcePtToSurface
    match = Match
