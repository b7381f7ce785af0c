"""`LaserOdometry` — the L3 driver of src/laser_odometry.cpp:416-683 without ROS: frames come
from a Python iterable instead of topics, poses go to a list / TUM file instead of a publisher.

Per frame (processData): frame 0 only seeds the target (:478, :668-670); every later frame is
registered against the accumulated target from rPose = identity (:484-485) and chained as
nowPose = prevLaserPose * rPose (:649-655).  Two loop modes:
  resident=True   plo_register: the whole loop on the device (the product path)
  resident=False  the reference's own structure: ProjSourcePtToSurface -> getXYZ/getNormals ->
                  solveMotionEstimationProblem -> rPose = deltaTrans * rPose, one host round
                  trip per iteration (boundary-parity mode, used by the tests)
Two target modes (accumulateTargetCloud, :116-136):
  map_mode="reference"   the reference's live code: the queued clouds are concatenated as they are
                         (only right for max_queue_size = 1, the config.json default)
  map_mode="consistent"  the TransformToEnd step the reference left commented out (:118-124) is applied: the
                         queue lives on the device (plo_map_push), every queued frame is moved into the newest
                         frame's coordinates with the pose still resident from plo_register, only the new frame
                         crosses the bus, the index is rebuilt on the device
"""
from __future__ import annotations

import collections

import numpy as np

from . import config as _config
from .context import Context
from .matcher import IMLSICPMatcher


class LaserOdometry:
    def __init__(self, cfg: dict | None = None, device: int = 0, resident: bool = True, ctx: Context | None = None,
                 map_mode: str = "reference", map_transform_normals: bool = True):
        self.cfg = cfg or _config.load_config()
        self.params = _config.params_from_config(self.cfg)
        self.ctx = ctx or Context(device, self.params)
        if ctx is not None:
            self.ctx.set_params(self.params)
        self.matcher = IMLSICPMatcher(ctx=self.ctx)
        self.resident = resident
        if map_mode not in ("reference", "consistent"):
            raise ValueError(f"map_mode {map_mode!r}: 'reference' or 'consistent'")
        self.map_mode = map_mode
        self.map_transform_normals = map_transform_normals
        if map_mode == "consistent":
            self.ctx.map_reset()
        self.max_queue_size = int(self.cfg["laser_odometry"].get("max_queue_size", 1))
        self.cloudQueue = collections.deque()
        self.prevLaserPose = np.eye(4)
        self.frameCount = 0
        self.poses = []          # global poses, one per processed frame (frame 0 = identity)
        self.frame_stats = []
        self._target = None

    # accumulateTargetCloud, src/laser_odometry.cpp:116-136 (no re-transformation, as upstream)
    def _accumulate(self, cloud):
        self.cloudQueue.append(cloud)
        while len(self.cloudQueue) > self.max_queue_size:
            self.cloudQueue.popleft()
        if len(self.cloudQueue) == 1:
            return self.cloudQueue[0]
        return np.concatenate([np.asarray(c) for c in self.cloudQueue], axis=0)

    def _icp_host_loop(self, T0=None):
        """src/laser_odometry.cpp:524-647, one host round trip per iteration"""
        p = self.params
        rPose = np.eye(4) if T0 is None else np.array(T0, np.float64)
        status, iters, pairs = 2, 0, 0
        for _ in range(p.iterations):
            pr = self.matcher.ProjSourcePtToSurface(rPose)
            pairs = pr["in_cloud"].shape[0]
            if pairs < p.correspond_number:            # :570-576
                status = 3
                break
            if p.solver == 2:                          # :609 on the device-resident pairs
                delta = self.ctx.solve_ransac()[0]
            else:
                delta, _rank = self.ctx.solve_ls() if p.solver == 1 else self.ctx.solve_wls()
            rPose = delta @ rPose                      # :619
            iters += 1
            dd = float(np.sqrt(delta[0, 3] ** 2 + delta[1, 3] ** 2 + delta[2, 3] ** 2))
            ct = min(1.0, max((np.trace(delta[:3, :3]) - 1.0) / 2.0, -1.0))
            if dd < p.delta_dist_threshold and np.arccos(ct) < p.delta_angle_threshold:   # :628-646
                status = 1
                break
        return rPose, dict(status=status, iters=iters, pairs=pairs)

    def process_frame(self, filtered_cloud, flat_cloud=None):
        """One pass of processData's body.  filtered_cloud = /laser_cloud_filtered (all points +
        normals), flat_cloud = /laser_cloud_flat (sampled source; default: the full cloud)."""
        flat_cloud = filtered_cloud if flat_cloud is None else flat_cloud
        stats = None
        consistent = self.map_mode == "consistent"
        if self.frameCount != 0:                                   # :478
            self.matcher.setSourcePointCloud(flat_cloud)           # :509
            if not consistent:
                self.matcher.setTargetPointCloud(self._target)     # :510 (consistent: the device map is the target)
            if self.resident:
                rPose, stats = self.ctx.register(None)             # :484-485 identity start
            else:
                rPose, stats = self._icp_host_loop()
            nowPose = self.prevLaserPose @ rPose                   # :652
            self.prevLaserPose = nowPose
            stats["rPose"] = rPose
        self.poses.append(self.prevLaserPose.copy())
        self.frame_stats.append(stats)
        if consistent:                                             # :668-670 with TransformToEnd (:118-124)
            if stats is None:
                self.ctx.map_push(filtered_cloud, None, max_queue=self.max_queue_size)
            elif self.resident:
                self.ctx.map_push(filtered_cloud, from_last_register=True, max_queue=self.max_queue_size,
                                  transform_normals=self.map_transform_normals)
            else:
                self.ctx.map_push(filtered_cloud, stats["rPose"], max_queue=self.max_queue_size,
                                  transform_normals=self.map_transform_normals)
        else:
            self._target = self._accumulate(filtered_cloud)        # :668-670
        self.frameCount += 1
        return self.prevLaserPose.copy(), stats

    def run(self, frames):
        for f in frames:
            self.process_frame(f)
        return np.stack(self.poses) if self.poses else np.zeros((0, 4, 4))


def save_poses_tum(path: str, poses, timestamps=None):
    """savePoseToFile, src/saver.cpp:46-54: `timestamp tx ty tz qx qy qz qw`, 6 decimals."""
    with open(path, "w", encoding="utf-8") as f:
        for i, T in enumerate(poses):
            R = np.asarray(T)[:3, :3]
            t = np.asarray(T)[:3, 3]
            qw = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
            if qw > 1e-8:
                qx, qy, qz = (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)
            else:
                qx = np.sqrt(max(0.0, 1.0 + R[0, 0] - R[1, 1] - R[2, 2])) / 2.0
                qy = np.sqrt(max(0.0, 1.0 - R[0, 0] + R[1, 1] - R[2, 2])) / 2.0
                qz = np.sqrt(max(0.0, 1.0 - R[0, 0] - R[1, 1] + R[2, 2])) / 2.0
            ts = f"{i / 10.0:.6f}" if timestamps is None else str(timestamps[i])
            f.write(f"{ts} {t[0]:.6f} {t[1]:.6f} {t[2]:.6f} {qx:.6f} {qy:.6f} {qz:.6f} {qw:.6f}\n")
