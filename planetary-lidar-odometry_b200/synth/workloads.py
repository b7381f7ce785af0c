"""Concrete synthetic inputs for BASELINE.json's configs.

Every builder is a pure function of its seed; the same bytes go to the oracle and
to the CUDA path.  Clouds are (n, 12) float32 arrays = 48-byte PointXYZINormal
records.  A *pair* is (source, target, T_gt) with T_gt the 4x4 that maps source
coordinates into target coordinates (what registration should recover).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import scenes, sensors


@dataclasses.dataclass
class Pair:
    name: str
    source: np.ndarray
    target: np.ndarray
    T_gt: np.ndarray


def _subsample(rec: np.ndarray, n: int | None, rng: np.random.Generator) -> np.ndarray:
    if n is None or rec.shape[0] <= n:
        return rec
    keep = np.sort(rng.choice(rec.shape[0], size=n, replace=False))
    return rec[keep]


def hdl64_pair(seed: int = 1001, azimuth_steps: int = 2083, max_source: int | None = None,
               max_target: int | None = None) -> Pair:
    """cfg-1: frame B (moved by t=(0.8,0.05,0.01) m, yaw 1.5 deg, pitch 0.2 deg) against frame A."""
    rng = np.random.default_rng(seed)
    scene = scenes.urban_scene(seed)
    sensor = sensors.hdl64(azimuth_steps)
    T_a = scenes.pose_matrix([0, 0, 0])
    T_b = scenes.pose_matrix([0.8, 0.05, 0.01], yaw_deg=1.5, pitch_deg=0.2)
    tgt = scenes.scan(scene, sensor, T_a, rng)
    src = scenes.scan(scene, sensor, T_b, rng)
    return Pair("hdl64_pair", _subsample(src, max_source, rng), _subsample(tgt, max_target, rng),
                np.linalg.inv(T_a) @ T_b)


def hdl64_vs_map(seed: int = 1002, map_points: int = 1_000_000, voxel: float = 0.02,
                 azimuth_steps: int = 2083, frame_spacing: float = 1.0,
                 max_source: int | None = None) -> Pair:
    """north-star headline (cfg-1b) / cfg-4: one HDL-64 frame against a local map made of
    consecutive frames `frame_spacing` m apart (alternating ahead/behind frame A),
    expressed in frame-A coordinates and voxel-deduplicated at `voxel` m; frames are added
    until `map_points` survive, then the map is cut to exactly `map_points` (seeded choice,
    original order kept).  SURVEY.md §8d asked for a 5 cm dedup of 8 frames, which leaves
    only ~0.48 M points; 2 cm (the cfg-4 value) reaches 1.00 M with ~10 frames."""
    rng = np.random.default_rng(seed)
    scene = scenes.urban_scene(seed)
    sensor = sensors.hdl64(azimuth_steps)
    T_a = scenes.pose_matrix([0, 0, 0])
    parts = []
    raw = 0
    total = None
    i = 0
    while True:
        off = ((i + 1) // 2) * frame_spacing * (1 if i % 2 == 1 else -1) if i else 0.0
        T_i = scenes.pose_matrix([off, 0.02 * np.sin(0.7 * i), 0.0], yaw_deg=0.3 * np.sin(0.5 * i))
        parts.append(scenes.scan(scene, sensor, T_i, rng, out_frame=T_a))
        raw += parts[-1].shape[0]
        i += 1
        if raw < 1.05 * map_points and i < 400:
            continue
        total = scenes.voxel_dedup(np.concatenate(parts, axis=0), voxel)
        if total.shape[0] >= map_points or i >= 400:
            break
    tgt = _subsample(total, map_points, rng)
    T_b = scenes.pose_matrix([0.8, 0.05, 0.01], yaw_deg=1.5, pitch_deg=0.2)
    src = scenes.scan(scene, sensor, T_b, rng)
    return Pair(f"hdl64_vs_map{tgt.shape[0]}", _subsample(src, max_source, rng), tgt, np.linalg.inv(T_a) @ T_b)


def hdl64_vs_dense_map(seed: int = 4001, map_points: int = 5_000_000, base_points: int = 1_000_000,
                       jitter: float = 0.01) -> Pair:
    """cfg-4 stress input: HDL-64 frame against a `map_points`-point accumulated map.  Ray casting 5 M
    deduplicated points would take minutes, so the map is the `base_points` local map plus jittered
    copies of it (sigma = `jitter` m, normals kept): same surfaces, 5x the density — the regime that
    stresses the index build and the radius search."""
    base = hdl64_vs_map(seed=seed, map_points=base_points)
    rng = np.random.default_rng(seed + 7)
    reps = int(np.ceil(map_points / base.target.shape[0]))
    parts = [base.target]
    for _ in range(reps - 1):
        c = base.target.copy()
        c[:, 0:3] += rng.normal(0.0, jitter, size=(c.shape[0], 3)).astype(np.float32)
        parts.append(c)
    tgt = np.concatenate(parts, axis=0)[:map_points]
    return Pair(f"hdl64_vs_dense_map{tgt.shape[0]}", base.source, tgt, base.T_gt)


def planetary_pair(seed: int = 3001, azimuth_steps: int = 1800) -> Pair:
    """cfg-3: sparse VLP-16 scans of fractal terrain with rocks (neighbour-starved path)."""
    rng = np.random.default_rng(seed)
    scene = scenes.planetary_scene(seed)
    sensor = sensors.vlp16(azimuth_steps)
    T_a = scenes.pose_matrix([0, 0, 0])
    T_b = scenes.pose_matrix([0.5, -0.04, 0.02], yaw_deg=-1.0, pitch_deg=0.3, roll_deg=-0.2)
    tgt = scenes.scan(scene, sensor, T_a, rng)
    src = scenes.scan(scene, sensor, T_b, rng)
    return Pair("vlp16_planetary_pair", src, tgt, np.linalg.inv(T_a) @ T_b)


def trajectory(n_frames: int, seed: int) -> list[np.ndarray]:
    """Smooth S-curve, 0.5–1.0 m per frame, <= 2 deg per frame (cfg-2 / cfg-5)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=3)
    poses = []
    x = y = 0.0
    yaw = 0.0
    for k in range(n_frames):
        step = 0.75 + 0.25 * np.sin(0.05 * k + ph[0])
        yaw_rate = 1.5 * np.sin(0.03 * k + ph[1])          # deg / frame
        yaw += yaw_rate
        x += step * np.cos(np.deg2rad(yaw))
        y += step * np.sin(np.deg2rad(yaw))
        pitch = 0.3 * np.sin(0.11 * k + ph[2])
        poses.append(scenes.pose_matrix([x, y, 0.02 * np.sin(0.2 * k)], yaw_deg=yaw, pitch_deg=pitch))
    return poses


class Sequence:
    """cfg-2 / cfg-5: a VLP-32C-shaped odometry sequence; frames are generated on demand.

    The scene tiles with the trajectory: boxes are re-seeded every 100 m so that every
    frame sees structure (the ground is an infinite analytic surface)."""

    def __init__(self, seed: int = 2001, n_frames: int = 1000, sensor: str = "vlp32c",
                 azimuth_steps: int = 1800, max_points: int | None = None):
        self.seed = seed
        self.n_frames = n_frames
        self.sensor = sensors.SENSORS[sensor](azimuth_steps)
        self.poses = trajectory(n_frames, seed)
        self.max_points = max_points
        self._base = scenes.urban_scene(seed)

    def _scene_at(self, T: np.ndarray) -> scenes.Scene:
        # periodic world (period 100 m): the 3x3 tiles around the sensor are instantiated,
        # so consecutive frames always see the same geometry within ~100 m
        cx = 100.0 * np.round(T[0, 3] / 100.0)
        cy = 100.0 * np.round(T[1, 3] / 100.0)
        tiles = []
        for dx in (-100.0, 0.0, 100.0):
            for dy in (-100.0, 0.0, 100.0):
                b = self._base.boxes.copy()
                b[-2:, 0] = -50.0          # walls span exactly one period
                b[-2:, 3] = 50.0
                b[:, [0, 3]] += cx + dx
                b[:, [1, 4]] += cy + dy
                tiles.append(b)
        return scenes.Scene(self._base.ground, np.concatenate(tiles, axis=0), self._base.spheres)

    def frame(self, k: int) -> np.ndarray:
        rng = np.random.default_rng([self.seed, k])
        rec = scenes.scan(self._scene_at(self.poses[k]), self.sensor, self.poses[k], rng)
        return _subsample(rec, self.max_points, rng)

    def relative_gt(self, k: int) -> np.ndarray:
        """T that maps frame k coordinates into frame k-1 coordinates."""
        return np.linalg.inv(self.poses[k - 1]) @ self.poses[k]


def rigid_copy_pair(seed: int = 7, n: int = 4000, noise: float = 0.0) -> Pair:
    """Known-answer case: source = rigidly transformed copy of the target,
    sampled on a plane + two walls; the recovered pose must be the inverse transform."""
    rng = np.random.default_rng(seed)
    n3 = n // 3
    g = np.zeros((n3, 12), np.float32)
    g[:, 0:2] = rng.uniform(-10, 10, size=(n3, 2))
    g[:, 2] = -1.5
    g[:, 4:7] = [0, 0, 1]
    w1 = np.zeros((n3, 12), np.float32)
    w1[:, 0] = 8.0
    w1[:, 1] = rng.uniform(-10, 10, size=n3)
    w1[:, 2] = rng.uniform(-1.5, 3, size=n3)
    w1[:, 4:7] = [-1, 0, 0]
    w2 = np.zeros((n - 2 * n3, 12), np.float32)
    w2[:, 1] = 9.0
    w2[:, 0] = rng.uniform(-10, 10, size=w2.shape[0])
    w2[:, 2] = rng.uniform(-1.5, 3, size=w2.shape[0])
    w2[:, 4:7] = [0, -1, 0]
    tgt = np.concatenate([g, w1, w2], axis=0)
    tgt[:, 3] = 1.0
    T = scenes.pose_matrix([0.12, -0.07, 0.03], yaw_deg=0.8, pitch_deg=-0.3, roll_deg=0.2)
    Ti = np.linalg.inv(T)
    src = tgt.copy()
    p = tgt[:, 0:3].astype(np.float64)
    if noise > 0:
        p = p + rng.normal(0, noise, size=p.shape)
    src[:, 0:3] = (p @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    src[:, 4:7] = (tgt[:, 4:7].astype(np.float64) @ Ti[:3, :3].T).astype(np.float32)
    return Pair("rigid_copy", src, tgt, T)


# ---- materialised sequences (cfg-2 at length, cfg-5): frames generated once, in parallel ---------------

class FrameSet:
    """A sequence whose frames are already in memory (same interface as `Sequence`)."""

    def __init__(self, seed: int, poses, frames=None):
        self.seed = seed
        self.poses = poses
        self.n_frames = len(poses)
        self._frames = frames          # list of (n, 12) float32 arrays, or None for a shard this process does not own

    def frame(self, k: int) -> np.ndarray:
        if self._frames is None:
            raise RuntimeError(f"sequence {self.seed}: frames were not generated in this process (not its shard)")
        return self._frames[k]

    def relative_gt(self, k: int) -> np.ndarray:
        return np.linalg.inv(self.poses[k - 1]) @ self.poses[k]


def _gen_frame(job):
    seed, n_frames, sensor, k = job
    return Sequence(seed=seed, n_frames=n_frames, sensor=sensor).frame(k)


def generate_sequences(seeds, n_frames: int, own=None, sensor: str = "vlp32c", workers: int = 0) -> list[FrameSet]:
    """`FrameSet`s for `seeds`; frames are ray-cast (in `workers` forked processes) only for the indices in
    `own` (default: all) — the other entries carry poses and length only.  Call before CUDA is initialised
    in this process (fork).  Every frame is a pure function of (seed, k): identical bytes whatever the split."""
    import multiprocessing as mp
    import os
    own = list(range(len(seeds))) if own is None else list(own)
    jobs = [(seeds[i], n_frames, sensor, k) for i in own for k in range(n_frames)]
    if workers <= 0:
        workers = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(workers, len(jobs)))
    if workers == 1 or not jobs:
        frames = [_gen_frame(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            frames = pool.map(_gen_frame, jobs, chunksize=1)
    out = []
    it = iter(frames)
    for i, s in enumerate(seeds):
        poses = trajectory(n_frames, s)
        out.append(FrameSet(s, poses, [next(it) for _ in range(n_frames)] if i in own else None))
    return out
