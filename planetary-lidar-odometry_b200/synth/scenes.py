"""Analytic scenes + ray casting for the synthetic inputs of SURVEY.md §8d.

Output contract = what the reference's front-end publishes on
/laser_cloud_filtered (src/scan_registration.cpp:1196-1222): one 48-byte
pcl::PointXYZINormal record per return — float32 xyz at byte 0, unit normal
(oriented +z) at byte 16, intensity = ring + 0.1*relTime at byte 32
(src/scan_registration.cpp:1041-1042).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .sensors import SensorModel

POINT_STRIDE = 48          # bytes, pcl::PointXYZINormal
POINT_FLOATS = POINT_STRIDE // 4


@dataclasses.dataclass
class Heightfield:
    """z = base + sum_k amp_k * sin(kx_k*x + px_k) * cos(ky_k*y + py_k)."""
    base: float
    amp: np.ndarray
    kx: np.ndarray
    ky: np.ndarray
    px: np.ndarray
    py: np.ndarray

    def z(self, x, y):
        out = np.full(np.shape(x), self.base, dtype=np.float64)
        for a, kx, ky, px, py in zip(self.amp, self.kx, self.ky, self.px, self.py):
            out = out + a * np.sin(kx * x + px) * np.cos(ky * y + py)
        return out

    def grad(self, x, y):
        gx = np.zeros(np.shape(x), dtype=np.float64)
        gy = np.zeros(np.shape(x), dtype=np.float64)
        for a, kx, ky, px, py in zip(self.amp, self.kx, self.ky, self.px, self.py):
            gx = gx + a * kx * np.cos(kx * x + px) * np.cos(ky * y + py)
            gy = gy - a * ky * np.sin(kx * x + px) * np.sin(ky * y + py)
        return gx, gy


@dataclasses.dataclass
class Scene:
    ground: Heightfield
    boxes: np.ndarray      # (nb, 6) xmin ymin zmin xmax ymax zmax
    spheres: np.ndarray    # (ns, 4) cx cy cz r

    def cast(self, origin: np.ndarray, dirs: np.ndarray, tmin: float, tmax: float):
        """First hit per ray.  Returns (t, normal[K,3]); t = inf where nothing is hit."""
        K = dirs.shape[0]
        t_best = np.full(K, np.inf)
        n_best = np.zeros((K, 3))
        # --- ground: bracket by the terrain's amplitude band, march, bisect ---------
        amp = float(np.sum(np.abs(self.ground.amp)))
        z_hi, z_lo = self.ground.base + amp, self.ground.base - amp
        n_s = 48
        frac = (np.arange(n_s) / (n_s - 1.0))[None, :]
        down = np.nonzero(dirs[:, 2] < -1e-9)[0]
        chunk = 1 << 16
        for s in range(0, down.shape[0], chunk):
            sel = down[s:s + chunk]
            d = dirs[sel]
            ta = np.maximum((origin[2] - z_hi) / (-d[:, 2]), tmin)   # first possible contact
            tb = np.minimum((origin[2] - z_lo) / (-d[:, 2]), tmax * 1.05)
            ok = tb > ta
            ts = ta[:, None] + (tb - ta)[:, None] * frac
            f = (origin[2] + d[:, 2:3] * ts) - self.ground.z(origin[0] + d[:, 0:1] * ts, origin[1] + d[:, 1:2] * ts)
            below = f <= 0.0
            first = np.argmax(below, axis=1)
            rows = np.arange(d.shape[0])
            has = below[rows, first] & ok
            lo = np.where(first > 0, ts[rows, np.maximum(first - 1, 0)], ta)
            hi = ts[rows, first]
            for _ in range(26):
                mid = 0.5 * (lo + hi)
                fm = (origin[2] + d[:, 2] * mid) - self.ground.z(origin[0] + d[:, 0] * mid, origin[1] + d[:, 1] * mid)
                neg = fm <= 0.0
                hi = np.where(neg, mid, hi)
                lo = np.where(neg, lo, mid)
            tg = np.where(has, 0.5 * (lo + hi), np.inf)
            tsafe = np.where(has, tg, 0.0)
            gx, gy = self.ground.grad(origin[0] + d[:, 0] * tsafe, origin[1] + d[:, 1] * tsafe)
            nn = np.stack([-gx, -gy, np.ones_like(gx)], axis=1)
            nn /= np.linalg.norm(nn, axis=1, keepdims=True)
            t_best[sel] = tg
            n_best[sel] = nn
        # --- boxes: slab method -----------------------------------------------------
        # Per box only the rays inside the azimuth interval its footprint subtends from the origin are tested
        # (a hit point lies in the footprint, so its ray does too); the arithmetic per tested ray is unchanged,
        # i.e. the result is bit-identical to testing every ray against every box.
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / dirs
        az = np.arctan2(dirs[:, 1], dirs[:, 0])
        steep = np.hypot(dirs[:, 0], dirs[:, 1]) < 1e-6
        for b in self.boxes:
            if b[0] <= origin[0] <= b[3] and b[1] <= origin[1] <= b[4]:
                idx = np.arange(K)
            else:
                ac = np.arctan2(0.5 * (b[1] + b[4]) - origin[1], 0.5 * (b[0] + b[3]) - origin[0])
                ca = np.arctan2(np.array([b[1], b[1], b[4], b[4]]) - origin[1], np.array([b[0], b[3], b[0], b[3]]) - origin[0])
                rel_c = (ca - ac + np.pi) % (2 * np.pi) - np.pi
                rel = (az - ac + np.pi) % (2 * np.pi) - np.pi
                idx = np.nonzero(((rel >= rel_c.min() - 1e-6) & (rel <= rel_c.max() + 1e-6)) | steep)[0]
            if idx.shape[0] == 0:
                continue
            Ks = idx.shape[0]
            rows = np.arange(Ks)
            inv_s = inv[idx]
            t0 = (b[0:3][None, :] - origin[None, :]) * inv_s
            t1 = (b[3:6][None, :] - origin[None, :]) * inv_s
            tn = np.minimum(t0, t1)
            tf = np.maximum(t0, t1)
            axis = np.argmax(tn, axis=1)
            tnear = tn[rows, axis]
            tfar = np.min(tf, axis=1)
            hit = (tnear <= tfar) & (tnear > tmin) & (tnear < t_best[idx])
            if not hit.any():
                continue
            nb = np.zeros((Ks, 3))
            nb[rows, axis] = -np.sign(dirs[idx, axis])
            t_best[idx] = np.where(hit, tnear, t_best[idx])
            n_best[idx] = np.where(hit[:, None], nb, n_best[idx])
        # --- spheres (rocks) --------------------------------------------------------
        for sp in self.spheres:
            oc = origin - sp[0:3]
            bq = dirs @ oc
            cq = oc @ oc - sp[3] * sp[3]
            disc = bq * bq - cq
            with np.errstate(invalid="ignore"):
                ts_ = -bq - np.sqrt(disc)
            hit = (disc > 0) & (ts_ > tmin) & (ts_ < t_best)
            hp = origin[None, :] + dirs * np.where(hit, ts_, 0.0)[:, None]
            ns = (hp - sp[0:3][None, :]) / sp[3]
            t_best = np.where(hit, ts_, t_best)
            n_best = np.where(hit[:, None], ns, n_best)
        t_best = np.where(t_best <= tmax, t_best, np.inf)
        return t_best, n_best


def urban_scene(seed: int) -> Scene:
    """cfg-1 scene: bumpy ground z=0.3 sin(0.2x)cos(0.15y) (sensor ~1.8 m above it),
    12 axis-aligned boxes and 2 long walls."""
    rng = np.random.default_rng(seed)
    ground = Heightfield(-1.8, np.array([0.3]), np.array([0.2]), np.array([0.15]), np.zeros(1), np.zeros(1))
    boxes = []
    for _ in range(12):
        r = rng.uniform(8.0, 45.0)
        a = rng.uniform(0, 2 * np.pi)
        cx, cy = r * np.cos(a), r * np.sin(a)
        sx, sy, sz = rng.uniform(1.5, 6.0), rng.uniform(1.5, 6.0), rng.uniform(1.5, 5.0)
        boxes.append([cx - sx / 2, cy - sy / 2, -2.5, cx + sx / 2, cy + sy / 2, -1.8 + sz])
    boxes.append([-70.0, 14.0, -2.5, 90.0, 14.6, 2.5])     # two long walls either side of the track
    boxes.append([-70.0, -17.6, -2.5, 90.0, -17.0, 3.0])
    return Scene(ground, np.asarray(boxes), np.zeros((0, 4)))


def planetary_scene(seed: int) -> Scene:
    """cfg-3 scene: fractal (fBm-like) terrain, amplitude ~1.5 m, sparse rocks, no walls."""
    rng = np.random.default_rng(seed)
    octaves = 5
    lam0 = 60.0
    amp, kx, ky, px, py = [], [], [], [], []
    for o in range(octaves):
        lam = lam0 / (2.0 ** o)
        th = rng.uniform(0, 2 * np.pi)
        k = 2 * np.pi / lam
        amp.append(1.5 * 0.5 ** o * 0.55)
        kx.append(k * np.cos(th))
        ky.append(k * np.sin(th))
        px.append(rng.uniform(0, 2 * np.pi))
        py.append(rng.uniform(0, 2 * np.pi))
    ground = Heightfield(-1.8, np.asarray(amp), np.asarray(kx), np.asarray(ky), np.asarray(px), np.asarray(py))
    spheres = []
    for _ in range(25):
        r = rng.uniform(6.0, 60.0)
        a = rng.uniform(0, 2 * np.pi)
        cx, cy = r * np.cos(a), r * np.sin(a)
        rad = rng.uniform(0.3, 1.2)
        cz = float(ground.z(np.asarray(cx), np.asarray(cy))) + 0.3 * rad
        spheres.append([cx, cy, cz, rad])
    return Scene(ground, np.zeros((0, 6)), np.asarray(spheres))


def pose_matrix(t, yaw_deg=0.0, pitch_deg=0.0, roll_deg=0.0) -> np.ndarray:
    """Sensor->world 4x4 (R = Rz(yaw) Ry(pitch) Rx(roll))."""
    y, p, r = np.deg2rad([yaw_deg, pitch_deg, roll_deg])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T


def scan(scene: Scene, sensor: SensorModel, T_ws: np.ndarray, rng: np.random.Generator,
         sigma_range: float = 0.02, out_frame: np.ndarray | None = None) -> np.ndarray:
    """One sweep from sensor pose T_ws (sensor->world).  Points/normals are
    expressed in `out_frame` (world->frame 4x4 inverse is applied; default: the
    sensor's own frame, as the reference's clouds are).  Returns (n, 12) float32
    records (48 bytes each)."""
    dirs_s, ring = sensor.directions()
    R, o = T_ws[:3, :3], T_ws[:3, 3]
    dirs_w = dirs_s @ R.T
    t, n_w = scene.cast(o, dirs_w, sensor.min_range, sensor.max_range)
    hit = np.isfinite(t)
    t = t[hit]
    t_noisy = t + rng.normal(0.0, sigma_range, size=t.shape) if sigma_range > 0 else t
    keep = (t_noisy >= sensor.min_range) & (t_noisy <= sensor.max_range)
    p_w = o[None, :] + dirs_w[hit][keep] * t_noisy[keep, None]
    n_w = n_w[hit][keep]
    ring = ring[hit][keep]
    rel = (np.nonzero(hit)[0][keep] % sensor.azimuth_steps) / float(sensor.azimuth_steps)
    T_fw = np.linalg.inv(T_ws if out_frame is None else out_frame)
    p_f = p_w @ T_fw[:3, :3].T + T_fw[:3, 3][None, :]
    n_f = n_w @ T_fw[:3, :3].T
    # orientation: +z (scan_registration.cpp:1196-1200); near-vertical faces are oriented
    # towards the viewpoint so both frames of a pair agree on the sign
    o_f = T_fw[:3, :3] @ o + T_fw[:3, 3]
    flip_z = n_f[:, 2] < 0
    vertical = np.abs(n_f[:, 2]) < 0.05
    to_view = np.einsum("ij,ij->i", n_f, o_f[None, :] - p_f)
    flip = np.where(vertical, to_view < 0, flip_z)
    n_f = np.where(flip[:, None], -n_f, n_f)
    rec = np.zeros((p_f.shape[0], POINT_FLOATS), dtype=np.float32)
    rec[:, 0:3] = p_f.astype(np.float32)
    rec[:, 3] = 1.0
    rec[:, 4:7] = n_f.astype(np.float32)
    rec[:, 8] = (ring + 0.1 * rel).astype(np.float32)
    return rec


def voxel_dedup(rec: np.ndarray, voxel: float) -> np.ndarray:
    """Keep the first record per `voxel`-sized cell (stable)."""
    q = np.floor(rec[:, 0:3].astype(np.float64) / voxel).astype(np.int64)
    q -= q.min(axis=0, keepdims=True)
    ext = q.max(axis=0) + 1
    key = (q[:, 0] * ext[1] + q[:, 1]) * ext[2] + q[:, 2]
    _, first = np.unique(key, return_index=True)
    return rec[np.sort(first)]
