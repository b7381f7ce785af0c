"""Spinning-LiDAR beam models used by the synthetic generators.

The elevation tables restate the ring-assignment rules of the reference's
front-end (src/scan_registration.cpp:926-929 bounds, :948-956 VLP-16,
:957-964 VLP-32C table, :990-996 HDL-64 two-block layout); the front-end itself
is out of scope — only its output contract (float32 xyz + unit normal per
point, 48-byte pcl::PointXYZINormal records, include/common.h:17) is reproduced.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class SensorModel:
    name: str
    elevations_deg: np.ndarray  # (n_beams,)
    azimuth_steps: int
    min_range: float = 2.0      # planetary_slam_VLP_32.launch:11
    max_range: float = 150.0    # planetary_slam_VLP_32.launch:13

    @property
    def n_beams(self) -> int:
        return int(self.elevations_deg.shape[0])

    def directions(self) -> tuple[np.ndarray, np.ndarray]:
        """Unit ray directions in the sensor frame, ring-major order.

        Returns (dirs[n_beams*azimuth_steps, 3] float64, ring[...] int32)."""
        el = np.deg2rad(self.elevations_deg.astype(np.float64))
        az = np.arange(self.azimuth_steps, dtype=np.float64) * (2.0 * np.pi / self.azimuth_steps)
        ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
        # velodyne spins clockwise seen from above: ori = -atan2(y, x) (scan_registration.cpp:1010)
        x = ce * np.cos(-az)[None, :]
        y = ce * np.sin(-az)[None, :]
        z = np.broadcast_to(se, x.shape)
        dirs = np.stack([x, y, z], axis=-1).reshape(-1, 3)
        ring = np.repeat(np.arange(self.n_beams, dtype=np.int32), self.azimuth_steps)
        return dirs, ring


def hdl64(azimuth_steps: int = 2083) -> SensorModel:
    # upper block: scanID = int((2 - angle)*3 + .5)  -> angle = 2 - id/3,    id 0..31
    # lower block: scanID = 32 + int((-8.83 - angle)*2 + .5) -> -8.83 - (id-32)/2, id 32..63
    upper = 2.0 - np.arange(32) / 3.0
    lower = -8.83 - np.arange(32) * 0.5
    return SensorModel("HDL-64", np.concatenate([upper, lower]), azimuth_steps)


def vlp32c(azimuth_steps: int = 1800) -> SensorModel:
    table = [-25.000, -15.639, -11.310, -8.843, -7.254, -6.148, -5.333, -4.667, -4.000,
             -3.667, -3.333, -3.000, -2.667, -2.333, -2.000, -1.667, -1.333, -1.000,
             -0.667, -0.333, 0.000, 0.333, 0.667, 1.000, 1.333, 1.667, 2.333]
    # the reference's table stops at 27 entries; the remaining five beams are the
    # VLP-32C data-sheet values
    table += [3.333, 4.667, 7.000, 10.333, 15.000]
    return SensorModel("VLP-32C", np.asarray(table), azimuth_steps)


def vlp16(azimuth_steps: int = 1800) -> SensorModel:
    return SensorModel("VLP-16", -15.0 + 2.0 * np.arange(16), azimuth_steps)


SENSORS = {"hdl64": hdl64, "vlp32c": vlp32c, "vlp16": vlp16}
