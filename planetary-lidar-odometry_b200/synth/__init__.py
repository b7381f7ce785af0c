"""Synthetic sensor/scene generators standing in for the reference's front-end
(src/scan_registration.cpp), which is out of scope."""
from . import scenes, sensors, workloads  # noqa: F401
from .scenes import POINT_FLOATS, POINT_STRIDE  # noqa: F401
