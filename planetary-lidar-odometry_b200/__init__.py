"""B200-native IMLS-ICP scan-to-map registration hot path.

Drop-in for the reference's IMLS matcher (include/imls_icp.h:45-147) + weighted-LS
solver (include/solver.h:92-98) inside the ICP loop of src/laser_odometry.cpp:524-647.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/plo/plo_c_api.h (csrc/libplo_cuda.so); this Python layer mirrors the reference's
matcher/solver/driver interface over ctypes.  There is no CPU fallback: importing the
compute modules without the built library, or calling them without a GPU, fails loudly.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
