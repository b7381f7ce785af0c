"""B200-native IMLS-ICP scan-to-map registration hot path.

Drop-in for the reference's IMLS matcher (include/imls_icp.h:45-147) + weighted-LS solver
(include/solver.h:92-98) inside the ICP loop of src/laser_odometry.cpp:524-647.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/plo/plo_c_api.h (csrc/libplo_cuda.so); this Python layer mirrors the reference's
matcher/solver/driver interface over ctypes.  There is no CPU fallback: using the compute
classes without the built library, or without a GPU, fails loudly.
"""
from . import synth  # noqa: F401
from . import _lib, config, distributed  # noqa: F401
from ._lib import PloError, PloFrontendParams, PloParams, default_params, frontend_default_params  # noqa: F401
from .context import Context  # noqa: F401
from .matcher import IMLSICPMatcher  # noqa: F401
from .odometry import LaserOdometry, save_poses_tum  # noqa: F401
from . import solver  # noqa: F401
from .solver import (SolveMotionEstimationProblemDRPM_CUDA, SolveMotionEstimationProblemLS_CUDA,  # noqa: F401
                     SolveMotionEstimationProblemRANSAC_CUDA, SolveMotionEstimationProblemWeightedLS_CUDA,
                     solveMotionEstimationProblem)

__all__ = ["synth", "config", "Context", "IMLSICPMatcher", "LaserOdometry", "PloError", "PloParams",
           "default_params", "frontend_default_params", "PloFrontendParams", "SolveMotionEstimationProblemWeightedLS_CUDA", "solveMotionEstimationProblem",
           "save_poses_tum"]
