"""ctypes binding of csrc/libplo_cuda.so (the C ABI of include/plo/plo_c_api.h).

The library is built in-tree (`make -C csrc`, nvcc, sm_100a only) and loaded from there;
nothing here falls back to a CPU implementation: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# PLO_LIB: another build of the SAME library (tuning variants from tools/build_variants.sh); never a different backend
LIB_PATH = os.environ.get("PLO_LIB") or os.path.join(CSRC, "libplo_cuda.so")

PLO_OK = 0
ERRORS = {-1: "INVALID_ARG", -2: "CUDA", -3: "NO_DEVICE", -4: "UNSUPPORTED", -5: "STATE"}
POINT_STATUS = ["ok", "no_normal", "too_far", "invalid_normal", "normal_constraint", "mls_fail", "nan_inf_height"]
REG_STATUS = {1: "CONVERGED", 2: "MAX_ITERS", 3: "TOO_FEW_PAIRS", 4: "SOLVE_FAILED"}
W_UNIT, W_HUBER_EXP = 0, 1
SOLVER_WLS, SOLVER_LS, SOLVER_RANSAC = 0, 1, 2
FINAL_LS, FINAL_WLS, FINAL_DRPM = 0, 1, 2


class PloError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"plo error {ERRORS.get(code, code)}: {msg}")
        self.code = code


class PloParams(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("h", C.c_double), ("r", C.c_double), ("r_normal", C.c_double),
        ("is_get_normals", C.c_int32), ("search_number_normal", C.c_int32), ("search_number", C.c_int32),
        ("normal_angle_constraint", C.c_int32), ("angle_diff_threshold", C.c_double),
        ("transform_normal", C.c_int32), ("correspond_number", C.c_int32),
        ("delta_dist_threshold", C.c_double), ("delta_angle_threshold", C.c_double),
        ("weight_mode", C.c_int32), ("ransac_distance_threshold", C.c_double), ("huber_threshold", C.c_double),
        ("solver", C.c_int32), ("ls_threshold", C.c_double),
        ("ransac_max_iterations", C.c_int32), ("ransac_min_inliers_percentage", C.c_double),
        ("ransac_final", C.c_int32), ("drpm_threshold", C.c_double), ("drpm_stdev_points", C.c_double),
        ("drpm_stdev_normals", C.c_double), ("ransac_seed", C.c_uint64),
    ]


class PloProjStats(C.Structure):
    _fields_ = [("n_source", C.c_int64), ("n_pairs", C.c_int64), ("dropped", C.c_int64 * 6)]


class PloRegStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("pairs", C.c_int64), ("rms", C.c_double),
                ("dropped", C.c_int64 * 6), ("delta_dist", C.c_double), ("delta_angle", C.c_double),
                ("rank", C.c_int32), ("reserved", C.c_int32)]


class PloFrontendParams(C.Structure):
    """plo_frontend_params (include/plo/plo_c_api.h)"""
    _fields_ = [("n_scans", C.c_int32), ("min_range", C.c_float), ("max_range", C.c_float), ("scan_period", C.c_float),
                ("window_size", C.c_int32), ("iter_step", C.c_int32), ("knn_distance_threshold", C.c_float),
                ("plane_distance_threshold", C.c_float), ("valid_points_threshold", C.c_float), ("use_all_points", C.c_int32),
                ("planarity_threshold", C.c_float)]


class PloFrontendStats(C.Structure):
    _fields_ = [("n_out", C.c_int64), ("gated", C.c_int64), ("ringed", C.c_int64), ("pca_failures", C.c_int64),
                ("plane_failures", C.c_int64), ("candidates", C.c_int64)]


EXPORTS = [
    "plo_create", "plo_destroy", "plo_last_error", "plo_version", "plo_set_stream", "plo_synchronize",
    "plo_default_params", "plo_set_params", "plo_set_target", "plo_set_source", "plo_set_target_device",
    "plo_set_source_device", "plo_target_size", "plo_source_size", "plo_project", "plo_get_pairs",
    "plo_get_neighbors", "plo_last_project_times", "plo_get_search_stats", "plo_get_query_results", "plo_get_target_normals", "plo_solve_wls", "plo_solve_ls", "plo_solve_ransac",
    "plo_solve_wls_host", "plo_get_normal_equations", "plo_register", "plo_register_batch",
    "plo_launch_count", "plo_last_timings", "plo_time_project_kernel", "plo_set_profiling",
    "plo_last_kernel_timings", "plo_map_reset", "plo_map_push", "plo_map_push_device", "plo_map_info", "plo_map_get",
    "plo_frontend_default_params", "plo_frontend", "plo_frontend_device", "plo_frontend_get", "plo_frontend_device_records",
    "plo_set_tuning", "plo_last_tile_misses", "plo_debug_loop_stamps", "plo_solve_ls_host", "plo_solve_ransac_host", "plo_solve_drpm_host",
    "plo_imls_height", "plo_compute_normal", "plo_stream_wait_event",
]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libplo_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-j4"] + (["-B"] if force else [])
    subprocess.check_call(args, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.plo_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.plo_destroy.argtypes = [vp]
    L.plo_destroy.restype = None
    L.plo_last_error.argtypes = [vp]
    L.plo_last_error.restype = C.c_char_p
    L.plo_set_stream.argtypes = [vp, vp]
    L.plo_synchronize.argtypes = [vp]
    L.plo_set_tuning.argtypes = [vp, C.c_char_p, i32]
    L.plo_stream_wait_event.argtypes = [vp, vp]
    L.plo_last_tile_misses.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.plo_debug_loop_stamps.argtypes = [vp, vp]
    L.plo_solve_ls_host.argtypes = [vp, vp, vp, vp, i64, C.c_double, vp, C.POINTER(i32)]
    L.plo_solve_ransac_host.argtypes = [vp, vp, vp, vp, i64, C.POINTER(PloParams), vp, vp, C.POINTER(i64), C.POINTER(i32)]
    L.plo_solve_drpm_host.argtypes = [vp, vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, vp, vp]
    L.plo_imls_height.argtypes = [vp, vp, i64, vp, vp]
    L.plo_compute_normal.argtypes = [vp, vp, i64, vp]
    L.plo_default_params.argtypes = [C.POINTER(PloParams)]
    L.plo_default_params.restype = None
    L.plo_set_params.argtypes = [vp, C.POINTER(PloParams)]
    for f in (L.plo_set_target, L.plo_set_source, L.plo_set_target_device, L.plo_set_source_device):
        f.argtypes = [vp, vp, i64, i32]
    for f in (L.plo_target_size, L.plo_source_size):
        f.argtypes = [vp]
        f.restype = i64
    L.plo_project.argtypes = [vp, vp, i32, C.POINTER(PloProjStats)]
    L.plo_get_pairs.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(i64)]
    L.plo_get_neighbors.argtypes = [vp, vp, vp, vp, vp]
    L.plo_get_query_results.argtypes = [vp, vp, vp]
    L.plo_get_search_stats.argtypes = [vp, vp]
    L.plo_get_target_normals.argtypes = [vp, vp]
    L.plo_solve_wls.argtypes = [vp, vp, C.POINTER(i32)]
    L.plo_solve_ls.argtypes = [vp, vp, C.POINTER(i32)]
    L.plo_solve_ransac.argtypes = [vp, vp, vp, C.POINTER(i64), C.POINTER(i32)]
    L.plo_solve_wls_host.argtypes = [vp, vp, vp, vp, vp, i64, vp, C.POINTER(i32)]
    L.plo_get_normal_equations.argtypes = [vp, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.plo_register.argtypes = [vp, vp, vp, C.POINTER(PloRegStats)]
    L.plo_register_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, vp, vp]
    L.plo_launch_count.argtypes = [vp]
    L.plo_launch_count.restype = i64
    L.plo_last_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.plo_time_project_kernel.argtypes = [vp, vp, i32, C.POINTER(C.c_float)]
    L.plo_set_profiling.argtypes = [vp, i32]
    L.plo_last_kernel_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.plo_last_project_times.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.plo_map_reset.argtypes = [vp]
    L.plo_map_push.argtypes = [vp, vp, i64, i32, vp, i32, i32, i32]
    L.plo_map_push_device.argtypes = [vp, vp, i64, i32, vp, i32, i32, i32]
    L.plo_map_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64)]
    L.plo_map_get.argtypes = [vp, vp, i64]
    L.plo_frontend_default_params.argtypes = [C.POINTER(PloFrontendParams)]
    L.plo_frontend_default_params.restype = None
    L.plo_frontend.argtypes = [vp, vp, i64, i32, C.POINTER(PloFrontendParams), C.POINTER(PloFrontendStats)]
    L.plo_frontend_device.argtypes = [vp, vp, i64, i32, C.POINTER(PloFrontendParams), C.POINTER(PloFrontendStats)]
    L.plo_frontend_get.argtypes = [vp, vp, vp, vp, vp, i64]
    L.plo_frontend_device_records.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    _lib = L
    return L


def frontend_default_params(**over) -> PloFrontendParams:
    p = PloFrontendParams()
    lib().plo_frontend_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(f"plo_frontend_params has no field {k!r}")
        setattr(p, k, v)
    return p


def default_params(**over) -> PloParams:
    p = PloParams()
    lib().plo_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p
