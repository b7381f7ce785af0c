"""config.json handling — the reference's selection surface.

The reference loads one global json (src/common.cpp:8-17) and dispatches on strings
(`matching_method.method` at src/laser_odometry.cpp:491,557,561; `solve_method.method` at
:606 -> :173-275).  The same file is read here, once, validated, and flattened into the
POD `plo_params` that crosses the C ABI.  Unknown method strings fail loudly (the reference
only prints, :567 and :271).

Accepted selections
  matching_method.method : "IMLS" (with top-level "backend": "cuda", the default here) or
                           "IMLS_CUDA"
  solve_method.method    : "WeightedLS_CUDA" / "Weighted LS" (weighted LS, unit weights unless
                           "weights": "huber_exp"), "LS" / "LS_CUDA" (the reference's trimmed LS,
                           src/solver.cpp:74-166, threshold from solve_method.LS.threshold), or
                           "RANSAC" (the config.json default chain, src/solver.cpp:222-385) with
                           final_solve_method "LS" (trim fraction RANSAC.LS_threshold, src/laser_odometry.cpp:205),
                           "Weighted LS" or "DRPM" (src/solver.cpp:499-603)
Everything else the reference lists (plane_ICP, Ceres, ICP, Teaser,
tensor voting, projected distance) is outside the hot-path scope and raises.
"""
from __future__ import annotations

import copy
import json

from . import _lib

# config.json of the reference, laser_odometry section, verbatim defaults (config.json:83-169),
# with the solver selection switched to the in-scope weighted LS
DEFAULT_CONFIG = {
    "backend": "cuda",
    "laser_odometry": {
        "max_queue_size": 1,
        "transform_normal": False,
        "matching_method": {
            "method": "IMLS",
            "correspond_number": 6,
            "IMLS": {
                "h": 1, "r": 3,
                "use_tensor_voting": {"enabled": False, "k": 50, "sigma": 0.2, "distance_threshold": 0.6},
                "get_normals": {"enabled": True, "r_normal": 1, "search_number_normal": 10},
                "use_projected_distance": {"enabled": False, "r_proj": 0.8},
                "normal_angle_constraint": {"enabled": True, "angle_diff_threshold": 30},
                "IMLS function": {"search_number": 20},
            },
        },
        "solve_method": {
            "method": "WeightedLS_CUDA",
            "iterations": 30,
            "delta_dist_threshold": 0.001,
            "delta_angle_threshold": 0.0001745353,
            "LS": {"threshold": 0.02},
            "RANSAC": {"max_iterations": 5000, "distance_threshold": 0.8, "min_inliers_percentage": 0.95, "huber_threshold": 0.648,
                       "final_solve_method": "Weighted LS", "LS_threshold": 0.02, "DRPM_threshold": 0.05,
                       "DRPM_stdev_points": 0.02, "DRPM_stdev_normals": 0.05},
        },
    },
}


class ConfigError(ValueError):
    pass


# final_solve_method strings of SolveMotionEstimationProblemRANSAC (src/solver.cpp:366-384)
RANSAC_FINALS = {"LS": _lib.FINAL_LS, "Weighted LS": _lib.FINAL_WLS, "DRPM": _lib.FINAL_DRPM}


def load_config(path: str | None = None) -> dict:
    """loadConfig(), src/common.cpp:8-17 — but with an explicit path and a loud failure."""
    if path is None:
        return copy.deepcopy(DEFAULT_CONFIG)
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def _get(d, *keys, default=None, required=True):
    cur = d
    for k in keys:
        if not isinstance(cur, dict) or k not in cur:
            if required and default is None:
                raise ConfigError("config.json: missing key " + ".".join(keys))
            return default
        cur = cur[k]
    return cur


def params_from_config(cfg: dict) -> _lib.PloParams:
    backend = cfg.get("backend", "cuda")
    lo = _get(cfg, "laser_odometry")
    mm = _get(lo, "matching_method")
    method = _get(mm, "method")
    if method == "IMLS_CUDA" or (method == "IMLS" and backend == "cuda"):
        pass
    elif method == "IMLS":
        raise ConfigError(f'matching_method "IMLS" with backend "{backend}": only the cuda backend exists here (no CPU fallback)')
    elif method == "plane_ICP":
        raise ConfigError('matching_method "plane_ICP" is outside the hot-path scope (SURVEY.md §2.1 row 4)')
    else:
        raise ConfigError(f"Invalid MATCHING_METHOD! ({method!r})")
    im = _get(mm, "IMLS")
    if _get(im, "use_tensor_voting", "enabled", default=False, required=False):
        raise ConfigError("use_tensor_voting is outside the hot-path scope (SURVEY.md §8a a13)")
    if _get(im, "use_projected_distance", "enabled", default=False, required=False):
        raise ConfigError("use_projected_distance is outside the hot-path scope (SURVEY.md §8a a13)")
    sm = _get(lo, "solve_method")
    smethod = _get(sm, "method")
    weight_mode = _lib.W_UNIT
    solver = _lib.SOLVER_WLS
    ransac_final = _lib.FINAL_DRPM
    if smethod in ("WeightedLS_CUDA", "Weighted LS"):
        weight_mode = _lib.W_HUBER_EXP if _get(sm, "weights", default="unit", required=False) == "huber_exp" else _lib.W_UNIT
    elif smethod in ("LS", "LS_CUDA"):
        solver = _lib.SOLVER_LS
    elif smethod == "RANSAC":
        solver = _lib.SOLVER_RANSAC
        final = _get(sm, "RANSAC", "final_solve_method")
        if final == "LS":
            ransac_final = _lib.FINAL_LS
        elif final == "Weighted LS":
            ransac_final = _lib.FINAL_WLS
        elif final == "DRPM":
            ransac_final = _lib.FINAL_DRPM
        else:   # the reference prints "Invalid FINAL_SOLVE_METHOD in RANSAC!" and returns false (src/solver.cpp:380-384)
            raise ConfigError(f'solve_method RANSAC -> unknown final_solve_method "{final}"')
    elif smethod in ("Ceres", "ICP", "Teaser"):
        raise ConfigError(f'solve_method "{smethod}" is outside the hot-path scope (SURVEY.md §2.1 row 2)')
    else:
        raise ConfigError(f"Invalid SOLVE_METHOD! ({smethod!r})")
    rs = _get(sm, "RANSAC", default={}, required=False) or {}
    ls_cfg = _get(sm, "LS", default={}, required=False) or {}
    return _lib.default_params(
        iterations=int(_get(sm, "iterations")),
        h=float(_get(im, "h")), r=float(_get(im, "r")),
        r_normal=float(_get(im, "get_normals", "r_normal")),
        is_get_normals=int(bool(_get(im, "get_normals", "enabled"))),
        search_number_normal=int(_get(im, "get_normals", "search_number_normal")),
        search_number=int(_get(im, "IMLS function", "search_number")),
        normal_angle_constraint=int(bool(_get(im, "normal_angle_constraint", "enabled"))),
        angle_diff_threshold=float(_get(im, "normal_angle_constraint", "angle_diff_threshold")),
        transform_normal=int(bool(_get(lo, "transform_normal", default=False, required=False))),
        correspond_number=int(_get(mm, "correspond_number")),
        delta_dist_threshold=float(_get(sm, "delta_dist_threshold")),
        delta_angle_threshold=float(_get(sm, "delta_angle_threshold")),
        weight_mode=weight_mode,
        ransac_distance_threshold=float(rs.get("distance_threshold", 0.8)),
        huber_threshold=float(rs.get("huber_threshold", 0.648)),
        solver=solver,
        # the trim fraction of the final "LS" inside RANSAC is its own key (src/laser_odometry.cpp:205 -> src/solver.cpp:366-371)
        ls_threshold=float(rs.get("LS_threshold", ls_cfg.get("threshold", 0.02)) if solver == _lib.SOLVER_RANSAC
                           else ls_cfg.get("threshold", 0.02)),
        ransac_max_iterations=int(rs.get("max_iterations", 5000)),
        ransac_min_inliers_percentage=float(rs.get("min_inliers_percentage", 0.95)),
        ransac_final=ransac_final,
        drpm_threshold=float(rs.get("DRPM_threshold", 0.05)),
        drpm_stdev_points=float(rs.get("DRPM_stdev_points", 0.02)),
        drpm_stdev_normals=float(rs.get("DRPM_stdev_normals", 0.05)),
        ransac_seed=int(rs.get("seed", 1)),
    )
