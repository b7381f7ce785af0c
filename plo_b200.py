"""Import shim: the package directory is named after the upstream repo
(`planetary-lidar-odometry_b200/`, hyphenated, hence not importable by name).
`import plo_b200` loads it as `planetary_lidar_odometry_b200` and aliases it."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_NAME = "planetary_lidar_odometry_b200"
_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "planetary-lidar-odometry_b200")

if _NAME not in _sys.modules:
    _spec = _ilu.spec_from_file_location(_NAME, _os.path.join(_DIR, "__init__.py"),
                                         submodule_search_locations=[_DIR])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_sys.modules[__name__] = _sys.modules[_NAME]
