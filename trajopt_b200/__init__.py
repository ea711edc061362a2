"""Importable alias of the ``trajectoryoptimization.jl_b200/`` package (its directory name is not a valid
Python identifier).  ``import trajopt_b200 as TO`` gives the mirror of the reference's API."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trajectoryoptimization.jl_b200")
_name = "trajectoryoptimization_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_name]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
capi = _mod._capi
PACKAGE_DIR = _pkg_dir
