"""trajectoryoptimization.jl_b200 -- B200-native (sm_100a) batched problem-evaluation hot path of
TrajectoryOptimization.jl behind the reference's Problem / Objective / AbstractConstraint surface.

The directory name contains a dot, so it is imported through the ``trajopt_b200`` alias package at the repo
root (``import trajopt_b200 as TO``).  Layout:
    csrc/          hand-written CUDA kernels + the C ABI (include/trajopt_b200.h)  -> libtrajopt_b200.so
    _capi.py       ctypes binding of the C ABI
    api.py         host-side mirror of the reference's Julia API
    julia/         the ccall shim a Julia host uses (cannot run in this image: no Julia)
"""
from .api import *  # noqa: F401,F403
from . import _capi  # noqa: F401
from . import problems  # noqa: F401,E402
from . import multi_gpu  # noqa: F401,E402
