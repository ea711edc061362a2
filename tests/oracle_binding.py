"""Test-side binding of the CPU oracle (oracle/_build/liboracle.so).

``OracleProblem`` subclasses the product's ``Problem`` description class but opens the ORACLE library instead of
the CUDA one, so every API function of ``trajopt_b200`` (``rollout(prob)``, ``cost(prob)``, ...) can be run on
either side with identical inputs.  This file lives in tests/: the product never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import trajopt_b200 as TO

K = TO.capi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_lib = None


def build_oracle():
    if not os.path.exists(ORACLE_LIB) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_LIB) for f in ("oracle.hpp", "models.hpp", "oracle_capi.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_LIB


def load_oracle():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_last_error.argtypes = [C.c_void_p]
    return _lib


class OracleProblem(TO.Problem):
    def _open(self):
        self._lib = load_oracle()
        self._h = C.c_void_p()
        rc = self._lib.orc_create(C.byref(self.spec.c), C.byref(self._h))
        if rc:
            msg = self._lib.orc_last_error(None).decode()
            raise {K.TO_EDIM: TO.DimensionMismatch, K.TO_EINVAL: TO.ArgumentError}.get(rc, TO.TrajOptError)(msg)

    def _raw_call(self, name, *args):
        if name == "to_backward_algebra":
            args[0]._obj.value = self._lib.orc_get_backward_variant(self._h)
            return
        fn = getattr(self._lib, "orc_" + name[3:])
        fn.restype = C.c_int
        conv = [C.c_double(a) if isinstance(a, float) else a for a in args]
        rc = fn(self._h, *conv)
        if rc:
            raise TO.TrajOptError(self._lib.orc_last_error(self._h).decode())

    def _default_options(self, o):
        self._lib.orc_default_options(C.byref(o))

    def set_backward_variant(self, v):
        """arithmetic form of the backward pass (oracle.hpp Options::backward_variant): 0 = Cholesky solve, 1 = the block-inverse
        algebra of csrc/riccati_frag.cu"""
        assert self._lib.orc_set_backward_variant(self._h, int(v)) == 0
        return self

    def set_gain_noise(self, rel):
        """test instrument (oracle.hpp Options::gain_noise): every backward pass returns gains perturbed by the relative amount `rel`"""
        self._lib.orc_set_gain_noise.argtypes = [C.c_void_p, C.c_double]
        assert self._lib.orc_set_gain_noise(self._h, float(rel)) == 0
        return self

    def set_integrator(self, order):
        assert self._lib.orc_set_integrator(self._h, int(order)) == 0
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orc_destroy(self._h)
            self._h = C.c_void_p()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def match_algebra(g, o):
    """make the oracle problem `o` use the arithmetic form of the backward pass that the CUDA problem `g` will use, so that the
    comparison is like with like (tests/test_oracle_variants.py measures what the OTHER form differs by)"""
    o.set_backward_variant(TO.backward_algebra(g))
    return o


def oracle_projection(cone, x):
    lib = load_oracle()
    x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
    out = np.empty_like(x)
    rc = lib.orc_projection(cone.code, x.shape[1], x.shape[0], _dp(x), _dp(out))
    return out, rc


def oracle_grad_projection(cone, x):
    lib = load_oracle()
    x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
    out = np.empty((x.shape[0], x.shape[1], x.shape[1]))
    rc = lib.orc_grad_projection(cone.code, x.shape[1], x.shape[0], _dp(x), _dp(out))
    return np.swapaxes(out, -1, -2), rc


def oracle_hess_projection(cone, x, b):
    lib = load_oracle()
    x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
    b = np.ascontiguousarray(np.atleast_2d(np.asarray(b, dtype=np.float64)))
    out = np.empty((x.shape[0], x.shape[1], x.shape[1]))
    rc = lib.orc_hess_projection(cone.code, x.shape[1], x.shape[0], _dp(x), _dp(b), _dp(out))
    return np.swapaxes(out, -1, -2), rc


def _model_args(model):
    p = np.ascontiguousarray(np.asarray(model.params, dtype=np.float64))
    dim = model.m if isinstance(model, TO.DoubleIntegrator) else 1
    return model.model_id, dim, _dp(p), len(p), p


def oracle_dynamics(model, x, u):
    lib = load_oracle()
    mid, dim, pp, npar, keep = _model_args(model)
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    xd = np.empty(model.n)
    lib.orc_dynamics(mid, dim, pp, npar, _dp(x), _dp(u), _dp(xd))
    return xd


def oracle_discrete_dynamics(model, x, u, h):
    lib = load_oracle()
    mid, dim, pp, npar, keep = _model_args(model)
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    xn = np.empty(model.n)
    lib.orc_discrete_dynamics(mid, dim, pp, npar, _dp(x), _dp(u), C.c_double(h), _dp(xn))
    return xn


def oracle_discrete_jacobian(model, x, u, h):
    lib = load_oracle()
    mid, dim, pp, npar, keep = _model_args(model)
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    AB = np.empty((model.n + model.m, model.n))
    lib.orc_discrete_jacobian(mid, dim, pp, npar, _dp(x), _dp(u), C.c_double(h), _dp(AB))
    return AB.T.copy()   # n x (n+m)
