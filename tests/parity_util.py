"""Conditioning-aware GPU-vs-oracle comparison (test infrastructure).

Why: the Riccati recursion and the closed-loop rollouts of the BASELINE Quadrotor problem amplify rounding differences by several
orders of magnitude per stage (cheap controls R = 0.01 against Qf = 100: the cost-to-go update is a Schur complement with heavy
cancellation; the open-loop initial guess tumbles).  Two CPU evaluations of the SAME mathematics in fp64 -- the oracle's two
arithmetic forms of the backward pass, oracle.hpp Options::backward_variant -- already differ by 1e-7 in the gains and 1e-3 in
the trajectory after one forward pass on that problem, while on Cartpole / Acrobot everything agrees to 1e-11 after six iterations
(profiles/r02_notes.md).  So every comparison is made twice:

  * like with like: the oracle runs the arithmetic form the CUDA kernel uses (oracle_binding.match_algebra) -- this is the parity
    statement, at kernel tolerance;
  * against the TWIN (the oracle in the other form): the per-instance divergence D_b of the two oracle runs is the intrinsic fp64
    uncertainty of instance b at that stage; after closed-loop iterations the CUDA result has to stay inside FACTOR x D_b (or the
    tight tolerance, whichever is larger).  Discrete decisions (step size, restarts) are compared where the twins agree on them.
"""
import numpy as np

import trajopt_b200 as TO
from oracle_binding import OracleProblem, match_algebra

FACTOR = 20.0


def inst_err(a, b):
    """per-instance max |a - b| / max(1, max |b|)   (arrays with the batch on axis 0)"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    B = b.shape[0]
    d = np.abs(a - b).reshape(B, -1).max(axis=1) if b.size else np.zeros(B)
    s = np.maximum(1.0, np.abs(b).reshape(B, -1).max(axis=1)) if b.size else np.ones(B)
    d = np.where(np.isfinite(d), d, np.inf)
    return d / s


def triple(build, opts=None):
    """(cuda problem, oracle in the same arithmetic form, oracle twin in the other form)"""
    g = build(TO.Problem)
    if opts:
        TO.set_options(g, **opts)
    o = match_algebra(g, build(OracleProblem))
    t = build(OracleProblem).set_backward_variant(1 - TO.backward_algebra(g))
    return g, o, t


def check(what, a_gpu, a_orc, a_twin, tight, sel=None):
    """every instance: err(gpu, oracle) <= max(tight, FACTOR * err(twin, oracle)); returns (worst gpu error, worst twin error)"""
    e, d = inst_err(a_gpu, a_orc), inst_err(a_twin, a_orc)
    if sel is not None:
        e, d = e[sel], d[sel]
    tol = np.maximum(tight, FACTOR * d)
    bad = np.nonzero(~(e <= tol))[0]
    assert bad.size == 0, (f"{what}: {bad.size} of {e.size} instances outside the budget; worst gpu-vs-oracle {e[bad].max():.3e} "
                           f"with twin divergence {d[bad][np.argmax(e[bad])]:.3e} (tight tolerance {tight:.0e})")
    return float(e.max()) if e.size else 0.0, float(d.max()) if d.size else 0.0


def decisions_agree(what, v_gpu, v_orc, v_twin, sel=None):
    """discrete per-instance results are compared where the two oracle forms agree on them"""
    v_gpu, v_orc, v_twin = np.asarray(v_gpu), np.asarray(v_orc), np.asarray(v_twin)
    m = (v_orc == v_twin)
    if sel is not None:
        m &= sel
    assert np.array_equal(v_gpu[m], v_orc[m]), f"{what}: {int(np.sum(v_gpu[m] != v_orc[m]))} mismatches among {int(m.sum())} decidable instances"
    return float(m.mean())
