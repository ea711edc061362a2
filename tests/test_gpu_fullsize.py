"""Element-wise GPU-vs-oracle parity at the FULL sizes of BASELINE.json's configurations (`-m gpu`).

The per-kernel and small-batch tests live in test_gpu_parity.py; this file runs the benchmarked sizes themselves: every instance
of the batch, every knot, gains / expected decrease / step sizes / regularisation / iterates, through the C ABI, against the CPU
oracle on the same seeded inputs.  More instances than resident warps (4096 > 148 x 16 or 148 x 28) exercises the atomic work
queues of the persistent Riccati kernels, which the small batches cannot.

Tolerances are the measured error budget of profiles/parity_budget.py: one backward pass agrees to
K_RTOL (relative to the largest gain of the batch); the closed-loop rollout that follows multiplies a gain perturbation by the
sensitivity of the trajectory to the gains, which is what ITER_RTOL covers after several iterations."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem

pytestmark = pytest.mark.gpu
P = TO.problems

K_RTOL = 1e-9       # gains / expected decrease of one backward pass
F_RTOL = 1e-9       # trajectory and merit after the first forward pass
ITER_RTOL = 1e-6    # merit / iterates after 3 iterations + AL update + 2 iterations


def close(a, b, rtol, what):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.max(np.abs(b))))
    err = float(np.max(np.abs(a - b)))
    assert np.all(np.isfinite(a)), f"{what}: non-finite GPU result"
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} > {rtol:.0e} * {scale:.3e}"


FULL = {
    # BASELINE.json configs[1]: Cartpole swing-up batch 1024, unconstrained (warp kernel is the automatic choice below 2048 instances)
    "cartpole_B1024": (lambda cls: P.cartpole(B=1024, N=101, cls=cls), {}),
    # configs[2]: Quadrotor point-to-point batch 4096, goal + control bounds -- full-state recursion (k_riccati, DMMA n = 13)
    "quadrotor_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls), {}),
    # ... and on the Lie-group error state, what Altro does for this model (k_riccati_frag, register-resident n_e = 12)
    "quadrotor_errstate_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls, error_state=True), {}),
    # configs[3]: Acrobot batch 8192 N=201, AL + dense second-order cost expansion -- both Riccati kernels
    "acrobot_B8192_warp": (lambda cls: P.acrobot(B=8192, N=201, cls=cls), {"backward_kernel": 1}),
    "acrobot_B8192_diag_thread": (lambda cls: P.acrobot(B=8192, N=201, cls=cls, dense_cost=False), {"backward_kernel": 2}),
    # configs[4]: the MPC sweep's longest horizon
    "quadrotor_N401_B1024": (lambda cls: P.quadrotor(B=1024, N=401, cls=cls, dt=0.05), {}),
    "quadrotor_errstate_N401_B1024": (lambda cls: P.quadrotor(B=1024, N=401, cls=cls, dt=0.05, error_state=True), {}),
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_elementwise(name):
    build, opts = FULL[name]
    g, o = build(TO.Problem), build(OracleProblem)
    if opts:
        TO.set_options(g, **opts)
    for p in (g, o):
        TO.rollout(p); TO.expand(p)
    close(TO.states(g), TO.states(o), 1e-10, "rollout X")
    sg, so = TO.backward(g), TO.backward(o)
    assert np.array_equal(sg, so), "regularisation restarts differ"
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, K_RTOL, "K"); close(dg, do, K_RTOL, "d")
    close(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], K_RTOL, "dV")
    Jg, ag = TO.forward(g); Jo, ao = TO.forward(o)
    assert np.array_equal(ag, ao), "accepted step sizes differ"
    close(Jg, Jo, F_RTOL, "merit after the forward pass")
    close(TO.states(g), TO.states(o), F_RTOL, "X after the forward pass")
    close(TO.controls(g), TO.controls(o), F_RTOL, "U after the forward pass")
    for p in (g, o):
        TO.ilqr_step(p, 2)
        if len(p.constraints):
            TO.al_update(p)
            TO.ilqr_step(p, 2)
    stg, sto = TO.solver_state(g), TO.solver_state(o)
    # an instance that has converged takes its discrete decisions on the last bits of J (test_gpu_parity.py)
    live = np.abs(sto["dV"][:, 0]) > 1e-9 * np.maximum(1.0, np.abs(TO.merit(o)))
    assert live.mean() > 0.5
    for key in ("alpha", "ls_iters", "bp_status"):
        assert np.array_equal(stg[key][live], sto[key][live]), key
    close(stg["rho"][live], sto["rho"][live], 1e-12, "rho")
    close(TO.merit(g)[live], TO.merit(o)[live], ITER_RTOL, "merit after the iterations")
    close(TO.states(g)[live], TO.states(o)[live], ITER_RTOL, "X after the iterations")
    close(TO.controls(g)[live], TO.controls(o)[live], ITER_RTOL, "U after the iterations")
    for i in range(len(g.constraints)):
        close(TO.multipliers(g, i)[live], TO.multipliers(o, i)[live], ITER_RTOL, f"multipliers {i}")
    g.close(); o.close()
