"""Element-wise GPU-vs-oracle parity at the FULL sizes of BASELINE.json's configurations (`-m gpu`).

The per-kernel and small-batch tests live in test_gpu_parity.py; this file runs the benchmarked sizes themselves: every instance
of the batch, every knot, gains / expected decrease / step sizes / regularisation / iterates, through the C ABI, against the CPU
oracle on the same seeded inputs.  More instances than resident warps (4096 > 148 x 16 or 148 x 28) exercises the atomic work
queues of the persistent Riccati kernels, which the small batches cannot.

Tolerances (parity_util.py): one kernel application on identical inputs agrees with the oracle to kernel tolerance;
downstream of a backward pass every instance has to stay within FACTOR x what a relative perturbation of the gains by that kernel
tolerance does to it in the oracle itself (the twin) -- Cartpole / Acrobot: that is the tight tolerance; the tumbling Quadrotor
instances: whatever the closed loop's amplification leaves."""
import numpy as np
import pytest

import trajopt_b200 as TO
from parity_util import GAIN_TOL, check, decisions_agree, triple

pytestmark = pytest.mark.gpu
P = TO.problems

ROLLOUT_TOL = 1e-12  # rollout!, [A B]: no recursion, no amplification
K_TOL = GAIN_TOL     # gains / expected decrease of one backward pass, like with like (measured: 4e-11 full state, 1e-9 register-resident kernel)
F_TOL = 1e-9         # one forward pass: only where the twins say the instance is that well conditioned
ITER_TOL = 1e-8      # after the iterations: ditto

FULL = {
    # BASELINE.json configs[1]: Cartpole swing-up batch 1024, unconstrained (warp kernel is the automatic choice below 2048 instances)
    "cartpole_B1024": (lambda cls: P.cartpole(B=1024, N=101, cls=cls), {}),
    # configs[2]: Quadrotor point-to-point batch 4096, goal + control bounds -- on the Lie-group error state (what Altro does for this
    # model; k_riccati_frag, register-resident n_e = 12) ...
    "quadrotor_errstate_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls, error_state=True), {}),
    # ... and the full-state recursion (k_riccati, DMMA n = 13; round-1 headline)
    "quadrotor_fullstate_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls), {}),
    # configs[3]: Acrobot batch 8192 N=201, AL + dense second-order cost expansion -- both Riccati kernels
    "acrobot_B8192_warp": (lambda cls: P.acrobot(B=8192, N=201, cls=cls), {"backward_kernel": 1}),
    "acrobot_B8192_diag_thread": (lambda cls: P.acrobot(B=8192, N=201, cls=cls, dense_cost=False), {"backward_kernel": 2}),
    # configs[4]: the MPC sweep's longest horizon
    "quadrotor_errstate_N401_B1024": (lambda cls: P.quadrotor(B=1024, N=401, cls=cls, dt=0.05, error_state=True), {}),
    "quadrotor_fullstate_N401_B1024": (lambda cls: P.quadrotor(B=1024, N=401, cls=cls, dt=0.05), {}),
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_elementwise(name):
    build, opts = FULL[name]
    g, o, t = triple(build, opts)
    probs = (g, o, t)
    for p in probs:
        TO.rollout(p); TO.expand(p)
    check("rollout X", TO.states(g), TO.states(o), TO.states(t), ROLLOUT_TOL)
    AB = (lambda p: TO.error_dynamics(p)) if g.error_state else (lambda p: TO.dynamics_jacobians(p))
    check("[A B]", AB(g), AB(o), AB(t), ROLLOUT_TOL)
    sg, so, st_ = TO.backward(g), TO.backward(o), TO.backward(t)
    assert np.array_equal(sg, so), "regularisation restarts differ"
    (Kg, dg), (Ko, do) = TO.gains(g), TO.gains(o)
    # one kernel application on identical inputs: kernel tolerance, no allowance
    check("K", Kg, Ko, Ko, K_TOL); check("d", dg, do, do, K_TOL)
    check("dV", TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], TO.solver_state(o)["dV"], K_TOL)
    (Jg, ag), (Jo, ao), (Jt, at) = TO.forward(g), TO.forward(o), TO.forward(t)
    ok = decisions_agree("accepted step sizes", ag, ao, at, allow=0.04)     # (N = 401 on the error state: 2.5 % of the instances sit within GAIN_TOL of an acceptance threshold)
    OUT1 = 0.01     # (one noise draw of the twin is a coarse yardstick for the instances whose open-loop guess tumbles over a 20 s horizon)
    check("merit after the forward pass", Jg, Jo, Jt, F_TOL, ok, OUT1)
    check("X after the forward pass", TO.states(g), TO.states(o), TO.states(t), F_TOL, ok, OUT1)
    check("U after the forward pass", TO.controls(g), TO.controls(o), TO.controls(t), F_TOL, ok, OUT1)
    for p in probs:
        TO.ilqr_step(p, 2)
        if len(p.constraints):
            TO.al_update(p)
            TO.ilqr_step(p, 2)
    stg, sto, stt = TO.solver_state(g), TO.solver_state(o), TO.solver_state(t)
    # an instance that has converged takes its discrete decisions on the last bits of J (test_gpu_parity.py)
    live = np.abs(sto["dV"][:, 0]) > 1e-9 * np.maximum(1.0, np.abs(TO.merit(o)))
    assert live.mean() > 0.5
    # a different discrete decision anywhere along the way (one more restart, one more halving of the step) sends the twins to different
    # iterates: compare the instances whose twins ended with the same last decisions, the others are counted
    dec = live & ok & (sto["alpha"] == stt["alpha"]) & (sto["bp_status"] == stt["bp_status"]) & (stg["alpha"] == sto["alpha"]) & (stg["bp_status"] == sto["bp_status"])
    assert dec.sum() > 0.5 * live.sum(), "too few instances with a common decision history"
    OUT = 0.01
    e_x, d_x = check("X after the iterations", TO.states(g), TO.states(o), TO.states(t), ITER_TOL, dec, OUT)
    check("U after the iterations", TO.controls(g), TO.controls(o), TO.controls(t), ITER_TOL, dec, OUT)
    check("merit after the iterations", TO.merit(g), TO.merit(o), TO.merit(t), ITER_TOL, dec, OUT)
    for i in range(len(g.constraints)):
        check(f"multipliers {i}", TO.multipliers(g, i), TO.multipliers(o, i), TO.multipliers(t, i), ITER_TOL, dec, OUT)
    print(f"{name}: decidable instances {dec.mean():.3f}, worst X error {e_x:.2e} (twin divergence up to {d_x:.2e})")
    for p in probs:
        p.close()
