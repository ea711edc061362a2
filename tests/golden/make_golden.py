#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the CPU oracle (oracle/), the pinned restatement of the
reference path.  (The Julia reference itself cannot run in this image -- no Julia -- so fixtures cannot be generated
from it; the oracle is pinned against the reference's own test KATs in tests/test_oracle_kats.py.)

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import trajopt_b200 as TO  # noqa: E402
from oracle_binding import OracleProblem  # noqa: E402

CASES = {
    "cartpole_B2_N21": lambda cls: TO.problems.cartpole(B=2, N=21, cls=cls, u_bound=3.0, goal=True),
    "quadrotor_B2_N11": lambda cls: TO.problems.quadrotor(B=2, N=11, cls=cls, dt=0.05),
    "acrobot_dense_B2_N21": lambda cls: TO.problems.acrobot(B=2, N=21, cls=cls, dense_cost=True),
    "double_integrator_B1_N21": lambda cls: TO.problems.double_integrator(B=1, N=21, dim=2, cls=cls),
    # Lie-group error state with QuatLQRCost + QuatVecEq, and the BASELINE objective on the error state (compact expansion)
    "quadrotor_lie_B2_N11": lambda cls: TO.problems.quadrotor_lie(B=2, N=11, cls=cls),
    "quadrotor_lie_lqr_B2_N11": lambda cls: TO.problems.quadrotor_lie(B=2, N=11, cls=cls, quat_cost=False, quat_goal=False),
    # the docs' nonlinear CartpoleCost as a recorded program + a recorded control-norm constraint
    "cartpole_autodiff_B2_N21": lambda cls: _cartpole_autodiff(cls),
}


def _cartpole_autodiff(cls):
    n, m, N = 4, 1, 21

    def stage(x, u):   # docs/src/costfunction_interface.md:38-50 with Q = (0.1, 5, 0.1, 0.1), R = 0.05
        return 5.0 * TO.cos(x[1] / 2) + 0.5 * (0.1 * x[0] ** 2 + 0.1 * x[2] ** 2 + 0.1 * x[3] ** 2) + 0.5 * 0.05 * u[0] ** 2
    xf = np.array([0, np.pi, 0, 0])
    obj = TO.Objective(TO.AutodiffCost(n, m, stage), TO.LQRCost(np.full(n, 100.0), np.full(m, 0.05), xf, terminal=True), N)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda u: [TO.sqrt(u[0] * u[0] + 1e-12) - 4.0], TO.Inequality(), "control"), (1, N - 1))
    TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    x0 = np.array([[0.0, 0.1, 0, 0], [0.1, -0.2, 0.05, 0]])
    prob = cls(TO.Cartpole(), obj, x0, 1.0, constraints=cons)
    TO.initial_controls(prob, np.linspace(0.5, 1.5, 2 * (N - 1)).reshape(2, N - 1, 1))
    return prob


def run_case(prob):
    out = {"x0": prob.x0.copy(), "U0": TO.controls(prob)}
    TO.rollout(prob)
    out["X_rollout"] = TO.states(prob)
    out["J_knots"] = TO.cost_knots(prob)
    out["merit0"] = TO.merit(prob)
    for i in range(len(prob.constraints)):
        out[f"c{i}"] = TO.evaluate_constraints(prob, i)
    TO.expand(prob)
    out["AB"] = TO.dynamics_jacobians(prob)
    out["bp_status"] = TO.backward(prob)
    out["K"], out["d"] = TO.gains(prob)
    out["dV"] = TO.solver_state(prob)["dV"]
    J, alpha = TO.forward(prob)
    out["J1"], out["alpha1"] = J, alpha
    out["X1"], out["U1"] = TO.states(prob), TO.controls(prob)
    TO.ilqr_step(prob, 2)
    out["merit3"] = TO.merit(prob)
    TO.al_update(prob)
    for i in range(len(prob.constraints)):
        out[f"lambda{i}"] = TO.multipliers(prob, i)
    return out


if __name__ == "__main__":
    only = sys.argv[1:]          # optional: names of the cases to (re)write
    for name, build in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(build(OracleProblem)))
        print("wrote", name)
