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
}


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
    for name, build in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(build(OracleProblem)))
        print("wrote", name)
