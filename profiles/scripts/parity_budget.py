#!/usr/bin/env python
"""Measured GPU-vs-oracle error budget at the full BASELINE sizes.

For every configuration: the error of ONE kernel application on identical inputs (rollout, [A B], one backward pass: K, d, dV; one
forward pass: X, U, J) and the growth of the iterate error over consecutive closed-loop iterations (each iteration starts from the
previous one's slightly different trajectory).  Errors are max |gpu - oracle| / max(1, max |oracle|) over the whole batch, the
metric the tests use.  Writes gpurun_out/r02_parity_budget.json and a markdown table on stdout."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as TO
from oracle_binding import OracleProblem

P = TO.problems
CONFIGS = {
    "cartpole_B1024": (lambda cls: P.cartpole(B=1024, N=101, cls=cls), {}),
    "quadrotor_fullstate_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls), {}),
    "quadrotor_errstate_B4096": (lambda cls: P.quadrotor(B=4096, N=101, cls=cls, error_state=True), {}),
    "quadrotor_errstate_smem_kernel_B1024": (lambda cls: P.quadrotor(B=1024, N=101, cls=cls, error_state=True), {"backward_kernel": 5}),
    "acrobot_dense_B8192": (lambda cls: P.acrobot(B=8192, N=201, cls=cls), {}),
    "quadrotor_errstate_N401_B1024": (lambda cls: P.quadrotor(B=1024, N=401, cls=cls, dt=0.05, error_state=True), {}),
}


def err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def main():
    out = {}
    names = sys.argv[1:] or list(CONFIGS)
    for name in names:
        build, opts = CONFIGS[name]
        g, o = build(TO.Problem), build(OracleProblem)
        if opts:
            TO.set_options(g, **opts)
        r = {}
        for p in (g, o):
            TO.rollout(p); TO.expand(p)
        r["rollout_X"] = err(TO.states(g), TO.states(o))
        r["AB"] = err(TO.error_dynamics(g), TO.error_dynamics(o)) if g.error_state else err(TO.dynamics_jacobians(g), TO.dynamics_jacobians(o))
        sg, so = TO.backward(g), TO.backward(o)
        r["restarts_equal"] = bool(np.array_equal(sg, so))
        Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
        r["K"], r["d"] = err(Kg, Ko), err(dg, do)
        r["K_scale"], r["d_scale"] = float(np.max(np.abs(Ko))), float(np.max(np.abs(do)))
        r["dV"] = err(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"])
        Jg, ag = TO.forward(g); Jo, ao = TO.forward(o)
        r["alpha_equal"] = bool(np.array_equal(ag, ao))
        r["fwd_J"], r["fwd_X"], r["fwd_U"] = err(Jg, Jo), err(TO.states(g), TO.states(o)), err(TO.controls(g), TO.controls(o))
        growth = []
        for it in range(2, 7):
            for p in (g, o):
                TO.ilqr_step(p, 1)
                if it == 4 and len(p.constraints):
                    TO.al_update(p)
            sto, stg = TO.solver_state(o), TO.solver_state(g)
            live = np.abs(sto["dV"][:, 0]) > 1e-9 * np.maximum(1.0, np.abs(TO.merit(o)))
            growth.append({"iteration": it, "J": err(TO.merit(g)[live], TO.merit(o)[live]), "X": err(TO.states(g)[live], TO.states(o)[live]),
                           "U": err(TO.controls(g)[live], TO.controls(o)[live]), "alpha_mismatch": int(np.sum(stg["alpha"][live] != sto["alpha"][live])),
                           "live": int(live.sum())})
        r["iterations"] = growth
        out[name] = r
        g.close(); o.close()
        print(name, json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity_budget.json"), "w"), indent=1)
    print("\n| config | rollout X | [A B] | K | d | dV | fwd J | fwd X | fwd U | after 6 it: J | X | U |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n, r in out.items():
        l = r["iterations"][-1]
        print(f"| {n} | {r['rollout_X']:.1e} | {r['AB']:.1e} | {r['K']:.1e} | {r['d']:.1e} | {r['dV']:.1e} | {r['fwd_J']:.1e} | {r['fwd_X']:.1e} | {r['fwd_U']:.1e} | {l['J']:.1e} | {l['X']:.1e} | {l['U']:.1e} |")


if __name__ == "__main__":
    main()
