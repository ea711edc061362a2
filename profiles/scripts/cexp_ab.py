"""A/B of the record cost-expansion kernels: run a few iLQR iterations of the error-state Quadrotor and dump gains, trajectory and merit.
usage: cexp_ab.py out.npz   (TO_CEXP_V1=1 selects k_expansion_rec16, default k_expansion_rec16b); compare two dumps with --cmp a.npz b.npz"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    worst = 0.0
    for k in a.files:
        d = float(np.max(np.abs(a[k] - b[k]))) if a[k].size else 0.0
        worst = max(worst, d)
        print(f"{k:12s} max|a-b| = {d:.3e}  identical={np.array_equal(a[k], b[k])}")
    print("A/B", "IDENTICAL" if worst == 0.0 else f"DIFFER (max {worst:.3e})")
    sys.exit(0)
import trajopt_b200 as TO


def bounded(B, N):
    """error-state Quadrotor with Bound rows on states AND controls (three AL terms on some entries at the last stage knots) + goal"""
    model = TO.Quadrotor(); n, m = 13, 4
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    obj = TO.LQRObjective(np.full(n, 0.1), np.full(m, 0.01), np.full(n, 100.0), xf, N)
    cons = TO.ConstraintList(n, m, N)
    x_max = np.full(n, np.inf); x_min = np.full(n, -np.inf)
    x_max[:3] = 2.5; x_min[:3] = -0.5; x_max[7:10] = 1.0; x_min[7:10] = -1.0; x_max[12] = 0.3
    TO.add_constraint(cons, TO.BoundConstraint(n, m, x_min=x_min, x_max=x_max, u_min=np.zeros(4), u_max=np.full(4, 10.0)), (1, N - 1))
    TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    r = np.random.default_rng(5)
    x0 = np.broadcast_to(np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), (B, n)).copy(); x0[:, :3] += r.uniform(-1, 1, (B, 3))
    prob = TO.Problem(model, obj, x0, 0.05 * (N - 1), xf=xf, constraints=cons, error_state=True)
    TO.initial_controls(prob, model.hover_control()[None, None, :] + 0.01 * r.standard_normal((B, N - 1, m)))
    return prob


out = {}
for (B, N) in ((37, 101), (5, 33), (64, 16), (-6, 40)):
    p = bounded(-B, N) if B < 0 else TO.problems.quadrotor(B=B, N=N, error_state=True, u_noise=0.01)
    TO.rollout(p)
    TO.ilqr_step(p, 3)          # iteration 1: full expansion; 2, 3: the overlapped (mode 1 / 2) launches
    K, d = TO.gains(p)
    out[f"K_{B}_{N}"] = K; out[f"d_{B}_{N}"] = d; out[f"X_{B}_{N}"] = TO.states(p); out[f"J_{B}_{N}"] = TO.merit(p)
    TO.al_update(p)
    TO.ilqr_step(p, 2)
    out[f"X2_{B}_{N}"] = TO.states(p); out[f"J2_{B}_{N}"] = TO.merit(p)
    p.close()
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1])
