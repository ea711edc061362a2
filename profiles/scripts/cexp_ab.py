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
out = {}
for (B, N) in ((37, 101), (5, 33), (64, 16)):
    p = TO.problems.quadrotor(B=B, N=N, error_state=True, u_noise=0.01)
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
