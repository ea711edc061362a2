"""The benchmark / parity problem set of BASELINE.json, built through the mirror API with deterministic
synthetic inputs (seed 1 like the reference's test/runtests.jl:13, NumPy PCG64).

Every builder takes ``cls`` (the Problem class to instantiate) so the tests can build the identical problem
on the CPU oracle; the default is the CUDA-backed ``Problem``.
"""
import numpy as np

from . import api as TO


def _rng(seed):
    return np.random.default_rng(seed)


def double_integrator(B=1, N=21, dim=2, seed=1, cls=None, constrained=True, **kw):
    """examples/quickstart.jl:28-64 (dim=2: n=4, m=2, N=21, tf=3) or the 1-D variant of BASELINE.json configs[0]
    (dim=1: n=2, m=1, N=51).  Goal + control bounds (the solver kernels take Goal/Bound constraints)."""
    cls = cls or TO.Problem
    model = TO.DoubleIntegrator(dim)
    n, m = model.dims()
    tf = 3.0
    xf = np.zeros(n); xf[dim - 1] = 2.0
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n) * (N - 1), xf, N)
    cons = TO.ConstraintList(n, m, N)
    if constrained:
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
        TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-10, u_max=10), (1, N - 1))
    r = _rng(seed)
    prob = cls(model, obj, np.zeros(n), tf, xf=xf, constraints=cons, batch=B, **kw)
    TO.initial_controls(prob, r.standard_normal((B, N - 1, m)))
    return prob


def cartpole(B=1024, N=101, seed=1, cls=None, u_bound=None, goal=False, dt_scaled_cost=False, **kw):
    """Cartpole swing-up, examples/Cartpole.ipynb:82-83,123-125,187: Q=1e-2 I, R=1e-1 I, Qf=100 I, tf=5,
    x0=0, xf=(0,pi,0,0), U0=0.01; batch: x0_b += N(0,0.1^2) on (y,theta), U0_b += N(0,0.01^2).
    Unconstrained by default (BASELINE configs[1]); u_bound=3 + goal=True gives the notebook's ALTRO problem.
    ``dt_scaled_cost`` multiplies Q and R by dt: the notebook's outputs were saved with TrajectoryOptimization v0.3,
    whose stage costs were integrated with dt; v0.7.1 (the reference) sums them unscaled (src/objective.jl:104-106)."""
    cls = cls or TO.Problem
    model = TO.Cartpole()
    n, m = 4, 1
    xf = np.array([0, np.pi, 0, 0])
    sc = 5.0 / (N - 1) if dt_scaled_cost else 1.0
    obj = TO.LQRObjective(1e-2 * sc * np.eye(n), 1e-1 * sc * np.eye(m), 100.0 * np.eye(n), xf, N)
    cons = TO.ConstraintList(n, m, N)
    if u_bound is not None:
        TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-u_bound, u_max=u_bound), (1, N - 1))
    if goal:
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    r = _rng(seed)
    x0 = np.zeros((B, n))
    U0 = np.full((B, N - 1, m), 0.01)
    if B > 1:
        x0[:, :2] += 0.1 * r.standard_normal((B, 2))
        U0 += 0.01 * r.standard_normal((B, N - 1, m))
    prob = cls(model, obj, x0, 5.0, xf=xf, constraints=cons, **kw)
    TO.initial_controls(prob, U0)
    return prob


def quadrotor(B=4096, N=101, seed=1, cls=None, constrained=True, dt=None, dense_cost=False, u_noise=0.05, **kw):
    """Quadrotor point-to-point, test/internal_api.jl:20-34: Q=.1 I, R=.01 I, Qf=100 I, x0=[1,2,1;1,0,0,0;0;0],
    xf=[0,0,2;1,0,0,0;0;0], u in [0,10] at 1..N-1, Goal(xf) at N, tf=5; U0 = hover + N(0,0.05^2);
    batch: r0_b = r0 + U(-1,1)^3.  ``dt`` fixes the step (MPC sweep: dt = .05 for every N)."""
    cls = cls or TO.Problem
    model = TO.Quadrotor()
    n, m = 13, 4
    x0 = np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    if dense_cost:   # QuadraticCost with off-diagonal Q and an x-u cross term: exercises the dense second-order expansion
        Q = 0.1 * np.eye(n) + 0.002 * (np.ones((n, n)) - np.eye(n)); R = 0.01 * np.eye(m) + 0.001 * (np.ones((m, m)) - np.eye(m))
        H = 0.001 * np.ones((m, n)); uf = np.full(m, 1.2)
        stage = TO.QuadraticCost(Q, R, H=H, q=-Q @ xf - H.T @ uf, r=-R @ uf - H @ xf, c=0.5 * xf @ Q @ xf + 0.5 * uf @ R @ uf + uf @ H @ xf)
        term = TO.QuadraticCost(100.0 * np.eye(n), R, q=-100.0 * xf, c=50.0 * xf @ xf, terminal=True)
        obj = TO.Objective(stage, term, N)
    else:
        obj = TO.LQRObjective(np.full(n, 0.1), np.full(m, 0.01), np.full(n, 100.0), xf, N)
    cons = TO.ConstraintList(n, m, N)
    if constrained:
        TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=np.zeros(4), u_max=np.full(4, 10.0)), (1, N - 1))
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    r = _rng(seed)
    x0b = np.broadcast_to(x0, (B, n)).copy()
    if B > 1:
        x0b[:, :3] += r.uniform(-1, 1, (B, 3))
    U0 = model.hover_control()[None, None, :] + u_noise * r.standard_normal((B, N - 1, m))
    tf = 5.0 if dt is None else dt * (N - 1)
    prob = cls(model, obj, x0b, tf, xf=xf, constraints=cons, **kw)
    TO.initial_controls(prob, U0)
    return prob


def acrobot(B=8192, N=201, seed=1, cls=None, dense_cost=True, **kw):
    """Acrobot swing-up (no in-reference definition; RobotZoo defaults): x0=(-pi/2,0,0,0) -> xf=(pi/2,0,0,0), tf=10,
    Q=I, R=.01, Qf=100 I, |u|<=15 + goal, AL on.  ``dense_cost`` uses a QuadraticCost with a small x-u cross term so
    the full second-order cost expansion (BASELINE configs[3]) is exercised."""
    cls = cls or TO.Problem
    model = TO.Acrobot()
    n, m = 4, 1
    x0 = np.array([-np.pi / 2, 0, 0, 0]); xf = np.array([np.pi / 2, 0, 0, 0])
    Q, R, Qf = np.eye(n), 0.01 * np.eye(m), 100.0 * np.eye(n)
    if dense_cost:
        Q = Q + 0.05 * (np.ones((n, n)) - np.eye(n))
        H = 0.01 * np.ones((m, n))
        stage = TO.QuadraticCost(Q, R, H=H, q=-Q @ xf, r=-H @ xf, c=0.5 * xf @ Q @ xf)
        term = TO.QuadraticCost(Qf, R, q=-Qf @ xf, c=0.5 * xf @ Qf @ xf, terminal=True)
        obj = TO.Objective(stage, term, N)
    else:
        obj = TO.LQRObjective(Q, R, Qf, xf, N)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-15.0, u_max=15.0), (1, N - 1))
    TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    r = _rng(seed)
    x0b = np.broadcast_to(x0, (B, n)).copy()
    U0 = np.zeros((B, N - 1, m))
    if B > 1:
        x0b[:, :2] += 0.05 * r.standard_normal((B, 2))
        U0 += 0.1 * r.standard_normal((B, N - 1, m))
    prob = cls(model, obj, x0b, 10.0, xf=xf, constraints=cons, **kw)
    TO.initial_controls(prob, U0)
    return prob


def quadrotor_lie(B=64, N=51, seed=1, cls=None, quat_cost=True, quat_goal=True, error_state=True, **kw):
    """Quadrotor with the Lie-group handling of SURVEY 8(f2): Riccati on the 12-dimensional error state (``error_state``), a
    ``QuatLQRCost`` objective (src/lie_costs.jl:129-139, zero quadratic weight on the quaternion, geodesic weight w) and, instead of
    a full-state goal, ``GoalConstraint`` on position / velocities + ``QuatVecEq`` on the attitude (src/constraints.jl:938-965).
    Inputs as in ``quadrotor`` plus a random initial attitude per instance."""
    cls = cls or TO.Problem
    model = TO.Quadrotor()
    n, m = 13, 4
    r = _rng(seed)
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    Qd = np.array([0.1] * 3 + [0.0] * 4 + [0.1] * 6); Qfd = np.array([100.0] * 3 + [0.0] * 4 + [100.0] * 6)
    Rd = np.full(m, 0.01)
    uf = model.hover_control()
    if quat_cost:
        stage = TO.QuatLQRCost(Qd, Rd, xf, uf, w=0.5)
        term = TO.QuatLQRCost(Qfd, Rd, xf, uf, w=20.0, terminal=True)
    else:   # plain LQR cost with isotropic quaternion weights (the BASELINE objective)
        stage = TO.LQRCost(np.full(n, 0.1), Rd, xf, uf)
        term = TO.LQRCost(np.full(n, 100.0), Rd, xf, uf, terminal=True)
    obj = TO.Objective(stage, term, N)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=np.zeros(4), u_max=np.full(4, 10.0)), (1, N - 1))
    if quat_goal:
        TO.add_constraint(cons, TO.GoalConstraint(xf, inds=[1, 2, 3, 8, 9, 10, 11, 12, 13]), N)
        TO.add_constraint(cons, TO.QuatVecEq(n, m, xf[3:7]), N)
    else:
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
    x0 = np.zeros((B, n))
    x0[:, :3] = np.array([1, 2, 1.0]) + (r.uniform(-1, 1, (B, 3)) if B > 1 else 0.0)
    q0 = np.array([1.0, 0, 0, 0]) + 0.3 * r.standard_normal((B, 4))
    x0[:, 3:7] = q0 / np.linalg.norm(q0, axis=1, keepdims=True)
    U0 = uf[None, None, :] + 0.05 * r.standard_normal((B, N - 1, m))
    prob = cls(model, obj, x0, 0.05 * (N - 1), xf=xf, constraints=cons, error_state=error_state, **kw)
    TO.initial_controls(prob, U0)
    return prob


def quadrotor_zigzag(cls=None, dt_scaled_cost=True, error_state=True, **kw):
    """examples/Quadrotor.ipynb cells 10-20: 20 m flight through two waypoints, N=101, tf=5, u in [0,12], initial control
    0.5*mass/m.  ``dt_scaled_cost``: the notebook's recorded output (cost 0.2992834848449584, :374-376) was produced with
    TrajectoryOptimization v0.3, which integrated the stage costs with dt (see ``cartpole``)."""
    cls = cls or TO.Problem
    model = TO.Quadrotor()
    n, m, N = 13, 4, 101
    sc = 5.0 / (N - 1) if dt_scaled_cost else 1.0
    fill = lambda p, q, v, w: np.array([p] * 3 + [q] * 4 + [v] * 3 + [w] * 3, dtype=float)      # RobotDynamics.fill_state
    build = lambda r: np.array([*r, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=float)                   # RobotDynamics.build_state, identity attitude
    x0, xf = build([0, -10, 1.0]), build([0, 10, 1.0])
    R = np.full(m, 1e-4) * sc
    cost_nom = TO.LQRCost(fill(1e-5, 1e-5, 1e-3, 1e-3) * sc, R, build([0, 0, 0.0]))
    wpts, times = [[10, 0, 1.0], [-10, 0, 1.0], [0, 10, 1.0]], [33, 66, 101]
    Qw, Qf = fill(1e3, 1, 1, 1), fill(10.0, 100, 10, 10)
    costs = [TO.LQRCost(Qf if t == N else 1e-3 * Qw * sc, R, build(r), terminal=(t == N)) for r, t in zip(wpts, times)]
    obj = TO.Objective([costs[times.index(k)] if k in times else cost_nom for k in range(1, N + 1)])
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=0.0, u_max=12.0), (1, N - 1))
    prob = cls(model, obj, x0, 5.0, xf=xf, constraints=cons, error_state=error_state, **kw)
    TO.initial_controls(prob, np.full((1, N - 1, m), 0.5 * model.mass / m))
    return prob


CONFIGS = {
    "quadrotor_lie": quadrotor_lie,
    "quadrotor_zigzag": quadrotor_zigzag,
    "double_integrator": double_integrator,
    "cartpole": cartpole,
    "quadrotor": quadrotor,
    "acrobot": acrobot,
}
