"""CPU tests (no GPU): pin the oracle against every closed form / literal KAT the reference's own tests hold
for the hot path, plus independent checks for the parts whose arithmetic lives in un-vendored
Julia packages (RK4, dynamics Jacobians, Riccati, line search).  Each test cites the reference test it
re-expresses.  Tolerance: the reference's `≈` (rtol sqrt(eps) ~ 1.5e-8) unless stated.
"""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import (OracleProblem, oracle_discrete_dynamics, oracle_discrete_jacobian, oracle_dynamics,
                            oracle_grad_projection, oracle_hess_projection, oracle_projection)

RTOL = 1.5e-8


def rng():
    return np.random.default_rng(1)   # Random.seed!(1), test/runtests.jl:13


def psd(r, k):
    A = r.standard_normal((k, k))
    return A @ A.T + k * np.eye(k)


# ------------------------------------------------------------------------------------------------------------
# costs: test/cost_tests.jl:229-279
@pytest.mark.parametrize("kind", ["quadratic", "diagonal"])
def test_cost_value_gradient_hessian_closed_forms(kind):
    r = rng()
    model = TO.Cartpole()
    n, m = 4, 1
    if kind == "quadratic":
        Q, R, H = psd(r, n), psd(r, m), r.standard_normal((m, n))
        cost = TO.QuadraticCost(Q, R, H=H, q=r.standard_normal(n), r=r.standard_normal(m), c=r.standard_normal())
    else:
        Q, R, H = np.diag(r.random(n) + 0.1), np.diag(r.random(m) + 0.1), np.zeros((m, n))
        cost = TO.DiagonalCost(Q, R, q=r.standard_normal(n), r=r.standard_normal(m), c=r.standard_normal())
    q, rr, c = cost.q, cost.r, cost.c
    # knot 1 = stage knot z, knot 2 = terminal knot zterm (dt = 0 => is_terminal, test/cost_tests.jl:235-236)
    prob = OracleProblem(model, TO.Objective(cost, 2), np.zeros(n), 0.1)
    x, u = r.random(n), r.random(m)
    TO.initial_states(prob, np.stack([x, x]))
    TO.initial_controls(prob, u)
    Jk = TO.cost_knots(prob)[0]
    assert np.isclose(Jk[0], 0.5 * (x @ Q @ x + u @ R @ u) + q @ x + rr @ u + c + u @ H @ x, rtol=RTOL)   # :238-241
    grad = TO.cost_gradient(prob)[0]
    assert np.allclose(grad[0, :n], Q @ x + q + H.T @ u, rtol=RTOL)        # :249-251
    assert np.allclose(grad[0, n:], R @ u + rr + H @ x, rtol=RTOL)
    assert np.allclose(grad[1, :n], Q @ x + q, rtol=RTOL)                  # terminal :245-247
    assert np.allclose(grad[1, n:], 0.0)
    hess = TO.cost_hessian(prob)[0]
    assert np.allclose(hess[1, :n, :n], Q) and np.allclose(hess[1, n:, n:], 0.0)    # :253-256
    assert np.allclose(hess[0, n:, n:], R) and np.allclose(hess[0, n:, :n], H)      # :257-258
    assert np.allclose(hess[0, :n, n:], H.T)   # this build writes the symmetric counterpart too


# test/objective_tests.jl:86-96: LQRObjective parameters (host logic)
def test_lqr_objective_parameters():
    r = rng()
    n, m, N = 4, 1, 11
    Q, R, Qf = np.diag(r.random(n)), np.diag(r.random(m)), np.diag(r.random(n))
    xf = r.random(n)
    obj = TO.LQRObjective(Q, R, Qf, xf, N)
    assert len(obj) == N and isinstance(obj[0], TO.DiagonalCost)
    assert np.allclose(obj[0].Q, Q) and np.allclose(obj[0].q, -Q @ xf)
    assert np.allclose(obj[1].r, 0.0)
    assert np.isclose(obj[1].c, 0.5 * xf @ Q @ xf)
    assert np.isclose(obj[-1].c, 0.5 * xf @ Qf @ xf)
    assert np.allclose(obj[-1].q, -Qf @ xf) and np.allclose(obj[-1].Q, Qf) and np.allclose(obj[-1].R, R)
    obj2 = TO.LQRObjective(Q + 0.01 * np.ones((n, n)), R, Qf, xf, N)     # dense Q -> QuadraticCost (:98-99)
    assert isinstance(obj2[0], TO.QuadraticCost)


# test/objective_tests.jl:124-140: trajectory cost, N = 101, explicit (u - uref) LQR sum
def test_trajectory_cost_closed_form():
    r = rng()
    n, m, N = 4, 1, 101
    Q, R, Qf = np.diag(r.random(n)), np.diag(r.random(m)), np.diag(r.random(n))
    xf, uref = r.random(n), r.random(m)
    obj = TO.LQRObjective(Q, R, Qf, xf, N, uf=uref)
    prob = OracleProblem(TO.Cartpole(), obj, np.zeros(n), 1.0)
    X, U = r.random((N, n)), r.random((N - 1, m))
    TO.initial_states(prob, X)
    TO.initial_controls(prob, U)
    J = sum(0.5 * (X[k] - xf) @ Q @ (X[k] - xf) + 0.5 * (U[k] - uref) @ R @ (U[k] - uref) for k in range(N - 1))
    J += 0.5 * (X[-1] - xf) @ Qf @ (X[-1] - xf) + 0.5 * uref @ R @ uref   # terminal control = 0 keeps R, r (src/objective.jl:154,180)
    # the reference's terminal cost omits the constant 1/2 uf'R uf (cf = 1/2 xf'Qf xf) and evaluates u_N = 0:
    J -= 0.5 * uref @ R @ uref
    assert np.isclose(TO.cost(prob)[0], J, rtol=RTOL)
    assert np.isclose(TO.cost_knots(prob)[0].sum(), J, rtol=RTOL)


# ------------------------------------------------------------------------------------------------------------
# constraints: test/constraint_tests.jl
def _cartpole_problem(cons_builder, N=5):
    n, m = 4, 1
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n), np.zeros(n), N)
    cons = TO.ConstraintList(n, m, N)
    cons_builder(cons, n, m, N)
    return OracleProblem(TO.Cartpole(), obj, np.zeros(n), 1.0, constraints=cons)


def test_goal_constraint():   # :17-39
    r = rng()
    xf = r.random(4)
    goal = TO.GoalConstraint(xf)
    prob = _cartpole_problem(lambda c, n, m, N: TO.add_constraint(c, goal, N))
    X = r.random((5, 4))
    TO.initial_states(prob, X)
    assert np.allclose(TO.evaluate_constraints(prob, goal)[0, 0], X[-1] - xf)
    assert np.allclose(TO.constraint_jacobians(prob, goal)[0, 0], np.hstack([np.eye(4), np.zeros((4, 1))]))
    assert TO.output_dim(goal) == 4 and TO.is_bound(goal)
    assert np.allclose(TO.upper_bound(goal), 0) and np.allclose(TO.lower_bound(goal), 0)
    assert TO.sense(goal) == TO.Equality()
    part = TO.GoalConstraint(xf, inds=[1, 3])   # src/constraints.jl:27-29
    prob2 = _cartpole_problem(lambda c, n, m, N: TO.add_constraint(c, part, N))
    TO.initial_states(prob2, X)
    assert np.allclose(TO.evaluate_constraints(prob2, part)[0, 0], X[-1, [0, 2]] - xf[[0, 2]])


def test_bound_constraint():   # :209-266
    r = rng()
    n, m = 4, 1
    xmin, xmax, umin, umax = -r.random(n), r.random(n), -r.random(m), r.random(m)
    bnd = TO.BoundConstraint(n, m, x_min=xmin, x_max=xmax, u_min=umin, u_max=umax)
    prob = _cartpole_problem(lambda c, n_, m_, N: TO.add_constraint(c, bnd, (1, N - 1)))
    X, U = r.random((5, n)), r.random((4, m))
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    x, u = X[0], U[0]
    assert np.allclose(TO.evaluate_constraints(prob, bnd)[0, 0], np.concatenate([x - xmax, u - umax, xmin - x, umin - u]))   # :217
    assert np.allclose(TO.constraint_jacobians(prob, bnd)[0, 0], np.vstack([np.eye(n + m), -np.eye(n + m)]))                # :221
    assert TO.output_dim(bnd) == 2 * (n + m)
    assert np.array_equal(TO.upper_bound(bnd), np.concatenate([xmax, umax])) and np.array_equal(TO.lower_bound(bnd), np.concatenate([xmin, umin]))
    # +-Inf pruning :228-245
    xmin2 = xmin.copy(); xmin2[0] = -np.inf
    umax2 = np.array([np.inf])
    bnd2 = TO.BoundConstraint(n, m, x_min=xmin2, x_max=xmax, u_min=umin, u_max=umax2)
    prob2 = _cartpole_problem(lambda c, n_, m_, N: TO.add_constraint(c, bnd2, (1, N - 1)))
    TO.initial_states(prob2, X); TO.initial_controls(prob2, U)
    assert TO.output_dim(bnd2) == 2 * (n + m) - 2
    assert np.allclose(TO.evaluate_constraints(prob2, bnd2)[0, 0], np.concatenate([x - xmax, xmin[1:] - x[1:], umin - u]))
    iz = np.ones(2 * (n + m), dtype=bool); iz[n] = False; iz[n + m] = False
    assert np.allclose(TO.constraint_jacobians(prob2, bnd2)[0, 0], np.vstack([np.eye(n + m), -np.eye(n + m)])[iz])
    with pytest.raises(TO.ArgumentError):   # :265
        TO.BoundConstraint(n, m, x_min=10, x_max=-10, u_min=umin, u_max=umax)


def test_state_and_control_bound_integer_kats():
    """StateBound / ControlBound literal KATs (test/constraint_tests.jl:280,292,302,319,331,341), expressed through
    BoundConstraint on the quadrotor dims with every other entry unbounded."""
    n, m, N = 13, 4, 3
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n), np.zeros(n), N)

    def run(bnd, x3=None, u3=None):
        cons = TO.ConstraintList(n, m, N)
        TO.add_constraint(cons, bnd, (1, N - 1))
        prob = OracleProblem(TO.Quadrotor(), obj, np.zeros(n), 1.0, constraints=cons)
        X, U = np.zeros((N, n)), np.zeros((N - 1, m))
        if x3 is not None: X[:, :3] = x3
        if u3 is not None: U[:, :3] = u3
        TO.initial_states(prob, X); TO.initial_controls(prob, U)
        return TO.evaluate_constraints(prob, bnd)[0, 0], TO.constraint_jacobians(prob, bnd)[0, 0]

    inf = np.inf
    xm = lambda a: np.concatenate([a, np.full(10, inf)])
    xn = lambda a: np.concatenate([a, np.full(10, -inf)])
    c, J = run(TO.BoundConstraint(n, m, x_max=xm([10, 2, 5.]), x_min=xn([0, -3, -4.])), x3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -1, -3, 0, -4, -6])                                          # :280
    assert np.array_equal(J[:, :3], np.vstack([np.eye(3), -np.eye(3)])) and not J[:, 3:].any()   # :282-285
    c, _ = run(TO.BoundConstraint(n, m, x_max=xm([10, 2, 5.]), x_min=xn([-inf, -3, -4.])), x3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -1, -3, -4, -6])                                             # :292
    c, _ = run(TO.BoundConstraint(n, m, x_max=xm([10, 10, 10.])), x3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -9, -8])                                                     # :302
    um = lambda a: np.concatenate([a, [inf]])
    un = lambda a: np.concatenate([a, [-inf]])
    c, _ = run(TO.BoundConstraint(n, m, u_max=um([10, 2, 5.]), u_min=un([0, -3, -4.])), u3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -1, -3, 0, -4, -6])                                          # :319
    c, _ = run(TO.BoundConstraint(n, m, u_max=um([10, 2, 5.]), u_min=un([-inf, -3, -4.])), u3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -1, -3, -4, -6])                                             # :331
    c, _ = run(TO.BoundConstraint(n, m, u_max=um([10, 10, 10.])), u3=[0, 1, 2.])
    assert np.array_equal(c, [-10, -9, -8])                                                     # :341


def test_collision_state_control_bound_api():
    """CollisionConstraint KAT (test/constraint_tests.jl:155-174) and the StateBound / ControlBound classes (:268-345)."""
    r = rng()
    n, m, N = 13, 4, 3
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n), np.zeros(n), N)
    col = TO.CollisionConstraint(n, [1, 2], [3, 4], 2.0)
    sb = TO.StateBound(n, x_max=np.concatenate([[10, 2, 5.], np.full(10, np.inf)]), x_min=np.concatenate([[0, -3, -4.], np.full(10, -np.inf)]))
    cb = TO.ControlBound(m, u_max=10)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, col, (1, N)); TO.add_constraint(cons, sb, (1, N)); TO.add_constraint(cons, cb, (1, N - 1))
    prob = OracleProblem(TO.Quadrotor(), obj, np.zeros(n), 1.0, constraints=cons)
    X, U = r.random((N, n)), np.tile([0, 1, 2, 3.], (N - 1, 1))
    X[:, :3] = [0, 1, 2.]
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    x = X[0]; d = x[[0, 1]] - x[[2, 3]]
    assert np.allclose(TO.evaluate_constraints(prob, col)[0, 0, 0], 4 - d @ d)                    # :161-163
    J = TO.constraint_jacobians(prob, col)[0, 0]
    assert np.allclose(J[0, :4], np.concatenate([-2 * d, 2 * d])) and not J[0, 4:].any()           # :166
    assert TO.output_dim(col) == 1 and isinstance(TO.sense(col), TO.Inequality)                     # :167
    with pytest.raises(TO.DimensionMismatch):
        TO.CollisionConstraint(n, [1, 2], [1, 2, 3], 1.0)                                           # :173
    assert np.array_equal(TO.evaluate_constraints(prob, sb)[0, 0], [-10, -1, -3, 0, -4, -6])        # :280
    assert TO.output_dim(sb) == 6 and TO.is_bound(sb)
    assert np.array_equal(TO.upper_bound(sb)[:3], [10, 2, 5]) and np.array_equal(TO.lower_bound(sb)[:3], [0, -3, -4])   # :277-278
    assert np.array_equal(TO.evaluate_constraints(prob, cb)[0, 0], [-10, -9, -8, -7])               # :341 (m = 4 here)
    assert np.array_equal(TO.upper_bound(cb), np.full(m, 10.0)) and np.all(np.isneginf(TO.lower_bound(cb)))
    Jb = TO.constraint_jacobians(prob, cb)[0, 0]
    assert np.array_equal(Jb[:, n:], np.eye(m)) and not Jb[:, :n].any()
    with pytest.raises(TO.ArgumentError):
        TO.StateBound(n, x_max=-10, x_min=10)                                                       # :304
    with pytest.raises(TO.ArgumentError):
        TO.ControlBound(m, u_max=-10, u_min=10)                                                     # :343
    assert np.array_equal(TO.num_constraints(prob), [1 + 6 + 4, 1 + 6 + 4, 1 + 6])


def test_circle_sphere_norm_linear_closed_forms():   # :43-205
    r = rng()
    n, m, N = 13, 4, 3
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n), np.zeros(n), N)
    xc, yc, zc, rad = r.random(3), r.random(3), r.random(3), r.random(3)
    circ = TO.CircleConstraint(n, xc, yc, rad)
    sph = TO.SphereConstraint(n, xc, yc, zc, rad)
    A, b = r.standard_normal((5, m)), r.standard_normal(5)
    lin = TO.LinearConstraint(n, m, A, b, TO.Inequality(), "control")
    nrm = TO.NormConstraint(n, m, 2.0, TO.SecondOrderCone(), "control")
    nrm2 = TO.NormConstraint(n, m, 3.0, TO.Inequality(), [1, 2, 3])
    cons = TO.ConstraintList(n, m, N)
    for c in (circ, sph, lin, nrm, nrm2):
        TO.add_constraint(cons, c, (1, N - 1))
    prob = OracleProblem(TO.Quadrotor(), obj, np.zeros(n), 1.0, constraints=cons)
    X, U = r.random((N, n)), r.random((N - 1, m))
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    x, u = X[0], U[0]
    assert np.allclose(TO.evaluate_constraints(prob, circ)[0, 0], -(x[0] - xc) ** 2 - (x[1] - yc) ** 2 + rad ** 2)      # :107-109
    Jc = TO.constraint_jacobians(prob, circ)[0, 0]
    assert np.allclose(Jc[:, 0], -2 * (x[0] - xc)) and np.allclose(Jc[:, 1], -2 * (x[1] - yc)) and not Jc[:, 2:].any()
    assert np.allclose(TO.evaluate_constraints(prob, sph)[0, 0], -(x[0] - xc) ** 2 - (x[1] - yc) ** 2 - (x[2] - zc) ** 2 + rad ** 2)
    assert np.allclose(TO.evaluate_constraints(prob, lin)[0, 0], A @ u - b)                                                   # :60-66
    assert np.allclose(TO.constraint_jacobians(prob, lin)[0, 0], np.hstack([np.zeros((5, n)), A]))
    assert np.allclose(TO.evaluate_constraints(prob, nrm)[0, 0], np.concatenate([u, [2.0]]))                                  # :187-188
    assert np.allclose(TO.evaluate_constraints(prob, nrm2)[0, 0], [x[:3] @ x[:3] - 9.0])                                     # :180
    Jn = TO.constraint_jacobians(prob, nrm2)[0, 0]
    assert np.allclose(Jn[0, :3], 2 * x[:3]) and not Jn[0, 3:].any()
    assert np.array_equal(TO.num_constraints(prob), [3 + 3 + 5 + 5 + 1] * 2 + [0])


# ------------------------------------------------------------------------------------------------------------
# cones: test/cone_tests.jl:8-75, reference closed forms test/socp.jl:6-50
def Pi_soc(x):
    v, s = x[:-1], x[-1]
    a = np.linalg.norm(v)
    if a <= -s: return np.zeros_like(x)
    if a <= s: return x.copy()
    return 0.5 * (1 + s / a) * np.concatenate([v, [a]])


def numjac(f, x, h=1e-6):
    J = np.zeros((f(x).size, x.size))
    for i in range(x.size):
        e = np.zeros_like(x); e[i] = h
        J[:, i] = (f(x + e) - f(x - e)) / (2 * h)
    return J


@pytest.mark.parametrize("x", [[2, 3, 1, 1.0], [2, 3, 1, -10.0], [2, 3, 1, 10.0]])   # :51,58,64 (outside / below / in)
def test_soc_projection_and_derivatives(x):
    x = np.array(x)
    b = rng().standard_normal(4)
    cone = TO.SecondOrderCone()
    px, rc = oracle_projection(cone, x)
    assert rc == 0 and np.allclose(px[0], Pi_soc(x), rtol=RTOL)
    J, _ = oracle_grad_projection(cone, x)
    assert np.allclose(J[0], numjac(Pi_soc, x), atol=1e-6)                             # J ~ ForwardDiff.jacobian(Pi, x) :39
    H, _ = oracle_hess_projection(cone, x, b)
    assert np.allclose(H[0], numjac(lambda y: numjac(Pi_soc, y, 1e-5).T @ b, x, 1e-4), atol=1e-4)   # :40


def test_orthant_projection_and_derivatives():   # :70-75
    x = np.array([1, 2, -3.0])
    cone = TO.Inequality()
    px, _ = oracle_projection(cone, x)
    assert np.array_equal(px[0], np.minimum(0, x))
    J, _ = oracle_grad_projection(cone, x)
    assert np.array_equal(J[0], np.diag([0, 0, 1.0]))
    H, _ = oracle_hess_projection(cone, x, np.ones(3))
    assert not H.any()
    pz, _ = oracle_projection(TO.ZeroCone(), x)
    assert not pz.any()
    assert TO.dualcone(TO.ZeroCone()) == TO.IdentityCone() and TO.dualcone(TO.Inequality()) == TO.Inequality()   # src/cones.jl:65-69


# ------------------------------------------------------------------------------------------------------------
# examples/quickstart.jl (the reference's end-to-end API tour with @tests)
def test_quickstart():
    r = rng()
    model = TO.DoubleIntegrator(2)
    n, m = model.dims()
    tf, N = 3.0, 21
    dt = tf / (N - 1)
    assert (n, m) == (4, 2)                                                                   # :32
    x0, xf = np.zeros(4), np.array([0, 2.0, 0, 0])
    Q, R = np.eye(n), np.eye(m)
    Qf = Q * (N - 1)
    obj = TO.LQRObjective(Q, R, Qf, xf, N)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.GoalConstraint(xf), N)                                          # :53-56
    TO.add_constraint(cons, TO.CircleConstraint(n, [0.0], [1.0], [0.5]), (2, N - 1))
    TO.add_constraint(cons, TO.NormConstraint(n, m, 5.0, TO.SecondOrderCone(), "control"), (1, N - 1))
    TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-10, u_max=10), (1, N - 1))
    prob = OracleProblem(model, obj, x0, tf, xf=xf, constraints=cons)
    U0 = r.standard_normal((N - 1, m))
    TO.initial_controls(prob, U0)
    X0 = np.zeros((N, n))
    TO.initial_states(prob, X0)
    J = sum(0.5 * (X0[k] - xf) @ Q @ (X0[k] - xf) + 0.5 * U0[k] @ R @ U0[k] for k in range(N - 1)) + 0.5 * (X0[-1] - xf) @ Qf @ (X0[-1] - xf)
    assert np.isclose(TO.cost(prob)[0], J, rtol=RTOL)                                          # :71-80
    TO.rollout(prob)
    # hand-rolled RK4 of the (linear) double integrator: x+ = x + dt*[v; u] + dt^2/2*[u; 0]
    Xr = np.zeros((N, n))
    for k in range(N - 1):
        x, u = Xr[k], U0[k]
        Xr[k + 1] = np.concatenate([x[:2] + dt * x[2:] + 0.5 * dt * dt * u, x[2:] + dt * u])
    assert np.allclose(TO.states(prob)[0], Xr, rtol=RTOL, atol=1e-12)                          # :83-96
    assert np.allclose(TO.gettimes(prob), np.arange(N) * dt)
    assert [TO.sense(c) for c in cons] == [TO.ZeroCone(), TO.NegativeOrthant(), TO.SecondOrderCone(), TO.NegativeOrthant()]   # :121-122
    assert [TO.is_bound(c) for c in cons] == [True, False, False, True]                        # :128-129
    lower = np.concatenate([TO.lower_bound(c) for c in cons])
    upper = np.concatenate([TO.upper_bound(c) for c in cons])
    assert np.allclose(lower, np.concatenate([np.zeros(n), [-np.inf], np.full(m + 1, -np.inf), np.full(n, -np.inf), np.full(m, -10)]))   # :134
    assert np.allclose(upper, np.concatenate([np.zeros(n), [0.0], np.full(m + 1, np.inf), np.full(n, np.inf), np.full(m, 10)]))          # :135
    # trajectory-level sweeps :169-197
    vals = TO.evaluate_constraints(prob, 0)
    assert vals.shape == (1, 1, n) and np.allclose(vals[0, 0], Xr[-1] - xf)
    nv = TO.evaluate_constraints(prob, 2)
    assert nv.shape == (1, N - 1, m + 1) and np.allclose(nv[0, :, :m], U0) and np.allclose(nv[0, :, m], 5.0)
    jac = TO.constraint_jacobians(prob, 2)
    assert np.allclose(jac[0, 3], np.vstack([np.hstack([np.zeros((m, n)), np.eye(m)]), np.zeros((1, n + m))]))
    assert np.array_equal(TO.num_constraints(prob), [m + 1 + 2 * m] + [1 + m + 1 + 2 * m] * (N - 2) + [n])


# ------------------------------------------------------------------------------------------------------------
# dynamics / RK4 / Jacobians (parity unpinned in the reference -> independent checks)
def cartpole_numpy(x, u, mc=1.0, mp=0.2, l=0.5, g=9.81):   # docs/src/model.md:32-51 with numpy linear algebra
    q, qd = x[:2], x[2:]
    s, c = np.sin(q[1]), np.cos(q[1])
    H = np.array([[mc + mp, mp * l * c], [mp * l * c, mp * l ** 2]])
    Cm = np.array([[0, -mp * qd[1] * l * s], [0, 0]])
    G = np.array([0, mp * g * l * s])
    Bv = np.array([1.0, 0])
    qdd = -np.linalg.solve(H, Cm @ qd + G - Bv * u[0])
    return np.concatenate([qd, qdd])


def rk4_numpy(f, x, u, h):
    k1 = f(x, u) * h; k2 = f(x + k1 / 2, u) * h; k3 = f(x + k2 / 2, u) * h; k4 = f(x + k3, u) * h
    return x + (k1 + 2 * k2 + 2 * k3 + k4) / 6


def test_cartpole_dynamics_and_rk4_vs_numpy():
    r = rng()
    for _ in range(5):
        x, u = r.standard_normal(4), r.standard_normal(1)
        assert np.allclose(oracle_dynamics(TO.Cartpole(), x, u), cartpole_numpy(x, u), rtol=1e-12)
        assert np.allclose(oracle_discrete_dynamics(TO.Cartpole(), x, u, 0.05), rk4_numpy(cartpole_numpy, x, u, 0.05), rtol=1e-12)


def test_rk4_order_against_scipy():
    from scipy.integrate import solve_ivp
    x, u = np.array([0.1, 0.5, -0.2, 0.3]), np.array([0.7])
    ref = solve_ivp(lambda t, y: cartpole_numpy(y, u), (0, 0.02), x, rtol=1e-13, atol=1e-15, method="DOP853").y[:, -1]
    e1 = np.abs(oracle_discrete_dynamics(TO.Cartpole(), x, u, 0.02) - ref).max()
    half = oracle_discrete_dynamics(TO.Cartpole(), oracle_discrete_dynamics(TO.Cartpole(), x, u, 0.01), u, 0.01)
    e2 = np.abs(half - ref).max()
    assert e1 < 1e-6 and 8 < e1 / e2 < 40   # 4th-order: halving h cuts the error ~16x


def quadrotor_numpy(x, u, mass=0.5, J=np.diag([0.0023, 0.0023, 0.004]), g=np.array([0, 0, -9.81]), L=0.175, kf=1.0, km=0.0245):
    """examples/Quadrotor.ipynb cells 4,8 + rigid-body equations, written with rotation matrices (independent of
    the oracle's quaternion-product form)."""
    q, v, w = x[3:7], x[7:10], x[10:13]
    qw, qv = q[0], q[1:]
    F = np.maximum(0, kf * u)
    Fb = np.array([0, 0, F.sum()])
    skew = lambda a: np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    Rm = (qw * qw - qv @ qv) * np.eye(3) + 2 * np.outer(qv, qv) + 2 * qw * skew(qv)   # rotation (scaled if |q| != 1)
    Fw = mass * g + Rm @ Fb
    M = km * u
    tau = np.array([L * (F[1] - F[3]), L * (F[2] - F[0]), M[0] - M[1] + M[2] - M[3]])
    qdot = 0.5 * np.concatenate([[-qv @ w], qw * w + np.cross(qv, w)])
    return np.concatenate([v, qdot, Fw / mass, np.linalg.solve(J, tau - np.cross(w, J @ w))])


def test_quadrotor_dynamics_vs_numpy_and_hover():
    r = rng()
    model = TO.Quadrotor()
    for _ in range(5):
        x, u = r.standard_normal(13), r.random(4) * 3
        assert np.allclose(oracle_dynamics(model, x, u), quadrotor_numpy(x, u), rtol=1e-11, atol=1e-12)
    x0 = np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    assert np.allclose(oracle_dynamics(model, x0, model.hover_control()), 0.0, atol=1e-14)


def test_hover_rollout_stays_at_x0():   # test/internal_api.jl:50-56 (N = 51, tf = 5)
    model = TO.Quadrotor()
    n, m, N = 13, 4, 51
    x0 = np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    obj = TO.LQRObjective(np.full(n, 0.1), np.full(m, 0.01), np.full(n, 100.0), xf, N)
    prob = OracleProblem(model, obj, x0, 5.0, xf=xf)
    assert np.isnan(TO.states(prob)).all()                                   # X0 = NaN before rollout (test/problems_tests.jl:209)
    TO.initial_controls(prob, model.hover_control())
    TO.rollout(prob)
    X = TO.states(prob)[0]
    assert np.array_equal(X[0], x0) and np.allclose(X[-1], x0, rtol=RTOL)    # :51-52


@pytest.mark.parametrize("model", [TO.Cartpole(), TO.Quadrotor(), TO.Acrobot(), TO.DoubleIntegrator(2), TO.DoubleIntegrator(1)])
def test_discrete_jacobian_vs_central_differences(model):   # stale test/dynamics_constraints.jl:57-71 (AD vs FD at 1e-6)
    r = rng()
    n, m = model.dims()
    x, u = r.standard_normal(n) * 0.5, r.random(m) + 0.5
    AB = oracle_discrete_jacobian(model, x, u, 0.05)
    f = lambda z: oracle_discrete_dynamics(model, z[:n], z[n:], 0.05)
    assert AB.shape == (n, n + m)
    assert np.allclose(AB, numjac(f, np.concatenate([x, u]), 1e-6), atol=1e-6)


def test_relu_tie_convention():
    """max(0, kf*w) at w == 0: derivative 0 (the constant wins the tie), SURVEY.md section 7."""
    model = TO.Quadrotor()
    x = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    AB = oracle_discrete_jacobian(model, x, np.zeros(4), 0.05)
    # thrust derivative vanishes at w = 0, the motor-torque (km*w) derivative of the yaw rate does not
    assert np.allclose(AB[7:10, 13:], 0.0) and np.allclose(AB[10:12, 13:], 0.0) and np.all(np.abs(AB[12, 13:]) > 0)
    ABp = oracle_discrete_jacobian(model, x, np.full(4, 1e-9), 0.05)
    assert np.all(ABp[9, 13:] > 0)


# ---- IndexedConstraint / change_dimension (src/constraints.jl:785-936; test/constraint_tests.jl:346-407) -----------------------------

def test_indexed_constraint_reference_test():
    """test/constraint_tests.jl:346-407: a BoundConstraint of (n, m) applied to the leading slices of a (2n, 2m) knot point
    z2 = ([x; 2x], [u; 2u]), and a CircleConstraint applied to the trailing state slice (where it sees 2x)."""
    r = np.random.default_rng(4)
    n, m, N, B = 4, 2, 3, 2
    n2, m2 = 2 * n, 2 * m
    x, u = r.standard_normal((B, N, n)), r.standard_normal((B, N - 1, m))
    X2, U2 = np.concatenate([x, 2 * x], axis=-1), np.concatenate([u, 2 * u], axis=-1)
    xmin, xmax, umin, umax = -r.random(n), r.random(n), -r.random(m), r.random(m)
    bnd = TO.BoundConstraint(n, m, x_min=xmin, x_max=xmax, u_min=umin, u_max=umax)
    cir = TO.CircleConstraint(n, [1.0, 1, 1], [1.0, 2, 3], [1.0, 1, 1])
    idx = TO.IndexedConstraint(n2, m2, bnd)                                   # :357
    idc = TO.IndexedConstraint(n2, m2, cir, (n + 1, 2 * n), (m + 1, 2 * m))   # :383
    assert TO.output_dim(idx) == TO.output_dim(bnd) and TO.sense(idx) == TO.sense(bnd) and (idx.n, idx.m) == (n2, m2)      # :372-375
    assert np.array_equal(TO.upper_bound(idx), TO.upper_bound(bnd)) and np.array_equal(TO.lower_bound(idx), TO.lower_bound(bnd))   # :376-377
    assert TO.is_bound(idx) and TO.is_bound(idc) == TO.is_bound(cir)                                                        # :378, :404

    def problem(nn, mm, cons_list, X, U):
        cons = TO.ConstraintList(nn, mm, N)
        for c in cons_list:
            TO.add_constraint(cons, c, (1, N - 1))
        obj = TO.LQRObjective(np.ones(nn), np.ones(mm), np.ones(nn), np.zeros(nn), N)
        p = OracleProblem(_Dummy(nn, mm), obj, np.zeros(nn), 1.0, constraints=cons, batch=B)
        TO.initial_states(p, X); TO.initial_controls(p, U)
        return p
    big = problem(n2, m2, [idx, idc], X2, U2)
    small = problem(n, m, [bnd], x, u)
    small2 = problem(n, m, [cir], 2 * x, 2 * u)
    assert np.allclose(TO.evaluate_constraints(big, 0), TO.evaluate_constraints(small, 0), rtol=1e-14)      # :360-363
    J, J0 = TO.constraint_jacobians(big, 0), TO.constraint_jacobians(small, 0)
    p = bnd.p
    want = np.concatenate([J0[..., :n], np.zeros((B, N - 1, p, n)), J0[..., n:], np.zeros((B, N - 1, p, m))], axis=-1)     # :369
    assert np.array_equal(J, want)
    assert np.allclose(TO.evaluate_constraints(big, 1), TO.evaluate_constraints(small2, 0), rtol=1e-14)     # :388-390
    Jc, Jc0 = TO.constraint_jacobians(big, 1), TO.constraint_jacobians(small2, 0)
    wantc = np.concatenate([np.zeros((B, N - 1, 3, n)), Jc0[..., :n], np.zeros((B, N - 1, 3, m2))], axis=-1)               # :396
    assert np.allclose(Jc, wantc, rtol=1e-14)


class _Dummy(TO.DoubleIntegrator):
    """a double integrator of dimension n/2 when m == n/2, otherwise only a carrier of (n, m) for constraint evaluation"""

    def __init__(self, n, m):
        super().__init__(m)
        assert n == 2 * m


def test_change_dimension_of_constraints_lists_and_costs():
    r = np.random.default_rng(6)
    n, m, N = 4, 2, 4
    n2, m2 = 8, 4
    ix, iu = (5, 8), (3, 4)
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.GoalConstraint(np.arange(1.0, 5.0), inds=[1, 3]), N)
    TO.add_constraint(cons, TO.NormConstraint(n, m, 2.0, TO.SecondOrderCone(), [2, 5, 6]), (1, N - 1))
    TO.add_constraint(cons, TO.LinearConstraint(n, m, r.standard_normal((2, m)), r.standard_normal(2), TO.Inequality(), "control"), (1, N - 1))
    TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x, u: [x[0] * u[1] - TO.sin(x[3])], TO.Equality()), (2, 3))
    TO.add_constraint(cons, TO.StateBound(n, x_max=[1.0, np.inf, 2.0, np.inf]), (1, N))
    TO.add_constraint(cons, TO.ControlBound(m, u_min=-1.0), (1, N - 1))
    big_cons = TO.change_dimension(cons, n2, m2, ix, iu)                     # src/constraint_list.jl:208-217
    assert len(big_cons) == len(cons) and np.array_equal(TO.num_constraints(big_cons), TO.num_constraints(cons))
    cost = TO.DiagonalCost(r.random(n), r.random(m), q=r.standard_normal(n), r=r.standard_normal(m), c=0.3)
    big_cost = TO.change_dimension(cost, n2, m2, ix, iu)                     # src/cost_functions.jl:391-401
    assert np.allclose(np.diag(big_cost.Q)[4:], np.diag(cost.Q)) and not np.diag(big_cost.Q)[:4].any() and np.allclose(big_cost.r[2:], cost.r)
    qc = TO.change_dimension(TO.QuatLQRCost(np.ones(4), np.ones(2), np.array([1.0, 0, 0, 0]), w=3.0, quat_ind=[1, 2, 3, 4]), n2, m2, ix, iu)
    assert isinstance(qc, TO.DiagonalQuatCost) and np.array_equal(qc.q_ind, [5, 6, 7, 8]) and qc.w == 3.0          # src/lie_costs.jl:144-159
    B = 2
    x, u = r.standard_normal((B, N, n)), r.standard_normal((B, N - 1, m))
    X2, U2 = r.standard_normal((B, N, n2)), r.standard_normal((B, N - 1, m2))
    X2[..., 4:] = x; U2[..., 2:] = u
    small = OracleProblem(_Dummy(n, m), TO.Objective(cost, N), np.zeros(n), 1.0, constraints=cons, batch=B)
    big = OracleProblem(_Dummy(n2, m2), TO.Objective(big_cost, N), np.zeros(n2), 1.0, constraints=big_cons, batch=B)
    TO.initial_states(small, x); TO.initial_controls(small, u); TO.initial_states(big, X2); TO.initial_controls(big, U2)
    assert np.allclose(TO.cost(big), TO.cost(small), rtol=1e-13)
    for i in range(len(cons)):
        assert np.allclose(TO.evaluate_constraints(big, i), TO.evaluate_constraints(small, i), rtol=1e-13, atol=1e-15), i
        Jb, Js = TO.constraint_jacobians(big, i), TO.constraint_jacobians(small, i)
        assert np.allclose(Jb[..., 4:8], Js[..., :4], atol=1e-15) and np.allclose(Jb[..., 10:12], Js[..., 4:6], atol=1e-15)
        assert not Jb[..., :4].any() and not Jb[..., 8:10].any()
    assert np.allclose(TO.merit(big), TO.merit(small), rtol=1e-13) and np.allclose(TO.max_violation(big), TO.max_violation(small), rtol=1e-13)
    with pytest.raises(TO.DimensionMismatch):
        TO.IndexedConstraint(n2, m2, TO.BoundConstraint(n, m, u_max=1.0), (1, 3), (1, 2))
    with pytest.raises(TO.ArgumentError):
        TO.IndexedConstraint(n2, m2, TO.GoalConstraint(np.zeros(n)), [4, 3, 2, 1])


def test_problem_constructor_and_setters_reference_test():
    """test/problems_tests.jl:1-215 against the host API (cartpole, N = 11, tf = 5): fields, defaults, time steps, initial / goal
    state setters and what they modify, dims, cost before / after the rollout, setinitialtime!"""
    import copy
    n, m, N, tf = 4, 1, 11, 5.0
    model = TO.Cartpole()
    x0, xf = np.zeros(n), np.array([0, np.pi, 0, 0])
    Q, Qf, R = 1e-2 * np.ones(n), 100.0 * np.ones(n), 1e-1 * np.ones(m)
    obj = TO.LQRObjective(Q, R, Qf, xf, N)
    conSet = TO.ConstraintList(n, m, N)
    bnd, goal = TO.BoundConstraint(n, m, u_min=-3.0, u_max=3.0), TO.GoalConstraint(xf)
    TO.add_constraint(conSet, bnd, (1, N - 1)); TO.add_constraint(conSet, goal, (N, N))
    X0, U0 = np.zeros((N, n)), np.full((N - 1, m), 0.01)
    prob = OracleProblem(model, obj, x0, tf, xf=xf, constraints=conSet, X0=X0, U0=U0)                    # keyword constructor :64
    assert np.array_equal(prob.x0[0], x0) and np.array_equal(prob.xf, xf) and TO.get_constraints(prob) is conSet and prob.obj is obj
    assert prob.N == N and np.allclose(TO.states(prob)[0], X0) and np.allclose(TO.controls(prob)[0], U0)   # :65-76
    assert np.isclose(TO.gettimes(prob)[-1], tf)
    dts = np.concatenate([np.full(N // 2, 1.0), np.full(N - N // 2 - 1, 0.5)]); dts *= tf / dts.sum()
    p2 = OracleProblem(model, obj, x0, tf, xf=xf, constraints=conSet, dt=dts)
    assert np.allclose(TO.gettimes(p2), np.concatenate([[0], np.cumsum(dts)]))                             # :79-82
    with pytest.raises(TO.ArgumentError):                                                                   # :85 (AssertionError in the reference)
        OracleProblem(model, obj, x0, tf, xf=xf, constraints=conSet, dt=1.0)
    p3 = OracleProblem(model, obj, x0, tf)                                                                  # defaults :92-99
    assert not p3.x0.any() and p3.N == N and np.isnan(TO.states(p3)).all() and not TO.controls(p3).any()
    assert len(TO.get_constraints(p3)) == 0 and not TO.is_constrained(p3)
    assert np.allclose(TO.gettimes(p3), np.arange(N) * tf / (N - 1))
    TO.initial_states(p3, 2 * X0 + 1); TO.initial_controls(p3, 2 * U0)                                     # :102-105
    assert np.allclose(TO.states(p3)[0], 2 * X0 + 1) and np.allclose(TO.controls(p3)[0], 2 * U0)
    x0_ = np.random.default_rng(0).random(n)
    TO.set_initial_state(p3, x0_)                                                                           # :145-148
    assert np.allclose(TO.get_initial_state(p3)[0], x0_)
    with pytest.raises(TO.DimensionMismatch):
        TO.set_initial_state(p3, np.zeros(2 * n))
    # goal state: objective and terminal constraint of COPIES (:151-170)
    p4 = OracleProblem(model, obj.copy(), x0, tf, xf=xf, constraints=copy.deepcopy(conSet))
    xf_new = np.random.default_rng(1).random(n)
    TO.set_goal_state(p4, xf_new)
    assert np.allclose(p4.xf, xf_new) and np.allclose(p4.obj[0].q, -Q * xf_new) and np.allclose(p4.constraints[1].xf, xf_new)
    assert np.allclose(obj[0].q, -Q * xf) and np.allclose(conSet[1].xf, xf)                                # the originals are untouched
    TO.rollout(p4)
    lam_free_merit = TO.cost(p4)
    assert np.allclose(TO.evaluate_constraints(p4, 1)[0, 0], TO.states(p4)[0, -1] - xf_new)                # the device-side constraint moved too
    p5 = OracleProblem(model, obj.copy(), x0, tf, xf=xf, constraints=copy.deepcopy(conSet))
    TO.set_goal_state(p5, xf_new, constraint=False)                                                         # :173-177
    assert np.allclose(p5.obj[0].q, -Q * xf_new) and np.allclose(p5.obj[-1].q, -Qf * xf_new) and np.allclose(p5.constraints[1].xf, xf)
    p6 = OracleProblem(model, obj.copy(), x0, tf, xf=xf)
    TO.set_goal_state(p6, xf_new, objective=False)                                                          # :180-184
    assert np.allclose(p6.xf, xf_new) and np.allclose(p6.obj[0].q, -Q * xf) and np.allclose(p6.obj[-1].q, -Qf * xf)
    cs = copy.deepcopy(conSet)
    p7 = OracleProblem(model, obj, x0, tf, xf=xf, constraints=cs)                                           # not copied: the originals change :187-191
    TO.set_goal_state(p7, xf_new)
    assert np.allclose(obj[0].q, -Q * xf_new) and np.allclose(obj[-1].q, -Qf * xf_new) and np.allclose(cs[1].xf, xf_new)
    assert TO.dims(p7) == (n, m, N)                                                                         # :200-207
    p8 = OracleProblem(model, TO.LQRObjective(Q, R, Qf, xf, N), x0, tf)
    assert np.isnan(TO.cost(p8)).all()                                                                      # :210-212
    TO.rollout(p8)
    assert np.isfinite(TO.cost(p8)).all() and np.isfinite(lam_free_merit).all()
    assert np.isclose(TO.setinitialtime(p8, 1.0), 1.0 + tf) and np.isclose(TO.gettimes(p8)[0], 1.0)        # :215-216


def test_problem_copy_getters_and_cost_checks():
    """Problem(p; ...) / copy(prob) (src/problem.jl:125-128, :342-345), get_initial_time / get_final_time (:189-196), RD.state_dim(prob, k)
    (:149-150), initial_trajectory! (:242-245), and the definiteness warnings of the cost constructors (src/cost_functions.jl:337-343)"""
    prob = TO.problems.cartpole(B=2, N=11, cls=OracleProblem, u_bound=3.0, goal=True)
    TO.rollout(prob); TO.ilqr_step(prob, 1)
    cp = TO.copy_problem(prob)
    assert cp.obj is not prob.obj and cp.constraints is not prob.constraints and cp.constraints[0] is prob.constraints[0]
    assert np.array_equal(TO.states(cp), TO.states(prob)) and np.array_equal(TO.controls(cp), TO.controls(prob))
    assert np.allclose(TO.cost(cp), TO.cost(prob), rtol=1e-14) and np.allclose(TO.gettimes(cp), TO.gettimes(prob))
    assert TO.get_initial_time(cp) == 0.0 and np.isclose(TO.get_final_time(cp), 5.0)
    assert TO.state_dim(cp, 3) == 4 and TO.control_dim(cp, 3) == 1 and TO.horizonlength(cp) == 11
    TO.set_goal_state(cp, np.array([0.0, 1.0, 0, 0]))                    # the copy's objective moves, the original's does not
    assert not np.allclose(cp.obj[0].q, prob.obj[0].q)
    X, U, t = TO.get_trajectory(prob)
    TO.initial_trajectory(cp, 2 * X, 3 * U)
    assert np.allclose(TO.states(cp), 2 * X) and np.allclose(TO.controls(cp), 3 * U) and t.shape == (11,)
    with pytest.warns(UserWarning, match="positive semi-definite"):
        TO.DiagonalCost([-1.0, 1.0], [1.0])
    with pytest.warns(UserWarning, match="positive definite"):
        TO.DiagonalCost([1.0, 1.0], [0.0])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        TO.DiagonalCost([1.0, 1.0], [0.0], terminal=True)               # no check on R for terminal costs
        TO.DiagonalCost([-1.0, 1.0], [0.0], checks=False)


def test_cost_constructors_and_math_reference_test():
    """test/cost_tests.jl:44-146 (constructors, definiteness warnings, LQRCost) and :148-196 (addition, inversion) on the host API"""
    import warnings
    r = np.random.default_rng(8)
    n, m = 12, 6
    Q, R, Qf = np.full(n, 0.1), np.full(m, 0.01), np.full(n, 10.0)
    H, q, rr, c = r.random((m, n)), r.random(n), r.random(m), 0.37
    xf = np.ones(n)
    qc = TO.QuadraticCost(np.diag(Q), np.diag(R))                                         # :45-53
    assert np.array_equal(qc.Q, np.diag(Q)) and not qc.q.any() and not qc.r.any() and qc.c == 0
    assert (qc.state_dim, qc.control_dim) == (n, m) and TO.is_blockdiag(qc)
    qc = TO.QuadraticCost(np.diag(Q), np.diag(R), H=H, q=q, r=rr, c=c, terminal=True)     # :60-67
    assert np.allclose(qc.H, H) and np.array_equal(qc.q, q) and np.array_equal(qc.r, rr) and qc.c == c and qc.terminal and not TO.is_blockdiag(qc)
    with pytest.warns(UserWarning, match="R is not positive definite"):                   # :95
        TO.QuadraticCost(np.diag(Q), np.diag(R) * 0, H=H, q=q, r=rr, c=c)
    with pytest.warns(UserWarning, match="Q is not positive semidefinite"):               # :97-104
        TO.QuadraticCost(np.diag(Q) - 0.2, np.diag(R), H=H, q=q, r=rr, c=c)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        TO.QuadraticCost(np.diag(Q) * 0, np.diag(R), H=H, q=q, r=rr, c=c)                 # :96 @test_nowarn
        TO.QuadraticCost(np.diag(Q) - 0.2, np.diag(R), q=q, checks=False)                 # :111
    dc = TO.DiagonalCost(Q, R, q=q)                                                       # :114-135 (diagonal vectors or matrices)
    dc2 = TO.DiagonalCost(np.diag(Q), np.diag(R), q=q)
    assert np.array_equal(dc.Q, dc2.Q) and np.array_equal(dc.R, np.diag(R)) and not dc.r.any() and np.array_equal(dc.q, q)
    lq = TO.LQRCost(Q, R, xf)                                                             # :138-146
    assert isinstance(lq, TO.DiagonalCost) and np.allclose(lq.q, -Q * xf) and not lq.r.any() and np.isclose(lq.c, 0.5 * xf @ (Q * xf))
    assert TO.is_blockdiag(lq) and TO.is_diag(lq)
    # addition :150-168
    qcost = TO.QuadraticCost(np.diag(Q), np.diag(R), H=H, q=q, r=rr, c=c, checks=False)
    dcost = TO.DiagonalCost(Q, R, q=q)
    add = dcost + qcost
    assert isinstance(add, TO.QuadraticCost) and np.allclose(add.Q, 2 * np.diag(Q)) and np.allclose(add.R, 2 * np.diag(R))
    assert np.allclose(add.q, 2 * q) and np.allclose(add.r, rr) and np.isclose(add.c, c) and np.allclose(add.H, H)
    dcost = TO.DiagonalCost(Q, R, q=q, r=rr, c=c)
    add = dcost + dcost
    assert isinstance(add, TO.DiagonalCost) and np.allclose(add.r, 2 * rr) and np.isclose(add.c, 2 * c)
    with pytest.raises(TO.DimensionMismatch):                                             # :170-171 (AssertionError in the reference)
        TO.DiagonalCost(Q, Q, q=q, r=q, c=c) + dcost
    # inversion :173-195
    dinv = dcost.inv()
    assert np.allclose(np.diag(dinv.Q), 1 / Q) and np.allclose(np.diag(dinv.R), 1 / R) and np.array_equal(dinv.q, q) and TO.is_diag(dinv)
    qinv = qcost.inv()
    G = np.linalg.inv(np.block([[np.diag(Q), H.T], [H, np.diag(R)]]))
    assert np.allclose(qinv.Q, G[:n, :n]) and np.allclose(qinv.H, G[n:, :n]) and np.allclose(qinv.R, G[n:, n:])
    assert not TO.is_diag(qinv) and not TO.is_blockdiag(qinv) and np.array_equal(qinv.r, rr) and qinv.c == c
    binv = TO.QuadraticCost(np.diag(Q), np.diag(R)).inv()
    assert np.allclose(binv.Q, np.diag(1 / Q)) and TO.is_blockdiag(binv)


def test_objective_constructors_reference_test():
    """test/objective_tests.jl:17-84: Objective(cost, N) aliases one cost object, Objective(cost, term, N), Objective(costs), get_J"""
    r = np.random.default_rng(9)
    n, m, N = 4, 1, 11
    Q, R, Qf = np.full(n, 0.1), np.full(m, 0.01), np.full(n, 10.0)
    H, q, rr, c = r.random((m, n)), r.random(n), r.random(m), 0.4
    qcost = TO.QuadraticCost(np.diag(Q), np.diag(R), H=H, q=q, r=rr, c=c)
    obj = TO.Objective(qcost, N)
    assert len(obj) == N and TO.state_dim(obj, 1) == n and TO.control_dim(obj, 1) == m and TO.get_J(obj) is obj.J and not obj[-1].terminal   # :21-25
    assert obj[0] is obj[1] and obj[0].Q is obj[1].Q                                                                                          # :28-29
    qcost.r[0] = 1.0
    assert obj[2].r[0] == 1.0                                                                                                                 # :30-31
    qterm = TO.QuadraticCost(np.diag(Qf), np.zeros((m, m)), q=q, c=c, terminal=True)
    obj = TO.Objective(qcost, qterm, N)                                                                                                       # :36-44
    assert obj[0] is obj[1] and np.allclose(obj[-1].Q, np.diag(Qf)) and not obj[-1].R.any() and obj[-1].terminal and np.array_equal(obj[-1].q, q)
    costs = [(qcost if k < N - 1 else qterm).copy() for k in range(N)]
    obj = TO.Objective(costs)                                                                                                                 # :67-79
    assert len(obj) == N and obj[0] is not obj[1] and obj[0].Q is not obj[1].Q
    prob = OracleProblem(TO.Cartpole(), obj, np.zeros(n), 1.0)
    TO.initial_controls(prob, r.standard_normal((1, N - 1, m))); TO.rollout(prob)
    Jk = TO.cost_knots(prob)
    assert np.isclose(TO.get_J(prob.obj).sum(), TO.cost(prob)[0]) and np.array_equal(TO.get_J(prob.obj), Jk[0])                               # :139-140


def test_second_order_constraint_term_matches_finite_differences_of_the_jacobian():
    """`∇constraint_jacobians!` (src/abstract_constraint.jl:267-280): H = d/dz (∇c' λ).  Zero for Goal / Bound / Linear (src/constraints.jl:70-73,
    :767-770 set `H .= 0`); Circle / Sphere / Norm / Collision / QuatVecEq in closed form, user constraints by second-order forward mode -- all
    checked against central differences of the constraint Jacobian the same library returns."""
    from test_oracle_nlcost import control_norm
    n, m, N, B = 13, 4, 6, 2
    r = np.random.default_rng(8)
    base = TO.problems.quadrotor(B=B, N=N, cls=OracleProblem, dt=0.05)
    cons = base.constraints
    extra = [TO.SphereConstraint(n, [0.5, 0.2], [1.0, 0.5], [1.5, 1.2], [0.3, 0.25]), TO.NormConstraint(n, m, 3.0, TO.Inequality(), [8, 9, 10]),
             TO.NormConstraint(n, m, 12.0, TO.SecondOrderCone(), "control"), TO.CollisionConstraint(n, [1, 2, 3], [8, 9, 10], 0.4),
             TO.QuatVecEq(n, m, np.array([0.6, 0.0, 0.8, 0.0])), TO.AutodiffConstraint(n, m, control_norm(7.0), TO.Inequality(), "control"),
             TO.AutodiffConstraint(n, m, lambda x, u: [TO.sin(x[0]) * u[1] + x[3] * x[4] * x[5], TO.exp(0.1 * u[0]) - x[7] ** 3], TO.Equality())]
    for c in extra:
        TO.add_constraint(cons, c, (1, N - 1))
    prob = OracleProblem(base.model, base.obj, base.x0, 0.05 * (N - 1), xf=base.xf, constraints=cons)
    X = r.standard_normal((B, N, n)); X[..., 3:7] += [2.0, 0, 0, 0]
    U = 1.0 + r.standard_normal((B, N - 1, m))
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    eps = 1e-6
    for i, con in enumerate(prob.constraints.constraints):
        first, last = prob.constraints.inds[i]
        L = last - first + 1
        lam = r.standard_normal((B, L, con.p))
        H = TO.constraint_hessians(prob, i, lam)
        assert np.allclose(H, np.swapaxes(H, -1, -2), atol=1e-12)
        if isinstance(con, (TO.GoalConstraint, TO.BoundConstraint)) or (isinstance(con, TO.NormConstraint) and con.p > 1):
            assert not H.any()
            continue
        Hfd = np.zeros_like(H)
        for j in range(n + m):
            for sgn in (1.0, -1.0):
                Xp, Up = X.copy(), U.copy()
                if j < n: Xp[..., j] += sgn * eps
                else: Up[..., j - n] += sgn * eps
                TO.initial_states(prob, Xp); TO.initial_controls(prob, Up)
                J = TO.constraint_jacobians(prob, i)                     # [B, L, p, n+m]
                Hfd[:, :, j, :] += sgn * np.einsum("blp,blpz->blz", lam, J) / (2 * eps)
        TO.initial_states(prob, X); TO.initial_controls(prob, U)
        if last == N:                                                    # the terminal knot has no controls
            Hfd[:, -1, n:, :] = 0; Hfd[:, -1, :, n:] = 0; H = H.copy(); H[:, -1, n:, :] = 0; H[:, -1, :, n:] = 0
        assert np.allclose(H, Hfd, rtol=1e-6, atol=1e-6), (type(con).__name__, np.abs(H - Hfd).max())
    prob.close(); base.close()
