"""CPU tests of user-defined (automatically differentiated) costs, SURVEY.md 8 f4: the reference's RD.@autodiff CostFunction
interface (docs/src/costfunction_interface.md:30-50), pinned by the closed forms of test/nlcosts.jl:22-45."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem

P = TO.problems
rng = np.random.default_rng(11)


def cartpole_cost(Q, R):
    """test/nlcosts.jl:4-19 / docs/src/costfunction_interface.md:38-50 (0-based indices)"""
    def fun(x, u):
        y, th, ydot, thdot = x[0], x[1], x[2], x[3]
        J = Q[1] * np.cos(th / 2)
        J += 0.5 * (Q[0] * y ** 2 + Q[2] * ydot ** 2 + Q[3] * thdot ** 2)
        J += 0.5 * R[0] * u[0] ** 2
        return J
    return fun


def test_nonlinear_cost_reference_test():
    """test/nlcosts.jl:22-45: evaluate(cst, 0, 0) = 2; ForwardDiff gradient / Hessian == the closed forms."""
    n, m, N, B = 4, 1, 5, 3
    Q, R = [1, 2, 3, 4.0], [2.0]
    cst = TO.AutodiffCost(n, m, cartpole_cost(Q, R))
    assert cst.state_dim == 4 and cst.control_dim == 1
    prob = OracleProblem(TO.Cartpole(), TO.Objective(cst, N), np.zeros(n), 1.0, batch=B)
    TO.initial_states(prob, np.zeros((B, N, n))); TO.initial_controls(prob, np.zeros((B, N - 1, m)))
    assert np.allclose(TO.cost_knots(prob), 2.0)                                   # :27
    X, U = rng.random((B, N, n)), rng.random((B, N - 1, m))
    X[0, 0, 1] = np.pi
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    g, H = TO.cost_gradient(prob), TO.cost_hessian(prob)
    for b in range(B):
        for k in range(N - 1):
            x, u = X[b, k], U[b, k]
            assert np.allclose(H[b, k, :n, :n], np.diag([1, -0.5 * np.cos(x[1] / 2), 3, 4]), rtol=1e-13, atol=1e-15)    # :41
            assert np.allclose(H[b, k, n:, n:], [[2.0]]) and np.allclose(H[b, k, n:, :n], 0) and np.allclose(H[b, k, :n, n:], 0)   # :42
            assert np.allclose(g[b, k, :n], [x[0], -np.sin(x[1] / 2), 3 * x[2], 4 * x[3]], rtol=1e-13)                      # :43
            assert np.allclose(g[b, k, n:], 2 * u, rtol=1e-13)                                                              # :44
        # terminal knot: u = 0, state derivatives only
        x = X[b, -1]
        assert np.allclose(g[b, -1, :n], [x[0], -np.sin(x[1] / 2), 3 * x[2], 4 * x[3]]) and np.allclose(H[b, -1, n:, :], 0)
    Jk = TO.cost_knots(prob)
    Jx = 2 * np.cos(X[..., 1] / 2) + 0.5 * (X[..., 0] ** 2 + 3 * X[..., 2] ** 2 + 4 * X[..., 3] ** 2)
    assert np.allclose(Jk[:, :-1], Jx[:, :-1] + U[..., 0] ** 2, rtol=1e-13) and np.allclose(Jk[:, -1], Jx[:, -1], rtol=1e-13)


def test_recorded_program_ops_against_finite_differences():
    """every recorded operation, first and second derivatives vs central differences"""
    n, m, N = 4, 2, 3

    def fun(x, u):
        a = TO.sin(x[0]) * TO.cos(x[1]) + TO.exp(0.3 * x[2]) / (2.0 + x[3] ** 2)
        b = TO.log(1.5 + u[0] ** 2) + TO.sqrt(2.0 + x[0] * x[0]) - TO.tanh(u[1] - x[1])
        c = (1.0 + x[2] ** 2) ** 1.5 + (3.0 - u[0]) ** 3 - (-x[3]) + 2.0 / (1.0 + u[1] ** 2)
        return a + b * 0.7 + c

    def f(z):
        x, u = z[:n], z[n:]
        return fun(x, u)
    cst = TO.AutodiffCost(n, m, fun)
    prob = OracleProblem(TO.DoubleIntegrator(2), TO.Objective(cst, N), np.zeros(n), 1.0)
    X, U = 0.5 * rng.standard_normal((1, N, n)), 0.5 * rng.standard_normal((1, N - 1, m))
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    Jk, g, H = TO.cost_knots(prob), TO.cost_gradient(prob), TO.cost_hessian(prob)
    z = np.concatenate([X[0, 0], U[0, 0]])
    assert np.isclose(Jk[0, 0], f(z), rtol=1e-14)
    h = 1e-5
    E = h * np.eye(n + m)
    gfd = np.array([(f(z + e) - f(z - e)) / (2 * h) for e in E])
    Hfd = np.array([[(f(z + ei + ej) - f(z + ei - ej) - f(z - ei + ej) + f(z - ei - ej)) / (4 * h * h) for ej in E] for ei in E])
    assert np.allclose(g[0, 0], gfd, rtol=1e-7, atol=1e-8)
    assert np.allclose(H[0, 0], Hfd, rtol=1e-4, atol=1e-5)
    assert np.allclose(H[0, 0], H[0, 0].T)


def test_recording_limits_and_errors():
    with pytest.raises(TO.ArgumentError):
        TO.AutodiffCost(2, 1, lambda x, u: x[0] if x[0] else u[0])          # branching on a value
    with pytest.raises(TO.ArgumentError):
        def long(x, u):
            s = x[0]
            for _ in range(200):
                s = s * x[1] + 1.0
            return s
        TO.AutodiffCost(2, 1, long)
    c = TO.AutodiffCost(2, 1, lambda x, u: 3.0)                             # constant cost
    prob = OracleProblem(TO.DoubleIntegrator(1), TO.Objective(c, 3), np.zeros(2), 1.0)
    TO.rollout(prob)
    assert np.allclose(TO.cost(prob), 9.0) and not TO.cost_gradient(prob).any()


def lqr_as_autodiff(prob_quadratic):
    """the same objective as ``prob_quadratic`` (DiagonalCost stage + terminal), written as user functions"""
    stage, term = prob_quadratic.obj[0], prob_quadratic.obj[-1]
    n, m = prob_quadratic.n, prob_quadratic.m

    def make(c):
        Qd, Rd, q, r, cc = np.diag(c.Q), np.diag(c.R), c.q, c.r, c.c
        return lambda x, u: 0.5 * sum(Qd[i] * x[i] * x[i] for i in range(n)) + sum(q[i] * x[i] for i in range(n)) + cc + \
            0.5 * sum(Rd[j] * u[j] * u[j] for j in range(m)) + sum(r[j] * u[j] for j in range(m))
    return TO.Objective(TO.AutodiffCost(n, m, make(stage)), TO.AutodiffCost(n, m, make(term), terminal=True), prob_quadratic.N)


def test_autodiff_lqr_cost_reproduces_the_quadratic_solver_iterates():
    """an LQR objective written as a user function must drive AL-iLQR through the same iterates as the built-in DiagonalCost"""
    ref = P.cartpole(B=3, N=31, cls=OracleProblem, u_bound=3.0, goal=True)
    prob = OracleProblem(ref.model, lqr_as_autodiff(ref), ref.x0, 5.0, xf=ref.xf, constraints=ref.constraints)
    TO.initial_controls(prob, TO.controls(ref))
    for p in (ref, prob):
        TO.rollout(p)
    assert np.allclose(TO.cost(prob), TO.cost(ref), rtol=1e-12)
    g1, H1 = TO.al_expansion(prob); g2, H2 = TO.al_expansion(ref)
    assert np.allclose(g1, g2, rtol=1e-11, atol=1e-12) and np.allclose(H1, H2, rtol=1e-11, atol=1e-12)
    for p in (ref, prob):
        TO.ilqr_step(p, 4); TO.al_update(p); TO.ilqr_step(p, 2)
    assert np.allclose(TO.merit(prob), TO.merit(ref), rtol=1e-7)
    assert np.allclose(TO.controls(prob), TO.controls(ref), rtol=1e-6, atol=1e-8)
    with pytest.raises(TO.ArgumentError):
        TO.update_trajectory(prob, np.zeros((40, 4)), np.zeros((40, 1)))       # only for quadratic objectives (src/objective.jl:207)


def test_cartpole_swingup_with_the_docs_nonlinear_cost():
    """the docs' CartpoleCost (cos(theta/2) attitude term) as stage cost + quadratic terminal cost: iLQR decreases it monotonically"""
    n, m, N = 4, 1, 51
    stage = TO.AutodiffCost(n, m, cartpole_cost([0.1, 5.0, 0.1, 0.1], [0.05]))
    xf = np.array([0, np.pi, 0, 0])
    term = TO.LQRCost(np.full(n, 100.0), np.full(m, 0.05), xf, terminal=True)
    prob = OracleProblem(TO.Cartpole(), TO.Objective(stage, term, N), np.zeros((2, n)), 2.5)
    TO.initial_controls(prob, 0.5 + 0.1 * rng.standard_normal((2, N - 1, m))); TO.rollout(prob)
    J = [TO.merit(prob)]
    for _ in range(25):
        TO.ilqr_step(prob, 1); J.append(TO.merit(prob))
    J = np.array(J)
    assert np.all(np.diff(J, axis=0) <= 1e-9) and np.all(J[-1] < 0.5 * J[0])


# ---- user-defined constraints (docs/src/constraint_interface.md:52-72) ------------------------------------------------------------

def control_norm(a):
    """the docs' ControlNorm: ||u|| - a <= 0 (a ControlConstraint)"""
    return lambda u: [TO.sqrt(sum(ui * ui for ui in u)) - a]


def test_control_norm_constraint_of_the_docs():
    """value ||u|| - a and the analytic Jacobian u'/||u|| the docs give as the optional jacobian! (:69-71)"""
    n, m, N, B = 4, 2, 6, 3
    con = TO.AutodiffConstraint(n, m, control_norm(1.5), TO.Inequality(), inputs="control")
    assert con.p == 1 and TO.sense(con) == TO.Inequality()
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, con, (1, N - 1))
    obj = TO.LQRObjective(np.ones(n), np.ones(m), np.ones(n), np.zeros(n), N)
    prob = OracleProblem(TO.DoubleIntegrator(2), obj, np.zeros(n), 1.0, constraints=cons, batch=B)
    U = rng.standard_normal((B, N - 1, m))
    TO.initial_controls(prob, U); TO.rollout(prob)
    c, J = TO.evaluate_constraints(prob, con), TO.constraint_jacobians(prob, con)
    nu = np.linalg.norm(U, axis=-1)
    assert np.allclose(c[..., 0], nu - 1.5, rtol=1e-14)
    assert np.allclose(J[..., 0, n:], U / nu[..., None], rtol=1e-13) and not J[..., 0, :n].any()
    assert np.allclose(TO.max_violation(prob), np.maximum(nu - 1.5, 0).max(axis=1), rtol=1e-13)


def test_program_constraints_equal_the_builtin_ones():
    """CircleConstraint, a linear equality and an SOC norm constraint written as user functions give the values, Jacobians, AL merit
    and solver iterates of the built-in types"""
    n, m, N = 4, 2, 21
    xf = np.array([0, 2.0, 0, 0])
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n) * (N - 1), xf, N)
    U0 = np.random.default_rng(1).standard_normal((N - 1, m))

    def build(user):
        cons = TO.ConstraintList(n, m, N)
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
        if user:
            TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x: [-(x[0] - 0.0) ** 2 - (x[1] - 1.0) ** 2 + 0.5 * 0.5], TO.Inequality(), "state"), (2, N - 1))
            TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda u: [u[0], u[1], 5.0 + 0.0 * u[0]], TO.SecondOrderCone(), "control"), (1, N - 1))
            TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x, u: [1.0 * x[2] + 0.5 * u[0] - 0.1], TO.Equality()), (3, 5))
        else:
            TO.add_constraint(cons, TO.CircleConstraint(n, [0.0], [1.0], [0.5]), (2, N - 1))
            TO.add_constraint(cons, TO.NormConstraint(n, m, 5.0, TO.SecondOrderCone(), "control"), (1, N - 1))
            # (the reference's LinearConstraint acts on x or on u only, src/constraints.jl:103-150: the mixed row is a program on both sides)
            TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x, u: [1.0 * x[2] + 0.5 * u[0] - 0.1], TO.Equality()), (3, 5))
        p = OracleProblem(TO.DoubleIntegrator(2), obj, np.zeros(n), 3.0, xf=xf, constraints=cons)
        TO.initial_controls(p, U0); TO.rollout(p)
        return p

    a, b = build(True), build(False)
    for i in range(3):
        assert np.allclose(TO.evaluate_constraints(a, i), TO.evaluate_constraints(b, i), rtol=1e-13, atol=1e-15)
        assert np.allclose(TO.constraint_jacobians(a, i), TO.constraint_jacobians(b, i), rtol=1e-13, atol=1e-15)
    assert np.allclose(TO.merit(a), TO.merit(b), rtol=1e-13)
    for p in (a, b):
        TO.ilqr_step(p, 3); TO.al_update(p); TO.ilqr_step(p, 2)
    assert np.allclose(TO.merit(a), TO.merit(b), rtol=1e-9) and np.allclose(TO.controls(a), TO.controls(b), rtol=1e-8, atol=1e-10)
    assert np.allclose(TO.multipliers(a, 2), TO.multipliers(b, 2), rtol=1e-8, atol=1e-10)
    with pytest.raises(TO.ArgumentError):
        TO.AutodiffConstraint(n, m, lambda x: [x[0]] * 17, TO.Inequality(), "state")      # at most 16 rows per general constraint
