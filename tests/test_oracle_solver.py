"""CPU tests of the solver-side rows of the hot path, whose arithmetic lives in Altro.jl and is
NOT under /root/reference ("parity unpinned"): the oracle's Riccati backward pass, forward line search and AL
update are checked against independent mathematics (dense KKT solve) and against the end-to-end numbers the
reference's notebooks recorded (soft pins: same local optimum, not bitwise)."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem

P = TO.problems


def test_lq_problem_one_ilqr_step_equals_kkt_solution():
    """Double integrator (linear) + quadratic cost: one iLQR iteration with alpha = 1 is the exact minimiser.
    Compare with a dense solve of the equality-constrained QP (numpy), independent of the Riccati recursion."""
    prob = P.double_integrator(B=1, N=21, dim=2, cls=OracleProblem, constrained=False)
    n, m, N = prob.n, prob.m, prob.N
    TO.rollout(prob)
    TO.expand(prob)
    AB = TO.dynamics_jacobians(prob)[0]
    A, Bm = AB[0, :, :n], AB[0, :, n:]
    TO.ilqr_step(prob, 1)
    st = TO.solver_state(prob)
    assert st["alpha"][0] == 1.0 and st["bp_status"][0] == 0
    X, U = TO.states(prob)[0], TO.controls(prob)[0]
    # dense QP: min sum 1/2 (x-xf)'Q(x-xf) + 1/2 u'Ru + terminal, s.t. x_{k+1} = A x_k + B u_k, x_0 = 0 -> eliminate x
    xf = prob.xf
    Q, R, Qf = np.eye(n), np.eye(m), np.eye(n) * (N - 1)
    # x_k = sum_j A^{k-1-j} B u_j
    G = np.zeros((N * n, (N - 1) * m))
    for k in range(1, N):
        for j in range(k):
            G[k * n:(k + 1) * n, j * m:(j + 1) * m] = np.linalg.matrix_power(A, k - 1 - j) @ Bm
    Qbar = np.kron(np.eye(N), Q); Qbar[-n:, -n:] = Qf
    Rbar = np.kron(np.eye(N - 1), R)
    xref = np.tile(xf, N)
    Hm = G.T @ Qbar @ G + Rbar
    g = G.T @ Qbar @ xref
    Ustar = np.linalg.solve(Hm, g).reshape(N - 1, m)
    assert np.allclose(U, Ustar, rtol=1e-9, atol=1e-10)
    assert np.allclose(X.reshape(-1), G @ Ustar.reshape(-1), rtol=1e-9, atol=1e-10)
    # a second iteration cannot improve: expected decrease ~ 0
    J1 = TO.merit(prob)[0]
    TO.ilqr_step(prob, 1)
    assert abs(TO.merit(prob)[0] - J1) < 1e-9 * max(1.0, abs(J1))


def test_backward_pass_gains_satisfy_riccati_identities():
    """K, d, dV of one backward pass against a direct numpy evaluation of the recursion (different code path,
    explicit matrix inverses) on the cartpole."""
    prob = P.cartpole(B=1, N=31, cls=OracleProblem)
    n, m, N = prob.n, prob.m, prob.N
    TO.rollout(prob); TO.expand(prob)
    status = TO.backward(prob)
    assert status[0] == 0
    Kg, dg = TO.gains(prob)
    AB = TO.dynamics_jacobians(prob)[0]
    g, H = TO.al_expansion(prob)
    S, s = H[0, -1, :n, :n].copy(), g[0, -1, :n].copy()
    dV = np.zeros(2)
    for k in range(N - 2, -1, -1):
        A, Bm = AB[k, :, :n], AB[k, :, n:]
        lxx, luu, lux = H[0, k, :n, :n], H[0, k, n:, n:], H[0, k, n:, :n]
        Qxx, Quu, Qux = lxx + A.T @ S @ A, luu + Bm.T @ S @ Bm, lux + Bm.T @ S @ A
        Qx, Qu = g[0, k, :n] + A.T @ s, g[0, k, n:] + Bm.T @ s
        Kk, dk = -np.linalg.solve(Quu, Qux), -np.linalg.solve(Quu, Qu)
        assert np.allclose(Kg[0, k], Kk, rtol=1e-8, atol=1e-10) and np.allclose(dg[0, k], dk, rtol=1e-8, atol=1e-10)
        s = Qx + Kk.T @ Quu @ dk + Kk.T @ Qu + Qux.T @ dk
        S = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
        S = 0.5 * (S + S.T)
        dV += [dk @ Qu, 0.5 * dk @ Quu @ dk]
    assert np.allclose(TO.solver_state(prob)["dV"][0], dV, rtol=1e-8)


def _notebook_cartpole(**kw):
    """examples/Cartpole.ipynb's problem as the versions that produced its recorded outputs saw it: TrajectoryOptimization v0.3 /
    RobotDynamics 0.2 / Altro 0.3 integrated the stage costs with dt (v0.7.1, the reference, sums them unscaled, src/objective.jl:104-106)
    and discretised with RK3 by default (the reference: RK4, src/problem.jl:119-123; the integrator is a constructor argument there).
    RK3 exists in the ORACLE only, for exactly this purpose."""
    prob = P.cartpole(B=1, N=101, cls=OracleProblem, dt_scaled_cost=True, **kw)
    assert prob._lib.orc_set_integrator(prob._h, 3) == 0
    TO.rollout(prob)
    return prob


def test_cartpole_rollout_and_stage_costs_match_the_notebooks_ipopt_log():
    """HARD PIN of Cartpole dynamics + RK3 rollout + stage-cost sum: Ipopt's iteration 0 in examples/Cartpole.ipynb (cell 29, "0  2.4696994e-01")
    is the objective of the rolled-out initial guess (U0 = 0.01) with x_N fixed at xf by the goal bound (remove_bounds=true), i.e. the
    sum of the dt-integrated stage costs of knots 1..N-1.  All eight printed digits are reproduced (RK4 would give 0.24697002)."""
    prob = _notebook_cartpole(u_bound=3.0, goal=True)
    Jk = TO.cost_knots(prob)[0]
    assert abs(Jk[:-1].sum() - 0.24696994) < 5e-9
    prob.close()


def test_cartpole_ilqr_reproduces_altros_recorded_solve():
    """HARD PIN of the solver path of the oracle (dual-number Jacobians, Riccati backward pass, regularisation, forward pass + line search):
    Altro's iLQRSolver on the unconstrained cartpole swing-up recorded 84 iterations, terminal cost 1.4497436179031664 and terminal
    dJ 6.889787558717053e-5 (examples/Cartpole.ipynb:378-382, cost_tolerance 1e-4).  The oracle, iterated with the same stopping rule
    (0 <= dJ < cost_tolerance after an accepted step), stops at the SAME iteration with the same cost to 1e-9 -- 84 closed-loop
    iterations amplify any difference in a gain, a step size or a regularisation decision far beyond that."""
    prob = _notebook_cartpole()
    prev = TO.merit(prob)[0]
    for it in range(1, 301):
        TO.ilqr_step(prob, 1)
        J = TO.merit(prob)[0]
        dJ = prev - J
        assert dJ >= -1e-12                      # the line search never accepts an increase
        if 0.0 <= dJ < 1e-4 and TO.solver_state(prob)["alpha"][0] > 0:
            break
        prev = J
    assert it == 84
    assert abs(TO.cost(prob)[0] - 1.4497436179031664) < 1e-9
    assert abs(dJ - 6.889787558717053e-5) < 1e-11
    X = TO.states(prob)[0]
    assert np.allclose(X[-1], [0, np.pi, 0, 0], atol=5e-2)
    prob.close()


def test_cartpole_ilqr_rk4_same_optimum():
    """the reference's own integrator (RK4) on the same problem: same local optimum (the discretisation error of RK3 vs RK4 at dt = 0.05)"""
    prob = P.cartpole(B=1, N=101, cls=OracleProblem, dt_scaled_cost=True)
    TO.rollout(prob)
    prev = TO.merit(prob)[0]
    for it in range(1, 301):
        TO.ilqr_step(prob, 1)
        J = TO.merit(prob)[0]
        if 0.0 <= prev - J < 1e-4 and TO.solver_state(prob)["alpha"][0] > 0:
            break
        prev = J
    assert abs(J - 1.4497436179031664) < 2e-2 and 40 <= it <= 150     # the 1e-4 stopping rule fires on a different plateau of the slow tail
    prob.close()


def test_cartpole_altro_style_al_converges_to_notebook_cost():
    """Cartpole with |u| <= 3 and a goal constraint, AL outer loop: ALTRO recorded cost 1.552558743680986 and
    violation 3.4e-9 (examples/Cartpole.ipynb:216-223).  Soft pin: cost to 2e-2, violation < 1e-4 (no
    projected-Newton polish here)."""
    prob = _notebook_cartpole(u_bound=3.0, goal=True)
    for outer in range(12):
        prev = TO.merit(prob)[0]
        for it in range(100):
            TO.ilqr_step(prob, 1)
            J = TO.merit(prob)[0]
            if abs(prev - J) < 1e-7 and TO.solver_state(prob)["alpha"][0] > 0:   # a failed line search only raises rho
                break
            prev = J
        if TO.max_violation(prob)[0] < 1e-6:
            break
        TO.al_update(prob)
    assert TO.max_violation(prob)[0] < 1e-4
    # ALTRO stopped at cost_tolerance_intermediate = 1e-2 per outer loop (1.5526); Ipopt's optimum of the same problem
    # is 1.4959 (examples/Cartpole.ipynb:780).  The converged AL-iLQR optimum must sit in that neighbourhood.
    assert 1.45 < TO.cost(prob)[0] < 1.552558743680986 + 1e-3
    U = TO.controls(prob)[0]
    assert np.abs(U).max() <= 3.0 + 1e-4


def test_quadrotor_constrained_iterations_decrease_merit_and_respect_status():
    prob = P.quadrotor(B=3, N=31, cls=OracleProblem)
    TO.rollout(prob)
    J0 = TO.merit(prob)
    for _ in range(5):
        TO.ilqr_step(prob, 1)
    st = TO.solver_state(prob)
    assert np.all(st["bp_status"] >= 0)
    assert np.all(TO.merit(prob) < J0)
    TO.al_update(prob)
    lam = TO.multipliers(prob, 0)
    assert np.all(lam <= 0.0)     # inequality multipliers live in the dual cone (negative orthant), src/cones.jl:67
    assert TO.penalty(prob, 0) == 10.0


def test_regularisation_restart_on_indefinite_quu():
    """A cost with negative R makes Quu indefinite: the backward pass must raise rho and restart, not fail."""
    n, m, N = 4, 1, 11
    stage = TO.DiagonalCost(np.ones(n), -0.5 * np.ones(m))
    term = TO.DiagonalCost(np.ones(n), -0.5 * np.ones(m), terminal=True)
    prob = OracleProblem(TO.Cartpole(), TO.Objective(stage, term, N), np.array([0, 0.1, 0, 0]), 0.5)
    TO.rollout(prob); TO.expand(prob)
    status = TO.backward(prob)
    assert status[0] > 0
    assert TO.solver_state(prob)["rho"][0] > 0


def test_mpc_update_trajectory_and_shift():
    """f3: update_trajectory!(obj, Z, start) (src/objective.jl:198-212) == a fresh TrackingObjective on the shifted window up to the
    constant terms set_LQR_goal! leaves untouched (src/cost_functions.jl:238-254); shift_trajectory moves X, U, multipliers, x0, t0."""
    r = np.random.default_rng(5)
    n, m, N, nref = 4, 1, 11, 30
    Xref, Uref = r.standard_normal((nref, n)), r.standard_normal((nref, m))
    Q, R, Qf = np.array([1.0, 2.0, 0.5, 0.1]), np.array([0.3]), np.array([10.0, 10.0, 1.0, 1.0])
    x0 = r.standard_normal((3, n))
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-2.0, u_max=2.0), (1, N - 1))
    TO.add_constraint(cons, TO.GoalConstraint(np.zeros(n)), N)
    prob = OracleProblem(TO.Cartpole(), TO.TrackingObjective(Q, R, Xref[:N], Uref[:N - 1], Qf=Qf), x0, 1.0, constraints=cons)
    U0 = 0.2 * r.standard_normal((3, N - 1, m))
    TO.initial_controls(prob, U0); TO.rollout(prob)
    start = 7
    TO.update_trajectory(prob, Xref, Uref, start)
    fresh = OracleProblem(TO.Cartpole(), TO.TrackingObjective(Q, R, Xref[start - 1:start - 1 + N], Uref[start - 1:start - 2 + N], Qf=Qf), x0, 1.0)
    TO.initial_controls(fresh, U0); TO.rollout(fresh)
    g1, g2 = TO.cost_gradient(prob), TO.cost_gradient(fresh)
    assert np.allclose(g1[:, :N - 1], g2[:, :N - 1], rtol=1e-13, atol=1e-13)
    assert np.allclose(g1[:, N - 1, :n], g2[:, N - 1, :n], rtol=1e-13, atol=1e-13)      # the terminal cost has no control part
    # c is left as is: J differs from the fresh objective by the constants of the old and new references
    half = lambda X, W: 0.5 * np.sum(X * X * W, axis=-1)
    c_old = half(Xref[:N - 1], Q).sum() + half(Uref[:N - 1], R).sum() + half(Xref[N - 1], Qf)
    c_new = half(Xref[start - 1:start - 2 + N], Q).sum() + half(Uref[start - 1:start - 2 + N], R).sum() + half(Xref[start - 2 + N], Qf)
    assert np.allclose(TO.cost(prob) - c_old, TO.cost(fresh) - c_new, rtol=1e-12)
    with pytest.raises(TO.DimensionMismatch):
        TO.update_trajectory(prob, Xref, Uref, nref - N + 2)
    # shift: solve a little so multipliers are non-trivial, then shift by 2 knots
    TO.ilqr_step(prob, 2); TO.al_update(prob)
    X, U, t = TO.states(prob), TO.controls(prob), TO.gettimes(prob)
    lam = TO.multipliers(prob, 0)
    TO.shift_trajectory(prob, 2)
    Xs, Us = TO.states(prob), TO.controls(prob)
    assert np.array_equal(Xs[:, :N - 2], X[:, 2:]) and np.array_equal(Xs[:, N - 2:], np.repeat(X[:, -1:], 2, axis=1))
    assert np.array_equal(Us[:, :N - 3], U[:, 2:]) and np.array_equal(Us[:, N - 3:], np.repeat(U[:, -1:], 2, axis=1))
    ls = TO.multipliers(prob, 0)
    assert np.array_equal(ls[:, :N - 3], lam[:, 2:]) and np.array_equal(ls[:, N - 3:], lam[:, N - 3:])
    assert np.allclose(TO.gettimes(prob), t + (t[2] - t[0]))
    TO.rollout(prob)                                   # x0 <- X[2]: the re-rolled trajectory starts there
    assert np.array_equal(TO.states(prob)[:, 0], X[:, 2])
