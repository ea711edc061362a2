"""CPU tests of the Lie-group rows of the oracle: DiagonalQuatCost / QuatLQRCost (src/lie_costs.jl), QuatVecEq
(src/constraints.jl:938-965, pinned by test/constraint_tests.jl:412-444) and the error-state machinery Altro.jl builds on
RobotDynamics' LieState (state_diff, the state-difference Jacobian G, error_expansion! of dynamics and costs -- arithmetic outside
/root/reference: "parity unpinned", checked here against finite differences of the group operation and a numpy recursion)."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem

P = TO.problems
rng = np.random.default_rng(7)


def qmul(a, b):
    w1, v1, w2, v2 = a[0], a[1:], b[0], b[1:]
    return np.concatenate([[w1 * w2 - v1 @ v2], w1 * v2 + w2 * v1 + np.cross(v1, v2)])


def cayley(phi):
    return np.concatenate([[1.0], phi]) / np.sqrt(1.0 + phi @ phi)


def oplus(x, dx):
    """x (+) dx: vector parts add, attitude <- q (x) cayley(phi)  (the retraction Altro's error state linearises)."""
    y = x.copy()
    y[:3] += dx[:3]; y[7:] += dx[6:]
    y[3:7] = qmul(x[3:7], cayley(dx[3:6]))
    return y


def rand_state(unit=True):
    x = rng.standard_normal(13)
    if unit:
        x[3:7] /= np.linalg.norm(x[3:7])
    return x


def quat_problem(B=2, N=6, **kw):
    return P.quadrotor_lie(B=B, N=N, cls=OracleProblem, **kw)


# ---- DiagonalQuatCost ---------------------------------------------------------------------------------------------------------

def test_diagonal_quat_cost_value_gradient_hessian_closed_forms():
    """RD.evaluate / gradient! of DiagonalQuatCost (src/lie_costs.jl:68-95; the closed forms of the stale test/quatcosts.jl:70-105):
    J = quadratic + w min(1 + dq, 1 - dq), grad_q = -+ w q_ref by the sign of dq = q_ref'q, Hessian = the diagonal Q, R."""
    n, m, N = 13, 4, 4
    Q, R, q, r = rng.random(n), rng.random(m), rng.standard_normal(n), rng.standard_normal(m)
    qref = rng.standard_normal(4); qref /= np.linalg.norm(qref)
    c, w = 0.7, 1.9
    cost = TO.DiagonalQuatCost(Q, R, q=q, r=r, c=c, w=w, q_ref=qref)
    prob = OracleProblem(TO.Quadrotor(), TO.Objective(cost, N), np.zeros(n), 1.0, batch=3)
    X, U = rng.standard_normal((3, N, n)), rng.standard_normal((3, N - 1, m))
    X[1, :, 3:7] = -qref + 0.1 * rng.standard_normal((N, 4))     # dq < 0 branch
    X[2, :, 3:7] = qref + 0.1 * rng.standard_normal((N, 4))      # dq > 0 branch
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    Jk, g, H = TO.cost_knots(prob), TO.cost_gradient(prob), TO.cost_hessian(prob)
    dq = X[..., 3:7] @ qref
    Jx = 0.5 * np.sum(Q * X * X, -1) + X @ q + c + w * np.minimum(1 + dq, 1 - dq)
    Ju = 0.5 * np.sum(R * U * U, -1) + U @ r
    assert np.allclose(Jk[:, :-1], Jx[:, :-1] + Ju, rtol=1e-13) and np.allclose(Jk[:, -1], Jx[:, -1], rtol=1e-13)
    gx = Q * X + q
    gx[..., 3:7] += np.where(dq[..., None] < 0, w, -w) * qref
    assert np.allclose(g[..., :n], gx, rtol=1e-13)
    assert np.allclose(g[:, :-1, n:], R * U + r, rtol=1e-13)
    assert np.array_equal(H[0, 0], np.diag(np.concatenate([Q, R])))
    assert (dq[1] < 0).all() and (dq[2] > 0).all()


def test_quat_lqr_cost_parameters_and_sum():
    """QuatLQRCost (src/lie_costs.jl:129-139): q = -Q xf, r = -R uf, c = 1/2 xf'Q xf + 1/2 uf'R uf, q_ref = xf[quat_ind];
    +(DiagonalQuatCost, QuadraticCostFunction) keeps w, q_ref, q_ind (:152-163)."""
    n, m = 13, 4
    Q, R, xf, uf = rng.random(n), rng.random(m), rng.standard_normal(n), rng.standard_normal(m)
    c = TO.QuatLQRCost(Q, R, xf, uf, w=10.0)
    assert np.allclose(c.q, -Q * xf) and np.allclose(c.r, -R * uf) and np.isclose(c.c, 0.5 * xf @ (Q * xf) + 0.5 * uf @ (R * uf))
    assert np.array_equal(c.q_ref, xf[3:7]) and np.array_equal(c.q_ind, [4, 5, 6, 7]) and c.w == 10.0
    c2 = TO.QuatLQRCost(Q, R, xf, w=2.0, quat_ind=[1, 2, 3, 4])
    assert np.array_equal(c2.q_ref, xf[:4]) and np.allclose(c2.r, 0)
    with pytest.raises(TO.DimensionMismatch):
        TO.QuatLQRCost(Q, R, xf, quat_ind=[4, 5, 6])           # @assert length(quat_ind) == 4
    s = c + TO.DiagonalCost(np.ones(n), np.ones(m), q=np.ones(n))
    assert isinstance(s, TO.DiagonalQuatCost) and s.w == 10.0 and np.allclose(np.diag(s.Q), Q + 1) and np.allclose(s.q, c.q + 1)
    s2 = TO.DiagonalCost(np.ones(n), np.ones(m)) + c
    assert isinstance(s2, TO.DiagonalQuatCost) and np.allclose(np.diag(s2.R), R + 1)
    assert isinstance(c.copy(), TO.DiagonalQuatCost) and c.copy().w == c.w


# ---- QuatVecEq ----------------------------------------------------------------------------------------------------------------

def test_quatveceq_matches_reference_test():
    """test/constraint_tests.jl:412-444: qf = expm([1,0,0] 45 deg); evaluate == -(sign(dq) vec(qf) - vec(q)) with q the
    (normalised) attitude, Jacobian == ForwardDiff (central differences here, atol 1e-6 like the reference's FD check)."""
    n, m, N = 13, 4, 3
    ang = np.deg2rad(45)
    qf = np.array([np.cos(ang / 2), np.sin(ang / 2), 0, 0])
    con = TO.QuatVecEq(n, m, qf)
    assert np.array_equal(con.qind, [4, 5, 6, 7]) and TO.sense(con) == TO.Equality() and con.p == 3
    con2 = TO.QuatVecEq(n, m, qf, [1, 2, 3, 4])
    assert np.array_equal(con2.qind, [1, 2, 3, 4])
    cons = TO.ConstraintList(n, m, N)
    TO.add_constraint(cons, con, (1, N))
    obj = TO.LQRObjective(np.ones(n), np.ones(m), np.ones(n), np.zeros(n), N)
    B = 6
    prob = OracleProblem(TO.Quadrotor(), obj, np.zeros(n), 1.0, constraints=cons, batch=B)
    X = rng.standard_normal((B, N, n))
    X[:3, :, 3:7] /= np.linalg.norm(X[:3, :, 3:7], axis=-1, keepdims=True)      # rand(model): unit attitude; the rest: non-unit inputs
    TO.initial_states(prob, X); TO.initial_controls(prob, rng.standard_normal((B, N - 1, m)))
    c = TO.evaluate_constraints(prob, con)
    q = X[..., 3:7] / np.linalg.norm(X[..., 3:7], axis=-1, keepdims=True)
    sg = np.sign(q @ qf)
    assert np.allclose(c, -(sg[..., None] * qf[1:] - q[..., 1:]), rtol=1e-13, atol=1e-15)
    assert (sg < 0).any() and (sg > 0).any()
    J = TO.constraint_jacobians(prob, con)
    assert J.shape == (B, N, 3, n + m)

    def f(x):
        qq = x[3:7] / np.linalg.norm(x[3:7]); s = -1.0 if qq @ qf < 0 else 1.0
        return -(s * qf[1:] - qq[1:])
    for b in range(B):
        Jfd = np.zeros((3, n + m))
        for j in range(n):
            e = np.zeros(n); e[j] = 1e-6
            Jfd[:, j] = (f(X[b, 1] + e) - f(X[b, 1] - e)) / 2e-6
        assert np.allclose(J[b, 1], Jfd, atol=1e-6)


# ---- error state ----------------------------------------------------------------------------------------------------------------

def test_state_diff_inverts_the_retraction_and_G_is_its_derivative():
    prob = quat_problem(B=4, N=5)
    assert TO.errstate_dim(prob) == 12 and prob.model.errstate_dim() == 12 and TO.Cartpole().errstate_dim() == 4
    X = np.stack([[rand_state() for _ in range(5)] for _ in range(4)])
    TO.initial_states(prob, X)
    dx = 0.3 * rng.standard_normal((4, 5, 12))
    Xbar = np.stack([[oplus(X[b, k], dx[b, k]) for k in range(5)] for b in range(4)])
    assert np.allclose(TO.state_diff(prob, Xbar), dx, rtol=1e-12, atol=1e-14)          # state_diff(x (+) d, x) == d
    assert np.allclose(TO.state_diff(prob, X), 0, atol=1e-15)
    # vector-space problem: plain difference
    cp = P.cartpole(B=2, N=5, cls=OracleProblem)
    TO.rollout(cp)
    Xc = TO.states(cp)
    assert np.allclose(TO.state_diff(cp, Xc + 0.5), 0.5)
    with pytest.raises(TO.ArgumentError):
        OracleProblem(TO.Cartpole(), cp.obj, np.zeros(4), 1.0, error_state=True)


def _fd_jac(fun, nin, h=1e-6):
    cols = []
    for j in range(nin):
        e = np.zeros(nin); e[j] = h
        cols.append((fun(e) - fun(-e)) / (2 * h))
    return np.stack(cols, axis=1)


def test_error_dynamics_equal_the_derivative_of_the_error_map():
    """[A_e B_e] = G_{k+1}' [A G_k | B] must be the derivative of (d, du) -> state_diff(f(x (+) d, u + du), f(x, u)) at 0."""
    from oracle_binding import oracle_discrete_dynamics
    prob = quat_problem(B=2, N=4)
    TO.rollout(prob); TO.expand(prob)
    ABe = TO.error_dynamics(prob)
    assert ABe.shape == (2, 3, 12, 16)
    X, U = TO.states(prob), TO.controls(prob)
    model, h = prob.model, 0.05
    chk = P.quadrotor_lie(B=1, N=2, cls=OracleProblem)          # used for its state_diff only
    for b, k in ((0, 0), (1, 2)):
        x, u = X[b, k], U[b, k]
        xn = oracle_discrete_dynamics(model, x, u, h)
        assert np.allclose(xn, X[b, k + 1], rtol=1e-13, atol=1e-13)

        def fun(e):
            xb = oracle_discrete_dynamics(model, oplus(x, e[:12]), u + e[12:], h)
            TO.initial_states(chk, np.stack([xn, xn])[None]); 
            return TO.state_diff(chk, np.stack([xb, xb])[None])[0, 0]
        J = _fd_jac(fun, 16)
        assert np.allclose(ABe[b, k], J, rtol=1e-6, atol=2e-7)


def test_error_expansion_is_the_second_order_model_in_error_coordinates():
    """E.x = G'g, E.xx = G'HG - (q'g_q) I3 (Altro error_expansion! + Rotations grad^2-differential): for a quadratic cost in x the
    function d -> l(x (+) d, u) has exactly this gradient and Hessian at d = 0 (central differences)."""
    prob = quat_problem(B=2, N=4, quat_goal=False)
    TO.rollout(prob)
    X, U = TO.states(prob), TO.controls(prob)
    ge, He = TO.error_expansion(prob)
    assert ge.shape == (2, 4, 16) and He.shape == (2, 4, 16, 16)
    assert np.allclose(He, np.swapaxes(He, -1, -2), atol=1e-12)
    cost = prob.obj[0]
    Q, R = np.diag(cost.Q), np.diag(cost.R)

    def l(x, u):
        dq = cost.q_ref @ x[3:7]
        return 0.5 * x @ (Q * x) + cost.q @ x + 0.5 * u @ (R * u) + cost.r @ u + cost.w * min(1 + dq, 1 - dq)
    b, k = 1, 1                                   # a stage knot without active bounds (multipliers are zero, mu c inactive)
    x, u = X[b, k], U[b, k]
    assert (u > 0).all() and (u < 10).all()
    h = 1e-4
    g_fd = np.array([(l(oplus(x, e[:12]), u + e[12:]) - l(oplus(x, -e[:12]), u - e[12:])) / (2 * h) for e in h * np.eye(16)])
    assert np.allclose(ge[b, k], g_fd, rtol=1e-6, atol=1e-8)
    H_fd = np.zeros((16, 16))
    E = h * np.eye(16)
    f0 = l(x, u)
    for i in range(16):
        for j in range(16):
            fpp = l(oplus(x, (E[i] + E[j])[:12]), u + (E[i] + E[j])[12:]); fmm = l(oplus(x, -(E[i] + E[j])[:12]), u - (E[i] + E[j])[12:])
            fpm = l(oplus(x, (E[i] - E[j])[:12]), u + (E[i] - E[j])[12:]); fmp = l(oplus(x, (E[j] - E[i])[:12]), u + (E[j] - E[i])[12:])
            H_fd[i, j] = (fpp + fmm - fpm - fmp) / (4 * h * h)
    assert np.allclose(He[b, k], H_fd, rtol=1e-4, atol=2e-6)
    assert f0 == f0


def test_error_state_backward_pass_against_numpy_recursion_and_forward_feedback():
    prob = quat_problem(B=2, N=9)
    TO.rollout(prob); TO.expand(prob)
    assert (TO.backward(prob) == 0).all()
    Kg, dg = TO.gains(prob)
    assert Kg.shape == (2, 8, 4, 12)
    AB = TO.error_dynamics(prob)
    g, H = TO.error_expansion(prob)
    n = 12
    for b in range(2):
        S, s = H[b, -1, :n, :n].copy(), g[b, -1, :n].copy()
        dV = np.zeros(2)
        for k in range(7, -1, -1):
            A, Bm = AB[b, k, :, :n], AB[b, k, :, n:]
            Qxx, Quu, Qux = H[b, k, :n, :n] + A.T @ S @ A, H[b, k, n:, n:] + Bm.T @ S @ Bm, H[b, k, n:, :n] + Bm.T @ S @ A
            Qx, Qu = g[b, k, :n] + A.T @ s, g[b, k, n:] + Bm.T @ s
            Kk, dk = -np.linalg.solve(Quu, Qux), -np.linalg.solve(Quu, Qu)
            assert np.allclose(Kg[b, k], Kk, rtol=1e-8, atol=1e-10) and np.allclose(dg[b, k], dk, rtol=1e-8, atol=1e-10)
            s = Qx + Kk.T @ Quu @ dk + Kk.T @ Qu + Qux.T @ dk
            S = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
            S = 0.5 * (S + S.T)
            dV += [dk @ Qu, 0.5 * dk @ Quu @ dk]
        assert np.allclose(TO.solver_state(prob)["dV"][b], dV, rtol=1e-8)
    # forward pass: ubar_k = u_k + K_k state_diff(xbar_k, x_k) + alpha d_k on the accepted trajectory
    X, U = TO.states(prob), TO.controls(prob)
    J0 = TO.merit(prob)
    J, alpha = TO.forward(prob)
    assert (alpha > 0).all() and (J < J0).all()
    Xn, Un = TO.states(prob), TO.controls(prob)
    chk = P.quadrotor_lie(B=2, N=9, cls=OracleProblem)
    TO.initial_states(chk, X)
    dx = TO.state_diff(chk, Xn)
    for b in range(2):
        for k in range(8):
            assert np.allclose(Un[b, k], U[b, k] + Kg[b, k] @ dx[b, k] + alpha[b] * dg[b, k], rtol=1e-10, atol=1e-12)


def test_error_state_al_ilqr_converges_with_attitude_constraint():
    """QuatLQRCost + control bounds + position/velocity goal + QuatVecEq attitude goal, from random initial attitudes."""
    prob = quat_problem(B=4, N=41)
    TO.rollout(prob)
    J0 = TO.merit(prob)
    for outer in range(8):
        for it in range(40):
            TO.ilqr_step(prob, 1)
        if TO.max_violation(prob).max() < 1e-5:
            break
        TO.al_update(prob)
    assert (TO.merit(prob) < J0).all()
    assert TO.max_violation(prob).max() < 1e-4
    X = TO.states(prob)
    q = X[:, -1, 3:7] / np.linalg.norm(X[:, -1, 3:7], axis=-1, keepdims=True)
    assert np.allclose(np.abs(q[:, 0]), 1.0, atol=1e-4) and np.allclose(X[:, -1, :3], [0, 0, 2], atol=1e-4)
    assert np.allclose(np.linalg.norm(X[..., 3:7], axis=-1), 1.0, atol=5e-3)     # RK4 keeps the attitude near the unit sphere


def test_quadrotor_notebook_zigzag_soft_pin():
    """examples/Quadrotor.ipynb:374-376: ALTRO (Altro 0.3, RK3, projected-Newton polish) recorded cost 0.2992834848449584,
    violation 7.6e-10, 90 iterations.  The error-state AL-iLQR of this oracle (RK4, no polish) converges to the same
    zig-zag optimum within a few percent; WITHOUT the error state iLQR stalls an order of magnitude above it."""
    prob = P.quadrotor_zigzag(cls=OracleProblem)
    TO.set_options(prob, penalty_scaling=100.0, penalty_initial=0.1)        # the notebook's SolverOptions (:366-369)
    TO.rollout(prob)
    for outer in range(8):
        for it in range(100):
            J0 = TO.merit(prob)[0]
            TO.ilqr_step(prob, 1)
            if TO.solver_state(prob)["alpha"][0] > 0 and abs(J0 - TO.merit(prob)[0]) < 1e-6:
                break
        if TO.max_violation(prob)[0] < 1e-6:
            break
        TO.al_update(prob)
    assert TO.max_violation(prob)[0] < 1e-5
    J = TO.cost(prob)[0]
    assert abs(J - 0.2992834848449584) < 0.05 * 0.2992834848449584, J
    X = TO.states(prob)[0]
    assert np.linalg.norm(X[32, :3] - [10, 0, 1]) < 1.0 and np.linalg.norm(X[65, :3] - [-10, 0, 1]) < 1.0     # through the waypoints
    assert np.linalg.norm(X[-1, :3] - [0, 10, 1]) < 0.2


def test_error_quadratic_cost():
    """ErrorQuadratic (src/lie_costs.jl:170-240): 1/2 dx'Q dx + c + 1/2 u'Ru + r'u with dx = state_diff(x, x_ref) (Cayley map)"""
    n, m, N, B = 13, 4, 4, 3
    model = TO.Quadrotor()
    Q13, R = rng.random(n), rng.random(m)
    x_ref, u_ref = rand_state(), rng.standard_normal(m)
    cost = TO.ErrorQuadratic(model, Q13, R, x_ref, u_ref)
    assert len(cost.prog) <= 128
    Q12 = np.delete(Q13, 3)                                   # deleteat(Q.diag, 4)  :216
    prob = OracleProblem(model, TO.Objective(cost, N), np.zeros(n), 1.0, batch=B)
    X = np.stack([[rand_state() for _ in range(N)] for _ in range(B)])
    U = rng.standard_normal((B, N - 1, m))
    TO.initial_states(prob, X); TO.initial_controls(prob, U)
    chk = P.quadrotor_lie(B=B, N=N, cls=OracleProblem)
    TO.initial_states(chk, np.broadcast_to(x_ref, (B, N, n)))
    dx = TO.state_diff(chk, X)                                 # state_diff(x, x_ref): x relative to the reference
    Jk = TO.cost_knots(prob)
    Jx = 0.5 * np.sum(Q12 * dx * dx, axis=-1)
    Ju = 0.5 * np.sum(R * (U - u_ref) ** 2, axis=-1)          # 1/2 u'Ru - (R u_ref)'u + 1/2 u_ref'R u_ref
    assert np.allclose(Jk[:, :-1], Jx[:, :-1] + Ju, rtol=1e-12)
    assert np.allclose(Jk[:, -1], Jx[:, -1] + 0.5 * u_ref @ (R * u_ref), rtol=1e-12)      # terminal knot: u = 0
    # gradient against central differences of the closed form
    g = TO.cost_gradient(prob)
    b, k = 1, 1

    def f(z):
        TO.initial_states(chk, np.broadcast_to(x_ref, (B, N, n)))
        d = TO.state_diff(chk, np.broadcast_to(z[:n], (B, N, n)))[0, 0]
        return 0.5 * d @ (Q12 * d) + 0.5 * np.sum(R * (z[n:] - u_ref) ** 2)
    z = np.concatenate([X[b, k], U[b, k]])
    h = 1e-6
    gfd = np.array([(f(z + e) - f(z - e)) / (2 * h) for e in h * np.eye(n + m)])
    assert np.allclose(g[b, k], gfd, rtol=1e-6, atol=1e-7)
    with pytest.raises(TO.ArgumentError):
        TO.ErrorQuadratic(TO.Cartpole(), np.ones(4), np.ones(1), np.zeros(4))
    # it drives the error-state solver (user cost -> materialised expansion -> dense Riccati pass)
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    hover = model.hover_control()
    stage = TO.ErrorQuadratic(model, np.full(12, 0.1), np.full(m, 0.01), xf, hover)
    term = TO.ErrorQuadratic(model, np.full(12, 100.0), np.full(m, 0.01), xf, hover, terminal=True)
    base = P.quadrotor_lie(B=2, N=21, cls=OracleProblem)
    p2 = OracleProblem(model, TO.Objective(stage, term, 21), base.x0, 1.0, error_state=True)
    TO.initial_controls(p2, TO.controls(base)); TO.rollout(p2)
    J0 = TO.merit(p2)
    for _ in range(15):
        TO.ilqr_step(p2, 1)
    assert np.all(TO.merit(p2) < 0.2 * J0)


def test_constraint_error_jacobians_follow_the_reference_hook():
    """error_expansion!(jac, jac0, con, model, G, inds) (src/abstract_constraint.jl:282-303; usage in the stale test/internal_api.jl:74-78,
    :120-126): state block = jac0_x G_k, control block copied; G from RD.errstate_jacobian!.  Consistency: the error-state AL expansion
    built by the solver path equals G' (cost + AL) G assembled from these Jacobians for an equality constraint."""
    prob = quat_problem(B=2, N=6)
    TO.rollout(prob)
    G = TO.errstate_jacobian(prob)
    assert G.shape == (2, 6, 13, 12)
    X = TO.states(prob)
    for b, k in ((0, 0), (1, 4)):
        q = X[b, k, 3:7]
        Lq = np.array([[q[0], -q[1], -q[2], -q[3]], [q[1], q[0], -q[3], q[2]], [q[2], q[3], q[0], -q[1]], [q[3], -q[2], q[1], q[0]]])
        assert np.allclose(G[b, k, 3:7, 3:6], Lq[:, 1:]) and np.allclose(G[b, k, :3, :3], np.eye(3)) and np.allclose(G[b, k, 7:, 6:], np.eye(6))
        # derivative of the retraction: d/dd (x (+) d) at 0
        Jfd = _fd_jac(lambda e: oplus(X[b, k], e), 12)
        assert np.allclose(G[b, k], Jfd, atol=1e-8)
    for i, con in enumerate(prob.constraints):
        J0, Je = TO.constraint_jacobians(prob, i), TO.constraint_error_jacobians(prob, i)
        first, last = prob.constraints.inds[i]
        assert Je.shape == J0.shape[:-1] + (16,)
        assert np.allclose(Je[..., :12], J0[..., :13] @ G[:, first - 1:last]) and np.array_equal(Je[..., 12:], J0[..., 13:])
    # QuatVecEq at the terminal knot: its error-state Jacobian is (close to) the identity on the attitude error when q ~ qf
    cp = P.cartpole(B=1, N=4, cls=OracleProblem, goal=True)
    TO.rollout(cp)
    assert np.array_equal(TO.constraint_error_jacobians(cp, 0), TO.constraint_jacobians(cp, 0))       # no Lie group: unchanged
