"""GPU parity tests (`-m gpu`): every kernel of the hot path, called through the C ABI, against the CPU oracle
on identical seeded inputs.  Tolerances (fp64, north_star "stated fp64 tolerance"): 1e-10 relative per kernel
(reduction / FMA-contraction order differs), 1e-8 after full iLQR iterations."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem, match_algebra, oracle_grad_projection, oracle_hess_projection, oracle_projection
from parity_util import GAIN_TOL, check, decisions_agree, triple

pytestmark = pytest.mark.gpu
P = TO.problems

KERNEL_RTOL = 1e-10
# Closed-loop iterations amplify the rounding differences of one backward pass.  Where a test compares iterates of the (ill-conditioned)
# Quadrotor problems it uses parity_util.check: per instance, against the divergence of the oracle's own two arithmetic forms.  The global
# tolerance below is kept for the well-conditioned small configurations only.
ITER_RTOL = 1e-6


def close(a, b, rtol, what=""):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    err = float(np.max(np.abs(a - b))) if b.size else 0.0
    assert np.all(np.isfinite(a)) or not np.all(np.isfinite(b)), f"{what}: non-finite GPU result"
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} > {rtol:.0e} * {scale:.3e}"


CONFIGS = {
    "double_integrator_1d": lambda cls: P.double_integrator(B=3, N=51, dim=1, cls=cls),
    "double_integrator_2d": lambda cls: P.double_integrator(B=2, N=21, dim=2, cls=cls),
    "cartpole": lambda cls: P.cartpole(B=37, N=101, cls=cls),
    "cartpole_altro": lambda cls: P.cartpole(B=5, N=101, cls=cls, u_bound=3.0, goal=True, dt_scaled_cost=True),
    "quadrotor": lambda cls: P.quadrotor(B=33, N=101, cls=cls),
    "quadrotor_short": lambda cls: P.quadrotor(B=4, N=4, cls=cls, dt=0.05),
    "quadrotor_dense_cost": lambda cls: P.quadrotor(B=5, N=31, cls=cls, dt=0.05, dense_cost=True),
    "quadrotor_long": lambda cls: P.quadrotor(B=2, N=401, cls=cls, dt=0.05),
    "cartpole_single": lambda cls: P.cartpole(B=1, N=2, cls=cls),
    "acrobot_dense": lambda cls: P.acrobot(B=9, N=201, cls=cls, dense_cost=True),
    "acrobot_diag": lambda cls: P.acrobot(B=3, N=51, cls=cls, dense_cost=False),
    # Lie-group error state (Riccati on n_e = 12, lie.cu), DiagonalQuatCost, QuatVecEq
    "quadrotor_lie": lambda cls: P.quadrotor_lie(B=7, N=31, cls=cls),                                           # quat cost + QuatVecEq + error state
    "quadrotor_lie_lqr": lambda cls: P.quadrotor_lie(B=9, N=41, cls=cls, quat_cost=False, quat_goal=False),     # BASELINE objective on the error state (fast forward path)
    "quadrotor_quatcost_fullstate": lambda cls: P.quadrotor_lie(B=5, N=21, cls=cls, quat_goal=False, error_state=False),
    "quadrotor_quatveceq_fullstate": lambda cls: P.quadrotor_lie(B=4, N=21, cls=cls, quat_cost=False, error_state=False),
}


@pytest.fixture(params=sorted(CONFIGS))
def pair(request):
    g, o = CONFIGS[request.param](TO.Problem), CONFIGS[request.param](OracleProblem)
    match_algebra(g, o)          # the oracle evaluates the backward pass in the arithmetic form of the kernel that will run (parity_util.py)
    yield g, o
    g.close(); o.close()


@pytest.fixture(params=sorted(CONFIGS))
def trio(request):
    g, o, t = triple(CONFIGS[request.param])
    yield g, o, t
    for p in (g, o, t):
        p.close()


def test_rollout_cost_constraints(pair):
    g, o = pair
    for p in pair:
        TO.rollout(p)
    close(TO.states(g), TO.states(o), KERNEL_RTOL, "rollout X")
    close(TO.controls(g), TO.controls(o), 0.0, "controls round trip")
    close(TO.cost_knots(g), TO.cost_knots(o), KERNEL_RTOL, "cost knots")
    close(TO.cost(g), TO.cost(o), KERNEL_RTOL, "cost")
    close(TO.cost_gradient(g), TO.cost_gradient(o), KERNEL_RTOL, "cost gradient")
    close(TO.cost_hessian(g), TO.cost_hessian(o), KERNEL_RTOL, "cost hessian")
    assert np.array_equal(TO.num_constraints(g), TO.num_constraints(o))
    for i in range(len(g.constraints)):
        close(TO.evaluate_constraints(g, i), TO.evaluate_constraints(o, i), KERNEL_RTOL, f"constraint {i} values")
        close(TO.constraint_jacobians(g, i), TO.constraint_jacobians(o, i), KERNEL_RTOL, f"constraint {i} jacobians")
    close(TO.max_violation(g), TO.max_violation(o), KERNEL_RTOL, "max violation")
    close(TO.merit(g), TO.merit(o), KERNEL_RTOL, "merit")
    gg, gh = TO.al_expansion(g)
    og, oh = TO.al_expansion(o)
    close(gg, og, KERNEL_RTOL, "AL gradient"); close(gh, oh, KERNEL_RTOL, "AL hessian")


def test_expansion_backward_forward(trio):
    g, o, t = trio
    for p in trio:
        TO.rollout(p); TO.expand(p)
    close(TO.dynamics_jacobians(g), TO.dynamics_jacobians(o), KERNEL_RTOL, "[A B]")
    sg, so = TO.backward(g), TO.backward(o); TO.backward(t)
    assert np.array_equal(sg, so)
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, GAIN_TOL, "K"); close(dg, do, GAIN_TOL, "d")
    close(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], GAIN_TOL, "dV")
    (Jg, ag), (Jo, ao), (Jt, at) = TO.forward(g), TO.forward(o), TO.forward(t)
    ok = decisions_agree("accepted step sizes", ag, ao, at)
    check("J after forward pass", Jg, Jo, Jt, 1e-10, ok)
    check("X after forward pass", TO.states(g), TO.states(o), TO.states(t), 1e-10, ok)
    check("U after forward pass", TO.controls(g), TO.controls(o), TO.controls(t), 1e-10, ok)
    decisions_agree("line-search trials", TO.solver_state(g)["ls_iters"], TO.solver_state(o)["ls_iters"], TO.solver_state(t)["ls_iters"])


def test_ilqr_iterations_and_al_update(trio):
    g, o, t = trio
    for p in trio:
        TO.rollout(p)
        TO.ilqr_step(p, 3)
    sg, so, st_ = TO.solver_state(g), TO.solver_state(o), TO.solver_state(t)
    # discrete line-search decisions are compared where the expected decrease is not at round-off level (an instance
    # that has converged accepts or rejects a step on the last bits of J) and where the oracle's two arithmetic forms agree
    live = np.abs(so["dV"][:, 0]) > 1e-9 * np.maximum(1.0, np.abs(TO.merit(o)))
    dec = live & (so["alpha"] == st_["alpha"]) & (so["bp_status"] == st_["bp_status"]) & (sg["alpha"] == so["alpha"]) & (sg["bp_status"] == so["bp_status"])
    check("merit after 3 iterations", TO.merit(g), TO.merit(o), TO.merit(t), 1e-8, dec)
    check("X after 3 iterations", TO.states(g), TO.states(o), TO.states(t), 1e-8, dec)
    check("U after 3 iterations", TO.controls(g), TO.controls(o), TO.controls(t), 1e-8, dec)
    for k in ("alpha", "ls_iters", "bp_status"):
        decisions_agree(k, sg[k], so[k], st_[k], live, allow=0.05)
    check("rho", sg["rho"], so["rho"], st_["rho"], 1e-12, dec)
    if len(g.constraints):
        for p in trio:
            TO.al_update(p)
        for i in range(len(g.constraints)):
            check(f"multipliers {i}", TO.multipliers(g, i), TO.multipliers(o, i), TO.multipliers(t, i), 1e-8, dec)
            assert TO.penalty(g, i) == TO.penalty(o, i)
        for p in trio:
            TO.ilqr_step(p, 2)
        sg, so, st_ = TO.solver_state(g), TO.solver_state(o), TO.solver_state(t)
        dec = dec & (so["alpha"] == st_["alpha"]) & (so["bp_status"] == st_["bp_status"]) & (sg["alpha"] == so["alpha"]) & (sg["bp_status"] == so["bp_status"])
        check("merit after AL update + 2 iterations", TO.merit(g), TO.merit(o), TO.merit(t), 1e-8, dec)
        check("violation", TO.max_violation(g), TO.max_violation(o), TO.max_violation(t), 1e-8, dec)


@pytest.mark.parametrize("name", ["double_integrator_1d", "double_integrator_2d", "cartpole", "cartpole_altro", "acrobot_dense", "acrobot_diag"])
@pytest.mark.parametrize("kernel", [1, 2])
def test_small_model_riccati_kernels(name, kernel):
    """both Riccati kernels of the small models (1 = warp per instance, 2 = thread per instance, riccati_small.cu; the automatic
    choice switches at 2048 instances) against the oracle: gains, expected decrease, iterates, restart bookkeeping"""
    g, o = CONFIGS[name](TO.Problem), CONFIGS[name](OracleProblem)
    TO.set_options(g, backward_kernel=kernel)
    for p in (g, o):
        TO.rollout(p); TO.expand(p)
    assert np.array_equal(TO.backward(g), TO.backward(o))
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, 1e-9, "K"); close(dg, do, 1e-9, "d")
    close(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], 1e-9, "dV")
    for p in (g, o):
        TO.forward(p); TO.ilqr_step(p, 3)
        if len(p.constraints):
            TO.al_update(p); TO.ilqr_step(p, 2)
    # converged instances take their discrete decisions on the last bits of J (see test_ilqr_iterations_and_al_update)
    so = TO.solver_state(o)
    live = np.abs(so["dV"][:, 0]) > 1e-9 * np.maximum(1.0, np.abs(TO.merit(o)))
    close(TO.merit(g)[live], TO.merit(o)[live], 1e-5, "merit")
    g.close(); o.close()


@pytest.mark.parametrize("name", ["quadrotor_lie", "quadrotor_lie_lqr", "quadrotor_quatcost_fullstate"])
def test_error_state_kernels(name):
    """lie.cu: RD.state_diff, [A_e B_e] = G' [A G | B], the error-state cost + AL expansion (Altro error_expansion!) against the oracle"""
    g, o, t = triple(CONFIGS[name])
    assert TO.errstate_dim(g) == TO.errstate_dim(o) == (12 if g.error_state else 13)
    for p in (g, o, t):
        TO.rollout(p); TO.expand(p)
    close(TO.error_dynamics(g), TO.error_dynamics(o), KERNEL_RTOL, "[A_e B_e]")
    r = np.random.default_rng(5)
    Xbar = TO.states(o) + 0.2 * r.standard_normal((g.B, g.N, g.n))
    close(TO.state_diff(g, Xbar), TO.state_diff(o, Xbar), KERNEL_RTOL, "state_diff")
    for it in range(2):
        (ge, He), (oe, Ho), (te, Ht) = TO.error_expansion(g), TO.error_expansion(o), TO.error_expansion(t)
        # second round: the trajectories themselves differ by the first iteration's amplified round-off -> budget from the oracle twin
        check("error-state gradient", ge, oe, te, KERNEL_RTOL); check("error-state Hessian", He, Ho, Ht, KERNEL_RTOL)
        for p in (g, o, t):
            TO.ilqr_step(p, 1); TO.al_update(p)      # non-zero multipliers, active bounds
    for p in (g, o, t):
        p.close()


def test_error_state_riccati_kernel_variants():
    """the backward passes of the error-state path against the oracle in the matching arithmetic form: register-resident fragment kernel
    (automatic choice for diagonal costs + Goal/Bound; block-inverse form), shared-memory tensor-core kernel on the compact expansion (5)
    and on the full materialised expansion (QuatVecEq present), generic DFMA kernel (3) -- all LDL' form"""
    for name, kernel in (("quadrotor_lie_lqr", 0), ("quadrotor_lie_lqr", 5), ("quadrotor_lie_lqr", 3), ("quadrotor_lie", 0), ("quadrotor_lie", 3)):
        g, o, t = triple(CONFIGS[name], {"backward_kernel": kernel})
        assert TO.backward_algebra(g) == (1 if (name, kernel) == ("quadrotor_lie_lqr", 0) else 0)
        for p in (g, o, t):
            TO.rollout(p); TO.expand(p)
        assert np.array_equal(TO.backward(g), TO.backward(o)); TO.backward(t)
        (Kg, dg), (Ko, do) = TO.gains(g), TO.gains(o)
        close(Kg, Ko, GAIN_TOL, f"K {name} kernel {kernel}"); close(dg, do, GAIN_TOL, f"d {name} kernel {kernel}")
        for p in (g, o, t):     # ... and again from iterates that carry multipliers and a regularisation history
            TO.forward(p); TO.ilqr_step(p, 1); TO.al_update(p); TO.expand(p)
        sg, so, st_ = TO.backward(g), TO.backward(o), TO.backward(t)
        same = decisions_agree("restarts", sg, so, st_)
        (Kg, dg), (Ko, do), (Kt, dt) = TO.gains(g), TO.gains(o), TO.gains(t)
        check(f"K {name} kernel {kernel}", Kg, Ko, Kt, GAIN_TOL, same); check(f"d {name} kernel {kernel}", dg, do, dt, GAIN_TOL, same)
        for p in (g, o, t):
            p.close()


def test_fragment_riccati_kernel_regularisation_and_queue():
    """riccati_frag.cu: (a) restarts -- a negative-definite control cost makes Quu + rho I indefinite until rho has grown: same restart
    counts, rho and gains as the oracle; (b) more instances than one wave of resident warps (148 SMs x 28) goes through the work queue"""
    n, m, N = 13, 4, 21
    xf = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    probs = []
    for cls in (TO.Problem, OracleProblem):
        r = np.random.default_rng(4)
        stage = TO.LQRCost(np.full(n, 0.1), np.full(m, -0.05), xf, TO.Quadrotor().hover_control())
        term = TO.LQRCost(np.full(n, 10.0), np.full(m, -0.05), xf, TO.Quadrotor().hover_control(), terminal=True)
        cons = TO.ConstraintList(n, m, N)
        TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=np.zeros(4), u_max=np.full(4, 10.0)), (1, N - 1))
        x0 = np.tile(np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), (6, 1)); x0[:, :3] += r.uniform(-1, 1, (6, 3))
        p = cls(TO.Quadrotor(), TO.Objective(stage, term, N), x0, 1.0, xf=xf, constraints=cons, error_state=True)
        TO.initial_controls(p, TO.Quadrotor().hover_control() + 0.05 * r.standard_normal((6, N - 1, m)))
        TO.rollout(p); TO.expand(p)
        probs.append(p)
    g, o = probs
    match_algebra(g, o)
    assert TO.backward_algebra(g) == 1
    sg, so = TO.backward(g), TO.backward(o)
    assert np.array_equal(sg, so) and np.all(so > 0), (sg, so)
    close(TO.solver_state(g)["rho"], TO.solver_state(o)["rho"], 1e-12, "rho")
    close(TO.gains(g)[0], TO.gains(o)[0], 1e-8, "K after restarts"); close(TO.gains(g)[1], TO.gains(o)[1], 1e-8, "d after restarts")
    g.close(); o.close()
    B = 148 * 28 + 300
    g, o = P.quadrotor(B=B, N=11, dt=0.05, error_state=True), P.quadrotor(B=B, N=11, dt=0.05, error_state=True, cls=OracleProblem)
    match_algebra(g, o)
    for p in (g, o):
        TO.rollout(p); TO.expand(p)
    assert np.array_equal(TO.backward(g), TO.backward(o))
    close(TO.gains(g)[0], TO.gains(o)[0], 1e-9, "K (queue)"); close(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], 1e-9, "dV (queue)")
    g.close(); o.close()


@pytest.mark.parametrize("B,N", [(1, 2), (3, 5), (5, 3), (33, 2), (2, 64)])
def test_error_state_edge_sizes(B, N):
    """the record / fragment path at the smallest and at odd sizes: one knot pair, fewer instances than a CTA has warps, a ring deeper than
    the horizon, the 16-lane groups of the cost expansion ending inside a warp"""
    g, o, t = triple(lambda cls: P.quadrotor(B=B, N=N, dt=0.05, error_state=True, cls=cls))
    for p in (g, o, t):
        TO.rollout(p); TO.expand(p)
    close(TO.error_dynamics(g), TO.error_dynamics(o), KERNEL_RTOL, "[A_e B_e]")
    assert np.array_equal(TO.backward(g), TO.backward(o)); TO.backward(t)
    close(TO.gains(g)[0], TO.gains(o)[0], GAIN_TOL, "K"); close(TO.gains(g)[1], TO.gains(o)[1], GAIN_TOL, "d")
    close(TO.solver_state(g)["dV"], TO.solver_state(o)["dV"], GAIN_TOL, "dV")
    for p in (g, o, t):
        TO.ilqr_step(p, 2); TO.al_update(p); TO.ilqr_step(p, 1)
    check("X after 3 iterations", TO.states(g), TO.states(o), TO.states(t), 1e-9)
    check("merit", TO.merit(g), TO.merit(o), TO.merit(t), 1e-9)
    for p in (g, o, t):
        p.close()


def test_error_state_full_size_properties():
    """BASELINE-size batch on the error state: merit monotone, backward pass never fails, attitude stays on the unit sphere,
    instance 0..7 equal to an 8-instance problem with the same inputs (instances never interact)."""
    B, N = 2048, 101
    prob = P.quadrotor_lie(B=B, N=N, quat_cost=False, quat_goal=False)
    small = P.quadrotor_lie(B=8, N=N, quat_cost=False, quat_goal=False)
    TO.set_initial_state(small, prob.x0[:8]); TO.initial_controls(small, TO.controls(prob)[:8])
    for p in (prob, small):
        TO.rollout(p)
    J0 = TO.merit(prob)
    for p in (prob, small):
        TO.ilqr_step(p, 3)
    J3 = TO.merit(prob)
    # (a handful of the random attitudes make Quu indefinite beyond bp_reg_max in the 2nd / 3rd iteration -- the oracle reports the same
    #  instances; a failed backward pass keeps the trajectory, so the merit stays monotone)
    assert np.all(J3 <= J0 + 1e-9) and np.mean(TO.solver_state(prob)["bp_status"] < 0) < 0.02
    assert np.median(np.abs(np.linalg.norm(TO.states(prob)[..., 3:7], axis=-1) - 1.0)) < 1e-2     # RK4 does not renormalise (nor does the reference)
    np.testing.assert_array_equal(TO.controls(prob)[:8], TO.controls(small))
    prob.close(); small.close()


def _autodiff_pair(kind):
    from test_oracle_nlcost import cartpole_cost, lqr_as_autodiff
    r = np.random.default_rng(21)
    probs = []
    for cls in (TO.Problem, OracleProblem):
        if kind == "cartpole_docs_cost":          # docs/src/costfunction_interface.md:38-50 as stage cost, |u| <= 4 + goal
            n, m, N, B = 4, 1, 41, 6
            stage = TO.AutodiffCost(n, m, cartpole_cost([0.1, 5.0, 0.1, 0.1], [0.05]))
            xf = np.array([0, np.pi, 0, 0])
            term = TO.LQRCost(np.full(n, 100.0), np.full(m, 0.05), xf, terminal=True)
            cons = TO.ConstraintList(n, m, N)
            TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-4.0, u_max=4.0), (1, N - 1))
            TO.add_constraint(cons, TO.GoalConstraint(xf), N)
            p = cls(TO.Cartpole(), TO.Objective(stage, term, N), 0.1 * r.standard_normal((B, n)), 2.0, constraints=cons)
            TO.initial_controls(p, 0.5 + 0.1 * r.standard_normal((B, N - 1, m)))
        elif kind == "quadrotor_lqr_as_program":   # the BASELINE objective recorded as a user function: 17 variables, 153 second-order passes
            ref = P.quadrotor(B=3, N=15, cls=OracleProblem, dt=0.05)
            p = cls(ref.model, lqr_as_autodiff(ref), ref.x0, 0.05 * 14, xf=ref.xf, constraints=ref.constraints)
            TO.initial_controls(p, TO.controls(ref)); ref.close()
        else:                                      # every recorded operation, on the 2-D double integrator, terminal program cost too
            n, m, N, B = 4, 2, 11, 4

            def fun(x, u):
                a = TO.sin(x[0]) * TO.cos(x[1]) + TO.exp(0.3 * x[2]) / (2.0 + x[3] ** 2)
                b = TO.log(1.5 + u[0] ** 2) + TO.sqrt(2.0 + x[0] * x[0]) - TO.tanh(u[1] - x[1])
                c = (1.0 + x[2] ** 2) ** 1.5 + (3.0 - u[0]) ** 3 - (-x[3]) + 2.0 / (1.0 + u[1] ** 2)
                return a + b * 0.7 + c + 0.5 * (u[0] * u[0] + u[1] * u[1]) + 2.0 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3])
            term = TO.AutodiffCost(n, m, lambda x, u: 10.0 * ((x[0] - 1.0) ** 2 + x[1] ** 2) + TO.cos(x[2]) + x[3] ** 4, terminal=True)
            p = cls(TO.DoubleIntegrator(2), TO.Objective(TO.AutodiffCost(n, m, fun), term, N), 0.3 * r.standard_normal((B, n)), 1.0)
            TO.initial_controls(p, 0.3 * r.standard_normal((B, N - 1, m)))
        probs.append(p)
        r = np.random.default_rng(21)
    return probs


@pytest.mark.parametrize("kind", ["cartpole_docs_cost", "quadrotor_lqr_as_program", "double_integrator_all_ops"])
def test_autodiff_costs_in_all_kernels(kind):
    """user costs recorded as programs (RD.@autodiff CostFunction, docs/src/costfunction_interface.md:30-50) -- value,
    ForwardAD gradient / Hessian, AL expansion, and the solver kernels (materialised expansion + dense Riccati pass, lie.cu)"""
    g, o = _autodiff_pair(kind)
    for p in (g, o):
        TO.rollout(p)
    close(TO.cost_knots(g), TO.cost_knots(o), KERNEL_RTOL, "cost")
    close(TO.cost_gradient(g), TO.cost_gradient(o), KERNEL_RTOL, "gradient")
    close(TO.cost_hessian(g), TO.cost_hessian(o), KERNEL_RTOL, "Hessian")
    gg, gh = TO.al_expansion(g); og, oh = TO.al_expansion(o)
    close(gg, og, KERNEL_RTOL, "AL gradient"); close(gh, oh, KERNEL_RTOL, "AL Hessian")
    if kind == "cartpole_docs_cost":   # the closed forms of test/nlcosts.jl:41-44 with Q = (0.1, 5, 0.1, 0.1), R = 0.05
        X, U, H, gr = TO.states(g), TO.controls(g), TO.cost_hessian(g), TO.cost_gradient(g)
        assert np.allclose(H[:, :-1, 1, 1], -0.25 * 5.0 * np.cos(X[:, :-1, 1] / 2), rtol=1e-12)
        assert np.allclose(gr[:, :-1, 1], -0.5 * 5.0 * np.sin(X[:, :-1, 1] / 2), rtol=1e-12) and np.allclose(gr[:, :-1, 4], 0.05 * U[..., 0], rtol=1e-12)
    for p in (g, o):
        TO.expand(p)
    assert np.array_equal(TO.backward(g), TO.backward(o))
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, 1e-8, "K"); close(dg, do, 1e-8, "d")
    Jg, ag = TO.forward(g); Jo, ao = TO.forward(o)
    assert np.array_equal(ag, ao)
    close(Jg, Jo, 1e-8, "J")
    for p in (g, o):
        TO.ilqr_step(p, 2)
        if len(p.constraints):
            TO.al_update(p); TO.ilqr_step(p, 1)
    close(TO.merit(g), TO.merit(o), 1e-5, "merit")
    g.close(); o.close()


def test_autodiff_constraints_in_all_kernels():
    """SURVEY 8(f4), constraints: user constraints recorded as programs (docs/src/constraint_interface.md:52-72) -- values, forward-mode
    Jacobians, AL merit / expansion and the general-constraint path of the solver kernels, against the oracle; the docs' ControlNorm KAT"""
    from test_oracle_nlcost import control_norm
    n, m, N, B = 13, 4, 21, 4
    probs = []
    for cls in (TO.Problem, OracleProblem):
        base = P.quadrotor(B=B, N=N, cls=cls, dt=0.05)
        cons = base.constraints
        TO.add_constraint(cons, TO.AutodiffConstraint(n, m, control_norm(7.0), TO.Inequality(), "control"), (1, N - 1))
        TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x: [-(x[0] - 0.5) ** 2 - (x[1] - 1.0) ** 2 - (x[2] - 1.5) ** 2 + 0.09,
                                                                        TO.tanh(x[7]) + x[8] * x[9] - 2.0], TO.Inequality(), "state"), (2, N))
        TO.add_constraint(cons, TO.AutodiffConstraint(n, m, lambda x, u: [u[0] - u[2] + 0.1 * TO.sin(x[10])], TO.Equality()), (1, 5))
        p = cls(base.model, base.obj, base.x0, 0.05 * (N - 1), xf=base.xf, constraints=cons)
        TO.initial_controls(p, TO.controls(base)); base.close(); TO.rollout(p)
        probs.append(p)
    g, o = probs
    U = TO.controls(g)
    close(TO.evaluate_constraints(g, 2)[..., 0], np.linalg.norm(U, axis=-1) - 7.0, 1e-13, "ControlNorm value (docs KAT)")
    close(TO.constraint_jacobians(g, 2)[..., 0, n:], U / np.linalg.norm(U, axis=-1)[..., None], 1e-13, "ControlNorm Jacobian (docs KAT)")
    rl = np.random.default_rng(2)
    for i in range(len(g.constraints)):
        close(TO.evaluate_constraints(g, i), TO.evaluate_constraints(o, i), KERNEL_RTOL, f"constraint {i} values")
        close(TO.constraint_jacobians(g, i), TO.constraint_jacobians(o, i), KERNEL_RTOL, f"constraint {i} jacobians")
        lam = rl.standard_normal(TO.evaluate_constraints(o, i).shape)       # second-order term d/dz (grad c' lambda), src/abstract_constraint.jl:267-280
        close(TO.constraint_hessians(g, i, lam), TO.constraint_hessians(o, i, lam), KERNEL_RTOL, f"constraint {i} second-order term")
    close(TO.constraint_hessians(g, 2), TO.constraint_hessians(o, 2), KERNEL_RTOL, "second-order term with the current multipliers")
    close(TO.merit(g), TO.merit(o), KERNEL_RTOL, "merit"); close(TO.max_violation(g), TO.max_violation(o), KERNEL_RTOL, "violation")
    gg, gh = TO.al_expansion(g); og, oh = TO.al_expansion(o)
    close(gg, og, KERNEL_RTOL, "AL gradient"); close(gh, oh, KERNEL_RTOL, "AL hessian")
    for p in (g, o):
        TO.expand(p)
    assert np.array_equal(TO.backward(g), TO.backward(o))
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, 1e-8, "K"); close(dg, do, 1e-8, "d")
    for p in (g, o):
        TO.forward(p); TO.ilqr_step(p, 2); TO.al_update(p); TO.ilqr_step(p, 1)
    close(TO.merit(g), TO.merit(o), 1e-5, "merit after iterations")
    for i in range(len(g.constraints)):
        close(TO.multipliers(g, i), TO.multipliers(o, i), 1e-6, f"multipliers {i}")
    g.close(); o.close()


def test_regularisation_restart_matches_oracle():
    n, m, N = 4, 1, 11
    stage = TO.DiagonalCost(np.ones(n), -0.5 * np.ones(m))
    term = TO.DiagonalCost(np.ones(n), -0.5 * np.ones(m), terminal=True)
    probs = [cls(TO.Cartpole(), TO.Objective(stage, term, N), np.array([[0, 0.1, 0, 0], [0, -0.2, 0.1, 0]]), 0.5) for cls in (TO.Problem, OracleProblem)]
    for p in probs:
        TO.rollout(p); TO.expand(p)
    sg, so = TO.backward(probs[0]), TO.backward(probs[1])
    assert np.array_equal(sg, so) and np.all(sg > 0)
    third = TO.Problem(TO.Cartpole(), TO.Objective(stage, term, N), np.array([[0, 0.1, 0, 0], [0, -0.2, 0.1, 0]]), 0.5)
    TO.set_options(third, backward_kernel=2); TO.rollout(third); TO.expand(third)
    assert np.array_equal(TO.backward(third), so)                                  # thread-per-instance kernel: same restarts
    close(TO.solver_state(third)["rho"], TO.solver_state(probs[1])["rho"], 1e-12, "rho (thread kernel)")
    close(TO.gains(third)[0], TO.gains(probs[1])[0], 1e-9, "K (thread kernel)")
    close(TO.solver_state(probs[0])["rho"], TO.solver_state(probs[1])["rho"], 1e-12, "rho")
    Kg, dg = TO.gains(probs[0]); Ko, do = TO.gains(probs[1])
    close(Kg, Ko, 1e-9, "K"); close(dg, do, 1e-9, "d")


def test_line_search_backtracking_matches_oracle():
    """a poor initial guess forces alpha < 1 on some instances: the parallel ladder must pick the same step."""
    probs = [P.cartpole(B=16, N=101, cls=cls) for cls in (TO.Problem, OracleProblem)]
    for p in probs:
        TO.rollout(p)
        TO.ilqr_step(p, 6)
    sg, so = TO.solver_state(probs[0]), TO.solver_state(probs[1])
    assert np.array_equal(sg["alpha"], so["alpha"]) and np.array_equal(sg["ls_iters"], so["ls_iters"])
    assert np.any(so["ls_iters"] > 1)
    close(TO.merit(probs[0]), TO.merit(probs[1]), ITER_RTOL, "merit")


def test_cones_match_oracle_and_reference_kats():
    r = np.random.default_rng(1)
    pts = np.array([[2, 3, 1, 1.0], [2, 3, 1, -10.0], [2, 3, 1, 10.0]])   # test/cone_tests.jl:51,58,64
    x = np.vstack([pts, r.standard_normal((200, 4)) * 2])
    b = r.standard_normal(x.shape)
    for cone in (TO.SecondOrderCone(), TO.NegativeOrthant(), TO.ZeroCone(), TO.IdentityCone(), TO.PositiveOrthant()):
        close(TO.projection(cone, x), oracle_projection(cone, x)[0], 1e-14, f"projection {cone}")
        close(TO.grad_projection(cone, x), oracle_grad_projection(cone, x)[0], 1e-13, f"grad projection {cone}")
        close(TO.hess_projection(cone, x, b), oracle_hess_projection(cone, x, b)[0], 1e-12, f"hess projection {cone}")
    assert np.array_equal(TO.projection(TO.Inequality(), np.array([1, 2, -3.0])), [0, 0, -3.0])   # test/cone_tests.jl:70-75


def test_quickstart_problem_with_general_constraints():
    """the full examples/quickstart.jl problem (Goal + Circle + SOC norm + bounds): evaluation AND iLQR iterations on the
    device (general-constraint AL path of the solver kernels) against the oracle."""
    r = np.random.default_rng(1)
    model = TO.DoubleIntegrator(2)
    n, m, N = 4, 2, 21
    xf = np.array([0, 2.0, 0, 0])
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n) * (N - 1), xf, N)
    U0 = r.standard_normal((N - 1, m))
    probs = []
    for cls in (TO.Problem, OracleProblem):
        cons = TO.ConstraintList(n, m, N)
        TO.add_constraint(cons, TO.GoalConstraint(xf), N)
        TO.add_constraint(cons, TO.CircleConstraint(n, [0.0], [1.0], [0.5]), (2, N - 1))
        TO.add_constraint(cons, TO.NormConstraint(n, m, 5.0, TO.SecondOrderCone(), "control"), (1, N - 1))
        TO.add_constraint(cons, TO.BoundConstraint(n, m, u_min=-10, u_max=10), (1, N - 1))
        p = cls(model, obj, np.zeros(n), 3.0, xf=xf, constraints=cons)
        TO.initial_controls(p, U0); TO.rollout(p)
        probs.append(p)
    g, o = probs
    for i in range(4):
        close(TO.evaluate_constraints(g, i), TO.evaluate_constraints(o, i), KERNEL_RTOL, f"values {i}")
        close(TO.constraint_jacobians(g, i), TO.constraint_jacobians(o, i), KERNEL_RTOL, f"jacobians {i}")
    close(TO.merit(g), TO.merit(o), KERNEL_RTOL, "merit with SOC penalty")
    close(TO.max_violation(g), TO.max_violation(o), KERNEL_RTOL, "violation")
    gg, gh = TO.al_expansion(g); og, oh = TO.al_expansion(o)
    close(gg, og, KERNEL_RTOL, "AL gradient"); close(gh, oh, KERNEL_RTOL, "AL hessian")
    for p in probs:
        TO.ilqr_step(p, 2)
    close(TO.merit(g), TO.merit(o), ITER_RTOL, "merit after 2 iterations (SOC + circle + bounds + goal)")
    close(TO.states(g), TO.states(o), 1e-6, "X")
    for p in probs:
        TO.al_update(p); TO.ilqr_step(p, 1)
    close(TO.merit(g), TO.merit(o), 1e-5, "merit after AL update")
    for i in range(4):
        close(TO.multipliers(g, i), TO.multipliers(o, i), 1e-6, f"multipliers {i}")


def _general_pair(kind):
    r = np.random.default_rng(3)
    probs = []
    for cls in (TO.Problem, OracleProblem):
        if kind == "cartpole_linear_circle":
            prob = P.cartpole(B=4, N=41, cls=cls, u_bound=3.0, goal=True, dt_scaled_cost=True)
            n, m, N = 4, 1, 41
            cons = prob.constraints
            extra = [(TO.LinearConstraint(n, m, [[1.0], [-1.0]], [2.5, 2.5], TO.Inequality(), "control"), (1, N - 1)),
                     (TO.CircleConstraint(n, [1.5, -1.5], [2.0, 1.0], [0.3, 0.2]), (2, N - 1)),
                     (TO.LinearConstraint(n, m, [[0.0, 0.0, 1.0, 0.0]], [0.0], TO.Equality(), "state"), N)]
        else:
            prob = P.quadrotor(B=3, N=21, cls=cls, dt=0.05)
            n, m, N = 13, 4, 21
            cons = prob.constraints
            extra = [(TO.SphereConstraint(n, [0.5, 0.2], [1.0, 0.5], [1.5, 1.2], [0.3, 0.25]), (2, N - 1)),
                     (TO.NormConstraint(n, m, 12.0, TO.SecondOrderCone(), "control"), (1, N - 1)),
                     (TO.NormConstraint(n, m, 3.0, TO.Inequality(), [8, 9, 10]), (1, N))]
            if kind == "quadrotor_collision_statebound":
                extra = [(TO.CollisionConstraint(n, [1, 2, 3], [8, 9, 10], 0.4), (1, N)),          # position vs (scaled) velocity: any index pair works
                         (TO.StateBound(n, x_max=np.concatenate([[1.2, 1.2, 2.5], np.full(10, np.inf)]), x_min=np.concatenate([[-0.2], np.full(12, -np.inf)])), (2, N)),
                         (TO.ControlBound(m, u_max=9.0), (1, N - 1))]
        prob.close()
        for con, inds in extra:
            TO.add_constraint(cons, con, inds)
        prob2 = cls(prob.model, prob.obj, prob.x0, 5.0 if kind.startswith("cart") else 1.0, xf=prob.xf, constraints=cons)
        TO.initial_controls(prob2, 0.3 if kind.startswith("cart") else TO.Quadrotor().hover_control() + 0.1)
        probs.append(prob2)
    return probs


@pytest.mark.parametrize("kind", ["cartpole_linear_circle", "quadrotor_sphere_soc_norm", "quadrotor_collision_statebound"])
def test_general_constraints_in_solver_kernels(kind):
    g, o = _general_pair(kind)
    for p in (g, o):
        TO.rollout(p); TO.expand(p)
    rl = np.random.default_rng(2)
    for i in range(len(g.constraints)):
        close(TO.evaluate_constraints(g, i), TO.evaluate_constraints(o, i), KERNEL_RTOL, f"constraint {i} values")
        close(TO.constraint_jacobians(g, i), TO.constraint_jacobians(o, i), KERNEL_RTOL, f"constraint {i} jacobians")
        lam = rl.standard_normal(TO.evaluate_constraints(o, i).shape)
        close(TO.constraint_hessians(g, i, lam), TO.constraint_hessians(o, i, lam), KERNEL_RTOL, f"constraint {i} second-order term")
    gg, gh = TO.al_expansion(g); og, oh = TO.al_expansion(o)
    close(gg, og, KERNEL_RTOL, "AL gradient"); close(gh, oh, KERNEL_RTOL, "AL hessian")
    assert np.array_equal(TO.backward(g), TO.backward(o))
    Kg, dg = TO.gains(g); Ko, do = TO.gains(o)
    close(Kg, Ko, 1e-8, "K"); close(dg, do, 1e-8, "d")
    Jg, ag = TO.forward(g); Jo, ao = TO.forward(o)
    assert np.array_equal(ag, ao)
    close(Jg, Jo, 1e-8, "J")
    for p in (g, o):
        TO.ilqr_step(p, 2); TO.al_update(p); TO.ilqr_step(p, 1)
    close(TO.merit(g), TO.merit(o), 1e-5, "merit")
    close(TO.max_violation(g), TO.max_violation(o), 1e-5, "violation")


def test_mpc_update_trajectory_and_shift_match_oracle():
    """f3 (BASELINE config 5): tracking objective re-targeted with update_trajectory!, receding-horizon shift, re-solve"""
    r = np.random.default_rng(11)
    n, m, N, nref, B = 13, 4, 51, 80, 6
    hover = TO.Quadrotor().hover_control()
    t = np.linspace(0, 4, nref)
    Xref = np.zeros((nref, n)); Xref[:, 0] = np.sin(t); Xref[:, 1] = 0.5 * t; Xref[:, 2] = 1.0; Xref[:, 3] = 1.0
    Uref = np.tile(hover, (nref, 1))
    Q, R, Qf = np.full(n, 1.0), np.full(m, 0.1), np.full(n, 10.0)
    x0 = np.tile(Xref[0], (B, 1)) + 0.05 * r.standard_normal((B, n)); x0[:, 3:7] = [1, 0, 0, 0]
    probs = []
    for cls in (TO.Problem, OracleProblem):
        cons = TO.ConstraintList(n, m, N)
        TO.add_constraint(cons, TO.ControlBound(m, u_min=0.0, u_max=8.0), (1, N - 1))
        p = cls(TO.Quadrotor(), TO.TrackingObjective(Q, R, Xref[:N], Uref[:N - 1], Qf=Qf), x0, 2.5, constraints=cons)
        TO.initial_controls(p, hover); TO.rollout(p)
        probs.append(p)
    g, o = probs
    for start in (1, 4):
        for p in probs:
            if start > 1:
                TO.shift_trajectory(p, 3)
                TO.update_trajectory(p, Xref, Uref, start)
                TO.set_initial_state(p, TO.states(p)[:, 0] + 0.01)       # "measured" state
                TO.rollout(p)
            TO.ilqr_step(p, 3); TO.al_update(p); TO.ilqr_step(p, 1)
        close(TO.gettimes(g), TO.gettimes(o), 1e-15, "times")
        close(TO.cost_gradient(g), TO.cost_gradient(o), KERNEL_RTOL, "tracking gradient")
        close(TO.multipliers(g, 0), TO.multipliers(o, 0), 1e-6, "multipliers")
        close(TO.states(g), TO.states(o), 1e-6, "X"); close(TO.controls(g), TO.controls(o), 1e-6, "U")
        close(TO.merit(g), TO.merit(o), ITER_RTOL, "merit")
    for p in probs:
        p.close()


def test_error_behaviour():
    n, m, N = 4, 1, 5
    obj = TO.LQRObjective(np.eye(n), np.eye(m), np.eye(n), np.zeros(n), N)
    with pytest.raises(TO.DimensionMismatch):       # src/problem.jl:64-68
        TO.Problem(TO.Quadrotor(), obj, np.zeros(13), 1.0)
    with pytest.raises(TO.ArgumentError):           # src/problem.jl:87-91
        TO.Problem(TO.Cartpole(), obj, np.zeros(n), 1.0, x0=np.zeros(n))
    prob = TO.Problem(TO.Cartpole(), obj, np.zeros(n), 1.0)
    assert np.isnan(TO.states(prob)).all() and not TO.controls(prob).any()   # X0 = NaN, U0 = 0 (src/problem.jl:83-84)
    with pytest.raises(TO.TrajOptError):
        TO.backward(prob)                           # backward pass before the expansion
    with pytest.raises(TO.DimensionMismatch):
        TO.initial_controls(prob, np.zeros((3, 7)))


def test_full_size_properties_quadrotor():
    """BASELINE full size (B=4096, N=101): size-independent properties instead of an element-wise oracle run:
    hover invariance, batch permutation equivariance, merit monotonicity, cost = sum of knot costs."""
    B = 4096
    prob = P.quadrotor(B=B, N=101)
    TO.rollout(prob)
    J0 = TO.merit(prob)
    assert np.allclose(TO.cost_knots(prob).sum(axis=1), TO.cost(prob), rtol=1e-12)
    TO.ilqr_step(prob, 2)
    J2 = TO.merit(prob)
    assert np.all(J2 <= J0 + 1e-9) and np.all(TO.solver_state(prob)["bp_status"] >= 0)
    # permutation equivariance: the same instances in reversed order give the reversed result
    prob2 = P.quadrotor(B=B, N=101)
    TO.set_initial_state(prob2, prob2.x0[::-1]); TO.initial_controls(prob2, TO.controls(prob2)[::-1])
    TO.rollout(prob2); TO.ilqr_step(prob2, 2)
    assert np.allclose(TO.merit(prob2)[::-1], J2, rtol=1e-12)
    # hover: x0 with hover controls stays put (test/internal_api.jl:50-56)
    x0 = np.array([1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    TO.set_initial_state(prob, x0); TO.initial_controls(prob, TO.Quadrotor().hover_control()); TO.rollout(prob)
    assert np.allclose(TO.states(prob)[:, -1], x0, rtol=1.5e-8)


def test_reduce_merit_buffer_tracks_line_search_results():
    """{sum J, max violation} maintained by the line search (no extra sweep) equals a fresh cost + constraint sweep"""
    import torch
    prob = P.quadrotor(B=257, N=101)
    TO.rollout(prob)
    TO.ilqr_step(prob, 3)
    t2 = TO.multi_gpu.merit_device_tensor(prob, "cuda:0")
    TO.multi_gpu.global_merit(prob, device_tensor=t2)
    torch.cuda.synchronize()
    got = t2.cpu().numpy().copy()
    assert np.isclose(got[0], TO.merit(prob).sum(), rtol=1e-12)
    assert np.isclose(got[1], TO.max_violation(prob).max(), rtol=1e-12)
    # on a consumer stream of its own, interleaved with further iterations (the bench's multi-GPU step)
    import ctypes as C
    side = torch.cuda.Stream()
    ref = P.quadrotor(B=257, N=101); TO.rollout(ref); TO.ilqr_step(ref, 3)
    for _ in range(3):
        TO.ilqr_step(prob, 1); TO.ilqr_step(ref, 1)
        assert prob._lib.to_reduce_merit_async(prob._h, C.c_void_p(side.cuda_stream)) == 0
        side.synchronize()
        assert np.isclose(t2.cpu().numpy()[0], TO.merit(ref).sum(), rtol=1e-12)
    np.testing.assert_array_equal(TO.controls(prob), TO.controls(ref))
    TO.al_update(prob)                       # invalidates J: the next reduce recomputes it
    TO.multi_gpu.global_merit(prob, device_tensor=t2)
    torch.cuda.synchronize()
    assert np.isclose(t2.cpu().numpy()[0], TO.merit(prob).sum(), rtol=1e-12)
