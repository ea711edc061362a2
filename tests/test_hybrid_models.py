"""Hybrid / variable-dimension models: the reference's ``Problem(models::Vector{<:DiscreteDynamics}, ...)`` (src/problem.jl:36-73,
src/dynamics.jl:15-31, src/constraint_list.jl:25-66) on its own example test/hybrid_dynamics_model.jl -- a 4-state / 2-control integrator for five
steps, a jump map onto 2 states, a 2-state / 1-control integrator for four more.  Host checks mirror that file's @test lines; the
oracle's recorded-program path is checked against closed forms and finite differences here, the CUDA path against the oracle in the GPU half."""
import numpy as np
import pytest

import trajopt_b200 as TO
from oracle_binding import OracleProblem


def hybrid_models():
    model1 = TO.AutodiffDynamics(4, 2, lambda x, u: [x[2], x[3], u[0], u[1]])                         # test/hybrid_dynamics_model.jl:27-29
    jumpmap = TO.AutodiffDynamics(4, 2, lambda x, u: [(x[2] + x[3]) / 2, (u[0] + u[1]) / 2], output_dim=2, discrete=True)   # :33-35
    model2 = TO.AutodiffDynamics(2, 1, lambda x, u: [x[1], u[0]])                                       # :37-39
    return model1, jumpmap, model2, [model1] * 5 + [jumpmap] + [model2] * 4


def hybrid_problem(cls=TO.Problem, constrained=True, batch=1, xg=(0.3, -0.2), **kw):
    model1, jumpmap, model2, models = hybrid_models()
    nx, nu = TO.dims(models)
    obj = TO.Objective([TO.LQRCost(np.full(nx[k], 1.0), np.full(nu[k], 0.1), np.zeros(nx[k])) for k in range(11)])   # :69-72
    cons = TO.ConstraintList(models)
    if constrained:
        TO.add_constraint(cons, TO.BoundConstraint(4, 2, u_max=4, u_min=-4), (1, 5))                  # :85-91
        TO.add_constraint(cons, TO.BoundConstraint(2, 1, u_max=2, u_min=-2, x_max=[10, np.inf]), (7, 10))
        TO.add_constraint(cons, TO.GoalConstraint(np.asarray(xg, dtype=float)), 11)
    return cls(models, obj, np.zeros(4), 2.0, constraints=cons, batch=batch, **kw), models


def test_model_dimension_vectors_and_the_reference_error_cases():
    model1, jumpmap, model2, models = hybrid_models()
    nx, nu = TO.dims(models)
    assert nx == [4, 4, 4, 4, 4, 4, 2, 2, 2, 2, 2]                    # test/hybrid_dynamics_model.jl:53
    assert nu == [2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1]                    # :54
    models_bad = [model1] * 5 + [model2] * 5
    with pytest.raises(TO.DimensionMismatch, match="Model mismatch at time step 5"):   # :61-65
        TO.dims(models_bad)
    cons = TO.ConstraintList(models)
    bnd1 = TO.BoundConstraint(4, 2, u_max=4, u_min=-4)
    bnd2 = TO.BoundConstraint(2, 1, u_max=2, u_min=-2, x_max=[10, np.inf])
    TO.add_constraint(cons, bnd1, (1, 5)); TO.add_constraint(cons, bnd2, (7, 10)); TO.add_constraint(cons, TO.GoalConstraint(np.zeros(2)), 11)
    assert list(TO.num_constraints(cons)) == [4, 4, 4, 4, 4, 0, 3, 3, 3, 3, 2]        # :96
    with pytest.raises(TO.DimensionMismatch, match="time step 7"):                    # :99
        TO.add_constraint(cons, bnd1, (3, 8))
    with pytest.raises(TO.DimensionMismatch):                                         # :100
        TO.add_constraint(cons, bnd2, (1, 3))
    assert TO.ConstraintList(nx, nu).nx == nx and TO.ConstraintList(4, 2, 11).nx == [4] * 11


def test_problem_construction_checks_and_rollout_on_the_oracle():
    prob, models = hybrid_problem(OracleProblem)
    nx, nu = TO.dims(models)
    assert TO.dims(prob, 5)[:2] == (4, 2) and TO.dims(prob, 6)[:2] == (4, 2) and TO.dims(prob, 11)[:2] == (2, 1)   # :80-82
    assert list(TO.num_constraints(prob)) == [4, 4, 4, 4, 4, 0, 3, 3, 3, 3, 2]
    model1, jumpmap, model2, _ = hybrid_models()
    obj = prob.obj
    with pytest.raises(TO.DimensionMismatch):                                          # :103  (models without the jump map)
        OracleProblem([model1] * 5 + [model2] * 5, obj, np.zeros(4), 2.0)
    obj_bad = TO.LQRObjective(np.ones(4), np.ones(2), np.ones(4), np.zeros(4), 11)
    with pytest.raises(TO.DimensionMismatch, match="Objective state"):                 # :105-106
        OracleProblem(models, obj_bad, np.zeros(4), 2.0)
    cons_bad = TO.ConstraintList(4, 2, 11)
    TO.add_constraint(cons_bad, TO.BoundConstraint(4, 2, u_max=4, u_min=-4), (1, 5))
    with pytest.raises(TO.DimensionMismatch, match="Constraint state"):                # :108-110
        OracleProblem(models, obj, np.zeros(4), 2.0, constraints=cons_bad)
    nu_bad = list(nu); nu_bad[6] = 2
    cons_bad2 = TO.ConstraintList(nx, nu_bad)
    with pytest.raises(TO.DimensionMismatch, match="Constraint control"):              # :113-117
        OracleProblem(models, obj, np.zeros(4), 2.0, constraints=cons_bad2)
    obj_bad2 = TO.Objective([TO.LQRCost(np.ones(n), np.ones(m), np.zeros(n)) for n, m in zip(nx, nu_bad)])
    with pytest.raises(TO.DimensionMismatch, match="Objective control"):               # :119-120
        OracleProblem(models, obj_bad2, np.zeros(4), 2.0)

    # rollout: double integrators are integrated exactly by RK4 -> closed forms
    rng = np.random.default_rng(0)
    U = rng.standard_normal((1, 10, 2)); U[:, 6:, 1] = 0.0            # the second control does not exist after the jump
    TO.set_initial_state(prob, np.array([0.1, -0.2, 0.3, 0.4, ]))
    TO.initial_controls(prob, U)
    TO.rollout(prob)
    X = TO.states(prob)[0]
    h = 0.2
    x = np.array([0.1, -0.2, 0.3, 0.4])
    for k in range(5):
        u = U[0, k]
        x = np.concatenate([x[:2] + h * x[2:] + 0.5 * h * h * u, x[2:] + h * u])
        assert np.allclose(X[k + 1], x, atol=1e-14)
    xj = np.array([(x[2] + x[3]) / 2, (U[0, 5, 0] + U[0, 5, 1]) / 2, 0.0, 0.0])
    assert np.allclose(X[6], xj, atol=1e-15)                           # the jump map, padded slots zero
    x = xj[:2]
    for k in range(6, 10):
        u = U[0, k, 0]
        x = np.array([x[0] + h * x[1] + 0.5 * h * h * u, x[1] + h * u])
        assert np.allclose(X[k + 1, :2], x, atol=1e-14) and np.all(X[k + 1, 2:] == 0.0)
    # cost: the padded unit weights act on controls that are zero
    J = TO.cost(prob)[0]
    Jref = sum(0.5 * X[k, :nx[k]] @ X[k, :nx[k]] + (0.05 * U[0, k, :nu[k]] @ U[0, k, :nu[k]] if k < 10 else 0.0) for k in range(11))
    assert abs(J - Jref) < 1e-13
    # Jacobians against the closed forms
    TO.expand(prob)
    AB = TO.dynamics_jacobians(prob)[0]
    A1 = np.eye(4); A1[0, 2] = A1[1, 3] = h
    B1 = np.array([[0.5 * h * h, 0], [0, 0.5 * h * h], [h, 0], [0, h]])
    assert np.allclose(AB[0], np.hstack([A1, B1]), atol=1e-15)
    Aj = np.zeros((4, 6)); Aj[0, 2] = Aj[0, 3] = 0.5; Aj[1, 4] = Aj[1, 5] = 0.5
    assert np.allclose(AB[5], Aj, atol=1e-15)
    A2 = np.zeros((4, 6)); A2[0, 0] = A2[1, 1] = A2[2, 2] = A2[3, 3] = 1; A2[0, 1] = h;   # (the unused state slots: x+ = x = 0)
    A2[0, 4] = 0.5 * h * h; A2[1, 4] = h
    assert np.allclose(AB[7], A2, atol=1e-15)


def test_oracle_solves_the_hybrid_problem():
    prob, models = hybrid_problem(OracleProblem)
    TO.rollout(prob)
    J0 = TO.merit(prob)[0]
    for outer in range(6):
        TO.ilqr_step(prob, 8)
        TO.al_update(prob)
    X, U = TO.states(prob)[0], TO.controls(prob)[0]
    assert np.all(np.abs(U[:5]) <= 4 + 1e-6) and np.all(np.abs(U[6:, 0]) <= 2 + 1e-6)
    assert np.all(U[6:, 1] == 0.0) and np.all(X[6:, 2:] == 0.0)       # the padded entries never move
    assert np.linalg.norm(X[10, :2] - np.array([0.3, -0.2])) < 1e-4    # the goal constraint of the 2-state phase
    assert TO.max_violation(prob)[0] < 1e-4


@pytest.mark.gpu
def test_gpu_hybrid_problem_matches_the_oracle():
    B = 64
    rng = np.random.default_rng(5)
    gp, _ = hybrid_problem(TO.Problem, batch=B)
    op, _ = hybrid_problem(OracleProblem, batch=B)
    x0 = 0.3 * rng.standard_normal((B, 4))
    U = rng.standard_normal((B, 10, 2)); U[:, 6:, 1] = 0.0
    for p in (gp, op):
        TO.set_initial_state(p, x0); TO.initial_controls(p, U); TO.rollout(p); TO.expand(p)
    assert np.allclose(TO.states(gp), TO.states(op), atol=1e-13)
    assert np.allclose(TO.dynamics_jacobians(gp), TO.dynamics_jacobians(op), atol=1e-13)
    assert np.allclose(TO.cost(gp), TO.cost(op), rtol=1e-13)
    for outer in range(4):
        for p in (gp, op):
            TO.ilqr_step(p, 6); TO.al_update(p)
    assert np.allclose(TO.states(gp), TO.states(op), atol=1e-7)
    assert np.allclose(TO.controls(gp), TO.controls(op), atol=1e-7)
    assert np.allclose(TO.max_violation(gp), TO.max_violation(op), atol=1e-8)
    assert np.all(TO.controls(gp)[:, 6:, 1] == 0.0) and np.all(TO.states(gp)[:, 6:, 2:] == 0.0)
    # the library's own dimension check (what a C caller sees)
    model1, jumpmap, model2, models = hybrid_models()
    bad = TO.Problem.__new__(TO.Problem)
    K = TO.capi
    spec = gp.spec
    nx_bad = list(spec.nx); nx_bad[6] = 4
    s2 = K.Spec(K.MODEL_EXPR, 4, 2, 11, 1, np.full(10, 0.2), spec.costs, spec.cost_index, spec.cons, dyn=spec.dyn, dyn_index=spec.dyn_index, nx=nx_bad, nu=spec.nu)
    import ctypes as C
    h = C.c_void_p()
    rc = gp._lib.to_create(C.byref(s2.c), C.byref(h))
    assert rc == K.TO_EDIM and b"Model mismatch at time step 6" in gp._lib.to_last_error(None)
