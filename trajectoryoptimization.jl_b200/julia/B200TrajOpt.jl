# B200TrajOpt.jl -- thin ccall shim over libtrajopt_b200.so (include/trajopt_b200.h).
#
# NOT RUNNABLE IN THIS REPO'S IMAGE (no Julia there); kept syntactically careful and reviewed against the header.
# It gives Julia host code (and Altro.jl) a `BatchedProblem` that is built from an ordinary
# TrajectoryOptimization.Problem and overloads the functions a solver calls on it
# (rollout!, cost, evaluate_constraints!, ... -- SURVEY.md 2.3), so the hot path runs on the GPU for a whole batch.
module B200TrajOpt

using TrajectoryOptimization
using RobotDynamics
using LinearAlgebra
using Rotations
const TO = TrajectoryOptimization
const RD = RobotDynamics

const libb200 = get(ENV, "LIBTRAJOPT_B200", "libtrajopt_b200.so")

# ---- C structs (must match include/trajopt_b200.h field for field) ------------------------------------------
struct ToCostSpec
    kind::Int32; terminal::Int32
    Q::Ptr{Float64}; R::Ptr{Float64}; H::Ptr{Float64}; q::Ptr{Float64}; r::Ptr{Float64}
    c::Float64
    w::Float64; q_ref::Ptr{Float64}; q_ind::Ptr{Int32}       # DiagonalQuatCost (kind 2), else 0 / C_NULL
    prog_len::Int32; nconst::Int32; prog::Ptr{Int32}; consts::Ptr{Float64}   # recorded program (kind 3), else 0 / C_NULL
end
struct ToConstraintSpec
    kind::Int32; first::Int32; last::Int32; sense::Int32; p::Int32; flag::Int32; ninds::Int32
    inds::Ptr{Int32}; a::Ptr{Float64}; b::Ptr{Float64}; c::Ptr{Float64}; rad::Ptr{Float64}
    val::Float64
end
struct ToDynamicsSpec         # one recorded model of a hybrid problem (to_dynamics_spec; models given as programs: Python host API only for now)
    n_in::Int32; m_in::Int32; n_out::Int32; discrete::Int32; prog_len::Int32; nconst::Int32
    prog::Ptr{Int32}; consts::Ptr{Float64}
end
struct ToSpec
    model::Int32; n::Int32; m::Int32; N::Int32; B::Int32; device::Int32; nparams::Int32
    params::Ptr{Float64}; dt::Ptr{Float64}; t0::Float64
    ncost::Int32; costs::Ptr{ToCostSpec}; cost_index::Ptr{Int32}
    ncon::Int32; cons::Ptr{ToConstraintSpec}
    error_state::Int32       # 1: solver kernels on the Lie-group error state (RD.errstate_dim(model) != n), as Altro does
    ndyn::Int32; dyn::Ptr{ToDynamicsSpec}; dyn_index::Ptr{Int32}; nx::Ptr{Int32}; nu::Ptr{Int32}    # TO_MODEL_EXPR only, else 0 / C_NULL
end

const TO_EDIM = -2
const TO_EINVAL = -1

function check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:to_last_error, libb200), Cstring, (Ptr{Cvoid},), h))
    rc == TO_EDIM && throw(DimensionMismatch(msg))     # same exception types as src/problem.jl:64-68, :87-91
    rc == TO_EINVAL && throw(ArgumentError(msg))
    error("libtrajopt_b200 [$rc]: $msg")
end

model_id(::Any) = error("model not available on the device; supported: DoubleIntegrator, Cartpole, Quadrotor, Acrobot")

mutable struct BatchedProblem
    h::Ptr{Cvoid}
    prob::TO.Problem          # the template instance (objective / constraint objects stay the reference's)
    B::Int
    keep::Vector{Any}         # GC roots of every array whose pointer went into the spec
end

sense_code(::TO.Equality) = Int32(0)
sense_code(::TO.Inequality) = Int32(1)
sense_code(::TO.SecondOrderCone) = Int32(2)
sense_code(::TO.IdentityCone) = Int32(3)
sense_code(::TO.PositiveOrthant) = Int32(4)

# ---- ConstraintList entry -> to_constraint_spec (include/trajopt_b200.h to_con_kind), every constraint type of src/constraints.jl ----
# `root` keeps the arrays whose pointers go into the spec alive until to_create has copied them.
const NULLI = Ptr{Int32}(C_NULL); const NULLD = Ptr{Float64}(C_NULL)
function constraint_spec(con::TO.GoalConstraint, f, l, n, m, root)                    # src/constraints.jl:22-87
    ii = root(Vector{Int32}(con.inds)); a = root(Vector{Float64}(con.xf))
    ToConstraintSpec(0, f, l, 0, 0, 0, length(ii), pointer(ii), pointer(a), NULLD, NULLD, NULLD, 0.0)
end
function constraint_spec(con::TO.BoundConstraint, f, l, n, m, root)                   # :644-783
    a = root(Vector{Float64}(con.z_max)); b = root(Vector{Float64}(con.z_min))
    ToConstraintSpec(1, f, l, 1, 0, 0, 0, NULLI, pointer(a), pointer(b), NULLD, NULLD, 0.0)
end
function bound_base_spec(bnd, f, l, n, m, root, control::Bool)                        # StateBound / ControlBound :547-631 = Bound with the other block open
    zmax = fill(Inf, n + m); zmin = fill(-Inf, n + m); off = control ? n : 0
    zmax[off .+ bnd.i_max] .= bnd.x_max; zmin[off .+ bnd.i_min] .= bnd.x_min
    a = root(zmax); b = root(zmin)
    ToConstraintSpec(1, f, l, 1, 0, 0, 0, NULLI, pointer(a), pointer(b), NULLD, NULLD, 0.0)
end
constraint_spec(con::TO.StateBound, f, l, n, m, root) = bound_base_spec(con.bnd, f, l, n, m, root, false)
constraint_spec(con::TO.ControlBound, f, l, n, m, root) = bound_base_spec(con.bnd, f, l, n, m, root, true)
function constraint_spec(con::TO.LinearConstraint, f, l, n, m, root)                  # :103-150 -- A acts on z[inds]; the device takes the x or the u block
    inds = collect(con.inds); P = length(con.b)
    A = Matrix{Float64}(con.A)
    if all(i -> i <= n, inds)
        Af = zeros(P, n); Af[:, inds] .= A; flag = 0
    elseif all(i -> i > n, inds)
        Af = zeros(P, m); Af[:, inds .- n] .= A; flag = 1
    else
        throw(ArgumentError("LinearConstraint across states and controls: split it into a state and a control constraint for the device"))
    end
    a = root(Af); b = root(Vector{Float64}(con.b))                                     # column-major p x w, as the ABI wants
    ToConstraintSpec(2, f, l, sense_code(con.sense), P, flag, 0, NULLI, pointer(a), pointer(b), NULLD, NULLD, 0.0)
end
function constraint_spec(con::TO.CircleConstraint, f, l, n, m, root)                  # :168-233
    a = root(Vector{Float64}(con.x)); b = root(Vector{Float64}(con.y)); r = root(Vector{Float64}(con.radius))
    ii = root(Int32[con.xi, con.yi])
    ToConstraintSpec(3, f, l, 1, length(a), 0, 2, pointer(ii), pointer(a), pointer(b), NULLD, pointer(r), 0.0)
end
function constraint_spec(con::TO.SphereConstraint, f, l, n, m, root)                  # :249-326
    a = root(Vector{Float64}(con.x)); b = root(Vector{Float64}(con.y)); c = root(Vector{Float64}(con.z)); r = root(Vector{Float64}(con.radius))
    ii = root(Int32[con.xi, con.yi, con.zi])
    ToConstraintSpec(4, f, l, 1, length(a), 0, 3, pointer(ii), pointer(a), pointer(b), pointer(c), pointer(r), 0.0)
end
function constraint_spec(con::TO.NormConstraint, f, l, n, m, root)                    # :438-521 (the static evaluate: z[inds[j]], SURVEY 2.4)
    ii = root(Vector{Int32}(con.inds))
    ToConstraintSpec(5, f, l, sense_code(con.sense), 0, 0, length(ii), pointer(ii), NULLD, NULLD, NULLD, NULLD, Float64(con.val))
end
function constraint_spec(con::TO.CollisionConstraint, f, l, n, m, root)               # :341-389
    ii = root(Int32[con.x1; con.x2])
    ToConstraintSpec(6, f, l, 1, 0, 0, length(ii), pointer(ii), NULLD, NULLD, NULLD, NULLD, con.radius)
end
function constraint_spec(con::TO.QuatVecEq, f, l, n, m, root)                         # :938-965
    a = root(Vector{Float64}(Rotations.params(con.qf))); ii = root(Vector{Int32}(con.qind))
    ToConstraintSpec(7, f, l, 0, 3, 0, 4, pointer(ii), pointer(a), NULLD, NULLD, NULLD, 0.0)
end
function constraint_spec(con::TO.IndexedConstraint, f, l, n, m, root)                 # :820-936: the inner constraint with its indices moved into the new z
    ix, iu = collect(con.ix), collect(con.iu) .- con.n                                # positions of the old x / u inside the new x / u
    constraint_spec(TO.change_dimension(con.con, n, m, ix, iu), f, l, n, m, root)
end
function constraint_spec(con::TO.StageConstraint, f, l, n, m, root)                   # RD.@autodiff user constraint: record RD.evaluate
    tape = root(record((x, u) -> RD.evaluate(con, x, u), n, m))
    P = RD.output_dim(con)
    ToConstraintSpec(8, f, l, sense_code(TO.sense(con)), P, length(tape.consts), length(tape.prog), pointer(tape.prog), pointer(tape.consts), NULLD, NULLD, NULLD, 0.0)
end

"""
    BatchedProblem(prob::TO.Problem, model_id, B; device=0, params=Float64[])

Describe `prob` (objective = vector of QuadraticCostFunctions, ConstraintList of Goal/Bound/Linear/Circle/Sphere/Norm)
to the library and allocate a batch of `B` instances on `device`.
"""
function BatchedProblem(prob::TO.Problem, mid::Integer, B::Integer; device::Integer=0, params::Vector{Float64}=Float64[],
                        error_state::Bool=(RD.errstate_dim(TO.get_model(prob)[1]) != RD.state_dim(TO.get_model(prob)[1])))
    n, m, N = RD.dims(prob, 1)
    keep = Any[]
    root(x) = (push!(keep, x); x)
    obj = TO.get_objective(prob)
    costs = ToCostSpec[]
    index = Int32[]
    seen = IdDict{Any,Int32}()
    for k = 1:N
        c = obj[k]
        if !haskey(seen, c) && !(c isa TO.QuadraticCostFunction)      # RD.@autodiff user cost: record RD.evaluate
            tape = root(record((x, u) -> RD.evaluate(c, x, u), n, m))
            push!(costs, expr_cost_spec(tape, k == N))
            seen[c] = Int32(length(costs) - 1)
        elseif !haskey(seen, c)
            isdiag = TO.is_diag(c)
            Q = root(isdiag ? Vector{Float64}(diag(c.Q)) : Matrix{Float64}(c.Q))
            R = root(isdiag ? Vector{Float64}(diag(c.R)) : Matrix{Float64}(c.R))
            H = isdiag ? C_NULL : pointer(root(Matrix{Float64}(c.H)))
            q = root(Vector{Float64}(c.q)); r = root(Vector{Float64}(c.r))
            if c isa TO.DiagonalQuatCost                         # src/lie_costs.jl:33-56
                qref = root(Vector{Float64}(c.q_ref)); qind = root(Vector{Int32}(c.q_ind))
                push!(costs, ToCostSpec(2, c.terminal ? 1 : 0, pointer(Q), pointer(R), C_NULL, pointer(q), pointer(r), c.c, c.w, pointer(qref), pointer(qind), 0, 0, C_NULL, C_NULL))
            else
                push!(costs, ToCostSpec(isdiag ? 0 : 1, c.terminal ? 1 : 0, pointer(Q), pointer(R), H, pointer(q), pointer(r), c.c, 0.0, C_NULL, C_NULL, 0, 0, C_NULL, C_NULL))
            end
            seen[c] = Int32(length(costs) - 1)
        end
        push!(index, seen[c])
    end
    cons = ToConstraintSpec[]
    for (inds, con) in zip(TO.get_constraints(prob))
        f, l = Int32(first(inds)), Int32(last(inds))
        push!(cons, constraint_spec(con, f, l, n, m, root))
    end
    dt = root(Vector{Float64}([RD.timestep(z) for z in TO.get_trajectory(prob)][1:N-1]))
    root(costs); root(index); root(cons); root(params)
    spec = Ref(ToSpec(mid, n, m, N, B, device, length(params), isempty(params) ? C_NULL : pointer(params), pointer(dt),
                      TO.get_initial_time(prob), length(costs), pointer(costs), pointer(index), length(cons),
                      isempty(cons) ? C_NULL : pointer(cons), error_state ? 1 : 0, 0, C_NULL, C_NULL, C_NULL, C_NULL))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve keep ccall((:to_create, libb200), Cint, (Ref{ToSpec}, Ref{Ptr{Cvoid}}), spec, h)
    check(C_NULL, rc)
    bp = BatchedProblem(h[], prob, B, keep)
    finalizer(p -> ccall((:to_destroy, libb200), Cint, (Ptr{Cvoid},), p.h), bp)
    return bp
end

# ---- user-defined costs / constraints (RD.@autodiff types): record RD.evaluate once, ship the tape ---------------------
# The device differentiates a straight-line program with second-order forward-mode duals (include/trajopt_b200.h to_expr_op).
# `Rec` is a number type that appends one instruction per arithmetic operation -- the trick ForwardDiff.Dual uses to see the
# user's function, applied to recording instead of differentiating.
mutable struct Tape
    prog::Vector{Int32}      # op, a, b triples (0-based operand indices)
    consts::Vector{Float64}
end
Tape() = Tape(Int32[], Float64[])
struct Rec <: Real
    tape::Tape
    idx::Int32               # 0-based index of the instruction that produced this value
end
function emit!(t::Tape, op, a, b)
    push!(t.prog, Int32(op), Int32(a), Int32(b))
    Rec(t, Int32(length(t.prog) ÷ 3 - 1))
end
function constindex!(t::Tape, v::Real)
    i = findfirst(c -> c === Float64(v), t.consts)
    i === nothing ? (push!(t.consts, Float64(v)); length(t.consts) - 1) : i - 1
end
Base.promote_rule(::Type{Rec}, ::Type{<:Real}) = Rec
for (f, op, opc, ropc) in ((:+, 3, 15, 15), (:*, 5, 16, 16))             # commutative: ADD / ADDC, MUL / MULC
    @eval Base.$f(a::Rec, b::Rec) = emit!(a.tape, $op, a.idx, b.idx)
    @eval Base.$f(a::Rec, b::Real) = emit!(a.tape, $opc, a.idx, constindex!(a.tape, b))
    @eval Base.$f(a::Real, b::Rec) = emit!(b.tape, $ropc, b.idx, constindex!(b.tape, a))
end
Base.:-(a::Rec, b::Rec) = emit!(a.tape, 4, a.idx, b.idx)
Base.:-(a::Rec, b::Real) = emit!(a.tape, 15, a.idx, constindex!(a.tape, -b))    # a + (-b)
Base.:-(a::Real, b::Rec) = emit!(b.tape, 19, b.idx, constindex!(b.tape, a))     # RSUBC
Base.:-(a::Rec) = emit!(a.tape, 7, a.idx, 0)
Base.:/(a::Rec, b::Rec) = emit!(a.tape, 6, a.idx, b.idx)
Base.:/(a::Rec, b::Real) = emit!(a.tape, 17, a.idx, constindex!(a.tape, b))     # DIVC
Base.:/(a::Real, b::Rec) = emit!(b.tape, 18, b.idx, constindex!(b.tape, a))     # RDIVC
Base.:^(a::Rec, p::Integer) = p == 2 ? a * a : emit!(a.tape, 13, a.idx, constindex!(a.tape, p))
Base.:^(a::Rec, p::Real) = emit!(a.tape, 13, a.idx, constindex!(a.tape, p))     # POWC
for (f, op) in ((:sin, 8), (:cos, 9), (:exp, 10), (:log, 11), (:sqrt, 12), (:tanh, 14))
    @eval Base.$f(a::Rec) = emit!(a.tape, $op, a.idx, 0)
end

"""
    record(f, n, m) -> Tape

Call `f(x, u)` (e.g. `(x, u) -> RD.evaluate(cost, x, u)`) on recording vectors and return the tape; the last instruction is the
value (scalar costs) or the last `p` instructions are the outputs (constraints: `f` returns a vector, each entry is re-emitted).
"""
function record(f, n::Integer, m::Integer)
    t = Tape()
    x = [emit!(t, 1, i - 1, 0) for i = 1:n]          # TO_OP_X
    u = [emit!(t, 2, j - 1, 0) for j = 1:m]          # TO_OP_U
    out = f(x, u)
    for o in (out isa AbstractVector ? out : (out,))
        o isa Rec ? emit!(t, 15, o.idx, constindex!(t, 0.0)) : emit!(t, 0, constindex!(t, o), 0)
    end
    return t
end
expr_cost_spec(t::Tape, terminal::Bool) =            # ToCostSpec of kind TO_COST_EXPR (keep `t` alive: GC roots)
    ToCostSpec(3, terminal ? 1 : 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0.0, 0.0, C_NULL, C_NULL,
               Int32(length(t.prog) ÷ 3), Int32(length(t.consts)), pointer(t.prog), pointer(t.consts))

# ---- the operator surface a solver calls (SURVEY.md 2.3) ------------------------------------------------------
# host arrays are Array{Float64,3}: X (n, N, B), U (m, N-1, B) -- exactly the library's instance-major layout.
TO.set_initial_state!(p::BatchedProblem, x0::Matrix{Float64}) =          # src/problem.jl:270
    check(p.h, ccall((:to_set_initial_state, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, x0))
TO.initial_controls!(p::BatchedProblem, U::Array{Float64,3}) =           # src/problem.jl:261
    check(p.h, ccall((:to_set_controls, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, U))
TO.initial_states!(p::BatchedProblem, X::Array{Float64,3}) =             # src/problem.jl:253
    check(p.h, ccall((:to_set_states, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, X))
TO.rollout!(p::BatchedProblem) =                                         # src/problem.jl:330-340
    check(p.h, ccall((:to_rollout, libb200), Cint, (Ptr{Cvoid},), p.h))
function TO.cost(p::BatchedProblem)                                      # src/problem.jl:321
    J = Vector{Float64}(undef, p.B)
    check(p.h, ccall((:to_cost, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, J)); J
end
function TO.states(p::BatchedProblem)                                    # src/problem.jl:175
    n, m, N = RD.dims(p.prob, 1); X = Array{Float64,3}(undef, n, N, p.B)
    check(p.h, ccall((:to_get_states, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, X)); X
end
function TO.controls(p::BatchedProblem)                                  # src/problem.jl:168
    n, m, N = RD.dims(p.prob, 1); U = Array{Float64,3}(undef, m, N - 1, p.B)
    check(p.h, ccall((:to_get_controls, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, U)); U
end
function TO.evaluate_constraints!(p::BatchedProblem, con_index::Integer, vals::Array{Float64,3})   # src/abstract_constraint.jl:200-225
    check(p.h, ccall((:to_eval_constraints, libb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), p.h, con_index - 1, vals)); vals
end
function TO.constraint_jacobians!(p::BatchedProblem, con_index::Integer, jac::Array{Float64,4})    # src/abstract_constraint.jl:236-248
    check(p.h, ccall((:to_constraint_jacobians, libb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), p.h, con_index - 1, jac)); jac
end

# ---- the same sweeps with the REFERENCE'S OWN SIGNATURES (src/abstract_constraint.jl:200-280), dispatching on a batched trajectory ----
# Altro calls   evaluate_constraints!(sig, con, vals, Z, inds) / constraint_jacobians!(sig, dif, con, jac, vals, Z, inds)
# with Z = get_trajectory(prob).  `get_trajectory(::BatchedProblem)` returns a BatchedTrajectory, so those calls land here unchanged; the
# output containers are 3- / 4-dimensional arrays (p, length(inds), B) / (p, n+m, length(inds), B) instead of vectors of vectors.
struct BatchedTrajectory
    p::BatchedProblem
end
TO.get_trajectory(p::BatchedProblem) = BatchedTrajectory(p)
function con_index(p::BatchedProblem, con, inds)
    for (i, (ii, c)) in enumerate(zip(TO.get_constraints(p.prob)))
        c === con && first(ii) == first(inds) && last(ii) == last(inds) && return i
    end
    throw(ArgumentError("constraint / knot range is not part of the batched problem's ConstraintList"))
end
TO.evaluate_constraints!(::RD.FunctionSignature, con::TO.StageConstraint, vals::Array{Float64,3}, Z::BatchedTrajectory, inds) =
    TO.evaluate_constraints!(Z.p, con_index(Z.p, con, inds), vals)
TO.constraint_jacobians!(::RD.FunctionSignature, ::RD.DiffMethod, con::TO.StageConstraint, jac::Array{Float64,4}, vals, Z::BatchedTrajectory, inds) =
    TO.constraint_jacobians!(Z.p, con_index(Z.p, con, inds), jac)
# second-order term: H[:, :, j, b] = d/dz (grad c' lambda) at knot inds[j] of instance b  (to_constraint_hessians; lambda = (p, length(inds), B) or nothing = the current multipliers)
function TO.∇constraint_jacobians!(::RD.FunctionSignature, ::RD.DiffMethod, con::TO.StageConstraint, H::Array{Float64,4}, λ, vals, Z::BatchedTrajectory, inds)
    lp = λ === nothing ? Ptr{Float64}(C_NULL) : pointer(λ)
    GC.@preserve λ check(Z.p.h, ccall((:to_constraint_hessians, libb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), Z.p.h, con_index(Z.p, con, inds) - 1, lp, H))
    H
end
# cost expansion of every knot of every instance: RD.gradient!(cost, grad, z) / RD.hessian!(cost, hess, z)  (src/cost_functions.jl:137-233)
function RD.gradient!(p::BatchedProblem, grad::Array{Float64,3})            # (n+m, N, B)
    check(p.h, ccall((:to_cost_gradient, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, grad)); grad
end
function RD.hessian!(p::BatchedProblem, hess::Array{Float64,4})             # (n+m, n+m, N, B), written symmetric
    check(p.h, ccall((:to_cost_hessian, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, hess)); hess
end
function TO.cost!(p::BatchedProblem, J::Matrix{Float64})                    # cost!(obj, Z) src/objective.jl:104-106: (N, B) knot costs
    check(p.h, ccall((:to_cost_knots, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, J)); J
end
# cones (src/cones.jl:71-276) on `count` vectors of length p at once: x, px (p, count); J, H (p, p, count)
TO.projection!(p::BatchedProblem, cone::TO.ConstraintSense, px::Matrix{Float64}, x::Matrix{Float64}) =
    (check(p.h, ccall((:to_projection, libb200), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}), p.h, sense_code(cone), size(x, 1), size(x, 2), x, px)); px)
TO.∇projection!(p::BatchedProblem, cone::TO.ConstraintSense, J::Array{Float64,3}, x::Matrix{Float64}) =
    (check(p.h, ccall((:to_grad_projection, libb200), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}), p.h, sense_code(cone), size(x, 1), size(x, 2), x, J)); J)
TO.∇²projection!(p::BatchedProblem, cone::TO.ConstraintSense, H::Array{Float64,3}, x::Matrix{Float64}, b::Matrix{Float64}) =
    (check(p.h, ccall((:to_hess_projection, libb200), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), p.h, sense_code(cone), size(x, 1), size(x, 2), x, b, H)); H)

TO.set_goal_state!(p::BatchedProblem, xf::Vector{Float64}; objective=true, constraint=true) =      # src/problem.jl:294-310
    check(p.h, ccall((:to_set_goal_state, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint), p.h, xf, objective, constraint))

# ---- what Altro.jl's iLQR / AL loop does with the API above, fused on the device ------------------------------
expand!(p::BatchedProblem) = check(p.h, ccall((:to_expand, libb200), Cint, (Ptr{Cvoid},), p.h))
backwardpass!(p::BatchedProblem) = check(p.h, ccall((:to_backward, libb200), Cint, (Ptr{Cvoid}, Ptr{Int32}), p.h, C_NULL))
forwardpass!(p::BatchedProblem) = check(p.h, ccall((:to_forward, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, C_NULL, C_NULL))
ilqr_step!(p::BatchedProblem, iters::Integer=1) = check(p.h, ccall((:to_ilqr_step, libb200), Cint, (Ptr{Cvoid}, Int32), p.h, iters))
al_update!(p::BatchedProblem) = check(p.h, ccall((:to_al_update, libb200), Cint, (Ptr{Cvoid},), p.h))
function max_violation(p::BatchedProblem)
    v = Vector{Float64}(undef, p.B)
    check(p.h, ccall((:to_max_violation, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, v)); v
end

# MPC plumbing: update_trajectory!(obj, Z, start) src/objective.jl:198-212 on the batched problem; Xref (n, nref), Uref (m, nref)
TO.update_trajectory!(p::BatchedProblem, Xref::Matrix{Float64}, Uref::Matrix{Float64}, start::Integer=1) =
    check(p.h, ccall((:to_update_trajectory, libb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32), p.h, Xref, Uref, size(Xref, 2), start))
shift_trajectory!(p::BatchedProblem, steps::Integer=1) = check(p.h, ccall((:to_shift_trajectory, libb200), Cint, (Ptr{Cvoid}, Int32), p.h, steps))

# multi-GPU (one process per GPU, e.g. under MPI.jl + NCCL.jl): the only collective is the {sum J, max violation} all-reduce.
# `to_reduce_merit_async` queues the per-GPU reduction behind the iteration in flight and makes `stream` (the CUDA.jl
# stream the NCCL call is issued on) wait for it; `merit_device_ptr` is the 2-double buffer to all-reduce in place.
reduce_merit_async!(p::BatchedProblem, stream::Ptr{Cvoid}) = check(p.h, ccall((:to_reduce_merit_async, libb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), p.h, stream))
function merit_device_ptr(p::BatchedProblem)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(p.h, ccall((:to_merit_device_ptr, libb200), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), p.h, r)); r[]
end

end # module
