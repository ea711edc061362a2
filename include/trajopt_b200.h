/* trajopt_b200.h -- C ABI of the B200-native batched problem-evaluation hot path of TrajectoryOptimization.jl.
 *
 * The reference (/root/reference, v0.7.1) is pure Julia and has no FFI: its "operator API" is the set of
 * Julia functions a solver (Altro.jl) calls on a `Problem`.  Each entry point below replaces one of those
 * calls for a BATCH of B independent problem instances that share model / objective / constraints and
 * differ in x0, X, U, multipliers.  Every function cites the reference interface it stands in for.  The
 * Julia-side binding (ccall) that a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all functions return 0 on success or a negative TO_E* code; to_last_error() gives the message
 *     (the analogue of the reference's DimensionMismatch / ArgumentError exceptions, src/problem.jl:64-68,87-91).
 *     Nothing throws across the ABI.
 *   - host arrays are caller-owned, instance-major, Julia column-major within an instance:
 *       X[B][N][n]  == Julia Array{Float64,3}(n, N, B)        U[B][N-1][m] == Array(m, N-1, B)
 *       matrices are column-major (Julia), knot ranges and indices are 1-based like the reference.
 *   - device memory is owned by the library behind the opaque handle; one handle <-> one GPU <-> one stream;
 *     a handle is not thread-safe.  Multi-GPU = one process/handle per device, batch sharded (no data-path
 *     collective; the global merit all-reduce runs on to_merit_device_ptr()).
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels or fails.
 */
#ifndef TRAJOPT_B200_H
#define TRAJOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TO_OK 0
#define TO_EINVAL (-1)   /* invalid argument (ArgumentError) */
#define TO_EDIM (-2)     /* dimension mismatch (DimensionMismatch, src/problem.jl:64-68, src/constraint_list.jl:108-110) */
#define TO_ECUDA (-3)    /* CUDA runtime failure */
#define TO_ENOMEM (-4)
#define TO_ESTATE (-5)   /* call order violated (e.g. backward pass before expansion) */
#define TO_ECONE (-6)    /* "Invalid second-order cone projection" (src/cones.jl:91,124) */

/* RobotZoo / example models on the path */
enum to_model_id {
    TO_MODEL_DOUBLE_INTEGRATOR = 0, /* examples/quickstart.jl:11-23 ; n = 2*dim, m = dim, params[0] = mass */
    TO_MODEL_CARTPOLE = 1,          /* docs/src/model.md:14-51 ; params = mc, mp, l, g */
    TO_MODEL_QUADROTOR = 2,         /* examples/Quadrotor.ipynb cells 4,8 ; params = mass, J1..3, g1..3, motor_dist, kf, km */
    TO_MODEL_ACROBOT = 3,           /* RobotZoo.Acrobot ; params = l1,l2,m1,m2,J1,J2,friction,g */
    TO_MODEL_EXPR = 4               /* user dynamics recorded as programs, one per knot allowed (hybrid / variable-dimension models, src/dynamics.jl:15-31,
                                       test/hybrid_dynamics_model.jl): to_spec.dyn, dyn_index, nx, nu */
};

/* QuadraticCostFunction (src/cost_functions.jl:326-347 DiagonalCost, :417-454 QuadraticCost) */
/* DiagonalQuatCost / QuatLQRCost (src/lie_costs.jl:33-95, :129-139): a DiagonalCost plus w * min(1 + q_ref'p, 1 - q_ref'p), p = x[q_ind] */
/* Generic (user-defined) cost: the reference lets users subtype CostFunction and get gradient!/hessian! from ForwardDiff through
 * RD.@autodiff (docs/src/costfunction_interface.md:30-50, test/nlcosts.jl:4-19).  Here the user function arrives as a straight-line
 * PROGRAM (SSA, one value per instruction, the last one is the cost) recorded by the host API, and the kernels evaluate it with
 * second-order forward-mode dual numbers.  Instruction i = {op, a, b}: operands are indices of earlier instructions,
 * of the state / control vector (TO_OP_X / TO_OP_U), or of the constant table (TO_OP_CONST; TO_OP_POWC: b).  At the terminal knot the
 * program is evaluated with u = 0 and only the state derivatives are taken (the reference's zero terminal control). */
enum to_cost_kind { TO_COST_DIAGONAL = 0, TO_COST_QUADRATIC = 1, TO_COST_DIAGONAL_QUAT = 2, TO_COST_EXPR = 3 };
enum to_expr_op { TO_OP_CONST = 0, TO_OP_X = 1, TO_OP_U = 2, TO_OP_ADD = 3, TO_OP_SUB = 4, TO_OP_MUL = 5, TO_OP_DIV = 6, TO_OP_NEG = 7,
                  TO_OP_SIN = 8, TO_OP_COS = 9, TO_OP_EXP = 10, TO_OP_LOG = 11, TO_OP_SQRT = 12, TO_OP_POWC = 13 /* a ^ const[b] */, TO_OP_TANH = 14,
                  /* one operand taken from the constant table: a (op) const[b] */
                  TO_OP_ADDC = 15, TO_OP_MULC = 16, TO_OP_DIVC = 17 /* a / c */, TO_OP_RDIVC = 18 /* c / a */, TO_OP_RSUBC = 19 /* c - a */ };
#define TO_EXPR_MAXLEN 128
#define TO_EXPR_MAXCONST 64
typedef struct {
    int32_t kind;     /* to_cost_kind */
    int32_t terminal; /* `terminal` flag of the cost (LQRObjective sets it on the last cost, src/objective.jl:154,180) */
    const double* Q;  /* DIAGONAL: n diagonal entries ; QUADRATIC: n*n column-major */
    const double* R;  /* DIAGONAL: m ; QUADRATIC: m*m column-major */
    const double* H;  /* QUADRATIC: m*n column-major cross term u'Hx, or NULL (zero) */
    const double* q;  /* n */
    const double* r;  /* m */
    double c;
    double w;               /* DIAGONAL_QUAT: weight of the geodesic term */
    const double* q_ref;    /* DIAGONAL_QUAT: reference quaternion (4, scalar first); else NULL */
    const int32_t* q_ind;   /* DIAGONAL_QUAT: 1-based state indices of the quaternion (4); NULL = 4:7 (src/lie_costs.jl:64) */
    int32_t prog_len;       /* EXPR: number of instructions (<= TO_EXPR_MAXLEN); Q, R, q, r may be NULL */
    int32_t nconst;         /* EXPR: number of constants (<= TO_EXPR_MAXCONST) */
    const int32_t* prog;    /* EXPR: prog_len x {op, a, b} */
    const double* consts;   /* EXPR */
} to_cost_spec;

/* One dynamics model of a hybrid problem: `RD.@autodiff struct M <: ContinuousDynamics` + `RD.dynamics(::M, x, u)` recorded as a program
 * (to_expr_op; inputs x[0..n_in), u[0..m_in); the LAST n_out instructions are the outputs), discretised with RK4 like every model of the path,
 * or -- `discrete` = 1 -- a jump map x+ = g(x, u) applied as is (n_out may differ from n_in: the state dimension changes there). */
typedef struct {
    int32_t n_in, m_in, n_out, discrete;
    int32_t prog_len, nconst;
    const int32_t* prog;
    const double* consts;
} to_dynamics_spec;

/* ConstraintSense (src/cones.jl:17-61) */
enum to_cone { TO_CONE_ZERO = 0 /* Equality */, TO_CONE_NEGATIVE_ORTHANT = 1 /* Inequality */, TO_CONE_SECOND_ORDER = 2,
               TO_CONE_IDENTITY = 3, TO_CONE_POSITIVE_ORTHANT = 4 };

/* AbstractConstraint subtypes (src/constraints.jl) */
enum to_con_kind {
    TO_CON_GOAL = 0,   /* GoalConstraint :22-87      a = xf[ninds], inds = 1-based state indices             */
    TO_CON_BOUND = 1,  /* BoundConstraint :644-783   a = z_max[n+m], b = z_min[n+m] (+-Inf = unbounded)        */
    TO_CON_LINEAR = 2, /* LinearConstraint :103-150  a = A[p x w] col-major, b = b[p], flag = 0 state | 1 control, sense */
    TO_CON_CIRCLE = 3, /* CircleConstraint :168-233  a = xc[p], b = yc[p], rad = r[p], inds = {xi, yi} (1-based) */
    TO_CON_SPHERE = 4, /* SphereConstraint :249-326  a,b,c = centers, rad, inds = {xi, yi, zi}                 */
    TO_CON_NORM = 5,   /* NormConstraint :438-521    val, inds = 1-based indices into z, sense (orthant | SOC)   */
    TO_CON_COLLISION = 6, /* CollisionConstraint :341-389  val = radius, inds = {x1[D], x2[D]} 1-based state indices (ninds = 2D): r^2 - |x[x1]-x[x2]|^2 <= 0.
                            StateBound / ControlBound :547-631 are TO_CON_BOUND with the other block unbounded. */
    TO_CON_EXPR = 8,     /* user constraint recorded as a program (RD.@autodiff struct ... <: StageConstraint, docs/src/constraint_interface.md:52-72):
                            inds = prog_len x {op, a, b} (ninds = 3 prog_len, to_expr_op as for TO_COST_EXPR), a = constants (flag = their number),
                            p = output dimension: the values of the LAST p instructions; sense; Jacobian by forward-mode AD on the device */
    TO_CON_QUATVEC = 7   /* QuatVecEq :938-965  a = qf (4, scalar first), inds = qind (4, 1-based; NULL = 4:7): with q = normalize(x[qind]) and
                            qf flipped when qf'q < 0, c = q[2:4] - qf[2:4]; Equality, p = 3 */
};
typedef struct {
    int32_t kind;        /* to_con_kind */
    int32_t first, last; /* knot range first:last, 1-based inclusive (add_constraint!, src/constraint_list.jl:103-134) */
    int32_t sense;       /* to_cone; ignored for GOAL (Equality) and BOUND/CIRCLE/SPHERE (Inequality) */
    int32_t p;           /* rows for LINEAR / number of obstacles for CIRCLE, SPHERE; ignored otherwise */
    int32_t flag;
    int32_t ninds;
    const int32_t* inds;
    const double* a;
    const double* b;
    const double* c;
    const double* rad;
    double val;
} to_constraint_spec;

/* Problem(model, obj, x0, tf; constraints, ...)  src/problem.jl:79-123 */
typedef struct {
    int32_t model;           /* to_model_id */
    int32_t n, m;            /* state / control dimension (checked against the model, src/problem.jl:64-68) */
    int32_t N;               /* knot points == length(obj) (src/problem.jl:95) */
    int32_t B;               /* batch: independent problem instances on this device */
    int32_t device;          /* CUDA device ordinal */
    int32_t nparams;
    const double* params;    /* model parameters, NULL = the model's defaults */
    const double* dt;        /* N-1 time steps (vector dt allowed, test/problems_tests.jl:79-82) */
    double t0;
    int32_t ncost;           /* distinct cost functions */
    const to_cost_spec* costs;
    const int32_t* cost_index; /* N entries, 0-based index into costs (Objective.cost, src/objective.jl:27-45) */
    int32_t ncon;
    const to_constraint_spec* cons; /* ConstraintList, in add_constraint! order */
    int32_t error_state;     /* 1: the solver kernels (backward / forward pass) work on the ERROR STATE of a Lie-group model, as Altro does when
                                RD.errstate_dim(model) != n: the Quadrotor's quaternion (x[4:7]) contributes 3 dimensions, n_e = 12.  State-difference
                                Jacobian G(x) = blkdiag(I3, L(q) H, I6) (Rotations.jl grad-differential), dynamics A_e = G_{k+1}' A G_k, B_e = G_{k+1}' B,
                                cost expansion G'lxx G + grad^2-differential, dx = state_diff(xbar, x) with the Cayley map.  The reference's hooks for
                                it: src/abstract_constraint.jl:282-303 (error_expansion! of constraint Jacobians), src/lie_costs.jl.  0: full state. */
    /* TO_MODEL_EXPR only (else 0 / NULL): `Problem(models::Vector{<:DiscreteDynamics}, ...)`, src/problem.jl:36-73 with RD.dims(models), src/dynamics.jl:15-31.
       n, m are the LARGEST state / control dimensions; knot k has nx[k] <= n states and nu[k] <= m controls, stored in the first entries of the
       n- / m-sized slots (the rest stays zero: costs and constraints are described on the padded [x(n); u(m)] layout, with unit weights on the
       unused controls so that Quu stays positive definite).  Model dyn[dyn_index[k]] maps knot k to k+1: n_in = nx[k], m_in = nu[k], n_out = nx[k+1]
       (checked: the reference's DimensionMismatch "Model mismatch at time step k"). */
    int32_t ndyn;
    const to_dynamics_spec* dyn;
    const int32_t* dyn_index;   /* N-1 entries, 0-based */
    const int32_t* nx;          /* N entries */
    const int32_t* nu;          /* N entries (nu[N-1] = nu[N-2], as RD.dims does) */
} to_spec;

/* Solver options on the path (Altro.jl SolverOptions, restated in oracle/oracle.hpp `Options`) */
typedef struct {
    double bp_reg_increase_factor, bp_reg_max, bp_reg_min, bp_reg_initial, bp_reg_fp;
    double line_search_lower_bound, line_search_upper_bound;
    int32_t iterations_linesearch;
    int32_t backward_kernel;   /* 0 = automatic; 1 = warp-per-instance Riccati kernel; 2 = thread-per-instance kernel (n <= 4, m <= 2,
                                  Goal/Bound constraints only, else ignored); error-state problems: 3 = generic DFMA kernel on the full
                                  materialised expansion, 5 = shared-memory tensor kernel on the compact expansion (automatic = the
                                  register-resident fragment kernel when the problem is compact). Not a solver option of the reference:
                                  a tuning / test knob. */
    double max_state_value, max_control_value;
    double penalty_initial, penalty_scaling, penalty_max, dual_max;
} to_options;

typedef struct to_handle to_handle;

/* ---- lifecycle -------------------------------------------------------------------------------------- */
int to_create(const to_spec* spec, to_handle** out);                 /* Problem(...)           src/problem.jl:79-111 */
int to_destroy(to_handle* h);
const char* to_last_error(const to_handle* h);                       /* h may be NULL: error of the last failed to_create */
int to_default_options(to_options* o);
int to_set_options(to_handle* h, const to_options* o);
int to_set_stream(to_handle* h, void* cuda_stream);                  /* run on the caller's stream (e.g. torch's current stream) */
int to_synchronize(to_handle* h);
int to_dims(const to_handle* h, int32_t* n, int32_t* m, int32_t* N, int32_t* B); /* RD.dims(prob)  src/problem.jl:139-147 */
int to_num_constraints(const to_handle* h, int32_t* p_per_knot /*[N]*/);         /* num_constraints(prob) src/problem.jl:206 */
int to_constraint_info(const to_handle* h, int32_t con, int32_t* p, int32_t* sense, int32_t* first, int32_t* last);
int to_bounds(const to_handle* h, int32_t con, double* lower /*[p]*/, double* upper /*[p]*/); /* lower_bound/upper_bound src/abstract_constraint.jl:97-123 */

/* ---- setters / getters (host arrays, instance-major) ------------------------------------------------ */
/* The setters enqueue their host-to-device copy on the handle's stream and return: the host buffer must stay valid (and unchanged) until the next
 * synchronising call on the handle -- to_synchronize or any getter.  Getters copy back and synchronise the handle's stream before returning. */
int to_set_initial_state(to_handle* h, const double* x0 /*[B][n]*/);          /* set_initial_state! src/problem.jl:270 */
int to_set_controls(to_handle* h, const double* U /*[B][N-1][m]*/);           /* initial_controls!  src/problem.jl:261 */
int to_set_states(to_handle* h, const double* X /*[B][N][n]*/);               /* initial_states!    src/problem.jl:253 */
int to_set_goal_state(to_handle* h, const double* xf /*[n]*/, int objective, int constraint); /* set_goal_state! src/problem.jl:294-310 */
int to_set_initial_time(to_handle* h, double t0, double* tf_out);             /* setinitialtime!    src/problem.jl:280 */
int to_get_states(to_handle* h, double* X /*[B][N][n]*/);                     /* states(prob)       src/problem.jl:175 */
int to_get_controls(to_handle* h, double* U /*[B][N-1][m]*/);                 /* controls(prob)     src/problem.jl:168 */
int to_get_times(to_handle* h, double* t /*[N]*/);                            /* gettimes(prob)     src/problem.jl:182 */

/* ---- MPC plumbing (BASELINE config 5) ---------------------------------------------------------
 * update_trajectory!(obj, Z, start) src/objective.jl:207-212: knot i = 1..N of a tracking objective follows row
 * (start-1+i) of the reference: set_LQR_goal! (src/cost_functions.jl:245-254) q = -Q xf, r = -R uf, c untouched.
 * Xref [nref][n], Uref [nref][m] host arrays (start + N - 1 <= nref). Costs shared by several knots end up tracking
 * the last of them, exactly as the reference's aliased cost objects do. */
int to_update_trajectory(to_handle* h, const double* Xref, const double* Uref, int32_t nref, int32_t start);
/* Receding-horizon warm start, on the device: X_k <- X_{k+steps}, U_k <- U_{k+steps} (the tail repeats the last
 * control and state), multipliers move with their knots (the tail keeps its last value), x0 <- X_{1+steps},
 * t0 += the skipped dt. The caller then sets the measured state (to_set_initial_state) and re-rolls out. */
int to_shift_trajectory(to_handle* h, int32_t steps);

/* ---- kernel 1: batched RK4 rollout (+ dual-number Jacobians) ---------------------------------------- */
int to_rollout(to_handle* h);                                                 /* rollout!           src/problem.jl:330-340 */
int to_expand(to_handle* h);                                                  /* RD.jacobian!(ForwardAD) on the discretized dynamics at every knot */
int to_get_dynamics_jacobians(to_handle* h, double* AB /*[B][N-1][n+m][n]: n x (n+m) col-major*/);

/* ---- kernel 2: cost + constraint + AL sweep --------------------------------------------------------- */
int to_cost(to_handle* h, double* J /*[B]*/);                                 /* cost(prob)         src/problem.jl:321, src/objective.jl:89-93 */
int to_cost_knots(to_handle* h, double* Jk /*[B][N]*/);                       /* cost! / get_J      src/objective.jl:104-110 */
int to_cost_gradient(to_handle* h, double* grad /*[B][N][n+m]*/);             /* RD.gradient!       src/cost_functions.jl:137-172 */
int to_cost_hessian(to_handle* h, double* hess /*[B][N][n+m][n+m]*/);         /* RD.hessian!        src/cost_functions.jl:212-233 (written symmetric) */
int to_eval_constraints(to_handle* h, int32_t con, double* vals /*[B][last-first+1][p]*/);      /* evaluate_constraints! src/abstract_constraint.jl:200-225 */
int to_constraint_jacobians(to_handle* h, int32_t con, double* jac /*[B][last-first+1][n+m][p]*/); /* constraint_jacobians! src/abstract_constraint.jl:236-248 */
int to_constraint_hessians(to_handle* h, int32_t con, const double* lambda /*[B][last-first+1][p], NULL = the current multipliers*/,
                           double* H /*[B][last-first+1][n+m][n+m]*/);   /* grad-constraint_jacobians! src/abstract_constraint.jl:267-280: d/dz (cz' lambda) of every knot
                                                                            (`∇jacobian!`: zero for Goal / Bound, src/constraints.jl:70-73, :767-770; second-order AD otherwise) */
int to_max_violation(to_handle* h, double* v /*[B]*/);
int to_merit(to_handle* h, double* J /*[B]*/);                                /* cost + AL penalty of the current trajectory */
int to_al_expansion(to_handle* h, double* grad /*[B][N][n+m]*/, double* hess /*[B][N][n+m][n+m]*/); /* cost expansion incl. AL terms */

/* cones (stand-alone, batched over `count` vectors of length p)  src/cones.jl */
int to_projection(to_handle* h, int32_t cone, int32_t p, int32_t count, const double* x, double* px);      /* projection!  :96-127 */
int to_grad_projection(to_handle* h, int32_t cone, int32_t p, int32_t count, const double* x, double* J);  /* grad-projection! :129-188, p x p col-major each */
int to_hess_projection(to_handle* h, int32_t cone, int32_t p, int32_t count, const double* x, const double* b, double* H); /* :201-276 */

/* ---- kernel 3 + forward pass (what Altro.jl's iLQR does with the API above) --------------------------- */
int to_backward(to_handle* h, int32_t* status /*[B] or NULL*/);               /* Riccati backward pass (needs to_expand) */
int to_forward(to_handle* h, double* J /*[B] or NULL*/, double* alpha /*[B] or NULL*/); /* closed-loop rollout + line search */
int to_ilqr_step(to_handle* h, int32_t iters);                                /* iters x (expand, backward, forward), no host sync */
int to_al_update(to_handle* h);                                               /* dual + penalty update */
int to_get_gains(to_handle* h, double* K /*[B][N-1][n_e][m]: m x n_e col-major (n_e = n unless error_state)*/, double* d /*[B][N-1][m]*/);
/* ---- Lie-group error state ------------------------------------------------------------------- */
int to_backward_algebra(const to_handle* h, int32_t* variant);             /* which arithmetic the next to_backward will use: 0 = pivot-by-pivot LDL' solve
                                                                             (riccati.cu, riccati_small.cu, lie.cu), 1 = 2 x 2 block inverse + W'K update
                                                                             (riccati_frag.cu).  Same mathematics (Altro backwardpass!); the oracle mirrors
                                                                             either so that parity tests compare like with like */
int to_error_state_dim(const to_handle* h, int32_t* ne);                      /* RD.errstate_dim(model): n, or n - 1 with spec.error_state */
/* RD.state_diff(model, xbar, x) of every knot against the current trajectory: Xbar [B][N][n] (host) -> dx [B][N][n_e] */
int to_state_diff(to_handle* h, const double* Xbar, double* dx);
/* error-state dynamics Jacobians [A_e B_e] = G_{k+1}' [A G_k | B] after to_expand: [B][N-1][n_e+m][n_e], n_e x (n_e+m) col-major */
int to_get_error_dynamics(to_handle* h, double* ABe);
/* error-state cost + AL expansion of every knot (Altro error_expansion!): grad [B][N][n_e+m], hess [B][N][n_e+m][n_e+m] */
int to_error_expansion(to_handle* h, double* grad, double* hess);
int to_get_multipliers(to_handle* h, int32_t con, double* lambda /*[B][last-first+1][p]*/);
int to_set_multipliers(to_handle* h, int32_t con, const double* lambda);
int to_get_penalty(to_handle* h, int32_t con, double* mu);
int to_set_penalty(to_handle* h, int32_t con, double mu);
int to_get_solver_state(to_handle* h, double* rho /*[B]*/, double* dV /*[B][2]*/, double* alpha /*[B]*/, int32_t* ls_iters /*[B]*/, int32_t* bp_status /*[B]*/);

/* ---- multi-GPU / measurement plumbing --------------------------------------------------------------- */
/* device pointer to {sum_b J_b, max_b violation_b} (2 doubles) refreshed by to_reduce_merit(); the host
 * framework all-reduces it (NCCL: sum on [0], max on [1]).  SURVEY 8(e). */
int to_reduce_merit(to_handle* h);
/* Same reduction, ordered after whatever part of the last iteration is still in flight on the library's side
 * stream, and handed to `consumer_stream` (a cudaStream_t, e.g. the stream the NCCL all-reduce is issued on) through
 * an event: the handle's main stream is NOT made to wait, so the next iteration keeps overlapping. */
int to_reduce_merit_async(to_handle* h, void* consumer_stream);
int to_merit_device_ptr(to_handle* h, void** ptr);
/* per-phase device timing (CUDA events on the handle's stream) for the roofline report */
enum to_phase { TO_PHASE_EXPAND = 0, TO_PHASE_BACKWARD = 1, TO_PHASE_FORWARD = 2, TO_PHASE_LADDER = 3, TO_PHASE_ACCEPT = 4, TO_PHASE_COSTEXP = 5 /* cost + AL expansion kernel of the record / materialised-expansion paths, when it is a launch of its own */,
               TO_PHASE_LATE = 6 /* overlapped iterations: dynamics + cost expansion of the instances the late line-search trials moved (side stream) */, TO_PHASE_COUNT = 8 };
int to_set_phase_timing(to_handle* h, int enable);
int to_get_phase_times(to_handle* h, double* ms /*[TO_PHASE_COUNT] accumulated*/, int64_t* launches /*[TO_PHASE_COUNT]*/, int reset);
int64_t to_launch_count(const to_handle* h);                                  /* kernels launched by this handle so far */
int to_algorithmic_bytes(const to_handle* h, int64_t* E, int64_t* R, int64_t* F); /* per instance per iteration, SURVEY 8(d) */

#ifdef __cplusplus
}
#endif
#endif
