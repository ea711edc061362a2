// oracle/oracle.hpp -- TEST INFRASTRUCTURE ONLY.  CPU (fp64) restatement of the hot path of
// TrajectoryOptimization.jl v0.7.1 + the solver pieces Altro.jl drives through it.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it; the product
// (trajectoryoptimization.jl_b200/csrc) never includes, links or calls anything in this directory.
//
// PARITY STATUS
//   pinned   : quadratic cost value/gradient/Hessian, LQRObjective parameters, Bound/Goal constraint
//              values + Jacobians, orthant/zero/SOC cone projections -- every closed form and literal
//              KAT the reference's tests hold for them is re-expressed in tests/test_oracle_*.py
//              (test/cost_tests.jl:238-279, test/objective_tests.jl:86-140,
//              test/constraint_tests.jl:17-39,209-344, test/cone_tests.jl:26-75, examples/quickstart.jl:71-137).
//              Also pinned: QuatVecEq value + Jacobian (test/constraint_tests.jl:412-444), user costs through AD
//              (test/nlcosts.jl:22-45), the docs' ControlNorm constraint (docs/src/constraint_interface.md:52-72),
//              IndexedConstraint (test/constraint_tests.jl:346-407, host side), DiagonalQuatCost closed forms
//              (src/lie_costs.jl:68-95; its test file test/quatcosts.jl is stale and not run by the reference).
//   UNPINNED : RK4, dual-number dynamics Jacobians, Riccati backward pass, forward line search, AL update,
//              and the Lie-group error state (state_diff, G, error_expansion: RobotDynamics / Rotations / Altro).
//              Their arithmetic lives in RobotDynamics.jl 0.4.8 / ForwardDiff 0.10 / RobotZoo 0.3 /
//              Altro.jl 0.3-0.5, none of which is vendored under /root/reference and none of which can
//              run here (no Julia).  They are restated from the published algorithms; independent checks
//              in tests/: RK4 vs scipy, Jacobians vs central differences, one iLQR step on an LQ problem vs
//              a dense KKT solve (numpy), Cartpole iLQR converging to the notebook's recorded cost
//              (examples/Cartpole.ipynb:378-382).  "parity unpinned" for those rows.
//
// Layouts (shared with the C ABI in include/trajopt_b200.h): instance-major, Julia column-major inside
// a knot:  X[B][N][n], U[B][N-1][m], AB[B][N-1][(n+m)][n] (n x (n+m) col-major), K[B][N-1][n][m]
// (m x n col-major), d[B][N-1][m].
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "models.hpp"

namespace oracle {

constexpr int MAXP = 32;   // rows of one constraint at one knot
static const double ZERO_U[MAXM] = {0};

// ------------------------------------------------------------------------------------------------
// Quadratic cost  (src/cost_functions.jl)
struct Cost {
    int n = 0, m = 0;
    bool diag = true;       // DiagonalCost (:326-347) vs QuadraticCost (:417-454)
    bool terminal = false;
    bool zeroH = true;      // is_blockdiag (:455, :382)
    std::vector<double> Q;  // n*n col-major (dense storage also for diagonal costs)
    std::vector<double> R;  // m*m col-major
    std::vector<double> H;  // m*n col-major
    std::vector<double> q, r;
    double c = 0;
    // DiagonalQuatCost (src/lie_costs.jl:33-56): + w * min(1 + q_ref'p, 1 - q_ref'p), p = x[q_ind]
    bool quat = false;
    double w = 0;
    double q_ref[4] = {1, 0, 0, 0};
    int q_ind[4] = {3, 4, 5, 6};   // 0-based
    // generic cost recorded as a straight-line program (RD.@autodiff user costs, docs/src/costfunction_interface.md:30-50)
    bool expr = false;
    std::vector<int> prog;          // {op, a, b} per instruction
    std::vector<double> consts;
};

// ---- straight-line program evaluation (to_expr_op in include/trajopt_b200.h) ----------------------------------------
// second-order forward mode: (v, d1, d2, d12) = value, d/ds, d/dt, d2/dsdt for seeds s, t (ForwardDiff's nested Duals restated)
struct Hyper {
    double v = 0, d1 = 0, d2 = 0, d12 = 0;
};
inline Hyper hyp_unary(const Hyper& a, double f, double f1, double f2) {
    Hyper r; r.v = f; r.d1 = f1 * a.d1; r.d2 = f1 * a.d2; r.d12 = f1 * a.d12 + f2 * a.d1 * a.d2; return r;
}
inline Hyper hyp_mul(const Hyper& a, const Hyper& b) {
    Hyper r; r.v = a.v * b.v; r.d1 = a.d1 * b.v + a.v * b.d1; r.d2 = a.d2 * b.v + a.v * b.d2;
    r.d12 = a.d12 * b.v + a.d1 * b.d2 + a.d2 * b.d1 + a.v * b.d12; return r;
}
enum ExprOp { OP_CONST = 0, OP_X = 1, OP_U = 2, OP_ADD = 3, OP_SUB = 4, OP_MUL = 5, OP_DIV = 6, OP_NEG = 7, OP_SIN = 8, OP_COS = 9,
              OP_EXP = 10, OP_LOG = 11, OP_SQRT = 12, OP_POWC = 13, OP_TANH = 14, OP_ADDC = 15, OP_MULC = 16, OP_DIVC = 17, OP_RDIVC = 18, OP_RSUBC = 19 };
constexpr int EXPR_MAXLEN = 128;
// evaluate the program with z_s1 seeded in d1 and z_s2 in d2 (index into [x;u]; -1 = no seed); has_u = false -> u = 0
// run a program: fills reg[0..L-1]
inline void expr_run(const int* prog, int L, const double* consts, int n, const double* x, const double* u, bool has_u, int s1, int s2, Hyper* reg) {
    for (int i = 0; i < L; i++) {
        const int op = prog[3 * i], a = prog[3 * i + 1], b = prog[3 * i + 2];
        Hyper r;
        switch (op) {
            case OP_CONST: r.v = consts[a]; break;
            case OP_X: r.v = x[a]; r.d1 = (a == s1); r.d2 = (a == s2); break;
            case OP_U: r.v = has_u ? u[a] : 0.0; r.d1 = (n + a == s1); r.d2 = (n + a == s2); break;
            case OP_ADD: r.v = reg[a].v + reg[b].v; r.d1 = reg[a].d1 + reg[b].d1; r.d2 = reg[a].d2 + reg[b].d2; r.d12 = reg[a].d12 + reg[b].d12; break;
            case OP_SUB: r.v = reg[a].v - reg[b].v; r.d1 = reg[a].d1 - reg[b].d1; r.d2 = reg[a].d2 - reg[b].d2; r.d12 = reg[a].d12 - reg[b].d12; break;
            case OP_MUL: r = hyp_mul(reg[a], reg[b]); break;
            case OP_DIV: { const double iv = 1.0 / reg[b].v; r = hyp_mul(reg[a], hyp_unary(reg[b], iv, -iv * iv, 2 * iv * iv * iv)); break; }
            case OP_NEG: r.v = -reg[a].v; r.d1 = -reg[a].d1; r.d2 = -reg[a].d2; r.d12 = -reg[a].d12; break;
            case OP_SIN: { const double sv = std::sin(reg[a].v), cv = std::cos(reg[a].v); r = hyp_unary(reg[a], sv, cv, -sv); break; }
            case OP_COS: { const double sv = std::sin(reg[a].v), cv = std::cos(reg[a].v); r = hyp_unary(reg[a], cv, -sv, -cv); break; }
            case OP_EXP: { const double e = std::exp(reg[a].v); r = hyp_unary(reg[a], e, e, e); break; }
            case OP_LOG: { const double iv = 1.0 / reg[a].v; r = hyp_unary(reg[a], std::log(reg[a].v), iv, -iv * iv); break; }
            case OP_SQRT: { const double sq = std::sqrt(reg[a].v); r = hyp_unary(reg[a], sq, 0.5 / sq, -0.25 / (sq * reg[a].v)); break; }
            case OP_POWC: { const double e = consts[b], v = reg[a].v; r = hyp_unary(reg[a], std::pow(v, e), e * std::pow(v, e - 1), e * (e - 1) * std::pow(v, e - 2)); break; }
            case OP_TANH: { const double t = std::tanh(reg[a].v); r = hyp_unary(reg[a], t, 1 - t * t, -2 * t * (1 - t * t)); break; }
            case OP_ADDC: r = reg[a]; r.v += consts[b]; break;
            case OP_MULC: { const double k = consts[b]; r.v = reg[a].v * k; r.d1 = reg[a].d1 * k; r.d2 = reg[a].d2 * k; r.d12 = reg[a].d12 * k; break; }
            case OP_DIVC: { const double k = consts[b]; r.v = reg[a].v / k; r.d1 = reg[a].d1 / k; r.d2 = reg[a].d2 / k; r.d12 = reg[a].d12 / k; break; }
            case OP_RDIVC: { const double k = consts[b], iv = 1.0 / reg[a].v; r = hyp_unary(reg[a], k * iv, -k * iv * iv, 2 * k * iv * iv * iv); break; }
            case OP_RSUBC: r.v = consts[b] - reg[a].v; r.d1 = -reg[a].d1; r.d2 = -reg[a].d2; r.d12 = -reg[a].d12; break;
        }
        reg[i] = r;
    }
}
inline Hyper expr_eval(const Cost& c, const double* x, const double* u, bool has_u, int s1, int s2) {
    Hyper reg[EXPR_MAXLEN];
    const int L = (int)c.prog.size() / 3;
    expr_run(c.prog.data(), L, c.consts.data(), c.n, x, u, has_u, s1, s2, reg);
    return reg[L - 1];
}

// RD.evaluate(::QuadraticCostFunction, x, u)  src/cost_functions.jl:89-104.  `has_u == false` is the
// `isempty(u)` branch; the batched layout has no control at the terminal knot, which equals the
// reference's convention of a zero terminal control (test/objective_tests.jl:128).
inline double cost_value(const Cost& c, const double* x, const double* u, bool has_u) {
    const int n = c.n, m = c.m;
    if (c.expr) return expr_eval(c, x, u, has_u, -1, -1).v;
    double J = 0;
    for (int j = 0; j < n; j++) {
        double qx = 0;
        for (int i = 0; i < n; i++) qx += c.Q[j * n + i] * x[i];  // (x'Q)_j
        J += 0.5 * qx * x[j];
    }
    double lin = 0;
    for (int i = 0; i < n; i++) lin += c.q[i] * x[i];
    J += lin + c.c;
    if (has_u) {
        double Ju = 0;
        for (int j = 0; j < m; j++) {
            double ru = 0;
            for (int i = 0; i < m; i++) ru += c.R[j * m + i] * u[i];
            Ju += 0.5 * ru * u[j];
        }
        double linu = 0;
        for (int i = 0; i < m; i++) linu += c.r[i] * u[i];
        J += Ju + linu;
        if (!c.zeroH) {
            double h = 0;
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) h += u[i] * c.H[j * m + i] * x[j];
            J += h;
        }
    }
    if (c.quat) {   // RD.evaluate(::DiagonalQuatCost)  src/lie_costs.jl:68-77
        double dq = 0;
        for (int i = 0; i < 4; i++) dq += c.q_ref[i] * x[c.q_ind[i]];
        J += c.w * std::min(1 + dq, 1 - dq);
    }
    return J;
}

// RD.gradient!  src/cost_functions.jl:137-172: grad = [Qx+q (+H'u); Ru+r (+Hx)], u-part untouched at terminal.
inline void cost_gradient(const Cost& c, const double* x, const double* u, bool is_terminal, double* grad) {
    const int n = c.n, m = c.m;
    if (c.expr) {   // RD.gradient!(ForwardAD) of a user cost; the control part is left untouched at the terminal knot
        for (int i = 0; i < (is_terminal ? n : n + m); i++) grad[i] = expr_eval(c, x, u, !is_terminal, i, -1).d1;
        return;
    }
    for (int i = 0; i < n; i++) {
        double g = c.q[i];
        for (int j = 0; j < n; j++) g += c.Q[j * n + i] * x[j];
        grad[i] = g;
    }
    if (c.quat) {   // gradient!(::DiagonalQuatCost)  src/lie_costs.jl:79-95: Qx -+ w Iq q_ref by the sign of q_ref'p
        double dq = 0;
        for (int i = 0; i < 4; i++) dq += c.q_ref[i] * x[c.q_ind[i]];
        for (int i = 0; i < 4; i++) grad[c.q_ind[i]] += (dq < 0 ? c.w : -c.w) * c.q_ref[i];
    }
    if (!is_terminal) {
        for (int i = 0; i < m; i++) {
            double g = c.r[i];
            for (int j = 0; j < m; j++) g += c.R[j * m + i] * u[j];
            grad[n + i] = g;
        }
        if (!c.zeroH) {
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) {
                    grad[j] += c.H[j * m + i] * u[i];
                    grad[n + i] += c.H[j * m + i] * x[j];
                }
        }
    }
}

// RD.hessian!  src/cost_functions.jl:212-233.  hess is (n+m)x(n+m) col-major.  Diagonal costs zero the
// matrix first (:216); dense costs only write blocks Q, R and the lower-left H -- here the
// caller passes a zeroed matrix, and `symmetric` additionally mirrors H' into the upper-right block
// (what a solver consumes).
inline void cost_hessian(const Cost& c, const double* x, const double* u, bool is_terminal, double* hess, bool symmetric) {
    const int n = c.n, m = c.m, nm = n + m;
    if (c.expr) {   // RD.hessian!(ForwardAD): full symmetric (n+m)^2 block (state block only at the terminal knot)
        std::fill(hess, hess + nm * nm, 0.0);
        const int lim = is_terminal ? n : nm;
        for (int j = 0; j < lim; j++)
            for (int i = j; i < lim; i++) { const double h = expr_eval(c, x, u, !is_terminal, i, j).d12; hess[j * nm + i] = h; hess[i * nm + j] = h; }
        return;
    }
    if (c.diag) std::fill(hess, hess + nm * nm, 0.0);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
            if (!c.diag || i == j) hess[j * nm + i] = c.Q[j * n + i];
    if (!is_terminal) {
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++)
                if (!c.diag || i == j) hess[(n + j) * nm + (n + i)] = c.R[j * m + i];
        if (!c.zeroH) {
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) {
                    hess[j * nm + (n + i)] = c.H[j * m + i];
                    if (symmetric) hess[(n + i) * nm + j] = c.H[j * m + i];
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cones  (src/cones.jl)
enum Cone { CONE_ZERO = 0, CONE_NEGATIVE_ORTHANT = 1, CONE_SECOND_ORDER = 2, CONE_IDENTITY = 3, CONE_POSITIVE_ORTHANT = 4 };

inline int dualcone(int cone) {  // src/cones.jl:65-69
    switch (cone) {
        case CONE_IDENTITY: return CONE_ZERO;
        case CONE_ZERO: return CONE_IDENTITY;
        default: return cone;
    }
}

// projection!  src/cones.jl:96-127.  returns 0, or -1 for the "Invalid second-order cone projection" error (:124)
inline int projection(int cone, const double* x, int p, double* px) {
    switch (cone) {
        case CONE_IDENTITY: for (int i = 0; i < p; i++) px[i] = x[i]; return 0;
        case CONE_ZERO: for (int i = 0; i < p; i++) px[i] = 0; return 0;
        case CONE_NEGATIVE_ORTHANT: for (int i = 0; i < p; i++) px[i] = std::min(0.0, x[i]); return 0;
        case CONE_POSITIVE_ORTHANT: for (int i = 0; i < p; i++) px[i] = std::max(0.0, x[i]); return 0;
        case CONE_SECOND_ORDER: {
            double s = x[p - 1], a = 0;
            for (int i = 0; i < p - 1; i++) a += x[i] * x[i];
            a = std::sqrt(a);
            if (a <= -s) { for (int i = 0; i < p; i++) px[i] = 0; }
            else if (a <= s) { for (int i = 0; i < p; i++) px[i] = x[i]; }
            else if (a >= std::fabs(s)) {
                double sc = 0.5 * (1 + s / a);
                for (int i = 0; i < p - 1; i++) px[i] = sc * x[i];
                px[p - 1] = sc * a;
            } else return -1;
            return 0;
        }
    }
    return -1;
}

// grad-projection!  src/cones.jl:129-188.  J is p x p col-major.  Orthant: only the diagonal is written by
// the reference (:138-145); here J is fully defined (off-diagonals zero).
inline int grad_projection(int cone, const double* x, int p, double* J) {
    std::fill(J, J + p * p, 0.0);
    switch (cone) {
        case CONE_IDENTITY: for (int i = 0; i < p; i++) J[i * p + i] = 1; return 0;
        case CONE_ZERO: return 0;
        case CONE_NEGATIVE_ORTHANT: for (int i = 0; i < p; i++) J[i * p + i] = x[i] <= 0 ? 1 : 0; return 0;
        case CONE_POSITIVE_ORTHANT: for (int i = 0; i < p; i++) J[i * p + i] = x[i] >= 0 ? 1 : 0; return 0;
        case CONE_SECOND_ORDER: {
            const int n = p;
            double s = x[n - 1], a = 0;
            for (int i = 0; i < n - 1; i++) a += x[i] * x[i];
            a = std::sqrt(a);
            if (a <= -s) return 0;
            if (a <= s) { for (int i = 0; i < n; i++) J[i * n + i] = 1; return 0; }
            if (a >= std::fabs(s)) {
                double c = 0.5 * (1 + s / a);
                for (int i = 0; i < n - 1; i++)
                    for (int j = 0; j < n - 1; j++) {
                        J[j * n + i] = -0.5 * s / (a * a * a) * x[i] * x[j];
                        if (i == j) J[j * n + i] += c;
                    }
                for (int i = 0; i < n - 1; i++) J[(n - 1) * n + i] = 0.5 * x[i] / a;
                for (int i = 0; i < n - 1; i++) J[i * n + (n - 1)] = ((-0.5 * s / (a * a)) + c / a) * x[i];
                J[(n - 1) * n + (n - 1)] = 0.5;
                return 0;
            }
            return -1;
        }
    }
    return -1;
}

// hess-projection!  src/cones.jl:198-276: Hessian of x -> Pi(x)'b.  Zero for every cone but the SOC.
inline int hess_projection(int cone, const double* x, const double* b, int p, double* hess) {
    std::fill(hess, hess + p * p, 0.0);
    if (cone != CONE_SECOND_ORDER) return 0;
    const int n = p - 1;
    double s = x[n], bs = b[n], a = 0, vbv = 0;
    for (int i = 0; i < n; i++) { a += x[i] * x[i]; vbv += x[i] * b[i]; }
    a = std::sqrt(a);
    if (a <= -s) return 0;
    if (a <= s) return 0;
    if (a > std::fabs(s)) {
        for (int i = 0; i < n; i++) {
            double hi = 0;
            for (int j = 0; j < n; j++) {
                double Hij = -x[i] * x[j] / (a * a);
                if (i == j) Hij += 1;
                hi += Hij * b[j];
            }
            hess[n * p + i] = hi / (2 * a);
            hess[i * p + n] = hi / (2 * a);
            for (int j = 0; j <= i; j++) {
                double vij = x[i] * x[j];
                double H1 = hi * x[j] * (-s / (a * a * a));
                double H2 = vij * (2 * vbv) / (a * a * a * a) - x[i] * b[j] / (a * a);
                double H3 = -vij / (a * a);
                if (i == j) { H2 -= vbv / (a * a); H3 += 1; }
                H2 *= s / a;
                H3 *= bs / a;
                hess[j * p + i] = (H1 + H2 + H3) / 2;
                hess[i * p + j] = hess[j * p + i];
            }
        }
        hess[n * p + n] = 0;
        return 0;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Constraints  (src/constraints.jl).  All are functions of one knot z = [x;u].
enum ConKind { CON_GOAL = 0, CON_BOUND = 1, CON_LINEAR = 2, CON_CIRCLE = 3, CON_SPHERE = 4, CON_NORM = 5, CON_COLLISION = 6, CON_QUATVEC = 7, CON_EXPR = 8 };

struct Constraint {
    int kind = CON_GOAL;
    int first = 1, last = 1;  // 1-based inclusive knot range (add_constraint!, src/constraint_list.jl:103-134)
    int n = 0, m = 0;
    int p = 0;                // output_dim
    int sense = CONE_ZERO;
    // GOAL: xf[p], inds[p] (0-based into x)                      src/constraints.jl:22-68
    // BOUND: zmax[n+m], zmin[n+m]; a_max/a_min = finite entries   src/constraints.jl:644-765
    // LINEAR: A[p x w] col-major, b[p], inds offset (state|control) src/constraints.jl:103-150
    // CIRCLE/SPHERE: centers + radii, xi,yi(,zi)                   src/constraints.jl:168-326
    // NORM: val, inds (into z)                                     src/constraints.jl:438-521
    std::vector<double> a, b, c3, rad;
    std::vector<int> inds;
    std::vector<int> a_max, a_min;
    int xi = 0, yi = 1, zi = 2;
    int lin_on_control = 0;
    double val = 0;
    std::vector<int> prog;          // CON_EXPR: recorded program, outputs = the last p instructions
    std::vector<double> consts;
    int width() const { return n + m; }
    int nknots() const { return last - first + 1; }
};

inline void bound_finalize(Constraint& con) {  // BoundConstraint ctor, src/constraints.jl:660-687
    con.a_max.clear(); con.a_min.clear();
    for (int i = 0; i < con.n + con.m; i++) if (std::isfinite(con.a[i])) con.a_max.push_back(i);
    for (int i = 0; i < con.n + con.m; i++) if (std::isfinite(con.b[i])) con.a_min.push_back(i);
    con.p = (int)(con.a_max.size() + con.a_min.size());
    con.sense = CONE_NEGATIVE_ORTHANT;
}

// RD.evaluate(con, z): c[p]
inline void con_evaluate(const Constraint& con, const double* x, const double* u, double* c) {
    switch (con.kind) {
        case CON_GOAL:  // :55-61
            for (int i = 0; i < con.p; i++) c[i] = x[con.inds[i]] - con.a[i];
            break;
        case CON_BOUND: {  // :738-755: upper block first (x then u), then the lower block
            int i = 0;
            for (int j : con.a_max) c[i++] = (j < con.n ? x[j] : u[j - con.n]) - con.a[j];
            for (int j : con.a_min) c[i++] = con.b[j] - (j < con.n ? x[j] : u[j - con.n]);
            break;
        }
        case CON_LINEAR: {  // :135-139: A*y - b, y = x or u
            const double* y = con.lin_on_control ? u : x;
            const int w = con.lin_on_control ? con.m : con.n;
            for (int i = 0; i < con.p; i++) {
                double s = -con.b[i];
                for (int j = 0; j < w; j++) s += con.a[j * con.p + i] * y[j];
                c[i] = s;
            }
            break;
        }
        case CON_CIRCLE:  // :190-199: -(x-xc)^2 - (y-yc)^2 + r^2
            for (int i = 0; i < con.p; i++) {
                double dx = x[con.xi] - con.a[i], dy = x[con.yi] - con.b[i];
                c[i] = -(dx * dx) - (dy * dy) + con.rad[i] * con.rad[i];
            }
            break;
        case CON_SPHERE:  // :278-290
            for (int i = 0; i < con.p; i++) {
                double dx = x[con.xi] - con.a[i], dy = x[con.yi] - con.b[i], dz = x[con.zi] - con.c3[i];
                c[i] = -(dx * dx) - (dy * dy) - (dz * dz) + con.rad[i] * con.rad[i];
            }
            break;
        case CON_NORM: {  // static evaluate :462-465 (the in-place version :470-472 is the known-buggy one)
            auto zj = [&](int j) { return j < con.n ? x[j] : u[j - con.n]; };
            if (con.sense == CONE_SECOND_ORDER) {
                for (size_t i = 0; i < con.inds.size(); i++) c[i] = zj(con.inds[i]);
                c[con.inds.size()] = con.val;
            } else {
                double s = 0;
                for (int j : con.inds) s += zj(j) * zj(j);
                c[0] = s - con.val * con.val;
            }
            break;
        }
        case CON_COLLISION: {  // src/constraints.jl:367-376 (in-place form): c = r^2, then c -= d_i^2 for i = 1..D
            const size_t D = con.inds.size() / 2;
            double s = con.val * con.val;
            for (size_t i = 0; i < D; i++) { const double d = x[con.inds[i]] - x[con.inds[D + i]]; s -= d * d; }
            c[0] = s;
            break;
        }
        case CON_EXPR: {  // user constraint (RD.@autodiff StageConstraint, docs/src/constraint_interface.md:52-72)
            Hyper reg[EXPR_MAXLEN];
            const int L = (int)con.prog.size() / 3;
            expr_run(con.prog.data(), L, con.consts.data(), con.n, x, u, true, -1, -1, reg);
            for (int i = 0; i < con.p; i++) c[i] = reg[L - con.p + i].v;
            break;
        }
        case CON_QUATVEC: {  // QuatVecEq  src/constraints.jl:947-956: q = normalize(x[qind]); qf *= -1 when qf'q < 0; -(qf[2:4] - q[2:4])
            double q[4], nrm = 0, dq = 0;
            for (int i = 0; i < 4; i++) { q[i] = x[con.inds[i]]; nrm += q[i] * q[i]; }
            nrm = std::sqrt(nrm);
            for (int i = 0; i < 4; i++) { q[i] /= nrm; dq += con.a[i] * q[i]; }
            const double sg = dq < 0 ? -1.0 : 1.0;
            for (int i = 0; i < 3; i++) c[i] = -(sg * con.a[i + 1] - q[i + 1]);
            break;
        }
    }
}

// RD.jacobian!(con, jac, c, z): jac is p x (n+m) col-major, fully written (zero-initialised here; the
// reference's Goal/Bound methods only set the +-1 entries and rely on gen_jacobian's zeros).
inline void con_jacobian(const Constraint& con, const double* x, const double* u, double* jac) {
    const int p = con.p, w = con.n + con.m;
    std::fill(jac, jac + p * w, 0.0);
    switch (con.kind) {
        case CON_GOAL: for (int i = 0; i < p; i++) jac[con.inds[i] * p + i] = 1; break;        // :62-68
        case CON_BOUND: {                                                                        // :757-765
            int i = 0;
            for (int j : con.a_max) { jac[j * p + i] = 1; i++; }
            for (int j : con.a_min) { jac[j * p + i] = -1; i++; }
            break;
        }
        case CON_LINEAR: {                                                                       // :140-144
            const int off = con.lin_on_control ? con.n : 0, wd = con.lin_on_control ? con.m : con.n;
            for (int j = 0; j < wd; j++) for (int i = 0; i < p; i++) jac[(off + j) * p + i] = con.a[j * p + i];
            break;
        }
        case CON_CIRCLE:                                                                         // :201-213
            for (int i = 0; i < p; i++) {
                jac[con.xi * p + i] = -2 * (x[con.xi] - con.a[i]);
                jac[con.yi * p + i] = -2 * (x[con.yi] - con.b[i]);
            }
            break;
        case CON_SPHERE:                                                                         // :292-306
            for (int i = 0; i < p; i++) {
                jac[con.xi * p + i] = -2 * (x[con.xi] - con.a[i]);
                jac[con.yi * p + i] = -2 * (x[con.yi] - con.b[i]);
                jac[con.zi * p + i] = -2 * (x[con.zi] - con.c3[i]);
            }
            break;
        case CON_NORM: {                                                                         // :493-517
            auto zj = [&](int j) { return j < con.n ? x[j] : u[j - con.n]; };
            if (con.sense == CONE_SECOND_ORDER) {
                for (size_t i = 0; i < con.inds.size(); i++) jac[con.inds[i] * p + i] = 1;
            } else {
                for (int j : con.inds) jac[j * p + 0] = 2 * zj(j);
            }
            break;
        }
        case CON_COLLISION: {                                                                    // :378-389
            const size_t D = con.inds.size() / 2;
            for (size_t i = 0; i < D; i++) {
                const double d = x[con.inds[i]] - x[con.inds[D + i]];
                jac[con.inds[i] * p] = -2 * d;
                jac[con.inds[D + i] * p] = 2 * d;
            }
            break;
        }
        case CON_EXPR: {  // RD.jacobian!(ForwardAD): one first-order pass per input
            Hyper reg[EXPR_MAXLEN];
            const int L = (int)con.prog.size() / 3;
            for (int j = 0; j < w; j++) {
                expr_run(con.prog.data(), L, con.consts.data(), con.n, x, u, true, j, -1, reg);
                for (int i = 0; i < p; i++) jac[j * p + i] = reg[L - p + i].d1;
            }
            break;
        }
        case CON_QUATVEC: {  // ForwardAD of the above (src/constraints.jl:938,962): d normalize(q)/dq = (I - qh qh')/|q|, rows 2:4
            double q[4], nrm = 0;
            for (int i = 0; i < 4; i++) { q[i] = x[con.inds[i]]; nrm += q[i] * q[i]; }
            nrm = std::sqrt(nrm);
            for (int i = 0; i < 4; i++) q[i] /= nrm;
            for (int j = 0; j < 4; j++)
                for (int i = 0; i < 3; i++) jac[con.inds[j] * p + i] = ((i + 1 == j ? 1.0 : 0.0) - q[i + 1] * q[j]) / nrm;
            break;
        }
    }
}

// RD.grad-jacobian!(con, H, lambda, c, z)  (reference: `∇jacobian!`, src/constraints.jl:70-73, :767-770 define it as zero for Goal / Bound; every
// other type gets RobotDynamics' ForwardDiff default): H[(n+m) x (n+m)] = d/dz (cz' lambda) = sum_i lambda_i Hess c_i(z), overwritten.
// Consumed through grad-constraint_jacobians! (src/abstract_constraint.jl:267-280) by solvers that keep the second-order constraint term.
inline void con_hess_vec(const Constraint& con, const double* x, const double* u, const double* lam, double* H) {
    const int w = con.n + con.m;
    std::fill(H, H + w * w, 0.0);
    auto add = [&](int i, int j, double v) { H[j * w + i] += v; };
    switch (con.kind) {
        case CON_GOAL: case CON_BOUND: case CON_LINEAR: break;                                   // linear in z
        case CON_CIRCLE:
            for (int i = 0; i < con.p; i++) { add(con.xi, con.xi, -2 * lam[i]); add(con.yi, con.yi, -2 * lam[i]); }
            break;
        case CON_SPHERE:
            for (int i = 0; i < con.p; i++) { add(con.xi, con.xi, -2 * lam[i]); add(con.yi, con.yi, -2 * lam[i]); add(con.zi, con.zi, -2 * lam[i]); }
            break;
        case CON_NORM:
            if (con.sense != CONE_SECOND_ORDER) for (int j : con.inds) add(j, j, 2 * lam[0]);   // c = |z_inds|^2 - val^2
            break;
        case CON_COLLISION: {                                                                    // c = r^2 - |x[x1] - x[x2]|^2
            const size_t D = con.inds.size() / 2;
            for (size_t i = 0; i < D; i++) {
                const int a = con.inds[i], b = con.inds[D + i];
                add(a, a, -2 * lam[0]); add(b, b, -2 * lam[0]); add(a, b, 2 * lam[0]); add(b, a, 2 * lam[0]);
            }
            break;
        }
        case CON_EXPR: {                                                                         // second-order forward mode: one pass per pair (j <= k)
            Hyper reg[EXPR_MAXLEN];
            const int L = (int)con.prog.size() / 3;
            for (int j = 0; j < w; j++)
                for (int k = j; k < w; k++) {
                    expr_run(con.prog.data(), L, con.consts.data(), con.n, x, u, true, j, k, reg);
                    double v = 0;
                    for (int i = 0; i < con.p; i++) v += lam[i] * reg[L - con.p + i].d12;
                    H[k * w + j] = v; H[j * w + k] = v;
                }
            break;
        }
        case CON_QUATVEC: {   // c_i = q_{i+1} / |q| - const:  d2 (q_a/|q|) / dq_j dq_k = -(d_aj q_k + d_ak q_j + d_jk q_a)/|q|^3 + 3 q_a q_j q_k / |q|^5
            double q[4], n2 = 0;
            for (int i = 0; i < 4; i++) { q[i] = x[con.inds[i]]; n2 += q[i] * q[i]; }
            const double n1 = std::sqrt(n2), i3 = 1.0 / (n2 * n1), i5 = i3 / n2;
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 4; k++) {
                    double v = 0;
                    for (int i = 0; i < 3; i++) {
                        const int a = i + 1;
                        v += lam[i] * (-((a == j ? q[k] : 0.0) + (a == k ? q[j] : 0.0) + (j == k ? q[a] : 0.0)) * i3 + 3 * q[a] * q[j] * q[k] * i5);
                    }
                    add(con.inds[j], con.inds[k], v);
                }
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Solver options -- Altro.jl `SolverOptions` defaults, restated (Altro is not under /root/reference).
struct Options {
    double bp_reg_increase_factor = 1.6;
    double bp_reg_max = 1e8;
    double bp_reg_min = 1e-8;
    double bp_reg_initial = 0.0;
    double bp_reg_fp = 10.0;
    double line_search_lower_bound = 1e-8;
    double line_search_upper_bound = 10.0;
    int iterations_linesearch = 10;   // alphas tried: 1, 1/2, ..., 2^-10  (11 trials)
    double max_state_value = 1e8;
    double max_control_value = 1e8;
    double penalty_initial = 1.0;
    double penalty_scaling = 10.0;
    double penalty_max = 1e8;
    double dual_max = 1e8;
    // 0: Cholesky solve + the long form of the cost-to-go update (below).  1 (m = 4 only): the ALGEBRA of the register-resident GPU
    // kernel (csrc/riccati_frag.cu) in plain loops -- (Quu + rho I)^-1 by 2 x 2 block elimination, K = -Minv Qux, W = (I + rho Minv) Qux,
    // S <- sym(Qxx + W'K), s <- Qx + K'w_d, dV2 = -(dV1 + rho d'd)/2.  Mathematically identical; the two differ by rounding only,
    // which the Riccati recursion of an ill-conditioned trajectory amplifies (tests/test_oracle_variants.py measures by how much).
    int backward_variant = 0;
    // test instrument: after a successful backward pass multiply every gain K, d by (1 + gain_noise * u), u in (-1, 1) deterministic in
    // (instance, entry, pass).  A problem run with gain_noise = the kernel tolerance is the yardstick for what a backward pass that is accurate
    // to that tolerance may do to the iterates downstream (tests/parity_util.py): the closed loop amplifies it by orders of magnitude.
    double gain_noise = 0.0;
    unsigned noise_epoch = 0;
};

struct Problem {
    ModelParams model;
    std::vector<DynProg> dyn;      // MODEL_EXPR: the models of a hybrid problem and which one steps knot k -> k+1
    std::vector<int> dyn_index;    // N-1
    ModelParams model_at(int k) const { ModelParams mp = model; if (model.id == MODEL_EXPR) mp.prog = &dyn[dyn_index[k]]; return mp; }
    int n = 0, m = 0, N = 0, B = 0;
    // Lie-group error state (to_spec.error_state): the Riccati recursion, the gains and the feedback act on ne = n - 1 dimensions,
    // the quaternion x[qs..qs+3] contributing its 3-dimensional differential (Altro.jl + RobotDynamics LieState, restated below)
    bool lie = false;
    int ne = 0, qs = 3;
    std::vector<double> ABe;       // B*(N-1)*ne*(ne+m): [A_e B_e] = G_{k+1}' [A G_k | B]
    std::vector<double> dt;        // N-1
    double t0 = 0;
    std::vector<Cost> costs;
    std::vector<int> cost_index;   // N
    std::vector<Constraint> cons;
    std::vector<int> con_offset;   // offset of each constraint's multipliers inside one instance's Lambda
    int lambda_len = 0;            // sum_c nknots_c * p_c
    Options opts;
    std::vector<double> mu;        // penalty per constraint
    // per-instance data
    std::vector<double> x0;        // B*n
    std::vector<double> X, U;      // B*N*n, B*(N-1)*m
    std::vector<double> Xc, Uc;    // line-search candidates
    std::vector<double> AB;        // B*(N-1)*n*(n+m)
    std::vector<double> K, d;      // B*(N-1)*m*n, B*(N-1)*m
    std::vector<double> lambda;    // B*lambda_len
    std::vector<double> rho, drho; // B
    std::vector<double> dV;        // B*2
    std::vector<double> J;         // B  (AL merit of the current trajectory)
    std::vector<double> alpha;     // B  (accepted step, 0 = line search failed)
    std::vector<int> bp_status;    // B  (0 ok, >0 = number of regularisation restarts, -1 = failed)
    std::vector<int> ls_iters;     // B  trials used by the last forward pass
    bool J_valid = false;

    void finalize() {
        n = model.n; m = model.m;
        ne = lie ? n - 1 : n;
        con_offset.clear(); lambda_len = 0;
        for (auto& c : cons) { con_offset.push_back(lambda_len); lambda_len += c.nknots() * c.p; }
        mu.assign(cons.size(), opts.penalty_initial);
        x0.assign((size_t)B * n, 0.0);
        X.assign((size_t)B * N * n, std::numeric_limits<double>::quiet_NaN());   // X0 = NaN, src/problem.jl:83
        U.assign((size_t)B * (N - 1) * m, 0.0);                                  // U0 = 0,   src/problem.jl:84
        Xc = X; Uc = U;
        AB.assign((size_t)B * (N - 1) * n * (n + m), 0.0);
        ABe.assign(lie ? (size_t)B * (N - 1) * ne * (ne + m) : 0, 0.0);
        K.assign((size_t)B * (N - 1) * m * ne, 0.0);
        d.assign((size_t)B * (N - 1) * m, 0.0);
        lambda.assign((size_t)B * lambda_len, 0.0);
        rho.assign(B, opts.bp_reg_initial); drho.assign(B, 0.0);
        dV.assign((size_t)B * 2, 0.0);
        J.assign(B, 0.0); alpha.assign(B, 0.0); bp_status.assign(B, 0); ls_iters.assign(B, 0);
        J_valid = false;
    }
    double* Xb(int b) { return &X[(size_t)b * N * n]; }
    double* Ub(int b) { return &U[(size_t)b * (N - 1) * m]; }
    double* ABb(int b) { return &AB[(size_t)b * (N - 1) * n * (n + m)]; }
    double* ABeb(int b) { return &ABe[(size_t)b * (N - 1) * ne * (ne + m)]; }
    double* Kb(int b) { return &K[(size_t)b * (N - 1) * m * ne]; }
    double* db(int b) { return &d[(size_t)b * (N - 1) * m]; }
    double* lamb(int b) { return &lambda[(size_t)b * lambda_len]; }
};

// ------------------------------------------------------------------------------------------------
// Lie-group error state.  None of this arithmetic is under /root/reference (RobotDynamics.jl LieState / Rotations.jl / Altro.jl);
// restated from their published formulas -- PARITY UNPINNED, checked in tests/ against finite differences of the group operation.
//   grad-differential(q) = L(q) H  (4 x 3, Rotations.jl): d/dphi of q (x) cayley(phi) at phi = 0; columns (-x,w,z,-y), (-y,-z,w,x), (-z,y,-x,w)
//   grad^2-differential(q, b) = -(q'b) I3
//   state_diff(xbar, x): vector parts xbar - x; rotation part = inverse Cayley map of q^-1 (x) qbar = vec / scalar
//   (the reference's own hook: error_expansion! multiplies constraint Jacobians by G, src/abstract_constraint.jl:282-303)
inline void quat_G(const double* q, double* G /*4x3 col-major*/) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    G[0] = -x; G[1] = w;  G[2] = z;  G[3] = -y;
    G[4] = -y; G[5] = -z; G[6] = w;  G[7] = x;
    G[8] = -z; G[9] = y;  G[10] = -x; G[11] = w;
}
// dx[ne] = state_diff(xbar, x)
inline void state_diff(const Problem& P, const double* xbar, const double* x, double* dx) {
    if (!P.lie) { for (int i = 0; i < P.n; i++) dx[i] = xbar[i] - x[i]; return; }
    const int qs = P.qs;
    for (int i = 0; i < qs; i++) dx[i] = xbar[i] - x[i];
    const double* q = x + qs; const double* p = xbar + qs;       // dq = conj(q) (x) p
    const double dw = q[0] * p[0] + q[1] * p[1] + q[2] * p[2] + q[3] * p[3];
    const double d1 = q[0] * p[1] - p[0] * q[1] - (q[2] * p[3] - q[3] * p[2]);
    const double d2 = q[0] * p[2] - p[0] * q[2] - (q[3] * p[1] - q[1] * p[3]);
    const double d3 = q[0] * p[3] - p[0] * q[3] - (q[1] * p[2] - q[2] * p[1]);
    dx[qs] = d1 / dw; dx[qs + 1] = d2 / dw; dx[qs + 2] = d3 / dw;
    for (int i = qs + 4; i < P.n; i++) dx[i - 1] = xbar[i] - x[i];
}
// row index map error state -> full state outside the quaternion
inline int lie_full_index(const Problem& P, int e) { return e < P.qs ? e : e + 1; }

// M (rows x n, col-major, leading dim ld) times E(x) = blkdiag(I, G(q), I): out (rows x ne, leading dim ldo)
inline void times_E(const Problem& P, const double* x, const double* M, int rows, int ld, double* out, int ldo) {
    double G[12]; quat_G(x + P.qs, G);
    for (int e = 0; e < P.ne; e++) {
        if (e >= P.qs && e < P.qs + 3) {
            const double* g = &G[(e - P.qs) * 4];
            for (int i = 0; i < rows; i++) {
                double t = 0;
                for (int r = 0; r < 4; r++) t += M[(P.qs + r) * ld + i] * g[r];
                out[e * ldo + i] = t;
            }
        } else {
            const int j = lie_full_index(P, e);
            for (int i = 0; i < rows; i++) out[e * ldo + i] = M[j * ld + i];
        }
    }
}
// E(x)' times M (n x cols, col-major, leading dim ld): out (ne x cols, leading dim ldo)
inline void Et_times(const Problem& P, const double* x, const double* M, int cols, int ld, double* out, int ldo) {
    double G[12]; quat_G(x + P.qs, G);
    for (int j = 0; j < cols; j++)
        for (int e = 0; e < P.ne; e++) {
            if (e >= P.qs && e < P.qs + 3) {
                const double* g = &G[(e - P.qs) * 4];
                double t = 0;
                for (int r = 0; r < 4; r++) t += g[r] * M[j * ld + P.qs + r];
                out[j * ldo + e] = t;
            } else out[j * ldo + e] = M[j * ld + lie_full_index(P, e)];
        }
}
// Altro error_expansion!(D, model, G): A_e = G_{k+1}' A G_k, B_e = G_{k+1}' B.  AB: n x (n+m), ABe: ne x (ne+m), both col-major
inline void error_dynamics(const Problem& P, const double* xk, const double* xk1, const double* AB, double* ABe) {
    const int n = P.n, m = P.m, ne = P.ne;
    double tmp[MAXN * (MAXN + MAXM)];                      // [A G_k | B] : n x (ne+m)
    times_E(P, xk, AB, n, n, tmp, n);
    for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) tmp[(ne + j) * n + i] = AB[(n + j) * n + i];
    Et_times(P, xk1, tmp, ne + m, n, ABe, ne);
}
// Altro error_expansion!(E, Q, model, Z, G) on the full-state expansion (grad[n+m], hess (n+m)^2) of knot x:
//   E.x = G'q ; E.u = r ; E.xx = G'Q G + grad^2-differential(x, q) ; E.ux = H G ; E.uu = R.   Output (ne+m) sized, col-major symmetric.
inline void error_expansion(const Problem& P, const double* x, const double* grad, const double* hess, double* ge, double* He) {
    const int n = P.n, m = P.m, nm = n + m, ne = P.ne, nme = ne + m;
    double T1[(MAXN + MAXM) * (MAXN + MAXM)], T2[(MAXN + MAXM) * (MAXN + MAXM)];
    // columns: hess (nm x nm) * blkdiag(E, I_m) -> T1 (nm x nme)
    times_E(P, x, hess, nm, nm, T1, nm);
    for (int j = 0; j < m; j++) for (int i = 0; i < nm; i++) T1[(ne + j) * nm + i] = hess[(n + j) * nm + i];
    // rows: blkdiag(E, I_m)' * T1 -> He (nme x nme); the u rows are copied
    Et_times(P, x, T1, nme, nm, T2, nme);
    for (int j = 0; j < nme; j++) {
        for (int e = 0; e < ne; e++) He[j * nme + e] = T2[j * nme + e];
        for (int a = 0; a < m; a++) He[j * nme + ne + a] = T1[j * nm + n + a];
    }
    Et_times(P, x, grad, 1, nm, ge, nme);
    for (int a = 0; a < m; a++) ge[ne + a] = grad[n + a];
    double qb = 0;
    for (int r = 0; r < 4; r++) qb += x[P.qs + r] * grad[P.qs + r];
    for (int i = 0; i < 3; i++) He[(P.qs + i) * nme + P.qs + i] -= qb;
}

// rollout!  src/problem.jl:334-340
inline void rollout(Problem& P, int b) {
    const int n = P.n, m = P.m, N = P.N;
    double* X = P.Xb(b); const double* U = P.Ub(b);
    for (int i = 0; i < n; i++) X[i] = P.x0[(size_t)b * n + i];
    for (int k = 1; k < N; k++) rk4_step<double>(P.model_at(k - 1), &X[(k - 1) * n], &U[(k - 1) * m], P.dt[k - 1], &X[k * n]);
}

// cost! / get_J  src/objective.jl:104-110 : per-knot J_k
inline void cost_knots(const Problem& P, const double* X, const double* U, double* Jk) {
    for (int k = 0; k < P.N; k++) {
        const bool last = (k == P.N - 1);
        Jk[k] = cost_value(P.costs[P.cost_index[k]], &X[k * P.n], last ? nullptr : &U[k * P.m], !last);
    }
}
// cost  src/objective.jl:89-93
inline double cost_total(const Problem& P, const double* X, const double* U) {
    double s = 0;   // same left-to-right summation as sum(get_J(obj))
    for (int k = 0; k < P.N; k++) {
        const bool last = (k == P.N - 1);
        s += cost_value(P.costs[P.cost_index[k]], &X[k * P.n], last ? nullptr : &U[k * P.m], !last);
    }
    return s;
}


// evaluate_constraints!  src/abstract_constraint.jl:200-225 : vals[nknots][p] for one constraint
inline void evaluate_constraints(const Problem& P, int ci, const double* X, const double* U, double* vals) {
    const Constraint& con = P.cons[ci];
    for (int k = con.first; k <= con.last; k++) {
        const double* u = (k == P.N) ? ZERO_U : &U[(k - 1) * P.m];
        con_evaluate(con, &X[(k - 1) * P.n], u, &vals[(k - con.first) * con.p]);
    }
}
// constraint_jacobians!  src/abstract_constraint.jl:236-248 : jacs[nknots][p x (n+m)]
inline void constraint_jacobians(const Problem& P, int ci, const double* X, const double* U, double* jacs) {
    const Constraint& con = P.cons[ci];
    const int sz = con.p * (P.n + P.m);
    for (int k = con.first; k <= con.last; k++) {
        const double* u = (k == P.N) ? ZERO_U : &U[(k - 1) * P.m];
        con_jacobian(con, &X[(k - 1) * P.n], u, &jacs[(k - con.first) * sz]);
    }
}

// max violation: equality |c|, inequality max(0,c), SOC distance-to-cone ||c - Pi(c)||_inf
inline double max_violation(const Problem& P, const double* X, const double* U) {
    double v = 0;
    std::vector<double> c, pc;
    for (size_t ci = 0; ci < P.cons.size(); ci++) {
        const Constraint& con = P.cons[ci];
        c.assign((size_t)con.nknots() * con.p, 0.0); pc.resize(con.p);
        evaluate_constraints(P, (int)ci, X, U, c.data());
        for (int k = 0; k < con.nknots(); k++) {
            projection(con.sense, &c[k * con.p], con.p, pc.data());
            for (int i = 0; i < con.p; i++) v = std::max(v, std::fabs(c[k * con.p + i] - pc[i]));
        }
    }
    return v;
}

// Augmented-Lagrangian merit in conic form (Altro.jl ALConstraint; building blocks src/cones.jl, write-up test/socp.jl:52-82):
//   lbar = lambda - mu c ; lp = Pi_{K*}(lbar) ; J_AL = (|lp|^2 - |lambda|^2) / (2 mu)
inline double al_penalty(const Problem& P, const double* X, const double* U, const double* lam) {
    double J = 0;
    double c[MAXP], lbar[MAXP], lp[MAXP];
    for (size_t ci = 0; ci < P.cons.size(); ci++) {
        const Constraint& con = P.cons[ci];
        const int p = con.p; const double mu = P.mu[ci];
        const double* l = lam + P.con_offset[ci];
        for (int k = 0; k < con.nknots(); k++) {
            const int k1 = con.first + k;
            con_evaluate(con, &X[(k1 - 1) * P.n], (k1 == P.N) ? ZERO_U : &U[(k1 - 1) * P.m], c);
            double a = 0, bsum = 0;
            for (int i = 0; i < p; i++) lbar[i] = l[k * p + i] - mu * c[i];
            projection(dualcone(con.sense), lbar, p, lp);
            for (int i = 0; i < p; i++) { a += lp[i] * lp[i]; bsum += l[k * p + i] * l[k * p + i]; }
            J += (a - bsum) / (2 * mu);
        }
    }
    return J;
}

inline double merit(const Problem& P, const double* X, const double* U, const double* lam) {
    return cost_total(P, X, U) + al_penalty(P, X, U, lam);
}

// Dynamics expansion: [A B] = d x_{k+1} / d [x_k; u_k] by forward-mode duals through the RK4 step
// (RobotDynamics jacobian!(ForwardAD); shape pinned by test/dynamics_constraints.jl:35,57-62).
template <int NM>
inline void dynamics_jacobian_t(const ModelParams& mp, const double* x, const double* u, double h, double* AB) {
    using D = Dual<NM>;
    const int n = mp.n, m = mp.m;
    D xd[MAXN], ud[MAXM], xn[MAXN];
    for (int i = 0; i < n; i++) { xd[i] = D(x[i]); xd[i].d[i] = 1; }
    for (int i = 0; i < m; i++) { ud[i] = D(u[i]); ud[i].d[n + i] = 1; }
    rk4_step<D>(mp, xd, ud, h, xn);
    for (int j = 0; j < n + m; j++)
        for (int i = 0; i < n; i++) AB[j * n + i] = xn[i].d[j];
}
inline void dynamics_jacobian(const ModelParams& mp, const double* x, const double* u, double h, double* AB) {
    switch (mp.n + mp.m) {
        case 3: dynamics_jacobian_t<3>(mp, x, u, h, AB); break;
        case 5: dynamics_jacobian_t<5>(mp, x, u, h, AB); break;
        case 6: dynamics_jacobian_t<6>(mp, x, u, h, AB); break;
        case 17: dynamics_jacobian_t<17>(mp, x, u, h, AB); break;
        default: dynamics_jacobian_t<MAXN + MAXM>(mp, x, u, h, AB); break;
    }
}
inline void expand_dynamics(Problem& P, int b) {
    const int n = P.n, m = P.m;
    for (int k = 0; k < P.N - 1; k++) {
        dynamics_jacobian(P.model_at(k), &P.Xb(b)[k * n], &P.Ub(b)[k * m], P.dt[k], &P.ABb(b)[(size_t)k * n * (n + m)]);
        if (P.lie) error_dynamics(P, &P.Xb(b)[k * n], &P.Xb(b)[(k + 1) * n], &P.ABb(b)[(size_t)k * n * (n + m)], &P.ABeb(b)[(size_t)k * P.ne * (P.ne + m)]);
    }
}

// Cost expansion of knot k (0-based) including the AL terms:
//   grad += -cz' D' lp ; hess += mu cz' D' D cz   with D = grad Pi_{K*}(lbar)   (Gauss-Newton; the
//   second-order projection term hess-projection! is zero for orthant / zero cones, src/cones.jl:201-206)
inline void cost_expansion(const Problem& P, const double* X, const double* U, const double* lam, int k,
                           double* grad /*n+m*/, double* hess /*(n+m)^2 col-major, symmetric*/) {
    const int n = P.n, m = P.m, nm = n + m;
    const bool last = (k == P.N - 1);
    const double* x = &X[k * n];
    const double* u = last ? ZERO_U : &U[k * m];
    std::fill(grad, grad + nm, 0.0);
    std::fill(hess, hess + nm * nm, 0.0);
    const Cost& c = P.costs[P.cost_index[k]];
    cost_gradient(c, x, u, last, grad);
    cost_hessian(c, x, u, last, hess, true);
    double cv[MAXP], jac[MAXP * (MAXN + MAXM)], lbar[MAXP], lp[MAXP], Dm[MAXP * MAXP], tmp[MAXP * (MAXN + MAXM)];   // no heap traffic in the hot loop
    for (size_t ci = 0; ci < P.cons.size(); ci++) {
        const Constraint& con = P.cons[ci];
        if (k + 1 < con.first || k + 1 > con.last) continue;
        const int p = con.p; const double mu = P.mu[ci];
        con_evaluate(con, x, u, cv);
        if (con.kind == CON_GOAL || con.kind == CON_BOUND) {
            // selector Jacobians (+-1 entries, src/constraints.jl:62-68, :757-765): same arithmetic as the dense path
            // below, written row by row so the CPU baseline is not handicapped by multiplying structural zeros
            const double* l0 = lam + P.con_offset[ci] + (size_t)(k + 1 - con.first) * p;
            const int dc0 = dualcone(con.sense);
            int row = 0;
            auto add_row = [&](int j, double sgn) {
                const double lb = l0[row] - mu * cv[row];
                const bool active = (dc0 == CONE_IDENTITY) || (lb <= 0);
                if (active && (j < n || !last)) { grad[j] -= sgn * lb; hess[j * nm + j] += mu; }
                row++;
            };
            if (con.kind == CON_GOAL) for (int i = 0; i < p; i++) add_row(con.inds[i], 1.0);
            else { for (int j : con.a_max) add_row(j, 1.0); for (int j : con.a_min) add_row(j, -1.0); }
            continue;
        }
        con_jacobian(con, x, u, jac);
        const double* l = lam + P.con_offset[ci] + (size_t)(k + 1 - con.first) * p;
        for (int i = 0; i < p; i++) lbar[i] = l[i] - mu * cv[i];
        const int dc = dualcone(con.sense);
        projection(dc, lbar, p, lp);
        grad_projection(dc, lbar, p, Dm);
        // tmp = D * cz  (p x nm)
        for (int j = 0; j < nm; j++)
            for (int i = 0; i < p; i++) {
                double s = 0;
                for (int r = 0; r < p; r++) s += Dm[r * p + i] * jac[j * p + r];
                tmp[j * p + i] = s;
            }
        const int lim = last ? n : nm;   // terminal knot: state part only
        for (int j = 0; j < lim; j++) {
            double g = 0;
            for (int i = 0; i < p; i++) g += tmp[j * p + i] * lp[i];
            grad[j] -= g;
            for (int j2 = 0; j2 < lim; j2++) {
                double hsum = 0;
                for (int i = 0; i < p; i++) hsum += tmp[j * p + i] * tmp[j2 * p + i];
                hess[j2 * nm + j] += mu * hsum;
            }
        }
    }
}

// Altro.jl regularization_update!, restated.
inline void reg_increase(const Options& o, double& rho, double& drho) {
    drho = std::max(drho * o.bp_reg_increase_factor, o.bp_reg_increase_factor);
    rho = std::max(rho * drho, o.bp_reg_min);
}
inline void reg_decrease(const Options& o, double& rho, double& drho) {
    drho = std::min(drho / o.bp_reg_increase_factor, 1.0 / o.bp_reg_increase_factor);
    rho = rho * drho * ((rho * drho > o.bp_reg_min) ? 1.0 : 0.0);
}

// Riccati backward pass (Altro.jl backwardpass!, restated).  Regularisation on Quu (bp_reg_type = :control).
//   Qx = lx + A's ; Qu = lu + B's ; Qxx = lxx + A'SA ; Quu = luu + B'SB ; Qux = lux + B'SA
//   K = -(Quu + rho I)^-1 Qux ; d = -(Quu + rho I)^-1 Qu
//   S <- Qxx + K'Quu K + K'Qux + Qux'K (symmetrised) ; s <- Qx + K'Quu d + K'Qu + Qux'd
//   dV += (d'Qu, 1/2 d'Quu d)
// A non-positive Cholesky pivot of Quu + rho I increases rho and restarts from the terminal knot.
// With the Lie-group error state (P.lie) the same recursion runs on n = ne dimensions: [A B] -> [A_e B_e] (error_dynamics),
// the expansion of every knot -> error_expansion of the full-state one.
inline int backward_pass(Problem& P, int b) {
    const int n = P.ne, m = P.m, nm = n + m, N = P.N, nf = P.n, nmf = nf + m;
    const double* X = P.Xb(b); const double* U = P.Ub(b); const double* lam = P.lamb(b);
    double* Kall = P.Kb(b); double* dall = P.db(b);
    std::vector<double> gradF(nmf), hessF((size_t)nmf * nmf);
    auto knot_expansion = [&](int k, double* g, double* H) {
        if (!P.lie) { cost_expansion(P, X, U, lam, k, g, H); return; }
        cost_expansion(P, X, U, lam, k, gradF.data(), hessF.data());
        error_expansion(P, &X[k * nf], gradF.data(), hessF.data(), g, H);
    };
    std::vector<double> grad(nm), hess((size_t)nm * nm), S((size_t)n * n), s(n), Sn((size_t)n * n), sn(n);
    std::vector<double> SAB((size_t)n * nm), Qzz((size_t)nm * nm), Qz(nm), L((size_t)m * m), Kd((size_t)m * (n + 1));
    std::vector<double> QuuK((size_t)m * n), Quud(m);
    int restarts = 0;
    for (;;) {
        knot_expansion(N - 1, grad.data(), hess.data());
        for (int j = 0; j < n; j++) { s[j] = grad[j]; for (int i = 0; i < n; i++) S[j * n + i] = hess[j * nm + i]; }
        double dV1 = 0, dV2 = 0;
        bool ok = true;
        const double rho = P.rho[b];
        for (int k = N - 2; k >= 0; k--) {
            const double* AB = P.lie ? &P.ABeb(b)[(size_t)k * n * nm] : &P.ABb(b)[(size_t)k * n * nm];
            knot_expansion(k, grad.data(), hess.data());
            // SAB = S * [A B]
            for (int j = 0; j < nm; j++)
                for (int i = 0; i < n; i++) {
                    double t = 0;
                    for (int r = 0; r < n; r++) t += S[r * n + i] * AB[j * n + r];
                    SAB[j * n + i] = t;
                }
            // Qzz = lzz + [A B]' S [A B] ; Qz = lz + [A B]' s
            for (int j = 0; j < nm; j++) {
                for (int i = 0; i < nm; i++) {
                    double t = 0;
                    for (int r = 0; r < n; r++) t += AB[i * n + r] * SAB[j * n + r];
                    Qzz[j * nm + i] = hess[j * nm + i] + t;
                }
                double t = 0;
                for (int r = 0; r < n; r++) t += AB[j * n + r] * s[r];
                Qz[j] = grad[j] + t;
            }
            if (P.opts.backward_variant == 1 && m == 4) {
                auto Mq = [&](int a, int c) { return Qzz[(n + (a > c ? a : c)) * nm + n + (a > c ? c : a)]; };   // upper-triangle entry of Quu
                const double a = Mq(0, 0) + rho, bq = Mq(0, 1), c = Mq(1, 1) + rho;
                const double M20 = Mq(0, 2), M30 = Mq(0, 3), M21 = Mq(1, 2), M31 = Mq(1, 3);
                const double R00 = Mq(2, 2) + rho, R01 = Mq(2, 3), R11 = Mq(3, 3) + rho;
                const double detP = std::fma(a, c, -bq * bq);
                const double iP = 1.0 / detP;
                const double yt00 = std::fma(M20, c, -M21 * bq), yt01 = std::fma(M21, a, -M20 * bq);
                const double yt10 = std::fma(M30, c, -M31 * bq), yt11 = std::fma(M31, a, -M30 * bq);
                const double z00 = std::fma(yt00, M20, yt01 * M21), z01 = std::fma(yt00, M30, yt01 * M31), z11 = std::fma(yt10, M30, yt11 * M31);
                const double s00 = std::fma(-iP, z00, R00), s01 = std::fma(-iP, z01, R01), s11 = std::fma(-iP, z11, R11);
                const double detS = std::fma(s00, s11, -s01 * s01);
                const double iS = 1.0 / detS;
                if (!((a > 0) && (detP > 0) && (s00 > 0) && (detS > 0) && (detP < 1e300) && (detS < 1e300))) { ok = false; break; }
                const double v00 = s11 * iS, v01 = -s01 * iS, v11 = s00 * iS;
                const double y00 = yt00 * iP, y01 = yt01 * iP, y10 = yt10 * iP, y11 = yt11 * iP;
                const double n00 = -std::fma(v00, y00, v01 * y10), n01 = -std::fma(v00, y01, v01 * y11);
                const double n10 = -std::fma(v01, y00, v11 * y10), n11 = -std::fma(v01, y01, v11 * y11);
                const double p00 = std::fma(c, iP, -std::fma(y00, n00, y10 * n10));
                const double p01 = std::fma(-bq, iP, -std::fma(y00, n01, y10 * n11));
                const double p11 = std::fma(a, iP, -std::fma(y01, n01, y11 * n11));
                const double Mi[4][4] = {{p00, p01, n00, n10}, {p01, p11, n01, n11}, {n00, n01, v00, v01}, {n10, n11, v01, v11}};
                double* Kk = &Kall[(size_t)k * m * n]; double* dk = &dall[(size_t)k * m];
                std::vector<double> Wk((size_t)m * n);
                double wd[4];
                for (int cc = 0; cc <= n; cc++)
                    for (int i = 0; i < 4; i++) {
                        double kk = 0, ww = 0;
                        for (int r = 0; r < 4; r++) {
                            const double q = cc < n ? Qzz[(n + r) * nm + cc] : Qz[n + r];      // Qxu[cc][r] (the kernel reads the x rows of the u columns) | Qu[r]
                            kk = std::fma(q, -Mi[r][i], kk);
                            ww = std::fma(q, (r == i ? 1.0 : 0.0) + rho * Mi[r][i], ww);
                        }
                        if (cc < n) { Kk[cc * m + i] = kk; Wk[cc * m + i] = ww; } else { dk[i] = kk; wd[i] = ww; }
                    }
                for (int j = 0; j < n; j++) {
                    for (int i = 0; i < n; i++) {
                        double t = Qzz[j * nm + i];
                        for (int r = 0; r < m; r++) t = std::fma(Wk[i * m + r], Kk[j * m + r], t);
                        Sn[j * n + i] = t;
                    }
                    double t = Qz[j];
                    for (int r = 0; r < m; r++) t = std::fma(wd[r], Kk[j * m + r], t);
                    sn[j] = t;
                }
                // (the antisymmetric part of S is an unstable mode of the recursion: every implementation has to remove it)
                for (int j = 0; j < n; j++) { s[j] = sn[j]; for (int i = 0; i < n; i++) S[j * n + i] = 0.5 * (Sn[j * n + i] + Sn[i * n + j]); }
                for (int i = 0; i < m; i++) { dV1 = std::fma(dk[i], Qz[n + i], dV1); dV2 = std::fma(dk[i], dk[i], dV2); }   // dV2 holds sum d'd until the end
                continue;
            }
            // Cholesky of Quu + rho I (lower L, col-major m x m)
            for (int j = 0; j < m; j++) {
                for (int i = j; i < m; i++) {
                    double t = Qzz[(n + j) * nm + (n + i)] + (i == j ? rho : 0.0);
                    for (int r = 0; r < j; r++) t -= L[r * m + i] * L[r * m + j];
                    if (i == j) {
                        if (!(t > 0) || !std::isfinite(t)) { ok = false; break; }
                        L[j * m + j] = std::sqrt(t);
                    } else L[j * m + i] = t / L[j * m + j];
                }
                if (!ok) break;
            }
            if (!ok) break;
            // solve (Quu+rho I) [K d] = -[Qux Qu]   (Kd is m x (n+1) col-major)
            for (int c = 0; c <= n; c++) {
                double y[MAXM];
                for (int i = 0; i < m; i++) {
                    double t = -(c < n ? Qzz[c * nm + (n + i)] : Qz[n + i]);
                    for (int r = 0; r < i; r++) t -= L[r * m + i] * y[r];
                    y[i] = t / L[i * m + i];
                }
                for (int i = m - 1; i >= 0; i--) {
                    double t = y[i];
                    for (int r = i + 1; r < m; r++) t -= L[i * m + r] * Kd[c * m + r];
                    Kd[c * m + i] = t / L[i * m + i];
                }
            }
            double* Kk = &Kall[(size_t)k * m * n]; double* dk = &dall[(size_t)k * m];
            for (int i = 0; i < m * n; i++) Kk[i] = Kd[i];
            for (int i = 0; i < m; i++) dk[i] = Kd[n * m + i];
            // QuuK = Quu K, Quud = Quu d (unregularised Quu)
            for (int c = 0; c < n; c++)
                for (int i = 0; i < m; i++) {
                    double t = 0;
                    for (int r = 0; r < m; r++) t += Qzz[(n + r) * nm + (n + i)] * Kk[c * m + r];
                    QuuK[c * m + i] = t;
                }
            for (int i = 0; i < m; i++) {
                double t = 0;
                for (int r = 0; r < m; r++) t += Qzz[(n + r) * nm + (n + i)] * dk[r];
                Quud[i] = t;
            }
            // S <- Qxx + K'QuuK + K'Qux + Qux'K ; s <- Qx + K'Quud + K'Qu + Qux'd
            for (int j = 0; j < n; j++) {
                for (int i = 0; i < n; i++) {
                    double t = Qzz[j * nm + i];
                    for (int r = 0; r < m; r++)
                        t += Kk[i * m + r] * QuuK[j * m + r] + Kk[i * m + r] * Qzz[j * nm + (n + r)] + Qzz[i * nm + (n + r)] * Kk[j * m + r];
                    Sn[j * n + i] = t;
                }
                double t = Qz[j];
                for (int r = 0; r < m; r++) t += Kk[j * m + r] * Quud[r] + Kk[j * m + r] * Qz[n + r] + Qzz[j * nm + (n + r)] * dk[r];
                sn[j] = t;
            }
            for (int j = 0; j < n; j++) { s[j] = sn[j]; for (int i = 0; i < n; i++) S[j * n + i] = 0.5 * (Sn[j * n + i] + Sn[i * n + j]); }
            for (int i = 0; i < m; i++) { dV1 += dk[i] * Qz[n + i]; dV2 += 0.5 * dk[i] * Quud[i]; }
        }
        if (ok && P.opts.backward_variant == 1 && m == 4) dV2 = -0.5 * std::fma(rho, dV2, dV1);   // 1/2 d'Quu d = -(d'Qu + rho d'd)/2
        if (ok) { P.dV[2 * b] = dV1; P.dV[2 * b + 1] = dV2; break; }
        reg_increase(P.opts, P.rho[b], P.drho[b]);
        restarts++;
        if (P.rho[b] > P.opts.bp_reg_max) { P.bp_status[b] = -1; return -1; }
    }
    if (P.opts.gain_noise > 0.0) {
        auto unit = [&](unsigned long long i) {     // splitmix64 -> (-1, 1)
            unsigned long long z = i + 0x9E3779B97F4A7C15ULL * (1ULL + P.opts.noise_epoch) + ((unsigned long long)b << 32);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
            return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        };
        const size_t nk = (size_t)(N - 1) * m * n, nd = (size_t)(N - 1) * m;
        for (size_t i = 0; i < nk; i++) Kall[i] *= 1.0 + P.opts.gain_noise * unit(i);
        for (size_t i = 0; i < nd; i++) dall[i] *= 1.0 + P.opts.gain_noise * unit(nk + i);
    }
    reg_decrease(P.opts, P.rho[b], P.drho[b]);
    P.bp_status[b] = restarts;
    return restarts;
}

// Closed-loop rollout for step size alpha (Altro.jl rollout!(solver, alpha), restated):
//   dx = xbar_k - x_k ; ubar_k = u_k + K_k dx + alpha d_k ; xbar_{k+1} = f(xbar_k, ubar_k)
// returns false when a state/control exceeds max_state_value / max_control_value (or is NaN).
inline bool forward_rollout(const Problem& P, int b, double alpha, double* Xc, double* Uc) {
    const int n = P.n, m = P.m, N = P.N, ne = P.ne;
    const double* X = &P.X[(size_t)b * N * n]; const double* U = &P.U[(size_t)b * (N - 1) * m];
    const double* K = &P.K[(size_t)b * (N - 1) * m * ne]; const double* d = &P.d[(size_t)b * (N - 1) * m];
    for (int i = 0; i < n; i++) Xc[i] = P.x0[(size_t)b * n + i];
    for (int k = 0; k < N - 1; k++) {
        double dx[MAXN];
        state_diff(P, &Xc[k * n], &X[k * n], dx);     // RD.state_diff(model, xbar, x): plain difference without a Lie group
        for (int a = 0; a < m; a++) {
            double t = U[k * m + a] + alpha * d[k * m + a];
            for (int i = 0; i < ne; i++) t += K[(size_t)k * m * ne + i * m + a] * dx[i];
            Uc[k * m + a] = t;
            if (!(std::fabs(t) <= P.opts.max_control_value)) return false;
        }
        rk4_step<double>(P.model_at(k), &Xc[k * n], &Uc[k * m], P.dt[k], &Xc[(k + 1) * n]);
        for (int i = 0; i < n; i++) if (!(std::fabs(Xc[(k + 1) * n + i]) <= P.opts.max_state_value)) return false;
    }
    return true;
}

// Forward pass with backtracking line search (Altro.jl forwardpass!, restated):
//   accept the first alpha in 1, 1/2, ... with  lower < z <= upper  or  J < J_prev,
//   z = (J_prev - J) / expected, expected = -alpha (dV1 + alpha dV2)  (z = -1 when expected <= 0).
// No acceptable alpha after the last trial: keep the trajectory, rho increase + bp_reg_fp.
inline void forward_pass(Problem& P, int b) {
    const int n = P.n, m = P.m, N = P.N;
    double* Xc = &P.Xc[(size_t)b * N * n]; double* Uc = &P.Uc[(size_t)b * (N - 1) * m];
    if (P.bp_status[b] < 0) { P.alpha[b] = 0; P.ls_iters[b] = 0; return; }
    const double J_prev = P.J[b];
    const double dV1 = P.dV[2 * b], dV2 = P.dV[2 * b + 1];
    double alpha = 1.0;
    bool accepted = false;
    int it = 0;
    for (; it <= P.opts.iterations_linesearch; it++, alpha *= 0.5) {
        if (!forward_rollout(P, b, alpha, Xc, Uc)) continue;
        const double J = merit(P, Xc, Uc, P.lamb(b));
        const double expected = -alpha * (dV1 + alpha * dV2);
        const double z = expected > 0 ? (J_prev - J) / expected : -1.0;
        if ((z > P.opts.line_search_lower_bound && z <= P.opts.line_search_upper_bound) || J < J_prev) {
            std::memcpy(P.Xb(b), Xc, sizeof(double) * N * n);
            std::memcpy(P.Ub(b), Uc, sizeof(double) * (N - 1) * m);
            P.J[b] = J; P.alpha[b] = alpha; accepted = true; it++;
            break;
        }
    }
    P.ls_iters[b] = it;
    if (!accepted) {
        P.alpha[b] = 0;
        reg_increase(P.opts, P.rho[b], P.drho[b]);
        P.rho[b] += P.opts.bp_reg_fp;
    }
}

inline void ensure_merit(Problem& P) {
    if (P.J_valid) return;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < P.B; b++) P.J[b] = merit(P, P.Xb(b), P.Ub(b), P.lamb(b));
    P.J_valid = true;
}

// one iLQR iteration for every instance: expansion + backward pass + forward pass
inline void ilqr_step(Problem& P, int iters) {
    ensure_merit(P);
    for (int it = 0; it < iters; it++) {
        P.opts.noise_epoch++;
#pragma omp parallel for schedule(dynamic, 4)
        for (int b = 0; b < P.B; b++) {
            expand_dynamics(P, b);
            backward_pass(P, b);
            forward_pass(P, b);
        }
    }
}

// AL outer update (Altro.jl dual_update! / penalty_update!, conic form): lambda <- Pi_{K*}(lambda - mu c), mu <- min(mu*phi, mu_max)
inline void al_update(Problem& P) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < P.B; b++) {
        std::vector<double> c, lbar;
        for (size_t ci = 0; ci < P.cons.size(); ci++) {
            const Constraint& con = P.cons[ci];
            const int p = con.p; const double mu = P.mu[ci];
            c.assign((size_t)con.nknots() * p, 0.0); lbar.resize(p);
            evaluate_constraints(P, (int)ci, P.Xb(b), P.Ub(b), c.data());
            double* l = P.lamb(b) + P.con_offset[ci];
            for (int k = 0; k < con.nknots(); k++) {
                for (int i = 0; i < p; i++) lbar[i] = l[k * p + i] - mu * c[k * p + i];
                projection(dualcone(con.sense), lbar.data(), p, &l[k * p]);
                for (int i = 0; i < p; i++) l[k * p + i] = std::max(-P.opts.dual_max, std::min(P.opts.dual_max, l[k * p + i]));
            }
        }
    }
    for (auto& mu : P.mu) mu = std::min(mu * P.opts.penalty_scaling, P.opts.penalty_max);
    for (int b = 0; b < P.B; b++) { P.rho[b] = P.opts.bp_reg_initial; P.drho[b] = 0; }
    P.J_valid = false;
}

}  // namespace oracle
