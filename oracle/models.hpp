// oracle/models.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Continuous-time dynamics xdot = f(x,u) of the models the hot path is benchmarked on, written as
// templates over the scalar type so the same text evaluates with `double` (rollout) and with
// forward-mode dual numbers (dynamics Jacobians, what RobotDynamics' ForwardAD does).
//
// Sources restated (cited as /root/reference/<file>:<line>):
//   * Cartpole          docs/src/model.md:32-51 (params mc=1, mp=.2, l=.5, g=9.81 at :27)
//   * Quadrotor         examples/Quadrotor.ipynb cell 4 (params) + cell 8 (forces/moments);
//                       the rigid-body kinematics around them live in RobotDynamics.jl `RigidBody`
//                       (un-vendored; Project.toml:11,20) -> restated from its published formulas:
//                       rdot = v (world-frame velocity, bodyframe=false), qdot = 0.5 * q (x) [0;w],
//                       vdot = F/m, wdot = Jinv*(tau - w x (J w)).  PARITY UNPINNED beyond the
//                       hover KAT of test/internal_api.jl:50-56.
//   * DoubleIntegrator  examples/quickstart.jl:11-23 (2-D: n=4,m=2) and the 1-D variant named in
//                       BASELINE.json configs[0] (n=2,m=1).
//   * Acrobot           appears nowhere in /root/reference; restated from RobotZoo.jl v0.3
//                       `Acrobot` defaults (l=(1,1), m=(1,1), J=m l^2/12, friction c=1, g=9.81).
//                       PARITY UNPINNED.
#pragma once
#include <cmath>
#include <vector>

namespace oracle {

enum ModelId { MODEL_DOUBLE_INTEGRATOR = 0, MODEL_CARTPOLE = 1, MODEL_QUADROTOR = 2, MODEL_ACROBOT = 3, MODEL_EXPR = 4 };

// ---------------------------------------------------------------------------------------------
// Forward-mode dual number with P partials (ForwardDiff.Dual restated).
template <int P>
struct Dual {
    double v;
    double d[P];
    Dual() : v(0) { for (int i = 0; i < P; i++) d[i] = 0; }
    Dual(double a) : v(a) { for (int i = 0; i < P; i++) d[i] = 0; }
};
template <int P> inline Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; r.v = a.v + b.v; for (int i = 0; i < P; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; r.v = a.v - b.v; for (int i = 0; i < P; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a) { Dual<P> r; r.v = -a.v; for (int i = 0; i < P; i++) r.d[i] = -a.d[i]; return r; }
template <int P> inline Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; r.v = a.v * b.v; for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int P> inline Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; double inv = 1.0 / b.v; r.v = a.v * inv; for (int i = 0; i < P; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
template <int P> inline Dual<P> operator+(const Dual<P>& a, double b) { Dual<P> r = a; r.v += b; return r; }
template <int P> inline Dual<P> operator+(double b, const Dual<P>& a) { Dual<P> r = a; r.v += b; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a, double b) { Dual<P> r = a; r.v -= b; return r; }
template <int P> inline Dual<P> operator-(double b, const Dual<P>& a) { Dual<P> r = -a; r.v += b; return r; }
template <int P> inline Dual<P> operator*(const Dual<P>& a, double b) { Dual<P> r; r.v = a.v * b; for (int i = 0; i < P; i++) r.d[i] = a.d[i] * b; return r; }
template <int P> inline Dual<P> operator*(double b, const Dual<P>& a) { return a * b; }
template <int P> inline Dual<P> operator/(const Dual<P>& a, double b) { return a * (1.0 / b); }
template <int P> inline Dual<P> operator/(double a, const Dual<P>& b) { return Dual<P>(a) / b; }
template <int P> inline Dual<P> sin(const Dual<P>& a) { Dual<P> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < P; i++) r.d[i] = c * a.d[i]; return r; }
template <int P> inline Dual<P> cos(const Dual<P>& a) { Dual<P> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < P; i++) r.d[i] = s * a.d[i]; return r; }
// max(0, x) as Julia's generic max(x,y)=ifelse(isless(x,y),y,x) with the constant first: on a tie
// the constant 0 is returned, so the derivative is 0 for x <= 0 and 1 for x > 0.
template <int P> inline Dual<P> relu(const Dual<P>& a) { return a.v > 0 ? a : Dual<P>(0.0); }
inline double relu(double a) { return a > 0 ? a : 0.0; }
// first-order chain rule for the unary functions a recorded dynamics program may use (MODEL_EXPR)
template <int P> inline Dual<P> chain(const Dual<P>& a, double f, double df) { Dual<P> r; r.v = f; for (int i = 0; i < P; i++) r.d[i] = df * a.d[i]; return r; }
inline double chain(double, double f, double) { return f; }
template <class S> inline double value_of(const S& a) { return a.v; }
template <> inline double value_of<double>(const double& a) { return a; }
using std::sin;
using std::cos;

// ---------------------------------------------------------------------------------------------
struct ModelParams {
    int id = MODEL_CARTPOLE;
    int n = 4, m = 1;
    double p[16] = {0};
    const struct DynProg* prog = nullptr;   // MODEL_EXPR: the recorded program of the knot being stepped (Problem::model_at)
    int integrator = 4;   // 4 = RK4 (the reference's default, src/problem.jl:119-123), 3 = RK3 (oracle only: the default of the
                          // TrajectoryOptimization v0.3 / Altro 0.3 versions that produced the recorded outputs of examples/*.ipynb)
};

// One model of a hybrid problem (include/trajopt_b200.h to_dynamics_spec; the reference: `Problem(models::Vector{<:DiscreteDynamics}, ...)`,
// src/problem.jl:36-73, RD.dims(models) src/dynamics.jl:15-31, example test/hybrid_dynamics_model.jl:15-54): the user's `RD.dynamics(model, x, u)`
// as a straight-line program; `discrete` = a jump map applied as is (its output dimension may differ from its state dimension).
struct DynProg {
    int n_in = 0, m_in = 0, n_out = 0, discrete = 0;
    std::vector<int> prog;        // prog_len x {op, a, b}  (to_expr_op)
    std::vector<double> consts;
};

inline ModelParams default_model(int id, int dim = 1) {
    ModelParams mp; mp.id = id;
    switch (id) {
        case MODEL_DOUBLE_INTEGRATOR: mp.n = 2 * dim; mp.m = dim; mp.p[0] = 1.0; break;      // mass
        case MODEL_CARTPOLE: mp.n = 4; mp.m = 1; mp.p[0] = 1.0; mp.p[1] = 0.2; mp.p[2] = 0.5; mp.p[3] = 9.81; break;
        case MODEL_QUADROTOR:
            mp.n = 13; mp.m = 4;
            mp.p[0] = 0.5;                                   // mass
            mp.p[1] = 0.0023; mp.p[2] = 0.0023; mp.p[3] = 0.004;  // J diag
            mp.p[4] = 0.0; mp.p[5] = 0.0; mp.p[6] = -9.81;   // gravity
            mp.p[7] = 0.1750;                                // motor_dist
            mp.p[8] = 1.0;                                   // kf
            mp.p[9] = 0.0245;                                // km
            break;
        case MODEL_EXPR: mp.n = 4; mp.m = 2; break;   // padded dimensions of a recorded hybrid problem
        case MODEL_ACROBOT:
            mp.n = 4; mp.m = 1;
            mp.p[0] = 1.0; mp.p[1] = 1.0;                    // l1,l2
            mp.p[2] = 1.0; mp.p[3] = 1.0;                    // m1,m2
            mp.p[4] = 1.0 / 12.0; mp.p[5] = 1.0 / 12.0;      // J1,J2
            mp.p[6] = 1.0;                                   // friction
            mp.p[7] = 9.81;
            break;
    }
    return mp;
}

// docs/src/model.md:32-51
template <class S>
inline void cartpole_dynamics(const double* p, const S* x, const S* u, S* xd) {
    const double mc = p[0], mp = p[1], l = p[2], g = p[3];
    S s = sin(x[1]), c = cos(x[1]);
    S qd1 = x[2], qd2 = x[3];
    // H = [mc+mp  mp*l*c; mp*l*c  mp*l^2]; C*qd = [-mp*qd2*l*s*qd2; 0]; G = [0; mp*g*l*s]; B = [1;0]
    S h11 = S(mc + mp), h12 = mp * l * c, h22 = S(mp * l * l);
    S r1 = (-mp * l) * (qd2 * s) * qd2 - u[0];   // (C*qd + G - B*u)[1]
    S r2 = (mp * g * l) * s;                     // (C*qd + G - B*u)[2]
    S det = h11 * h22 - h12 * h12;
    // qdd = -H \ r
    S qdd1 = -(h22 * r1 - h12 * r2) / det;
    S qdd2 = -(h11 * r2 - h12 * r1) / det;
    xd[0] = qd1; xd[1] = qd2; xd[2] = qdd1; xd[3] = qdd2;
}

// examples/quickstart.jl:15-20
template <class S>
inline void double_integrator_dynamics(const double* p, int dim, const S* x, const S* u, S* xd) {
    for (int i = 0; i < dim; i++) { xd[i] = x[dim + i]; xd[dim + i] = u[i] / p[0]; }
}

// examples/Quadrotor.ipynb cell 8 (forces :131-150, moments :152-175) + RobotDynamics RigidBody.
// state = [r(3); q(4) scalar-first; v(3) world frame; w(3) body frame]
template <class S>
inline void quadrotor_dynamics(const double* p, const S* x, const S* u, S* xd) {
    const double mass = p[0], J1 = p[1], J2 = p[2], J3 = p[3];
    const double gx = p[4], gy = p[5], gz = p[6], L = p[7], kf = p[8], km = p[9];
    S qw = x[3], qx = x[4], qy = x[5], qz = x[6];
    S wx = x[10], wy = x[11], wz = x[12];
    S F1 = relu(kf * u[0]), F2 = relu(kf * u[1]), F3 = relu(kf * u[2]), F4 = relu(kf * u[3]);
    S Fz = F1 + F2 + F3 + F4;                 // body-frame thrust [0,0,Fz]
    // q*F for (possibly non-unit) quaternion: (w^2 - v'v) r + 2 v (v'r) + 2 w (v x r), r = [0,0,Fz]
    S vv = qx * qx + qy * qy + qz * qz;
    S ww = qw * qw - vv;
    S vr = qz * Fz;
    S Fwx = 2.0 * (qx * vr) + 2.0 * (qw * (qy * Fz));
    S Fwy = 2.0 * (qy * vr) - 2.0 * (qw * (qx * Fz));
    S Fwz = ww * Fz + 2.0 * (qz * vr);
    // moments (body frame)
    S M1 = km * u[0], M2 = km * u[1], M3 = km * u[2], M4 = km * u[3];
    S t1 = L * (F2 - F4), t2 = L * (F3 - F1), t3 = (M1 - M2 + M3 - M4);
    // rdot = v
    xd[0] = x[7]; xd[1] = x[8]; xd[2] = x[9];
    // qdot = 0.5 * q (x) [0; w]
    xd[3] = -0.5 * (qx * wx + qy * wy + qz * wz);
    xd[4] = 0.5 * (qw * wx + qy * wz - qz * wy);
    xd[5] = 0.5 * (qw * wy + qz * wx - qx * wz);
    xd[6] = 0.5 * (qw * wz + qx * wy - qy * wx);
    // vdot = (m g + q F)/m
    xd[7] = (mass * gx + Fwx) / mass;
    xd[8] = (mass * gy + Fwy) / mass;
    xd[9] = (mass * gz + Fwz) / mass;
    // wdot = Jinv (tau - w x (J w))
    S Jw1 = J1 * wx, Jw2 = J2 * wy, Jw3 = J3 * wz;
    xd[10] = (t1 - (wy * Jw3 - wz * Jw2)) / J1;
    xd[11] = (t2 - (wz * Jw1 - wx * Jw3)) / J2;
    xd[12] = (t3 - (wx * Jw2 - wy * Jw1)) / J3;
}

// RobotZoo.jl Acrobot (restated from its published source; absent from /root/reference)
template <class S>
inline void acrobot_dynamics(const double* p, const S* x, const S* u, S* xd) {
    const double l1 = p[0], l2 = p[1], m1 = p[2], m2 = p[3], J1 = p[4], J2 = p[5], fr = p[6], g = p[7];
    S th1 = x[0], th2 = x[1], th1d = x[2], th2d = x[3];
    S c1 = cos(th1), s2 = sin(th2), c2 = cos(th2), c12 = cos(th1 + th2);
    S m11 = (m1 * l1 * l1 + J1 + J2) + m2 * ((l1 * l1 + l2 * l2) + (2.0 * l1 * l2) * c2);
    S m12 = m2 * ((l2 * l2 + J2) + (l1 * l2) * c2);
    S m22 = S(l2 * l2 * m2 + J2);
    S tmp = (l1 * l2 * m2) * s2;
    S b1 = -(2.0 * (th1d * th2d) + th2d * th2d) * tmp;
    S b2 = tmp * (th1d * th1d);
    S f1 = fr * th1d, f2 = fr * th2d;
    S g1 = (((m1 + m2) * l2) * c1 + (m2 * l2) * c12) * g;
    S g2 = (m2 * l2 * g) * c12;
    S r1 = -b1 - g1 - f1;
    S r2 = u[0] - b2 - g2 - f2;
    S det = m11 * m22 - m12 * m12;
    xd[0] = th1d; xd[1] = th2d;
    xd[2] = (m22 * r1 - m12 * r2) / det;
    xd[3] = (m11 * r2 - m12 * r1) / det;
}

// evaluate a recorded program with the scalar type S; outputs = the last n_out instructions, the remaining (padded) slots are zero
template <class S>
inline void expr_dynamics(const DynProg& dp, int n, const S* x, const S* u, S* xd) {
    const int len = (int)dp.prog.size() / 3;
    std::vector<S> reg(len);
    for (int i = 0; i < len; i++) {
        const int op = dp.prog[3 * i], a = dp.prog[3 * i + 1], b = dp.prog[3 * i + 2];
        S r = S(0.0);
        switch (op) {
            case 0: r = S(dp.consts[a]); break;
            case 1: r = x[a]; break;
            case 2: r = u[a]; break;
            case 3: r = reg[a] + reg[b]; break;
            case 4: r = reg[a] - reg[b]; break;
            case 5: r = reg[a] * reg[b]; break;
            case 6: r = reg[a] / reg[b]; break;
            case 7: r = -reg[a]; break;
            case 8: r = sin(reg[a]); break;
            case 9: r = cos(reg[a]); break;
            case 10: { const double e = std::exp(value_of(reg[a])); r = chain(reg[a], e, e); break; }
            case 11: { const double v = value_of(reg[a]); r = chain(reg[a], std::log(v), 1.0 / v); break; }
            case 12: { const double q = std::sqrt(value_of(reg[a])); r = chain(reg[a], q, 0.5 / q); break; }
            case 13: { const double v = value_of(reg[a]), c = dp.consts[b]; r = chain(reg[a], std::pow(v, c), c * std::pow(v, c - 1.0)); break; }
            case 14: { const double t = std::tanh(value_of(reg[a])); r = chain(reg[a], t, 1.0 - t * t); break; }
            case 15: r = reg[a] + dp.consts[b]; break;
            case 16: r = reg[a] * dp.consts[b]; break;
            case 17: r = reg[a] * (1.0 / dp.consts[b]); break;
            case 18: r = S(dp.consts[b]) / reg[a]; break;
            case 19: r = dp.consts[b] - reg[a]; break;
        }
        reg[i] = r;
    }
    for (int i = 0; i < n; i++) xd[i] = (i < dp.n_out) ? reg[len - dp.n_out + i] : S(0.0);
}

template <class S>
inline void dynamics(const ModelParams& mp, const S* x, const S* u, S* xd) {
    switch (mp.id) {
        case MODEL_DOUBLE_INTEGRATOR: double_integrator_dynamics<S>(mp.p, mp.m, x, u, xd); break;
        case MODEL_CARTPOLE: cartpole_dynamics<S>(mp.p, x, u, xd); break;
        case MODEL_QUADROTOR: quadrotor_dynamics<S>(mp.p, x, u, xd); break;
        case MODEL_ACROBOT: acrobot_dynamics<S>(mp.p, x, u, xd); break;
        case MODEL_EXPR: expr_dynamics<S>(*mp.prog, mp.n, x, u, xd); break;
    }
}

constexpr int MAXN = 16;  // largest state dimension the oracle handles
constexpr int MAXM = 8;

// RobotDynamics.jl RK4 (zero-order hold on u), restated:
//   k1 = f(x,u) h; k2 = f(x + k1/2,u) h; k3 = f(x + k2/2,u) h; k4 = f(x + k3,u) h;
//   x+ = x + (k1 + 2 k2 + 2 k3 + k4)/6.       Sole call site: src/problem.jl:338.
// RobotDynamics.jl RK3 (Kutta's third-order rule, zero-order hold): k1 = f(x,u) h; k2 = f(x + k1/2,u) h; k3 = f(x - k1 + 2 k2,u) h;
//   x+ = x + (k1 + 4 k2 + k3)/6.  The integrator argument of the reference's Problem constructor (src/problem.jl:119-123) selects it.
template <class S>
inline void rk3_step(const ModelParams& mp, const S* x, const S* u, double h, S* xn) {
    const int n = mp.n;
    S k1[MAXN], k2[MAXN], k3[MAXN], xt[MAXN];
    dynamics<S>(mp, x, u, k1);
    for (int i = 0; i < n; i++) { k1[i] = k1[i] * h; xt[i] = x[i] + k1[i] * 0.5; }
    dynamics<S>(mp, xt, u, k2);
    for (int i = 0; i < n; i++) { k2[i] = k2[i] * h; xt[i] = x[i] - k1[i] + 2.0 * k2[i]; }
    dynamics<S>(mp, xt, u, k3);
    for (int i = 0; i < n; i++) { k3[i] = k3[i] * h; xn[i] = x[i] + (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0; }
}

template <class S>
inline void rk4_step(const ModelParams& mp, const S* x, const S* u, double h, S* xn) {
    if (mp.id == MODEL_EXPR && mp.prog->discrete) { dynamics<S>(mp, x, u, xn); return; }   // a jump map is already discrete
    if (mp.integrator == 3) { rk3_step<S>(mp, x, u, h, xn); return; }
    const int n = mp.n;
    S k1[MAXN], k2[MAXN], k3[MAXN], k4[MAXN], xt[MAXN];
    dynamics<S>(mp, x, u, k1);
    for (int i = 0; i < n; i++) { k1[i] = k1[i] * h; xt[i] = x[i] + k1[i] * 0.5; }
    dynamics<S>(mp, xt, u, k2);
    for (int i = 0; i < n; i++) { k2[i] = k2[i] * h; xt[i] = x[i] + k2[i] * 0.5; }
    dynamics<S>(mp, xt, u, k3);
    for (int i = 0; i < n; i++) { k3[i] = k3[i] * h; xt[i] = x[i] + k3[i]; }
    dynamics<S>(mp, xt, u, k4);
    for (int i = 0; i < n; i++) { k4[i] = k4[i] * h; xn[i] = x[i] + (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0; }
}

}  // namespace oracle
