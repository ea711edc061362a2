// costcon.cuh -- per-knot cost / constraint / cone device functions (pointer based, runtime n and m).
// Every function is a pure map over ONE knot z = [x;u] (reference docs/src/costfunctions.md:16) so any
// (instance, knot) pair can be evaluated by any thread.
//
// Mirrors (reference file:line):
//   cost_value        RD.evaluate(::QuadraticCostFunction,x,u)   src/cost_functions.jl:89-104
//   cost_gradient     RD.gradient!                               src/cost_functions.jl:137-172
//   cost_hessian      RD.hessian!                                src/cost_functions.jl:212-233
//   con_evaluate      RD.evaluate(con, z)                        src/constraints.jl:55-61 (Goal), :738-755 (Bound),
//                                                                :135-139 (Linear), :190-199 (Circle), :278-290 (Sphere), :462-465 (Norm)
//   con_jacobian      RD.jacobian!(con, jac, c, z)               src/constraints.jl:62-68, :757-765, :140-144, :201-213, :292-306, :493-517
//   cone_projection / cone_grad_projection / cone_hess_projection  src/cones.jl:96-127, :129-188, :201-276
//   dualcone          src/cones.jl:65-69
#pragma once
#include "common.cuh"

__device__ __forceinline__ int dualcone(int cone) {
    return cone == CONE_IDENTITY ? CONE_ZERO : (cone == CONE_ZERO ? CONE_IDENTITY : cone);
}

__device__ inline double cost_value(const DevCost& c, int n, int m, const double* x, const double* u, bool has_u) {
    double J = 0;
    if (c.diag) {
        double a = 0, l = 0;
        for (int i = 0; i < n; i++) { a = fma(c.Qd[i] * x[i], x[i], a); l = fma(c.q[i], x[i], l); }
        J = 0.5 * a + l + c.c;
        if (has_u) {
            double au = 0, lu = 0;
            for (int i = 0; i < m; i++) { au = fma(c.Rd[i] * u[i], u[i], au); lu = fma(c.r[i], u[i], lu); }
            J += 0.5 * au + lu;
        }
        return J;
    }
    for (int j = 0; j < n; j++) {
        double qx = 0;
        for (int i = 0; i < n; i++) qx = fma(c.Q[j * n + i], x[i], qx);
        J = fma(0.5 * qx, x[j], J);
    }
    double lin = 0;
    for (int i = 0; i < n; i++) lin = fma(c.q[i], x[i], lin);
    J += lin + c.c;
    if (has_u) {
        double Ju = 0, linu = 0;
        for (int j = 0; j < m; j++) {
            double ru = 0;
            for (int i = 0; i < m; i++) ru = fma(c.R[j * m + i], u[i], ru);
            Ju = fma(0.5 * ru, u[j], Ju);
        }
        for (int i = 0; i < m; i++) linu = fma(c.r[i], u[i], linu);
        J += Ju + linu;
        if (!c.zeroH) {
            double h = 0;
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) h = fma(u[i] * c.H[j * m + i], x[j], h);
            J += h;
        }
    }
    return J;
}

// grad[n+m]; the u-part is left untouched at the terminal knot (the reference skips it when is_terminal(z))
__device__ inline void cost_gradient(const DevCost& c, int n, int m, const double* x, const double* u, bool is_terminal, double* grad) {
    for (int i = 0; i < n; i++) {
        double g = c.q[i];
        if (c.diag) g = fma(c.Qd[i], x[i], g);
        else for (int j = 0; j < n; j++) g = fma(c.Q[j * n + i], x[j], g);
        grad[i] = g;
    }
    if (!is_terminal) {
        for (int i = 0; i < m; i++) {
            double g = c.r[i];
            if (c.diag) g = fma(c.Rd[i], u[i], g);
            else for (int j = 0; j < m; j++) g = fma(c.R[j * m + i], u[j], g);
            grad[n + i] = g;
        }
        if (!c.zeroH)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) {
                    grad[j] = fma(c.H[j * m + i], u[i], grad[j]);
                    grad[n + i] = fma(c.H[j * m + i], x[j], grad[n + i]);
                }
    }
}

// hess (n+m)x(n+m) col-major, written in full and symmetric (the reference writes only the lower-left H block
// and leaves the rest to the caller's zero initialisation)
__device__ inline void cost_hessian(const DevCost& c, int n, int m, bool is_terminal, double* hess) {
    const int nm = n + m;
    for (int i = 0; i < nm * nm; i++) hess[i] = 0;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) if (!c.diag || i == j) hess[j * nm + i] = c.Q[j * n + i];
    if (!is_terminal) {
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) if (!c.diag || i == j) hess[(n + j) * nm + n + i] = c.R[j * m + i];
        if (!c.zeroH)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++) { hess[j * nm + n + i] = c.H[j * m + i]; hess[(n + i) * nm + j] = c.H[j * m + i]; }
    }
}

__device__ __forceinline__ double zget(int n, const double* x, const double* u, int j) { return j < n ? x[j] : u[j - n]; }

__device__ inline void con_evaluate(const DevCon& con, int n, int m, const double* x, const double* u, double* c) {
    switch (con.kind) {
        case CON_GOAL:
            for (int i = 0; i < con.p; i++) c[i] = x[con.inds[i]] - con.a[i];
            break;
        case CON_BOUND: {   // upper block first, then the lower block
            int i = 0;
            for (int r = 0; r < con.n_max; r++, i++) { int j = con.a_max[r]; c[i] = zget(n, x, u, j) - con.a[j]; }
            for (int r = 0; r < con.n_min; r++, i++) { int j = con.a_min[r]; c[i] = con.b[j] - zget(n, x, u, j); }
            break;
        }
        case CON_LINEAR: {
            const double* y = con.flag ? u : x;
            const int w = con.flag ? m : n;
            for (int i = 0; i < con.p; i++) {
                double s = -con.b[i];
                for (int j = 0; j < w; j++) s = fma(con.a[j * con.p + i], y[j], s);
                c[i] = s;
            }
            break;
        }
        case CON_CIRCLE:
            for (int i = 0; i < con.p; i++) {
                double dx = x[con.inds[0]] - con.a[i], dy = x[con.inds[1]] - con.b[i];
                c[i] = -(dx * dx) - (dy * dy) + con.rad[i] * con.rad[i];
            }
            break;
        case CON_SPHERE:
            for (int i = 0; i < con.p; i++) {
                double dx = x[con.inds[0]] - con.a[i], dy = x[con.inds[1]] - con.b[i], dz = x[con.inds[2]] - con.c3[i];
                c[i] = -(dx * dx) - (dy * dy) - (dz * dz) + con.rad[i] * con.rad[i];
            }
            break;
        case CON_NORM:
            if (con.sense == CONE_SECOND_ORDER) {
                for (int i = 0; i < con.ninds; i++) c[i] = zget(n, x, u, con.inds[i]);
                c[con.ninds] = con.val;
            } else {
                double s = 0;
                for (int i = 0; i < con.ninds; i++) { double z = zget(n, x, u, con.inds[i]); s = fma(z, z, s); }
                c[0] = s - con.val * con.val;
            }
            break;
        case CON_COLLISION: {   // src/constraints.jl:367-376: r^2 - sum_i (x[x1_i] - x[x2_i])^2, accumulated in that order
            const int D = con.ninds / 2;
            double s = con.val * con.val;
            for (int i = 0; i < D; i++) { const double d = x[con.inds[i]] - x[con.inds[D + i]]; s -= d * d; }
            c[0] = s;
            break;
        }
    }
}

// jac: p x (n+m) col-major, fully written
__device__ inline void con_jacobian(const DevCon& con, int n, int m, const double* x, const double* u, double* jac) {
    const int p = con.p, w = n + m;
    for (int i = 0; i < p * w; i++) jac[i] = 0;
    switch (con.kind) {
        case CON_GOAL: for (int i = 0; i < p; i++) jac[con.inds[i] * p + i] = 1; break;
        case CON_BOUND: {
            int i = 0;
            for (int r = 0; r < con.n_max; r++, i++) jac[con.a_max[r] * p + i] = 1;
            for (int r = 0; r < con.n_min; r++, i++) jac[con.a_min[r] * p + i] = -1;
            break;
        }
        case CON_LINEAR: {
            const int off = con.flag ? n : 0, wd = con.flag ? m : n;
            for (int j = 0; j < wd; j++) for (int i = 0; i < p; i++) jac[(off + j) * p + i] = con.a[j * p + i];
            break;
        }
        case CON_CIRCLE:
            for (int i = 0; i < p; i++) {
                jac[con.inds[0] * p + i] = -2 * (x[con.inds[0]] - con.a[i]);
                jac[con.inds[1] * p + i] = -2 * (x[con.inds[1]] - con.b[i]);
            }
            break;
        case CON_SPHERE:
            for (int i = 0; i < p; i++) {
                jac[con.inds[0] * p + i] = -2 * (x[con.inds[0]] - con.a[i]);
                jac[con.inds[1] * p + i] = -2 * (x[con.inds[1]] - con.b[i]);
                jac[con.inds[2] * p + i] = -2 * (x[con.inds[2]] - con.c3[i]);
            }
            break;
        case CON_NORM:
            if (con.sense == CONE_SECOND_ORDER) for (int i = 0; i < con.ninds; i++) jac[con.inds[i] * p + i] = 1;
            else for (int i = 0; i < con.ninds; i++) jac[con.inds[i] * p + 0] = 2 * zget(n, x, u, con.inds[i]);
            break;
        case CON_COLLISION: {   // :378-389 (assignments, as in the reference)
            const int D = con.ninds / 2;
            for (int i = 0; i < D; i++) {
                const double d = x[con.inds[i]] - x[con.inds[D + i]];
                jac[con.inds[i] * p] = -2 * d;
                jac[con.inds[D + i] * p] = 2 * d;
            }
            break;
        }
    }
}

// returns 0, or 1 for the reference's "Invalid second-order cone projection" error branch
__device__ inline int cone_projection(int cone, const double* x, int p, double* px) {
    switch (cone) {
        case CONE_IDENTITY: for (int i = 0; i < p; i++) px[i] = x[i]; return 0;
        case CONE_ZERO: for (int i = 0; i < p; i++) px[i] = 0; return 0;
        case CONE_NEGATIVE_ORTHANT: for (int i = 0; i < p; i++) px[i] = fmin(0.0, x[i]); return 0;
        case CONE_POSITIVE_ORTHANT: for (int i = 0; i < p; i++) px[i] = fmax(0.0, x[i]); return 0;
        case CONE_SECOND_ORDER: {
            double s = x[p - 1], a = 0;
            for (int i = 0; i < p - 1; i++) a = fma(x[i], x[i], a);
            a = sqrt(a);
            if (a <= -s) { for (int i = 0; i < p; i++) px[i] = 0; }
            else if (a <= s) { for (int i = 0; i < p; i++) px[i] = x[i]; }
            else if (a >= fabs(s)) {
                double sc = 0.5 * (1 + s / a);
                for (int i = 0; i < p - 1; i++) px[i] = sc * x[i];
                px[p - 1] = sc * a;
            } else return 1;
            return 0;
        }
    }
    return 1;
}

// J: p x p col-major, fully written
__device__ inline int cone_grad_projection(int cone, const double* x, int p, double* J) {
    for (int i = 0; i < p * p; i++) J[i] = 0;
    switch (cone) {
        case CONE_IDENTITY: for (int i = 0; i < p; i++) J[i * p + i] = 1; return 0;
        case CONE_ZERO: return 0;
        case CONE_NEGATIVE_ORTHANT: for (int i = 0; i < p; i++) J[i * p + i] = x[i] <= 0 ? 1 : 0; return 0;
        case CONE_POSITIVE_ORTHANT: for (int i = 0; i < p; i++) J[i * p + i] = x[i] >= 0 ? 1 : 0; return 0;
        case CONE_SECOND_ORDER: {
            const int n = p;
            double s = x[n - 1], a = 0;
            for (int i = 0; i < n - 1; i++) a = fma(x[i], x[i], a);
            a = sqrt(a);
            if (a <= -s) return 0;
            if (a <= s) { for (int i = 0; i < n; i++) J[i * n + i] = 1; return 0; }
            if (a >= fabs(s)) {
                double c = 0.5 * (1 + s / a);
                for (int i = 0; i < n - 1; i++)
                    for (int j = 0; j < n - 1; j++) {
                        double v = -0.5 * s / (a * a * a) * x[i] * x[j];
                        if (i == j) v += c;
                        J[j * n + i] = v;
                    }
                for (int i = 0; i < n - 1; i++) J[(n - 1) * n + i] = 0.5 * x[i] / a;
                for (int i = 0; i < n - 1; i++) J[i * n + (n - 1)] = ((-0.5 * s / (a * a)) + c / a) * x[i];
                J[(n - 1) * n + (n - 1)] = 0.5;
                return 0;
            }
            return 1;
        }
    }
    return 1;
}

__device__ inline int cone_hess_projection(int cone, const double* x, const double* b, int p, double* hess) {
    for (int i = 0; i < p * p; i++) hess[i] = 0;
    if (cone != CONE_SECOND_ORDER) return 0;
    const int n = p - 1;
    double s = x[n], bs = b[n], a = 0, vbv = 0;
    for (int i = 0; i < n; i++) { a = fma(x[i], x[i], a); vbv = fma(x[i], b[i], vbv); }
    a = sqrt(a);
    if (a <= -s) return 0;
    if (a <= s) return 0;
    if (a > fabs(s)) {
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
    return 1;
}

// AL penalty of one knot (conic form): sum_c (|Pi_{K*}(lambda - mu c)|^2 - |lambda|^2) / (2 mu); also the
// knot's constraint violation |c - Pi_K(c)|_inf.  k1 = 1-based knot.  x/u may be registers, local or global.
__device__ inline double al_knot_penalty(const DevProblem& P, int k1, const double* x, const double* u,
                                         const double* lam_b, double& viol) {
    double pen = 0;
    for (int ci = 0; ci < P.ncon; ci++) {
        const DevCon& con = P.cons[ci];
        if (k1 < con.first || k1 > con.last) continue;
        const double mu = P.mu[ci];
        const double* lam = lam_b + con.offset + (size_t)(k1 - con.first) * con.p;
        double c[TO_MAXP], lbar[TO_MAXP], lp[TO_MAXP];
        con_evaluate(con, P.n, P.m, x, u, c);
        for (int i = 0; i < con.p; i++) lbar[i] = lam[i] - mu * c[i];
        cone_projection(dualcone(con.sense), lbar, con.p, lp);
        double a = 0, l2 = 0;
        for (int i = 0; i < con.p; i++) { a = fma(lp[i], lp[i], a); l2 = fma(lam[i], lam[i], l2); }
        pen += (a - l2) / (2 * mu);
        cone_projection(con.sense, c, con.p, lp);
        for (int i = 0; i < con.p; i++) viol = fmax(viol, fabs(c[i] - lp[i]));
    }
    return pen;
}

// Cost expansion of one knot with the AL terms (Gauss-Newton):
//   grad += -cz' D' lp ; hess += mu cz' D'D cz,  D = grad Pi_{K*}(lambda - mu c), lp = Pi_{K*}(lambda - mu c)
// k0 = 0-based knot.  grad[n+m], hess[(n+m)^2] col-major symmetric.
__device__ inline void al_knot_expansion(const DevProblem& P, int k0, const double* x, const double* u, const double* lam_b,
                                         double* grad, double* hess) {
    const int n = P.n, m = P.m, nm = n + m;
    const bool last = (k0 == P.N - 1);
    const DevCost& cost = P.costs[P.cost_index[k0]];
    for (int i = 0; i < nm; i++) grad[i] = 0;
    cost_gradient(cost, n, m, x, u, last, grad);
    cost_hessian(cost, n, m, last, hess);
    const int lim = last ? n : nm;
    for (int ci = 0; ci < P.ncon; ci++) {
        const DevCon& con = P.cons[ci];
        if (k0 + 1 < con.first || k0 + 1 > con.last) continue;
        const int p = con.p;
        const double mu = P.mu[ci];
        const double* lam = lam_b + con.offset + (size_t)(k0 + 1 - con.first) * p;
        double c[TO_MAXP], lbar[TO_MAXP], lp[TO_MAXP];
        double jac[TO_MAXP * TO_MAXNM], Dm[TO_MAXP * TO_MAXP], tmp[TO_MAXP * TO_MAXNM];
        con_evaluate(con, n, m, x, u, c);
        con_jacobian(con, n, m, x, u, jac);
        for (int i = 0; i < p; i++) lbar[i] = lam[i] - mu * c[i];
        const int dc = dualcone(con.sense);
        cone_projection(dc, lbar, p, lp);
        cone_grad_projection(dc, lbar, p, Dm);
        for (int j = 0; j < nm; j++)
            for (int i = 0; i < p; i++) {
                double s = 0;
                for (int r = 0; r < p; r++) s = fma(Dm[r * p + i], jac[j * p + r], s);
                tmp[j * p + i] = s;
            }
        for (int j = 0; j < lim; j++) {
            double g = 0;
            for (int i = 0; i < p; i++) g = fma(tmp[j * p + i], lp[i], g);
            grad[j] -= g;
            for (int j2 = 0; j2 < lim; j2++) {
                double hsum = 0;
                for (int i = 0; i < p; i++) hsum = fma(tmp[j * p + i], tmp[j2 * p + i], hsum);
                hess[j2 * nm + j] += mu * hsum;
            }
        }
    }
}
