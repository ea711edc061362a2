// forward.cu -- the forward pass of iLQR: closed-loop RK4 rollout fused with the cost + constraint + AL-penalty
// sweep (kernel 1 without partials + kernel 2), and the per-instance backtracking line search.
//
// What it computes (Altro.jl forwardpass! / rollout!(solver, alpha), restated in oracle/oracle.hpp
// `forward_rollout` / `forward_pass`; it drives the reference's rollout (src/problem.jl:334-340), cost
// (src/objective.jl:89-106) and constraint evaluation (src/abstract_constraint.jl:200-225)):
//     dx = xbar_k - x_k ; ubar_k = u_k + K_k dx + alpha d_k ; xbar_{k+1} = RK4(xbar_k, ubar_k)
//     J(alpha) = sum_k l_k(xbar_k, ubar_k) + AL penalty ;  z = (J_prev - J) / -(alpha (dV1 + alpha dV2))
//     accept the first alpha in 1, 1/2, ..., 2^-ls_iters with  lower < z <= upper  or  J < J_prev.
// The line search is per instance (no communication).
//
// B200 mapping: the recursion is serial in k and an instance has no parallelism worth a warp, so
//   k_forward : one thread per instance tries alpha = 1 (the common case) and writes the candidate trajectory
//               into the instance's spare trajectory buffer;
//   k_ladder  : 16 lanes per instance; for instances that rejected alpha = 1 each lane evaluates one of the
//               remaining step sizes concurrently (same latency as one trial), a ballot picks the first
//               acceptable one -- exactly the sequential backtracking result -- and that lane re-runs its
//               rollout storing the trajectory.  Lane 0 of every group then commits the instance: flip the
//               live buffer + J on acceptance, regularisation increase on failure.
#include "costcon.cuh"
#include "kernels.h"
#include "models.cuh"

namespace {

// closed-loop rollout of one instance for step size alpha; returns the merit, `ok` = no blow-up.
// STORE: write the candidate trajectory.  FAST: diagonal costs + Goal/Bound constraints with x,u in registers.
template <int MODEL, bool STORE, bool FAST>
__device__ __forceinline__ double rollout_merit(const DevProblem& P, int b, double alpha, bool& ok) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m;
    const int N = P.N, buf = P.cur[b];
    const double* X = traj_X(P, buf, b);
    const double* U = traj_U(P, buf, b);
    double* Xc = traj_Xw(P, buf ^ 1, b);
    double* Uc = traj_Uw(P, buf ^ 1, b);
    const double* Kg = P.K + (size_t)b * (N - 1) * n * m;
    const double* dg = P.d + (size_t)b * (N - 1) * m;
    const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
    double x[n], u[m], xn[n];
    double J = 0.0, viol = 0.0;
    ok = true;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = P.x0[(size_t)b * n + i];
    for (int k = 0; k < N; k++) {
        const bool last = (k == N - 1);
        if (!last) {
            double dx[n];
#pragma unroll
            for (int i = 0; i < n; i++) dx[i] = x[i] - X[(size_t)k * n + i];
#pragma unroll
            for (int a = 0; a < m; a++) {
                double t = fma(alpha, dg[(size_t)k * m + a], U[(size_t)k * m + a]);
#pragma unroll
                for (int i = 0; i < n; i++) t = fma(Kg[(size_t)k * n * m + i * m + a], dx[i], t);
                u[a] = t;
                if (!(fabs(t) <= P.opt.max_control_value)) ok = false;
            }
        } else {
#pragma unroll
            for (int a = 0; a < m; a++) u[a] = 0.0;
        }
        if (STORE) {
#pragma unroll
            for (int i = 0; i < n; i++) Xc[(size_t)k * n + i] = x[i];
            if (!last) {
#pragma unroll
                for (int a = 0; a < m; a++) Uc[(size_t)k * m + a] = u[a];
            }
        }
        // ---- cost + AL penalty of knot k ------------------------------------------------------------
        const DevCost& cost = P.costs[P.cost_index[k]];
        if (FAST) {
            double a2 = 0.0, l1 = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) { a2 = fma(cost.Qd[i] * x[i], x[i], a2); l1 = fma(cost.q[i], x[i], l1); }
            double Jk = 0.5 * a2 + l1 + cost.c;
            if (!last) {
                double au = 0.0, lu = 0.0;
#pragma unroll
                for (int i = 0; i < m; i++) { au = fma(cost.Rd[i] * u[i], u[i], au); lu = fma(cost.r[i], u[i], lu); }
                Jk += 0.5 * au + lu;
            }
            double pen = 0.0;
            for (int ci = 0; ci < P.ncon; ci++) {
                const DevCon& con = P.cons[ci];
                if (k + 1 < con.first || k + 1 > con.last) continue;
                const double mu = P.mu[ci];
                const double* lam = lam_b + con.offset + (size_t)(k + 1 - con.first) * con.p;
                double a = 0.0, l2 = 0.0;
                if (con.kind == CON_GOAL) {
#pragma unroll
                    for (int i = 0; i < n; i++) {
                        const int row = con.row_max[i];
                        if (row >= 0) { const double l = lam[row]; const double lp = l - mu * (x[i] - con.a[row]); a = fma(lp, lp, a); l2 = fma(l, l, l2); }
                    }
                } else {   // CON_BOUND
#pragma unroll
                    for (int i = 0; i < n + m; i++) {
                        const double zi = (i < n) ? x[i < n ? i : 0] : u[i >= n ? i - n : 0];
                        int row = con.row_max[i];
                        if (row >= 0) { const double l = lam[row]; const double lp = fmin(0.0, l - mu * (zi - con.a[i])); a = fma(lp, lp, a); l2 = fma(l, l, l2); }
                        row = con.row_min[i];
                        if (row >= 0) { const double l = lam[row]; const double lp = fmin(0.0, l - mu * (con.b[i] - zi)); a = fma(lp, lp, a); l2 = fma(l, l, l2); }
                    }
                }
                pen += (a - l2) / (2 * mu);
            }
            J += Jk + pen;
        } else {
            J += cost_value(cost, n, m, x, u, !last);
            J += al_knot_penalty(P, k + 1, x, u, lam_b, viol);
        }
        if (!last) {
            rk4_step<MODEL, double>(P.params, x, u, P.dt[k], xn);
#pragma unroll
            for (int i = 0; i < n; i++) { x[i] = xn[i]; if (!(fabs(xn[i]) <= P.opt.max_state_value)) ok = false; }
            if (!ok) break;   // Altro stops the rollout at the first blow-up; the trial is rejected
        }
    }
    return J;
}

__device__ __forceinline__ bool ls_accept(const DevProblem& P, double J, double J_prev, double alpha, double dV1, double dV2, bool ok) {
    if (!ok) return false;
    const double expected = -alpha * (dV1 + alpha * dV2);
    const double z = expected > 0.0 ? (J_prev - J) / expected : -1.0;
    return (z > P.opt.ls_lower && z <= P.opt.ls_upper) || (J < J_prev);
}

template <int MODEL, bool FAST>
__global__ void __launch_bounds__(32) k_forward(const DevProblem P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    if (P.bp_status[b] < 0) { P.accepted[b] = 0; return; }
    bool ok;
    const double J = rollout_merit<MODEL, true, FAST>(P, b, 1.0, ok);
    const bool acc = ls_accept(P, J, P.J[b], 1.0, P.dV[2 * b], P.dV[2 * b + 1], ok);
    P.accepted[b] = acc ? 1 : 0;
    if (acc) { P.Jc[b] = J; P.alpha[b] = 1.0; P.ls_iters[b] = 1; }
}

template <int MODEL, bool FAST>
__global__ void __launch_bounds__(128) k_ladder(const DevProblem P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t >> 4, l = t & 15;
    if (b >= P.B) return;                      // whole 16-lane groups leave together
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
    const int status = P.bp_status[b];
    int accepted = P.accepted[b];
    if (status >= 0 && !accepted) {
        const int ntry = P.opt.ls_iters;       // alphas 2^-1 .. 2^-ntry
        const double alpha = ldexp(1.0, -(l + 1));
        bool ok = false, good = false;
        double J = 0.0;
        if (l < ntry) {
            J = rollout_merit<MODEL, false, FAST>(P, b, alpha, ok);
            good = ls_accept(P, J, P.J[b], alpha, P.dV[2 * b], P.dV[2 * b + 1], ok);
        }
        const unsigned votes = __ballot_sync(gmask, good) & gmask;
        if (votes) {
            const int win = __ffs(votes) - 1 - (threadIdx.x & 16);
            if (l == win) {
                bool ok2;
                const double J2 = rollout_merit<MODEL, true, FAST>(P, b, alpha, ok2);
                P.Jc[b] = J2; P.alpha[b] = alpha; P.ls_iters[b] = win + 2;
            }
            accepted = 1;
        }
        __syncwarp(gmask);
    }
    if (l == 0) {
        if (status < 0) { P.alpha[b] = 0.0; P.ls_iters[b] = 0; }
        else if (accepted) { P.cur[b] ^= 1; P.J[b] = P.Jc[b]; P.accepted[b] = 1; }
        else {
            double rho = P.rho[b], drho = P.drho[b];
            reg_increase(P.opt, rho, drho);
            rho += P.opt.bp_reg_fp;
            P.rho[b] = rho; P.drho[b] = drho;
            P.alpha[b] = 0.0; P.ls_iters[b] = P.opt.ls_iters + 1;
        }
    }
}

}  // namespace

cudaError_t launch_forward(const DevProblem& P, cudaStream_t s) {
    const int threads = 32, blocks = (P.B + threads - 1) / threads;
    const bool fast = P.all_diag_cost && P.all_diag_con;
    if (fast) { TO_DISPATCH_MODEL(P.model, P.m, (k_forward<MODEL, true><<<blocks, threads, 0, s>>>(P))); }
    else { TO_DISPATCH_MODEL(P.model, P.m, (k_forward<MODEL, false><<<blocks, threads, 0, s>>>(P))); }
    return cudaGetLastError();
}

cudaError_t launch_ladder(const DevProblem& P, cudaStream_t s) {
    const int threads = 128;
    const long long total = (long long)P.B * 16;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    const bool fast = P.all_diag_cost && P.all_diag_con;
    if (fast) { TO_DISPATCH_MODEL(P.model, P.m, (k_ladder<MODEL, true><<<blocks, threads, 0, s>>>(P))); }
    else { TO_DISPATCH_MODEL(P.model, P.m, (k_ladder<MODEL, false><<<blocks, threads, 0, s>>>(P))); }
    return cudaGetLastError();
}

cudaError_t launch_accept(const DevProblem& P, cudaStream_t s) { return cudaSuccess; }   // folded into k_ladder
