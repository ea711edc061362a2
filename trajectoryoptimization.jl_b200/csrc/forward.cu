#include <cstdlib>
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
// B200 mapping.  The recursion is serial in k and one rollout has no parallelism worth a warp, so the kernel is
// bound by the LATENCY of the per-knot dependency chain (feedback -> 4 dynamics evaluations -> next knot).  Two
// things follow:
//   * backtracking trials are evaluated CONCURRENTLY: a group of G lanes owns one instance and lane j rolls out
//     step size 2^-(trial0+j), writing its candidate into trajectory buffer (cur+1+j) % NBUF.  A ballot picks the
//     first acceptable lane -- exactly the sequential backtracking result -- and acceptance only moves cur[b]
//     (no copy).  Pass 1 (G=4: alpha = 1..1/8) covers ~95% of the instances, pass 2 (G=8) the remaining trials and
//     commits failures (regularisation increase, Altro's bp_reg_fp).
//   * everything off the chain is kept off it: the group's per-knot operands (K_k, d_k, x_k, u_k, lambda_k, ~650 B,
//     identical for all lanes of the group) are prefetched one knot ahead with cp.async (LDGSTS) by the lanes
//     cooperatively into a double-buffered shared-memory stage and read back as broadcasts; cost / constraint
//     descriptors are copied once per CTA into shared memory; x, u live in registers; no fp64 division on the chain.
#include "costcon.cuh"
#include "kernels.h"
#include "models.cuh"

namespace {

constexpr int FWD_MAX_COST = 4;    // cost functions cached in shared memory (more -> read from global)
constexpr int FWD_MAX_N = 512;
constexpr int FWD_THREADS = 32;

struct FwdCon {
    int kind, first, last, p, offset;
    unsigned mask_max, mask_min;     // bit j set: z_j has a finite upper / lower bound (Goal: x_j constrained)
    int ubox;                        // Bound with both bounds finite on every control and none on the state: rows are static
    double mu, inv2mu;
    int row_max[TO_MAXNM], row_min[TO_MAXNM];
    double a[TO_MAXNM], b[TO_MAXNM];
};
struct FwdCost {
    double Qd[TO_MAXN], Rd[TO_MAXM], q[TO_MAXN], r[TO_MAXM], c;
};
struct alignas(16) FwdTab {
    int ncon, ncost_cached, pad0, pad1;
    FwdCon con[TO_MAXCON];
    FwdCost cost[FWD_MAX_COST];
    double dt[FWD_MAX_N];
    int cost_index[FWD_MAX_N];
    int lam_off[FWD_MAX_N][2];       // multipliers to stage for knot k: up to two (offset, count) segments of lambda_b
    int lam_cnt[FWD_MAX_N][2];
};

__device__ inline void load_tables(const DevProblem& P, FwdTab& tab) {
    const int t = threadIdx.x, T = blockDim.x;
    if (t == 0) { tab.ncon = P.ncon; tab.ncost_cached = P.ncost <= FWD_MAX_COST ? P.ncost : 0; }
    for (int ci = 0; ci < P.ncon; ci++) {
        const DevCon& c = P.cons[ci];
        FwdCon& f = tab.con[ci];
        if (t == 0) {
            f.kind = c.kind; f.first = c.first; f.last = c.last; f.p = c.p; f.offset = c.offset;
            f.mu = P.mu[ci]; f.inv2mu = 1.0 / (2.0 * P.mu[ci]);
            unsigned mx = 0, mn = 0;
            for (int j = 0; j < TO_MAXNM; j++) { if (c.row_max[j] >= 0) mx |= 1u << j; if (c.row_min[j] >= 0) mn |= 1u << j; }
            f.mask_max = mx; f.mask_min = mn;
            const unsigned ubits = ((1u << P.m) - 1u) << P.n;
            f.ubox = (c.kind == CON_BOUND && mx == ubits && mn == ubits) ? 1 : 0;
        }
        for (int j = t; j < TO_MAXNM; j += T) {
            f.row_max[j] = c.row_max[j]; f.row_min[j] = c.row_min[j];
            f.a[j] = c.a[j]; f.b[j] = c.b[j];
        }
    }
    if (P.ncost <= FWD_MAX_COST)
        for (int ci = 0; ci < P.ncost; ci++) {
            const DevCost& c = P.costs[ci];
            FwdCost& f = tab.cost[ci];
            for (int j = t; j < TO_MAXN; j += T) { f.Qd[j] = c.Qd[j]; f.q[j] = c.q[j]; }
            for (int j = t; j < TO_MAXM; j += T) { f.Rd[j] = c.Rd[j]; f.r[j] = c.r[j]; }
            if (t == 0) f.c = c.c;
        }
    for (int k = t; k < P.N && k < FWD_MAX_N; k += T) {
        tab.dt[k] = (k < P.N - 1) ? P.dt[k] : 0.0; tab.cost_index[k] = P.cost_index[k];
        int ns = 0;
        tab.lam_cnt[k][0] = tab.lam_cnt[k][1] = 0; tab.lam_off[k][0] = tab.lam_off[k][1] = 0;
        for (int ci = 0; ci < P.ncon; ci++) {
            const DevCon& c = P.cons[ci];
            if (k + 1 < c.first || k + 1 > c.last) continue;
            if (ns < 2) { tab.lam_off[k][ns] = c.offset + (k + 1 - c.first) * c.p; tab.lam_cnt[k][ns] = c.p; }
            ns++;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// Depth of the operand ring of the closed-loop rollout: knot k + FWD_STAGES - 1 is in flight while knot k is consumed.  One knot ahead
// (2 stages) is enough: a knot takes ~2 us.  Deeper rings were tried against the slow-down of the late trials next to the expansion
// kernels (r02t): 3 stages no change, 4 / 6 stages cost registers (128 -> 152) and shared memory and make pass 1 37 % slower.
#ifndef TO_FWD_STAGES
#define TO_FWD_STAGES 2
#endif
constexpr int FWD_STAGES = TO_FWD_STAGES;
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

// shared-memory stage of one knot's operands for the IPB instances of a CTA.
// K_k is stored as 16-byte pairs [pair][IPB][2] (when n*m is even), everything else as 8-byte slots [slot][IPB].
template <int n, int m, int IPB, int NE = n>
struct Stage {
    static constexpr int KSLOTS = NE * m;     // K_k is m x NE (NE = n - 1 on the Lie-group error state)
    static constexpr bool K16 = (KSLOTS % 2) == 0;
    static constexpr int OFF_D = KSLOTS, OFF_U = OFF_D + m, OFF_X = OFF_U + m, OFF_L = OFF_X + n;
    static constexpr int LAM_SLOTS = 2 * (n + m);
    static constexpr int NSLOT = OFF_L + LAM_SLOTS;
    static constexpr int DOUBLES = NSLOT * IPB;          // per stage buffer
    __device__ static __forceinline__ int kidx(int e, int g) { return K16 ? ((e >> 1) * IPB + g) * 2 + (e & 1) : e * IPB + g; }
    __device__ static __forceinline__ int sidx(int slot, int g) { return slot * IPB + g; }
};

// lanes of a group cooperatively issue the async copies of knot k's operands of instance g
template <int n, int m, int IPB, int G, int NE>
__device__ __forceinline__ void prefetch_knot(double* base, int g, int l, int k, const double* Kg, const double* dg, const double* X,
                                              const double* U, const double* lam_b, const FwdTab& tab, int N) {
    using S = Stage<n, m, IPB, NE>;
    if (k < N - 1) {
        const double* Kk = Kg + (size_t)k * NE * m;
        if (S::K16) {
#pragma unroll
            for (int c = 0; c < (S::KSLOTS / 2 + G - 1) / G; c++) {
                const int cc = c * G + l;
                if (cc < S::KSLOTS / 2) cp_async16(base + S::kidx(2 * cc, g), Kk + 2 * cc);
            }
        } else {
#pragma unroll
            for (int c = 0; c < (S::KSLOTS + G - 1) / G; c++) {
                const int cc = c * G + l;
                if (cc < S::KSLOTS) cp_async8(base + S::kidx(cc, g), Kk + cc);
            }
        }
        // d, u, x : 2m + n consecutive 8-byte slots
#pragma unroll
        for (int c = 0; c < (2 * m + n + G - 1) / G; c++) {
            const int s = c * G + l;
            if (s < m) cp_async8(base + S::sidx(S::OFF_D + s, g), dg + (size_t)k * m + s);
            else if (s < 2 * m) cp_async8(base + S::sidx(S::OFF_U + s - m, g), U + (size_t)k * m + (s - m));
            else if (s < 2 * m + n) cp_async8(base + S::sidx(S::OFF_X + s - 2 * m, g), X + (size_t)k * n + (s - 2 * m));
        }
    }
    // multipliers of the (at most two) constraints active at knot k+1, packed in constraint order
    {
        const int c0 = tab.lam_cnt[k][0], c1 = tab.lam_cnt[k][1];
        const double* l0 = lam_b + tab.lam_off[k][0];
        const double* l1 = lam_b + tab.lam_off[k][1];
        for (int i = l; i < c0; i += G) cp_async8(base + S::sidx(S::OFF_L + i, g), l0 + i);
        for (int i = l; i < c1; i += G) cp_async8(base + S::sidx(S::OFF_L + c0 + i, g), l1 + i);
    }
}

// closed-loop rollout of instance b (group g of the CTA) with step size alpha, diagonal costs, Goal/Bound constraints.
// The candidate trajectory goes to buffer `cbuf`.  Returns the merit; `ok` = no blow-up.
template <int MODEL, int IPB, int G, bool LIE>
__device__ __forceinline__ double rollout_fast(const DevProblem& P, const FwdTab& tab, double* stage, int b, int g, int l, unsigned gmask,
                                               double alpha, int cbuf, bool& ok, double& viol) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, NE = LIE ? n - 1 : n;
    using S = Stage<n, m, IPB, NE>;
    const int N = P.N, buf = P.cur[b];
    const double* X = traj_X(P, buf, b);
    const double* U = traj_U(P, buf, b);
    double* Xc = traj_Xw(P, cbuf, b);
    double* Uc = traj_Uw(P, cbuf, b);
    const double* Kg = P.K + (size_t)b * (N - 1) * NE * m;
    const double* dg = P.d + (size_t)b * (N - 1) * m;
    const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
    double x[n], u[m], xn[n];
    double J = 0.0;
    ok = true; viol = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = P.x0[(size_t)b * n + i];
    constexpr int D = FWD_STAGES - 1;           // prefetch distance in knots
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (j < N) prefetch_knot<n, m, IPB, G, NE>(stage + j * S::DOUBLES, g, l, j, Kg, dg, X, U, lam_b, tab, N);
        cp_async_commit();
    }
    int sb = 0, sp = D;                         // stage of knot k, stage knot k + D goes to (= the one knot k - 1 has just left)
    for (int k = 0; k < N; k++) {
        const bool last = (k == N - 1);
        __syncwarp(gmask);                      // every lane of the group is done reading the stage of knot k-1
        if (k + D < N) prefetch_knot<n, m, IPB, G, NE>(stage + sp * S::DOUBLES, g, l, k + D, Kg, dg, X, U, lam_b, tab, N);
        cp_async_commit();
        cp_async_wait<D>();                     // this lane's copies for knot k have landed ...
        __syncwarp(gmask);                      // ... and so have the other lanes'
        const double* st = stage + sb * S::DOUBLES;
        sb = (sb + 1 == FWD_STAGES) ? 0 : sb + 1; sp = (sp + 1 == FWD_STAGES) ? 0 : sp + 1;
        if (!last) {
#pragma unroll
            for (int a = 0; a < m; a++) u[a] = fma(alpha, st[S::sidx(S::OFF_D + a, g)], st[S::sidx(S::OFF_U + a, g)]);
            double dxe[NE];   // RD.state_diff(xbar_k, x_k): plain difference, or the Cayley error of the attitude (LIE)
            if constexpr (LIE) {
                double xr[n];
#pragma unroll
                for (int i = 0; i < n; i++) xr[i] = st[S::sidx(S::OFF_X + i, g)];
                state_diff(true, n, 3, x, xr, dxe);
            } else {
#pragma unroll
                for (int i = 0; i < n; i++) dxe[i] = x[i] - st[S::sidx(S::OFF_X + i, g)];
            }
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const double dx = dxe[i];
                if (S::K16 && (m % 2 == 0)) {
#pragma unroll
                    for (int a = 0; a < m; a += 2) {
                        const double2 kv = *reinterpret_cast<const double2*>(&st[S::kidx(i * m + a, g)]);
                        u[a] = fma(kv.x, dx, u[a]);
                        u[a + 1] = fma(kv.y, dx, u[a + 1]);
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < m; a++) u[a] = fma(st[S::kidx(i * m + a, g)], dx, u[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < m; a++) if (!(fabs(u[a]) <= P.opt.max_control_value)) ok = false;
        } else {
#pragma unroll
            for (int a = 0; a < m; a++) u[a] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < n; i++) Xc[(size_t)k * n + i] = x[i];
        if (!last) {
#pragma unroll
            for (int a = 0; a < m; a++) Uc[(size_t)k * m + a] = u[a];
        }
        // ---- cost of knot k (DiagonalCost) -------------------------------------------------------------
        {
            const int cid = tab.cost_index[k];
            double a2 = 0.0, l1 = 0.0, cc;
            if (tab.ncost_cached) {
                const FwdCost& c = tab.cost[cid];
#pragma unroll
                for (int i = 0; i < n; i++) { a2 = fma(c.Qd[i] * x[i], x[i], a2); l1 = fma(c.q[i], x[i], l1); }
                if (!last) {
#pragma unroll
                    for (int i = 0; i < m; i++) { a2 = fma(c.Rd[i] * u[i], u[i], a2); l1 = fma(c.r[i], u[i], l1); }
                }
                cc = c.c;
            } else {
                const DevCost& c = P.costs[cid];
#pragma unroll
                for (int i = 0; i < n; i++) { a2 = fma(c.Qd[i] * x[i], x[i], a2); l1 = fma(c.q[i], x[i], l1); }
                if (!last) {
#pragma unroll
                    for (int i = 0; i < m; i++) { a2 = fma(c.Rd[i] * u[i], u[i], a2); l1 = fma(c.r[i], u[i], l1); }
                }
                cc = c.c;
            }
            J += fma(0.5, a2, l1) + cc;
        }
        // ---- AL penalty of knot k (Goal / Bound): (|Pi_K*(lambda - mu c)|^2 - |lambda|^2) / (2 mu) -------------
        {
            int slot = 0;
            for (int ci = 0; ci < tab.ncon; ci++) {
                const FwdCon& c = tab.con[ci];
                if (k + 1 < c.first || k + 1 > c.last) continue;
                const double mu = c.mu;
                const int lo = S::OFF_L + slot;
                double a = 0.0, l2 = 0.0;
                if (c.kind == CON_GOAL) {
                    const unsigned mk = c.mask_max;
#pragma unroll
                    for (int i = 0; i < n; i++) {
                        if (mk & (1u << i)) {
                            const int row = c.row_max[i];
                            const double lm = st[S::sidx(lo + row, g)];
                            const double cv = x[i] - c.a[row];
                            const double lp = fma(-mu, cv, lm);
                            a = fma(lp, lp, a); l2 = fma(lm, lm, l2); viol = fmax(viol, fabs(cv));
                        }
                    }
                } else if (c.ubox) {
                    // u_min <= u <= u_max on every control: rows 0..m-1 = upper, m..2m-1 = lower (src/constraints.jl:738-755)
#pragma unroll
                    for (int i = 0; i < m; i++) {
                        const double lu = st[S::sidx(lo + i, g)], ll = st[S::sidx(lo + m + i, g)];
                        const double cu = u[i] - c.a[n + i], cl = c.b[n + i] - u[i];
                        const double pu = fmin(0.0, fma(-mu, cu, lu)), pl = fmin(0.0, fma(-mu, cl, ll));
                        a = fma(pu, pu, a); a = fma(pl, pl, a); l2 = fma(lu, lu, l2); l2 = fma(ll, ll, l2);
                        viol = fmax(viol, fmax(cu, cl));
                    }
                } else {
                    const unsigned mx = c.mask_max, mn = c.mask_min;
                    if ((mx | mn) & ((1u << n) - 1u)) {
#pragma unroll
                        for (int i = 0; i < n; i++) {
                            if (mx & (1u << i)) { const double lm = st[S::sidx(lo + c.row_max[i], g)]; const double cv = x[i] - c.a[i]; const double lp = fmin(0.0, fma(-mu, cv, lm)); a = fma(lp, lp, a); l2 = fma(lm, lm, l2); viol = fmax(viol, cv); }
                            if (mn & (1u << i)) { const double lm = st[S::sidx(lo + c.row_min[i], g)]; const double cv = c.b[i] - x[i]; const double lp = fmin(0.0, fma(-mu, cv, lm)); a = fma(lp, lp, a); l2 = fma(lm, lm, l2); viol = fmax(viol, cv); }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < m; i++) {
                        if (mx & (1u << (n + i))) { const double lm = st[S::sidx(lo + c.row_max[n + i], g)]; const double cv = u[i] - c.a[n + i]; const double lp = fmin(0.0, fma(-mu, cv, lm)); a = fma(lp, lp, a); l2 = fma(lm, lm, l2); viol = fmax(viol, cv); }
                        if (mn & (1u << (n + i))) { const double lm = st[S::sidx(lo + c.row_min[n + i], g)]; const double cv = c.b[n + i] - u[i]; const double lp = fmin(0.0, fma(-mu, cv, lm)); a = fma(lp, lp, a); l2 = fma(lm, lm, l2); viol = fmax(viol, cv); }
                    }
                }
                J = fma(a - l2, c.inv2mu, J);
                slot += c.p;
            }
        }
        if (!last) {
            rk4_step<MODEL, double>(model_params<MODEL>(P, k), x, u, tab.dt[k], xn);
#pragma unroll
            for (int i = 0; i < n; i++) { x[i] = xn[i]; if (!(fabs(xn[i]) <= P.opt.max_state_value)) ok = false; }
            // a blown-up trial keeps integrating (the group stays in lock step); its result is rejected through `ok`
        }
    }
    cp_async_wait<0>();
    return J;
}

// generic path (dense costs or general constraints): pointer-based evaluation, operands read directly from global
template <int MODEL, bool LIE>
__device__ __forceinline__ double rollout_generic(const DevProblem& P, int b, double alpha, int cbuf, bool& ok, double& viol) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, NE = LIE ? n - 1 : n;
    const int N = P.N, buf = P.cur[b];
    const double* X = traj_X(P, buf, b);
    const double* U = traj_U(P, buf, b);
    double* Xc = traj_Xw(P, cbuf, b);
    double* Uc = traj_Uw(P, cbuf, b);
    const double* Kg = P.K + (size_t)b * (N - 1) * NE * m;
    const double* dg = P.d + (size_t)b * (N - 1) * m;
    const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
    double x[n], u[m], xn[n];
    double J = 0.0;
    ok = true; viol = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = P.x0[(size_t)b * n + i];
    for (int k = 0; k < N; k++) {
        const bool last = (k == N - 1);
        if (!last) {
#pragma unroll
            for (int a = 0; a < m; a++) u[a] = fma(alpha, dg[(size_t)k * m + a], U[(size_t)k * m + a]);
            double dxe[NE], xr[n];
#pragma unroll
            for (int i = 0; i < n; i++) xr[i] = X[(size_t)k * n + i];
            state_diff(LIE, n, 3, x, xr, dxe);
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const double dx = dxe[i];
#pragma unroll
                for (int a = 0; a < m; a++) u[a] = fma(Kg[(size_t)k * NE * m + i * m + a], dx, u[a]);
            }
#pragma unroll
            for (int a = 0; a < m; a++) if (!(fabs(u[a]) <= P.opt.max_control_value)) ok = false;
        } else {
#pragma unroll
            for (int a = 0; a < m; a++) u[a] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < n; i++) Xc[(size_t)k * n + i] = x[i];
        if (!last) {
#pragma unroll
            for (int a = 0; a < m; a++) Uc[(size_t)k * m + a] = u[a];
        }
        J += cost_value(P.costs[P.cost_index[k]], n, m, x, u, !last);
        J += al_knot_penalty(P, k + 1, x, u, lam_b, viol);
        if (!last) {
            rk4_step<MODEL, double>(model_params<MODEL>(P, k), x, u, P.dt[k], xn);
#pragma unroll
            for (int i = 0; i < n; i++) { x[i] = xn[i]; if (!(fabs(xn[i]) <= P.opt.max_state_value)) ok = false; }
            if (!ok) break;
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

extern __shared__ __align__(16) unsigned char fwd_smem[];

// One line-search pass: lane l of group g evaluates trial (trial0 + l) of instance b.
//   first_pass : ignore / reset accepted[b];   final_pass : commit failures (no acceptable step size).
template <int MODEL, int G, bool FAST, int LANES, bool LIE>
__global__ void __launch_bounds__(FWD_THREADS) k_linesearch(const DevProblem P, int trial0, int first_pass, int final_pass) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, NE = LIE ? n - 1 : n;
    // LANES = 16: only half of the warp carries groups.  The pass is a latency-bound FP64 chain at ~4 warps per SM, and
    // an FP64 instruction of a half-empty warp takes one pipe pass instead of two (profiles/r01_notes.md).
    constexpr int IPB = LANES / G;
    using S = Stage<n, m, IPB, NE>;
    FwdTab* tab = reinterpret_cast<FwdTab*>(fwd_smem);
    double* stage = reinterpret_cast<double*>(fwd_smem + sizeof(FwdTab));
    const int g = (threadIdx.x % LANES) / G, l = threadIdx.x % G;
    // pass 1 walks every instance; the later passes walk the list pass 1 left of the instances it did not accept, so that only CTAs with work
    // stay resident next to the kernels of the main stream (a CTA with one late instance of four used to hold its registers for the whole pass)
    int b = blockIdx.x * IPB + g;
    if (!first_pass && P.late_list) b = (b < *P.late_count) ? P.late_list[b] : P.B;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));
    const bool valid = b < P.B && threadIdx.x < LANES;
    const int status = valid ? P.bp_status[b] : -1;
    const int was_accepted = (valid && !first_pass) ? P.accepted[b] : 0;
    const bool work = valid && status >= 0 && !was_accepted && trial0 <= P.opt.ls_iters;
    // CTA-uniform: skip the table load when no group of this CTA has work
    const unsigned any = __ballot_sync(0xffffffffu, work);
    if (FAST && any) load_tables(P, *tab);
    if (!valid) return;
    bool accepted = was_accepted != 0;
    if (work) {
        const int trial = trial0 + l;
        const double alpha = ldexp(1.0, -trial);
        const int cbuf = (P.cur[b] + 1 + l) % TO_NBUF;
        bool ok = false;
        double J, viol = 0.0;
        if (FAST) J = rollout_fast<MODEL, IPB, G, LIE>(P, *tab, stage, b, g, l, gmask, alpha, cbuf, ok, viol);
        else J = rollout_generic<MODEL, LIE>(P, b, alpha, cbuf, ok, viol);
        const bool good = (trial <= P.opt.ls_iters) && ls_accept(P, J, P.J[b], alpha, P.dV[2 * b], P.dV[2 * b + 1], ok);
        const unsigned votes = __ballot_sync(gmask, good) & gmask;
        if (votes) {
            const int win = __ffs(votes) - 1 - g * G;
            if (l == win) {
                P.cur[b] = cbuf; P.J[b] = J; P.viol[b] = viol; P.alpha[b] = alpha; P.ls_iters[b] = trial + 1; P.accepted[b] = 1;
            }
            accepted = true;
        }
    }
    if (l == 0 && first_pass) {
        P.acc1[b] = accepted ? 1 : 0;
        if (!accepted && P.late_list) P.late_list[atomicAdd(P.late_count, 1)] = b;
    }
    if (l == 0 && !accepted) {
        if (first_pass) P.accepted[b] = 0;
        if (status < 0) { P.alpha[b] = 0.0; P.ls_iters[b] = 0; }
        else if (final_pass) {   // no acceptable step: keep the trajectory, raise the regularisation (Altro forwardpass!)
            double rho = P.rho[b], drho = P.drho[b];
            reg_increase(P.opt, rho, drho);
            rho += P.opt.bp_reg_fp;
            P.rho[b] = rho; P.drho[b] = drho;
            P.alpha[b] = 0.0; P.ls_iters[b] = P.opt.ls_iters + 1;
        }
    }
    (void)sizeof(S);
}

template <int MODEL, int G, bool FAST, int LANES, bool LIE = false>
cudaError_t launch_pass_l(const DevProblem& P, int trial0, int first_pass, int final_pass, cudaStream_t s) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, NE = LIE ? n - 1 : n;
    constexpr int IPB = LANES / G;
    const int blocks = (P.B + IPB - 1) / IPB;
    const size_t smem = FAST ? sizeof(FwdTab) + (size_t)FWD_STAGES * Stage<n, m, IPB, NE>::DOUBLES * sizeof(double) : 0;
    auto kern = k_linesearch<MODEL, G, FAST, LANES, LIE>;
    static bool configured[TO_MAXDEV] = {false};
    const int dev = current_device_slot();
    if (!configured[dev] && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    configured[dev] = true;
    { static bool done[TO_MAXDEV] = {false}; prefer_common_carveout(kern, done); }
    kern<<<blocks, FWD_THREADS, smem, s>>>(P, trial0, first_pass, final_pass);
    return cudaGetLastError();
}

template <int MODEL, int G, bool FAST>
cudaError_t launch_pass(const DevProblem& P, int trial0, int first_pass, int final_pass, cudaStream_t s) {
    // lanes of each warp that carry groups: 16 for the first pass (half-warp FP64 instructions take one pipe pass, profiles/r01_notes.md); the later passes
    // use 16 when they walk the compact late list (two-instance CTAs, see to_create) and 32 when they scan all instances.  TO_FWD_LANES_P1 / _P2 override.
    static int lanes1 = -1, lanes2 = -1;
    if (lanes1 < 0) { const char* v = getenv("TO_FWD_LANES_P1"); lanes1 = v ? atoi(v) : 16; v = getenv("TO_FWD_LANES_P2"); lanes2 = v ? atoi(v) : 0; }
    const int lanes = first_pass ? lanes1 : (lanes2 ? lanes2 : (P.late_list ? 16 : 32));
    if constexpr (MODEL == MODEL_QUADROTOR) {   // Lie-group error state: dx = state_diff(xbar, x), gains m x (n - 1)
        if (P.lie) {
            if (G <= 16 && lanes == 16) return launch_pass_l<MODEL, G, FAST, (G <= 16 ? 16 : 32), true>(P, trial0, first_pass, final_pass, s);
            return launch_pass_l<MODEL, G, FAST, 32, true>(P, trial0, first_pass, final_pass, s);
        }
    }
    if (G <= 16 && lanes == 16) return launch_pass_l<MODEL, G, FAST, (G <= 16 ? 16 : 32)>(P, trial0, first_pass, final_pass, s);
    return launch_pass_l<MODEL, G, FAST, 32>(P, trial0, first_pass, final_pass, s);
}

bool fast_path(const DevProblem& P) {
    return P.all_diag_cost && P.all_diag_con && P.N <= FWD_MAX_N && P.max_p_knot <= 2 * (P.n + P.m) && P.max_cons_knot <= 2;
}

}  // namespace

// pass 1: trials 0..3 (alpha = 1, 1/2, 1/4, 1/8), 4 lanes per instance.
// (Measured alternative: the whole ladder in one 16-lane pass costs 1.34 ms -- 11x the FLOPs, the uncoalesced candidate
//  stores saturate the LSU -- against 0.33 + 0.25 ms for the two latency-bound passes; profiles/r01_notes.md.)
cudaError_t launch_forward(const DevProblem& P, cudaStream_t s) {
    cudaError_t e = cudaErrorNotSupported;
    const int final_pass = P.opt.ls_iters < 4;
    if (P.late_list) { e = cudaMemsetAsync(P.late_count, 0, sizeof(int), s); if (e != cudaSuccess) return e; e = cudaErrorNotSupported; }
    if (fast_path(P)) { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_pass<MODEL, 4, true>(P, 0, 1, final_pass, s))); }
    else { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_pass<MODEL, 4, false>(P, 0, 1, final_pass, s))); }
    return e;
}

// pass 2 (+3 when ls_iters > 11): the remaining trials, 8 lanes per instance; commits failures
cudaError_t launch_ladder(const DevProblem& P, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    for (int trial0 = 4; trial0 <= P.opt.ls_iters && e == cudaSuccess; trial0 += 8) {
        const int final_pass = trial0 + 8 > P.opt.ls_iters;
        if (fast_path(P)) { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_pass<MODEL, 8, true>(P, trial0, 0, final_pass, s))); }
        else { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_pass<MODEL, 8, false>(P, trial0, 0, final_pass, s))); }
    }
    return e;
}

cudaError_t launch_accept(const DevProblem& P, cudaStream_t s) { return cudaSuccess; }   // acceptance is committed inside k_linesearch
