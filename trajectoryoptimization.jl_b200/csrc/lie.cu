// lie.cu -- Lie-group (quaternion) error state of the solver kernels and the Riccati pass that reads a
// MATERIALISED per-knot expansion.
//
// What it stands for.  Altro.jl solves rigid-body problems (RobotZoo.Quadrotor is a RobotDynamics `RigidBody` with a
// `LieState`) on the ERROR STATE: n_e = RD.errstate_dim(model) = 12 instead of n = 13.  Per iteration it builds
//     G_k   = errstate_jacobian(model, x_k) = blkdiag(I3, grad-differential(q_k), I6)            (Rotations.jl: L(q) H, 4 x 3)
//     [A_e B_e]_k = G_{k+1}' [A_k G_k | B_k]                                                     (error_expansion!(D, model, G))
//     E_k.x = G_k' lx ;  E_k.xx = G_k' lxx G_k + grad^2-differential(q_k, lx[q]) = ... - (q'lx[q]) I3 ;  E_k.ux = lux G_k
//                                                                                                (error_expansion!(E, Q, model, Z, G))
// runs the Riccati recursion on those, and feeds dx = RD.state_diff(model, xbar, x) (inverse Cayley map of q^-1 (x) qbar)
// through the gains in the forward pass.  None of that arithmetic is under /root/reference; what the reference holds is the
// constraint-side hook (error_expansion! of constraint Jacobians by G, src/abstract_constraint.jl:282-303), the quaternion cost
// DiagonalQuatCost (src/lie_costs.jl:33-95) and the attitude constraint QuatVecEq (src/constraints.jl:938-965) -- both of the
// latter are evaluated by costcon.cuh.  Restatement + finite-difference checks: oracle/oracle.hpp, tests/test_oracle_lie.py.
//
// Kernels
//   k_state_diff        RD.state_diff of every knot against the live trajectory (C ABI to_state_diff)
//   k_error_dynamics    [A_e B_e] from [A B] (rollout.cu k_expand) -- one thread per (instance, knot, column)
//   k_error_expansion   error-state cost + AL expansion from the full-state one (sweep.cu k_al_expansion) -- one thread per
//                       (instance, knot, column)
//   k_riccati_dense     backward pass, one warp per instance, reading [A_e B_e]_k, E_k from HBM: 3.7 KB per knot instead of the
//                       2 KB of the fused kernel (riccati.cu), in exchange for taking ANY cost / constraint type -- the expansion
//                       is whatever the sweep kernels wrote.  First correct version: DFMA on shared-memory operands, lane-strided.
//                       (n_e = 12, n_e + m = 16 tiles the FP64 MMA shapes exactly: the tensor-core variant is the next step.)
#include "costcon.cuh"
#include "kernels.h"

namespace {

inline unsigned nblk(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

__global__ void k_state_diff(const DevProblem P, const double* __restrict__ Xbar, double* __restrict__ dx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * P.N) return;
    const int k = (int)(t % P.N), b = (int)(t / P.N);
    const double* x = traj_X(P, P.cur[b], b) + (size_t)k * P.n;
    double xb[TO_MAXN], xx[TO_MAXN], d[TO_MAXN];
    for (int i = 0; i < P.n; i++) { xb[i] = Xbar[t * P.n + i]; xx[i] = x[i]; }
    state_diff(P.lie != 0, P.n, P.qs, xb, xx, d);
    for (int i = 0; i < P.ne; i++) dx[t * P.ne + i] = d[i];
}

// column e (0 .. ne+m-1) of [A_e B_e]_k.  [A B] is row-major with row stride ldab (common.cuh); the output is col-major ne x (ne+m).
__global__ void __launch_bounds__(128) k_error_dynamics(const DevProblem P) {
    const int n = P.n, m = P.m, ne = P.ne, nme = ne + m, qs = P.qs, ld = P.ldab;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * (P.N - 1) * nme) return;
    const int e = (int)(t % nme);
    const long long bk = t / nme;
    const int k = (int)(bk % (P.N - 1)), b = (int)(bk / (P.N - 1));
    const double* AB = P.AB + ((size_t)b * (P.N - 1) + k) * n * ld;
    const double* X = traj_X(P, P.cur[b], b);
    double col[TO_MAXN];                                   // column e of [A G_k | B]  (n entries)
    if (P.lie && e >= qs && e < qs + 3) {
        double G[12]; quat_G(X + (size_t)k * n + qs, G);
        const double* g = &G[(e - qs) * 4];
        for (int i = 0; i < n; i++) {
            double s = 0;
            for (int r = 0; r < 4; r++) s += AB[i * ld + qs + r] * g[r];
            col[i] = s;
        }
    } else {
        const int j = (!P.lie || e < qs) ? e : e + 1;       // full-state column ([A B] columns n.. are B)
        for (int i = 0; i < n; i++) col[i] = AB[i * ld + j];
    }
    double* out = P.ABe + (((size_t)b * (P.N - 1) + k) * nme + e) * ne;
    if (!P.lie) { for (int i = 0; i < n; i++) out[i] = col[i]; return; }
    double G1[12]; quat_G(X + (size_t)(k + 1) * n + qs, G1);
    for (int i = 0; i < qs; i++) out[i] = col[i];
    for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int r = 0; r < 4; r++) s += G1[c * 4 + r] * col[qs + r];
        out[qs + c] = s;
    }
    for (int i = qs + 4; i < n; i++) out[i - 1] = col[i];
}

// column e of the error-state expansion of knot k from the full-state (grad, hess) in `gfull`, `hfull` ([B][N][nm], [B][N][nm][nm])
__global__ void __launch_bounds__(128) k_error_expansion(const DevProblem P, const double* __restrict__ gfull, const double* __restrict__ hfull,
                                                         double* __restrict__ EG, double* __restrict__ EH) {
    const int n = P.n, m = P.m, nm = n + m, ne = P.ne, nme = ne + m, qs = P.qs;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * P.N * nme) return;
    const int e = (int)(t % nme);
    const long long bk = t / nme;
    const int k = (int)(bk % P.N), b = (int)(bk / P.N);
    const double* g = gfull + bk * nm;
    const double* H = hfull + bk * nm * nm;
    double* ge = EG + bk * nme;
    double* He = EH + bk * nme * nme + (size_t)e * nme;     // column e
    if (!P.lie) {
        for (int i = 0; i < nm; i++) He[i] = H[e * nm + i];
        ge[e] = g[e];
        return;
    }
    const double* q = traj_X(P, P.cur[b], b) + (size_t)k * n + qs;
    double G[12]; quat_G(q, G);
    double col[TO_MAXNM];                                    // column e of hess * blkdiag(E, I)  (nm entries)
    const bool qcol = (e >= qs && e < qs + 3);
    if (qcol) {
        const double* gq = &G[(e - qs) * 4];
        for (int i = 0; i < nm; i++) {
            double s = 0;
            for (int r = 0; r < 4; r++) s += H[(qs + r) * nm + i] * gq[r];
            col[i] = s;
        }
    } else {
        const int j = e < qs ? e : e + 1;
        for (int i = 0; i < nm; i++) col[i] = H[j * nm + i];
    }
    // rows: blkdiag(E, I)' col
    for (int i = 0; i < qs; i++) He[i] = col[i];
    for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int r = 0; r < 4; r++) s += G[c * 4 + r] * col[qs + r];
        He[qs + c] = s;
    }
    for (int i = qs + 4; i < nm; i++) He[i - 1] = col[i];
    double qb = 0;                                           // grad^2-differential: -(q'g_q) on the attitude diagonal
    for (int r = 0; r < 4; r++) qb += q[r] * g[qs + r];
    if (qcol) {
        He[e] -= qb;
        double s = 0;
        for (int r = 0; r < 4; r++) s += G[(e - qs) * 4 + r] * g[qs + r];
        ge[e] = s;
    } else {
        ge[e] = g[e < qs ? e : e + 1];
    }
}

// ---- Riccati backward pass on the materialised expansion: one warp per instance ---------------------------------------------
// Same recursion and restart / regularisation rules as riccati.cu (Altro backwardpass!, oracle/oracle.hpp backward_pass).
template <int NR, int M>
struct DenseSmem {
    static constexpr int NME = NR + M;
    double ab[NR * NME];        // [A_e B_e]_k, col-major NR x NME
    double Q[NME * NME];        // E_k.hess, then Qzz (col-major, full)
    double q[NME];              // E_k.grad, then Qz
    double S[NR * NR];          // cost-to-go Hessian (symmetric, full)
    double s[NR];
    double T[NR * NME];         // S [A B]
    double K[M * (NR + 1)];     // [K | d], row a = control
    double W[M * (NR + 1)];     // Qux - rho K | (unused)
};

template <int NR, int M, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_riccati_dense(const DevProblem P) {
    using SM = DenseSmem<NR, M>;
    constexpr int NME = NR + M;
    __shared__ SM smem[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * WARPS + warp;
    if (b >= P.B) return;
    SM& sm = smem[warp];
    const int N = P.N;
    const double* ABg = P.ABe + (size_t)b * (N - 1) * NR * NME;
    const double* EGg = P.EG + (size_t)b * N * NME;
    const double* EHg = P.EH + (size_t)b * N * NME * NME;
    double* Kg = P.K + (size_t)b * (N - 1) * NR * M;
    double* dg = P.d + (size_t)b * (N - 1) * M;
    double rho = P.rho[b], drho = P.drho[b];
    int restarts = 0;
    bool failed = false;

    for (;;) {
        // terminal knot: S = E_N.xx, s = E_N.x
        {
            const double* H = EHg + (size_t)(N - 1) * NME * NME;
            for (int e = lane; e < NR * NR; e += 32) sm.S[e] = H[(e / NR) * NME + (e % NR)];
            if (lane < NR) sm.s[lane] = EGg[(size_t)(N - 1) * NME + lane];
        }
        __syncwarp();
        double dV1 = 0.0, dV2 = 0.0;
        bool ok = true;
        for (int k = N - 2; k >= 0; k--) {
            {
                const double* src = ABg + (size_t)k * NR * NME;
                for (int e = lane; e < NR * NME; e += 32) sm.ab[e] = src[e];
                const double* H = EHg + (size_t)k * NME * NME;
                for (int e = lane; e < NME * NME; e += 32) sm.Q[e] = H[e];
                if (lane < NME) sm.q[lane] = EGg[(size_t)k * NME + lane];
            }
            __syncwarp();
            // T = S [A B]
            for (int e = lane; e < NR * NME; e += 32) {
                const int i = e % NR, j = e / NR;
                double t = 0.0;
#pragma unroll
                for (int r = 0; r < NR; r++) t = fma(sm.S[r * NR + i], sm.ab[j * NR + r], t);
                sm.T[e] = t;
            }
            __syncwarp();
            // Qzz = lzz + [A B]' T ; Qz = lz + [A B]' s
            for (int e = lane; e < NME * NME; e += 32) {
                const int i = e % NME, j = e / NME;
                double t = sm.Q[e];
#pragma unroll
                for (int r = 0; r < NR; r++) t = fma(sm.ab[i * NR + r], sm.T[j * NR + r], t);
                sm.Q[e] = t;
            }
            if (lane < NME) {
                double t = sm.q[lane];
#pragma unroll
                for (int r = 0; r < NR; r++) t = fma(sm.ab[lane * NR + r], sm.s[r], t);
                sm.q[lane] = t;
            }
            __syncwarp();
            // gains: LDL' of Quu + rho I (every lane factors the same M x M matrix), one lane per column of [Qux | Qu]
            double Quu[M * (M + 1) / 2], Lf[M * (M + 1) / 2], dj[M];
#pragma unroll
            for (int a = 0; a < M; a++)
#pragma unroll
                for (int c = 0; c <= a; c++) Quu[a * (a + 1) / 2 + c] = 0.5 * (sm.Q[(NR + c) * NME + NR + a] + sm.Q[(NR + a) * NME + NR + c]);
#pragma unroll
            for (int j = 0; j < M; j++) {
                double t = Quu[j * (j + 1) / 2 + j] + rho;
#pragma unroll
                for (int r = 0; r < j; r++) t = fma(-Lf[j * (j + 1) / 2 + r] * Lf[j * (j + 1) / 2 + r], dj[r], t);
                if (!(t > 0.0) || !isfinite(t)) ok = false;
                dj[j] = t;
                const double inv = 1.0 / t;
                Lf[j * (j + 1) / 2 + j] = inv;
#pragma unroll
                for (int i = j + 1; i < M; i++) {
                    double v = Quu[i * (i + 1) / 2 + j];
#pragma unroll
                    for (int r = 0; r < j; r++) v = fma(-Lf[i * (i + 1) / 2 + r] * Lf[j * (j + 1) / 2 + r], dj[r], v);
                    Lf[i * (i + 1) / 2 + j] = v * inv;
                }
            }
            if (!ok) break;   // uniform across the warp
            if (lane <= NR) {
                const int c = lane;
                double rhs[M], kc[M];
#pragma unroll
                for (int a = 0; a < M; a++) rhs[a] = (c < NR) ? sm.Q[c * NME + NR + a] : sm.q[NR + a];   // Qux[a][c] | Qu[a]
#pragma unroll
                for (int a = 0; a < M; a++) {
                    double t = -rhs[a];
#pragma unroll
                    for (int r = 0; r < a; r++) t = fma(-Lf[a * (a + 1) / 2 + r], kc[r], t);
                    kc[a] = t;
                }
#pragma unroll
                for (int a = 0; a < M; a++) kc[a] *= Lf[a * (a + 1) / 2 + a];
#pragma unroll
                for (int a = M - 1; a >= 0; a--) {
                    double t = kc[a];
#pragma unroll
                    for (int r = a + 1; r < M; r++) t = fma(-Lf[r * (r + 1) / 2 + a], kc[r], t);
                    kc[a] = t;
                }
#pragma unroll
                for (int a = 0; a < M; a++) { sm.K[a * (NR + 1) + c] = kc[a]; sm.W[a * (NR + 1) + c] = fma(-rho, kc[a], rhs[a]); }
                if (c < NR) {
#pragma unroll
                    for (int a = 0; a < M; a++) Kg[(size_t)k * NR * M + c * M + a] = kc[a];
                } else {
                    double t1 = 0.0, t2 = 0.0;
#pragma unroll
                    for (int a = 0; a < M; a++) {
                        dg[(size_t)k * M + a] = kc[a];
                        t1 = fma(kc[a], rhs[a], t1);
                        double qd = 0.0;
#pragma unroll
                        for (int r = 0; r < M; r++) qd = fma((r <= a) ? Quu[a * (a + 1) / 2 + r] : Quu[r * (r + 1) / 2 + a], kc[r], qd);
                        t2 = fma(0.5 * kc[a], qd, t2);
                    }
                    dV1 += t1; dV2 += t2;
                }
            }
            __syncwarp();
            // S <- Qxx + W'K (symmetrised), s <- Qx + W'd
            for (int e = lane; e < NR * NR; e += 32) {
                const int i = e % NR, j = e / NR;
                if (i > j) continue;
                double v = 0.5 * (sm.Q[j * NME + i] + sm.Q[i * NME + j]);
                double wk = 0.0;
#pragma unroll
                for (int a = 0; a < M; a++) wk += sm.W[a * (NR + 1) + i] * sm.K[a * (NR + 1) + j] + sm.W[a * (NR + 1) + j] * sm.K[a * (NR + 1) + i];
                v = fma(0.5, wk, v);
                sm.S[j * NR + i] = v; sm.S[i * NR + j] = v;
            }
            if (lane < NR) {
                double t = sm.q[lane];
#pragma unroll
                for (int a = 0; a < M; a++) t = fma(sm.W[a * (NR + 1) + lane], sm.K[a * (NR + 1) + NR], t);
                sm.s[lane] = t;
            }
            __syncwarp();
        }
        if (ok) {
            if (lane == NR) { P.dV[2 * b] = dV1; P.dV[2 * b + 1] = dV2; }
            break;
        }
        reg_increase(P.opt, rho, drho);
        restarts++;
        if (rho > P.opt.bp_reg_max) { failed = true; break; }
        __syncwarp();
    }
    if (!failed) reg_decrease(P.opt, rho, drho);
    if (lane == 0) { P.rho[b] = rho; P.drho[b] = drho; P.bp_status[b] = failed ? -1 : restarts; }
}

template <int NR, int M>
cudaError_t launch_dense_t(const DevProblem& P, cudaStream_t s) {
    constexpr int WARPS = 4;
    static_assert(sizeof(DenseSmem<NR, M>) * WARPS <= 48 * 1024, "static shared memory");
    k_riccati_dense<NR, M, WARPS><<<(P.B + WARPS - 1) / WARPS, 32 * WARPS, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_state_diff(const DevProblem& P, const double* Xbar, double* dx, cudaStream_t s) {
    k_state_diff<<<nblk((long long)P.B * P.N, 128), 128, 0, s>>>(P, Xbar, dx);
    return cudaGetLastError();
}
cudaError_t launch_error_dynamics(const DevProblem& P, cudaStream_t s) {
    k_error_dynamics<<<nblk((long long)P.B * (P.N - 1) * (P.ne + P.m), 128), 128, 0, s>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_error_expansion(const DevProblem& P, const double* gfull, const double* hfull, double* EG, double* EH, cudaStream_t s) {
    k_error_expansion<<<nblk((long long)P.B * P.N * (P.ne + P.m), 128), 128, 0, s>>>(P, gfull, hfull, EG, EH);
    return cudaGetLastError();
}
cudaError_t launch_backward_dense(const DevProblem& P, cudaStream_t s) {
    if (P.ne == 12 && P.m == 4) return launch_dense_t<12, 4>(P, s);
    if (P.ne == 13 && P.m == 4) return launch_dense_t<13, 4>(P, s);
    if (P.ne == 4 && P.m == 1) return launch_dense_t<4, 1>(P, s);
    if (P.ne == 4 && P.m == 2) return launch_dense_t<4, 2>(P, s);
    if (P.ne == 2 && P.m == 1) return launch_dense_t<2, 1>(P, s);
    return cudaErrorNotSupported;
}
