// riccati.cu -- kernel 3 of the hot path: the dense Riccati backward pass of iLQR for a batch of instances.
//
// What it computes (Altro.jl backwardpass!, which drives the reference's expansions; restated in
// oracle/oracle.hpp `backward_pass`), per instance, serial in k = N-1 .. 1:
//     Qzz = lzz + [A B]' S [A B]      Qz = lz + [A B]' s          (z = [x;u], lzz/lz = cost + AL expansion)
//     K = -(Quu + rho I)^-1 Qux       d = -(Quu + rho I)^-1 Qu    (Cholesky; non-PD -> rho increase + restart)
//     S <- Qxx + K'Quu K + K'Qux + Qux'K     s <- Qx + K'Quu d + K'Qu + Qux'd     dV += (d'Qu, 1/2 d'Quu d)
// With W = Qux - rho K the update collapses to S <- Qxx + W'K, s <- Qx + W'd (identical algebra, Quu K = -Qux - rho K).
// The cost expansion consumed here is the reference's RD.gradient!/RD.hessian! (src/cost_functions.jl:137-233)
// plus the AL terms built from projection!/grad-projection! (src/cones.jl:96-145) for Goal/Bound constraints
// (src/constraints.jl:55-68, :738-765), whose Jacobians are +-1 selectors, so both are diagonal in z.
//
// B200 mapping
//   * one WARP per instance (the recursion is serial in k; B=4096 gives ~28 instances per SM, so intra-instance
//     parallelism has to fill the FP64 pipe).  Persistent CTAs of one warp pull instances from an atomic queue.
//   * [A B]_k (n x LDAB doubles, contiguous per knot thanks to the instance-major layout) is streamed from HBM
//     by 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx) into a multi-stage shared-memory ring, issued
//     by lane 0 several knots ahead of use -- this is the dominant HBM traffic of the whole iteration.
//   * the n x n / n x (n+m) products run as 2x2 register micro-blocks per lane: per inner index one LDS.128 per
//     operand pair and 4 DFMA per block (DFMA issue is 2 cycles on sm_100, every other instruction costs an
//     issue slot, so operands are fetched as 16-byte pairs and addresses are compile-time immediates).
//     T = S [A B] (with s appended as an extra column), then Qzz|Qz = [A B]' T restricted to the upper blocks.
//     FP64 tensor MMA (DMMA m8n8k4) shares the DFMA pipe on B200 (measured: profiles/microbench) and wastes >50%
//     of its tile on n=13, so it is not used.
//   * Quu is m x m (m <= 8): Cholesky + triangular solves are done per right-hand-side column, one lane per
//     column of [Qux Qu], in registers.
#include "costcon.cuh"
#include "kernels.h"

namespace {

constexpr int even_up(int v) { return (v + 1) & ~1; }

__device__ __forceinline__ double2 lds128(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts128(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 2x2 micro-block outer-product accumulate
__device__ __forceinline__ void fma2x2(double (&acc)[4], const double2& a, const double2& b) {
    acc[0] = fma(a.x, b.x, acc[0]);
    acc[1] = fma(a.x, b.y, acc[1]);
    acc[2] = fma(a.y, b.x, acc[2]);
    acc[3] = fma(a.y, b.y, acc[3]);
}

template <int N_, int M_, int STAGES>
struct RiccatiSmem {
    static constexpr int NM = N_ + M_;
    static constexpr int LDAB = even_up(NM);        // row stride of [A B] (HBM and smem)
    static constexpr int LDT = even_up(NM + 1);     // T / Q row stride: one extra column carries s / Qz
    static constexpr int NP = even_up(N_);          // padded state dim
    static constexpr int LDK = even_up(N_ + 1);     // K|d row stride
    static constexpr int AB_BYTES = N_ * LDAB * 8;
    static constexpr int AB_STRIDE = (AB_BYTES + 127) / 128 * 16;   // doubles, 128-byte aligned stages
    double ab[STAGES][AB_STRIDE];
    double S[NP * NP];
    double T[NP * LDT];
    double Q[LDAB * LDT];
    double K[M_ * LDK];
    double W[M_ * LDK];
    double g[LDT];        // lz (cost + AL gradient), padded
    double h[LDT];        // diag(lzz)
    uint64_t bar[STAGES];
};

template <int N_, int M_, int STAGES>
__global__ void __launch_bounds__(32) k_riccati(const DevProblem P, int* __restrict__ work_counter) {
    using SM = RiccatiSmem<N_, M_, STAGES>;
    constexpr int n = N_, m = M_, NM = SM::NM, LDAB = SM::LDAB, LDT = SM::LDT, NP = SM::NP, LDK = SM::LDK;
    constexpr int RBT = NP / 2, CBT = LDAB / 2;              // T blocks: rows of S x column pairs of [A B]
    constexpr int NBT = RBT * CBT;
    constexpr int RT = (NBT + 31) / 32;
    constexpr int RBQ = LDAB / 2, CBQ = LDT / 2;             // Q blocks (upper: cb >= rb)
    constexpr int NBQ = RBQ * CBQ - RBQ * (RBQ - 1) / 2;
    constexpr int RQ = (NBQ + 31) / 32;
    constexpr int RBS = NP / 2;                              // S blocks (upper)
    constexpr int NBS = RBS * (RBS + 1) / 2;
    constexpr int RS = (NBS + 31) / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int lane = threadIdx.x;
    const int N = P.N;

    // ---- loop-invariant lane -> block assignments -------------------------------------------------------
    int t_rb[RT], t_cb[RT]; bool t_on[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) {
        int id = lane + 32 * r; t_on[r] = id < NBT; if (!t_on[r]) id = NBT - 1;
        t_rb[r] = id / CBT; t_cb[r] = id % CBT;
    }
    int q_rb[RQ], q_cb[RQ]; bool q_on[RQ];
#pragma unroll
    for (int r = 0; r < RQ; r++) {
        int id = lane + 32 * r; q_on[r] = id < NBQ; if (!q_on[r]) id = NBQ - 1;
        int rb = 0, rem = id;                                  // row rb holds CBQ - rb blocks
        while (rem >= CBQ - rb) { rem -= CBQ - rb; rb++; }
        q_rb[r] = rb; q_cb[r] = rb + rem;
    }
    int s_rb[RS], s_cb[RS]; bool s_on[RS];
#pragma unroll
    for (int r = 0; r < RS; r++) {
        int id = lane + 32 * r; s_on[r] = id < NBS; if (!s_on[r]) id = NBS - 1;
        int rb = 0, rem = id;
        while (rem >= RBS - rb) { rem -= RBS - rb; rb++; }
        s_rb[r] = rb; s_cb[r] = rb + rem;
    }

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase_bits = 0;   // per-stage parity of the next completion to wait for

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(work_counter, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= P.B) break;

        const int buf = P.cur[b];
        const double* X = traj_X(P, buf, b);
        const double* U = traj_U(P, buf, b);
        const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
        const double* ABg = P.AB + (size_t)b * (N - 1) * n * LDAB;
        double* Kg = P.K + (size_t)b * (N - 1) * n * m;
        double* dg = P.d + (size_t)b * (N - 1) * m;
        double rho = P.rho[b], drho = P.drho[b];
        int restarts = 0;
        bool failed = false;

        for (;;) {   // regularisation restart loop
            // ---- prologue: start streaming the last STAGES knots ----------------------------------------
            if (lane == 0) {
#pragma unroll
                for (int s = 0; s < STAGES; s++) {
                    const int k = N - 2 - s;
                    if (k >= 0) {
                        mbar_expect_tx(&sm.bar[s], SM::AB_BYTES);
                        bulk_g2s(sm.ab[s], ABg + (size_t)k * n * LDAB, SM::AB_BYTES, &sm.bar[s]);
                    }
                }
            }
            // ---- terminal knot: S = lxx_N, s = lx_N (cost + AL) ------------------------------------------
            for (int e = lane; e < NP * NP; e += 32) sm.S[e] = 0.0;
            __syncwarp();
            double s_reg = 0.0;   // lane i < n holds s_i
            {
                const DevCost& cost = P.costs[P.cost_index[N - 1]];
                if (lane < n) {
                    const int i = lane;
                    const double xi = X[(size_t)(N - 1) * n + i];
                    double gi = cost.q[i], hi = 0.0;
                    if (cost.diag) { gi = fma(cost.Qd[i], xi, gi); hi = cost.Qd[i]; }
                    else {
                        for (int j = 0; j < n; j++) { gi = fma(cost.Q[j * n + i], X[(size_t)(N - 1) * n + j], gi); sm.S[j * NP + i] = cost.Q[j * n + i]; }
                    }
                    for (int ci = 0; ci < P.ncon; ci++) {
                        const DevCon& con = P.cons[ci];
                        if (N < con.first || N > con.last) continue;
                        const double mu = P.mu[ci];
                        const double* lam = lam_b + con.offset + (size_t)(N - con.first) * con.p;
                        if (con.kind == CON_GOAL) {
                            const int row = con.row_max[i];
                            if (row >= 0) { const double lp = lam[row] - mu * (xi - con.a[row]); gi -= lp; hi += mu; }
                        } else if (con.kind == CON_BOUND) {
                            int row = con.row_max[i];
                            if (row >= 0) { const double lb = lam[row] - mu * (xi - con.a[i]); if (lb <= 0) { gi -= lb; hi += mu; } }
                            row = con.row_min[i];
                            if (row >= 0) { const double lb = lam[row] - mu * (con.b[i] - xi); if (lb <= 0) { gi += lb; hi += mu; } }
                        }
                    }
                    s_reg = gi;
                    if (cost.diag) sm.S[i * NP + i] = hi; else sm.S[i * NP + i] += hi;
                }
            }
            __syncwarp();

            double dV1 = 0.0, dV2 = 0.0;   // accumulated by lane n
            bool ok = true;
            int stage = 0;
            int k;
            for (k = N - 2; k >= 0; k--) {
                // ---- cost + AL expansion of knot k: lane i < NM handles z_i (diagonal terms) ------------
                {
                    const DevCost& cost = P.costs[P.cost_index[k]];
                    double gi = 0.0, hi = 0.0;
                    if (lane < NM) {
                        const int i = lane;
                        const double zi = (i < n) ? X[(size_t)k * n + i] : U[(size_t)k * m + (i - n)];
                        if (cost.diag) {
                            if (i < n) { gi = fma(cost.Qd[i], zi, cost.q[i]); hi = cost.Qd[i]; }
                            else { gi = fma(cost.Rd[i - n], zi, cost.r[i - n]); hi = cost.Rd[i - n]; }
                        } else {
                            if (i < n) {
                                gi = cost.q[i];
                                for (int j = 0; j < n; j++) gi = fma(cost.Q[j * n + i], X[(size_t)k * n + j], gi);
                                if (!cost.zeroH) for (int a = 0; a < m; a++) gi = fma(cost.H[i * m + a], U[(size_t)k * m + a], gi);
                            } else {
                                const int a = i - n;
                                gi = cost.r[a];
                                for (int j = 0; j < m; j++) gi = fma(cost.R[j * m + a], U[(size_t)k * m + j], gi);
                                if (!cost.zeroH) for (int j = 0; j < n; j++) gi = fma(cost.H[j * m + a], X[(size_t)k * n + j], gi);
                            }
                        }
                        for (int ci = 0; ci < P.ncon; ci++) {
                            const DevCon& con = P.cons[ci];
                            if (k + 1 < con.first || k + 1 > con.last) continue;
                            const double mu = P.mu[ci];
                            const double* lam = lam_b + con.offset + (size_t)(k + 1 - con.first) * con.p;
                            if (con.kind == CON_GOAL) {
                                const int row = (i < n) ? con.row_max[i] : -1;
                                if (row >= 0) { const double lp = lam[row] - mu * (zi - con.a[row]); gi -= lp; hi += mu; }
                            } else if (con.kind == CON_BOUND) {
                                int row = con.row_max[i];
                                if (row >= 0) { const double lb = lam[row] - mu * (zi - con.a[i]); if (lb <= 0) { gi -= lb; hi += mu; } }
                                row = con.row_min[i];
                                if (row >= 0) { const double lb = lam[row] - mu * (con.b[i] - zi); if (lb <= 0) { gi += lb; hi += mu; } }
                            }
                        }
                    }
                    if (lane < LDT) { sm.g[lane] = gi; sm.h[lane] = hi; }
                }
                // ---- wait for [A B]_k in the ring ------------------------------------------------------
                mbar_wait(&sm.bar[stage], (phase_bits >> stage) & 1u);
                phase_bits ^= (1u << stage);
                const double* sAB = sm.ab[stage];

                // ---- T = S [A B]  (2x2 blocks) ---------------------------------------------------------
                {
                    double acc[RT][4];
#pragma unroll
                    for (int r = 0; r < RT; r++) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0; }
#pragma unroll
                    for (int j = 0; j < n; j++) {
#pragma unroll
                        for (int r = 0; r < RT; r++) {
                            const double2 a = lds128(&sm.S[j * NP + 2 * t_rb[r]]);
                            const double2 bb = lds128(&sAB[j * LDAB + 2 * t_cb[r]]);
                            fma2x2(acc[r], a, bb);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RT; r++) {
                        if (t_on[r]) {
                            sts128(&sm.T[(2 * t_rb[r]) * LDT + 2 * t_cb[r]], acc[r][0], acc[r][1]);
                            sts128(&sm.T[(2 * t_rb[r] + 1) * LDT + 2 * t_cb[r]], acc[r][2], acc[r][3]);
                        }
                    }
                }
                __syncwarp();
                if (lane < n) sm.T[lane * LDT + NM] = s_reg;   // extra column: s
                __syncwarp();

                // ---- [Qzz | Qz] = [A B]' [T | s] + [lzz | lz]  (upper 2x2 blocks) -----------------------
                {
                    double acc[RQ][4];
#pragma unroll
                    for (int r = 0; r < RQ; r++) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0; }
#pragma unroll
                    for (int j = 0; j < n; j++) {
#pragma unroll
                        for (int r = 0; r < RQ; r++) {
                            const double2 a = lds128(&sAB[j * LDAB + 2 * q_rb[r]]);
                            const double2 bb = lds128(&sm.T[j * LDT + 2 * q_cb[r]]);
                            fma2x2(acc[r], a, bb);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RQ; r++) {
                        const int i0 = 2 * q_rb[r], j0 = 2 * q_cb[r];
                        if (q_rb[r] == q_cb[r]) { acc[r][0] += sm.h[i0]; acc[r][3] += sm.h[i0 + 1]; }
                        if (j0 == NM) { acc[r][0] += sm.g[i0]; acc[r][2] += sm.g[i0 + 1]; }
                        if (j0 + 1 == NM) { acc[r][1] += sm.g[i0]; acc[r][3] += sm.g[i0 + 1]; }
                        if (q_on[r]) {
                            sts128(&sm.Q[i0 * LDT + j0], acc[r][0], acc[r][1]);
                            sts128(&sm.Q[(i0 + 1) * LDT + j0], acc[r][2], acc[r][3]);
                        }
                    }
                }
                __syncwarp();
                // the stage buffer is free: refill it with the knot STAGES steps ahead
                if (lane == 0 && k - STAGES >= 0) {
                    mbar_expect_tx(&sm.bar[stage], SM::AB_BYTES);
                    bulk_g2s(sm.ab[stage], ABg + (size_t)(k - STAGES) * n * LDAB, SM::AB_BYTES, &sm.bar[stage]);
                }
                stage = (stage + 1 == STAGES) ? 0 : stage + 1;

                // dense cost Hessian (QuadraticCost): add the off-diagonal entries of lzz to the upper part of Q
                if (!P.all_diag_cost) {
                    const DevCost& cost = P.costs[P.cost_index[k]];
                    if (!cost.diag) {
                        for (int e = lane; e < NM * NM; e += 32) {
                            const int i = e / NM, j = e % NM;     // need (i,j) with block(i) <= block(j)
                            if (i == j || (i >> 1) > (j >> 1)) continue;
                            double v;
                            if (i < n && j < n) v = cost.Q[j * n + i];
                            else if (i >= n && j >= n) v = cost.R[(j - n) * m + (i - n)];
                            else if (i < n) v = cost.zeroH ? 0.0 : cost.H[i * m + (j - n)];   // (x_i, u_a): H[a][i]
                            else v = cost.zeroH ? 0.0 : cost.H[j * m + (i - n)];
                            sm.Q[i * LDT + j] += v;
                        }
                        // diagonal: sm.h carried only the AL part for dense costs -> add Q_ii / R_aa
                        if (lane < NM) sm.Q[lane * LDT + lane] += (lane < n) ? cost.Q[lane * n + lane] : cost.R[(lane - n) * m + (lane - n)];
                        __syncwarp();
                    }
                }

                // ---- gains: one lane per column of [Qux | Qu] -------------------------------------------
                double Lc[M_ * (M_ + 1) / 2];   // Cholesky factor of Quu + rho I, packed lower by rows
                double Quu[M_ * (M_ + 1) / 2];
                {
#pragma unroll
                    for (int a = 0; a < m; a++)
#pragma unroll
                        for (int c = 0; c <= a; c++) Quu[a * (a + 1) / 2 + c] = sm.Q[(n + c) * LDT + (n + a)];   // upper entry (c <= a)
#pragma unroll
                    for (int a = 0; a < m; a++) {
#pragma unroll
                        for (int c = 0; c <= a; c++) {
                            double t = Quu[a * (a + 1) / 2 + c] + ((a == c) ? rho : 0.0);
#pragma unroll
                            for (int r = 0; r < c; r++) t = fma(-Lc[a * (a + 1) / 2 + r], Lc[c * (c + 1) / 2 + r], t);
                            if (a == c) {
                                if (!(t > 0.0) || !isfinite(t)) ok = false;
                                Lc[a * (a + 1) / 2 + a] = sqrt(t);
                            } else Lc[a * (a + 1) / 2 + c] = t / Lc[c * (c + 1) / 2 + c];
                        }
                    }
                }
                if (!ok) break;   // uniform across the warp (every lane factors the same matrix)
                double kc[M_];    // column of K (lane < n) or d (lane == n)
                double rhs[M_];
                {
                    const int c = (lane <= n) ? lane : n;
#pragma unroll
                    for (int a = 0; a < m; a++) rhs[a] = (c < n) ? sm.Q[c * LDT + (n + a)] : sm.Q[(n + a) * LDT + NM];   // Qux[a][c] | Qu[a]
                    double y[M_];
#pragma unroll
                    for (int a = 0; a < m; a++) {
                        double t = -rhs[a];
#pragma unroll
                        for (int r = 0; r < a; r++) t = fma(-Lc[a * (a + 1) / 2 + r], y[r], t);
                        y[a] = t / Lc[a * (a + 1) / 2 + a];
                    }
#pragma unroll
                    for (int a = m - 1; a >= 0; a--) {
                        double t = y[a];
#pragma unroll
                        for (int r = a + 1; r < m; r++) t = fma(-Lc[r * (r + 1) / 2 + a], kc[r], t);
                        kc[a] = t / Lc[a * (a + 1) / 2 + a];
                    }
                    if (lane <= n) {
#pragma unroll
                        for (int a = 0; a < m; a++) {
                            sm.K[a * LDK + c] = kc[a];
                            sm.W[a * LDK + c] = fma(-rho, kc[a], rhs[a]);   // W = Qux - rho K
                        }
                    }
                    if (lane < n) {
#pragma unroll
                        for (int a = 0; a < m; a++) Kg[(size_t)k * n * m + lane * m + a] = kc[a];
                    } else if (lane == n) {
                        double t1 = 0.0, t2 = 0.0;
#pragma unroll
                        for (int a = 0; a < m; a++) {
                            dg[(size_t)k * m + a] = kc[a];
                            t1 = fma(kc[a], rhs[a], t1);
                            double qd = 0.0;   // (Quu d)_a
#pragma unroll
                            for (int r = 0; r < m; r++) qd = fma((r <= a) ? Quu[a * (a + 1) / 2 + r] : Quu[r * (r + 1) / 2 + a], kc[r], qd);
                            t2 = fma(0.5 * kc[a], qd, t2);
                        }
                        dV1 += t1; dV2 += t2;
                    }
                }
                __syncwarp();

                // ---- S <- Qxx + W'K (upper blocks, mirrored) ; s <- Qx + W'd ------------------------------
                {
                    double acc[RS][4];
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        const double2 q0 = lds128(&sm.Q[(2 * s_rb[r]) * LDT + 2 * s_cb[r]]);
                        const double2 q1 = lds128(&sm.Q[(2 * s_rb[r] + 1) * LDT + 2 * s_cb[r]]);
                        acc[r][0] = q0.x; acc[r][1] = q0.y; acc[r][2] = q1.x; acc[r][3] = q1.y;
                    }
#pragma unroll
                    for (int a = 0; a < m; a++) {
#pragma unroll
                        for (int r = 0; r < RS; r++) {
                            const double2 w = lds128(&sm.W[a * LDK + 2 * s_rb[r]]);
                            const double2 kk = lds128(&sm.K[a * LDK + 2 * s_cb[r]]);
                            fma2x2(acc[r], w, kk);
                        }
                    }
                    double snew = 0.0;
                    if (lane < n) {
                        snew = sm.Q[lane * LDT + NM];
#pragma unroll
                        for (int a = 0; a < m; a++) snew = fma(sm.W[a * LDK + lane], sm.K[a * LDK + n], snew);
                    }
                    s_reg = snew;
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        if (!s_on[r]) continue;
                        const int i0 = 2 * s_rb[r], j0 = 2 * s_cb[r];
                        if (s_rb[r] == s_cb[r]) {
                            const double off = 0.5 * (acc[r][1] + acc[r][2]);
                            sts128(&sm.S[i0 * NP + j0], acc[r][0], off);
                            sts128(&sm.S[(i0 + 1) * NP + j0], off, acc[r][3]);
                        } else {
                            sts128(&sm.S[i0 * NP + j0], acc[r][0], acc[r][1]);
                            sts128(&sm.S[(i0 + 1) * NP + j0], acc[r][2], acc[r][3]);
                            sts128(&sm.S[j0 * NP + i0], acc[r][0], acc[r][2]);
                            sts128(&sm.S[(j0 + 1) * NP + i0], acc[r][1], acc[r][3]);
                        }
                    }
                }
                __syncwarp();
            }   // knots

            if (ok) {
                if (lane == n) { P.dV[2 * b] = dV1; P.dV[2 * b + 1] = dV2; }
                break;
            }
            // ---- non-PD Quu at knot k: drain the copies still in flight (knots k-1 .. k-STAGES, already
            //      re-armed), increase rho (Altro regularization_update!(:increase)) and restart ------------
            {
                const int outstanding = (k < STAGES) ? k : STAGES;
                for (int i = 0; i < outstanding; i++) {
                    const int st = (stage + i) % STAGES;
                    mbar_wait(&sm.bar[st], (phase_bits >> st) & 1u);
                    phase_bits ^= (1u << st);
                }
            }
            __syncwarp();
            reg_increase(P.opt, rho, drho);
            restarts++;
            if (rho > P.opt.bp_reg_max) { failed = true; break; }
        }
        if (!failed) reg_decrease(P.opt, rho, drho);
        if (lane == 0) {
            P.rho[b] = rho; P.drho[b] = drho;
            P.bp_status[b] = failed ? -1 : restarts;
        }
        __syncwarp();
    }
}

template <int N_, int M_>
cudaError_t launch_riccati_t(const DevProblem& P, int* work_counter, cudaStream_t s) {
    constexpr int STAGES = 3;
    using SM = RiccatiSmem<N_, M_, STAGES>;
    auto kern = k_riccati<N_, M_, STAGES>;
    static bool configured = false;
    static int ctas_per_sm = 1, num_sms = 1;
    const int smem = (int)sizeof(SM);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 32, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured = true;
    }
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    int grid = num_sms * ctas_per_sm;     // persistent: a multiple of the SM count, instances pulled from a queue
    if (grid > P.B) grid = P.B;
    kern<<<grid, 32, smem, s>>>(P, work_counter);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_backward(const DevProblem& P, int* work_counter, cudaStream_t s) {
    if (P.n == 13 && P.m == 4) return launch_riccati_t<13, 4>(P, work_counter, s);
    if (P.n == 4 && P.m == 1) return launch_riccati_t<4, 1>(P, work_counter, s);
    if (P.n == 4 && P.m == 2) return launch_riccati_t<4, 2>(P, work_counter, s);
    if (P.n == 2 && P.m == 1) return launch_riccati_t<2, 1>(P, work_counter, s);
    return cudaErrorNotSupported;
}
