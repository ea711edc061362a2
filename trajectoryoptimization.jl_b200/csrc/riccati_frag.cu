// riccati_frag.cu -- kernel 3 for the error-state Quadrotor (n_e = 12, m = 4): the Riccati backward pass with the whole
// recursion state resident in FP64 tensor-core fragment registers.
//
// What it computes: Altro.jl backwardpass! on the error state (restated in oracle/oracle.hpp backward_pass),
// per instance, serial in k = N-1 .. 1, with z = [x_e; u] (16 entries):
//     Qzz = lzz + [A_e B_e]' S [A_e B_e]     Qz = lz + [A_e B_e]' s
//     K = -(Quu + rho I)^-1 Qux   d = -(Quu + rho I)^-1 Qu          (non-PD Quu + rho I -> rho increase + restart)
//     S <- Qxx + W'K   s <- Qx + W'd   with W = Qux - rho K          dV += (d'Qu, 1/2 d'Quu d)
// lzz / lz is the reference's cost expansion (RD.gradient!/hessian!, src/cost_functions.jl:137-233) plus the AL terms of Goal / Bound
// constraints (src/constraints.jl:55-68, :738-765; projection! src/cones.jl:96-145), projected on the error state
// (error_expansion!, intent at src/abstract_constraint.jl:282-303) by k_expansion_rec below.
//
// B200 mapping (why this kernel exists: profiles/r01_notes.md -- the shared-memory kernels spend their time on five smem hand-offs
// per knot and a 47-deep scalar LDL' chain at 4 warps per scheduler):
//   * one warp per instance; z is held in the PHYSICAL order of frag_layout.cuh, chosen so that the D fragments of every
//     mma.sync.m8n8k4.f64 (SASS DMMA) are exactly the A / B fragments of the next one:
//        T' = [A B]' S^      12 DMMA   B operand = the S accumulators of the previous knot (s rides in row 0 of S^)
//        Q^ = H^ + [A B]' T  12 DMMA   B operand = the T' accumulators; column 0 of T' is [A B]'s = Qz - lz for free
//        [K|W]' = Q^[:,u] [-Minv | I + rho Minv]   2 DMMA   A operand = register 0 of the Q^ tiles (u_a sits on p = 2a);
//                                                  row 0 <- Qu gives d and w_d = Qu - rho d in the same product
//        S^ <- Q^ + W'K      4 DMMA    A / B operands = the two result registers of the previous product (row 0 <- Qz gives s)
//     30 DMMA per knot, S / T / Q never leave the register file; per knot the warp touches shared memory for the record
//     (3 LDS.128 + 6 LDS.64), the 4 x 4 Quu (one STS, one LDS burst) and the 10 entries of its inverse.
//   * (Quu + rho I)^-1 by 2 x 2 block elimination (two Newton reciprocals, 20-deep chain instead of the 47 of a scalar LDL'),
//     evaluated by the lower half-warp only (an FP64 instruction of a half-empty warp takes one pipe pass);
//     positive-definiteness = positive leading minors a, det P, s00, det S (the pivots of LDL' are their ratios).
//   * the record of knot k (1920 B: fragments of [A_e B_e] + compact expansion) arrives by ONE 1-D bulk TMA copy (cp.async.bulk +
//     mbarrier, SASS UBLKCP) into a per-warp ring, issued one knot ahead.
//   * 72 registers, 4-warp CTAs x 7 per SM = 28 resident warps per SM: B = 4096 instances are a single wave on 148 SMs.
#include "costcon.cuh"
#include "frag_layout.cuh"
#include "kernels.h"

#ifndef TO_FRAG_STAGES
#define TO_FRAG_STAGES 2
#endif
#ifndef TO_FRAG_WARPS
#define TO_FRAG_WARPS 4
#endif
#ifndef TO_FRAG_MINB
#define TO_FRAG_MINB 7
#endif

namespace {

inline unsigned nblk(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ double2 lds128(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 1/x for a positive finite x: hardware seed + two Newton steps (<= 1 ulp), no IEEE-division slow path on the pivot chain
__device__ __forceinline__ double rcp_pos(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// ---- compact error-state expansion into the record --------------------------------------------------------------------------------
// One thread per (instance, knot).  Full-state expansion of a DiagonalCost + Goal / Bound AL rows is a gradient g and a DIAGONAL h;
// on the error state it is G'g, the same diagonal outside the attitude, and the 3 x 3 block G_q' diag(h_q) G_q - (q'g_q) I3
// (Altro error_expansion!; lie.cu k_expansion_compact computes the same numbers in logical order for the shared-memory kernel).
__global__ void __launch_bounds__(128) k_expansion_rec(const DevProblem P) {
    const int n = P.n, m = P.m, nm = n + m, qs = P.qs;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * P.N) return;
    const int k = (int)(t % P.N), b = (int)(t / P.N);
    const bool last = (k == P.N - 1);
    const double* xg = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* ug = traj_U(P, P.cur[b], b) + (size_t)k * m;
    const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
    double z[TO_MAXNM], g[TO_MAXNM], h[TO_MAXNM];
    for (int i = 0; i < n; i++) z[i] = xg[i];
    for (int a = 0; a < m; a++) z[n + a] = last ? 0.0 : ug[a];
    const DevCost& c = P.costs[P.cost_index[k]];
    for (int i = 0; i < n; i++) { g[i] = fma(c.Qd[i], z[i], c.q[i]); h[i] = c.Qd[i]; }
    for (int a = 0; a < m; a++) { g[n + a] = last ? 0.0 : fma(c.Rd[a], z[n + a], c.r[a]); h[n + a] = last ? 0.0 : c.Rd[a]; }
    const int lim = last ? n : nm;
    for (int ci = 0; ci < P.ncon; ci++) {
        const DevCon& con = P.cons[ci];
        if (k + 1 < con.first || k + 1 > con.last) continue;
        const double mu = P.mu[ci];
        const double* lam = lam_b + con.offset + (size_t)(k + 1 - con.first) * con.p;
        const bool eq = (con.kind == CON_GOAL);
        const int nrow = eq ? con.p : con.n_max + con.n_min;
        for (int r = 0; r < nrow; r++) {
            const int j = eq ? con.inds[r] : (r < con.n_max ? con.a_max[r] : con.a_min[r - con.n_max]);
            const bool lower = !eq && r >= con.n_max;
            const double cv = eq ? z[j] - con.a[r] : (lower ? con.b[j] - z[j] : z[j] - con.a[j]);
            const double lb = lam[r] - mu * cv;
            if ((eq || lb <= 0.0) && j < lim) { g[j] -= lower ? -lb : lb; h[j] += mu; }
        }
    }
    double* out = P.REC + t * TO_REC_LEN;
    double G[12]; quat_G(z + qs, G);
    double ge[16], hd[16];
    for (int e = 0; e < qs; e++) { ge[e] = g[e]; hd[e] = h[e]; }
    for (int e = qs + 3; e < n - 1 + m; e++) { ge[e] = g[e + 1]; hd[e] = h[e + 1]; }
    double qb = 0;
    for (int r = 0; r < 4; r++) qb += z[qs + r] * g[qs + r];
    for (int cc = 0; cc < 3; cc++) {
        double s = 0, d = 0;
        for (int r = 0; r < 4; r++) { s += G[cc * 4 + r] * g[qs + r]; d += G[cc * 4 + r] * h[qs + r] * G[cc * 4 + r]; }
        ge[qs + cc] = s; hd[qs + cc] = d - qb;
    }
    double b01 = 0, b02 = 0, b12 = 0;
    for (int r = 0; r < 4; r++) { b01 += G[r] * h[qs + r] * G[4 + r]; b02 += G[r] * h[qs + r] * G[8 + r]; b12 += G[4 + r] * h[qs + r] * G[8 + r]; }
#pragma unroll
    for (int j = 0; j < 16; j++) { out[TO_REC_G + fraglayout::phys_z(j)] = ge[j]; out[TO_REC_HD + fraglayout::phys_z(j)] = hd[j]; }
    // Hb[a][b] = H~[8+2a][8+2b]: a, b = 0..2 the attitude error (e = 3..5), a = 3 is p = 14 (e = 7)
    out[TO_REC_HB + 0] = hd[3]; out[TO_REC_HB + 1] = b01;   out[TO_REC_HB + 2] = b02;    out[TO_REC_HB + 3] = 0.0;
    out[TO_REC_HB + 4] = b01;   out[TO_REC_HB + 5] = hd[4]; out[TO_REC_HB + 6] = b12;    out[TO_REC_HB + 7] = 0.0;
    out[TO_REC_HB + 8] = b02;   out[TO_REC_HB + 9] = b12;   out[TO_REC_HB + 10] = hd[5]; out[TO_REC_HB + 11] = 0.0;
    out[TO_REC_HB + 12] = 0.0;  out[TO_REC_HB + 13] = 0.0;  out[TO_REC_HB + 14] = 0.0;   out[TO_REC_HB + 15] = hd[7];
}

// [A_e B_e] of the record back in the col-major 12 x 16 layout of P.ABe (to_get_error_dynamics, the shared-memory kernels of lie.cu)
__global__ void __launch_bounds__(128) k_export_abe(const DevProblem P) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * (P.N - 1) * 16) return;
    const int j = (int)(t & 15);
    const long long bk = t >> 4;
    const int k = (int)(bk % (P.N - 1)), b = (int)(bk / (P.N - 1));
    const double* rec = P.REC + ((size_t)b * P.N + k) * TO_REC_LEN;
    double* out = P.ABe + ((size_t)bk * 16 + j) * 12;
    for (int e = 0; e < 12; e++) {
        const int q = fraglayout::phys_x(e);
        const int ks = (q < 8) ? 0 : ((q & 1) ? 2 : 1);
        const int fc = (q < 8) ? (q - 1) / 2 : ((q & 1) ? (q - 9) / 2 : (q - 8) / 2);
        const int c = fraglayout::phys_z(j);
        out[e] = rec[(ks * 32 + 4 * (c & 7) + fc) * 2 + (c >> 3)];
    }
}

// ---- the backward pass ----------------------------------------------------------------------------------------------------------------
template <int STAGES, int WARPS>
struct FragSmem {
    double rec[WARPS][STAGES][TO_REC_LEN];    // per-warp ring of knot records (1920 B = 15 x 128 B each)
    double quu[WARPS][32];                    // [0,16) Quu (row-major 4 x 4) ; [16,27) the 10 entries of its inverse + the PD flag
    uint64_t bar[WARPS][STAGES];
};

// entries of the symmetric inverse in the order they are stored in shared memory
enum { MI_P00 = 0, MI_P01, MI_P11, MI_N00, MI_N01, MI_N10, MI_N11, MI_V00, MI_V01, MI_V11, MI_OK };
__device__ __forceinline__ int minv_slot(int i, int j) {   // slot of Minv[i][j]
    if (i > j) { const int t = i; i = j; j = t; }
    if (j < 2) return i == 0 ? (j == 0 ? MI_P00 : MI_P01) : MI_P11;
    if (i >= 2) return i == 2 ? (j == 2 ? MI_V00 : MI_V01) : MI_V11;
    // i < 2 <= j: Minv[j][i] = n_{j-2, i}
    return (j == 2) ? (i == 0 ? MI_N00 : MI_N01) : (i == 0 ? MI_N10 : MI_N11);
}

template <int STAGES, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) k_riccati_frag(const DevProblem P, int* __restrict__ work_counter) {
    using SM = FragSmem<STAGES, WARPS>;
    extern __shared__ __align__(128) unsigned char frag_smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(frag_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = lane >> 2, fc = lane & 3;
    const int N = P.N;
    double* const ring = &sm.rec[warp][0][0];
    double* const quu = sm.quu[warp];
    double* const minv = quu + 16;
    uint64_t* const bar = sm.bar[warp];

    // loop-invariant lane roles
    const bool d00 = (fr == 2 * fc), d01 = (fr == 2 * fc + 1);   // this lane holds a diagonal entry of a diagonal tile in reg 0 / reg 1
    const bool fr_even = (fr & 1) == 0;
    const bool row0 = (fr == 0);
    const int eoff0 = (fr & 1) ? fraglayout::e_of_p(fr) * 4 + fc : -1;          // K[a = fc][e(p = fr)]      (tile row 0: x_e only on odd p)
    const int eoff1 = fraglayout::e_of_p(8 + fr) * 4 + fc;                      // K[a = fc][e(p = 8 + fr)]
    const int mslot = minv_slot(fc, fr >> 1);                                     // this lane's entry Minv[fc][fr >> 1] of the B fragment
    const double bdelta = ((fr & 1) && fc == (fr >> 1)) ? 1.0 : 0.0;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase_bits = 0;

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(work_counter, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= P.B) break;
        const double* recg = P.REC + (size_t)b * N * TO_REC_LEN;
        double* Kg = P.K + (size_t)b * (N - 1) * 48;
        double* dg = P.d + (size_t)b * (N - 1) * 4;
        double rho = P.rho[b], drho = P.drho[b];
        int restarts = 0;
        bool failed = false;

        auto issue = [&](int st, int k) {
            if (lane == 0) {
                mbar_expect_tx(&bar[st], TO_REC_LEN * 8);
                bulk_g2s(ring + st * TO_REC_LEN, recg + (size_t)k * TO_REC_LEN, TO_REC_LEN * 8, &bar[st]);
            }
        };

        for (;;) {   // regularisation restart loop
#pragma unroll
            for (int s = 0; s < STAGES; s++) { const int k = N - 2 - s; if (k >= 0) issue(s, k); }
            // ---- terminal knot: S^ = H^_N with s = g_N in row 0 ---------------------------------------------------------------------
            double S[2][2][2];
            {
                const double* rN = recg + (size_t)(N - 1) * TO_REC_LEN;
                const double hd0 = rN[TO_REC_HD + fr], hd1 = rN[TO_REC_HD + 8 + fr], hb = rN[TO_REC_HB + (fr >> 1) * 4 + fc];
                S[0][0][0] = d00 ? hd0 : 0.0; S[0][0][1] = d01 ? hd0 : 0.0;
                S[0][1][0] = 0.0; S[0][1][1] = 0.0; S[1][0][0] = 0.0; S[1][0][1] = 0.0;
                S[1][1][0] = fr_even ? hb : 0.0; S[1][1][1] = d01 ? hd1 : 0.0;
                if (row0) {
                    S[0][0][0] = rN[TO_REC_G + 2 * fc]; S[0][0][1] = rN[TO_REC_G + 2 * fc + 1];
                    S[0][1][0] = rN[TO_REC_G + 8 + 2 * fc]; S[0][1][1] = rN[TO_REC_G + 9 + 2 * fc];
                }
            }
            double acc1 = 0.0, acc2 = 0.0;   // lanes (0, fc): sum_k d_a Qu_a, sum_k d_a^2
            bool ok = true;
            int stage = 0;
            int k;
            for (k = N - 2; k >= 0; k--) {
                mbar_wait(&bar[stage], (phase_bits >> stage) & 1u);
                phase_bits ^= (1u << stage);
                const double* r = ring + stage * TO_REC_LEN;
                // fragments of [A_e B_e]_k: abf[ks][mi] = AB[q(ks,fc)][p = 8 mi + fr]
                double abf[3][2];
#pragma unroll
                for (int ks = 0; ks < 3; ks++) { const double2 v = lds128(r + (ks * 32 + lane) * 2); abf[ks][0] = v.x; abf[ks][1] = v.y; }
                // ---- T'[c][j] = sum_q AB[q][c] S^[j][q] ------------------------------------------------------------------------------
                double T[2][2][2];
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int nj = 0; nj < 2; nj++) { T[mi][nj][0] = 0.0; T[mi][nj][1] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < 3; ks++) {
                    const int ni = (ks == 0) ? 0 : 1, rg = (ks == 1) ? 0 : 1;   // k-step class -> (column tile, register)
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int nj = 0; nj < 2; nj++) dmma(T[mi][nj][0], T[mi][nj][1], abf[ks][mi], S[nj][ni][rg]);
                }
                // ---- Q^ = H^ + [A B]' T ; Qz (column form, lanes fc == 0) = g~ + column 0 of T' ----------------------------------------
                double Q[2][2][2];
                {
                    const double hd0 = r[TO_REC_HD + fr], hd1 = r[TO_REC_HD + 8 + fr], hb = r[TO_REC_HB + (fr >> 1) * 4 + fc];
                    Q[0][0][0] = d00 ? hd0 : 0.0; Q[0][0][1] = d01 ? hd0 : 0.0;
                    Q[0][1][0] = 0.0; Q[0][1][1] = 0.0; Q[1][0][0] = 0.0; Q[1][0][1] = 0.0;
                    Q[1][1][0] = fr_even ? hb : 0.0; Q[1][1][1] = d01 ? hd1 : 0.0;
                }
                const double qzc0 = r[TO_REC_G + fr] + T[0][0][0], qzc1 = r[TO_REC_G + 8 + fr] + T[1][0][0];
#pragma unroll
                for (int ks = 0; ks < 3; ks++) {
                    const int ni = (ks == 0) ? 0 : 1, rg = (ks == 1) ? 0 : 1;
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int nc = 0; nc < 2; nc++) dmma(Q[mi][nc][0], Q[mi][nc][1], abf[ks][mi], T[nc][ni][rg]);
                }
                // Qz in row form for lanes (0, fc): entries 2fc, 2fc+1, 8+2fc, 9+2fc live in lanes (2fc, 0) / (2fc+1, 0)
                const double qx00 = __shfl_sync(0xffffffffu, qzc0, 8 * fc), qx01 = __shfl_sync(0xffffffffu, qzc0, 8 * fc + 4);
                const double qx10 = __shfl_sync(0xffffffffu, qzc1, 8 * fc), qx11 = __shfl_sync(0xffffffffu, qzc1, 8 * fc + 4);
                // ---- gains -------------------------------------------------------------------------------------------------------------
                if (fr_even) quu[(fr >> 1) * 4 + fc] = Q[0][0][0];      // Quu[a][b] = Q^[2a][2b] sits in lane (2a, b)
                __syncwarp();
                if (lane < 16) {
                    const double2 r0a = lds128(quu + 0), r0b = lds128(quu + 2), r1b = lds128(quu + 6), r2b = lds128(quu + 10);
                    const double m11 = quu[5], m33 = quu[15];
                    const double a = r0a.x + rho, bq = r0a.y, c = m11 + rho;
                    const double M20 = r0b.x, M30 = r0b.y, M21 = r1b.x, M31 = r1b.y;      // block (2:3, 0:1), upper-triangle copies
                    const double R00 = r2b.x + rho, R01 = r2b.y, R11 = m33 + rho;
                    const double detP = fma(a, c, -bq * bq);
                    const double iP = rcp_pos(detP);
                    const double yt00 = fma(M20, c, -M21 * bq), yt01 = fma(M21, a, -M20 * bq);     // Y~ = Q21 adj(P)
                    const double yt10 = fma(M30, c, -M31 * bq), yt11 = fma(M31, a, -M30 * bq);
                    const double z00 = fma(yt00, M20, yt01 * M21), z01 = fma(yt00, M30, yt01 * M31), z11 = fma(yt10, M30, yt11 * M31);
                    const double s00 = fma(-iP, z00, R00), s01 = fma(-iP, z01, R01), s11 = fma(-iP, z11, R11);   // Schur complement
                    const double detS = fma(s00, s11, -s01 * s01);
                    const double iS = rcp_pos(detS);
                    const double v00 = s11 * iS, v01 = -s01 * iS, v11 = s00 * iS;
                    const double y00 = yt00 * iP, y01 = yt01 * iP, y10 = yt10 * iP, y11 = yt11 * iP;
                    const double n00 = -fma(v00, y00, v01 * y10), n01 = -fma(v00, y01, v01 * y11);
                    const double n10 = -fma(v01, y00, v11 * y10), n11 = -fma(v01, y01, v11 * y11);
                    const double p00 = fma(c, iP, -fma(y00, n00, y10 * n10));
                    const double p01 = fma(-bq, iP, -fma(y00, n01, y10 * n11));
                    const double p11 = fma(a, iP, -fma(y01, n01, y11 * n11));
                    const bool pd = (a > 0.0) && (detP > 0.0) && (s00 > 0.0) && (detS > 0.0) && (detP < 1e300) && (detS < 1e300);
                    if (lane == 0) {
                        *reinterpret_cast<double2*>(minv + MI_P00) = make_double2(p00, p01);
                        *reinterpret_cast<double2*>(minv + MI_P11) = make_double2(p11, n00);
                        *reinterpret_cast<double2*>(minv + MI_N01) = make_double2(n01, n10);
                        *reinterpret_cast<double2*>(minv + MI_N11) = make_double2(n11, v00);
                        *reinterpret_cast<double2*>(minv + MI_V01) = make_double2(v01, v11);
                        minv[MI_OK] = pd ? 1.0 : 0.0;
                    }
                }
                __syncwarp();
                const double Eij = minv[mslot];
                ok = minv[MI_OK] != 0.0;
                if (!ok) break;                       // uniform
                const double bfrag = (fr & 1) ? fma(rho, Eij, bdelta) : -Eij;     // [-Minv | I + rho Minv], columns interleaved
                const double af0 = row0 ? qx00 : Q[0][0][0];                      // row 0 <- Qu (Qu[fc] = Qz[2 fc])
                double k0 = 0.0, w0 = 0.0, k1 = 0.0, w1 = 0.0;
                dmma(k0, w0, af0, bfrag);             // lane (fr, fc): K[fc][p = fr], W[fc][p = fr]; row 0: d[fc], w_d[fc]
                dmma(k1, w1, Q[1][0][0], bfrag);      //                K[fc][p = 8 + fr], W[fc][p = 8 + fr]
                acc1 = fma(k0, af0, acc1); acc2 = fma(k0, k0, acc2);
                if (eoff0 >= 0) Kg[(size_t)k * 48 + eoff0] = k0;
                Kg[(size_t)k * 48 + eoff1] = k1;
                if (row0) {
                    dg[(size_t)k * 4 + fc] = k0;
                    Q[0][0][0] = qx00; Q[0][0][1] = qx01; Q[0][1][0] = qx10; Q[0][1][1] = qx11;     // row 0 <- Qz: the product below leaves s there
                }
                // ---- S^ <- Q^ + W'K -----------------------------------------------------------------------------------------------------
                dmma(Q[0][0][0], Q[0][0][1], w0, k0);
                dmma(Q[0][1][0], Q[0][1][1], w0, k1);
                dmma(Q[1][0][0], Q[1][0][1], w1, k0);
                dmma(Q[1][1][0], Q[1][1][1], w1, k1);
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int nj = 0; nj < 2; nj++) { S[mi][nj][0] = Q[mi][nj][0]; S[mi][nj][1] = Q[mi][nj][1]; }
                // the record and the Minv slots have been consumed by every lane: refill the ring slot
                __syncwarp();
                if (k - STAGES >= 0) issue(stage, k - STAGES);
                stage = (stage + 1 == STAGES) ? 0 : stage + 1;
            }   // knots

            if (ok) {
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1); acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2);
                // 1/2 d'Quu d = -1/2 (d'Qu + rho d'd)   since (Quu + rho I) d = -Qu
                if (lane == 0) { P.dV[2 * b] = acc1; P.dV[2 * b + 1] = -0.5 * fma(rho, acc2, acc1); }
                break;
            }
            // ---- non-PD Quu + rho I at knot k: drain the copies in flight, increase rho (Altro regularization_update!(:increase)), restart ----
            {
                // slot `stage` (knot k) was consumed; the slots after it hold knots k-1 .. k-STAGES+1
                const int outstanding = (k < STAGES - 1) ? k : STAGES - 1;
                for (int i = 1; i <= outstanding; i++) {
                    const int st = (stage + i) % STAGES;
                    mbar_wait(&bar[st], (phase_bits >> st) & 1u);
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

}  // namespace

cudaError_t launch_expansion_rec(const DevProblem& P, cudaStream_t s) {
    k_expansion_rec<<<nblk((long long)P.B * P.N, 128), 128, 0, s>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_export_abe(const DevProblem& P, cudaStream_t s) {
    k_export_abe<<<nblk((long long)P.B * (P.N - 1) * 16, 128), 128, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_backward_frag(const DevProblem& P, int* work_counter, cudaStream_t s) {
    constexpr int STAGES = TO_FRAG_STAGES, WARPS = TO_FRAG_WARPS, MINB = TO_FRAG_MINB;
    using SM = FragSmem<STAGES, WARPS>;
    auto kern = k_riccati_frag<STAGES, WARPS, MINB>;
    const int smem = (int)sizeof(SM);
    // per-device launch configuration (one process may hold handles on several GPUs)
    static int ctas_per_sm[TO_MAXDEV] = {0}, num_sms[TO_MAXDEV] = {0};
    const int dev = current_device_slot();
    cudaError_t e = cudaSuccess;
    if (!ctas_per_sm[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        int c = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, 32 * WARPS, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm[dev] = c < 1 ? 1 : c;
    }
    e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    int grid = num_sms[dev] * ctas_per_sm[dev];       // persistent: warps pull instances from the queue
    const int need = (P.B + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<grid, 32 * WARPS, smem, s>>>(P, work_counter);
    return cudaGetLastError();
}
