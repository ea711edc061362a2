// rollout.cu -- kernel 1 of the hot path: batched RK4 rollout and dual-number dynamics expansion.
//
//   k_rollout : rollout!(prob)                         reference src/problem.jl:330-340
//               x_1 = x0 ; x_k = RK4(x_{k-1}, u_{k-1}, dt_{k-1}).  Serial in k, parallel over instances:
//               one thread per instance (the recursion has no intra-instance parallelism worth a warp).
//   k_expand  : RD.jacobian!(ForwardAD) on the discretised dynamics at every knot (no call site inside the
//               reference; shape [A B] = n x (n+m) pinned by test/dynamics_constraints.jl:35,57-62).
//               One thread per (instance, knot, seed direction j): the RK4 step is pushed through a
//               Dual<1> whose tangent is the one-hot e_j, i.e. the thread computes column j of [A B] with the
//               partial carried in registers.  Threads of one knot are adjacent, so row i of AB is written by
//               adjacent lanes; the pad columns of a row (LDAB > n+m) are never touched and stay zero.
#include <cstdlib>

#include "frag_layout.cuh"
#include "kernels.h"
#include "models.cuh"

// One thread integrates one instance; the warp writes its 32 states of a knot through a shared-memory transpose, so that the stores are runs of
// n contiguous doubles per instance (full 32-byte sectors) instead of 32 scattered 8-byte words per instruction.
template <int MODEL>
__global__ void __launch_bounds__(32) k_rollout(const DevProblem P) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m;
    __shared__ double stage[32 * n];
    __shared__ double* base[32];
    const int lane = threadIdx.x;
    const int b = blockIdx.x * 32 + lane;
    const bool valid = b < P.B;
    const int bc = valid ? b : P.B - 1;                     // (idle lanes of the last warp shadow a real instance and store nothing)
    base[lane] = valid ? traj_Xw(P, P.cur[bc], bc) : nullptr;
    const double* U = traj_U(P, P.cur[bc], bc);
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = P.x0[(size_t)bc * n + i];
    for (int k = 0; k < P.N; k++) {
#pragma unroll
        for (int i = 0; i < n; i++) stage[lane * n + i] = x[i];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < n; j++) {
            const int e = j * 32 + lane, ii = e / n, c = e - ii * n;
            double* Xi = base[ii];
            if (Xi) Xi[(size_t)k * n + c] = stage[e];
        }
        __syncwarp();
        if (k == P.N - 1) break;
#pragma unroll
        for (int i = 0; i < m; i++) u[i] = U[k * m + i];
        rk4_step<MODEL, double>(model_params<MODEL>(P, k), x, u, P.dt[k], xn);
#pragma unroll
        for (int i = 0; i < n; i++) x[i] = xn[i];
    }
}

// Seed pruning (full state).  The position r and the world-frame linear velocity v of the Quadrotor (a RobotDynamics RigidBody) enter the
// dynamics only through rdot = v, so their columns of [A B] are known in closed form -- d x+/d r = e_r, d x+/d v = h e_r + e_v (the RK4
// weights sum to one) -- and are written once when the problem is created (k_trivial_columns_full); 11 seeds (quaternion, angular velocity,
// controls) are pushed through the dual-number RK4 step instead of 17.  Other models: every seed.
template <int MODEL> struct SeedList {
    static constexpr int count = ModelDims<MODEL>::n + ModelDims<MODEL>::m;
    __host__ __device__ static constexpr int seed(int s) { return s; }
    __host__ __device__ static constexpr int ntrivial() { return 0; }
    __host__ __device__ static constexpr int trivial(int) { return 0; }
};
template <> struct SeedList<MODEL_QUADROTOR> {
    static constexpr int count = 11;
    __host__ __device__ static constexpr int seed(int s) { return s < 4 ? 3 + s : 6 + s; }       // 3..6 (q), 10..12 (omega), 13..16 (u)
    __host__ __device__ static constexpr int ntrivial() { return 6; }
    __host__ __device__ static constexpr int trivial(int s) { return s < 3 ? s : 4 + s; }        // 0..2 (r), 7..9 (v)
};

template <int MODEL, int NP>
__global__ void __launch_bounds__(128) k_expand(const DevProblem P, int mode) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, nm = n + m;
    constexpr int NSEED = SeedList<MODEL>::count;
    constexpr int TPK = (NSEED + NP - 1) / NP;     // threads per knot: each carries NP seed directions
    using D = Dual<NP>;
    const int ld = P.ldab;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.B * (P.N - 1) * TPK;
    if (t >= total) return;
    const int s0 = (int)(t % TPK) * NP;            // first seed slot of this thread
    const long long bk = t / TPK;
    const int k = (int)(bk % (P.N - 1));
    const int b = (int)(bk / (P.N - 1));
    if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) return;     // overlapped expansion: this launch covers the other group
    double* AB = P.AB + ((size_t)b * (P.N - 1) + k) * n * ld;     // pad columns nm..ld-1 stay zero from to_create
    const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
    int js[NP];                                    // the z index each seed slot differentiates with respect to (nm = none)
#pragma unroll
    for (int q = 0; q < NP; q++) js[q] = (s0 + q < NSEED) ? SeedList<MODEL>::seed(s0 + q) : nm;
    D x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) {
        x[i].v = X[i];
#pragma unroll
        for (int q = 0; q < NP; q++) x[i].d[q] = (i == js[q]) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < m; i++) {
        u[i].v = U[i];
#pragma unroll
        for (int q = 0; q < NP; q++) u[i].d[q] = (n + i == js[q]) ? 1.0 : 0.0;
    }
    rk4_step<MODEL, D>(model_params<MODEL>(P, k), x, u, P.dt[k], xn);
#pragma unroll
    for (int i = 0; i < n; i++) {
        if (NP == 2 && js[1] == js[0] + 1 && !(js[0] & 1)) *reinterpret_cast<double2*>(&AB[i * ld + js[0]]) = make_double2(xn[i].d[0], xn[i].d[1]);
        else {
#pragma unroll
            for (int q = 0; q < NP; q++) if (js[q] < nm) AB[i * ld + js[q]] = xn[i].d[q];
        }
    }
}

// the closed-form columns of [A B] (SeedList<MODEL>::trivial): thread = (instance, knot, one of them); run once per problem
template <int MODEL>
__global__ void __launch_bounds__(128) k_trivial_columns_full(const DevProblem P) {
    constexpr int n = ModelDims<MODEL>::n, NT = SeedList<MODEL>::ntrivial();
    if constexpr (NT > 0) {
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= (long long)P.B * (P.N - 1) * NT) return;
        const int jt = SeedList<MODEL>::trivial((int)(t % NT));
        const long long bk = t / NT;
        const int k = (int)(bk % (P.N - 1));
        double* AB = P.AB + (size_t)bk * n * P.ldab;
        const double h = P.dt[k];
        // x = [r(0..2); q(3..6); v(7..9); omega(10..12)]: d r+/d r = I, d r+/d v = h I, d v+/d v = I
        for (int i = 0; i < n; i++) AB[i * P.ldab + jt] = (i == jt) ? 1.0 : ((jt >= 7 && i == jt - 7) ? h : 0.0);
    }
}
cudaError_t launch_trivial_columns_full(const DevProblem& P, cudaStream_t s) {
    if (P.model == MODEL_QUADROTOR) {
        const long long total = (long long)P.B * (P.N - 1) * SeedList<MODEL_QUADROTOR>::ntrivial();
        k_trivial_columns_full<MODEL_QUADROTOR><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(P);
    }
    return cudaGetLastError();
}

// (g, h) = (lz_i, lzz_ii) of entry i of the full-state z = [x; u] at knot k (0-based): DiagonalCost (RD.gradient!/hessian!, src/cost_functions.jl:137-233)
// + the AL terms of the Goal / Bound rows acting on z_i (src/constraints.jl:55-68, :738-765; projection on the dual cone src/cones.jl:96-145).
// The compact problem class only (P.compact): every cost diagonal, every constraint Goal or Bound, at most TO_EXP_MAXT rows per entry
// (the host-built table P.exptab; walking the constraint descriptors per thread instead made this kernel 2x slower than its RK4 work).
__device__ __forceinline__ void compact_entry_expansion(const DevProblem& P, const ExpTab& tab, int k, int i, double zi, const double* __restrict__ lam_b, double& g, double& h) {
    const int n = P.n;
    const bool last = (k == P.N - 1);
    const DevCost& c = P.costs[P.cost_index[k]];
    if (i < n) { g = fma(c.Qd[i], zi, c.q[i]); h = c.Qd[i]; }
    else if (last) { g = 0.0; h = 0.0; return; }
    else { g = fma(c.Rd[i - n], zi, c.r[i - n]); h = c.Rd[i - n]; }
#pragma unroll
    for (int t = 0; t < TO_EXP_MAXT; t++) {
        const unsigned px = __ldg(&tab.pkx[t][i]);
        if ((unsigned)(k + 1) - (px & 0xfffu) <= ((px >> 12) & 0xfffu)) {
            const double nms = __ldg(&tab.nms[t][i]);
            const double lam = lam_b[(int)(__ldg(&tab.pky[t][i]) + (unsigned)(k + 1) * ((px >> 24) & 0x7fu))];
            const double lb = fma(nms, zi - __ldg(&tab.bound[t][i]), lam);          // lambda - mu c
            if ((px >> 31) || lb <= 0.0) { g += (nms < 0.0) ? -lb : lb; h += fabs(nms); }   // g -= sign lb ; h += mu
        }
    }
}

// Seed pruning.  The position r and the (world-frame) linear velocity v of a RigidBody enter the dynamics only through rdot = v: f does not
// depend on r, and on v only in rdot.  Their columns of the discrete Jacobian are therefore known in closed form -- d x+/d r = e_r and
// d x+/d v = h e_r + e_v (the RK4 weights sum to one) -- and need no dual-number sweep: 10 seeds (attitude, angular velocity, controls) are
// pushed through the RK4 step instead of 16, one thread each; the six trivial columns depend on the time steps only and are written once,
// when the problem is created (k_trivial_columns).
__device__ __forceinline__ int lie_seed(int s) { return (int)((0xFEDCBA9543ULL >> (4 * s)) & 15); }       // 3,4,5,9,10,11,12,13,14,15
__device__ __forceinline__ int lie_trivial(int s) { return (int)((0x876210ULL >> (4 * s)) & 15); }        // 0,1,2,6,7,8

// column j of [A_e B_e]_k: the RK4 step of knot k pushed through Dual<1> with the seed of error-state coordinate j (attitude: a column of G(q_k)),
// projected on the error state of knot k + 1 with G(q_{k+1})'
template <int MODEL>
__device__ __forceinline__ void expand_lie_column(const DevProblem& P, int k, int j, const double* __restrict__ X, const double* __restrict__ U, double (&col)[ModelDims<MODEL>::n - 1]) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, ne = n - 1, qs = 3;
    using D = Dual<1>;
    const double h = P.dt[k];
    D x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i].v = X[i]; x[i].d[0] = 0.0; }
#pragma unroll
    for (int i = 0; i < m; i++) { u[i].v = U[i]; u[i].d[0] = (ne + i == j) ? 1.0 : 0.0; }
    if (j < qs + 3) {
        const double w = X[qs], qx = X[qs + 1], qy = X[qs + 2], qz = X[qs + 3];
        const int c = j - qs;        // column c of L(q) H: (-x,w,z,-y), (-y,-z,w,x), (-z,y,-x,w)
        x[qs].d[0] = (c == 0) ? -qx : (c == 1) ? -qy : -qz;
        x[qs + 1].d[0] = (c == 0) ? w : (c == 1) ? -qz : qy;
        x[qs + 2].d[0] = (c == 0) ? qz : (c == 1) ? w : -qx;
        x[qs + 3].d[0] = (c == 0) ? -qy : (c == 1) ? qx : w;
    } else if (j < ne) {
#pragma unroll
        for (int i = qs + 4; i < n; i++) x[i].d[0] = (i == j + 1) ? 1.0 : 0.0;
    }
    rk4_step<MODEL, D>(model_params<MODEL>(P, k), x, u, h, xn);
    const double* q1 = X + n + qs;                                     // attitude of knot k + 1
    const double w1 = q1[0], x1 = q1[1], y1 = q1[2], z1 = q1[3];
#pragma unroll
    for (int i = 0; i < qs; i++) col[i] = xn[i].d[0];
    const double t0 = xn[qs].d[0], t1 = xn[qs + 1].d[0], t2 = xn[qs + 2].d[0], t3 = xn[qs + 3].d[0];
    col[qs] = -x1 * t0 + w1 * t1 + z1 * t2 - y1 * t3;
    col[qs + 1] = -y1 * t0 - z1 * t1 + w1 * t2 + x1 * t3;
    col[qs + 2] = -z1 * t0 + y1 * t1 - x1 * t2 + w1 * t3;
#pragma unroll
    for (int i = qs + 4; i < n; i++) col[i - 1] = xn[i].d[0];
}

// FRAG: the column goes into the fragment block of the knot's record (frag_layout.cuh) instead of P.ABe.
#ifndef TO_EXPAND_LIE_MINB
#define TO_EXPAND_LIE_MINB 4      // CTAs per SM the register allocation aims at: 128 registers, 16 warps per SM (r02g: 0.29 vs 0.37 ms at 168 registers / 12 warps)
#endif
// CTA size: the kernel fills the register file (128 registers x 512 threads), so every CTA of another kernel that becomes resident next to it (the
// late line-search trials on the side stream, 4096 registers each) evicts a whole CTA of this one: 64-thread CTAs lose 1/8 of an SM, not 1/4.
#ifndef TO_EXPAND_LIE_THREADS
#define TO_EXPAND_LIE_THREADS 64
#endif
template <int MODEL, bool FRAG>
__global__ void __launch_bounds__(TO_EXPAND_LIE_THREADS, TO_EXPAND_LIE_MINB * 128 / TO_EXPAND_LIE_THREADS) k_expand_lie(const DevProblem P, int mode) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, ne = n - 1, nme = ne + m, qs = 3, NS = 10;
    using D = Dual<1>;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.B * (P.N - 1) * NS;
    if (t >= total) return;
    const int sd = (int)(t % NS);
    const int j = lie_seed(sd);
    const long long bk = t / NS;
    const int k = (int)(bk % (P.N - 1));
    const int b = (int)(bk / (P.N - 1));
    if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) return;
    const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
    const double h = P.dt[k];
    // where column jj of [A_e B_e] goes: the fragment slots of the record, or 12 contiguous doubles of P.ABe
    auto store_column = [&](int jj, const double (&col)[ne]) {
        if constexpr (FRAG) {
            // element (row e, column jj) of the record: (ks(e)*32 + 4*(c & 7) + fc(e))*2 + (c >> 3), c = physical index of column jj
            const int c = (int)((0x6420FDB9E7CA8531ULL >> (4 * jj)) & 15);       // fraglayout::phys_z(jj) as a nibble table
            double* rec = P.REC + ((size_t)b * P.N + k) * TO_REC_LEN + 8 * (c & 7) + (c >> 3);
#pragma unroll
            for (int e = 0; e < ne; e++) rec[fraglayout::ab_index(e, 12)] = col[e];   // column 12 (u_0, c = 0) has a zero column offset
        } else {
            double* out = P.ABe + ((size_t)bk * nme + jj) * ne;
#pragma unroll
            for (int e = 0; e < ne; e++) out[e] = col[e];
        }
    };
    double col[ne];
    expand_lie_column<MODEL>(P, k, j, X, U, col);
    store_column(j, col);
}

// k_expand_lie with the knot's fragment block staged in shared memory.  k_expand_lie stores a column as 12 scattered 8-byte words: every one is
// a partial-sector write for the L2 (ncu r02z_expand: 368 MB of DRAM READS for 70 MB of inputs -- sector fills -- and 50 M sector transactions).
// Here the 10 seed threads of a knot sit in one CTA, drop their columns (and the six closed-form ones) into a 1536-byte image of the
// record's [A_e B_e] block, and the CTA writes the images out as whole lines.
template <int MODEL>
__global__ void __launch_bounds__(TO_EXPAND_LIE_THREADS, TO_EXPAND_LIE_MINB * 128 / TO_EXPAND_LIE_THREADS) k_expand_lie_staged(const DevProblem P, int mode) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, ne = n - 1, NS = 10, T = TO_EXPAND_LIE_THREADS, KPC = T / NS;
    static_assert(ne == 12 && m == 4, "record layout of the error-state Quadrotor");
    __shared__ __align__(16) double st[KPC][TO_REC_G];
    __shared__ int live[KPC];
    const int tid = threadIdx.x, kk = tid / NS, sd = tid - kk * NS;
    const long long nknots = (long long)P.B * (P.N - 1);
    const long long bk0 = (long long)blockIdx.x * KPC;
    bool work = kk < KPC && bk0 + kk < nknots;
    int b = 0, k = 0;
    if (work) {
        k = (int)((bk0 + kk) % (P.N - 1)); b = (int)((bk0 + kk) / (P.N - 1));
        if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) work = false;            // overlapped iterations: the other launch covers this instance
    }
    if (kk < KPC && sd == 0) live[kk] = work ? 1 : 0;
    for (int idx = tid; idx < KPC * 72; idx += T) {                                // the closed-form columns (positions, velocities; k_trivial_columns)
        const int q = idx / 72, r = idx - q * 72, s6 = r / 12, e = r - s6 * 12;
        if (bk0 + q < nknots) {
            const int kq = (int)((bk0 + q) % (P.N - 1));
            const int jt = lie_trivial(s6);
            st[q][fraglayout::ab_index(e, jt)] = (e == jt) ? 1.0 : ((jt >= 6 && e == jt - 6) ? P.dt[kq] : 0.0);
        }
    }
    if (work) {
        const int j = lie_seed(sd);
        const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
        const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
        double col[ne];
        expand_lie_column<MODEL>(P, k, j, X, U, col);
        const int c = (int)((0x6420FDB9E7CA8531ULL >> (4 * j)) & 15);               // fraglayout::phys_z(j)
        double* dst = &st[kk][8 * (c & 7) + (c >> 3)];
#pragma unroll
        for (int e = 0; e < ne; e++) dst[fraglayout::ab_index(e, 12)] = col[e];
    }
    __syncthreads();
    for (int idx = tid; idx < KPC * (TO_REC_G / 2); idx += T) {
        const int q = idx / (TO_REC_G / 2), w = idx - q * (TO_REC_G / 2);
        if (live[q]) {
            const int kq = (int)((bk0 + q) % (P.N - 1)), bq = (int)((bk0 + q) / (P.N - 1));
            double2* rec = reinterpret_cast<double2*>(P.REC + ((size_t)bq * P.N + kq) * TO_REC_LEN);
            rec[w] = reinterpret_cast<const double2*>(st[q])[w];
        }
    }
}

// the closed-form columns of [A_e B_e] (positions, velocities): thread = (instance, knot, one of the six)
template <bool FRAG>
__global__ void __launch_bounds__(128) k_trivial_columns(const DevProblem P) {
    constexpr int ne = 12, nme = 16;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P.B * (P.N - 1) * 6) return;
    const int sd = (int)(t % 6);
    const long long bk = t / 6;
    const int k = (int)(bk % (P.N - 1)), b = (int)(bk / (P.N - 1));
    const int jt = lie_trivial(sd);
    const double h = P.dt[k];
    for (int e = 0; e < ne; e++) {
        const double v = (e == jt) ? 1.0 : ((jt >= 6 && e == jt - 6) ? h : 0.0);
        if (FRAG) P.REC[((size_t)b * P.N + k) * TO_REC_LEN + fraglayout::ab_index(e, jt)] = v;
        else P.ABe[((size_t)bk * nme + jt) * ne + e] = v;
    }
}
cudaError_t launch_trivial_columns(const DevProblem& P, cudaStream_t s) {
    const long long total = (long long)P.B * (P.N - 1) * 6;
    if (P.frag) k_trivial_columns<true><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(P);
    else k_trivial_columns<false><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(P);
    return cudaGetLastError();
}

// Cost + AL expansion part of every record (frag_layout.cuh [192, 240)), the terminal knot included: 16 lanes per knot, lane i = entry i of
// the full-state z = [x; u] (lane 0 also takes the 17th entry u_3).  Outside the attitude the error-state expansion of a diagonal full-state
// one is the same entry; the quaternion block is projected by lanes 3..5 from the (g, h, q) of lanes 3..6, fetched by 16-lane shuffles:
// G'g, G' diag(h) G - (q'g_q) I3 (Altro error_expansion!; lie.cu k_expansion_compact is the one-thread-per-knot version of the same numbers).
// A light kernel (every load independent, ~40 registers) that runs at full occupancy; fused into the FP64-bound k_expand_lie it doubled that
// kernel's time (profiles/r02_notes.md).
#ifndef TO_CEXP_ITERS
#define TO_CEXP_ITERS 4          // knots per 16-lane group (the grid shrinks accordingly)
#endif
#ifndef TO_CEXP_MINB
#define TO_CEXP_MINB 4          // 64 registers: 32 warps per SM (r02m: 0.179 ms against 0.205 at 3)
#endif
#ifndef TO_CEXP_THREADS
#define TO_CEXP_THREADS 128
#endif
__global__ void __launch_bounds__(TO_CEXP_THREADS, TO_CEXP_MINB * 256 / TO_CEXP_THREADS) k_expansion_rec16(const DevProblem P, int mode) {
    constexpr int qs = 3;
    const int i = threadIdx.x & 15;                                                   // full-state entry of this lane
    const int n = P.n, N = P.N;
    const int ngroups = (int)((gridDim.x * blockDim.x) >> 4);
    const ExpTab& tab = *P.exptab;
    // the AL rows acting on z_i: loop-invariant, kept in registers (the first version re-read them per knot: 530 instructions per thread)
    unsigned px[TO_EXP_MAXT], py[TO_EXP_MAXT]; double nms[TO_EXP_MAXT], bnd[TO_EXP_MAXT];
#pragma unroll
    for (int t = 0; t < TO_EXP_MAXT; t++) { px[t] = tab.pkx[t][i]; py[t] = tab.pky[t][i]; nms[t] = tab.nms[t][i]; bnd[t] = tab.bound[t][i]; }
    const int e = (i < qs) ? i : i - 1;                                               // error-state coordinate of entry i (controls: 12 + a = i - 1)
    const int pme = (int)((0x6420FDB9E7CA8531ULL >> (4 * (e & 15))) & 15);            // its physical slot
    auto entry = [&](const DevCost& c, int k, int ii, double zi, const double* lam_b, const unsigned (&qx)[TO_EXP_MAXT], const unsigned (&qy)[TO_EXP_MAXT],
                     const double (&qn)[TO_EXP_MAXT], const double (&qb)[TO_EXP_MAXT], double& g, double& h) {
        const bool last = (k == N - 1);
        if (ii < n) { g = fma(c.Qd[ii], zi, c.q[ii]); h = c.Qd[ii]; }
        else if (last) { g = 0.0; h = 0.0; return; }
        else { g = fma(c.Rd[ii - n], zi, c.r[ii - n]); h = c.Rd[ii - n]; }
#pragma unroll
        for (int t = 0; t < TO_EXP_MAXT; t++) {
            if ((unsigned)(k + 1) - (qx[t] & 0xfffu) <= ((qx[t] >> 12) & 0xfffu)) {
                const double lam = lam_b[(int)(qy[t] + (unsigned)(k + 1) * ((qx[t] >> 24) & 0x7fu))];
                const double lb = fma(qn[t], zi - qb[t], lam);                         // lambda - mu c
                if ((qx[t] >> 31) || lb <= 0.0) { g += (qn[t] < 0.0) ? -lb : lb; h += fabs(qn[t]); }   // g -= sign lb ; h += mu
            }
        }
    };
    const int total = P.B * N;
    const unsigned gm = 0xFFFFu << (threadIdx.x & 16);                                // the two groups of a warp may take different trips
    for (int bk = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 4); bk < total; bk += ngroups) {      // (whole 16-lane groups iterate together)
        const int b = bk / N, k = bk - b * N;
        if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) continue;     // overlapped iterations: this launch covers the other group (see k_expand)
        const bool last = (k == N - 1);
        const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
        const double* U = traj_U(P, P.cur[b], b) + (size_t)(last ? 0 : k) * P.m;      // (not read at the terminal knot)
        const double* lam_b = P.lambda + (size_t)b * P.lambda_len;
        const DevCost& c = P.costs[P.cost_index[k]];
        double* rec = P.REC + (size_t)bk * TO_REC_LEN;
        const double zi = (i < n) ? X[i] : (last ? 0.0 : U[i - n]);
        double g, h;
        entry(c, k, i, zi, lam_b, px, py, nms, bnd, g, h);
        // the quaternion block: (g, h, q) of lanes 3..6 to every lane of the 16-lane group (lanes 3..5 use them)
        double gq[4], hq[4], q[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            gq[r] = __shfl_sync(gm, g, qs + r, 16); hq[r] = __shfl_sync(gm, h, qs + r, 16); q[r] = __shfl_sync(gm, zi, qs + r, 16);
        }
        if (i >= qs && i < qs + 3) {
            const int cc = i - qs;
            // rows of G' = (L(q) H)': (-x,w,z,-y), (-y,-z,w,x), (-z,y,-x,w)   (kept in registers: no run-time indexed arrays)
            const double G0[4] = {-q[1], q[0], q[3], -q[2]}, G1[4] = {-q[2], -q[3], q[0], q[1]}, G2[4] = {-q[3], q[2], -q[1], q[0]};
            double gc[4];
#pragma unroll
            for (int r = 0; r < 4; r++) gc[r] = (cc == 0) ? G0[r] : (cc == 1) ? G1[r] : G2[r];
            double qb = 0.0, ge = 0.0, hb0 = 0.0, hb1 = 0.0, hb2 = 0.0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                qb += q[r] * gq[r]; ge += gc[r] * gq[r];
                const double tt = gc[r] * hq[r];
                hb0 += tt * G0[r]; hb1 += tt * G1[r]; hb2 += tt * G2[r];
            }
            const double hd = ((cc == 0) ? hb0 : (cc == 1) ? hb1 : hb2) - qb;
            const int p = 8 + 2 * cc;                                                 // attitude error e = 3 + c sits on p = 8, 10, 12 (frag_layout.cuh)
            rec[TO_REC_G + p] = ge; rec[TO_REC_HD + p] = hd;
            rec[TO_REC_HB + 4 * cc + 0] = (cc == 0) ? hd : hb0;
            rec[TO_REC_HB + 4 * cc + 1] = (cc == 1) ? hd : hb1;
            rec[TO_REC_HB + 4 * cc + 2] = (cc == 2) ? hd : hb2;
            rec[TO_REC_HB + 4 * cc + 3] = 0.0;
        } else if (i != qs + 3) {                                                      // lane 6 (q_z) has no coordinate of its own
            rec[TO_REC_G + pme] = g; rec[TO_REC_HD + pme] = h;
            if (e == 7) { rec[TO_REC_HB + 12] = 0.0; rec[TO_REC_HB + 13] = 0.0; rec[TO_REC_HB + 14] = 0.0; rec[TO_REC_HB + 15] = h; }   // p = 14 is row 3 of Hb
        }
        if (i == 0) {                                                                  // the 17th entry: u_3 -> coordinate 15 (physical slot 6)
            unsigned rx[TO_EXP_MAXT], ry[TO_EXP_MAXT]; double rn[TO_EXP_MAXT], rb[TO_EXP_MAXT];
#pragma unroll
            for (int t = 0; t < TO_EXP_MAXT; t++) { rx[t] = __ldg(&tab.pkx[t][n + 3]); ry[t] = __ldg(&tab.pky[t][n + 3]); rn[t] = __ldg(&tab.nms[t][n + 3]); rb[t] = __ldg(&tab.bound[t][n + 3]); }
            entry(c, k, n + 3, last ? 0.0 : U[3], lam_b, rx, ry, rn, rb, g, h);
            rec[TO_REC_G + 6] = g; rec[TO_REC_HD + 6] = h;
        }
    }
}
// ---- k_expansion_rec16b: the same numbers, blocked by 16 knots (the default; TO_CEXP_V1=1 selects the kernel above) -----------------
// k_expansion_rec16 spends 270 instructions per lane and knot on 3 outputs (ncu r02z_costexp: 112 M warp instructions, issue slots 59 %
// busy, half of the stalls on two-level dependent loads cur[b] -> X, cost_index[k] -> DevCost): every 16-lane step pays for the
// attitude projection that 3 of its lanes need (24 shuffles + ~40 FP64), for the 17th entry u_3 that lane 0 takes in a divergent second
// call, and for the unpacking of the term table.  Here a 16-lane group owns a BLOCK of 16 consecutive knots of one instance:
//   phase A   16 steps, lane i < 13 = state entry x_i of one knot: diagonal cost + AL terms -> (g, h); loads (x_i and the multipliers of
//             its <= 3 terms) are issued for 4 knots at a time before the arithmetic; the cost coefficients of the lane are cached in
//             registers while the cost index does not change; lanes 3..6 (the quaternion) leave (g, h, q) in shared memory
//   phase B   lane j projects the attitude block of knot j (G'g, G' diag(h) G - (q'g_q) I3): once per knot instead of once per step
//   phase C   lane j takes the four control entries of knot j (their Bound rows are the AL terms that are active at every knot: in
//             phase A three lanes of sixteen would execute them at every step)
// ~55 instructions per lane and knot.  Same expressions in the same order as above: the records are bit-identical.
#ifndef TO_CEXP2_MINB
#define TO_CEXP2_MINB 9
#endif
#ifndef TO_CEXP2_THREADS
#define TO_CEXP2_THREADS 64
#endif
__global__ void __launch_bounds__(TO_CEXP2_THREADS, TO_CEXP2_MINB) k_expansion_rec16b(const DevProblem P, int mode) {
    constexpr int qs = 3, n = 13, m = 4;
    constexpr int ROW = 49;                                                          // 48 doubles of expansion per knot, padded: lane j works on row j (stride 98 words: conflict-free)
    // the block's 16 x 48 outputs are staged in shared memory and leave as whole 128-byte lines: written entry by entry they are 8-byte
    // stores scattered over the records, and every one of them costs the L2 a 32-byte sector transaction (r02v: 0.16 ms for 160 MB)
    __shared__ __align__(16) double stage_s[TO_CEXP2_THREADS / 16][16][ROW];
    const int i = threadIdx.x & 15;
    double (*stage)[ROW] = stage_s[threadIdx.x >> 4];
    const int N = P.N, NB = (N + 15) >> 4;
    const int ngroups = (int)((gridDim.x * blockDim.x) >> 4);
    const ExpTab& tab = *P.exptab;
    unsigned px[TO_EXP_MAXT], py[TO_EXP_MAXT]; double nms[TO_EXP_MAXT], bnd[TO_EXP_MAXT];
#pragma unroll
    for (int t = 0; t < TO_EXP_MAXT; t++) { px[t] = tab.pkx[t][i]; py[t] = tab.pky[t][i]; nms[t] = tab.nms[t][i]; bnd[t] = tab.bound[t][i]; }
    const int e = (i < qs) ? i : i - 1;                                               // error-state coordinate of entry i (controls: 12 + a = i - 1)
    const int pme = (int)((0x6420FDB9E7CA8531ULL >> (4 * (e & 15))) & 15);            // its physical slot
    const bool quat = (i >= qs && i <= qs + 3);
    const unsigned gm = 0xFFFFu << (threadIdx.x & 16);
    const int total = P.B * NB;
    constexpr int G_ = TO_REC_G - TO_REC_G, HD_ = TO_REC_HD - TO_REC_G, HB_ = TO_REC_HB - TO_REC_G;   // offsets inside the staged 48 doubles
    for (int unit = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 4); unit < total; unit += ngroups) {
        const int b = unit / NB, kb = (unit - b * NB) << 4;
        if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) continue;                   // (uniform over the group)
        const int buf = P.cur[b];
        const double* __restrict__ Xb = traj_X(P, buf, b);
        const double* __restrict__ Ub = traj_U(P, buf, b);
        const double* __restrict__ lam_b = P.lambda + (size_t)b * P.lambda_len;
        double* __restrict__ recb = P.REC + ((size_t)b * N + kb) * TO_REC_LEN + TO_REC_G;
        const int nk = (N - kb < 16) ? N - kb : 16;
        const int mycid = (i < nk) ? P.cost_index[kb + i] : 0;                        // lane j <-> knot kb + j (phases B, C; broadcast in phase A)
        int ccid = -1; double ca = 0.0, cb = 0.0;                                     // this lane's coefficients (Qd_i, q_i) of cost ccid
        // term t acts on the knots of the block whose bit is set in act[t]; lp[t] = its multiplier at the first knot of the block
        unsigned act[TO_EXP_MAXT]; const double* lp[TO_EXP_MAXT]; int ls[TO_EXP_MAXT];
#pragma unroll
        for (int t = 0; t < TO_EXP_MAXT; t++) {
            const int first = (int)(px[t] & 0xfffu), span = (int)((px[t] >> 12) & 0xfffu);
            int lo = first - 1 - kb, hi = first + span - kb;                          // knots kb + lo .. kb + hi - 1 (0-based) carry the row
            lo = lo < 0 ? 0 : lo; hi = hi > nk ? nk : hi;
            act[t] = (hi > lo && i < n) ? ((0xFFFFu >> (16 - (hi - lo))) << lo) : 0u;
            ls[t] = (int)((px[t] >> 24) & 0x7fu);
            lp[t] = lam_b + (int)(py[t] + (unsigned)(kb + 1) * (unsigned)ls[t]);
        }
        // ---- phase A: the state entries (lanes 0..12) ---------------------------------------------------------------------------------
        const double* __restrict__ zp = Xb + (size_t)kb * n + (i < n ? i : 0);
        for (int k0 = 0; k0 < nk; k0 += 4) {
            double zi[4], lam[4][TO_EXP_MAXT];
            const unsigned a0 = act[0] >> k0, a1 = act[1] >> k0, a2 = act[2] >> k0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int kk = k0 + u;
                zi[u] = 0.0; lam[u][0] = 0.0; lam[u][1] = 0.0; lam[u][2] = 0.0;
                if (kk < nk && i < n) zi[u] = __ldg(zp + kk * n);
                if ((a0 >> u) & 1u) lam[u][0] = __ldg(lp[0] + kk * ls[0]);
                if ((a1 >> u) & 1u) lam[u][1] = __ldg(lp[1] + kk * ls[1]);
                if ((a2 >> u) & 1u) lam[u][2] = __ldg(lp[2] + kk * ls[2]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k0 + u >= nk) break;                                               // (uniform over the group)
                const int cid = __shfl_sync(gm, mycid, k0 + u, 16);
                if (i < n) {
                    if (cid != ccid) { const DevCost& c = P.costs[cid]; ca = c.Qd[i]; cb = c.q[i]; ccid = cid; }
                    double g = fma(ca, zi[u], cb), h = ca;
#pragma unroll
                    for (int t = 0; t < TO_EXP_MAXT; t++) {
                        if ((((t == 0) ? a0 : (t == 1) ? a1 : a2) >> u) & 1u) {
                            const double lb = fma(nms[t], zi[u] - bnd[t], lam[u][t]);   // lambda - mu c
                            if ((px[t] >> 31) || lb <= 0.0) { g += (nms[t] < 0.0) ? -lb : lb; h += fabs(nms[t]); }   // g -= sign lb ; h += mu
                        }
                    }
                    double* row = stage[k0 + u];
                    if (quat) { double* a = row + HB_ + 3 * (i - qs); a[0] = g; a[1] = h; a[2] = zi[u]; }   // (g, h, q) of q_w..q_z wait in Hb[0..11] for phase B
                    else {
                        row[G_ + pme] = g; row[HD_ + pme] = h;
                        if (e == 7) { row[HB_ + 12] = 0.0; row[HB_ + 13] = 0.0; row[HB_ + 14] = 0.0; row[HB_ + 15] = h; }   // p = 14 is row 3 of Hb
                    }
                }
            }
        }
        __syncwarp(gm);
        if (i < nk) {
            double* row = stage[i];
            // ---- phase B: attitude block of knot kb + i ---------------------------------------------------------------------------
            double gq[4], hq[4], q[4];
#pragma unroll
            for (int r = 0; r < 4; r++) { gq[r] = row[HB_ + 3 * r]; hq[r] = row[HB_ + 3 * r + 1]; q[r] = row[HB_ + 3 * r + 2]; }
            // rows of G' = (L(q) H)': (-x,w,z,-y), (-y,-z,w,x), (-z,y,-x,w)
            const double G0[4] = {-q[1], q[0], q[3], -q[2]}, G1[4] = {-q[2], -q[3], q[0], q[1]}, G2[4] = {-q[3], q[2], -q[1], q[0]};
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                double gc[4];
#pragma unroll
                for (int r = 0; r < 4; r++) gc[r] = (cc == 0) ? G0[r] : (cc == 1) ? G1[r] : G2[r];
                double qb = 0.0, ge = 0.0, hb0 = 0.0, hb1 = 0.0, hb2 = 0.0;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    qb += q[r] * gq[r]; ge += gc[r] * gq[r];
                    const double tt = gc[r] * hq[r];
                    hb0 += tt * G0[r]; hb1 += tt * G1[r]; hb2 += tt * G2[r];
                }
                const double hd = ((cc == 0) ? hb0 : (cc == 1) ? hb1 : hb2) - qb;
                const int p = 8 + 2 * cc;                                             // attitude error e = 3 + c sits on p = 8, 10, 12 (frag_layout.cuh)
                row[G_ + p] = ge; row[HD_ + p] = hd;
                row[HB_ + 4 * cc + 0] = (cc == 0) ? hd : hb0;
                row[HB_ + 4 * cc + 1] = (cc == 1) ? hd : hb1;
                row[HB_ + 4 * cc + 2] = (cc == 2) ? hd : hb2;
                row[HB_ + 4 * cc + 3] = 0.0;
            }
            // ---- phase C: the control entries of knot kb + i (coordinate 12 + a, physical slot 2a) ----------------------------------
            const int k = kb + i;
            const DevCost& c = P.costs[mycid];
            double zu[m], lu[m][TO_EXP_MAXT];
#pragma unroll
            for (int a = 0; a < m; a++) {                                             // every load of the phase first
                zu[a] = (k != N - 1) ? __ldg(Ub + (size_t)k * m + a) : 0.0;
#pragma unroll
                for (int t = 0; t < TO_EXP_MAXT; t++) {
                    const unsigned rx = __ldg(&tab.pkx[t][n + a]);
                    lu[a][t] = (k != N - 1 && (unsigned)(k + 1) - (rx & 0xfffu) <= ((rx >> 12) & 0xfffu))
                                   ? __ldg(lam_b + (int)(__ldg(&tab.pky[t][n + a]) + (unsigned)(k + 1) * ((rx >> 24) & 0x7fu))) : 0.0;
                }
            }
#pragma unroll
            for (int a = 0; a < m; a++) {
                double g = 0.0, h = 0.0;
                if (k != N - 1) {
                    g = fma(c.Rd[a], zu[a], c.r[a]); h = c.Rd[a];
#pragma unroll
                    for (int t = 0; t < TO_EXP_MAXT; t++) {
                        const unsigned rx = __ldg(&tab.pkx[t][n + a]);
                        if ((unsigned)(k + 1) - (rx & 0xfffu) <= ((rx >> 12) & 0xfffu)) {
                            const double rn = __ldg(&tab.nms[t][n + a]);
                            const double lb = fma(rn, zu[a] - __ldg(&tab.bound[t][n + a]), lu[a][t]);
                            if ((rx >> 31) || lb <= 0.0) { g += (rn < 0.0) ? -lb : lb; h += fabs(rn); }
                        }
                    }
                }
                row[G_ + 2 * a] = g; row[HD_ + 2 * a] = h;
            }
        }
        __syncwarp(gm);
        // ---- write-out: 384 contiguous bytes per knot, 16 bytes per lane and store -------------------------------------------------------
        for (int kk = 0; kk < nk; kk++) {
            const double* row = stage[kk];
            double* dst = recb + (size_t)kk * TO_REC_LEN;
            *reinterpret_cast<double2*>(dst + 2 * i) = make_double2(row[2 * i], row[2 * i + 1]);
            if (i < 8) *reinterpret_cast<double2*>(dst + 32 + 2 * i) = make_double2(row[32 + 2 * i], row[33 + 2 * i]);
        }
        __syncwarp(gm);                                                               // the stage is free again
    }
}
cudaError_t launch_expansion_rec16(const DevProblem& P, cudaStream_t s, int mode) {
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int v1 = getenv("TO_CEXP_V1") ? atoi(getenv("TO_CEXP_V1")) : 0;
    if (!v1 && P.n == 13 && P.m == 4) {
        // blocks of 16 knots, one 16-lane group each, TO_CEXP2_UNITS blocks per group (grid-stride)
        static const int upg = getenv("TO_CEXP2_UNITS") ? atoi(getenv("TO_CEXP2_UNITS")) : 1;   // r02u: 1.369 / 1.384 / 1.411 ms per step at 1 / 2 / 4
        const long long units = (long long)P.B * ((P.N + 15) / 16);
        constexpr int GPB = TO_CEXP2_THREADS / 16;                                   // 16-lane groups per CTA
        long long blocks = ((units + GPB - 1) / GPB + upg - 1) / (upg < 1 ? 1 : upg);
        if (blocks < sms) blocks = sms;
        { static bool done[TO_MAXDEV] = {false}; prefer_common_carveout(k_expansion_rec16b, done); }
        k_expansion_rec16b<<<(unsigned)blocks, TO_CEXP2_THREADS, 0, s>>>(P, mode);
        return cudaGetLastError();
    }
    // 16-lane groups, a few knots each (grid-stride): the per-lane term table stays in registers
    const long long total = (long long)P.B * P.N * 16;
    constexpr int T = TO_CEXP_THREADS;
    long long blocks = ((total + T - 1) / T + TO_CEXP_ITERS - 1) / TO_CEXP_ITERS;
    if (blocks < sms) blocks = sms;
    { static bool done[TO_MAXDEV] = {false}; prefer_common_carveout(k_expansion_rec16, done); }
    k_expansion_rec16<<<(unsigned)blocks, T, 0, s>>>(P, mode);
    return cudaGetLastError();
}

cudaError_t launch_expand_lie(const DevProblem& P, cudaStream_t s, int mode) {
    if (P.model != MODEL_QUADROTOR) return cudaErrorNotSupported;
    const long long total = (long long)P.B * (P.N - 1) * 10;      // 10 dual-number seeds per knot (k_expand_lie: seed pruning)
    static_assert(fraglayout::phys_z(0) == 1 && fraglayout::phys_z(5) == 12 && fraglayout::phys_z(11) == 15 && fraglayout::phys_z(12) == 0 && fraglayout::phys_z(15) == 6, "nibble table of k_expand_lie");
    { static bool d1[TO_MAXDEV] = {false}, d2[TO_MAXDEV] = {false}; prefer_common_carveout(k_expand_lie<MODEL_QUADROTOR, true>, d1); prefer_common_carveout(k_expand_lie<MODEL_QUADROTOR, false>, d2); }
    constexpr int T = TO_EXPAND_LIE_THREADS;
    static const int staged = getenv("TO_EXPAND_STAGE") ? atoi(getenv("TO_EXPAND_STAGE")) : 0;
    if (P.frag && staged) {
        constexpr int KPC = T / 10;
        const long long nknots = (long long)P.B * (P.N - 1);
        { static bool d3[TO_MAXDEV] = {false}; prefer_common_carveout(k_expand_lie_staged<MODEL_QUADROTOR>, d3); }
        k_expand_lie_staged<MODEL_QUADROTOR><<<(unsigned)((nknots + KPC - 1) / KPC), T, 0, s>>>(P, mode);
    } else if (P.frag) k_expand_lie<MODEL_QUADROTOR, true><<<(unsigned)((total + T - 1) / T), T, 0, s>>>(P, mode);
    else k_expand_lie<MODEL_QUADROTOR, false><<<(unsigned)((total + T - 1) / T), T, 0, s>>>(P, mode);
    return cudaGetLastError();
}

cudaError_t launch_rollout(const DevProblem& P, cudaStream_t s) {
    const int threads = 32, blocks = (P.B + threads - 1) / threads;
    TO_DISPATCH_MODEL(P.model, P.m, (k_rollout<MODEL><<<blocks, threads, 0, s>>>(P)));
    return cudaGetLastError();
}

template <int MODEL, int NP>
static cudaError_t launch_expand_t(const DevProblem& P, cudaStream_t s, int mode) {
    constexpr int TPK = (SeedList<MODEL>::count + NP - 1) / NP;
    const long long total = (long long)P.B * (P.N - 1) * TPK;
    const int threads = 128;
    { static bool done[TO_MAXDEV] = {false}; prefer_common_carveout(k_expand<MODEL, NP>, done); }
    k_expand<MODEL, NP><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P, mode);
    return cudaGetLastError();
}

cudaError_t launch_expand(const DevProblem& P, cudaStream_t s, int mode) {
    // seeds per thread: 1 (value recomputed per seed) or 2 (value shared by two seeds, more registers); profiles/r01_notes.md
    static int np = -1;
    if (np < 0) { const char* v = getenv("TO_EXPAND_SEEDS"); np = v ? atoi(v) : 1; }
    cudaError_t e = cudaErrorNotSupported;
    if (np == 2) { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_expand_t<MODEL, 2>(P, s, mode))); }
    else { TO_DISPATCH_MODEL(P.model, P.m, (e = launch_expand_t<MODEL, 1>(P, s, mode))); }
    return e;
}
