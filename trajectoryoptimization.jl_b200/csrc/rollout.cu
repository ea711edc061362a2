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

template <int MODEL>
__global__ void __launch_bounds__(64) k_rollout(const DevProblem P) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    double* X = traj_Xw(P, P.cur[b], b);
    const double* U = traj_U(P, P.cur[b], b);
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i] = P.x0[(size_t)b * n + i]; X[i] = x[i]; }
    for (int k = 0; k < P.N - 1; k++) {
#pragma unroll
        for (int i = 0; i < m; i++) u[i] = U[k * m + i];
        rk4_step<MODEL, double>(P.params, x, u, P.dt[k], xn);
#pragma unroll
        for (int i = 0; i < n; i++) { x[i] = xn[i]; X[(k + 1) * n + i] = xn[i]; }
    }
}

template <int MODEL, int NP>
__global__ void __launch_bounds__(128) k_expand(const DevProblem P, int mode) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, nm = n + m;
    constexpr int TPK = (nm + NP - 1) / NP;        // threads per knot: each carries NP seed directions
    using D = Dual<NP>;
    const int ld = P.ldab;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.B * (P.N - 1) * TPK;
    if (t >= total) return;
    const int j0 = (int)(t % TPK) * NP;            // first seed of this thread
    const long long bk = t / TPK;
    const int k = (int)(bk % (P.N - 1));
    const int b = (int)(bk / (P.N - 1));
    if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) return;     // overlapped expansion: this launch covers the other group
    double* AB = P.AB + ((size_t)b * (P.N - 1) + k) * n * ld;     // pad columns nm..ld-1 stay zero from to_create
    const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
    D x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) {
        x[i].v = X[i];
#pragma unroll
        for (int s = 0; s < NP; s++) x[i].d[s] = (i == j0 + s) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < m; i++) {
        u[i].v = U[i];
#pragma unroll
        for (int s = 0; s < NP; s++) u[i].d[s] = (n + i == j0 + s) ? 1.0 : 0.0;
    }
    rk4_step<MODEL, D>(P.params, x, u, P.dt[k], xn);
#pragma unroll
    for (int i = 0; i < n; i++) {
        if (NP == 2 && j0 + 1 < nm) *reinterpret_cast<double2*>(&AB[i * ld + j0]) = make_double2(xn[i].d[0], xn[i].d[1]);
        else {
#pragma unroll
            for (int s = 0; s < NP; s++) if (j0 + s < nm) AB[i * ld + j0 + s] = xn[i].d[s];
        }
    }
}

// Error-state expansion of a Lie-group model (Quadrotor; SURVEY 8 f2, lie.cu): thread (instance, knot, j) pushes the j-th column of
// E(x_k) = blkdiag(I3, G(q_k), I6 | I4) through the RK4 step as the tangent of a Dual<1> -- the directional derivative [A G_k | B] e_j --
// and projects the result with G(q_{k+1})' (q_{k+1} from the stored trajectory, as Altro's errstate_jacobian! does).  It writes column j
// of [A_e B_e]_k (12 contiguous doubles, col-major 12 x 16): 1.5 KB per knot instead of the 2 KB of the padded full-state [A B].
// FRAG: the column goes into the fragment block of the knot's record (frag_layout.cuh) instead of P.ABe.
template <int MODEL, bool FRAG>
__global__ void __launch_bounds__(128) k_expand_lie(const DevProblem P, int mode) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m, ne = n - 1, nme = ne + m, qs = 3;
    using D = Dual<1>;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.B * (P.N - 1) * nme;
    if (t >= total) return;
    const int j = (int)(t % nme);
    const long long bk = t / nme;
    const int k = (int)(bk % (P.N - 1));
    const int b = (int)(bk / (P.N - 1));
    if (mode != 0 && (P.acc1[b] != 0) != (mode == 1)) return;
    const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
    D x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i].v = X[i]; x[i].d[0] = 0.0; }
#pragma unroll
    for (int i = 0; i < m; i++) { u[i].v = U[i]; u[i].d[0] = (ne + i == j) ? 1.0 : 0.0; }
    if (j < qs) {
#pragma unroll
        for (int i = 0; i < qs; i++) x[i].d[0] = (i == j) ? 1.0 : 0.0;
    } else if (j < qs + 3) {
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
    rk4_step<MODEL, D>(P.params, x, u, P.dt[k], xn);
    const double* q1 = X + n + qs;                                     // attitude of knot k + 1
    const double w1 = q1[0], x1 = q1[1], y1 = q1[2], z1 = q1[3];
    double col[ne];
#pragma unroll
    for (int i = 0; i < qs; i++) col[i] = xn[i].d[0];
    const double t0 = xn[qs].d[0], t1 = xn[qs + 1].d[0], t2 = xn[qs + 2].d[0], t3 = xn[qs + 3].d[0];
    col[qs] = -x1 * t0 + w1 * t1 + z1 * t2 - y1 * t3;
    col[qs + 1] = -y1 * t0 - z1 * t1 + w1 * t2 + x1 * t3;
    col[qs + 2] = -z1 * t0 + y1 * t1 - x1 * t2 + w1 * t3;
#pragma unroll
    for (int i = qs + 4; i < n; i++) col[i - 1] = xn[i].d[0];
    if constexpr (FRAG) {
        // element (row e, column j) of the record: (ks(e)*32 + 4*(c & 7) + fc(e))*2 + (c >> 3), c = physical index of column j
        const int c = (int)((0x6420FDB9E7CA8531ULL >> (4 * j)) & 15);       // fraglayout::phys_z(j) as a nibble table
        double* rec = P.REC + ((size_t)b * P.N + k) * TO_REC_LEN + 8 * (c & 7) + (c >> 3);
#pragma unroll
        for (int e = 0; e < ne; e++) rec[fraglayout::ab_index(e, 12)] = col[e];   // column 12 (u_0, c = 0) has a zero column offset
    } else {
        double* out = P.ABe + ((size_t)bk * nme + j) * ne;
#pragma unroll
        for (int e = 0; e < ne; e++) out[e] = col[e];
    }
}

cudaError_t launch_expand_lie(const DevProblem& P, cudaStream_t s, int mode) {
    if (P.model != MODEL_QUADROTOR) return cudaErrorNotSupported;
    const long long total = (long long)P.B * (P.N - 1) * (P.ne + P.m);
    static_assert(fraglayout::phys_z(0) == 1 && fraglayout::phys_z(5) == 12 && fraglayout::phys_z(11) == 15 && fraglayout::phys_z(12) == 0 && fraglayout::phys_z(15) == 6, "nibble table of k_expand_lie");
    if (P.frag) k_expand_lie<MODEL_QUADROTOR, true><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(P, mode);
    else k_expand_lie<MODEL_QUADROTOR, false><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(P, mode);
    return cudaGetLastError();
}

cudaError_t launch_rollout(const DevProblem& P, cudaStream_t s) {
    const int threads = 64, blocks = (P.B + threads - 1) / threads;
    TO_DISPATCH_MODEL(P.model, P.m, (k_rollout<MODEL><<<blocks, threads, 0, s>>>(P)));
    return cudaGetLastError();
}

template <int MODEL, int NP>
static cudaError_t launch_expand_t(const DevProblem& P, cudaStream_t s, int mode) {
    constexpr int nm = ModelDims<MODEL>::n + ModelDims<MODEL>::m;
    constexpr int TPK = (nm + NP - 1) / NP;
    const long long total = (long long)P.B * (P.N - 1) * TPK;
    const int threads = 128;
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
