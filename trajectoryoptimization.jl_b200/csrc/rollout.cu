// rollout.cu -- kernel 1 of the hot path: batched RK4 rollout and dual-number dynamics expansion.
//
//   k_rollout : rollout!(prob)                         reference src/problem.jl:330-340
//               x_1 = x0 ; x_k = RK4(x_{k-1}, u_{k-1}, dt_{k-1}).  Serial in k, parallel over instances:
//               one thread per instance (the recursion has no intra-instance parallelism worth a warp).
//   k_expand  : RD.jacobian!(ForwardAD) on the discretised dynamics at every knot (no call site inside the
//               reference; shape [A B] = n x (n+m) pinned by test/dynamics_constraints.jl:35,57-62).
//               One thread per (instance, knot, seed direction j): the RK4 step is pushed through a
//               Dual<1> whose tangent is the one-hot e_j, i.e. the thread computes column j of [A B] with the
//               partial carried in registers.  Threads of one knot are adjacent, so row i of AB (LDAB
//               contiguous doubles incl. the zero pad column) is written by LDAB adjacent lanes.
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

template <int MODEL>
__global__ void __launch_bounds__(128) k_expand(const DevProblem P) {
    constexpr int n = ModelDims<MODEL>::n, m = ModelDims<MODEL>::m;
    using D = Dual<1>;
    const int ld = P.ldab;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.B * (P.N - 1) * ld;
    if (t >= total) return;
    const int j = (int)(t % ld);
    const long long bk = t / ld;
    const int k = (int)(bk % (P.N - 1));
    const int b = (int)(bk / (P.N - 1));
    double* AB = P.AB + ((size_t)b * (P.N - 1) + k) * n * ld;
    if (j >= n + m) {   // pad column
#pragma unroll
        for (int i = 0; i < n; i++) AB[i * ld + j] = 0.0;
        return;
    }
    const double* X = traj_X(P, P.cur[b], b) + (size_t)k * n;
    const double* U = traj_U(P, P.cur[b], b) + (size_t)k * m;
    D x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i].v = X[i]; x[i].d[0] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int i = 0; i < m; i++) { u[i].v = U[i]; u[i].d[0] = (n + i == j) ? 1.0 : 0.0; }
    rk4_step<MODEL, D>(P.params, x, u, P.dt[k], xn);
#pragma unroll
    for (int i = 0; i < n; i++) AB[i * ld + j] = xn[i].d[0];
}

cudaError_t launch_rollout(const DevProblem& P, cudaStream_t s) {
    const int threads = 64, blocks = (P.B + threads - 1) / threads;
    TO_DISPATCH_MODEL(P.model, P.m, (k_rollout<MODEL><<<blocks, threads, 0, s>>>(P)));
    return cudaGetLastError();
}

cudaError_t launch_expand(const DevProblem& P, cudaStream_t s) {
    const long long total = (long long)P.B * (P.N - 1) * P.ldab;
    const int threads = 128;
    const long long blocks = (total + threads - 1) / threads;
    TO_DISPATCH_MODEL(P.model, P.m, (k_expand<MODEL><<<(unsigned)blocks, threads, 0, s>>>(P)));
    return cudaGetLastError();
}
