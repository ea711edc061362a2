// Microbenchmark: dependent-issue latencies (cycles) of the instructions on the critical chain of the Riccati kernels, one warp on one SM:
// DFMA, DMUL, DADD (full warp / lower half-warp), MUFU.RCP64H + Newton, DMMA.8x8x4 (dependent accumulator chain, and 4 interleaved chains),
// SHFL.IDX (64-bit = 2 x 32-bit), LDS.64, STS.64 -> __syncwarp -> LDS.64 round trip.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(long long* out, double* sink, double a, double b, int half) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = lane; sm[lane + 32] = lane;
    __syncwarp();
    double x = 1.0 + lane * 1e-9, y = 0.5, c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    long long t0, t1;
    int r = 0;
#define TIC t0 = clock64();
#define TOC if (lane == 0) out[r] = clock64() - t0; r++;
    TIC for (int i = 0; i < N; i++) x = fma(x, a, b); TOC                                         // 0 DFMA chain, full warp
    TIC if (lane < 16) for (int i = 0; i < N; i++) x = fma(x, a, b); TOC                           // 1 DFMA chain, half warp
    TIC for (int i = 0; i < N; i++) x = x * a; TOC                                                 // 2 DMUL
    TIC for (int i = 0; i < N; i++) x = x + b; TOC                                                 // 3 DADD
    TIC for (int i = 0; i < N / 8; i++) { double yy; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(yy) : "d"(x)); double e = fma(-x, yy, 1.0); yy = fma(yy, e, yy); e = fma(-x, yy, 1.0); x = fma(yy, e, yy) + 1.5; } TOC   // 4 reciprocal (MUFU + 4 FMA + add), N/8 of them
    TIC for (int i = 0; i < N; i++) dmma(c0, c1, a, b); TOC                                        // 5 DMMA dependent chain
    TIC for (int i = 0; i < N / 4; i++) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); } TOC   // 6 four interleaved chains (N DMMA)
    TIC for (int i = 0; i < N; i++) { c0 = 0; dmma(c0, c1, x, b); x = c0 + 1e-300; } TOC           // 7 DMMA whose A operand depends on the previous result (+ DADD)
    TIC for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); TOC          // 8 SHFL 64-bit
    TIC for (int i = 0; i < N; i++) x = sm[(int)x & 31] ; TOC                                      // 9 LDS dependent (address from the loaded value)
    TIC for (int i = 0; i < N; i++) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); } TOC   // 10 STS -> syncwarp -> LDS (+ DADD) -> syncwarp
    sink[lane] = x + y + c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
    (void)half;
}
int main() {
    long long* out; double* sink; cudaMallocManaged(&out, 64 * 8); cudaMalloc(&sink, 32 * 8);
    k<<<1, 32>>>(out, sink, 1.0000001, 1e-9, 0); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, sink, 1.0000001, 1e-9, 0); cudaDeviceSynchronize();
    const char* names[] = {"DFMA chain (full warp)", "DFMA chain (half warp)", "DMUL chain", "DADD chain", "rcp (MUFU + 4 DFMA + DADD)", "DMMA dependent chain",
                           "DMMA 4 interleaved chains", "DMMA -> DADD -> DMMA(A operand)", "SHFL.64 chain", "LDS.64 dependent", "STS->sync->LDS->DADD->sync"};
    const int div[] = {N, N, N, N, N / 8, N, N, N, N, N, N};
    for (int i = 0; i < 11; i++) printf("%-34s %8.1f cycles\n", names[i], (double)out[i] / div[i]);
    return 0;
}
