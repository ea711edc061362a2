// Microbenchmark: FP64 vector FMA (DFMA) vs FP64 tensor MMA (DMMA m8n8k4) throughput on sm_100a,
// alone and mixed, to decide which pipe the Riccati n x n products should use.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096

__global__ void k_dfma(double* out, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per iteration 8 DMMA (=2048 FMA/warp) + 64 DFMA (=2048 FMA/warp)
__global__ void k_mixed(double* out, double a, double b) {
    double c[8][2], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; acc[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = fma(acc[j], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory broadcast load + DFMA: how many LDS per DFMA can we afford
template <int NLD>
__global__ void k_lds_dfma(double* out, double a) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    int base = (threadIdx.x >> 5) * 8;
    for (int it = 0; it < ITERS; it++) {
        double v[NLD];
#pragma unroll
        for (int l = 0; l < NLD; l++) v[l] = sm[(base + it * NLD + l) & 1023];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, v[i % NLD]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 5; r++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int blocks = p.multiProcessorCount * 4, threads = 512;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    double nthr = (double)blocks * threads;
    float t;
    t = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("DFMA : %.3f ms  %.2f TFLOP/s\n", t, nthr * ITERS * 8 * 2 / t / 1e9);
    t = timeit([&] { k_dmma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("DMMA : %.3f ms  %.2f TFLOP/s\n", t, (nthr / 32) * ITERS * 256 * 2 / t / 1e9);
    t = timeit([&] { k_mixed<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("MIXED: %.3f ms  %.2f TFLOP/s (half DMMA half DFMA)\n", t,
           ((nthr / 32) * ITERS * 256 * 2 + nthr * ITERS * 8 * 2) / t / 1e9);
    t = timeit([&] { k_lds_dfma<1><<<blocks, threads>>>(out, 1.0000001); });
    printf("LDS1+8DFMA: %.3f ms  %.2f TFLOP/s\n", t, nthr * ITERS * 8 * 2 / t / 1e9);
    t = timeit([&] { k_lds_dfma<2><<<blocks, threads>>>(out, 1.0000001); });
    printf("LDS2+8DFMA: %.3f ms  %.2f TFLOP/s\n", t, nthr * ITERS * 8 * 2 / t / 1e9);
    t = timeit([&] { k_lds_dfma<4><<<blocks, threads>>>(out, 1.0000001); });
    printf("LDS4+8DFMA: %.3f ms  %.2f TFLOP/s\n", t, nthr * ITERS * 8 * 2 / t / 1e9);
    t = timeit([&] { k_lds_dfma<8><<<blocks, threads>>>(out, 1.0000001); });
    printf("LDS8+8DFMA: %.3f ms  %.2f TFLOP/s\n", t, nthr * ITERS * 8 * 2 / t / 1e9);
    return 0;
}
