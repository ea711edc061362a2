// green_ctx.cu -- does an SM partition (CUDA green contexts, driver API through cudaGetDriverEntryPoint: no link against libcuda) work with
// the runtime API the library uses?  Checks: kernels launched with <<<>>> on a green-context stream stay on the partition's SMs; events created
// by the runtime can be recorded on green streams and waited for / timed from primary-context streams.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o green_ctx green_ctx.cu     run: ./green_ctx [sms of the small partition]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

__global__ void k_smid(int* out, long long spin) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
}
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define DR(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("FAIL %s: CUresult %d\n", #x, (int)r_); return 1; } } while (0)
template <class F> static bool entry(const char* name, F& fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { printf("no driver entry point %s\n", name); return false; }
    fn = (F)p; return true;
}
int main(int argc, char** argv) {
    const int want = argc > 1 ? atoi(argv[1]) : 16;
    RT(cudaSetDevice(0)); RT(cudaFree(0));
    CUresult (*pDeviceGet)(CUdevice*, int); CUresult (*pGetRes)(CUdevice, CUdevResource*, CUdevResourceType);
    CUresult (*pSplit)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    CUresult (*pDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int); CUresult (*pCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    CUresult (*pStream)(CUstream*, CUgreenCtx, unsigned int, int); CUresult (*pDestroy)(CUgreenCtx);
    if (!entry("cuDeviceGet", pDeviceGet) || !entry("cuDeviceGetDevResource", pGetRes) || !entry("cuDevSmResourceSplitByCount", pSplit) ||
        !entry("cuDevResourceGenerateDesc", pDesc) || !entry("cuGreenCtxCreate", pCreate) || !entry("cuGreenCtxStreamCreate", pStream) || !entry("cuGreenCtxDestroy", pDestroy)) return 1;
    CUdevice dev; DR(pDeviceGet(&dev, 0));
    CUdevResource all, small, rest; DR(pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int nb = 1; DR(pSplit(&small, &nb, &all, &rest, 0, (unsigned)want));
    printf("SMs: all %u, small %u, rest %u (groups %u)\n", all.sm.smCount, small.sm.smCount, rest.sm.smCount, nb);
    CUdevResourceDesc d1, d2; DR(pDesc(&d1, &small, 1)); DR(pDesc(&d2, &rest, 1));
    CUgreenCtx g1, g2; DR(pCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM)); DR(pCreate(&g2, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    int lo = 0, hi = 0; RT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUstream s1, s2; DR(pStream(&s1, g1, CU_STREAM_NON_BLOCKING, hi)); DR(pStream(&s2, g2, CU_STREAM_NON_BLOCKING, 0));
    cudaStream_t main_s; RT(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
    const int nblk = 2048; int* out[3]; for (int i = 0; i < 3; i++) RT(cudaMalloc(&out[i], nblk * sizeof(int)));
    cudaEvent_t e0, e1, e2, ej; RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1)); RT(cudaEventCreate(&e2)); RT(cudaEventCreateWithFlags(&ej, cudaEventDisableTiming));
    // fork from the primary-context stream, kernels on the two partitions at once, join
    k_smid<<<nblk, 128, 0, main_s>>>(out[2], 1000); RT(cudaGetLastError());
    RT(cudaEventRecord(ej, main_s));
    RT(cudaStreamWaitEvent((cudaStream_t)s1, ej, 0)); RT(cudaStreamWaitEvent((cudaStream_t)s2, ej, 0));
    RT(cudaEventRecord(e0, (cudaStream_t)s1));
    k_smid<<<nblk, 128, 0, (cudaStream_t)s1>>>(out[0], 20000); RT(cudaGetLastError());
    RT(cudaEventRecord(e1, (cudaStream_t)s1));
    k_smid<<<nblk, 128, 0, (cudaStream_t)s2>>>(out[1], 20000); RT(cudaGetLastError());
    RT(cudaEventRecord(e2, (cudaStream_t)s2));
    RT(cudaStreamWaitEvent(main_s, e1, 0)); RT(cudaStreamWaitEvent(main_s, e2, 0));
    k_smid<<<nblk, 128, 0, main_s>>>(out[2], 1000); RT(cudaGetLastError());
    RT(cudaStreamSynchronize(main_s));
    float ms = 0; RT(cudaEventElapsedTime(&ms, e0, e1)); printf("small-partition kernel: %.3f ms between runtime events recorded on the green stream\n", ms);
    std::vector<int> h(nblk);
    const char* names[3] = {"small", "rest", "primary"};
    std::set<int> sets[3];
    for (int i = 0; i < 3; i++) { RT(cudaMemcpy(h.data(), out[i], nblk * sizeof(int), cudaMemcpyDeviceToHost)); sets[i] = std::set<int>(h.begin(), h.end()); printf("%s stream ran on %zu distinct SMs\n", names[i], sets[i].size()); }
    int common = 0; for (int s : sets[0]) common += (int)sets[1].count(s);
    printf("SMs shared by the two partitions: %d\n", common);
    printf("%s\n", (sets[0].size() <= small.sm.smCount && common == 0) ? "GREEN OK" : "GREEN NOT CONFINED");
    pDestroy(g1); pDestroy(g2);
    return 0;
}
