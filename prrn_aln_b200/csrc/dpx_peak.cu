// dpx_peak.cu -- register-only micro-benchmark of the DPX issue rate (the roofline denominator of
// the DP fill kernels): independent chains of __viaddmax_s32 / __viaddmax_s16x2, no memory traffic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pg_internal.h"

namespace {

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

template <bool PACKED>
__global__ void __launch_bounds__(256) dpx_chain_kernel(int seed, int* sink)
{
    int a[CHAINS];
    int b = seed * 3 + 1, c = seed - 5;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x + i * seed;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (PACKED) a[i] = (int)__viaddmax_s16x2((unsigned)a[i], (unsigned)b, (unsigned)c);
            else a[i] = __viaddmax_s32(a[i], b, c);
        }
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i];
    if (r == 0x7fffffff) sink[0] = r;   // never true in practice; keeps the chains alive
}

template <bool PACKED>
cudaError_t run_one(int sm_count, cudaStream_t st, double* gops)
{
    int* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(int));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sm_count * 8, threads = 256;
    dpx_chain_kernel<PACKED><<<blocks, threads, 0, st>>>(1, sink);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        dpx_chain_kernel<PACKED><<<blocks, threads, 0, st>>>(rep + 2, sink);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return e;
    double instr = (double)blocks * threads * (double)CHAINS * ITERS;
    *gops = instr / (best * 1e-3) / 1e9;
    return cudaGetLastError();
}

}  // namespace

cudaError_t dpx_peak_run(int sm_count, cudaStream_t st, double* gops_s32, double* gops_s16x2)
{
    cudaError_t e = run_one<false>(sm_count, st, gops_s32);
    if (e != cudaSuccess) return e;
    return run_one<true>(sm_count, st, gops_s16x2);
}
