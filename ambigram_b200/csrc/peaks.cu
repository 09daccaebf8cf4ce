// peaks.cu -- on-box micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not
// carry ("replace with a measured IADD/LDS micro-benchmark on the box").
//
//   k_peak_alu   independent IADD3/LOP3 chains (the alu pipe: what k_expand's integer work issues to)
//   k_peak_mix   the same with every other operation an IMAD (fma pipe), i.e. both integer pipes busy
//   k_peak_lds   conflict-free 128-bit shared-memory loads
//
// Each is timed with CUDA events on the ctx stream (best of a few launches after a warm-up).
#include "../../include/ambigram_b200.h"
#include "ambi_common.cuh"
#include <cstdio>

namespace ambi {

constexpr int PEAK_ITERS = 4096;
constexpr int PEAK_CHAINS = 8;

__global__ void __launch_bounds__(256) k_peak_alu(uint32_t *out, uint32_t seed)
{
    uint32_t v[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) v[k] = threadIdx.x * 2654435761u + k + seed;
    uint32_t b = seed | 1u, c = seed ^ 0x9e3779b9u;
#pragma unroll 1
    for (int i = 0; i < PEAK_ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k) {
            // one IADD3 and one LOP3 per chain and iteration (both alu pipe)
            asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(b));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[k]) : "r"(c));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s ^= v[k];
    if (s == 0x12345678u) out[0] = s;          // never true in practice; keeps the chains alive
}

__global__ void __launch_bounds__(256) k_peak_mix(uint32_t *out, uint32_t seed)
{
    uint32_t v[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) v[k] = threadIdx.x * 2654435761u + k + seed;
    uint32_t b = seed | 1u, c = (seed ^ 0x9e3779b9u) | 3u;
#pragma unroll 1
    for (int i = 0; i < PEAK_ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(b));                 // alu pipe
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(c), "r"(b));  // fma pipe
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s ^= v[k];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(256) k_peak_lds(uint32_t *out, uint32_t seed)
{
    __shared__ uint4 buf[1024];                // 16 KB
    for (int i = threadIdx.x; i < 1024; i += 256) buf[i] = make_uint4(i, seed, i ^ seed, 1u);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < PEAK_ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = buf[(idx + k * 256) & 1023];      // consecutive lanes -> consecutive 16 B: conflict free
            acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        }
        idx = (idx + (int)(acc.w & 0u) + 32) & 1023;            // data-dependent in form only
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) out[0] = acc.x;
}

template <class K> static int time_kernel(K kern, cudaStream_t st, int blocks, uint32_t *d_out, double *best_ms)
{
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
    *best_ms = 1e30;
    for (int r = 0; r < 6; ++r) {              // first two are warm-up
        cudaEventRecord(e0, st);
        kern<<<blocks, 256, 0, st>>>(d_out, 12345u + r);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return -1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < *best_ms) *best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int measure_peaks(cudaStream_t st, ambi_peaks *out)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t *d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return -1;
    const int blocks = sms * 8;                // 2048 threads per SM: every scheduler has 16 warps to pick from
    const double threads = (double)blocks * 256.0;
    double ms = 0;
    int rc = time_kernel(k_peak_alu, st, blocks, d_out, &ms);
    if (!rc) { out->alu_tops = threads * PEAK_ITERS * PEAK_CHAINS * 2.0 / (ms * 1e-3) / 1e12; rc = time_kernel(k_peak_mix, st, blocks, d_out, &ms); }
    if (!rc) { out->alu_fma_tops = threads * PEAK_ITERS * PEAK_CHAINS * 2.0 / (ms * 1e-3) / 1e12; rc = time_kernel(k_peak_lds, st, blocks, d_out, &ms); }
    if (!rc) out->smem_tbs = threads * PEAK_ITERS * 4.0 * 16.0 / (ms * 1e-3) / 1e12;
    out->sms = sms;
    cudaFree(d_out);
    return rc;
}

}  // namespace ambi
