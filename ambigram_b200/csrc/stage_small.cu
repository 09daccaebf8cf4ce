// stage_small.cu -- kernel wrapper of the warp-per-partition Stage-A search (small_core.cuh) and its host side.
//
// One launch solves every partition of the batch with 2..16 segments whose frontier fits the per-warp record
// capacity; the others are flagged for the level-synchronous global-memory path of stage_a.cu.
#include "ambi_common.cuh"
#include "small_core.cuh"
#include "stage_small.h"
#include <cstdio>

namespace ambi {

__global__ void __launch_bounds__(128, 4)
k_solve_small(small::Params P_param)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ small::Params s_P;
    if (threadIdx.x == 0) s_P = P_param;
    __syncthreads();
    const small::Params &P = s_P;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t *base = s_raw + small::warp_bytes(P.C, P.tsize) * wib;
    small::Fixed &S = *reinterpret_cast<small::Fixed *>(base);
    const small::Arr R = small::carve(base, P.C, P.tsize);
    const uint32_t total = P.pcount ? min(*P.pcount, P.count) : P.count;
    uint32_t *arena = P.arena + (size_t)(blockIdx.x * (blockDim.x >> 5) + wib) * small::MAXN_SMALL * P.C;
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(P.work, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= total) break;
        const uint32_t prob = P.plist ? P.plist[wi] : wi;
        small::solve_one(P, S, R, prob, arena, lane);
    }
}

__global__ void k_stamp(unsigned long long *t) { *t = globaltimer_ns(); }

int small_configure(SmallConfig *cfg, int C, int sms)
{
    if (C < 32) C = 32;
    if (C > small::MAXC) C = small::MAXC;
    int tsize = 64; while (tsize < 2 * C) tsize <<= 1;
    const size_t wb = small::warp_bytes(C, tsize);
    const size_t smem_max = 227 * 1024 - 1024;          // per SM, minus the static part and the per-block reservation
    int wps = (int)(smem_max / wb);                     // warps per SM
    if (wps < 1) return -1;
    if (wps > 16) wps = 16;
    // blocks of 4 warps (several blocks per SM) keep the tail of a launch short; a warp's records must fit one block
    int wpb = wps >= 4 ? 4 : wps;
    int bps = wps / wpb;
    cfg->C = C; cfg->tsize = tsize; cfg->warps_per_block = wpb; cfg->blocks = sms * bps;
    cfg->smem_bytes = wb * wpb;
    cfg->arena_words = (size_t)cfg->blocks * wpb * small::MAXN_SMALL * C;
    cudaError_t e = cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes);
    if (e != cudaSuccess) return -2;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_small, wpb * 32, cfg->smem_bytes);
    if (e != cudaSuccess || per_sm < 1) return -3;
    cfg->blocks = sms * per_sm;
    cfg->arena_words = (size_t)cfg->blocks * wpb * small::MAXN_SMALL * C;
    return 0;
}

void small_launch(const SmallConfig &cfg, const SmallLaunch &L, cudaStream_t st)
{
    small::Params P;
    P.probs = L.probs; P.plist = L.plist; P.pcount = nullptr; P.count = L.count; P.work = L.work;
    P.best_score = L.best_score; P.best_rows = L.best_rows; P.incumbent = L.incumbent;
    P.cand = L.cand; P.node = L.node; P.funnel = L.funnel;
    P.fb_list = nullptr; P.fb_count = L.fb_count; P.active = L.active; P.arena = L.arena; P.stamps = L.stamps;
    P.C = cfg.C; P.tsize = cfg.tsize; P.beam_width = L.beam_width; P.beam_growth = L.beam_growth;
    P.beam_delta = L.beam_delta; P.beam_shift = L.beam_shift;
    // several contexts of one process may size the kernel differently: (re)state the limit this launch needs
    cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
    k_solve_small<<<cfg.blocks, cfg.warps_per_block * 32, cfg.smem_bytes, st>>>(P);
}

void stamp_launch(unsigned long long *d_t, cudaStream_t st) { k_stamp<<<1, 1, 0, st>>>(d_t); }

}  // namespace ambi
