// stage_small.h -- host interface of the warp-per-partition Stage-A kernel (stage_small.cu, small_core.cuh).
#pragma once
#include "ambi_common.cuh"

namespace ambi {

struct SmallConfig {
    int C = 0, tsize = 0;                 // records per level and hash slots per warp
    int warps_per_block = 0, blocks = 0;
    size_t smem_bytes = 0, arena_words = 0;
};

struct SmallLaunch {
    const DevProblem *probs; const uint32_t *plist; uint32_t count; uint32_t *work;
    int32_t *best_score; uint32_t *best_rows; int32_t *incumbent;
    unsigned long long *cand, *node, *funnel;
    uint32_t *fb_count; uint8_t *active; uint32_t *arena; unsigned long long *stamps;
    int beam_width; float beam_growth; int beam_delta, beam_shift;
};

// sizes the launch for `C` records per level on a device with `sms` multiprocessors; 0 on success
int small_configure(SmallConfig *cfg, int C, int sms);
void small_launch(const SmallConfig &cfg, const SmallLaunch &L, cudaStream_t st);
void stamp_launch(unsigned long long *d_t, cudaStream_t st);

}  // namespace ambi
