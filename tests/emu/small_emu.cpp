// small_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles ambigram_b200/csrc/small_core.cuh (the warp-per-partition
// Stage-A kernel body) with g++ against the 32-lane warp emulator, so that the kernel's logic can be checked
// against the oracle on a box without a GPU (tests/test_small_core_emu.py).  The product never loads this.
#define AMBI_EMU 1
#include "simt_emu.h"
#include "../../ambigram_b200/csrc/host_flatten.h"
#include "../../ambigram_b200/csrc/small_core.cuh"
#include <vector>

using namespace ambi;

struct Job {
    small::Params P; uint8_t *smem; uint32_t prob; uint32_t *arena; int result;
};

static void lane_entry(int lane, void *arg)
{
    Job *J = static_cast<Job *>(arg);
    small::Fixed &S = *reinterpret_cast<small::Fixed *>(J->smem);
    const small::Arr R = small::carve(J->smem, J->P.C, J->P.tsize);
    const int r = small::solve_one(J->P, S, R, J->prob, J->arena, lane);
    if (lane == 0) J->result = r;
}

// one partition through the emulated warp; -> 0 solved, 1 frontier did not fit C records, <0 bad input
extern "C" int emu_small_solve(int32_t n, const int32_t *c2, const int32_t *j2, int32_t loop_ub, int32_t n_tgs, const int32_t *tgs_pairs,
                               int32_t junc_info, int32_t C, int32_t beam_width, int32_t *x_out, int64_t *score2, int32_t *incumbent,
                               uint64_t *cand, uint64_t *nodes, uint64_t *funnel4)
{
    ambi_bfb_problem p; memset(&p, 0, sizeof p);
    p.n = n; p.start_id = 0; p.seg_cn2 = c2; p.fbi_cn2 = j2; p.loop_ub = loop_ub; p.bias = 1; p.n_tgs = n_tgs; p.tgs_pairs = tgs_pairs;
    p.flags = junc_info ? AMBI_FLAG_JUNC_INFO : 0u;
    DevProblem d; std::string why;
    if (to_dev(p, d, why)) return -1;
    if (n < 2 || n > small::MAXN_SMALL || C > small::MAXC) return -2;
    int tsize = 64; while (tsize < 2 * C) tsize <<= 1;
    std::vector<uint8_t> smem(small::warp_bytes(C, tsize) + 64);
    std::vector<uint32_t> arena((size_t)AMBI_MAXN * C, 0), rows(AMBI_MAXN, 0);
    int32_t best = 0x7fffffff, inc = 0;
    unsigned long long cnd = 0, nod = 0, fun[4] = {0, 0, 0, 0}, stamps[4];
    Job J; memset(&J.P, 0, sizeof J.P);
    J.P.probs = &d; J.P.count = 1; J.P.best_score = &best; J.P.best_rows = rows.data(); J.P.incumbent = &inc;
    J.P.cand = &cnd; J.P.node = &nod; J.P.funnel = fun; J.P.stamps = stamps; J.P.arena = arena.data();
    J.P.C = C; J.P.tsize = tsize; J.P.beam_width = beam_width; J.P.beam_growth = 0.5f; J.P.beam_delta = 8; J.P.beam_shift = 2;
    J.smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15); J.prob = 0; J.arena = arena.data(); J.result = -9;
    emu::run_warp(lane_entry, &J);
    if (J.result == 0) rows_to_element_cn(n, rows.data(), x_out);
    *score2 = best; *incumbent = inc; *cand = cnd; *nodes = nod;
    for (int i = 0; i < 4; ++i) funnel4[i] = fun[i];
    funnel4[3] = stamps[3] >> 8;       // widest exact level instead of the batch count
    return J.result;
}
