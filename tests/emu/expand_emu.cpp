// expand_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the device bodies of the product's Stage-A kernels
// (ambigram_b200/csrc/stage_a.cu with AMBI_EMU: expand_body, dedup_body, finalize_body, beam_*_body,
// materialise_body, init_roots_body, set_limit_body -- everything on the launch side is compiled out) with g++ against
// the 32-lane warp emulator and plays the launch sequence of StageA::enqueue_incumbent / enqueue_exact (greedy dive,
// beam, exact pass; per level expand -> merge -> pick -> beam cut -> materialise) one emulated warp at a time, so that
// the logic of the level-synchronous GPU path can be checked against the oracle on a box without a GPU
// (tests/test_expand_emu.py).  The product never loads this.
#define AMBI_EMU 1
#include "simt_emu.h"
static inline unsigned long long globaltimer_ns() { return emu_globaltimer(); }
#include "../../ambigram_b200/csrc/host_flatten.h"
#include "../../ambigram_b200/csrc/stage_a.cu"
#include <functional>
#include <vector>

using namespace ambi;

namespace {

std::function<void()> g_body;
void lane_entry(int, void *) { g_body(); }

// one "kernel launch": a grid of one block of one warp (every body is a grid-stride or work-counter loop)
void launch(const std::function<void()> &body)
{
    emu::g_block_idx = 0; emu::g_grid_dim = 1; emu::g_block_dim = 32; emu::g_warp_in_block = 0;
    g_body = body;
    emu::run_warp(lane_entry, nullptr);
}

struct Search {
    int count, nmax, stride, kstride; uint32_t cap;
    bool scaled;
    std::vector<DevProblem> probs;
    std::vector<uint8_t> buf[2], kbuf, dbuf, active, smem;
    std::vector<uint2> lbuf;
    std::vector<uint32_t> list, best_rows, beam_rows, hist, ticket, cutq, c_work, c_child, c_front;
    std::vector<unsigned long long> table, cand, node, stamps;
    std::vector<int32_t> limit, best_score, beam_score, incumbent, seen, fmin;
    unsigned long long funnel[4] = {0, 0, 0, 0};
    uint32_t overflow = 0, gen = 1;
    int beam = 16;
    int shard_world = 1, shard_rank = 0, shard_level = -1;     // prefix sharding of the exact pass (StageA::enqueue_exact)

    void expand(const LevelArgs &L)
    {
        launch([&] {
            if (nmax <= 16) { if (scaled) expand_body<16, true>(L, smem.data()); else expand_body<16, false>(L, smem.data()); }
            else { if (scaled) expand_body<32, true>(L, smem.data()); else expand_body<32, false>(L, smem.data()); }
        });
    }

    // StageA::run_level
    void run_level(int level, int cur, int mode)
    {
        const bool merge = mode != 0, beam_pass = mode == 1;
        LevelArgs L; memset(&L, 0, sizeof L);
        L.probs = probs.data(); L.parents = buf[cur].data(); L.plist = nullptr; L.pcount = &c_front[level];
        L.children = kbuf.data(); L.links = lbuf.data(); L.dead = dbuf.data(); L.kstride = kstride; L.ccount = &c_child[level]; L.cap = cap;
        L.limit = limit.data(); L.cand_count = cand.data(); L.node_count = node.data(); L.funnel = funnel; L.work_counter = &c_work[level];
        L.overflow = &overflow; L.level = level; L.stride = stride; L.nmax = nmax; L.emit = mode == 0 ? EMIT_BEST : EMIT_ALL;
        L.beam_delta = beam_pass ? 8 : -1; L.beam_shift = 2; L.beam_delta_short = 16; L.beam_shift_short = 3; L.beam_short_n = 20;
        L.shard_world = mode == 2 ? shard_world : 1; L.shard_rank = shard_rank; L.shard_level = mode == 2 && shard_world > 1 ? shard_level : -1; L.gpw = 2;
        expand(L);
        int32_t *fmin_cur = fmin.data() + (size_t)level * count;
        if (beam_pass) std::fill(hist.begin(), hist.end(), 0u);
        if (merge) {
            DedupArgs D; memset(&D, 0, sizeof D);
            D.recs = kbuf.data(); D.dead = dbuf.data(); D.tshift = 1; D.count = &c_child[level]; D.cap = cap; D.table = table.data(); D.tmask = (uint32_t)table.size() - 1;
            D.gen_base = &gen; D.gen_off = (uint32_t)(mode * 40 + level); D.stride = kstride; D.level = level; D.overflow = &overflow;
            D.probs = probs.data(); D.fmin_cur = beam_pass ? fmin_cur : nullptr; D.nmax = nmax;
            D.links = lbuf.data(); D.parents = buf[cur].data(); D.pstride = stride;
            launch([&] { dedup_body(D); });
        }
        FinalizeArgs F; memset(&F, 0, sizeof F);
        F.probs = probs.data(); F.recs = kbuf.data(); F.dead = dbuf.data(); F.count = &c_child[level]; F.cap = cap;
        F.best_score = mode == 2 ? best_score.data() : beam_score.data(); F.best_rows = mode == 2 ? best_rows.data() : beam_rows.data();
        F.stride = kstride; F.level = level; F.limit = nullptr; F.merge_best = 0; F.overflow = &overflow; F.seen = seen.data(); F.stamps = stamps.data();
        F.links = lbuf.data(); F.parents = buf[cur].data(); F.pstride = stride;
        launch([&] { finalize_body(F); });
        MatArgs M; memset(&M, 0, sizeof M);
        M.probs = probs.data(); M.keys = kbuf.data(); M.dead = dbuf.data(); M.links = lbuf.data(); M.count = &c_child[level]; M.cap = cap;
        M.parents = buf[cur].data(); M.pstride = stride; M.out = buf[cur ^ 1].data(); M.out_count = &c_front[level + 1]; M.out_cap = cap;
        M.kstride = kstride; M.level = level; M.nmax = nmax; M.overflow = &overflow;
        if (beam_pass) {
            BeamArgs B; memset(&B, 0, sizeof B);
            B.probs = probs.data(); B.recs = kbuf.data(); B.dead = dbuf.data(); B.count = &c_child[level]; B.cap = cap;
            B.fmin = fmin_cur; B.hist = hist.data(); B.cutq = cutq.data() + (size_t)level * count; B.ticket = ticket.data() + (size_t)level * count;
            B.nprob = count; B.width = beam; B.width_floor = std::min(16, std::max(beam, beam * 4096 / std::max(count, 1)));
            B.stride = kstride; B.level = level; B.growth = 0.5f; B.width_div = 1;
            launch([&] { beam_hist_body(B); });
            launch([&] { beam_cut_body(B); });
            M.fmin = fmin_cur; M.cutq = B.cutq; M.ticket = B.ticket;
        }
        if (level + 1 < nmax) launch([&] { materialise_body(M); });
    }

    // StageA::run_pass
    void run_pass(int mode, bool first)
    {
        if (mode == 1) { std::fill(fmin.begin(), fmin.end(), 0x7f7f7f7f); std::fill(ticket.begin(), ticket.end(), 0u); }
        std::fill(c_work.begin(), c_work.end(), 0u); std::fill(c_child.begin(), c_child.end(), 0u); std::fill(c_front.begin(), c_front.end(), 0u);
        InitRootsArgs R; memset(&R, 0, sizeof R);
        R.probs = probs.data(); R.count = count; R.recs = buf[0].data(); R.stride = stride; R.nmax = nmax; R.list = list.data();
        R.lcount = &c_front[0]; R.limit = limit.data(); R.best_score = best_score.data(); R.first = first ? 1 : 0; R.active = active.data();
        launch([&] { init_roots_body(R); });
        int cur = 0;
        for (int a = 0; a < nmax; ++a) { run_level(a, cur, mode); cur ^= 1; }
    }

    void set_limit()
    {
        SetLimitArgs A; memset(&A, 0, sizeof A);
        A.probs = probs.data(); A.count = count; A.limit = limit.data(); A.best_score = beam_score.data(); A.incumbent = incumbent.data();
        A.active = active.data(); A.seen = seen.data(); A.stamps = stamps.data();
        launch([&] { set_limit_body(A); });
    }

    // StageA::enqueue_incumbent + enqueue_exact
    void run()
    {
        overflow = 0; gen += 128u;
        std::fill(beam_score.begin(), beam_score.end(), 0x7f7f7f7f);
        run_pass(0, true);
        set_limit();
        if (beam > 1) { std::fill(beam_score.begin(), beam_score.end(), 0x7f7f7f7f); run_pass(1, false); set_limit(); }
        run_pass(2, false);
    }
};

}  // namespace

// `count` partitions of 2..32 segments through the emulated level-synchronous search.
// c2 / j2 / tgs: [count][32] / [count][32] / [count][2 * max_tgs]; -> 0 solved, 1 a level did not fit `cap` records, < 0 bad input
static int solve_impl(int32_t count, const int32_t *n, const int32_t *c2, const int32_t *j2, const int32_t *loop_ub,
                      const int32_t *unit_shift, const int32_t *n_tgs, const int32_t *tgs, int32_t max_tgs, const int32_t *junc_info,
                      int32_t cap, int32_t beam_width, int32_t *x_out, int64_t *score2, int32_t *incumbent_out, uint64_t *cand_out,
                      uint64_t *funnel4, int world, int rank, int level, uint32_t *rows_out)
{
    Search S;
    S.shard_world = world; S.shard_rank = rank;
    S.count = count; S.probs.resize(count); S.nmax = 0; S.scaled = false; S.beam = beam_width > 0 ? beam_width : 16;
    for (int i = 0; i < count; ++i) {
        ambi_bfb_problem p; memset(&p, 0, sizeof p);
        p.n = n[i]; p.start_id = 0; p.seg_cn2 = c2 + (size_t)i * 32; p.fbi_cn2 = j2 + (size_t)i * 32; p.loop_ub = loop_ub[i]; p.bias = 1;
        p.n_tgs = n_tgs[i]; p.tgs_pairs = tgs + (size_t)i * 2 * max_tgs; p.flags = junc_info[i] ? AMBI_FLAG_JUNC_INFO : 0u; p.unit_shift = (uint32_t)unit_shift[i];
        std::string why;
        if (to_dev(p, S.probs[i], why) || p.n < 2) return -1;
        S.nmax = std::max(S.nmax, (int)p.n); S.scaled = S.scaled || p.unit_shift != 0;
    }
    S.stride = (int)ambi_node_stride(S.nmax); S.kstride = S.nmax <= 16 ? 32 : 48; S.cap = (uint32_t)cap;
    for (int k = 0; k < 2; ++k) S.buf[k].assign((size_t)cap * S.stride + 64, 0);
    S.kbuf.assign((size_t)cap * 48 + 64, 0); S.lbuf.assign((size_t)cap + 8, make_uint2(0, 0)); S.dbuf.assign((size_t)cap + 8, 0); S.list.assign((size_t)cap + 8, 0);
    size_t ts = 1024; while (ts < 4 * (size_t)cap) ts <<= 1;
    S.table.assign(ts, 0ull);
    S.active.assign(count, 1);
    S.smem.assign(std::max(sizeof(GroupShared<16>), sizeof(GroupShared<32>)) * WPB + 64, 0);
    S.best_rows.assign((size_t)count * AMBI_MAXN, 0); S.beam_rows.assign((size_t)count * AMBI_MAXN, 0);
    S.hist.assign((size_t)count * BEAM_BINS, 0); S.ticket.assign((size_t)count * AMBI_MAXN, 0); S.cutq.assign((size_t)count * AMBI_MAXN, 0);
    S.c_work.assign(40, 0); S.c_child.assign(40, 0); S.c_front.assign(41, 0);
    S.cand.assign(count, 0); S.node.assign(count, 0); S.stamps.assign((size_t)count * 4, 0);
    S.limit.assign(count, 0); S.best_score.assign(count, 0x7fffffff); S.beam_score.assign(count, 0x7fffffff);
    S.incumbent.assign(count, 0); S.seen.assign(count, 0x7fffffff); S.fmin.assign((size_t)count * AMBI_MAXN, 0x7f7f7f7f);
    for (int i = 0; i < count; ++i) S.incumbent[i] = S.probs[i].zero_score;
    S.shard_level = world > 1 ? std::min(level, std::max(0, S.nmax - 2)) : -1;
    S.run();
    if (S.overflow) return 1;
    if (rows_out) memcpy(rows_out, S.best_rows.data(), (size_t)count * AMBI_MAXN * 4);
    for (int i = 0; i < count; ++i) {
        score2[i] = S.best_score[i]; incumbent_out[i] = S.incumbent[i]; cand_out[i] = S.cand[i];
        if (S.best_score[i] != 0x7fffffff) rows_to_element_cn(S.probs[i].n, S.best_rows.data() + (size_t)i * AMBI_MAXN, x_out + (size_t)i * 32 * 33);
    }
    for (int k = 0; k < 4; ++k) funnel4[k] = S.funnel[k];
    return 0;
}

extern "C" int emu_expand_solve(int32_t count, const int32_t *n, const int32_t *c2, const int32_t *j2, const int32_t *loop_ub,
                                const int32_t *unit_shift, const int32_t *n_tgs, const int32_t *tgs, int32_t max_tgs, const int32_t *junc_info,
                                int32_t cap, int32_t beam_width, int32_t *x_out, int64_t *score2, int32_t *incumbent_out, uint64_t *cand_out,
                                uint64_t *funnel4)
{
    return solve_impl(count, n, c2, j2, loop_ub, unit_shift, n_tgs, tgs, max_tgs, junc_info, cap, beam_width, x_out, score2, incumbent_out,
                      cand_out, funnel4, 1, 0, -1, nullptr);
}

// The multi-GPU split of one search: every "rank" runs the incumbent part and the first
// `level` rows redundantly and keeps the children of row `level` whose prefix hashes to it; the ranks' optima are
// combined by (score, row codes in search order) like ambi_bfb_solve does after its all-gather.
extern "C" int emu_expand_solve_sharded(int32_t world, int32_t level, int32_t count, const int32_t *n, const int32_t *c2, const int32_t *j2,
                                        const int32_t *loop_ub, const int32_t *unit_shift, const int32_t *n_tgs, const int32_t *tgs, int32_t max_tgs,
                                        const int32_t *junc_info, int32_t cap, int32_t beam_width, int32_t *x_out, int64_t *score2, uint64_t *cand_total)
{
    std::vector<int64_t> best(count, 0x7fffffff);
    std::vector<uint32_t> best_rows((size_t)count * AMBI_MAXN, 0);
    std::vector<int32_t> x((size_t)count * 32 * 33), inc(count);
    std::vector<int64_t> sc(count);
    std::vector<uint64_t> cand(count);
    std::vector<uint32_t> rows((size_t)count * AMBI_MAXN);
    uint64_t fun[4];
    for (int i = 0; i < count; ++i) cand_total[i] = 0;
    for (int r = 0; r < world; ++r) {
        const int rc = solve_impl(count, n, c2, j2, loop_ub, unit_shift, n_tgs, tgs, max_tgs, junc_info, cap, beam_width, x.data(), sc.data(), inc.data(),
                                  cand.data(), fun, world, r, level, rows.data());
        if (rc) return rc;
        for (int i = 0; i < count; ++i) {
            cand_total[i] += cand[i];
            if (sc[i] == 0x7fffffff) continue;                     // this shard saw no completed assignment
            bool take = sc[i] < best[i];
            if (sc[i] == best[i]) {
                take = false;
                for (int k = 0; k < n[i]; ++k) { const uint32_t a = rows[(size_t)i * AMBI_MAXN + k], b = best_rows[(size_t)i * AMBI_MAXN + k]; if (a != b) { take = a < b; break; } }
            }
            if (take) {
                best[i] = sc[i];
                memcpy(&best_rows[(size_t)i * AMBI_MAXN], &rows[(size_t)i * AMBI_MAXN], AMBI_MAXN * 4);
                memcpy(x_out + (size_t)i * 32 * 33, x.data() + (size_t)i * 32 * 33, 32 * 33 * 4);
            }
        }
    }
    for (int i = 0; i < count; ++i) score2[i] = best[i];
    return 0;
}
