// small_core.cuh -- Stage A for partitions of up to 16 segments: ONE WARP PER PARTITION, the whole
// three-pass search (greedy dive -> beam -> exact, see DESIGN.md "Search formulation") inside shared memory.
//
// Replaces BFB_ILP + external cbc (reference src/LocalGenomicMap.cpp:4397-4752, localhap.cpp:173-212) for the
// partitions it takes; same model, same bound, same tie-break as the level-synchronous global-memory path in
// stage_a.cu (k_expand / k_dedup), which stays the path for longer partitions and for partitions whose frontier
// does not fit here.  What is different is where the frontier lives and who merges it:
//   * a frontier record is the 16 column bytes of the state (one uint4; byte 0 carries the running
//     third-generation count once column 0 is closed), its score and f -- 24 bytes in shared memory;
//   * children are merged ON ARRIVAL in a per-warp shared-memory hash table (same column state -> keep the
//     smaller (score, prefix in search order)), so only distinct states are ever stored;
//   * the prefix is not carried: a record knows its parent's RANK in prefix order (dense, recomputed per level by
//     counting), key = rank(parent) << 20 | row code orders children exactly as the full prefixes would, and one
//     back-pointer word per surviving record and level (global arena, L2 resident) rebuilds the winner's rows;
//   * candidate decode uses per-record position lists (nibble lists of the open columns, the class-0 loop
//     multisets listed once per record) and a run-length walk over the records' candidate counts instead of
//     population-count searches.
// No global-memory traffic inside a level, no kernel launch per level: one launch solves the whole batch.
//
// This header is compiled by nvcc for sm_100a (stage_small.cu) and -- TEST INFRASTRUCTURE -- by g++ against the
// 32-lane warp emulator tests/emu/simt_emu.h (tests/emu/small_emu.cpp), which is how the logic is checked
// against the oracle on a box without a GPU.
#pragma once
#include <stdint.h>
#include "decode_common.cuh"

namespace ambi {
namespace small {

#ifndef AMBI_BEAM_FLOOR
#define AMBI_BEAM_FLOOR 16           // narrowest beam (k_beam_cut of stage_a.cu uses the same floor)
#endif
constexpr int MAXN_SMALL = 16;          // one uint4 of column bytes
using dec::L0CAP;
using dec::below;
using dec::pack_cfg;
using dec::unpack_cfg;
constexpr uint32_t T_EMPTY = 0xffffffffu;
constexpr int MAXC = 4096;              // rank / parent index have 12 bits in key and back-pointer
constexpr int NBINS = 64;
enum { M_GREEDY = 0, M_BEAM = 1, M_EXACT = 2 };

struct Params {
    const DevProblem *probs;
    const uint32_t *plist;              // partitions to solve (indices into probs); nullptr = 0..count-1
    const uint32_t *pcount;             // number of entries (device counter) or nullptr -> count
    uint32_t count;
    uint32_t *work;                     // work counter (partitions are handed out one warp at a time)
    int32_t *best_score; uint32_t *best_rows; int32_t *incumbent;
    unsigned long long *cand, *node, *funnel;
    uint32_t *fb_list, *fb_count;       // partitions whose exact pass did not fit C records -> next tier
    uint8_t *active;                    // != nullptr: flag the partition for the global-memory path instead
    uint32_t *arena;                    // back-pointers [warp][level][C]
    unsigned long long *stamps;         // [prob][4] %globaltimer: picked up / final incumbent found / proven / spare
    int C, tsize;                       // records per level, hash slots (power of two, >= 2C)
    int beam_width; float beam_growth; int beam_delta, beam_shift;
};

// per-warp staging (fixed part); the record arrays follow it (see Arr)
struct Fixed {
    int32_t n, a, mode, m, mc, limit, lmax, g0, overflow, width, beam_shift, beam_delta;
    uint32_t tmask_row; int32_t max_mc;            // widest level of the exact pass (statistics)
    uint32_t cj[MAXN_SMALL];            // 2c_j | 2J_j << 16 of the partition
    uint32_t okP[32], parP[32], okL1[32], A2m[32], par1[32];
    int32_t  score[32], slack[32], lslack[32], d0[32], f0[32], ptgs[32], frec[32];
    uint32_t nl0[32], cstart[33], rank[32];
    uint8_t  ovt[32][MAXN_SMALL + 4];   // prefilter table of decode_light (decode_common.cuh fill_ovt)
    uint8_t  posP[32][MAXN_SMALL], posE[32][MAXN_SMALL];    // positions of the set bits of okP / E, lowest first
    uint16_t L0[32][L0CAP];                                 // class-0 loop multisets: low | high << 8 (0xff: none)
    uint32_t qh[64], qw_cfg[64];
    int32_t  qw_err[64], qw_tot[64];
    unsigned long long best;            // greedy: min (tot << 32 | code)
    unsigned long long ctr_cfg, ctr_heavy, ctr_child, ctr_batches, ctr_cand, ctr_node;
};

struct Arr {                            // views into the warp's shared memory
    uint4 *st;                          // [2][C] column state
    int32_t *sc, *f;                    // [2][C]
    uint16_t *rank;                     // [2][C] rank of the record's prefix among the level's records
    uint32_t *key;                      // [C] children: rank(parent) << 20 | row code
    uint16_t *pidx;                     // [C] children: index of the parent in the frontier
    uint32_t *table;                    // [tsize] record index or T_EMPTY
};

__host__ __device__ inline size_t warp_bytes(int C, int tsize)
{
    size_t s = (sizeof(Fixed) + 15) & ~(size_t)15;
    s += (size_t)C * (2 * 16 + 2 * 4 + 2 * 4 + 2 * 2 + 4 + 2) + (size_t)tsize * 4;
    return (s + 15) & ~(size_t)15;
}

__device__ __forceinline__ Arr carve(uint8_t *base, int C, int tsize)
{
    Arr R;
    uint8_t *p = base + ((sizeof(Fixed) + 15) & ~(size_t)15);
    R.st = reinterpret_cast<uint4 *>(p); p += (size_t)C * 32;
    R.sc = reinterpret_cast<int32_t *>(p); p += (size_t)C * 8;
    R.f = reinterpret_cast<int32_t *>(p); p += (size_t)C * 8;
    R.key = reinterpret_cast<uint32_t *>(p); p += (size_t)C * 4;
    R.table = reinterpret_cast<uint32_t *>(p); p += (size_t)tsize * 4;
    R.rank = reinterpret_cast<uint16_t *>(p); p += (size_t)C * 4;
    R.pidx = reinterpret_cast<uint16_t *>(p);
    return R;
}

__device__ __forceinline__ uint32_t mix32s(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t colbyte(const uint4 &s, int j)
{
    const uint32_t w = (j >> 2) == 0 ? s.x : ((j >> 2) == 1 ? s.y : ((j >> 2) == 2 ? s.z : s.w));
    return (w >> (8 * (j & 3))) & 0xffu;
}
// search-order code of a row configuration (ambi_common.cuh): ip << 10 | il
__device__ __forceinline__ uint32_t cfg_code(int n, int p1, int p2, int l1, int l2)
{
    const int ip = p1 < 0 ? 0 : ambi_ip_code(n - 1 - p1, p2 < 0 ? -1 : n - 1 - p2);
    const int il = l1 < 0 ? 0 : ambi_il_code(n - 1 - l1, l2 < 0 ? -1 : n - 1 - l2);
    return ((uint32_t)ip << 10) | (uint32_t)il;
}

// Admissible bound on segments j > a of the state `s` after adding a row with pattern ends pm and loop ends
// lm1 (lm2: the column that takes two) -- bit masks by column -- with `rows_left0` = rows that can still add to
// column n-1 (n-1-a for a child of row a; n-a for the record itself, whose row a is still open).
// Mirrors child_bound() of stage_a.cu and future_bound() of the oracle.  All lanes walk the same columns.
__device__ __forceinline__ int bound_cols(const Fixed &S, const uint4 &s, int a, int n, uint32_t pm, uint32_t lm1, uint32_t lm2,
                                          int rows_left0, int budget, bool active, bool early_exit)
{
    int lb = 0, cov = 0, capsum = 0;
    int rp_le = __popc(pm), rl_le = __popc(lm1) + __popc(lm2);
    int six = 6 * rows_left0;
#pragma unroll
    for (int j = MAXN_SMALL - 1; j >= 1; --j) {
        if (j < n && j > a) {                              // warp-uniform
            const uint32_t cb = colbyte(s, j);
            const uint32_t cjw = S.cj[j];
            const int c2 = (int)(cjw & 0xffffu), j2 = (int)(cjw >> 16);
            const int nP = cb & 3, nL = (cb >> 2) & 3, bP = (cb >> 4) & 3, bL = cb >> 6;
            const int pab = (pm >> j) & 1, lab = (int)((lm1 >> j) & 1) + (int)((lm2 >> j) & 1);
            const int rp = rp_le - pab, rl = rl_le - lab;              // ends strictly below column j
            rp_le = rp; rl_le = rl;
            const int kP = 2 - pab - max(rp, rl);                       // rows 2, 5a at (a,j)
            const int kL = 2 - max(lab + rl, pab + max(rl, rp));        // rows 4a, 4b, 5b
            const int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
            const int cap = nbP + 2 * nbL;
            const int wn = nP + 2 * nL + pab + 2 * lab;
            cov += wn; capsum += cap;
            const int fmax = min(six, capsum);
            six -= 6;
            const int d = 2 * cov - c2;
            lb += max(max(d, -d - 2 * fmax), 0);
            const int gmax = cap + min(capsum, 6);
            const int f = wn - j2;
            lb += max(max(f, -f - gmax), 0);
            if (early_exit && (j & 3) == 0 && __ballot_sync(0xffffffffu, active && lb <= budget) == 0) break;
        }
    }
    return lb;
}

// ---- phase 1: lane = frontier record g0 + lane; masks, candidate counts, position lists -----------------
__device__ __noinline__ uint32_t phase1(Fixed &S, const Arr &R, int lane)
{
    const int a = S.a, n = S.n;
    const int idx = S.g0 + lane;
    const bool have = idx < S.m;
    uint4 s = make_uint4(0, 0, 0, 0);
    int score = 0, frec = 0;
    if (have) { s = R.st[idx]; score = R.sc[idx]; frec = R.f[idx]; }
    const int ptgs = (have && a > 0) ? (int)(s.x & 0xffu) : 0;
    int slack = have ? S.limit - score : -1;
    // heuristic band (the beam only feeds the incumbent): a child may exceed its parent's f by beam_delta, or by a
    // fraction of what is left up to the current limit if that is more
    if (S.mode == M_BEAM && have) slack = min(slack, frec - score + max(S.beam_delta, (S.limit - frec) >> S.beam_shift));
    uint32_t okP = 0, parP = 0, okL1 = 0, okL2 = 0, par1 = 0, par2 = 0;
    int covC = 0, fbC = 0;
    const int lmax = S.lmax;
#pragma unroll
    for (int j = 0; j < MAXN_SMALL; ++j) {
        if (j >= a && j < n) {
            const uint32_t b = colbyte(s, j);
            const int nP = b & 3, nL = (b >> 2) & 3, bP = (b >> 4) & 3, bL = (b >> 6) & 3;
            const bool root = (a == 0 && j == n - 1);
            covC += nP + 2 * nL;
            if (j == a) fbC = nP + 2 * nL;
            okP |= (uint32_t)(bP >= 1) << j;
            parP |= (uint32_t)(nP >= 1 || root) << j;
            okL1 |= (uint32_t)(bL >= 1 && lmax >= 1) << j;
            okL2 |= (uint32_t)(bL >= 2 && lmax >= 2) << j;
            par1 |= (uint32_t)(nP + nL >= 1 || root) << j;
            par2 |= (uint32_t)(nP + nL >= 2 || root) << j;
        }
    }
    if (!have) { okP = parP = okL1 = okL2 = par1 = par2 = 0; }
    // the record's own bound on segments > a: every child's bound on those segments is at least this, so
    // err + pf > slack cuts a configuration before the O(n) bound is evaluated for it
    const int pf = bound_cols(S, s, a, n, 0u, 0u, 0u, n - a, 0, false, false);
    uint32_t over = 0, fover = 0;
#ifndef AMBI_NO_OVER
    {
        int cov = 0;
#pragma unroll
        for (int j = MAXN_SMALL - 1; j >= 1; --j) {
            if (j < n && j > a) {
                const uint32_t cb = colbyte(s, j);
                const int wn = (int)(cb & 3) + 2 * (int)((cb >> 2) & 3);
                cov += wn;
                over |= (uint32_t)(2 * cov >= (int)(S.cj[j] & 0xffffu)) << j;
                fover |= (uint32_t)(wn >= (int)(S.cj[j] >> 16)) << j;
            }
        }
    }
#endif
    dec::fill_ovt(S, lane, over, fover, n);
    const uint32_t cja = S.cj[a];
    uint32_t total = 0, nl0 = 0;
    if (have && slack >= 0) total = dec::count_configs(okP, parP, okL1, okL2, par1, par2, &nl0);
    dec::write_lists(S, lane, okP, parP, okL1, okL2, par1, par2, nl0);
    S.score[lane] = score; S.slack[lane] = slack; S.lslack[lane] = slack - pf; S.frec[lane] = frec;
    S.d0[lane] = 2 * covC - (int)(cja & 0xffffu); S.f0[lane] = fbC - (int)(cja >> 16);
    S.ptgs[lane] = ptgs;
    S.rank[lane] = have ? (uint32_t)R.rank[idx] : 0u;
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    S.cstart[lane + 1] = incl;
    if (lane == 0) S.cstart[0] = 0;
    const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    return T;
}

// One batch of light-test survivors through the O(n) bound; survivors are queued for the merge (beam / exact) or
// folded into the arg-min (greedy).  Returns the new write-queue fill.
__device__ __noinline__ int heavy_batch(Fixed &S, const Arr &R, int count, int qw, int lane)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int a = S.a, n = S.n;
    const bool act = lane < count;
    const uint32_t e = act ? S.qh[lane] : 0u;
    const int slot = (e >> 24) & 31;
    const uint32_t cfg = e & 0xffffffu;
    int p1, p2, l1, l2; unpack_cfg(cfg, p1, p2, l1, l2);
    const uint4 s = R.st[min(S.g0 + slot, S.m - 1)];
    const int slack = S.slack[slot];
    const int w = (p1 >= 0) + (p2 >= 0) + 2 * ((l1 >= 0) + (l2 >= 0));
    const int err = act ? abs(S.d0[slot] + 2 * w) + abs(S.f0[slot] + w) : 0;
    const uint32_t pm = (p1 >= 0 ? 1u << p1 : 0u) | (p2 >= 0 ? 1u << p2 : 0u);
    const uint32_t lm1 = (l1 >= 0 ? 1u << l1 : 0u) | (l2 >= 0 ? 1u << l2 : 0u);
    const uint32_t lm2 = (l1 >= 0 && l1 == l2) ? 1u << l1 : 0u;
    const int lb = bound_cols(S, s, a, n, pm, lm1, lm2, n - 1 - a, slack - err, act, true);
    const int tot = err + lb;
    const bool keep = act && tot <= slack;
    if (S.mode == M_GREEDY) {
        unsigned long long k = keep ? (((unsigned long long)(uint32_t)tot << 32) | ((unsigned long long)cfg_code(n, p1, p2, l1, l2) << 8) | (unsigned)slot) : ~0ull;
        // one record per level in the greedy dive: slot is always 0; min over lanes
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, k, o); k = v < k ? v : k; }
        if (lane == 0 && k < S.best) S.best = k;
        return qw;
    }
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    if (keep) { const int pos = qw + __popc(km & lt_mask); S.qw_cfg[pos] = e; S.qw_err[pos] = err; S.qw_tot[pos] = tot; }
    return qw + __popc(km);
}

// child state of record state `s` after row a takes configuration (p1,p2,l1,l2); th = third-generation count
__device__ __forceinline__ uint4 child_state(const uint4 &s, int a, int n, int p1, int p2, int l1, int l2, int th)
{
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 1; j < MAXN_SMALL; ++j) {
        if (j > a && j < n) {
            const uint32_t cb = colbyte(s, j);
            const int nP = cb & 3, nL = (cb >> 2) & 3, bP = (cb >> 4) & 3, bL = (cb >> 6) & 3;
            const int pab = (j == p1) + (j == p2), lab = (j == l1) + (j == l2);
            const int rp = (p1 >= 0 && p1 < j) + (p2 >= 0 && p2 < j);
            const int rl = (l1 >= 0 && l1 < j) + (l2 >= 0 && l2 < j);
            const int kP = 2 - pab - max(rp, rl);
            const int kL = 2 - max(lab + rl, max(pab + rl, pab + rp));
            const int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
            w[j >> 2] |= (uint32_t)(((nP + pab) | ((nL + lab) << 2) | (nbP << 4) | (nbL << 6)) & 0xff) << (8 * (j & 3));
        }
    }
    w[0] |= (uint32_t)th & 0xffu;          // column 0 is closed for every child: its byte carries the count
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ bool st_eq(const uint4 &x, const uint4 &y) { return x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w; }

// Merge `count` queued children into the level's table of distinct states (children live in the other half of
// the record arrays).  Lanes whose states hash alike go one per round, so that a state is only ever inserted or
// replaced by one lane at a time.
__device__ __noinline__ void merge_batch(Fixed &S, const Arr &R, int count, int C, uint32_t tmask, int lane)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int a = S.a, n = S.n;
    const bool act = lane < count;
    uint32_t cfg = 0; int err = 0, tot = 0, slot = 0;
    if (act) { const uint32_t e = S.qw_cfg[lane]; slot = (e >> 24) & 31; cfg = e & 0xffffffu; err = S.qw_err[lane]; tot = S.qw_tot[lane]; }
    int p1, p2, l1, l2; unpack_cfg(cfg, p1, p2, l1, l2);
    const int pidx = min(S.g0 + slot, S.m - 1);
    const uint4 ps = R.st[pidx];
    const uint32_t tm = S.tmask_row;
    int hits = 0;
    if (tm) hits = (p1 >= 0 && ((tm >> p1) & 1)) + (p2 >= 0 && ((tm >> p2) & 1)) + (l1 >= 0 && ((tm >> l1) & 1)) + (l2 >= 0 && ((tm >> l2) & 1));
    // completed assignments all share one (empty) state: drop the running third-generation count
    const uint4 cs = child_state(ps, a, n, p1, p2, l1, l2, (n == a + 1) ? 0 : S.ptgs[slot] + hits);
    const int sc = S.score[slot] + err, ff = S.score[slot] + tot;
    const uint32_t key = (S.rank[slot] << 20) | cfg_code(n, p1, p2, l1, l2);
    const uint32_t h = mix32s(cs.x ^ mix32s(cs.y + 0x9e3779b9u) ^ mix32s(cs.z ^ 0x85ebca6bu) ^ (cs.w * 0xc2b2ae35u));
    uint4 *stB = R.st + C; int32_t *scB = R.sc + C, *fB = R.f + C;      // children half (the caller keeps cur == 0 layout, see swap())
    bool pending = act;
    for (;;) {
        if (__ballot_sync(0xffffffffu, pending) == 0) break;
        const uint32_t grp = __match_any_sync(0xffffffffu, pending ? (unsigned long long)h : ((1ull << 32) | (unsigned)lane));
        const bool leader = pending && (__ffs((int)grp) - 1 == lane);
        int found = -1; uint32_t ts = h & tmask;
        if (leader) {
            for (;;) {
                const uint32_t cur = R.table[ts];
                if (cur == T_EMPTY) break;
                if (st_eq(stB[cur], cs)) { found = (int)cur; break; }
                ts = (ts + 1) & tmask;
            }
        }
        const bool isnew = leader && found < 0;
        const uint32_t nm = __ballot_sync(0xffffffffu, isnew);
        const int base = S.mc;
        if (isnew) {
            const int idx = base + __popc(nm & lt_mask);
            if (idx < C) {
                stB[idx] = cs; scB[idx] = sc; fB[idx] = ff; R.key[idx] = key; R.pidx[idx] = (uint16_t)pidx;
                while (atomicCAS(&R.table[ts], T_EMPTY, (uint32_t)idx) != T_EMPTY) ts = (ts + 1) & tmask;
            }
        } else if (leader) {
            // same state: keep the smaller (score, prefix in search order)
            const int os = scB[found];
            if (sc < os || (sc == os && key < R.key[found])) { scB[found] = sc; fB[found] = ff; R.key[found] = key; R.pidx[found] = (uint16_t)pidx; }
        }
        __syncwarp();
        if (lane == 0) {
            const int want = base + __popc(nm);
            if (want > C) { S.overflow = 1; S.mc = C; } else S.mc = want;
        }
        pending = pending && !leader;
        __syncwarp();
    }
}

// ---- one level: expand the frontier (records [0, m) of half 0) into distinct children (half 1) ---------------
__device__ __noinline__ void expand_level(Fixed &S, const Arr &R, int C, uint32_t tmask, int lane)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int m = S.m;
    if (S.mode != M_GREEDY) for (int i = lane; i <= (int)tmask; i += 32) R.table[i] = T_EMPTY;
    if (lane == 0) { S.mc = 0; S.best = ~0ull; }
    __syncwarp();
    unsigned long long n_cfg = 0, n_heavy = 0, n_child = 0, n_batches = 0, n_infeas = 0;
    for (int g0 = 0; g0 < m; g0 += 32) {
        if (lane == 0) S.g0 = g0;
        __syncwarp();
        const uint32_t T = phase1(S, R, lane);
        n_cfg += T;
        // ---- phase 2: lane = candidate, flattened over the group's records; three stages fed through two queues
        int qh = 0, qw = 0;
        uint32_t cb0 = 0, s0 = 0;          // s0: record that owns candidate cb0
        for (;;) {
            if (qw >= 32 || (cb0 >= T && qh == 0 && qw > 0)) {
                const int cnt = min(qw, 32);
                n_child += cnt;
                merge_batch(S, R, cnt, C, tmask, lane);
                qw -= cnt;
                uint32_t m0 = 0; int m1 = 0, m2 = 0;
                if (lane < qw) { m0 = S.qw_cfg[32 + lane]; m1 = S.qw_err[32 + lane]; m2 = S.qw_tot[32 + lane]; }
                __syncwarp();
                if (lane < qw) { S.qw_cfg[lane] = m0; S.qw_err[lane] = m1; S.qw_tot[lane] = m2; }
                __syncwarp();
                continue;
            }
            if (qh >= 32 || (cb0 >= T && qh > 0)) {
                const int cnt = min(qh, 32);
                n_heavy += cnt; ++n_batches;
                qw = heavy_batch(S, R, cnt, qw, lane);
                __syncwarp();
                qh -= cnt;
                const uint32_t mv = lane < qh ? S.qh[32 + lane] : 0u;
                __syncwarp();
                if (lane < qh) S.qh[lane] = mv;
                __syncwarp();
                continue;
            }
            if (cb0 >= T) break;
            // owner of candidate cb0 + lane by a run-length walk: bit (e - cb0 - 1) marks the record that ends at e;
            // every record has at least the empty configuration, so ends are distinct
            const uint32_t e = S.cstart[min(s0 + 1 + (uint32_t)lane, 32u)];
            const uint32_t d = e - cb0 - 1u;
            const uint32_t bm = __reduce_or_sync(0xffffffffu, (s0 + 1 + (uint32_t)lane <= 32u && d < 32u && e > cb0) ? 1u << d : 0u);
            const uint32_t c = cb0 + lane;
            bool pass = false;
            uint32_t entry = 0;
            if (c < T) {
                const int slot = (int)(s0 + __popc(bm & lt_mask));
                entry = dec::decode_light(S, slot, c - S.cstart[slot], S.tmask_row, 0, &pass);
            }
            n_infeas += __popc(__ballot_sync(0xffffffffu, c < T && !entry));
            s0 += __popc(bm);
            cb0 += 32;
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            if (pass) S.qh[qh + __popc(bal & lt_mask)] = entry;
            qh += __popc(bal);
            __syncwarp();
        }
        if (S.overflow && S.mode == M_EXACT) break;
    }
    if (lane == 0) {
        if (S.mode == M_EXACT && S.mc > S.max_mc) S.max_mc = S.mc;
        S.ctr_cfg += n_cfg; S.ctr_heavy += n_heavy; S.ctr_child += n_child; S.ctr_batches += n_batches;
        S.ctr_cand += n_cfg - n_infeas; S.ctr_node += (unsigned long long)m;
    }
    __syncwarp();
}

// children (half 1) become the frontier (half 0)
__device__ __forceinline__ void swap_halves(Fixed &S, const Arr &R, int C, int lane, bool with_rank)
{
    const int mc = S.mc;
    for (int i = lane; i < mc; i += 32) {
        R.st[i] = R.st[C + i]; R.sc[i] = R.sc[C + i]; R.f[i] = R.f[C + i];
        if (with_rank) R.rank[i] = R.rank[C + i];
    }
    __syncwarp();
    if (lane == 0) S.m = mc;
    __syncwarp();
}

// beam: keep the `width` best children by f (bins of width 1 from the level's minimum, the rest in the last bin;
// inside the cut bin the first come); compaction in place
__device__ __noinline__ void beam_select(Fixed &S, const Arr &R, int C, int lane)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int mc = S.mc, width = S.width;
    if (mc <= width) return;
    int32_t *fB = R.f + C;
    int fmin = 0x7fffffff;
    for (int i = lane; i < mc; i += 32) fmin = min(fmin, fB[i]);
    fmin = __reduce_min_sync(0xffffffffu, fmin);
    uint32_t *hist = R.table;                      // the merge table is free between levels
    for (int i = lane; i < NBINS; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < mc; i += 32) atomicAdd(&hist[max(0, min(fB[i] - fmin, NBINS - 1))], 1u);
    __syncwarp();
    int cut = NBINS; uint32_t quota = 0;
    {
        uint32_t below_ = 0;
        for (int k = 0; k < NBINS; ++k) { const uint32_t c = hist[k]; if (below_ + c > (uint32_t)width) { cut = k; quota = (uint32_t)width - below_; break; } below_ += c; }
    }
    __syncwarp();
    // stable in-place compaction, 32 records at a time (destination never passes the source)
    int out = 0; uint32_t taken = 0;
    uint4 *stB = R.st + C; int32_t *scB = R.sc + C;
    for (int i0 = 0; i0 < mc; i0 += 32) {
        const int i = i0 + lane;
        const bool have = i < mc;
        uint4 s = make_uint4(0, 0, 0, 0); int sc = 0, ff = 0;
        if (have) { s = stB[i]; sc = scB[i]; ff = fB[i]; }
        const int bin = have ? max(0, min(ff - fmin, NBINS - 1)) : NBINS;
        const uint32_t incut = __ballot_sync(0xffffffffu, have && bin == cut);
        const bool keep = have && (bin < cut || (bin == cut && taken + __popc(incut & lt_mask) < quota));
        taken += __popc(incut);
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) { const int o = out + __popc(km & lt_mask); stB[o] = s; scB[o] = sc; fB[o] = ff; }
        out += __popc(km);
        __syncwarp();
    }
    if (lane == 0) S.mc = out;
    __syncwarp();
}

// exact: dense ranks of the children in prefix order (= key order) and one back-pointer word per child
__device__ __noinline__ void rank_and_link(Fixed &S, const Arr &R, int C, uint32_t *arena_level, int lane)
{
    const int mc = S.mc;
    for (int i0 = 0; i0 < mc; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t k = i < mc ? R.key[i] : 0u;
        uint32_t r = 0;
        for (int j = 0; j < mc; ++j) r += R.key[j] < k;
        if (i < mc) {
            R.rank[C + i] = (uint16_t)r;
            arena_level[i] = (uint32_t)R.pidx[i] | ((k & 0xfffffu) << 12);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void put_root(Fixed &S, const Arr &R, const DevProblem &P, int lane)
{
    if (lane == 0) {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int b = 0; b < P.n; ++b) w[b >> 2] |= (uint32_t)((AMBI_BUD_INF << 4) | (AMBI_BUD_INF << 6)) << (8 * (b & 3));
        R.st[0] = make_uint4(w[0], w[1], w[2], w[3]); R.sc[0] = 0; R.f[0] = P.root_lb; R.rank[0] = 0;
        S.m = 1; S.overflow = 0;
    }
    __syncwarp();
}

#ifdef AMBI_EMU
#define AMBI_GLOBALTIMER() emu_globaltimer()
#else
#define AMBI_GLOBALTIMER() ::globaltimer_ns()
#endif

// One pass over the rows (greedy dive, beam, or exact within S.limit).  Returns the score of the completed
// assignment the pass ends in (0x7fffffff: none within the limit); *overflow = the exact pass did not fit C records.
__device__ __noinline__ int run_pass(const Params &P, Fixed &S, const Arr &R, const DevProblem &D, int mode, uint32_t *arena, int lane, int *overflow)
{
    const int n = D.n, C = P.C;
    const uint32_t tmask = (uint32_t)P.tsize - 1u;
    put_root(S, R, D, lane);
    if (lane == 0) { S.mode = mode; S.beam_shift = P.beam_shift; S.beam_delta = max(0, P.beam_delta); }
    __syncwarp();
    *overflow = 0;
    int leaf_score = 0x7fffffff;
    for (int a = 0; a < n; ++a) {
        if (lane == 0) { S.a = a; S.tmask_row = D.has_tgs ? D.tgs_mask[a] : 0u; }
        __syncwarp();
        expand_level(S, R, C, tmask, lane);
        if (mode == M_GREEDY) {
            // the record's arg-min child (tot, then search order) becomes the next record
            const unsigned long long bk = S.best;
            if (bk == ~0ull) break;
            if (lane == 0) {
                const uint32_t code = (uint32_t)(bk >> 8) & 0xfffffu;
                const int tot = (int)(bk >> 32);
                int p1 = -1, p2 = -1, l1 = -1, l2 = -1;
                const int ip = code >> 10, il = code & 1023;
                if (ip) { int q1 = 0; while (ambi_ip_code(q1 + 1, -1) <= ip) ++q1; const int off = ip - ambi_ip_code(q1, -1);
                          p1 = n - 1 - q1; if (off > 0) p2 = n - 1 - (off - 1); }
                if (il) { int q1 = 0; while (ambi_il_code(q1 + 1, -1) <= il) ++q1; const int off = il - ambi_il_code(q1, -1);
                          l1 = n - 1 - q1; if (off > 0) l2 = (off - 1 == q1) ? l1 : n - 1 - (off - 1); }
                const int w = (p1 >= 0) + (p2 >= 0) + 2 * ((l1 >= 0) + (l2 >= 0));
                const int err = abs(S.d0[0] + 2 * w) + abs(S.f0[0] + w);
                const uint32_t tm = S.tmask_row;
                int hits = 0;
                if (tm) hits = (p1 >= 0 && ((tm >> p1) & 1)) + (p2 >= 0 && ((tm >> p2) & 1)) + (l1 >= 0 && ((tm >> l1) & 1)) + (l2 >= 0 && ((tm >> l2) & 1));
                const uint4 ps = R.st[0];
                const int sc0 = R.sc[0];
                R.st[0] = child_state(ps, a, n, p1, p2, l1, l2, (n == a + 1) ? 0 : S.ptgs[0] + hits);
                R.sc[0] = sc0 + err; R.f[0] = sc0 + tot;
            }
            __syncwarp();
            if (a == n - 1) leaf_score = R.sc[0];
            continue;
        }
        if (mode == M_EXACT && S.overflow) { *overflow = 1; break; }
        if (S.mc == 0) break;
        if (a == n - 1) {
            leaf_score = R.sc[C];
            if (mode == M_EXACT && lane == 0) arena[(size_t)a * C] = (uint32_t)R.pidx[0] | ((R.key[0] & 0xfffffu) << 12);
            __syncwarp();
            break;
        }
        if (mode == M_BEAM) { beam_select(S, R, C, lane); swap_halves(S, R, C, lane, false); }
        else { rank_and_link(S, R, C, arena + (size_t)a * C, lane); swap_halves(S, R, C, lane, true); }
    }
    return leaf_score;
}

// The whole search of one partition.  Returns 0 when solved, 1 when the exact pass did not fit C records.
__device__ __noinline__ int solve_one(const Params &P, Fixed &S, const Arr &R, uint32_t prob, uint32_t *arena, int lane)
{
    const DevProblem &D = P.probs[prob];
    const int n = D.n, C = P.C;
    unsigned long long t_start = 0, t_best = 0;
    if (lane == 0) {
        t_start = AMBI_GLOBALTIMER();
        S.n = n; S.lmax = D.lmax; S.limit = D.zero_score; S.max_mc = 0;
        S.ctr_cfg = S.ctr_heavy = S.ctr_child = S.ctr_batches = S.ctr_cand = S.ctr_node = 0;
        S.width = max(AMBI_BEAM_FLOOR, (int)min(8192.0f, (float)P.beam_width * exp2f(P.beam_growth * (float)max(0, n - 12))));
        S.width = min(S.width, C);
    }
    if (lane < MAXN_SMALL) S.cj[lane] = lane < n ? D.cj[lane] : 0u;
    __syncwarp();
    // ---- incumbent: greedy dive, then the beam inside that limit (any completed assignment is a valid upper bound)
    int incumbent = D.zero_score, ov = 0;
    for (int mode = M_GREEDY; mode <= M_BEAM; ++mode) {
        if (mode == M_BEAM && P.beam_width <= 1) continue;
        const int leaf = run_pass(P, S, R, D, mode, arena, lane, &ov);
        if (leaf < incumbent) { incumbent = leaf; if (lane == 0) { S.limit = leaf; t_best = AMBI_GLOBALTIMER(); } }
        __syncwarp();
    }
    // ---- exact pass within the incumbent.  If its widest level does not fit C records the limit is lowered:
    // a pass within L that ends in a completed assignment has seen every assignment of score <= L, so what it
    // returns is the optimum (and the tie-break winner) whatever L was; a pass that ends in none proves
    // optimum > L.  Bisection between the root bound and the smallest limit known not to fit.
    int result = 1, best = 0x7fffffff;
    int lo = min(D.root_lb, incumbent), over_at = 0x7fffffff, L = incumbent;
    for (int it = 0; it < 12; ++it) {
        if (lane == 0) S.limit = L;
        __syncwarp();
        const int leaf = run_pass(P, S, R, D, M_EXACT, arena, lane, &ov);
        if (ov) { over_at = L; }
        else if (leaf != 0x7fffffff) { best = leaf; result = 0; break; }
        else lo = L + 1;
        if (lo > over_at - 1 || lo > incumbent) break;
        L = lo + (min(over_at - 1, incumbent) - lo) / 2;
    }
    if (result == 0 && lane == 0) {
        // walk the back-pointers from the single completed record
        if (best < incumbent) t_best = AMBI_GLOBALTIMER();
        uint32_t *rows = P.best_rows + (size_t)prob * AMBI_MAXN;
        uint32_t idx = 0;
        for (int lvl = n - 1; lvl >= 0; --lvl) {
            const uint32_t v = *reinterpret_cast<volatile uint32_t *>(&arena[(size_t)lvl * C + idx]);
            rows[lvl] = v >> 12; idx = v & 0xfffu;
        }
        P.best_score[prob] = best;
    }
    if (lane == 0) {
        if (P.incumbent) P.incumbent[prob] = incumbent;
        atomicAdd(&P.cand[prob], S.ctr_cand); atomicAdd(&P.node[prob], S.ctr_node);
        if (P.funnel) { atomicAdd(&P.funnel[0], S.ctr_cfg); atomicAdd(&P.funnel[1], S.ctr_heavy); atomicAdd(&P.funnel[2], S.ctr_child); atomicAdd(&P.funnel[3], S.ctr_batches); }
        if (P.stamps) {
            unsigned long long *st = P.stamps + (size_t)prob * 4;
            st[0] = t_start; st[1] = t_best ? t_best : t_start; st[2] = AMBI_GLOBALTIMER(); st[3] = (unsigned long long)result | ((unsigned long long)S.max_mc << 8);
        }
        if (result) {
            if (P.active) P.active[prob] = 1;
            if (P.fb_count) { const uint32_t s = atomicAdd(P.fb_count, 1u); if (P.fb_list) P.fb_list[s] = prob; }
        }
    }
    __syncwarp();
    return result;
}

}  // namespace small
}  // namespace ambi
