// stage_a.cu -- Stage A of the BFB fold-pattern search on sm_100a.
//
// Replaces BFB_ILP + external cbc (reference src/LocalGenomicMap.cpp:4397-4752,
// localhap.cpp:173-212).  The variable grid P[a,b], L[a,b] is fixed row by row
// (a = start segment); after row a both deviation terms of segment a are exact
// and every constraint of the reference model reduces to per-column counters
// (see DESIGN.md "Search formulation").
//
// Kernels (all INT32 + shared memory; no tensor cores -- nothing here is a
// dense contraction):
//   k_expand       K1+K2  one warp per group of up to 32 frontier records: enumerate the
//                         feasible row configurations, score them (exact deviation of
//                         segment a + admissible bound for the rest), write the survivors
//                         as key records + links + flags
//   k_dedup        K3a    hash-table merge of key records that reach the same column
//                         state: keep min (score, prefix in search order)
//   k_finalize     K3b    per-problem arg-min of completed assignments
//   k_beam_*       K3c    per-problem top-W cut by f-histogram (incumbent pass)
//   k_materialise  K3d    full records (key + the parent's rows + the row code) for the
//                         survivors only: the next level's dense frontier
//   k_unpack              compact host problems -> DevProblem
#include "ambi_common.cuh"
#ifndef AMBI_EMU            // tests/emu/expand_emu.cpp compiles the device bodies of this file with g++ (test infrastructure)
#include "stage_a.h"
#endif
#include "decode_common.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstddef>
#include <vector>
#include <algorithm>

namespace ambi {

constexpr int WPB = 8;              // warps per block in k_expand
constexpr int GRP = 32;             // frontier records handled together by one warp
constexpr int BEAM_BINS = 64;
#ifndef AMBI_CHUNK
#define AMBI_CHUNK 4
#endif
constexpr uint32_t CHUNK = AMBI_CHUNK;        // consecutive configurations of a record stepped through by one lane (k_expand phase 2)

enum { EMIT_ALL = 0, EMIT_BEST = 1 };

struct LevelArgs {
    const DevProblem *probs;
    const uint8_t *parents;
    const uint32_t *plist;
    const uint32_t *pcount;
    uint8_t *children;              // key records of the children: header + column state (kstride bytes each)
    uint2 *links;                   // per child: (parent's index in `parents`, row code) -- the prefix is not copied
    uint8_t *dead;                  // per child: 0 open, 2 completed assignment (set here), 1 once the merge dropped it -- the key records stay read-only
    int kstride;
    uint32_t *ccount;
    uint32_t cap;
    const int32_t *limit;
    unsigned long long *cand_count;
    unsigned long long *node_count;
    unsigned long long *funnel;      // [4] configurations enumerated / bounds evaluated / children written / bound batches (whole run)
    uint32_t *work_counter;
    uint32_t *overflow;
    int level;
    int stride;
    int nmax;
    int emit;                       // EMIT_ALL: every child within the limit; EMIT_BEST: arg-min child per record
    int beam_shift;
    int beam_delta_short, beam_shift_short, beam_short_n;   // the band of partitions of up to beam_short_n segments
    int beam_delta;                 // >= 0 (incumbent beam only): also drop children whose f exceeds the parent's by more
    int shard_world, shard_rank, shard_level;
    int gpw;                        // target number of record groups per resident warp
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t below(int b) { return (1u << b) - 1u; }

// Per-warp staging of a group of up to GRP frontier records (only records that have at least one
// configuration take a slot).  Column arrays are laid out [column][slot] so that lanes working on different
// records hit different banks and lanes working on the same record broadcast.  NMAX = 16 or 32 columns: a batch
// whose longest partition has at most 16 segments runs the narrow instantiation, whose staging is small enough
// for 4 blocks of 8 warps per SM (the 2c|2J words of the partitions are read through the read-only cache
// instead of being copied per slot).
template <int NMAX>
struct GroupShared {
    uint32_t colw[NMAX / 4][GRP];       // column bytes nP | nL<<2 | bP<<4 | bL<<6 of the record in the slot, 4 columns per word
    __device__ __forceinline__ uint32_t colb(int j, int slot) const { return reinterpret_cast<const uint8_t *>(&colw[j >> 2][slot])[j & 3]; }
    uint8_t  posP[GRP][NMAX];           // enumeration lists (decode_common.cuh, "Incremental enumeration")
    uint16_t LE[GRP][dec::L0CAP + NMAX];
    uint32_t okP[GRP], parP[GRP];       // column may take a pattern end / has a pattern parent (row 1)
    uint32_t okL1[GRP], A2m[GRP];       // column may take one loop end / two loop ends and has two parents
    uint32_t A1m[GRP];                  // column may take one loop end and already holds a parent for it (row 3)
    uint32_t tmask[GRP];                // third-generation cells of this row
    uint8_t  ovt[GRP][NMAX + 4];        // prefilter table of decode_light (decode_common.cuh fill_ovt)
    int32_t  score[GRP], slack[GRP], prob[GRP];
    int32_t  d0[GRP], f0[GRP];          // 2 cov_a - 2c_a and fb_a - 2J_a before the row's own entries
    int32_t  lslack[GRP];               // slack minus the record's own bound on segments > a (light-test limit)
    uint32_t pidx[GRP];
    uint16_t nl0[GRP];                  // loop multisets whose parents already exist (class 0)
    uint32_t ncfg[GRP];                 // configurations of the record (up to ~2^18 for 32 open columns: not 16 bits)
    uint32_t cstart[GRP + 1];           // exclusive prefix of the records' chunk counts (work items of phase 2)
    uint8_t  n[GRP];
    uint8_t  ptgs[GRP];                 // running third-generation count (at most 5)
    uint8_t  ush[GRP];                  // copy-number unit of the record's partition (SCALED instantiations only)
    uint32_t cnt[GRP];                  // structurally infeasible configurations per record (not scored)
    uint32_t qh[64];                    // light-test survivors awaiting the bound: slot<<24 | cfg
    uint32_t qw_cfg[64]; int32_t qw_err[64];       // children awaiting the write ...
    alignas(8) int32_t qw_tot[64];                 // ... (EMIT_BEST never queues children: the same bytes hold best[])
    // EMIT_BEST: per record min (tot << 32 | cfg-order key)
    __device__ __forceinline__ unsigned long long *best() { return reinterpret_cast<unsigned long long *>(qw_tot); }
};

using dec::pack_cfg;
using dec::unpack_cfg;
// search-order code of a configuration (see ambi_common.cuh)
__device__ __forceinline__ uint32_t cfg_code(int n, int p1, int p2, int l1, int l2) {
    int ip = p1 < 0 ? 0 : ambi_ip_code(n - 1 - p1, p2 < 0 ? -1 : n - 1 - p2);
    int il = l1 < 0 ? 0 : ambi_il_code(n - 1 - l1, l2 < 0 ? -1 : n - 1 - l2);
    return ((uint32_t)ip << 10) | (uint32_t)il;
}

// Admissible bound on segments j > a for a child of record `slot` (ends p1<p2, l1<=l2; -1 absent).
// Mirrors oracle/ambi_oracle.c future_bound(): committed coverage / fold-back can only grow; what
// can still be added is capped by the column budgets and by 6 per remaining row.
// All lanes walk the same columns (jtop-1 .. a+1) so the shared-memory reads stay conflict free.
// SCALED: copy numbers in units of 1/(2U) copies, U = 1 << ush per partition: counts of copies become counts of units
// where they meet 2c / 2J (the half-copy instantiation is unchanged: ush is the constant 0 there).
template <int NMAX, bool SCALED>
__device__ __forceinline__ int child_bound(const GroupShared<NMAX> &S, const DevProblem *probs, int slot, int jtop, int a,
                                           int p1, int p2, int l1, int l2, int budget, bool active)
{
    const int ush = SCALED ? (int)S.ush[slot] : 0;
    int lb = 0, cov = 0, capsum = 0;
    // ends as bit masks aligned so that column jtop-1 sits in bit 31; the walk shifts them left
    // (pattern ends are distinct columns; a loop column may hold two ends: lm2 = the doubled one)
    const int sh = 32 - jtop;
    uint32_t pm = ((p1 >= 0 ? 1u << p1 : 0u) | (p2 >= 0 ? 1u << p2 : 0u)) << sh;
    uint32_t lm1 = ((l1 >= 0 ? 1u << l1 : 0u) | (l2 >= 0 ? 1u << l2 : 0u)) << sh;
    uint32_t lm2 = ((l1 >= 0 && l1 == l2) ? 1u << l1 : 0u) << sh;
    // ends at or below the current column (the walk goes right to left, every end is < jtop)
    int rp_le = __popc(pm), rl_le = __popc(lm1) + __popc(lm2);
    int six = 6 * (jtop - 1 - a);                                // 6 * (j - a)
    const uint32_t *cjp = probs[S.prob[slot]].cj;           // 2c_j | 2J_j << 16 of the record's partition, through the read-only cache
    const int nrec = S.n[slot];
#pragma unroll 2
    for (int j = jtop - 1; j > a; --j) {
        const uint32_t cb = S.colb(j, slot);
        const uint32_t cjw = j < nrec ? __ldg(cjp + j) : 0u;   // columns beyond the record's partition are empty
        const int c2 = (int)(cjw & 0xffffu), j2 = (int)(cjw >> 16);
        const int nP = cb & 3, nL = (cb >> 2) & 3, bP = (cb >> 4) & 3, bL = cb >> 6;
        const int pab = pm >> 31, lab = (lm1 >> 31) + (lm2 >> 31);
        pm <<= 1; lm1 <<= 1; lm2 <<= 1;
        const int rp = rp_le - pab, rl = rl_le - lab;          // ends strictly below column j
        rp_le = rp; rl_le = rl;
        const int kP = 2 - pab - max(rp, rl);                   // rows 2, 5a at (a,j)
        const int kL = 2 - max(lab + rl, pab + max(rl, rp));    // rows 4a, 4b, 5b
        const int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
        const int cap = nbP + 2 * nbL;                          // both already <= 2
        const int wn = nP + 2 * nL + pab + 2 * lab;
        cov += wn; capsum += cap;
        const int fmax = min(six, capsum);
        six -= 6;
        const int d = (cov << (ush + 1)) - c2;
        lb += max(max(d, -d - (fmax << (ush + 1))), 0);
        const int gmax = cap + min(capsum, 6);
        const int f = (wn << ush) - j2;
        lb += max(max(f, -f - (gmax << ush)), 0);
        if ((j & 3) == 0 && __ballot_sync(0xffffffffu, active && lb <= budget) == 0) break;
    }
    return lb;
}

// One batch of light-test survivors through the O(n) bound; survivors are queued for the write
// (EMIT_ALL) or folded into the per-record arg-min (EMIT_BEST).  Returns the new write-queue fill.
template <int NMAX, bool SCALED>
__device__ __noinline__ int heavy_batch(GroupShared<NMAX> &S, const LevelArgs &A, int count, int jtop, int qw)
{
    const int lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int a = A.level;
    const bool act = lane < count;
    const uint32_t e = act ? S.qh[lane] : 0u;
    const int slot = (e >> 24) & 31;
    const uint32_t cfg = e & 0xffffffu;
    int p1, p2, l1, l2; unpack_cfg(cfg, p1, p2, l1, l2);
    int err = 0, tot = 0; bool keep = false;
    const int slack = S.slack[slot];
    if (act) {
        const int w = (p1 >= 0) + (p2 >= 0) + 2 * ((l1 >= 0) + (l2 >= 0));
        const int ush = SCALED ? (int)S.ush[slot] : 0;
        err = abs(S.d0[slot] + (w << (ush + 1))) + abs(S.f0[slot] + (w << ush));
    }
    const int lb = child_bound<NMAX, SCALED>(S, A.probs, slot, jtop, a, p1, p2, l1, l2, slack - err, act);
    if (act) { tot = err + lb; keep = tot <= slack; }
    if (A.emit == EMIT_BEST) {
        if (keep) atomicMin(&S.best()[slot], ((unsigned long long)(uint32_t)tot << 32) | cfg_code(S.n[slot], p1, p2, l1, l2));
        return qw;
    }
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    if (keep) { const int pos = qw + __popc(km & lt_mask); S.qw_cfg[pos] = e; S.qw_err[pos] = err; S.qw_tot[pos] = tot; }
    return qw + __popc(km);
}

// Children are written by one thread each (`count` queued entries from the head of the queue).
template <int NMAX>
__device__ __noinline__ void write_batch(GroupShared<NMAX> &S, const LevelArgs &A, int count)
{
    const int lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int a = A.level;
    const bool act = lane < count;
    uint32_t cfg = 0; int err = 0, tot = 0, slot = 0;
    if (act) { const uint32_t e = S.qw_cfg[lane]; slot = (e >> 24) & 31; cfg = e & 0xffffffu; err = S.qw_err[lane]; tot = S.qw_tot[lane]; }
    int p1, p2, l1, l2; unpack_cfg(cfg, p1, p2, l1, l2);
    const int n = S.n[slot];
    bool wr = act;
    const uint32_t code = cfg_code(n, p1, p2, l1, l2);
    const uint8_t *prec = A.parents + (size_t)S.pidx[slot] * A.stride;
    if (A.shard_world > 1 && a == A.shard_level && wr) {
        const uint32_t *prow = reinterpret_cast<const uint32_t *>(prec + sizeof(NodeHdr));
        uint32_t h = 0;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const uint32_t v = k < a ? prow[k] : (k == a ? code : 0u);
            h ^= mix32(v + 0x9e3779b9u * (uint32_t)(k + 1));
        }
        wr = (int)(mix32(h) % (uint32_t)A.shard_world) == A.shard_rank;
    }
    const uint32_t wm = __ballot_sync(0xffffffffu, wr);
    uint32_t cbase = 0;
    if (lane == 0 && wm) cbase = atomicAdd(A.ccount, (uint32_t)__popc(wm));
    cbase = __shfl_sync(0xffffffffu, cbase, 0);
    if (!wr) return;
    const uint32_t cslot = cbase + __popc(wm & lt_mask);
    if (cslot >= A.cap) {
        // the incumbent beam just loses the record (it is a heuristic); the exact pass records the level it
        // overflowed at, every later launch of the pass returns at once and the host continues depth-first
        if (A.beam_delta < 0) atomicExch(A.overflow, (uint32_t)a + 1u);
        return;
    }
    // A child leaves as a KEY record -- header and column state, what the merge and the beam selection look at -- plus
    // a link word pair (parent, row code).  Four children in five lose the merge: their prefix is never copied;
    // k_materialise builds full records (key + the parent's rows + the code) for the survivors only.
    uint8_t *crec = A.children + (size_t)cslot * A.kstride;
    uint32_t *ccol = reinterpret_cast<uint32_t *>(crec + offsetof(NodeHdr, col));
#pragma unroll 1
    for (int q = 0; q < NMAX / 4; ++q) {
        uint32_t wq = 0;
        if (4 * q + 3 > a && 4 * q < n) {
            const uint32_t pw = S.colw[q][slot];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 4 * q + k;
                const uint32_t cb = (pw >> (8 * k)) & 0xffu;
                const int nP = cb & 3, nL = (cb >> 2) & 3, bP = (cb >> 4) & 3, bL = (cb >> 6) & 3;
                const int pab = (j == p1) + (j == p2), lab = (j == l1) + (j == l2);
                const int rp = (p1 >= 0 && p1 < j) + (p2 >= 0 && p2 < j);
                const int rl = (l1 >= 0 && l1 < j) + (l2 >= 0 && l2 < j);
                const int kP = 2 - pab - max(rp, rl);
                const int kL = 2 - max(lab + rl, max(pab + rl, pab + rp));
                const int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
                const uint32_t nb = (uint32_t)(((nP + pab) | ((nL + lab) << 2) | (nbP << 4) | (nbL << 6)) & 0xff);
                if (j > a && j < n) wq |= nb << (8 * k);
            }
        }
        ccol[q] = wq;
    }
    const uint32_t tm = S.tmask[slot];
    int hits = 0;
    if (tm) hits = (p1 >= 0 && ((tm >> p1) & 1)) + (p2 >= 0 && ((tm >> p2) & 1)) +
                   (l1 >= 0 && ((tm >> l1) & 1)) + (l2 >= 0 && ((tm >> l2) & 1));
    const int sc = S.score[slot];
    // completed assignments all share one (empty) state: drop the running third-generation count
    reinterpret_cast<int4 *>(crec)[0] = make_int4(sc + err, sc + tot, S.prob[slot], (n == a + 1) ? 0 : (int)S.ptgs[slot] + hits);
    A.links[cslot] = make_uint2(S.pidx[slot], code);
    A.dead[cslot] = (n == a + 1) ? 2 : 0;
}

// Body of the expansion kernel: `A` lives in shared memory (the stage functions take it by reference), `s_raw` is the
// block's dynamic shared memory (one GroupShared per warp).  Called by k_expand and by the fused program kernel.
template <int NMAX, bool SCALED>
__device__ __forceinline__ void expand_body(const LevelArgs &A, uint8_t *s_raw)
{
    if (*A.overflow) return;                // an earlier level of this pass did not fit (block-uniform)
    GroupShared<NMAX> &S = reinterpret_cast<GroupShared<NMAX> *>(s_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int a = A.level;
    const uint32_t pcount = min(*A.pcount, A.cap);
    // records per warp group: spread a small frontier over all resident warps
    const uint32_t nwarps = gridDim.x * WPB;
    const uint32_t grp_n = min((uint32_t)GRP, max(1u, (pcount + A.gpw * nwarps - 1) / (A.gpw * nwarps)));   // >= gpw groups per warp when possible

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = *reinterpret_cast<volatile uint32_t *>(A.overflow) ? 0xffffffffu : atomicAdd(A.work_counter, grp_n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= pcount) break;          // no work left, or this level has overflowed: stop early
        __syncwarp();

        // ---- phase 1: lane = record; column state in registers, masks, the record's own bound, candidate count.
        //      Records with at least one configuration are then staged in consecutive slots.
        const bool have = (uint32_t)lane < grp_n && base + lane < pcount;
        const DevProblem *P = nullptr;
        int my_n = 0, slack = -1, lmax = 0;
        uint32_t pidx = 0;
        int4 hd = make_int4(0, 0, 0, 0);
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (have) {
            // no list: the frontier is the dense array of the previous level's survivors (k_materialise)
            pidx = A.plist ? A.plist[base + lane] : base + lane;
            const uint8_t *prec = A.parents + (size_t)pidx * A.stride;
            hd = reinterpret_cast<const int4 *>(prec)[0];                   // score, f, prob, tgs
            const uint4 c0 = reinterpret_cast<const uint4 *>(prec + offsetof(NodeHdr, col))[0];
            w[0] = c0.x; w[1] = c0.y; w[2] = c0.z; w[3] = c0.w;
            if (NMAX > 16) {
                const uint4 c1 = reinterpret_cast<const uint4 *>(prec + offsetof(NodeHdr, col))[1];
                w[4] = c1.x; w[5] = c1.y; w[6] = c1.z; w[7] = c1.w;
            }
            P = &A.probs[hd.z];
            my_n = P->n; lmax = P->lmax;
            slack = A.limit[hd.z] - hd.x;
            // heuristic band (the beam only feeds the incumbent): a child may exceed its parent's f by
            // beam_delta, or by half of what is left up to the current limit if that is more
            if (A.beam_delta >= 0) {
                const bool shortp = my_n <= A.beam_short_n;
                slack = min(slack, hd.y - hd.x + max(shortp ? A.beam_delta_short : A.beam_delta, (A.limit[hd.z] - hd.y) >> (shortp ? A.beam_shift_short : A.beam_shift)));
            }
        }
        // masks of all columns at once, 4 column bytes per word (bytes of closed columns and beyond the partition are
        // zero): a byte-wise "at least" test leaves bit 7 of each byte, one multiply gathers the four bits
        uint32_t okP = 0, parP = 0, okL1 = 0, okL2 = 0, par1 = 0, par2 = 0;
        int covC = 0;
#pragma unroll
        for (int q = 0; q < NMAX / 4; ++q) {
            const uint32_t x = w[q];
            const uint32_t vP = x & 0x03030303u, vL = (x >> 2) & 0x03030303u;
            const uint32_t g_okP = ((x & 0x30303030u) + 0x7f7f7f7fu) & 0x80808080u;          // bP >= 1
            const uint32_t g_parP = (vP + 0x7f7f7f7fu) & 0x80808080u;                          // nP >= 1
            const uint32_t g_okL1 = (((x >> 6) & 0x03030303u) + 0x7f7f7f7fu) & 0x80808080u;    // bL >= 1
            const uint32_t g_okL2 = (x & 0x80808080u);                                          // bL >= 2
            const uint32_t g_par1 = ((vP + vL) + 0x7f7f7f7fu) & 0x80808080u;                   // nP + nL >= 1
            const uint32_t g_par2 = ((vP + vL) + 0x7e7e7e7eu) & 0x80808080u;                   // nP + nL >= 2
            okP |= ((g_okP * 0x00204081u) >> 28) << (4 * q);
            parP |= ((g_parP * 0x00204081u) >> 28) << (4 * q);
            okL1 |= ((g_okL1 * 0x00204081u) >> 28) << (4 * q);
            okL2 |= ((g_okL2 * 0x00204081u) >> 28) << (4 * q);
            par1 |= ((g_par1 * 0x00204081u) >> 28) << (4 * q);
            par2 |= ((g_par2 * 0x00204081u) >> 28) << (4 * q);
            covC += (int)(((vP + 2 * vL) * 0x01010101u) >> 24);                                // committed coverage of segment a
        }
        if (a == 0 && my_n > 0) { const uint32_t rb = 1u << (my_n - 1); parP |= rb; par1 |= rb; par2 |= rb; }   // the root pair needs no parent
        if (lmax < 1) okL1 = 0;
        if (lmax < 2) okL2 = 0;
        const uint32_t wa = (a >> 2) == 0 ? w[0] : ((a >> 2) == 1 ? w[1] : ((a >> 2) == 2 ? w[2] : ((a >> 2) == 3 ? w[3] :
                            ((a >> 2) == 4 ? w[4] : ((a >> 2) == 5 ? w[5] : ((a >> 2) == 6 ? w[6] : w[7]))))));
        const uint32_t ba = (wa >> (8 * (a & 3))) & 0xffu;
        const int fbC = (int)(ba & 3) + 2 * (int)((ba >> 2) & 3);
        // every open record within the limit has at least the empty configuration: it takes a staging slot
        const bool staged = have && my_n > a && slack >= 0;
        const uint32_t nzm = __ballot_sync(0xffffffffu, staged);
        const int ns = __popc(nzm);                       // staged records
        const int slot = __popc(nzm & lt_mask);
        const int jtop = (int)__reduce_max_sync(0xffffffffu, staged ? (unsigned)my_n : 0u);
        uint32_t total = 0;
        if (staged) {
#pragma unroll
            for (int q = 0; q < NMAX / 4; ++q) S.colw[q][slot] = w[q];
            // The record's own bound on segments > a (what its parent level charged for them): every child's
            // bound on those segments is at least this (coverage only grows, budgets only shrink), so
            // err + pf > slack cuts a configuration before the O(n) bound is evaluated for it.
            int pf = 0, cov = 0, capsum = 0;
            uint32_t over = 0, fover = 0;
            const int ush = SCALED ? P->ush : 0;
            const uint32_t cja = __ldg(&P->cj[a]);
#pragma unroll 1
            for (int j = my_n - 1; j > a; --j) {
                const uint32_t b = S.colb(j, slot);
                const uint32_t cjw = __ldg(&P->cj[j]);
                const int nP = b & 3, nL = (b >> 2) & 3, bP = (b >> 4) & 3, bL = (b >> 6) & 3;
                const int cap = min(bP, 2) + 2 * min(bL, 2);
                const int wn = nP + 2 * nL;
                cov += wn; capsum += cap;
                const int fmax = min(6 * (j - a + 1), capsum);
                const int d = (cov << (ush + 1)) - (int)(cjw & 0xffff);
                pf += d > 0 ? d : max(0, -d - (fmax << (ush + 1)));
                const int f = (wn << ush) - (int)(cjw >> 16);
                pf += f > 0 ? f : max(0, -f - ((cap + min(capsum, 6)) << ush));
                over |= (uint32_t)(d >= 0) << j; fover |= (uint32_t)(f >= 0) << j;       // the sharper prefilter of the light test
            }
            dec::fill_ovt(S, slot, over, fover, my_n);
            const int lslack = slack - pf;
            const int d0 = (covC << (ush + 1)) - (int)(cja & 0xffffu), f0 = (fbC << ush) - (int)(cja >> 16);
            // Columns the prefilter would refuse in every configuration are taken out of the masks before anything is
            // counted or listed: an end at column c charges ovt[c] (a loop end twice that), and the exact deviation of
            // segment a is at least its minimum over the weights w the end can come with -- the configurations that
            // disappear are exactly ones the light test would drop one by one.  They still count as candidates.
            uint32_t nl0 = 0;
            const uint32_t total_all = dec::count_configs_call(okP, parP, okL1, okL2, par1, par2, &nl0);
            total = total_all;
            if (lslack < 0) total = 0;                    // the record's own bound is already beyond the limit
            else {
                if (dec::restrict_masks(over, fover, lslack, d0, f0, ush, &okP, &okL1, &okL2))
                    total = dec::count_configs_call(okP, parP, okL1, okL2, par1, par2, &nl0);
                dec::write_lists2(S, slot, okP, parP, okL1, okL2, par1, par2, nl0);
            }
            atomicAdd(&A.cand_count[hd.z], (unsigned long long)total_all);
            S.ush[slot] = (uint8_t)ush;
            S.pidx[slot] = pidx; S.score[slot] = hd.x; S.prob[slot] = hd.z; S.ptgs[slot] = (uint8_t)hd.w;
            S.slack[slot] = slack; S.lslack[slot] = lslack; S.n[slot] = (uint8_t)my_n;
            S.d0[slot] = d0; S.f0[slot] = f0;
            S.tmask[slot] = P->has_tgs ? P->tgs_mask[a] : 0u;
            S.ncfg[slot] = total;
        }
        // phase 2 works in chunks of CHUNK consecutive configurations of a record (a record left without any still owns
        // one, empty, chunk): exclusive prefix of the chunk counts over the staged records
        uint32_t incl = staged ? max(1u, (total + CHUNK - 1) / CHUNK) : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const uint32_t TI = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t T = __reduce_add_sync(0xffffffffu, total);
        if (staged) S.cstart[slot + 1] = incl;
        if (lane >= ns) S.cstart[lane + 1] = TI;          // beyond the last staged record
        if (lane == 0) S.cstart[0] = 0;
        if (A.emit == EMIT_BEST) S.best()[lane] = ~0ull;
        S.cnt[lane] = 0;
        __syncwarp();

        // ---- phase 2: lane = a chunk of consecutive configurations of one record (flattened over the group's
        //      records), stepped through by a cursor; three stages fed through two shared-memory queues so that
        //      each stage runs with full lanes
        int qh = 0, qw = 0;                 // queue fills (warp-uniform)
        uint32_t ib0 = 0, s0 = 0;           // s0: staged record that owns chunk ib0
        int st = 0, rsteps = 0;             // step within the current round of chunks / its length (warp-uniform)
        int own = 0, left = 0;              // this lane's record and configurations in its chunk
        uint32_t my_nl0 = 1;
        dec::Cursor cur;
        cur.m_rem = 0; cur.li = 0; cur.width = 1; cur.hi = -1; cur.pi = 0; cur.kP = 0; cur.pcode = 0; cur.pw = 0;
        uint32_t n_heavy = 0, n_write = 0, n_hcalls = 0;  // funnel statistics (warp-uniform)
        for (;;) {
            const bool enum_done = st >= rsteps && ib0 >= TI;
            if (qw >= 32 || (enum_done && qh == 0 && qw > 0)) {
                const int cnt = min(qw, 32);
                n_write += cnt;
                write_batch<NMAX>(S, A, cnt);
                __syncwarp();
                qw -= cnt;
                uint32_t m0 = 0; int m1 = 0, m2 = 0;
                if (lane < qw) { m0 = S.qw_cfg[32 + lane]; m1 = S.qw_err[32 + lane]; m2 = S.qw_tot[32 + lane]; }
                __syncwarp();
                if (lane < qw) { S.qw_cfg[lane] = m0; S.qw_err[lane] = m1; S.qw_tot[lane] = m2; }
                __syncwarp();
                continue;
            }
            if (qh >= 32 || (enum_done && qh > 0)) {
                const int cnt = min(qh, 32);
                n_heavy += cnt; ++n_hcalls;
                __syncwarp();               // the queue entries of the last steps
                qw = heavy_batch<NMAX, SCALED>(S, A, cnt, jtop, qw);
                __syncwarp();
                qh -= cnt;
                const uint32_t mv = lane < qh ? S.qh[32 + lane] : 0u;
                __syncwarp();
                if (lane < qh) S.qh[lane] = mv;
                __syncwarp();
                continue;
            }
            if (enum_done) break;
            if (st >= rsteps) {             // next round: 32 chunks, one per lane
                const uint32_t bm = dec::owner_mask(S, s0, ib0, lane);
                const uint32_t it = ib0 + lane;
                left = 0;
                if (it < TI) {
                    own = (int)(s0 + __popc(bm & lt_mask));
                    const uint32_t r0 = (it - S.cstart[own]) * CHUNK;
                    left = min((int)CHUNK, (int)S.ncfg[own] - (int)r0);
                    if (left > 0) { my_nl0 = S.nl0[own]; dec::cursor_start(S, own, r0, cur); }
                }
                s0 += __popc(bm);
                ib0 += 32;
                rsteps = (int)__reduce_max_sync(0xffffffffu, (unsigned)left);
                st = 0;
            }
            bool pass = false;
            uint32_t entry = 0;
            if (st < left) {
                entry = dec::cursor_eval<SCALED>(S, own, cur, my_nl0, &pass);
                // candidates scored = enumerated minus the (few) structurally infeasible ones, per record
                if (!entry) atomicAdd(&S.cnt[own], 1u);
                dec::cursor_advance(S, own, cur, my_nl0);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            if (pass) S.qh[qh + __popc(bal & lt_mask)] = entry;
            qh += __popc(bal);
            ++st;
        }
        if (A.emit == EMIT_BEST) {
            // every record emits its arg-min child (tot, then search order)
            __syncwarp();
            const unsigned long long bk = S.best()[lane];
            const bool has = lane < ns && bk != ~0ull;
            const uint32_t hm = __ballot_sync(0xffffffffu, has);
            if (has) {
                const int n = S.n[lane];
                const uint32_t code = (uint32_t)bk;       // recover the configuration from its search-order code
                const int ip = code >> 10, il = code & 1023;
                int p1 = -1, p2 = -1, l1 = -1, l2 = -1;
                if (ip) { int q1 = 0; while (ambi_ip_code(q1 + 1, -1) <= ip) ++q1; const int off = ip - ambi_ip_code(q1, -1);
                          p1 = n - 1 - q1; if (off > 0) p2 = n - 1 - (off - 1); }
                if (il) { int q1 = 0; while (ambi_il_code(q1 + 1, -1) <= il) ++q1; const int off = il - ambi_il_code(q1, -1);
                          l1 = n - 1 - q1; if (off > 0) l2 = (off - 1 == q1) ? l1 : n - 1 - (off - 1); }
                const int tot = (int)(bk >> 32);
                const int w2 = (p1 >= 0) + (p2 >= 0) + 2 * ((l1 >= 0) + (l2 >= 0));
                const int ush = SCALED ? (int)S.ush[lane] : 0;
                const int err = abs(S.d0[lane] + (w2 << (ush + 1))) + abs(S.f0[lane] + (w2 << ush));
                const int pos = __popc(hm & lt_mask);
                S.qw_cfg[pos] = ((uint32_t)lane << 24) | pack_cfg(p1, p2, l1, l2); S.qw_err[pos] = err; S.qw_tot[pos] = tot;
            }
            __syncwarp();
            if (hm) { n_write += __popc(hm); write_batch<NMAX>(S, A, __popc(hm)); __syncwarp(); }
        }
        // ---- statistics (per problem of each staged record)
        __syncwarp();
        if (lane < 4 && T) atomicAdd(&A.funnel[lane], (unsigned long long)(lane == 0 ? T : (lane == 1 ? n_heavy : (lane == 2 ? n_write : n_hcalls))));
        if (lane < ns) {
            // candidates of a record = its configurations (counted in phase 1) minus the structurally infeasible ones met
            if (S.cnt[lane]) atomicAdd(&A.cand_count[S.prob[lane]], 0ull - (unsigned long long)S.cnt[lane]);
            atomicAdd(&A.node_count[S.prob[lane]], 1ull);
        }
        __syncwarp();
    }
}

#ifndef AMBI_EMU
template <int NMAX, bool SCALED>
__global__ void __launch_bounds__(WPB * 32, NMAX <= 16 ? 4 : 3)
k_expand(LevelArgs A_param)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ LevelArgs s_args;
    if (threadIdx.x == 0) s_args = A_param;
    __syncthreads();
    expand_body<NMAX, SCALED>(s_args, s_raw);
}

#endif

// ---------------------------------------------------------------------------
// K3a: merge records that reached the same column state.
// A byte per key record says what it is: 0 an open record, 2 a completed assignment (both written by k_expand with
// the child), 1 dropped by the merge.  The keys themselves are never modified, a level's flags stay L2 resident, and
// the passes that follow (finalize, beam histogram, materialise) read a key only when they need the record.
// ---------------------------------------------------------------------------

struct DedupArgs {
    uint8_t *recs;
    const uint32_t *count;
    uint32_t cap;
    unsigned long long *table;
    uint32_t tmask;
    const uint32_t *gen_base;   // bumped once per run, so a replayed graph never sees stale slots
    uint32_t gen_off;
    int stride;
    int level;          // rows 0..level are fixed in these records
    const uint32_t *overflow;
    // incumbent beam only (fmin_cur != nullptr): the merge also finds the level's minimum f per problem (a record
    // displaced later has an f at least as large as its displacer's, so records alive at insertion suffice)
    const DevProblem *probs;
    int32_t *fmin_cur;          // [prob]
    int nmax;                   // longest partition of the batch
    const uint2 *links;         // per key record: (parent, row code)
    const uint8_t *parents;     // full records of the previous level (prefix rows for the tie-break)
    int pstride;
    uint8_t *dead;              // [record] dropped by the merge
    int tshift;                 // table slots per record: at least 1 << tshift
};

// beam bins: width 1 from the level's minimum f, everything beyond in the last bin (the cut for the W best
// records lies at the low end; measured: bins scaled to the whole [minimum, limit] range, or based on the
// previous level's minimum, lose the resolution that matters and cost hard partitions their incumbent)
__device__ __forceinline__ int beam_bin(int f, int base) { return max(0, min(f - base, BEAM_BINS - 1)); }


// `wide`: some partition of the batch has more than 16 segments (otherwise columns 16..31 are zero everywhere
// and only the first sector of either record is touched)
__device__ __forceinline__ bool same_state(const NodeHdr *x, const NodeHdr *y, bool wide) {
    if (x->prob != y->prob || x->tgs != y->tgs) return false;
    const uint4 *a = reinterpret_cast<const uint4 *>(x->col), *b = reinterpret_cast<const uint4 *>(y->col);
    const uint4 a0 = a[0], b0 = b[0];
    if (!(a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w)) return false;
    if (!wide) return true;
    const uint4 a1 = a[1], b1 = b[1];
    return a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
}

// strict order: lower score, then smaller prefix in search order.  A child's prefix is its parent's rows followed by
// its own row code; the parents are full records of the previous level (read only when the scores tie).
__device__ __forceinline__ bool rec_better(const DedupArgs &A, uint32_t i, uint32_t o)
{
    const NodeHdr *hx = reinterpret_cast<const NodeHdr *>(A.recs + (size_t)i * A.stride);
    const NodeHdr *hy = reinterpret_cast<const NodeHdr *>(A.recs + (size_t)o * A.stride);
    if (hx->score != hy->score) return hx->score < hy->score;
    const uint2 lx = A.links[i], ly = A.links[o];
    if (lx.x != ly.x) {
        const uint32_t *rx = reinterpret_cast<const uint32_t *>(A.parents + (size_t)lx.x * A.pstride + sizeof(NodeHdr));
        const uint32_t *ry = reinterpret_cast<const uint32_t *>(A.parents + (size_t)ly.x * A.pstride + sizeof(NodeHdr));
        for (int k = 0; k < A.level; ++k) if (rx[k] != ry[k]) return rx[k] < ry[k];
    }
    return lx.y < ly.y;
}

__device__ __forceinline__ void dedup_body(const DedupArgs &A)
{
    if (*A.overflow) return;
    uint32_t cnt = min(*A.count, A.cap);
    const uint32_t gen = *A.gen_base + A.gen_off;
    // use only as much of the table as this level needs (>= 1 << tshift slots per record), so that small
    // levels stay L2 resident; generation stamps make whatever is left there harmless
    uint32_t tmask = 1023u;
    while (tmask < A.tmask && (tmask >> A.tshift) < cnt) tmask = (tmask << 1) | 1u;
    tmask = min(tmask, A.tmask);
    const bool wide = A.nmax > 16;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        uint8_t *rec = A.recs + (size_t)i * A.stride;
        NodeHdr *h = reinterpret_cast<NodeHdr *>(rec);
        const uint32_t *cw = reinterpret_cast<const uint32_t *>(h->col);
        uint32_t hv = mix32((uint32_t)h->prob * 0x9e3779b9u + (uint32_t)h->tgs);
#pragma unroll
        for (int k = 0; k < 4; ++k) hv = mix32(hv ^ cw[k]) + 0x85ebca6bu * (k + 1);
        if (wide) {
#pragma unroll
            for (int k = 4; k < 8; ++k) hv = mix32(hv ^ cw[k]) + 0x85ebca6bu * (k + 1);
        }
        uint32_t slot = hv & tmask;
        const unsigned long long mine = ((unsigned long long)gen << 32) | i;
        bool alive = false;
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&A.table[slot]);
            if ((uint32_t)(cur >> 32) != gen) {
                unsigned long long old = atomicCAS(&A.table[slot], cur, mine);
                if (old == cur) { alive = true; break; }
                continue;                              // re-read the same slot
            }
            uint32_t o = (uint32_t)cur;
            uint8_t *orec = A.recs + (size_t)o * A.stride;
            if (same_state(h, reinterpret_cast<const NodeHdr *>(orec), wide)) {
                if (rec_better(A, i, o)) {
                    unsigned long long old = atomicCAS(&A.table[slot], cur, mine);
                    if (old == cur) { A.dead[o] = 1; alive = true; break; }
                    continue;
                } else { A.dead[i] = 1; break; }
            }
            slot = (slot + 1) & tmask;
        }
        if (A.fmin_cur && alive) {
            if (A.dead[i] != 2) atomicMin(&A.fmin_cur[h->prob], h->f);   // completed assignments are not part of the beam
        }
    }
}

// ---------------------------------------------------------------------------
// K3b: completed assignments -> per-problem best; the rest -> next frontier.
// ---------------------------------------------------------------------------
struct FinalizeArgs {
    const DevProblem *probs;
    const uint8_t *recs;
    const uint32_t *count;
    uint32_t cap;
    int32_t *best_score;        // [prob]
    uint32_t *best_rows;        // [prob][AMBI_MAXN]
    int32_t *limit;             // chunked mode: tighten the pruning limit as completed assignments arrive
    int stride;
    int level;
    int merge_best;             // chunked mode: keep the better of (stored best, this chunk's best)
    const uint32_t *overflow;
    int32_t *seen;              // [prob] best completed score seen so far in any pass ...
    unsigned long long *stamps; // ... and when (%globaltimer) it was first seen: [prob][4], word 1
    const uint2 *links;         // the key records' (parent, row code)
    const uint8_t *parents;     // full records of the previous level
    int pstride;
    const uint8_t *dead;
};

__device__ __forceinline__ void finalize_body(const FinalizeArgs &A)
{
    if (*A.overflow) return;
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        if (A.dead[i] != 2) continue;
        const uint8_t *rec = A.recs + (size_t)i * A.stride;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(rec);
        const int n = A.probs[h->prob].n;
        {
            // all completed assignments of a problem share the empty column state, so the
            // merge above left exactly one alive: the arg-min in (score, search order)
            const uint2 lk = A.links[i];
            const uint32_t *r = reinterpret_cast<const uint32_t *>(A.parents + (size_t)lk.x * A.pstride + sizeof(NodeHdr));   // rows 0..n-2; row n-1 = lk.y
            uint32_t *o = A.best_rows + (size_t)h->prob * AMBI_MAXN;
            bool take = true;
            if (A.merge_best) {     // one alive completed record per problem and launch: no race
                const int old = A.best_score[h->prob];
                if (old < h->score) take = false;
                else if (old == h->score) {
                    take = false;
                    for (int k = 0; k < n; ++k) { const uint32_t rk = k < n - 1 ? r[k] : lk.y; if (rk != o[k]) { take = rk < o[k]; break; } }
                }
            }
            if (take) {
                if (A.stamps && h->score < A.seen[h->prob]) { A.seen[h->prob] = h->score; A.stamps[(size_t)h->prob * 4 + 1] = globaltimer_ns(); }
                A.best_score[h->prob] = h->score;
                for (int k = 0; k < n; ++k) o[k] = k < n - 1 ? r[k] : lk.y;
                if (A.limit && h->score < A.limit[h->prob]) A.limit[h->prob] = h->score;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K3c: beam selection (incumbent pass): keep the W best-f records per problem.
// ---------------------------------------------------------------------------
struct BeamArgs {
    const DevProblem *probs;
    const uint8_t *recs;
    const uint32_t *count;
    uint32_t cap;
    const int32_t *fmin;        // [prob] minimum f of the level (k_dedup): base of the bins
    uint32_t *hist;             // [prob][BEAM_BINS]
    uint32_t *cutq;             // [prob] cut bin << 16 | records still to take from it
    uint32_t *ticket;           // [prob]
    int nprob;
    int width_floor;            // batch-scaled width (a fixed budget of beam records per batch)
    int width;                  // per-problem beam width of a 12-segment partition ...
    float growth;               // ... doubled every 1/growth further segments (a longer partition needs a wider beam)
    int width_div;              // cooperative beam: this rank runs 1/width_div of the width on its share of the prefixes
    int stride;
    int level;
    const uint8_t *dead;
};

__device__ __forceinline__ void beam_hist_body(const BeamArgs &A)
{
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        if (A.dead[i]) continue;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(A.recs + (size_t)i * A.stride);
        const int prob = h->prob;
        atomicAdd(&A.hist[(size_t)prob * BEAM_BINS + beam_bin(h->f, A.fmin[prob])], 1u);
    }
}

// per problem: the bin in which the cumulative count reaches the beam width, and how many records of that bin
// still fit (every record in a lower bin is kept, records of the cut bin draw tickets)
__device__ __forceinline__ void beam_cut_body(const BeamArgs &A)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < A.nprob; p += gridDim.x * blockDim.x) {
    const uint32_t width = max((uint32_t)min(A.width, 16), (uint32_t)max((float)A.width_floor, min(8192.0f, (float)A.width * exp2f(A.growth * (float)max(0, A.probs[p].n - 12)))) / (uint32_t)A.width_div);
    const uint32_t *hs = A.hist + (size_t)p * BEAM_BINS;
    uint32_t below = 0;
    int cut = BEAM_BINS;                                // keep everything
    uint32_t quota = 0;
    for (int k = 0; k < BEAM_BINS; ++k) {
        const uint32_t c = hs[k];
        if (below + c > width) { cut = k; quota = width - below; break; }
        below += c;
    }
    A.cutq[p] = ((uint32_t)cut << 16) | quota;
    }
}

// ---------------------------------------------------------------------------
// K3d: full records for the survivors of a level (alive after the merge, not completed, and -- in the beam pass --
// among the W best of their problem): key + the parent's rows + the row code, appended densely to the next
// frontier.  One warp-wide ballot per 32 key records keeps the appends at one atomic per warp.
// ---------------------------------------------------------------------------
struct MatArgs {
    const DevProblem *probs;
    const uint8_t *keys; const uint2 *links; const uint32_t *count; uint32_t cap;
    const uint8_t *parents; int pstride;
    uint8_t *out; uint32_t *out_count; uint32_t out_cap;
    int kstride, level, nmax;
    const uint32_t *overflow;
    // beam selection (cutq != nullptr): keep below the cut bin, draw a ticket in the cut bin
    const int32_t *fmin; const uint32_t *cutq; uint32_t *ticket;
    const uint8_t *dead;
};

__device__ __forceinline__ void materialise_body(const MatArgs &A)
{
    if (*A.overflow) return;
    const uint32_t cnt = min(*A.count, A.cap);
    const int lane = threadIdx.x & 31;
    const int nq = (A.nmax + 3) >> 2;
    const uint32_t wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    // a warp takes 128 key records at a time, four per lane: one word of flags per lane, then -- for the open
    // records only -- independent loads of keys, links and parent rows (the kernel waits on memory, not on issue
    // slots), and one append to the frontier counter per 128 records
    for (uint32_t w0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 128u; w0 < cnt; w0 += nwarps * 128u) {
        const uint32_t ib = w0 + 4u * (uint32_t)lane;
        uint32_t dw = 0x01010101u;
        if (ib < cnt) dw = *reinterpret_cast<const uint32_t *>(A.dead + ib);      // (the flag array is padded to a multiple of 4)
        uint32_t km = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (ib + k < cnt && ((dw >> (8 * k)) & 0xffu) == 0u) km |= 1u << k;
        if (A.cutq) {
            int4 hd[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) hd[k] = ((km >> k) & 1u) ? *reinterpret_cast<const int4 *>(A.keys + (size_t)(ib + k) * A.kstride) : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((km >> k) & 1u) {
                const int bin = beam_bin(hd[k].y, A.fmin[hd[k].z]);
                const uint32_t cq = A.cutq[hd[k].z];
                const int cut = (int)(cq >> 16);
                bool keep = bin < cut;
                if (bin == cut) keep = atomicAdd(&A.ticket[hd[k].z], 1u) < (cq & 0xffffu);
                if (!keep) km &= ~(1u << k);
            }
        }
        const int c = __popc(km);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (lane == 31 && total) base = atomicAdd(A.out_count, (uint32_t)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t o = base + (uint32_t)(incl - c);
        const int a = A.level;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((km >> k) & 1u)) continue;
            const uint32_t i = ib + k;
            if (o < A.out_cap) {                                 // always: out_cap >= cap
                uint8_t *dst = A.out + (size_t)o * A.pstride;
                const uint4 *ksrc = reinterpret_cast<const uint4 *>(A.keys + (size_t)i * A.kstride);
                const uint2 lk = A.links[i];
                const uint4 k0 = ksrc[0], k1 = ksrc[1];
                const uint4 k2 = A.kstride > 32 ? ksrc[2] : make_uint4(0, 0, 0, 0);
                const uint4 *src = reinterpret_cast<const uint4 *>(A.parents + (size_t)lk.x * A.pstride + sizeof(NodeHdr));
                reinterpret_cast<uint4 *>(dst)[0] = k0;
                reinterpret_cast<uint4 *>(dst)[1] = k1;
                reinterpret_cast<uint4 *>(dst)[2] = k2;
                uint4 *rd = reinterpret_cast<uint4 *>(dst + sizeof(NodeHdr));
                for (int q = 0; q < nq; ++q) {
                    uint4 v = src[q];
                    if ((a >> 2) == q) { const int r = a & 3; if (r == 0) v.x = lk.y; else if (r == 1) v.y = lk.y; else if (r == 2) v.z = lk.y; else v.w = lk.y; }
                    rd[q] = v;
                }
            }
            ++o;
        }
    }
}

#ifndef AMBI_EMU
__global__ void k_dedup(DedupArgs A) { dedup_body(A); }
__global__ void k_finalize(FinalizeArgs A) { finalize_body(A); }
__global__ void k_beam_hist(BeamArgs A) { beam_hist_body(A); }
__global__ void k_beam_cut(BeamArgs A) { beam_cut_body(A); }
__global__ void k_materialise(MatArgs A) { materialise_body(A); }
#endif

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
// `active`: partitions this (global-memory) path still has to solve; the others were solved by the
// warp-per-partition kernel (stage_small.cu) and keep their results
struct InitRootsArgs {
    const DevProblem *probs; int count; uint8_t *recs; int stride, nmax;
    uint32_t *list, *lcount; int32_t *limit, *best_score; int first; const uint8_t *active;
};

__device__ __forceinline__ void init_roots_body(const InitRootsArgs &A)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < A.count; p += gridDim.x * blockDim.x) {
        if (!A.active[p]) continue;
        const DevProblem &P = A.probs[p];
        if (A.first) { A.limit[p] = P.zero_score; A.best_score[p] = 0x7fffffff; }
        if (P.n < 2) continue;                // n == 1 handled by k_solve_n1
        uint32_t s = atomicAdd(A.lcount, 1u);
        uint8_t *rec = A.recs + (size_t)s * A.stride;
        NodeHdr *h = reinterpret_cast<NodeHdr *>(rec);
        for (int b = 0; b < AMBI_MAXN; ++b) h->col[b] = b < P.n ? (uint8_t)((AMBI_BUD_INF << 4) | (AMBI_BUD_INF << 6)) : 0;
        h->score = 0; h->f = P.root_lb; h->prob = p; h->tgs = 0;
        uint32_t *r = reinterpret_cast<uint32_t *>(rec + sizeof(NodeHdr));
        for (int k = 0; k < A.nmax; ++k) r[k] = 0;
        A.list[s] = s;
    }
}
#ifndef AMBI_EMU
__global__ void k_init_roots(InitRootsArgs A) { init_roots_body(A); }

// Single-segment partitions: no a<b rows exist; P[s,s] in {0,1}, L[s,s] in [0,loop_ub]
// (LGM:4708-4724); the pattern coefficient of the fold-back row is 0 (LGM:4464-4476).
__global__ void k_solve_n1(const DevProblem *probs, int count, int32_t *best_score, uint32_t *best_rows,
                           unsigned long long *cand_count)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count || probs[p].n != 1) return;
    const DevProblem &P = probs[p];
    int best = 0x7fffffff, bp = 0, bl = 0;
    for (int pp = 0; pp <= 1; ++pp)
        for (int l = 0; l <= P.loop_ub; ++l) {
            int e = abs(((pp + 2 * l) << (P.ush + 1)) - P.c2[0]) + abs((l << (P.ush + 1)) - P.j2[0]);
            if (e < best) { best = e; bp = pp; bl = l; }
        }
    best_score[p] = best;
    best_rows[(size_t)p * AMBI_MAXN] = 0x80000000u | ((uint32_t)bp << 20) | (uint32_t)bl;   // raw (P,L) for n==1
    cand_count[p] += 2ull * (unsigned long long)(P.loop_ub + 1);
}

__global__ void k_bump_gen(uint32_t *gen) { *gen += 128u; }

// start of a search: nothing solved yet, statistics cleared, the partitions that take the global-memory path flagged
__global__ void k_prepare(const DevProblem *probs, int count, int32_t *best_score, int32_t *seen, int32_t *incumbent, unsigned long long *stamps, unsigned long long *cand,
                          unsigned long long *node, uint8_t *active, const uint8_t *active_base)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    best_score[p] = 0x7fffffff; seen[p] = 0x7fffffff; incumbent[p] = probs[p].zero_score; cand[p] = 0; node[p] = 0; active[p] = active_base[p];
    stamps[(size_t)p * 4] = 0; stamps[(size_t)p * 4 + 1] = 0; stamps[(size_t)p * 4 + 2] = 0; stamps[(size_t)p * 4 + 3] = 0;
}

#endif
struct SetLimitArgs {
    const DevProblem *probs; int count; int32_t *limit; const int32_t *best_score; int32_t *incumbent;
    const uint8_t *active; int32_t *seen; unsigned long long *stamps;
};

__device__ __forceinline__ void set_limit_body(const SetLimitArgs &A)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < A.count; p += gridDim.x * blockDim.x) {
        if (!A.active[p]) continue;
        if (A.best_score[p] < A.seen[p]) { A.seen[p] = A.best_score[p]; A.stamps[(size_t)p * 4 + 1] = globaltimer_ns(); }   // incumbent of the pass that just ended
        int z = A.limit[p], b = A.best_score[p];       // limit starts at the all-zero score (k_init_roots)
        int l = b < z ? b : z;
        A.limit[p] = l; A.incumbent[p] = l;
    }
}
#ifndef AMBI_EMU            // everything below is the launch side (CUDA runtime)
__global__ void k_set_limit(SetLimitArgs A) { set_limit_body(A); }

// ---------------------------------------------------------------------------
// The launch sequence of a search as a PROGRAM: a list of operations (one per kernel launch or clear) that is
// either played as individual launches -- plain, or captured into the CUDA graph that large batches replay --
// or interpreted by ONE cooperative kernel with a grid-wide barrier between operations (k_program).  A search
// of a single partition is some 200 dependent operations on frontiers of a few hundred records: launch latency,
// not work, is what it costs, and the barrier of a small resident grid is several times cheaper than a launch.
// ---------------------------------------------------------------------------
enum { OP_MEMSET32 = 0, OP_INIT_ROOTS, OP_EXPAND, OP_DEDUP, OP_FINALIZE, OP_BEAM_HIST, OP_BEAM_CUT, OP_MATERIALISE, OP_SET_LIMIT, OP_BUMP_GEN };

struct MemsetArgs { uint32_t *ptr; unsigned long long words; uint32_t value; };

struct Op {                     // plain data: it is copied to the device and into shared memory word by word
    int kind;
    int sync_after;             // the next operation depends on this one (grid barrier / stream order)
    union {
        MemsetArgs ms; InitRootsArgs ir; LevelArgs ex; DedupArgs dd; FinalizeArgs fz; BeamArgs bm; MatArgs mt; SetLimitArgs sl; uint32_t *gen;
    } u;
};

__global__ void k_memset32(MemsetArgs A)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.words; i += (unsigned long long)gridDim.x * blockDim.x) A.ptr[i] = A.value;
}

}  // namespace ambi
#include <cooperative_groups.h>
namespace ambi {

// Interpreter: every block walks the same program; an operation is a grid-stride body, followed by a grid barrier.
// Before an expansion whose frontier has outgrown what this small grid should chew on, the kernel stops and reports
// the operation it stopped at: the host plays the rest as ordinary full-width launches (same buffers, same counters).
template <int NMAX, bool SCALED>
__global__ void __launch_bounds__(WPB * 32, NMAX <= 16 ? 4 : 3)
k_program(const Op *ops, int first, int n_ops, uint32_t big_frontier, uint32_t *stopped_at)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ Op s_op;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int i = first; i < n_ops; ++i) {
        __syncthreads();
        if (threadIdx.x < sizeof(Op) / 4) reinterpret_cast<uint32_t *>(&s_op)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&ops[i])[threadIdx.x];
        __syncthreads();
        const Op &op = s_op;
        switch (op.kind) {
        case OP_MEMSET32:
            for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < op.u.ms.words; k += (unsigned long long)gridDim.x * blockDim.x) op.u.ms.ptr[k] = op.u.ms.value;
            break;
        case OP_INIT_ROOTS: init_roots_body(op.u.ir); break;
        case OP_EXPAND:
            if (*op.u.ex.pcount > big_frontier) {               // the same value in every thread: written before the last barrier
                if (blockIdx.x == 0 && threadIdx.x == 0) *stopped_at = (uint32_t)i;
                return;
            }
            expand_body<NMAX, SCALED>(op.u.ex, s_raw);
            break;
        case OP_DEDUP: dedup_body(op.u.dd); break;
        case OP_FINALIZE: finalize_body(op.u.fz); break;
        case OP_BEAM_HIST: beam_hist_body(op.u.bm); break;
        case OP_BEAM_CUT: beam_cut_body(op.u.bm); break;
        case OP_MATERIALISE: materialise_body(op.u.mt); break;
        case OP_SET_LIMIT: set_limit_body(op.u.sl); break;
        case OP_BUMP_GEN: if (blockIdx.x == 0 && threadIdx.x == 0) *op.u.gen += 128u; break;
        }
        if (op.sync_after) grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *stopped_at = (uint32_t)n_ops;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    snprintf(err_, sizeof err_, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return -4; } } while (0)

static uint32_t next_pow2(uint64_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// device counters: one set per level so that a whole pass needs a single clear
enum { CTR_OVERFLOW = 0, CTR_GEN = 1, CTR_SMALL_WORK = 2, CTR_SMALL_FB = 3, CTR_STOP = 4, CTR_WORK = 8, CTR_CHILD = 8 + 40, CTR_FRONT = 8 + 80, CTR_TOTAL = 8 + 128 };

StageA::StageA() { err_[0] = 0; }
StageA::~StageA() { release(); }

void StageA::release()
{
    cudaFree(buf_[0]); cudaFree(buf_[1]); cudaFree(list_[0]); cudaFree(list_[1]); cudaFree(kbuf_); cudaFree(lbuf_); cudaFree(dbuf_); kbuf_ = nullptr; lbuf_ = nullptr; dbuf_ = nullptr;
    cudaFree(table_); cudaFree(counters_); cudaFree(d_funnel_); cudaFree(d_flag_); cudaFree(pool_); pool_ = nullptr; d_funnel_ = nullptr; d_flag_ = nullptr;
    cudaFree(d_prog_); d_prog_ = nullptr; d_prog_cap_ = 0; prog_valid_ = false;
    cudaFree(d_arena_); d_arena_ = nullptr; cudaFree(d_stamp0_); d_stamp0_ = nullptr; cudaFreeHost(h_flag_); h_flag_ = nullptr;
    for (int i = 0; i < 2; ++i) if (ev_small_[i]) { cudaEventDestroy(ev_small_[i]); ev_small_[i] = nullptr; }
    buf_[0] = buf_[1] = nullptr; list_[0] = list_[1] = nullptr; table_ = nullptr;
    counters_ = nullptr;
    free_problem_arrays();
    for (int i = 0; i < 4; ++i) if (ev_[i]) { cudaEventDestroy(ev_[i]); ev_[i] = nullptr; }
    for (cudaEvent_t e : prof_ev_) cudaEventDestroy(e);
    prof_ev_.clear();
    if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
    if (graph_exec2_) { cudaGraphExecDestroy(graph_exec2_); graph_exec2_ = nullptr; }
}

void StageA::free_problem_arrays()
{
    cudaFree(d_probs_); cudaFree(d_limit_); cudaFree(d_best_score_); cudaFree(d_best_rows_);
    cudaFree(d_beam_score_); cudaFree(d_beam_rows_); cudaFree(d_cand_); cudaFree(d_node_);
    cudaFree(d_fmin_); cudaFree(d_hist_); cudaFree(d_ticket_); cudaFree(d_cutq_); cudaFree(d_incumbent_); d_cutq_ = nullptr;
    cudaFreeHost(h_probs_); cudaFreeHost(h_res_); h_probs_ = nullptr; h_res_ = nullptr; h_cap_ = 0;
    cudaFreeHost(h_compact_); cudaFree(d_compact_); h_compact_ = nullptr; d_compact_ = nullptr;
    cudaFree(d_small_list_); cudaFree(d_active_); cudaFree(d_active_base_); cudaFree(d_seen_); cudaFree(d_stamps_);
    d_small_list_ = nullptr; d_active_ = d_active_base_ = nullptr; d_seen_ = nullptr; d_stamps_ = nullptr;
    d_probs_ = nullptr; d_limit_ = d_best_score_ = d_beam_score_ = d_fmin_ = d_incumbent_ = nullptr;
    d_best_rows_ = d_beam_rows_ = d_hist_ = d_ticket_ = nullptr; d_cand_ = d_node_ = nullptr;
    prob_cap_ = 0;
}

int StageA::init(cudaStream_t stream, int64_t node_capacity, int beam_width, int rank, int world, int shard_level, bool use_small)
{
    stream_ = stream;
    stride_ = (int)ambi_node_stride(AMBI_MAXN);
    if (node_capacity > 0) cap_base_ = (uint32_t)std::min<int64_t>(node_capacity, 1ll << 28);
    else {
        // automatic size: a quarter of what is free right now, at most 32 GB (64 M records) -- several processes may
        // share the GPU (one context per Ambigram process), and everything else the library allocates later (problem
        // arrays, the depth-first pool, Stage B) has to fit beside it.  ~540 bytes per record: two frontier buffers
        // for the widest record, keys, links, flags, list and the merge table.
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t budget = std::min<uint64_t>(free_b / 4, 32ull << 30);
        cap_base_ = (uint32_t)std::max<uint64_t>(1u << 20, std::min<uint64_t>(budget / 540, 64ull << 20));
    }
    cap_ = cap_base_;
    beam_ = beam_width > 0 ? beam_width : 16;
    if (const char *e = getenv("AMBI_BEAM_WIDTH")) { int v = atoi(e); if (v > 0) beam_ = v; }   // tuning knobs; results do not depend on them
    if (const char *e = getenv("AMBI_BEAM_DELTA")) beam_delta_ = atoi(e);
    if (const char *e = getenv("AMBI_DEDUP_SHIFT")) dedup_shift_ = std::max(0, std::min(3, atoi(e)));
    if (const char *e = getenv("AMBI_GPW")) gpw_ = std::max(1, atoi(e));
    if (const char *e = getenv("AMBI_BEAM_GROWTH")) beam_growth_ = (float)atof(e);
    if (const char *e = getenv("AMBI_BEAM_FLOOR_MAX")) beam_floor_max_ = std::max(1, atoi(e));
    if (const char *e = getenv("AMBI_BEAM_SHIFT")) beam_shift_ = std::max(0, std::min(8, atoi(e)));
    if (const char *e = getenv("AMBI_BEAM_DELTA_SHORT")) beam_delta_short_ = atoi(e);
    if (const char *e = getenv("AMBI_BEAM_SHIFT_SHORT")) beam_shift_short_ = std::max(0, std::min(8, atoi(e)));
    if (const char *e = getenv("AMBI_BEAM_SHORT_N")) beam_short_n_ = atoi(e);
    rank_ = rank; world_ = world > 0 ? world : 1; shard_level_ = shard_level;
    // buffers are sized in bytes for the widest record (n = 32); narrower batches fit more records.  An automatic size
    // backs off (halves) when the device cannot give it -- several processes may share the GPU, each sizing itself
    // from what was free when it looked; an explicit node_capacity is taken as given.
    for (;;) {
        buf_bytes_ = (size_t)cap_base_ * stride_;
        cap_max_ = (uint32_t)std::min<uint64_t>(2ull * cap_base_, buf_bytes_ / ambi_node_stride(2));
        tsize_ = next_pow2((uint64_t)cap_max_ * 2);
        cudaError_t e = cudaMalloc(&buf_[0], buf_bytes_);
        if (e == cudaSuccess) e = cudaMalloc(&buf_[1], buf_bytes_);
        if (e == cudaSuccess) e = cudaMalloc(&kbuf_, (size_t)cap_max_ * 48);
        if (e == cudaSuccess) e = cudaMalloc(&lbuf_, (size_t)cap_max_ * sizeof(uint2));
        if (e == cudaSuccess) e = cudaMalloc(&dbuf_, (size_t)cap_max_ + 4);
        if (e == cudaSuccess) e = cudaMalloc(&list_[0], (size_t)cap_max_ * 4);
        if (e == cudaSuccess) e = cudaMalloc(&table_, (size_t)tsize_ * 8);
        if (e == cudaSuccess) break;
        cudaGetLastError();
        cudaFree(buf_[0]); cudaFree(buf_[1]); cudaFree(kbuf_); cudaFree(lbuf_); cudaFree(dbuf_); cudaFree(list_[0]); cudaFree(table_);
        buf_[0] = buf_[1] = nullptr; kbuf_ = nullptr; lbuf_ = nullptr; dbuf_ = nullptr; list_[0] = nullptr; table_ = nullptr;
        if (e != cudaErrorMemoryAllocation || node_capacity > 0 || cap_base_ <= (1u << 20)) {
            snprintf(err_, sizeof err_, "candidate buffers for %u records: %s", cap_base_, cudaGetErrorString(e)); return -4;
        }
        cap_base_ = std::max<uint32_t>(1u << 20, cap_base_ / 2);
    }
    cap_ = cap_base_;
    CK(cudaMemsetAsync(table_, 0, (size_t)tsize_ * 8, stream_));
    CK(cudaMalloc(&counters_, CTR_TOTAL * sizeof(uint32_t)));
    CK(cudaMalloc(&d_funnel_, 4 * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_flag_, 4 * sizeof(int32_t)));
    CK(cudaMemsetAsync(counters_, 0, CTR_TOTAL * sizeof(uint32_t), stream_));
    CK(cudaStreamSynchronize(stream_));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ev_[i]));
    int dev = 0; CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev));
    gen_ = 1;
    // two instantiations of the expansion kernel: staging for 16 columns (batches whose longest partition has at
    // most 16 segments; 4 blocks per SM) and for 32
    expand_smem_[0] = sizeof(GroupShared<16>) * WPB; expand_smem_[1] = sizeof(GroupShared<32>) * WPB;
    CK(cudaFuncSetAttribute(k_expand<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[0]));
    CK(cudaFuncSetAttribute(k_expand<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[1]));
    CK(cudaFuncSetAttribute(k_expand<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[0]));
    CK(cudaFuncSetAttribute(k_expand<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[1]));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_expand<16, true>, WPB * 32, expand_smem_[0]));
    expand_blocks_[0] = sms_ * (per_sm > 0 ? per_sm : 1);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_expand<32, true>, WPB * 32, expand_smem_[1]));
    expand_blocks_[1] = sms_ * (per_sm > 0 ? per_sm : 1);
    // the fused program kernel (small batches): a small resident grid -- its barrier is what a step costs
    CK(cudaFuncSetAttribute(k_program<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[0]));
    CK(cudaFuncSetAttribute(k_program<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[0]));
    CK(cudaFuncSetAttribute(k_program<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[1]));
    CK(cudaFuncSetAttribute(k_program<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_[1]));
    {
        int coop = 0, dev2 = 0; CK(cudaGetDevice(&dev2));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev2));
        fused_on_ = coop != 0;
        int bps = 2;
        if (const char *e = getenv("AMBI_FUSED_BPS")) bps = std::max(1, atoi(e));
        if (const char *e = getenv("AMBI_FUSED_MAX")) fused_max_count_ = atoi(e);
        if (const char *e = getenv("AMBI_FUSED_BIG")) fused_big_ = (uint32_t)atoll(e);
        for (int i = 0; i < 2; ++i) {
            int p2 = 1;
            if (i == 0) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, k_program<16, true>, WPB * 32, expand_smem_[0]));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, k_program<32, true>, WPB * 32, expand_smem_[1]));
            if (p2 < 1) fused_on_ = false;
            fused_blocks_[i] = sms_ * std::max(1, std::min(bps, p2));
            if (const char *e = getenv("AMBI_FUSED_BLOCKS")) fused_blocks_[i] = std::max(1, std::min(atoi(e), sms_ * p2));
        }
    }
    // warp-per-partition kernel for partitions of up to 16 segments (opt-in: AMBI_CFG_SMALL_KERNEL or AMBI_SMALL=1;
    // AMBI_SMALL_C: records per level and warp)
    small_on_ = use_small || getenv("AMBI_SMALL") != nullptr;
    int C = 256;
    if (const char *e = getenv("AMBI_SMALL_C")) C = std::max(32, atoi(e));
    if (small_on_) {
        if (small_configure(&small_cfg_, C, sms_)) { snprintf(err_, sizeof err_, "small_configure(%d) failed: %s", C, cudaGetErrorString(cudaGetLastError())); return -4; }
        CK(cudaMalloc(&d_arena_, small_cfg_.arena_words * 4));
    }
    CK(cudaMalloc(&d_stamp0_, 8));
    CK(cudaMallocHost(&h_flag_, 16));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreate(&ev_small_[i]));
    return 0;
}

int StageA::ensure_problem_arrays(int count)
{
    if (count <= prob_cap_) return 0;
    free_problem_arrays();
    int c = std::max(count, 64);
    CK(cudaMalloc(&d_probs_, (size_t)c * sizeof(DevProblem)));
    CK(cudaMalloc(&d_limit_, (size_t)c * 4)); CK(cudaMalloc(&d_incumbent_, (size_t)c * 4));
    CK(cudaMalloc(&d_best_score_, (size_t)c * 4)); CK(cudaMalloc(&d_beam_score_, (size_t)c * 4));
    CK(cudaMalloc(&d_best_rows_, (size_t)c * AMBI_MAXN * 4)); CK(cudaMalloc(&d_beam_rows_, (size_t)c * AMBI_MAXN * 4));
    CK(cudaMalloc(&d_cand_, (size_t)c * 8)); CK(cudaMalloc(&d_node_, (size_t)c * 8));
    // beam statistics: one slice per level, cleared once per pass
    CK(cudaMalloc(&d_fmin_, (size_t)c * 4 * AMBI_MAXN)); CK(cudaMalloc(&d_hist_, (size_t)c * BEAM_BINS * 4));       // one level's histograms, cleared per beam level
    CK(cudaMalloc(&d_ticket_, (size_t)c * 4 * AMBI_MAXN));
    CK(cudaMalloc(&d_cutq_, (size_t)c * 4 * AMBI_MAXN));
    CK(cudaMalloc(&d_small_list_, (size_t)c * 4)); CK(cudaMalloc(&d_active_, (size_t)c)); CK(cudaMalloc(&d_active_base_, (size_t)c));
    CK(cudaMalloc(&d_seen_, (size_t)c * 4)); CK(cudaMalloc(&d_stamps_, (size_t)c * 32));
    CK(cudaMallocHost(&h_probs_, (size_t)c * sizeof(DevProblem)));
    CK(cudaMallocHost(&h_compact_, (size_t)c * compact_words(AMBI_MAXN, true) * 4)); CK(cudaMalloc(&d_compact_, (size_t)c * compact_words(AMBI_MAXN, true) * 4));
    CK(cudaMallocHost(&h_res_, (size_t)c * (4 + 4 + 8 + 8 + 4 * AMBI_MAXN + 32 + 1) + 64));
    h_cap_ = c;
    prob_cap_ = c;
    return 0;
}

void StageA::drop_graphs()
{
    prog_valid_ = false;
    if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
    if (graph_exec2_) { cudaGraphExecDestroy(graph_exec2_); graph_exec2_ = nullptr; }
}

uint32_t *StageA::compact_staging(int count)
{
    if (ensure_problem_arrays(count)) return nullptr;
    return h_compact_;
}

// one warp per problem, lane = segment: compact words -> DevProblem (coalesced 128-byte stores)
__global__ void k_unpack(const uint32_t *compact, int words, int nmax, int any_tgs, DevProblem *out, int count)
{
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= count) return;
    const uint32_t *c = compact + (size_t)p * words;
    DevProblem &d = out[p];
    if (lane == 0) {
        const uint32_t h = c[0];
        d.n = (int32_t)(h & 0xffu); d.lmax = (int32_t)((h >> 8) & 0xffu); d.ush = (int32_t)((h >> 16) & 0xffu); d.has_tgs = (int32_t)(h >> 24);
        d.loop_ub = (int32_t)c[1]; d.zero_score = (int32_t)c[2]; d.root_lb = (int32_t)c[3]; d.pad = 0;
    }
    const uint32_t cj = lane < nmax ? c[4 + lane] : 0u;
    d.c2[lane] = (int32_t)(cj & 0xffffu); d.j2[lane] = (int32_t)(cj >> 16); d.cj[lane] = cj;
    d.tgs_mask[lane] = (any_tgs && lane < nmax) ? c[4 + nmax + lane] : 0u;
}

int StageA::unpack_compact(int count, int nmax, bool any_tgs)
{
    const int wpb = 8;
    k_unpack<<<(count + wpb - 1) / wpb, wpb * 32, 0, stream_>>>(d_compact_, compact_words(nmax, any_tgs), nmax, any_tgs ? 1 : 0, d_probs_, count);
    CK(cudaGetLastError());
    return 0;
}

DevProblem *StageA::staging(int count)
{
    if (ensure_problem_arrays(count)) return nullptr;
    return h_probs_;
}

int StageA::upload(const DevProblem *host_probs, int count, bool copied, const int32_t *n_arr, const int8_t *ush_arr)
{
    if (int rc = ensure_problem_arrays(count)) return rc;
    if (count != count_) drop_graphs();
    count_ = count;
    // one pass over the staged problems: longest partition, finer units anywhere, levels at which a partition ends
    int nmax = 0;
    bool scaled = false;
    uint32_t leaf = 0;
    for (int i = 0; i < count; ++i) {
        const int n = n_arr ? n_arr[i] : host_probs[i].n;
        nmax = std::max(nmax, n);
        scaled = scaled || (ush_arr ? ush_arr[i] : host_probs[i].ush) != 0;
        if (n >= 2) leaf |= 1u << (n - 1);
    }
    if (nmax != nmax_) drop_graphs();
    nmax_ = nmax;
    if (scaled != scaled_) drop_graphs();
    scaled_ = scaled;
    if (leaf != leaf_levels_) drop_graphs();
    leaf_levels_ = leaf;
    stride_ = (int)ambi_node_stride(nmax_);
    kstride_ = nmax_ <= 16 ? 32 : 48;          // key record: header + 16 or 32 column bytes
    cap_ = (uint32_t)std::min<uint64_t>(cap_max_, buf_bytes_ / stride_);
    if ((uint64_t)count > cap_) { snprintf(err_, sizeof err_, "batch of %d partitions exceeds node capacity %u", count, cap_); return -5; }
    if (!copied) CK(cudaMemcpyAsync(d_probs_, host_probs, (size_t)count * sizeof(DevProblem), cudaMemcpyHostToDevice, stream_));
    {   // partitions of 2..16 segments go to the warp-per-partition kernel first; the others are active for the
        // global-memory path from the start (single-segment partitions are solved in closed form)
        std::vector<uint32_t> lst; std::vector<uint8_t> act((size_t)count, 0);
        n_large_ = 0;
        for (int i = 0; i < count; ++i) {
            const int n = n_arr ? n_arr[i] : host_probs[i].n;
            if (small_on_ && n >= 2 && n <= 16 && (ush_arr ? ush_arr[i] : host_probs[i].ush) == 0) lst.push_back((uint32_t)i);
            else if (n >= 2) { act[i] = 1; ++n_large_; }
        }
        n_small_ = (int)lst.size();
        // heaviest first would be better still; longest partitions first keeps the tail of the launch short
        std::stable_sort(lst.begin(), lst.end(), [&](uint32_t x, uint32_t y) { return n_arr ? n_arr[x] > n_arr[y] : host_probs[x].n > host_probs[y].n; });
        if (n_small_) CK(cudaMemcpyAsync(d_small_list_, lst.data(), (size_t)n_small_ * 4, cudaMemcpyHostToDevice, stream_));
        CK(cudaMemcpyAsync(d_active_base_, act.data(), (size_t)count, cudaMemcpyHostToDevice, stream_));
        CK(cudaStreamSynchronize(stream_));      // the host vectors go out of scope
    }
    return 0;
}

// ---- operations: recorded into a program (rec_ set) or launched at once -------------------------------------
int StageA::op(const Op &o)
{
    if (rec_) { rec_->push_back(o); return 0; }
    return launch_op(o);
}

int StageA::launch_op(const Op &o)
{
    const int tpb = 128, pblocks = (count_ + tpb - 1) / tpb;
    switch (o.kind) {
    case OP_MEMSET32: CK(cudaMemsetAsync(o.u.ms.ptr, (int)(o.u.ms.value & 0xffu), (size_t)o.u.ms.words * 4, stream_)); return 0;   // 0 or 0x7f7f7f7f
    case OP_INIT_ROOTS: k_init_roots<<<pblocks, tpb, 0, stream_>>>(o.u.ir); break;
    case OP_EXPAND: launch_expand(o.u.ex); return 0;
    case OP_DEDUP: k_dedup<<<sms_ * 8, 256, 0, stream_>>>(o.u.dd); break;
    case OP_FINALIZE: k_finalize<<<sms_ * 8, 256, 0, stream_>>>(o.u.fz); break;
    case OP_BEAM_HIST: k_beam_hist<<<sms_ * 8, 256, 0, stream_>>>(o.u.bm); break;
    case OP_BEAM_CUT: k_beam_cut<<<pblocks, tpb, 0, stream_>>>(o.u.bm); break;
    case OP_MATERIALISE: k_materialise<<<sms_ * 8, 256, 0, stream_>>>(o.u.mt); break;
    case OP_SET_LIMIT: k_set_limit<<<pblocks, tpb, 0, stream_>>>(o.u.sl); break;
    case OP_BUMP_GEN: k_bump_gen<<<1, 1, 0, stream_>>>(o.u.gen); break;
    }
    ++launches_;
    return 0;
}

int StageA::op_memset(void *ptr, int byte, size_t bytes, bool sync_after)
{
    Op o; memset(&o, 0, sizeof o);
    o.kind = OP_MEMSET32; o.sync_after = sync_after ? 1 : 0;
    o.u.ms.ptr = static_cast<uint32_t *>(ptr); o.u.ms.words = bytes / 4; o.u.ms.value = 0x01010101u * (uint32_t)(byte & 0xff);
    return op(o);
}

// The whole search as one cooperative kernel (k_program) for small batches; what it leaves (a frontier too wide for
// its small grid) is played as ordinary launches.  [first, last) index the recorded program.
int StageA::run_fused(int first, int last)
{
    const int idx = nmax_ <= 16 ? 0 : 1;
    const Op *d_ops = d_prog_;
    uint32_t *d_stop = counters_ + CTR_STOP;
    uint32_t big = fused_big_;
    void *args[] = {(void *)&d_ops, (void *)&first, (void *)&last, (void *)&big, (void *)&d_stop};
    const void *fn = nmax_ <= 16 ? (scaled_ ? (const void *)k_program<16, true> : (const void *)k_program<16, false>)
                                 : (scaled_ ? (const void *)k_program<32, true> : (const void *)k_program<32, false>);
    CK(cudaLaunchCooperativeKernel(fn, dim3(fused_blocks_[idx]), dim3(WPB * 32), args, expand_smem_[idx], stream_));
    ++launches_;
    CK(cudaMemcpyAsync(h_flag_ + 1, d_stop, 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    const int stopped = (int)h_flag_[1];
    fused_stops_ += stopped < last;
    for (int i = stopped; i < last; ++i) if (int rc = launch_op(prog_[i])) return rc;
    return 0;
}

void StageA::launch_expand(const LevelArgs &L)
{
    // four instantiations: 16 / 32 staged columns x copy numbers in half copies / in finer units (only when some
    // partition of the batch asks for one: the half-copy kernels carry no scaling arithmetic)
    if (nmax_ <= 16) {
        if (scaled_) k_expand<16, true><<<expand_blocks_[0], WPB * 32, expand_smem_[0], stream_>>>(L);
        else k_expand<16, false><<<expand_blocks_[0], WPB * 32, expand_smem_[0], stream_>>>(L);
    } else {
        if (scaled_) k_expand<32, true><<<expand_blocks_[1], WPB * 32, expand_smem_[1], stream_>>>(L);
        else k_expand<32, false><<<expand_blocks_[1], WPB * 32, expand_smem_[1], stream_>>>(L);
    }
    ++launches_;
}

// One level: expand -> (merge) -> finalize (-> beam selection).
//   mode 0: greedy incumbent (every record emits its arg-min child)
//   mode 1: beam incumbent (children within the limit, top-W per problem kept)
//   mode 2: exact (children within the limit, merged by column state)
int StageA::run_level(int level, int cur, int mode, int shard_world)
{
    uint32_t *c_over = counters_ + CTR_OVERFLOW;
    uint32_t *c_work = counters_ + CTR_WORK + level, *c_child = counters_ + CTR_CHILD + level;
    uint32_t *c_cur = counters_ + CTR_FRONT + level, *c_nxt = counters_ + CTR_FRONT + level + 1;
    const bool merge = mode != 0;
    LevelArgs L;
    // the frontier of a level is a dense array of full records (k_materialise); its children are key records + links
    L.probs = d_probs_; L.parents = buf_[cur]; L.plist = nullptr; L.pcount = c_cur;
    L.children = kbuf_; L.links = lbuf_; L.dead = dbuf_; L.kstride = kstride_; L.ccount = c_child; L.cap = cap_; L.limit = d_limit_;
    L.cand_count = d_cand_; L.node_count = d_node_; L.funnel = d_funnel_; L.work_counter = c_work; L.overflow = c_over;
    L.level = level; L.stride = stride_; L.nmax = nmax_; L.emit = mode == 0 ? EMIT_BEST : EMIT_ALL;
    L.beam_delta = mode == 1 ? beam_delta_ : -1; L.beam_shift = beam_shift_;
    L.beam_delta_short = beam_delta_short_; L.beam_shift_short = beam_shift_short_; L.beam_short_n = beam_short_n_;
    L.shard_world = shard_world; L.shard_rank = rank_; L.shard_level = shard_level_eff_; L.gpw = gpw_;
    const bool prof = profile_ && !capturing_;
    if (prof) {
        if (prof_used_ + 2 > prof_ev_.size()) { cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); prof_ev_.push_back(e0); prof_ev_.push_back(e1); }
        CK(cudaEventRecord(prof_ev_[prof_used_], stream_));
    }
    { Op o; memset(&o, 0, sizeof o); o.kind = OP_EXPAND; o.sync_after = 1; o.u.ex = L; if (int rc = op(o)) return rc; }
    if (prof) { CK(cudaEventRecord(prof_ev_[prof_used_ + 1], stream_)); prof_used_ += 2; }
    const bool beam = mode == 1;
    int32_t *fmin_cur = d_fmin_ + (size_t)level * count_;
    uint32_t *hist = d_hist_;
    if (beam) { if (int rc = op_memset(d_hist_, 0, (size_t)count_ * BEAM_BINS * 4, false)) return rc; }     // used after the merge's barrier
    if (merge) {
        DedupArgs D;
        D.recs = kbuf_; D.dead = dbuf_; D.tshift = dedup_shift_; D.count = c_child; D.cap = cap_; D.table = table_; D.tmask = tsize_ - 1;
        D.gen_base = counters_ + CTR_GEN; D.gen_off = (uint32_t)(mode * 40 + level); D.stride = kstride_; D.level = level; D.overflow = c_over;
        D.probs = d_probs_; D.fmin_cur = beam ? fmin_cur : nullptr; D.nmax = nmax_;
        D.links = lbuf_; D.parents = buf_[cur]; D.pstride = stride_;
        { Op o; memset(&o, 0, sizeof o); o.kind = OP_DEDUP; o.sync_after = 1; o.u.dd = D; if (int rc = op(o)) return rc; }
    }
    FinalizeArgs F;
    F.probs = d_probs_; F.recs = kbuf_; F.dead = dbuf_; F.count = c_child; F.cap = cap_;
    F.best_score = mode == 2 ? d_best_score_ : d_beam_score_; F.best_rows = mode == 2 ? d_best_rows_ : d_beam_rows_;
    F.stride = kstride_; F.level = level;
    F.limit = nullptr; F.merge_best = 0; F.overflow = c_over; F.seen = d_seen_; F.stamps = d_stamps_;
    F.links = lbuf_; F.parents = buf_[cur]; F.pstride = stride_;
    // this kernel only collects completed assignments: skip the levels where no problem ends
    // (it reads what the merge left and writes per-problem results only: no barrier needed before the next operation)
    if ((leaf_levels_ >> level) & 1u) { Op o; memset(&o, 0, sizeof o); o.kind = OP_FINALIZE; o.sync_after = 0; o.u.fz = F; if (int rc = op(o)) return rc; }
    MatArgs M;
    M.probs = d_probs_; M.keys = kbuf_; M.dead = dbuf_; M.links = lbuf_; M.count = c_child; M.cap = cap_; M.parents = buf_[cur]; M.pstride = stride_;
    M.out = buf_[cur ^ 1]; M.out_count = c_nxt; M.out_cap = cap_; M.kstride = kstride_; M.level = level; M.nmax = nmax_; M.overflow = c_over;
    M.fmin = nullptr; M.cutq = nullptr; M.ticket = nullptr;
    if (beam) {
        BeamArgs B;
        B.probs = d_probs_; B.recs = kbuf_; B.dead = dbuf_; B.count = c_child; B.cap = cap_;
        B.fmin = fmin_cur; B.hist = hist; B.cutq = d_cutq_ + (size_t)level * count_;
        B.ticket = d_ticket_ + (size_t)level * count_;
        B.nprob = count_; B.width = beam_; B.width_floor = std::min(beam_floor_max_, std::max(beam_, beam_ * 4096 / std::max(count_, 1)));
        B.stride = kstride_; B.level = level; B.growth = beam_growth_; B.width_div = beam_div_;
        { Op o; memset(&o, 0, sizeof o); o.kind = OP_BEAM_HIST; o.sync_after = 1; o.u.bm = B; if (int rc = op(o)) return rc; }
        { Op o; memset(&o, 0, sizeof o); o.kind = OP_BEAM_CUT; o.sync_after = 1; o.u.bm = B; if (int rc = op(o)) return rc; }
        M.fmin = fmin_cur; M.cutq = B.cutq; M.ticket = B.ticket;          // the selection itself happens while materialising
    }
    if (level + 1 < nmax_) { Op o; memset(&o, 0, sizeof o); o.kind = OP_MATERIALISE; o.sync_after = 1; o.u.mt = M; if (int rc = op(o)) return rc; }
    else if (rec_ && !rec_->empty()) rec_->back().sync_after = 1;      // the pass ends here
    return 0;
}

int StageA::run_pass(int mode, int shard_world, bool first)
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    // everything but the overflow flag
    (void)blocks;
    if (mode == 1) {        // (independent clears share the barrier of the last one)
        if (int rc = op_memset(d_fmin_, 0x7f, (size_t)count_ * 4 * nmax_, false)) return rc;
        if (int rc = op_memset(d_ticket_, 0, (size_t)count_ * 4 * nmax_, false)) return rc;
    }
    if (int rc = op_memset(counters_ + CTR_WORK, 0, (CTR_TOTAL - CTR_WORK) * sizeof(uint32_t))) return rc;
    {
        Op o; memset(&o, 0, sizeof o); o.kind = OP_INIT_ROOTS; o.sync_after = 1;
        o.u.ir.probs = d_probs_; o.u.ir.count = count_; o.u.ir.recs = buf_[0]; o.u.ir.stride = stride_; o.u.ir.nmax = nmax_; o.u.ir.list = list_[0];
        o.u.ir.lcount = counters_ + CTR_FRONT; o.u.ir.limit = d_limit_; o.u.ir.best_score = d_best_score_; o.u.ir.first = first ? 1 : 0; o.u.ir.active = d_active_;
        if (int rc = op(o)) return rc;
    }
    int cur = 0;
    for (int a = 0; a < nmax_; ++a) { if (int rc = run_level(a, cur, mode, shard_world)) return rc; cur ^= 1; }
    return 0;
}

// Incumbent part of a search: greedy dive, then the beam.  `coop` (prefix-sharded with a communicator): the
// beam is sharded by prefix like the exact pass -- every rank keeps the children of row `shard_level` that hash
// to it and runs a beam of 1/world of the width below them; the ranks' incumbents are min-reduced over NCCL
// before the exact part starts (run()).
int StageA::enqueue_incumbent(bool coop)
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    (void)blocks;
    if (int rc = op_memset(counters_ + CTR_OVERFLOW, 0, sizeof(uint32_t), false)) return rc;
    { Op o; memset(&o, 0, sizeof o); o.kind = OP_BUMP_GEN; o.sync_after = 0; o.u.gen = counters_ + CTR_GEN; if (int rc = op(o)) return rc; }
    if (int rc = op_memset(d_beam_score_, 0x7f, (size_t)count_ * 4, false)) return rc;      // (run_pass ends its clears with a barrier)
    shard_level_eff_ = -1;
    // ---- pass 0: greedy dive -> first incumbent (limit starts at the all-zero score)
    if (int rc = run_pass(0, 1, true)) return rc;
    { Op o; memset(&o, 0, sizeof o); o.kind = OP_SET_LIMIT; o.sync_after = 1; o.u.sl.probs = d_probs_; o.u.sl.count = count_; o.u.sl.limit = d_limit_;
      o.u.sl.best_score = d_beam_score_; o.u.sl.incumbent = d_incumbent_; o.u.sl.active = d_active_; o.u.sl.seen = d_seen_; o.u.sl.stamps = d_stamps_;
      if (int rc = op(o)) return rc; }
    // ---- pass 1: beam of width W inside that limit -> tighter incumbent
    if (beam_ > 1) {
        if (int rc = op_memset(d_beam_score_, 0x7f, (size_t)count_ * 4, false)) return rc;
        beam_div_ = 1;
        if (coop) {
            int sl = shard_level_ > 0 ? shard_level_ : 2;
            shard_level_eff_ = std::min(sl, std::max(0, nmax_ - 2));
            beam_div_ = world_;
        }
        if (int rc = run_pass(1, coop ? world_ : 1, false)) return rc;
        beam_div_ = 1;
    }
    return 0;
}

// Exact part: the limit becomes the best incumbent, then the exact pass (prefix-sharded when `sharded`).
int StageA::enqueue_exact(bool sharded)
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    (void)blocks;
    if (beam_ > 1) { Op o; memset(&o, 0, sizeof o); o.kind = OP_SET_LIMIT; o.sync_after = 1; o.u.sl.probs = d_probs_; o.u.sl.count = count_; o.u.sl.limit = d_limit_;
      o.u.sl.best_score = d_beam_score_; o.u.sl.incumbent = d_incumbent_; o.u.sl.active = d_active_; o.u.sl.seen = d_seen_; o.u.sl.stamps = d_stamps_;
      if (int rc = op(o)) return rc; }
    if (!capturing_) CK(cudaEventRecord(ev_[1], stream_));
    int sw = sharded ? world_ : 1;
    shard_level_eff_ = -1;
    if (sw > 1) {
        int sl = shard_level_ > 0 ? shard_level_ : 2;
        shard_level_eff_ = std::min(sl, std::max(0, nmax_ - 2));
    }
    if (int rc = run_pass(2, sw, false)) return rc;
    return 0;
}

// The launch sequence depends only on (count, nmax, sharding): every size that matters lives in
// device counters.  It is captured once into a CUDA graph and replayed, so a step costs one
// graph launch instead of ~150 kernel launches; the profiling pass uses plain launches.  A cooperative
// (prefix-sharded, NCCL) search is two graphs with the all-reduce(min) of the incumbents in between.
int StageA::capture(cudaGraphExec_t *exec, unsigned long long *nlaunch, int part, bool flag)
{
    cudaGraph_t g = nullptr;
    const unsigned long long l0 = launches_;
    CK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    capturing_ = true;
    int rc = 0;
    if (part == 0) { rc = enqueue_incumbent(false); if (!rc) rc = enqueue_exact(flag); }      // everything, one graph
    else if (part == 1) rc = enqueue_incumbent(true);
    else rc = enqueue_exact(true);
    capturing_ = false;
    cudaError_t e = cudaStreamEndCapture(stream_, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    CK(e);
    CK(cudaGraphInstantiate(exec, g, 0));
    cudaGraphDestroy(g);
    *nlaunch = launches_ - l0;
    launches_ = l0;
    return 0;
}

// Start of every search (plain launches, not part of the replayed graph): results and statistics cleared,
// single-segment partitions solved in closed form, the reference time stamp taken.
int StageA::prepare()
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    k_prepare<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, d_best_score_, d_seen_, d_incumbent_, d_stamps_, d_cand_, d_node_, d_active_, d_active_base_); ++launches_;
    CK(cudaMemsetAsync(d_funnel_, 0, 4 * sizeof(unsigned long long), stream_));
    CK(cudaMemsetAsync(counters_ + CTR_SMALL_WORK, 0, 2 * sizeof(uint32_t), stream_));
    CK(cudaMemsetAsync(counters_ + CTR_OVERFLOW, 0, sizeof(uint32_t), stream_));
    k_solve_n1<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, d_best_score_, d_best_rows_, d_cand_); ++launches_;
    stamp_launch(d_stamp0_, stream_); ++launches_;
    return 0;
}

int StageA::run(bool sharded)
{
    prof_used_ = 0;
    last_sharded_ = sharded;
    CK(cudaEventRecord(ev_[0], stream_));
    if (int rc = prepare()) return rc;
    // ---- partitions of up to 16 segments: one warp each, the whole search in shared memory (stage_small.cu).
    // Whatever does not fit there is flagged active and joins the longer partitions on the global-memory path.
    global_count_ = n_large_; small_solved_ = 0; ran_global_ = false;
    h_flag_[0] = 0;
    if (small_on_ && n_small_ > 0) {
        SmallLaunch L;
        L.probs = d_probs_; L.plist = d_small_list_; L.count = (uint32_t)n_small_; L.work = counters_ + CTR_SMALL_WORK;
        L.best_score = d_best_score_; L.best_rows = d_best_rows_; L.incumbent = d_incumbent_;
        L.cand = d_cand_; L.node = d_node_; L.funnel = d_funnel_; L.fb_count = counters_ + CTR_SMALL_FB; L.active = d_active_;
        L.arena = d_arena_; L.stamps = d_stamps_; L.beam_width = beam_; L.beam_growth = beam_growth_; L.beam_delta = beam_delta_; L.beam_shift = beam_shift_;
        CK(cudaEventRecord(ev_small_[0], stream_));
        small_launch(small_cfg_, L, stream_); ++launches_;
        CK(cudaEventRecord(ev_small_[1], stream_));
        CK(cudaGetLastError());
        // how many partitions were handed on decides whether the global-memory path runs at all (one host round trip)
        CK(cudaMemcpyAsync(h_flag_, counters_ + CTR_SMALL_FB, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        global_count_ += (int)h_flag_[0];
        small_solved_ = n_small_ - (int)h_flag_[0];
    }
    if (global_count_ == 0) { used_graph_ = false; CK(cudaEventRecord(ev_[1], stream_)); CK(cudaEventRecord(ev_[2], stream_)); return 0; }
    ran_global_ = true;
    const bool coop = sharded && world_ > 1 && allreduce_ != nullptr && beam_ > 1 && getenv("AMBI_REPLICATED_BEAM") == nullptr;
    const bool use_graph = !profile_ && getenv("AMBI_NO_GRAPH") == nullptr;
    const int mode = coop ? 2 : (sharded ? 1 : 0);
    if (use_graph && graph_exec_ && graph_mode_ != mode) {
        cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr;
        if (graph_exec2_) { cudaGraphExecDestroy(graph_exec2_); graph_exec2_ = nullptr; }
    }
    // small batches: the whole launch sequence as one cooperative kernel (latency, not work, is what they cost)
    const bool fused = fused_on_ && !profile_ && global_count_ <= fused_max_count_ && getenv("AMBI_NO_FUSED") == nullptr;
    if (fused) {
        if (!prog_valid_ || prog_mode_ != mode) {
            prog_.clear();
            rec_ = &prog_; capturing_ = true;
            int rc = enqueue_incumbent(coop);
            prog_split_ = (int)prog_.size();
            if (!rc) rc = enqueue_exact(sharded);
            rec_ = nullptr; capturing_ = false;
            if (rc) return rc;
            if (prog_.size() > d_prog_cap_) {
                cudaFree(d_prog_); d_prog_ = nullptr; d_prog_cap_ = 0;
                CK(cudaMalloc(&d_prog_, prog_.size() * sizeof(Op)));
                d_prog_cap_ = prog_.size();
            }
            CK(cudaMemcpyAsync(d_prog_, prog_.data(), prog_.size() * sizeof(Op), cudaMemcpyHostToDevice, stream_));
            CK(cudaStreamSynchronize(stream_));
            prog_valid_ = true; prog_mode_ = mode;
        }
        if (!coop) { if (int rc = run_fused(0, (int)prog_.size())) return rc; }
        else {
            if (int rc = run_fused(0, prog_split_)) return rc;
            if (allreduce_(allreduce_user_, d_beam_score_, count_, 0)) { snprintf(err_, sizeof err_, "incumbent all-reduce(min) failed"); return -4; }
            if (int rc = run_fused(prog_split_, (int)prog_.size())) return rc;
        }
        used_graph_ = true;                 // (no per-pass events in this mode either)
        CK(cudaEventRecord(ev_[2], stream_));
        return 0;
    }
    if (!coop) {
        if (!use_graph) {
            if (int rc = enqueue_incumbent(false)) return rc;
            if (int rc = enqueue_exact(sharded)) return rc;
        } else {
            if (!graph_exec_) { if (int rc = capture(&graph_exec_, &graph_launches_, 0, sharded)) return rc; graph_mode_ = mode; }
            CK(cudaGraphLaunch(graph_exec_, stream_));
            launches_ += graph_launches_;
        }
    } else {
        if (!use_graph) { if (int rc = enqueue_incumbent(true)) return rc; }
        else {
            if (!graph_exec_) { if (int rc = capture(&graph_exec_, &graph_launches_, 1, true)) return rc; graph_mode_ = mode; }
            CK(cudaGraphLaunch(graph_exec_, stream_));
            launches_ += graph_launches_;
        }
        // the ranks' beams covered disjoint prefixes: everybody continues with the best incumbent found anywhere
        if (allreduce_(allreduce_user_, d_beam_score_, count_, 0)) { snprintf(err_, sizeof err_, "incumbent all-reduce(min) failed"); return -4; }
        if (!use_graph) { if (int rc = enqueue_exact(true)) return rc; }
        else {
            if (!graph_exec2_) { if (int rc = capture(&graph_exec2_, &graph_launches2_, 2, true)) return rc; }
            CK(cudaGraphLaunch(graph_exec2_, stream_));
            launches_ += graph_launches2_;
        }
    }
    used_graph_ = use_graph;
    CK(cudaEventRecord(ev_[2], stream_));
    return 0;
}


// ---------------------------------------------------------------------------
// chunked (depth-first) exact pass: the fallback when one level of the breadth-first
// pass does not fit the candidate buffers.  Frontiers live in a pool used as a stack of
// levels; each level is expanded a chunk at a time, all the way down before the next
// chunk, so completed assignments arrive early and tighten the limit for what is left.
// ---------------------------------------------------------------------------
__global__ void k_iota(uint32_t *list, uint32_t start, uint32_t m, uint32_t *count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) list[i] = start + i;
    if (i == 0) *count = m;
}

__global__ void k_fold_best_into_limit(int count, int32_t *limit, const int32_t *best_score)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < count && best_score[p] < limit[p]) limit[p] = best_score[p];
}

// One round of the bound exchange between the ranks of a prefix-sharded search: every rank's pruning limit
// becomes the minimum over ranks (a completed assignment found anywhere prunes everywhere; the limit test is
// `<=`, so equal-score assignments survive and the tie-break winner does not depend on who found what), and
// the ranks learn how many of them are still searching.
int StageA::exchange_round(int active, int *total_active)
{
    int32_t a = active;
    if (allreduce_(allreduce_user_, d_limit_, count_, 0)) { snprintf(err_, sizeof err_, "bound exchange: all-reduce(min) failed"); return -4; }
    CK(cudaMemcpyAsync(d_flag_, &a, 4, cudaMemcpyHostToDevice, stream_));
    if (allreduce_(allreduce_user_, d_flag_, 1, 2)) { snprintf(err_, sizeof err_, "bound exchange: all-reduce(sum) failed"); return -4; }
    CK(cudaMemcpyAsync(&a, d_flag_, 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    *total_active = a;
    ++exchange_rounds_;
    return 0;
}

int StageA::run_chunked(bool sharded, int start_level, bool exchange)
{
    const int rc = run_chunked_core(sharded, start_level, exchange);
    if (exchange) {         // done (or failed) here: keep answering the other ranks' rounds until nobody is searching
        int total = 1;
        while (total > 0) if (exchange_round(0, &total)) return -4;
    }
    return rc;
}

int StageA::run_chunked_core(bool sharded, int start_level, bool exchange)
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    if (!pool_) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        pool_bytes_ = std::min<size_t>(free_b / 2, 8 * buf_bytes_);
        if (const char *e = getenv("AMBI_POOL_BYTES")) { long long v = atoll(e); if (v > 0) pool_bytes_ = std::min<size_t>(pool_bytes_, (size_t)v); }   // test knob: force "pool exhausted"
        CK(cudaMalloc(&pool_, pool_bytes_));
    }
    int sw = sharded ? world_ : 1;
    shard_level_eff_ = -1;
    if (sw > 1) { int sl = shard_level_ > 0 ? shard_level_ : 2; shard_level_eff_ = std::min(sl, std::max(0, nmax_ - 2)); }
    // the limit (incumbent from passes 0/1) and the n==1 results of the failed run are still valid
    CK(cudaMemsetAsync(counters_ + CTR_OVERFLOW, 0, sizeof(uint32_t), stream_));
    struct Seg { int level; size_t off; uint32_t cnt, pos; };
    std::vector<Seg> stack;
    uint32_t nroots = 0;
    if (start_level >= 1) {
        // continue from the last level the breadth-first pass completed: its frontier (dense full records) is still
        // intact in the parent buffer of the level that overflowed; every per-problem result found so far stays valid
        CK(cudaMemcpyAsync(&nroots, counters_ + CTR_FRONT + start_level, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        nroots = std::min(nroots, cap_);
        if ((size_t)nroots * stride_ > pool_bytes_) { snprintf(err_, sizeof err_, "chunked pass: frontier pool too small"); return -5; }
        CK(cudaMemcpyAsync(pool_, buf_[start_level & 1], (size_t)nroots * stride_, cudaMemcpyDeviceToDevice, stream_));
        CK(cudaMemsetAsync(counters_ + CTR_WORK, 0, (CTR_TOTAL - CTR_WORK) * sizeof(uint32_t), stream_));
        if (nroots > 0) stack.push_back({start_level, 0, nroots, 0});
    } else {
        // roots -> pool
        CK(cudaMemsetAsync(counters_ + CTR_WORK, 0, (CTR_TOTAL - CTR_WORK) * sizeof(uint32_t), stream_));
        {
            InitRootsArgs ir; ir.probs = d_probs_; ir.count = count_; ir.recs = pool_; ir.stride = stride_; ir.nmax = nmax_; ir.list = list_[0];
            ir.lcount = counters_ + CTR_FRONT; ir.limit = d_limit_; ir.best_score = d_best_score_; ir.first = 0; ir.active = d_active_;
            k_init_roots<<<blocks, tpb, 0, stream_>>>(ir); ++launches_;
        }
        CK(cudaMemcpyAsync(&nroots, counters_ + CTR_FRONT, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        if (nroots > 0) stack.push_back({0, 0, nroots, 0});
    }
    size_t pool_top = (size_t)nroots * stride_;
    std::vector<uint32_t> chunk(AMBI_MAXN + 1, cap_ / 16 ? cap_ / 16 : 1);   // parents per chunk, per level (adaptive)
    uint32_t *c_over = counters_ + CTR_OVERFLOW;
    uint32_t *c_work = counters_ + CTR_WORK, *c_child = counters_ + CTR_CHILD;
    uint32_t *c_par = counters_ + CTR_FRONT, *c_next = counters_ + CTR_FRONT + 1;
    unsigned long long chunks = 0, since_exchange = 0;
    const unsigned long long exchange_every = 8;
    while (!stack.empty()) {
        if (exchange && since_exchange >= exchange_every) {
            int total = 0;
            if (int rc = exchange_round(1, &total)) return rc;
            since_exchange = 0;
        }
        Seg &sg = stack.back();
        if (sg.pos >= sg.cnt) { pool_top = sg.off; stack.pop_back(); continue; }
        const int level = sg.level;
        const uint32_t m = std::min<uint32_t>(chunk[level], sg.cnt - sg.pos);
        CK(cudaMemsetAsync(c_work, 0, 4, stream_)); CK(cudaMemsetAsync(c_child, 0, 4, stream_));
        CK(cudaMemsetAsync(c_next, 0, 4, stream_)); CK(cudaMemsetAsync(c_over, 0, 4, stream_));
        k_iota<<<(m + 255) / 256, 256, 0, stream_>>>(list_[0], sg.pos, m, c_par); ++launches_;
        LevelArgs L;
        L.probs = d_probs_; L.parents = pool_ + sg.off; L.plist = list_[0]; L.pcount = c_par;
        L.children = kbuf_; L.links = lbuf_; L.dead = dbuf_; L.kstride = kstride_; L.ccount = c_child; L.cap = cap_; L.limit = d_limit_;
        L.cand_count = d_cand_; L.node_count = d_node_; L.funnel = d_funnel_; L.work_counter = c_work; L.overflow = c_over;
        L.level = level; L.stride = stride_; L.nmax = nmax_; L.emit = EMIT_ALL; L.beam_delta = -1; L.beam_shift = 1; L.beam_delta_short = 0; L.beam_shift_short = 1; L.beam_short_n = 0;
        L.shard_world = sw; L.shard_rank = rank_; L.shard_level = shard_level_eff_; L.gpw = gpw_;
        launch_expand(L);
        uint32_t ctr[2] = {0, 0};
        CK(cudaMemcpyAsync(&ctr[0], c_over, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaMemcpyAsync(&ctr[1], c_child, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        if (ctr[0]) {                               // this chunk's children do not fit: halve and retry
            if (m == 1) { snprintf(err_, sizeof err_, "chunked pass: one record's children exceed the buffer (%u)", cap_); return -5; }
            chunk[level] = std::max<uint32_t>(1, m / 2);
            // candidates of the aborted attempt were scored; they stay in the statistics
            continue;
        }
        if (ctr[1] * 4ull < cap_ && chunk[level] < cap_) chunk[level] = std::min<uint64_t>(cap_, 2ull * chunk[level]);
        sg.pos += m;
        ++chunks; ++since_exchange;
        DedupArgs D;
        D.recs = kbuf_; D.dead = dbuf_; D.tshift = dedup_shift_; D.count = c_child; D.cap = cap_; D.table = table_; D.tmask = tsize_ - 1;
        D.gen_base = counters_ + CTR_GEN; D.gen_off = 0; D.stride = kstride_; D.level = level; D.overflow = c_over;
        D.probs = d_probs_; D.fmin_cur = nullptr; D.nmax = nmax_;
        D.links = lbuf_; D.parents = pool_ + sg.off; D.pstride = stride_;
        k_bump_gen<<<1, 1, 0, stream_>>>(counters_ + CTR_GEN); ++launches_;
        k_dedup<<<sms_ * 8, 256, 0, stream_>>>(D); ++launches_;
        FinalizeArgs F;
        F.probs = d_probs_; F.recs = kbuf_; F.dead = dbuf_; F.count = c_child; F.cap = cap_;
        F.best_score = d_best_score_; F.best_rows = d_best_rows_;
        F.limit = d_limit_; F.stride = kstride_; F.level = level; F.merge_best = 1; F.overflow = c_over; F.seen = d_seen_; F.stamps = d_stamps_;
        F.links = lbuf_; F.parents = pool_ + sg.off; F.pstride = stride_;
        k_finalize<<<sms_ * 8, 256, 0, stream_>>>(F); ++launches_;
        // the survivors become full records on top of the pool (the next, deeper segment of the stack)
        const uint64_t room = (pool_bytes_ - pool_top) / (size_t)stride_;
        MatArgs M;
        M.probs = d_probs_; M.keys = kbuf_; M.dead = dbuf_; M.links = lbuf_; M.count = c_child; M.cap = cap_; M.parents = pool_ + sg.off; M.pstride = stride_;
        M.out = pool_ + pool_top; M.out_count = c_next; M.out_cap = (uint32_t)std::min<uint64_t>(room, 0xffffffffu);
        M.kstride = kstride_; M.level = level; M.nmax = nmax_; M.overflow = c_over; M.fmin = nullptr; M.cutq = nullptr; M.ticket = nullptr;
        k_materialise<<<sms_ * 8, 256, 0, stream_>>>(M); ++launches_;
        uint32_t nnext = 0;
        CK(cudaMemcpyAsync(&nnext, c_next, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        if (nnext) {
            if (nnext > room) { snprintf(err_, sizeof err_, "chunked pass: frontier pool exhausted (%zu bytes)", pool_bytes_); return -5; }
            stack.push_back({level + 1, pool_top, nnext, 0});
            pool_top += (size_t)nnext * stride_;
        }
    }
    chunked_runs_++;
    (void)chunks;
    CK(cudaEventRecord(ev_[2], stream_));
    return 0;
}

__global__ void k_reset_best(const DevProblem *probs, int count, int32_t *best_score, const uint8_t *active)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < count && probs[p].n >= 2 && active[p]) best_score[p] = 0x7fffffff;
}

int StageA::fetch_raw(Fetched *F)
{
    chunked_failed_ = false;
    if (ran_global_) {   // a level of the breadth-first exact pass did not fit: continue depth-first in chunks
        uint32_t over = 0;
        CK(cudaMemcpyAsync(&over, counters_ + CTR_OVERFLOW, 4, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        const bool chunked_ok = getenv("AMBI_NO_CHUNKED") == nullptr;
        // prefix-sharded with a communicator: the ranks agree whether anybody goes depth-first; if so, all of
        // them join the periodic bound exchange (those whose shard fitted contribute their final best score)
        bool exchange = false;
        if (allreduce_ && last_sharded_ && world_ > 1 && chunked_ok) {
            int32_t f = over ? 1 : 0;
            CK(cudaMemcpyAsync(d_flag_, &f, 4, cudaMemcpyHostToDevice, stream_));
            if (allreduce_(allreduce_user_, d_flag_, 1, 1)) { snprintf(err_, sizeof err_, "all-reduce(max) failed"); return -4; }
            CK(cudaMemcpyAsync(&f, d_flag_, 4, cudaMemcpyDeviceToHost, stream_));
            CK(cudaStreamSynchronize(stream_));
            exchange = f != 0;
        }
        if (over && chunked_ok) {
            // `over` = 1 + the level of the exact pass that did not fit
            const int lvl = (int)over - 1;
            const bool resume = lvl >= 1 && getenv("AMBI_CHUNKED_FROM_ROOT") == nullptr;
            if (!resume) { k_reset_best<<<(count_ + 127) / 128, 128, 0, stream_>>>(d_probs_, count_, d_best_score_, d_active_); ++launches_; }
            // -5 = the depth-first pass itself ran out of room (frontier pool exhausted / too small, or one record's
            // children exceed a buffer): whatever partial best exists is NOT proven optimal, so the failure is
            // reported as a frontier overflow (run_chunked_core clears the device flag per chunk; this host flag
            // is what fetch reports)
            if (int rc = run_chunked(last_sharded_, resume ? lvl : 0, exchange)) { if (rc != -5) return rc; chunked_failed_ = true; }
        } else if (exchange) {
            k_fold_best_into_limit<<<(count_ + 127) / 128, 128, 0, stream_>>>(count_, d_limit_, d_best_score_); ++launches_;
            int total = 1;
            while (total > 0) if (int rc = exchange_round(0, &total)) return rc;
        }
    }
    // one pinned staging area: score | incumbent | cand | node | rows (nmax words per problem)
    const size_t c = (size_t)count_;
    int32_t *h_score = reinterpret_cast<int32_t *>(h_res_);
    int32_t *h_inc = h_score + h_cap_;
    unsigned long long *h_cand = reinterpret_cast<unsigned long long *>(h_inc + h_cap_);
    unsigned long long *h_node = h_cand + h_cap_;
    uint32_t *h_rows = reinterpret_cast<uint32_t *>(h_node + h_cap_);
    const int rs = nmax_ > 0 ? nmax_ : 1;
    uint32_t ctr[8];   // CTR_OVERFLOW is word 0
    CK(cudaMemcpyAsync(h_score, d_best_score_, c * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(h_inc, d_incumbent_, c * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpy2DAsync(h_rows, (size_t)rs * 4, d_best_rows_, (size_t)AMBI_MAXN * 4, (size_t)rs * 4, c, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(h_cand, d_cand_, c * 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(h_node, d_node_, c * 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(ctr, counters_, sizeof ctr, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(funnel_, d_funnel_, sizeof funnel_, cudaMemcpyDeviceToHost, stream_));
    unsigned long long *h_stamps = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(h_rows) + (((size_t)h_cap_ * AMBI_MAXN * 4 + 15) & ~(size_t)15));
    uint8_t *h_act = reinterpret_cast<uint8_t *>(h_stamps + (size_t)h_cap_ * 4);
    unsigned long long stamp0 = 0;
    CK(cudaMemcpyAsync(h_stamps, d_stamps_, c * 32, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(h_act, d_active_, c, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(&stamp0, d_stamp0_, 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    float ms_beam = 0, ms_all = 0, ms_small = 0;
    if (!used_graph_) cudaEventElapsedTime(&ms_beam, ev_[0], ev_[1]);
    cudaEventElapsedTime(&ms_all, ev_[0], ev_[2]);
    if (small_on_ && n_small_ > 0) cudaEventElapsedTime(&ms_small, ev_small_[0], ev_small_[1]);
    prof_small_ms_ = ms_small;
    // per-partition times from the device clock: word 1 = when the returned optimum was first seen as a completed
    // assignment (any pass), word 2 = when the warp-per-partition kernel finished the partition; partitions of the
    // global-memory path are proven when the whole pass ends
    h_tbest_.resize(c); h_tproof_.resize(c);
    for (size_t p = 0; p < c; ++p) {
        const unsigned long long *st = h_stamps + p * 4;
        h_tbest_[p] = st[1] > stamp0 ? (double)(st[1] - stamp0) * 1e-9 : 0.0;
        h_tproof_[p] = (!h_act[p] && st[2] > stamp0) ? (double)(st[2] - stamp0) * 1e-9 : (double)ms_all * 1e-3;
        if (h_tbest_[p] > h_tproof_[p]) h_tbest_[p] = h_tproof_[p];
    }
    prof_total_ms_ = ms_all; prof_beam_ms_ = ms_beam;
    if (profile_) {
        prof_expand_ms_ = 0; prof_expand_launches_ = prof_used_ / 2;
        for (size_t i = 0; i + 1 < prof_used_; i += 2) { float m = 0; cudaEventElapsedTime(&m, prof_ev_[i], prof_ev_[i + 1]); prof_expand_ms_ += m; }
    }
    F->score = h_score; F->incumbent = h_inc; F->rows = h_rows; F->rows_stride = rs; F->cand = h_cand; F->node = h_node;
    F->t_best = h_tbest_.data(); F->t_proof = h_tproof_.data();
    F->overflow = ctr[CTR_OVERFLOW] != 0 || chunked_failed_; F->seconds = ms_all * 1e-3; F->seconds_beam = ms_beam * 1e-3; F->used_graph = used_graph_;
    if (F->overflow && !chunked_failed_) snprintf(err_, sizeof err_, "frontier overflow: more than %u candidate records in one level", cap_);
    return 0;
}

int StageA::fetch(HostResult *out)
{
    Fetched F;
    if (int rc = fetch_raw(&F)) return rc;
    for (int p = 0; p < count_; ++p) {
        out[p].score2 = F.score[p]; out[p].incumbent = F.incumbent[p];
        out[p].candidates = F.cand[p]; out[p].nodes = F.node[p];
        memset(out[p].rows, 0, sizeof out[p].rows);
        memcpy(out[p].rows, F.rows + (size_t)p * F.rows_stride, (size_t)F.rows_stride * 4);
        out[p].seconds = F.t_proof[p]; out[p].seconds_to_best = F.t_best[p];
        out[p].status = F.overflow ? -5 : (F.score[p] == 0x7fffffff ? -4 : 0);
    }
    return F.overflow ? -5 : 0;
}

#endif  // AMBI_EMU
}  // namespace ambi
