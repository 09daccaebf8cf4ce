// decode_common.cuh -- enumeration of a frontier record's row configurations, shared by the two Stage-A kernels
// (k_expand in stage_a.cu: level-synchronous, records in HBM; k_solve_small in stage_small.cu: one warp per
// partition, records in shared memory).  Reference model: BFB_ILP rows 1-6, src/LocalGenomicMap.cpp:4547-4703.
//
// A record at row a has, per open column j, the counters nP | nL<<2 | bP<<4 | bL<<6 (DESIGN.md "Search
// formulation").  From them phase 1 derives six column masks, and the row configurations that satisfy the
// budgets and the parent rows 1/3 are numbered 0..T-1:
//   [0, nl0)          empty pattern set x class-0 loop multisets (parents already in the column): {}, singles in
//                     A1 = okL1 & par1, doubles in A2 = okL2 & par2, pairs with the high end in A1 and the low end
//                     any budgeted column below it
//   then one block per pattern high end hi in okP & parP, (1 + k_hi) x (nl0 + e_hi) configurations:
//                     pattern sets {hi} and {lo, hi} (lo any budgeted column below hi -- hi is lo's parent) times the
//                     class-0 multisets plus the singles at E & below(hi), E = okL1 & ~par1 (their only parent is
//                     this row's hi; two entries below hi are cut by rows 4/5 anyway).
// Decoding a number back into (pattern ends, loop ends) used to be a population-count search per part
// (17 % of k_expand's instructions in round 1).  Here phase 1 lists, once per record, the positions of the set
// bits of okP and E and the class-0 multisets themselves, so the decode is a table read: the k-th budgeted column
// below hi is the k-th set bit of okP whatever hi is, and likewise for E.
#pragma once
#include <stdint.h>

namespace ambi {
namespace dec {

constexpr int L0CAP = 16;               // class-0 loop multisets listed per record (more: decoded the slow way)
constexpr uint32_t ENTRY_VALID = 0x80000000u;    // queue entry: ENTRY_VALID | slot << 24 | packed configuration

__device__ __forceinline__ uint32_t below(int b) { return (1u << b) - 1u; }
__device__ __forceinline__ int nth_bit_slow(uint32_t m, int k) { for (; k > 0; --k) m &= m - 1; return __ffs((int)m) - 1; }
// packed row configuration: four 6-bit fields holding column + 1 (0 = absent; p1<p2, l1<=l2), so that a field decodes
// with one subtraction
__device__ __forceinline__ uint32_t pack_cfg(int p1, int p2, int l1, int l2)
{
    return (uint32_t)(p1 + 1) | ((uint32_t)(p2 + 1) << 6) | ((uint32_t)(l1 + 1) << 12) | ((uint32_t)(l2 + 1) << 18);
}
__device__ __forceinline__ void unpack_cfg(uint32_t c, int &p1, int &p2, int &l1, int &l2)
{
    p1 = (int)(c & 63u) - 1; p2 = (int)((c >> 6) & 63u) - 1; l1 = (int)((c >> 12) & 63u) - 1; l2 = (int)((c >> 18) & 63u) - 1;
}

// Phase 1, lane = record: number of configurations of a record with these masks (nl0 = class-0 loop multisets).
__device__ __forceinline__ uint32_t count_configs(uint32_t okP, uint32_t parP, uint32_t okL1, uint32_t okL2, uint32_t par1,
                                                  uint32_t par2, uint32_t *nl0_out)
{
    const uint32_t A1 = okL1 & par1, A2 = okL2 & par2, E = okL1 & ~par1;
    uint32_t nl0 = 1 + __popc(A1) + __popc(A2);
    for (uint32_t m = A1; m; m &= m - 1) nl0 += __popc(okL1 & below(__ffs((int)m) - 1));
    uint32_t total = nl0;
    for (uint32_t m = okP & parP; m; m &= m - 1) {
        const int hi = __ffs((int)m) - 1;
        total += (1u + __popc(okP & below(hi))) * (nl0 + __popc(E & below(hi)));
    }
    *nl0_out = nl0;
    return total;
}

static __device__ __noinline__ uint32_t count_configs_call(uint32_t okP, uint32_t parP, uint32_t okL1, uint32_t okL2, uint32_t par1,
                                                    uint32_t par2, uint32_t *nl0_out)
{
    return count_configs(okP, parP, okL1, okL2, par1, par2, nl0_out);
}

// ... and its masks and decode lists, written to staging slot `slot`.  SH needs okP/parP/okL1/A2m/par1/nl0 [32],
// posP/posE [32][NCOL] (uint8), L0 [32][L0CAP] (uint16): a record's lists are contiguous, so lanes decoding
// candidates of the same record read the same few words.
template <class SH>
__device__ __forceinline__ void write_lists(SH &S, int slot, uint32_t okP, uint32_t parP, uint32_t okL1, uint32_t okL2,
                                            uint32_t par1, uint32_t par2, uint32_t nl0)
{
    const uint32_t A1 = okL1 & par1, A2 = okL2 & par2, E = okL1 & ~par1;
    if (nl0 >= 1 && nl0 <= (uint32_t)L0CAP) {          // nl0 == 0: a record without configurations (nothing is decoded)
        int k = 1;
        for (uint32_t m = A1; m; m &= m - 1) S.L0[slot][k++] = (uint16_t)((__ffs((int)m) - 1) | 0xff00);
        for (uint32_t m = A2; m; m &= m - 1) { const int c = __ffs((int)m) - 1; S.L0[slot][k++] = (uint16_t)(c | (c << 8)); }
        for (uint32_t m = A1; m; m &= m - 1) {
            const int hi = __ffs((int)m) - 1;
            for (uint32_t q = okL1 & below(hi); q; q &= q - 1) S.L0[slot][k++] = (uint16_t)((__ffs((int)q) - 1) | (hi << 8));
        }
    }
    if (nl0 >= 1) {
        int k = 0;
        for (uint32_t m = okP; m; m &= m - 1, ++k) S.posP[slot][k] = (uint8_t)(__ffs((int)m) - 1);
        k = 0;
        for (uint32_t m = E; m; m &= m - 1, ++k) S.posE[slot][k] = (uint8_t)(__ffs((int)m) - 1);
    }
    S.okP[slot] = okP; S.parP[slot] = parP; S.okL1[slot] = okL1; S.A2m[slot] = A2; S.par1[slot] = par1; S.nl0[slot] = nl0;
}

// Lane = candidate: decode configuration number r of record `slot` and run the light test (rows 2/4/5 inside the
// row, third-generation row `tm` with the record's running count, exact deviation of segment a through the
// record's d0 = 2 cov_a - 2c_a and f0 = fb_a - 2J_a, and the O(1) prefilter lslack = slack - own bound).
// Returns ENTRY_VALID | slot << 24 | cfg, or 0 when structurally infeasible; *pass = within the limit.
// `ush`: copy numbers are in units of 1/(2 << ush) copies (0 in the half-copy instantiations: the shifts fold away).
template <class SH>
__device__ __forceinline__ uint32_t decode_light(const SH &S, int slot, uint32_t r, uint32_t tm, int ush, bool *pass)
{
    const uint32_t okP = S.okP[slot], okL1 = S.okL1[slot], par1 = S.par1[slot];
    const uint32_t nl0 = S.nl0[slot];
    int p1 = -1, p2 = -1, l1 = -1, l2 = -1;
    uint32_t li = r;
    if (r >= nl0) {
        r -= nl0;
        const uint32_t E = okL1 & ~par1;
        int hi = 0;
        uint32_t width = 1;
        for (uint32_t m = okP & S.parP[slot]; m; m &= m - 1) {
            hi = __ffs((int)m) - 1;
            width = nl0 + __popc(E & below(hi));
            const uint32_t blk = (1u + __popc(okP & below(hi))) * width;
            if (r < blk) break;
            r -= blk;
        }
        // r < 2^16 and width < 2^11: (r + 0.5) / width is at least 2^-12 away from every integer, far more than the
        // error of the approximate float division, so truncation gives the exact quotient
        const uint32_t pi = __float2uint_rz(__fdividef((float)r + 0.5f, (float)width));
        li = r - pi * width;
        if (pi == 0) p1 = hi;
        else { p2 = hi; p1 = (int)S.posP[slot][pi - 1]; }
    }
    if (li >= nl0) l1 = (int)S.posE[slot][li - nl0];
    else if (li > 0) {
        if (nl0 <= (uint32_t)L0CAP) {
            const uint32_t e = S.L0[slot][li];
            l1 = (int)(e & 0xffu); l2 = (int)(e >> 8); if (l2 == 0xff) l2 = -1;
        } else {                                        // rare: more class-0 multisets than the list holds
            --li;
            const uint32_t M1 = okL1 & par1, M2 = S.A2m[slot];
            const uint32_t n1 = __popc(M1), n2 = __popc(M2);
            if (li < n1) l1 = nth_bit_slow(M1, (int)li);
            else if (li - n1 < n2) { l1 = nth_bit_slow(M2, (int)(li - n1)); l2 = l1; }
            else {
                li -= n1 + n2;
                for (uint32_t m = M1; m; m &= m - 1) {
                    const int hi = __ffs((int)m) - 1;
                    const uint32_t lm = okL1 & below(hi);
                    const uint32_t k = __popc(lm);
                    if (li < k) { l2 = hi; l1 = nth_bit_slow(lm, (int)li); break; }
                    li -= k;
                }
            }
        }
    }
    const int nl = (l1 >= 0) + (l2 >= 0);
    const int lhi = l2 >= 0 ? l2 : l1;
    const int pmax = p2 >= 0 ? p2 : p1;
    bool feas = !(nl == 2 && pmax > lhi);            // a pattern end right of both loop entries (rows 4/5)
    if (feas && tm) {
        const int hits = (p1 >= 0 && ((tm >> p1) & 1)) + (p2 >= 0 && ((tm >> p2) & 1)) +
                         (l1 >= 0 && ((tm >> l1) & 1)) + (l2 >= 0 && ((tm >> l2) & 1));
        if (S.ptgs[slot] + hits > 5) feas = false;
    }
    if (!feas) { *pass = false; return 0u; }
    const int w = (p1 >= 0) + (p2 >= 0) + 2 * nl;
    const int err = abs(S.d0[slot] + (w << (ush + 1))) + abs(S.f0[slot] + (w << ush));
    // O(1) prefilter.  lslack = slack - (the record's own bound on segments > a); a child's bound on those segments is
    // never below the record's.  Sharper: on a segment whose committed coverage already reaches its copy number
    // (over: 2 cov_j >= 2c_j), every copy this row adds costs exactly 2 more, and likewise for the fold-back count of
    // a column that already reaches 2J_j (fover) -- both terms of the record's bound are exact there.
    // (fill_ovt() below tabulates the two masks per record.)
    // ovt[slot][1 + e] = 2 * (such segments in (a, e]) + (column e is such a fold-back column), ovt[slot][0] = 0 for an
    // absent end (-1): four table reads on the load/store pipe instead of four population counts on the alu pipe
    const uint8_t *t = S.ovt[slot];
    const int extra = (int)t[1 + p1] + (int)t[1 + p2] + 2 * ((int)t[1 + l1] + (int)t[1 + l2]);
    *pass = err + (extra << ush) <= S.lslack[slot];
    return ENTRY_VALID | ((uint32_t)slot << 24) | pack_cfg(p1, p2, l1, l2);
}

// ---------------------------------------------------------------------------------------------------------------
// Incremental enumeration (k_expand).  Decoding every configuration number from scratch cost ~340 thread
// instructions per configuration (half of k_expand in the round-2 profile): the block scan over the pattern high
// ends, a division, list reads.  Consecutive numbers differ in one list position, so a lane now takes a CHUNK of
// consecutive numbers of one record: it decodes the first the slow way into a cursor (cursor_start) and steps the
// cursor for the rest (cursor_eval + cursor_advance: a list read, the light test, a counter).
//
// Lists per staged record (SH additionally needs LE [32][L0CAP + NCOL] uint16, A1m [32]):
//   LE[slot][li]  loop entry number li as (l1 + 1) | (l2 + 1) << 8 (0 = absent): li < nl0 the class-0 multisets in the
//                 order of the numbering above, then the singles of E in increasing column order -- under high end hi
//                 the list is walked up to width = nl0 + |E below hi|.  More than L0CAP class-0 multisets (rare): the
//                 singles start at L0CAP and the multisets are decoded by population counts (loop_entry_slow).
// ---------------------------------------------------------------------------------------------------------------
struct Cursor {
    uint32_t m_rem;         // pattern high ends still to come (bits above the current one)
    uint32_t li, width;     // position in / length of the loop-entry list under the current pattern set
    int hi, pi, kP;         // high end (-1: no pattern end), set number (0: {hi}; k: {k-th budgeted column below hi, hi}), last set number
    uint32_t pcode;         // pack_cfg(p1, p2, -1, -1)
    int pw;                 // pattern ends in the set | their prefilter charge << 8
};

template <class SH>
__device__ __noinline__ void write_lists2(SH &S, int slot, uint32_t okP, uint32_t parP, uint32_t okL1, uint32_t okL2,
                                          uint32_t par1, uint32_t par2, uint32_t nl0)
{
    // (not inlined, loops not unrolled: once per record, and k_expand has to stay inside the instruction cache)
    const uint32_t A1 = okL1 & par1, A2 = okL2 & par2, E = okL1 & ~par1;
    uint16_t *le = S.LE[slot];
    int k = L0CAP;
    if (nl0 <= (uint32_t)L0CAP) {
        k = 0;
        le[k++] = 0;
#pragma unroll 1
        for (uint32_t m = A1; m; m &= m - 1) le[k++] = (uint16_t)__ffs((int)m);
#pragma unroll 1
        for (uint32_t m = A2; m; m &= m - 1) { const int c = __ffs((int)m); le[k++] = (uint16_t)(c | (c << 8)); }
#pragma unroll 1
        for (uint32_t m = A1; m; m &= m - 1) {
            const int hi = __ffs((int)m) - 1;
#pragma unroll 1
            for (uint32_t q = okL1 & below(hi); q; q &= q - 1) le[k++] = (uint16_t)(__ffs((int)q) | ((hi + 1) << 8));
        }
    }
#pragma unroll 1
    for (uint32_t m = E; m; m &= m - 1) le[k++] = (uint16_t)__ffs((int)m);
    k = 0;
#pragma unroll 1
    for (uint32_t m = okP; m; m &= m - 1, ++k) S.posP[slot][k] = (uint8_t)(__ffs((int)m) - 1);
    S.okP[slot] = okP; S.parP[slot] = parP; S.okL1[slot] = okL1; S.A1m[slot] = A1; S.A2m[slot] = A2; S.nl0[slot] = (uint16_t)nl0;
}

// loop entry li of a record with more class-0 multisets than the list holds
template <class SH>
__device__ __noinline__ uint32_t loop_entry_slow(const SH &S, int slot, uint32_t li)
{
    const uint32_t nl0 = S.nl0[slot];
    if (li >= nl0) return S.LE[slot][L0CAP + li - nl0];
    if (li == 0) return 0u;
    --li;
    const uint32_t M1 = S.A1m[slot], M2 = S.A2m[slot], okL1 = S.okL1[slot];
    const uint32_t n1 = __popc(M1), n2 = __popc(M2);
    if (li < n1) return (uint32_t)(nth_bit_slow(M1, (int)li) + 1);
    if (li - n1 < n2) { const uint32_t c = (uint32_t)(nth_bit_slow(M2, (int)(li - n1)) + 1); return c | (c << 8); }
    li -= n1 + n2;
    for (uint32_t m = M1; m; m &= m - 1) {
        const int hi = __ffs((int)m) - 1;
        const uint32_t lm = okL1 & below(hi);
        const uint32_t k = __popc(lm);
        if (li < k) return (uint32_t)(nth_bit_slow(lm, (int)li) + 1) | ((uint32_t)(hi + 1) << 8);
        li -= k;
    }
    return 0u;
}

// pattern set (hi, pi) of the cursor -> its code and prefilter charge
template <class SH>
__device__ __forceinline__ void cursor_pattern(const SH &S, int slot, Cursor &c)
{
    const uint8_t *t = S.ovt[slot];
    if (c.hi < 0) { c.pcode = 0u; c.pw = 0; }
    else if (c.pi == 0) { c.pcode = (uint32_t)(c.hi + 1); c.pw = 1 | ((int)t[c.hi + 1] << 8); }
    else {
        const int lo = (int)S.posP[slot][c.pi - 1];
        c.pcode = (uint32_t)(lo + 1) | ((uint32_t)(c.hi + 1) << 6);
        c.pw = 2 | (((int)t[lo + 1] + (int)t[c.hi + 1]) << 8);
    }
}

// cursor at configuration number r of record `slot` (r < the record's count)
template <class SH>
__device__ __forceinline__ void cursor_start(const SH &S, int slot, uint32_t r, Cursor &c)
{
    const uint32_t okP = S.okP[slot], nl0 = S.nl0[slot];
    uint32_t m = okP & S.parP[slot];
    c.hi = -1; c.kP = 0; c.pi = 0; c.width = nl0; c.li = r;
    if (r >= nl0) {
        r -= nl0;
        const uint32_t E = S.okL1[slot] & ~S.A1m[slot];
        for (;; m &= m - 1) {                            // r < count: the scan ends at a block
            c.hi = __ffs((int)m) - 1;
            c.kP = __popc(okP & below(c.hi));
            c.width = nl0 + __popc(E & below(c.hi));
            const uint32_t blk = (1u + (uint32_t)c.kP) * c.width;
            if (r < blk) break;
            r -= blk;
        }
        m &= m - 1;
        // r < 2^18 (at most ~2.6e5 configurations for 32 open columns), width <= 1024 and the quotient <= 512:
        // (r + 0.5) / width is at least 2^-11 away from every integer, the approximate float division is off by at
        // most 2 ulp of a value below 512 (2^-13), so truncation gives the exact quotient
        const uint32_t pi = __float2uint_rz(__fdividef((float)r + 0.5f, (float)c.width));
        c.pi = (int)pi;
        c.li = r - pi * c.width;
    }
    c.m_rem = m;
    cursor_pattern(S, slot, c);
}

// the configuration under the cursor through the light test (see decode_light): ENTRY_VALID | slot << 24 | cfg, or 0
// when structurally infeasible; *pass = within the limit
template <bool SCALED, class SH>
__device__ __forceinline__ uint32_t cursor_eval(const SH &S, int slot, const Cursor &c, uint32_t nl0, bool *pass)
{
    const uint32_t e = nl0 <= (uint32_t)L0CAP ? (uint32_t)S.LE[slot][c.li] : loop_entry_slow(S, slot, c.li);
    const uint32_t c1 = e & 0xffu, c2 = e >> 8;
    bool feas = !(c2 != 0u && (uint32_t)(c.hi + 1) > c2);           // a pattern end right of both loop entries (rows 4/5)
    const uint32_t tm = S.tmask[slot];
    if (tm) {
        const unsigned long long t1 = (unsigned long long)tm << 1;  // bit (column + 1); code 0 (absent) hits nothing
        const int hits = (int)((t1 >> (c.pcode & 63u)) & 1u) + (int)((t1 >> ((c.pcode >> 6) & 63u)) & 1u) + (int)((t1 >> c1) & 1u) + (int)((t1 >> c2) & 1u);
        if (S.ptgs[slot] + hits > 5) feas = false;
    }
    const int w = (c.pw & 0xff) + 2 * ((c1 != 0u) + (c2 != 0u));
    const int ush = SCALED ? (int)S.ush[slot] : 0;
    const int err = abs(S.d0[slot] + (w << (ush + 1))) + abs(S.f0[slot] + (w << ush));
    const uint8_t *t = S.ovt[slot];
    const int extra = (c.pw >> 8) + 2 * ((int)t[c1] + (int)t[c2]);
    *pass = feas && err + (extra << ush) <= S.lslack[slot];
    return feas ? (ENTRY_VALID | ((uint32_t)slot << 24) | c.pcode | (c1 << 12) | (c2 << 18)) : 0u;
}

// ... and on to the next configuration number
template <class SH>
__device__ __forceinline__ void cursor_advance(const SH &S, int slot, Cursor &c, uint32_t nl0)
{
    if (++c.li < c.width) return;
    c.li = 0;
    if (++c.pi > c.kP) {
        c.pi = 0;
        if (!c.m_rem) return;                            // past the record's last configuration
        c.hi = __ffs((int)c.m_rem) - 1;
        c.m_rem &= c.m_rem - 1;
        c.kP = __popc(S.okP[slot] & below(c.hi));
        c.width = nl0 + __popc(S.okL1[slot] & ~S.A1m[slot] & below(c.hi));
    }
    cursor_pattern(S, slot, c);
}

// Phase 1, lane = record: the prefilter table of decode_light from the record's masks `over` (segments j > a whose
// committed coverage already reaches 2c_j) and `fover` (columns whose fold-back count already reaches 2J_j).
template <class SH>
__device__ __forceinline__ void fill_ovt(SH &S, int slot, uint32_t over, uint32_t fover, int n)
{
    uint8_t *t = S.ovt[slot];
    t[0] = 0;
    int run = 0;
#pragma unroll 1
    for (int e = 0; e < n; ++e) { run += (int)((over >> e) & 1u); t[1 + e] = (uint8_t)(2 * run + (int)((fover >> e) & 1u)); }
}

// Columns an end may sit at without its prefilter charge alone exceeding B: charge(c) = 2 * (segments of `over` in
// (a, c]) + (c in fover) <= B (the table of fill_ovt, evaluated on the masks).
__device__ __forceinline__ uint32_t cols_with_run_upto(uint32_t over, int k)
{
    uint32_t m = over;
#pragma unroll 1
    for (int i = 0; i < k && m; ++i) m &= m - 1;        // m: the (k+1)-th segment of `over` and everything above
    return m ? below(__ffs((int)m) - 1) : 0xffffffffu;
}
static __device__ __noinline__ uint32_t allowed_cols(uint32_t over, uint32_t fover, int B)
{
    if (B < 0) return 0u;
    const uint32_t mA = cols_with_run_upto(over, B >> 1);                     // 2 run <= B
    const uint32_t mB = B >= 1 ? cols_with_run_upto(over, (B - 1) >> 1) : 0u;   // 2 run + 1 <= B
    return (mA & ~fover) | (mB & fover);
}

// Phase 1 of k_expand: take out of the three column masks every column the light test would refuse in every
// configuration.  An end at column c charges ovt[c] << ush (a loop end twice, two loop ends at one column four times
// that); the exact deviation of segment a is at least its minimum over the weights the end can come with (a pattern
// end: w >= 1, a loop end: w >= 2, two loop ends: w >= 4; w <= 6).  Returns true when something was removed.
__device__ __forceinline__ bool restrict_masks(uint32_t over, uint32_t fover, int lslack, int d0, int f0, int ush,
                                               uint32_t *okP, uint32_t *okL1, uint32_t *okL2)
{
    int e[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) e[k] = abs(d0 + (k << (ush + 1))) + abs(f0 + (k << ush));
    const int m4 = min(e[4], min(e[5], e[6])), m2 = min(min(e[2], e[3]), m4), m1 = min(e[1], m2);
    const uint32_t rP = allowed_cols(over, fover, (lslack - m1) >> ush);
    const uint32_t rL1 = allowed_cols(over, fover, (lslack - m2) >> (ush + 1));
    const uint32_t rL2 = allowed_cols(over, fover, (lslack - m4) >> (ush + 2));
    if (!((*okP & ~rP) | (*okL1 & ~rL1) | (*okL2 & ~rL2))) return false;
    *okP &= rP; *okL1 &= rL1; *okL2 &= rL2;
    return true;
}

// Owner of candidate cb0 + lane by a run-length walk over the records' candidate ranges: s0 = record that owns
// candidate cb0; bit (e - cb0 - 1) of the returned mask marks a record that ends at e, so the owner of lane k is
// s0 + popc(mask & below(k)) and the next chunk starts at record s0 + popc(mask).  Every staged record must have at
// least one configuration (ends are then distinct); entries of cstart beyond the last record hold the total.
template <class SH>
__device__ __forceinline__ uint32_t owner_mask(const SH &S, uint32_t s0, uint32_t cb0, int lane)
{
    const uint32_t idx = s0 + 1u + (uint32_t)lane;
    const uint32_t e = S.cstart[idx <= 32u ? idx : 32u];
    const uint32_t d = e - cb0 - 1u;
    return __reduce_or_sync(0xffffffffu, (idx <= 32u && d < 32u) ? 1u << d : 0u);
}

}  // namespace dec
}  // namespace ambi
