// host_flatten.h -- host side of the C ABI: ambi_bfb_problem -> DevProblem, row codes -> element_cn in the
// reference's variableIdx numbering (localhap.cpp:122-133, 204-211).  Shared by capi.cu and by the CPU warp
// emulator build of the small-partition kernel (tests/emu, test infrastructure).
#pragma once
#include "../../include/ambigram_b200.h"
#include "ambi_common.cuh"
#include <string>
#include <cstring>
#include <algorithm>

namespace ambi {

// ---- problem flattening ----------------------------------------------------
inline int to_dev(const ambi_bfb_problem &p, DevProblem &d, std::string &why)
{
    memset(&d, 0, sizeof d);
    if (p.n < 1) { why = "n < 1"; return AMBI_ERR_BAD_ARG; }
    if (p.n > AMBI_MAX_SEGMENTS) { why = "n > AMBI_MAX_SEGMENTS"; return AMBI_ERR_TOO_LARGE; }
    if (!p.seg_cn2 || !p.fbi_cn2) { why = "null copy-number arrays"; return AMBI_ERR_BAD_ARG; }
    if (p.unit_shift > 6) { why = "unit_shift > 6"; return AMBI_ERR_NOT_HALF_INTEGRAL; }
    const int ush = (int)p.unit_shift;
    d.n = p.n; d.ush = ush;
    d.loop_ub = p.loop_ub < 0 ? 0 : p.loop_ub;
    d.lmax = d.loop_ub < 2 ? d.loop_ub : 2;
    long long z = 0;
    for (int i = 0; i < p.n; ++i) {
        if (p.seg_cn2[i] < 0 || p.fbi_cn2[i] < 0) { why = "negative copy number"; return AMBI_ERR_BAD_ARG; }
        if (p.seg_cn2[i] > 65535 || p.fbi_cn2[i] > 65535) { why = "2*copy number above 65535 (copy numbers are staged as 16-bit halves)"; return AMBI_ERR_TOO_LARGE; }
        d.c2[i] = p.seg_cn2[i]; d.j2[i] = p.fbi_cn2[i];
        d.cj[i] = (uint32_t)p.seg_cn2[i] | ((uint32_t)p.fbi_cn2[i] << 16);
        z += (long long)p.seg_cn2[i] + p.fbi_cn2[i];
    }
    if (z > 0x3fffffff) { why = "copy numbers too large"; return AMBI_ERR_BAD_ARG; }
    d.zero_score = (int32_t)z;
    {   // bound before any row is fixed: every column may still take 2 patterns + 2 loops (cap 6),
        // every row adds at most 6 (child_bound() with a = -1 and an empty state)
        long long lb = 0;
        for (int j = 0; j < p.n; ++j) {
            int fmax = std::min(6 * (j + 1), 6 * (p.n - j));
            lb += std::max(0, p.seg_cn2[j] - ((2 * fmax) << ush));
            lb += std::max(0, p.fbi_cn2[j] - ((6 + std::min(6 * (p.n - j), 6)) << ush));
        }
        d.root_lb = (int32_t)std::min<long long>(lb, z);
    }
    if (p.n_tgs > 0 && (p.flags & AMBI_FLAG_JUNC_INFO) && p.tgs_pairs) {
        for (int t = 0; t < p.n_tgs; ++t) {
            int s = p.tgs_pairs[2 * t] - p.start_id, e = p.tgs_pairs[2 * t + 1] - p.start_id;
            if (s > e) std::swap(s, e);
            // Documented divergence: for a span that leaves the partition the reference indexes variableIdx with a
            // key that does not exist (std::map creates it with value 0, so P[s,s] of the partition lands in the row
            // twice, LGM:4692-4697) -- an accident of operator[], not a model.  Such a span constrains no variable of
            // this partition and is dropped here, as in the oracle (orc_solve_units).
            if (s < 0 || e >= p.n) continue;
            if (s == 0 && e == p.n - 1) continue;                 // LGM:4690
            d.tgs_mask[s] |= 1u << e; d.has_tgs = 1;
        }
    }
    return AMBI_OK;
}

inline void decode_codes(uint32_t code, int n, int out[4])
{
    // inverse of ambi_ip_code / ambi_il_code: columns p1<p2, l1<=l2 (or -1)
    out[0] = out[1] = out[2] = out[3] = -1;
    int ip = code >> 10, il = code & 1023;
    if (ip) {
        int q1 = 0; while (ambi_ip_code(q1 + 1, -1) <= ip) ++q1;
        int off = ip - ambi_ip_code(q1, -1);
        out[0] = n - 1 - q1;
        if (off > 0) out[1] = n - 1 - (off - 1);
    }
    if (il) {
        int q1 = 0; while (ambi_il_code(q1 + 1, -1) <= il) ++q1;
        int off = il - ambi_il_code(q1, -1);
        out[2] = n - 1 - q1;
        if (off > 0) out[3] = (off - 1 == q1) ? out[2] : n - 1 - (off - 1);
    }
}

inline int pair_index(int n, int a, int b) { return a * n - a * (a - 1) / 2 + (b - a); }

inline void rows_to_element_cn(int n, const uint32_t *rows, int32_t *x)
{
    const int H = n * (n + 1) / 2;
    memset(x, 0, sizeof(int32_t) * 2 * H);
    if (n == 1) { x[0] = (rows[0] >> 20) & 1; x[1] = rows[0] & 0xfffff; return; }
    for (int a = 0; a < n; ++a) {
        int c[4]; decode_codes(rows[a], n, c);
        if (c[0] >= 0) x[pair_index(n, a, c[0])] += 1;
        if (c[1] >= 0) x[pair_index(n, a, c[1])] += 1;
        if (c[2] >= 0) x[H + pair_index(n, a, c[2])] += 1;
        if (c[3] >= 0) x[H + pair_index(n, a, c[3])] += 1;
    }
}


}  // namespace ambi
