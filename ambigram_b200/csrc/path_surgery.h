// path_surgery.h -- what getBFB does to a breakpoint path once an order has assembled (LGM:3656-3670), as one
// implementation for the host entry points (ambi_bfb_imperfect_fbi / ambi_bfb_expand_path) and for the device kernel
// that post-processes every valid order of --all (k_path_surgery, stage_b.cu).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define AMBI_HD __host__ __device__
#else
#define AMBI_HD
#endif

namespace ambi {

// imperfectFBI (LGM:3431-3512): rewrite fold vertices whose fold-back junction joins two different segments.
// bkp: signed segment ids (+ strand / - strand), pairs (first, last) of monotone runs; inv_src / inv_tgt are indexed
// by absolute segment id (0 = none) and carry the endpoints of inversions[id] as getJuncCN recorded them
// (LGM:4014-4048).
template <class T, class U>
AMBI_HD inline void imperfect_fbi(T *bkp, int len, const U *inv_src, const U *inv_tgt, int max_id)
{
    int pos = 0;
    while (pos < len) {
        int r = len;                                   // first q >= pos+3 holding the complement of bkp[pos]
        for (int q = pos + 3; q < len; ++q) if (bkp[q] == -bkp[pos]) { r = q; break; }
        const int l = r - 1;
        if (r == len || bkp[l] != -bkp[pos + 1]) {     // not the outer pair of a palindrome (LGM:3442-3472)
            int id = bkp[pos + 1] < 0 ? -bkp[pos + 1] : bkp[pos + 1];
            if (id >= 1 && id <= max_id && inv_src[id] != 0) {
                const int lo = inv_src[id] < inv_tgt[id] ? inv_src[id] : inv_tgt[id], hi = inv_src[id] < inv_tgt[id] ? inv_tgt[id] : inv_src[id];
                bkp[pos + 1] = (T)(bkp[pos + 1] > 0 ? lo : -hi);
            }
            if (pos > 0) {
                id = bkp[pos] < 0 ? -bkp[pos] : bkp[pos];
                const int prev = bkp[pos - 1] < 0 ? -bkp[pos - 1] : bkp[pos - 1];
                if (id >= 1 && id <= max_id && inv_src[id] != 0 && prev == id) {
                    const int other = inv_src[id] == id ? inv_tgt[id] : inv_src[id];
                    bkp[pos] = (T)(bkp[pos] > 0 ? other : -other);
                }
            }
            const int m0 = bkp[pos] < 0 ? -bkp[pos] : bkp[pos], m1 = bkp[pos + 1] < 0 ? -bkp[pos + 1] : bkp[pos + 1];
            if (bkp[pos] > 0 && m0 > m1) bkp[pos + 1] = bkp[pos];
            if (bkp[pos] < 0 && m0 < m1) bkp[pos + 1] = bkp[pos];
            pos += 2;
        } else {                                       // palindrome: walk outwards from its centre (LGM:3473-3509)
            int p1 = pos + (l - pos) / 2, p2 = p1 + 1;
            while (p1 >= pos - 1 && p1 > 0) {
                const int id = bkp[p1] < 0 ? -bkp[p1] : bkp[p1];
                if (id >= 1 && id <= max_id && inv_src[id] != 0) {
                    const int lo = inv_src[id] < inv_tgt[id] ? inv_src[id] : inv_tgt[id], hi = inv_src[id] < inv_tgt[id] ? inv_tgt[id] : inv_src[id];
                    if (bkp[p1] > 0) { bkp[p1] = (T)lo; bkp[p1 + 1] = (T)-hi; }
                    else { bkp[p1] = (T)-hi; bkp[p1 + 1] = (T)lo; }
                    if (p2 != p1 + 1) {
                        if (p1 > pos - 1) bkp[p2] = (T)-bkp[p1];
                        bkp[p2 - 1] = (T)-bkp[p1 + 1];
                    }
                }
                p1 -= 2; p2 += 2;
            }
            pos = r + 1;
        }
    }
}

// expansion (LGM:3661-3670): each pair (first,last) of a monotone run expands to every id in between; returns the
// full length (entries beyond `cap` are not written)
template <class T, class O>
AMBI_HD inline int expand_path(const T *bkp, int bkp_len, O *path_out, int cap)
{
    int len = 0;
    for (int j = 1; j < bkp_len; j += 2) {
        const int u = bkp[j - 1], v = bkp[j];
        if (u > 0) { for (int k = u; k <= v; ++k) { if (len < cap) path_out[len] = (O)k; ++len; } }
        else { for (int k = -u; k >= -v; --k) { if (len < cap) path_out[len] = (O)-k; ++len; } }
    }
    return len;
}

}  // namespace ambi
