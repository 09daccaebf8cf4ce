// ambi_common.cuh -- shared device/host definitions for libambigram_b200.
//
// Data layout in HBM (see DESIGN.md "Data layout"):
//   DevProblem[count]   per-partition inputs (2c_i, 2J_i, TGS masks), read-only
//   NodeRec buffers     two ping-pong arrays of fixed-stride candidate records
//   frontier list       u32 indices of the records that survive a level
//   dedup table         u64 open-addressing slots (generation<<32 | record idx)
#pragma once
#include <stdint.h>
#ifndef AMBI_EMU            // tests/emu compiles the shared definitions with g++ (no CUDA headers)
#include <cuda_runtime.h>
#endif

#define AMBI_MAXN 32
#define AMBI_BUD_INF 3          // "no cap yet" column budget (row 0 only)

// Row-configuration codes.  A row's pattern ends {b1<b2} and loop ends
// {l1<=l2} are numbered from the right end (q = n-1-b) so that the integer
// order of (ip<<10 | il) equals the search order used for tie-breaks
// (column a most significant, smaller multiplicity first).
//   ip: {} -> 0 ; {q1} -> 1+q1(q1+1)/2 ; {q1>q2} -> 1+q1(q1+1)/2+1+q2
//   il: {} -> 0 ; {q1:1} -> 1+2q1+q1(q1-1)/2 ; {q1:1,q2:1} -> ..+1+q2 ; {q1:2} -> ..+1+q1
#define AMBI_NP_CODES 529       // ip < 1 + 31*32/2 + 32
#define AMBI_NL_CODES 561       // il <= 1 + 62 + 465 + 32
#define AMBI_NONE 255

__host__ __device__ inline int ambi_np_for(int m) { return 1 + m + m * (m - 1) / 2; }       // codes with q1 < m
__host__ __device__ inline int ambi_nl_for(int m) { return 1 + 2 * m + m * (m - 1) / 2; }

__host__ __device__ inline int ambi_ip_code(int q1, int q2) {   // q1 > q2 or q2 < 0; q1 < 0 => empty
    if (q1 < 0) return 0;
    int base = 1 + q1 * (q1 + 1) / 2;
    return q2 < 0 ? base : base + 1 + q2;
}
__host__ __device__ inline int ambi_il_code(int q1, int q2) {   // q1 >= q2; q2==q1 => multiplicity 2
    if (q1 < 0) return 0;
    int base = 1 + 2 * q1 + q1 * (q1 - 1) / 2;
    if (q2 < 0) return base;
    if (q2 == q1) return base + 1 + q1;
    return base + 1 + q2;
}

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

struct DevProblem {
    int32_t n;
    int32_t lmax;            // min(2, loop_ub): per-cell loop multiplicity cap for n>=2
    int32_t loop_ub;
    int32_t zero_score;      // score2 of the all-zero assignment = sum(c2)+sum(j2)
    int32_t has_tgs;
    int32_t root_lb;         // admissible bound of the empty prefix (host, same formula as child_bound)
    int32_t ush;             // copy numbers in units of 1/(2 << ush) copies (ambi_bfb_problem.unit_shift)
    int32_t pad;
    int32_t c2[AMBI_MAXN];
    int32_t j2[AMBI_MAXN];
    uint32_t cj[AMBI_MAXN];         // c2 | j2<<16 (both < 65536), what k_expand stages in shared memory
    uint32_t tgs_mask[AMBI_MAXN];   // row a: bit b set if (a,b) is a third-generation span
};

// Candidate record (one partial assignment: rows 0..level-1 fixed).
//   col[b]  = nP | nL<<2 | bP<<4 | bL<<6 for columns b >= level, 0 elsewhere (at byte 16 + b)
//   rows[a] = ip<<10 | il of row a (a < level)
struct NodeHdr {             // the four header words and columns 0..15 share the record's first 32-byte sector:
    int32_t score;           // exact deviation (x2) of segments < level          partitions of <= 16 segments
    int32_t f;               // score + admissible bound on the rest             are merged, selected and staged
    int32_t prob;            //                                                  from that one sector
    int32_t tgs;             // running count of the third-generation row
    uint8_t col[AMBI_MAXN];
};                           // 48 bytes, followed by uint32 rows[nmax]

__host__ __device__ inline size_t ambi_node_stride(int nmax) {
    size_t s = sizeof(NodeHdr) + 4 * (size_t)nmax;
    return (s + 15) & ~(size_t)15;
}
