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
//   k_expand    K1+K2  one warp per frontier record: enumerate the feasible row
//                      configurations, score them (exact deviation of segment a
//                      + admissible bound for the rest), write the survivors
//   k_dedup     K3a    hash-table merge of records that reach the same column
//                      state: keep min (score, prefix in search order)
//   k_finalize  K3b    per-problem arg-min of completed assignments, frontier
//                      list for the next level, beam statistics
//   k_beam_*    K3c    per-problem top-W selection by f-histogram (incumbent pass)
#include "ambi_common.cuh"
#include "stage_a.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

namespace ambi {

constexpr int WPB = 8;              // warps per block in k_expand
constexpr int PL_CAP = 544;         // >= AMBI_NP_CODES, multiple of 32
constexpr int LL_CAP = 576;         // >= AMBI_NL_CODES, multiple of 32
constexpr int BEAM_BINS = 64;

struct LevelArgs {
    const DevProblem *probs;
    const uint8_t *parents;
    const uint32_t *plist;
    const uint32_t *pcount;
    uint8_t *children;
    uint32_t *ccount;
    uint32_t cap;
    const int32_t *limit;
    unsigned long long *cand_count;
    unsigned long long *node_count;
    uint32_t *work_counter;
    uint32_t *overflow;
    const uint8_t *ptab;            // [AMBI_NP_CODES][2] (q1,q2)
    const uint8_t *ltab;            // [AMBI_NL_CODES][2]
    int level;
    int stride;
    int nmax;
    int shard_world, shard_rank, shard_level;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

struct WarpShared {
    uint16_t pl[PL_CAP];
    uint16_t ll[LL_CAP];
    int32_t c2[AMBI_MAXN];
    int32_t j2[AMBI_MAXN];
    uint32_t col[AMBI_MAXN];        // w | bP<<8 | bL<<16 | (nP+nL)<<24
    uint32_t queue[64];
};

// Admissible bound on segments j > a for a child (ends p1<p2, l1<=l2; -1 = absent).
// Mirrors oracle/ambi_oracle.c future_bound(): committed coverage / fold-back can
// only grow; what can still be added is capped by column budgets and 6 per row.
__device__ __forceinline__ int child_bound(const WarpShared &S, int n, int a,
                                           int p1, int p2, int l1, int l2, int budget)
{
    int lb = 0, cov = 0, capsum = 0;
    for (int j = n - 1; j > a; --j) {
        uint32_t cw = S.col[j];
        int w = cw & 0xff, bP = (cw >> 8) & 0xff, bL = (cw >> 16) & 0xff;
        int pab = (j == p1) + (j == p2), lab = (j == l1) + (j == l2);
        int rp = (p1 >= 0 && p1 < j) + (p2 >= 0 && p2 < j);
        int rl = (l1 >= 0 && l1 < j) + (l2 >= 0 && l2 < j);
        int kP = 2 - pab - max(rp, rl);
        int kL = 2 - max(lab + rl, max(pab + rl, pab + rp));
        int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
        int cap = min(nbP, 2) + 2 * min(nbL, 2);
        int wn = w + pab + 2 * lab;
        cov += wn; capsum += cap;
        int fmax = min(6 * (j - a), capsum);
        int d = 2 * cov - S.c2[j];
        lb += d > 0 ? d : max(0, -d - 2 * fmax);
        int gmax = cap + min(capsum, 6);
        int f = wn - S.j2[j];
        lb += f > 0 ? f : max(0, -f - gmax);
        if (lb > budget) break;
    }
    return lb;
}

__global__ void __launch_bounds__(WPB * 32)
k_expand(LevelArgs A)
{
    __shared__ uint8_t s_ptab[AMBI_NP_CODES * 2 + 6];
    __shared__ uint8_t s_ltab[AMBI_NL_CODES * 2 + 6];
    __shared__ WarpShared s_w[WPB];

    for (int i = threadIdx.x; i < AMBI_NP_CODES * 2; i += blockDim.x) s_ptab[i] = A.ptab[i];
    for (int i = threadIdx.x; i < AMBI_NL_CODES * 2; i += blockDim.x) s_ltab[i] = A.ltab[i];
    __syncthreads();

    const int lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    WarpShared &S = s_w[threadIdx.x >> 5];
    const int a = A.level;
    const uint32_t pcount = *A.pcount;

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(A.work_counter, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= pcount) break;

        const uint32_t pidx = A.plist[wi];
        const uint8_t *prec = A.parents + (size_t)pidx * A.stride;
        const NodeHdr *ph = reinterpret_cast<const NodeHdr *>(prec);
        const uint32_t *prow = reinterpret_cast<const uint32_t *>(prec + sizeof(NodeHdr));
        const int score = ph->score, prob = ph->prob, ptgs = ph->tgs;
        const DevProblem &P = A.probs[prob];
        const int n = P.n;
        const int slack = A.limit[prob] - score;
        if (n <= a || slack < 0) continue;
        const int lmax = P.lmax;
        const uint32_t tmask = P.has_tgs ? P.tgs_mask[a] : 0u;

        // ---- stage the parent's column state in shared memory (lane j = column j)
        const uint32_t cb = ph->col[lane];
        const bool incol = lane >= a && lane < n;
        const int nP = cb & 3, nL = (cb >> 2) & 3, bP = (cb >> 4) & 3, bL = (cb >> 6) & 3;
        const int wj = incol ? nP + 2 * nL : 0;
        __syncwarp();
        S.col[lane] = incol ? (uint32_t)(wj | (bP << 8) | (bL << 16) | ((nP + nL) << 24)) : 0u;
        S.c2[lane] = lane < n ? P.c2[lane] : 0;
        S.j2[lane] = lane < n ? P.j2[lane] : 0;
        int covC = wj;
#pragma unroll
        for (int o = 16; o; o >>= 1) covC += __shfl_xor_sync(0xffffffffu, covC, o);
        const int fbC = __shfl_sync(0xffffffffu, wj, a);
        const bool rootcol = (a == 0 && lane == n - 1);
        const uint32_t okP = __ballot_sync(0xffffffffu, incol && bP >= 1);
        const uint32_t parP = __ballot_sync(0xffffffffu, incol && (nP >= 1 || rootcol));
        const uint32_t okL1 = __ballot_sync(0xffffffffu, incol && bL >= 1 && lmax >= 1);
        const uint32_t okL2 = __ballot_sync(0xffffffffu, incol && bL >= 2 && lmax >= 2);
        __syncwarp();

        const int m = n - a;
        // ---- feasible pattern sets (column budget, parent row ①)
        int npl = 0;
        {
            const int NPm = ambi_np_for(m);
            for (int base = 0; base < NPm; base += 32) {
                int ip = base + lane;
                bool ok = false;
                if (ip < NPm) {
                    if (ip == 0) ok = true;
                    else {
                        int q1 = s_ptab[2 * ip], q2 = s_ptab[2 * ip + 1];
                        int b1 = n - 1 - q1;
                        if (q2 == AMBI_NONE) ok = ((okP >> b1) & 1) && ((parP >> b1) & 1);
                        else { int b2 = n - 1 - q2; ok = ((okP >> b1) & 1) && ((okP >> b2) & 1) && ((parP >> b2) & 1); }
                    }
                }
                uint32_t bal = __ballot_sync(0xffffffffu, ok);
                if (ok) S.pl[npl + __popc(bal & lt_mask)] = (uint16_t)ip;
                npl += __popc(bal);
            }
        }
        // ---- feasible loop multisets (column budget); class 0 = parent row ③ already
        //      satisfied by earlier rows, class 1 = needs pattern ends to its right
        int nl0 = 0, nll = 0;
        {
            const int NLm = ambi_nl_for(m);
            for (int cls = 0; cls < 2; ++cls) {
                for (int base = 0; base < NLm; base += 32) {
                    int il = base + lane;
                    bool ok = false;
                    if (il < NLm) {
                        if (il == 0) ok = (cls == 0);
                        else {
                            int q1 = s_ltab[2 * il], q2 = s_ltab[2 * il + 1];
                            int l1 = n - 1 - q1, hi, need;
                            bool bud;
                            if (q2 == AMBI_NONE) { hi = l1; need = 1; bud = (okL1 >> l1) & 1; }
                            else if (q2 == q1) { hi = l1; need = 2; bud = (okL2 >> l1) & 1; }
                            else { hi = n - 1 - q2; need = 1; bud = ((okL1 >> l1) & 1) && ((okL1 >> hi) & 1); }
                            int par = (S.col[hi] >> 24);
                            bool c0 = (a == 0 && hi == n - 1) || par >= need;
                            ok = bud && (c0 == (cls == 0));
                        }
                    }
                    uint32_t bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) S.ll[nll + __popc(bal & lt_mask)] = (uint16_t)il;
                    nll += __popc(bal);
                }
                if (cls == 0) nl0 = nll;
            }
        }
        __syncwarp();

        // ---- enumerate the product, light test per lane, heavy test from a queue
        int qn = 0;                       // queue fill (warp-uniform)
        unsigned long long ncand = 0;     // per-lane count

        auto heavy = [&](uint32_t code, bool active) {
            // code = ip<<10 | il ; all lanes call, inactive lanes idle
            int p1 = -1, p2 = -1, l1 = -1, l2 = -1, err = 0, tot = 0;
            bool keep = false;
            if (active) {
                int ip = code >> 10, il = code & 1023;
                if (ip) { int q1 = s_ptab[2 * ip], q2 = s_ptab[2 * ip + 1];
                          p1 = n - 1 - q1; if (q2 != AMBI_NONE) p2 = n - 1 - q2; }
                if (il) { int q1 = s_ltab[2 * il], q2 = s_ltab[2 * il + 1];
                          l1 = n - 1 - q1; if (q2 != AMBI_NONE) l2 = n - 1 - q2; }
                int w = (p1 >= 0) + (p2 >= 0) + 2 * ((l1 >= 0) + (l2 >= 0));
                err = abs(2 * (covC + w) - S.c2[a]) + abs(fbC + w - S.j2[a]);
                int lb = child_bound(S, n, a, p1, p2, l1, l2, slack - err);
                tot = err + lb;
                keep = tot <= slack;
            }
            // ---- survivors are written one at a time by the whole warp
            uint32_t km = __ballot_sync(0xffffffffu, keep);
            while (km) {
                int src = __ffs(km) - 1; km &= km - 1;
                uint32_t ccode = __shfl_sync(0xffffffffu, code, src);
                int cp1 = __shfl_sync(0xffffffffu, p1, src), cp2 = __shfl_sync(0xffffffffu, p2, src);
                int cl1 = __shfl_sync(0xffffffffu, l1, src), cl2 = __shfl_sync(0xffffffffu, l2, src);
                int cerr = __shfl_sync(0xffffffffu, err, src), ctot = __shfl_sync(0xffffffffu, tot, src);
                // prefix value for this lane (row code of row `lane`)
                uint32_t myrow = lane < a ? prow[lane] : (lane == a ? ccode : 0u);
                if (A.shard_world > 1 && a == A.shard_level) {
                    uint32_t h = mix32(myrow + 0x9e3779b9u * (lane + 1));
#pragma unroll
                    for (int o = 16; o; o >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, o);
                    if ((int)(mix32(h) % (uint32_t)A.shard_world) != A.shard_rank) continue;
                }
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(A.ccount, 1u);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= A.cap) { if (lane == 0) atomicExch(A.overflow, 1u); continue; }
                uint8_t *crec = A.children + (size_t)slot * A.stride;
                // column byte of lane's column
                uint32_t nb = 0;
                if (lane > a && lane < n) {
                    int j = lane;
                    int pab = (j == cp1) + (j == cp2), lab = (j == cl1) + (j == cl2);
                    int rp = (cp1 >= 0 && cp1 < j) + (cp2 >= 0 && cp2 < j);
                    int rl = (cl1 >= 0 && cl1 < j) + (cl2 >= 0 && cl2 < j);
                    int kP = 2 - pab - max(rp, rl);
                    int kL = 2 - max(lab + rl, max(pab + rl, pab + rp));
                    int nbP = min(bP - pab, kP), nbL = min(bL - lab, kL);
                    nb = (uint32_t)((nP + pab) | ((nL + lab) << 2) | (nbP << 4) | (nbL << 6));
                }
                crec[lane] = (uint8_t)nb;
                if (lane == 0) {
                    NodeHdr *ch = reinterpret_cast<NodeHdr *>(crec);
                    int hits = 0;
                    if (tmask) hits = (cp1 >= 0 && ((tmask >> cp1) & 1)) + (cp2 >= 0 && ((tmask >> cp2) & 1)) +
                                      (cl1 >= 0 && ((tmask >> cl1) & 1)) + (cl2 >= 0 && ((tmask >> cl2) & 1));
                    ch->score = score + cerr; ch->f = score + ctot; ch->prob = prob; ch->tgs = ptgs + hits;
                }
                uint32_t *crow = reinterpret_cast<uint32_t *>(crec + sizeof(NodeHdr));
                if (lane < A.nmax) crow[lane] = myrow;
            }
        };

        for (int pi = 0; pi < npl; ++pi) {
            const int ip = S.pl[pi];
            int p1 = -1, p2 = -1;
            if (ip) { int q1 = s_ptab[2 * ip], q2 = s_ptab[2 * ip + 1];
                      p1 = n - 1 - q1; if (q2 != AMBI_NONE) p2 = n - 1 - q2; }
            const int pmax = p2 >= 0 ? p2 : p1;           // highest pattern end
            const int np = (p1 >= 0) + (p2 >= 0);
            const int range = ip ? nll : nl0;
            const int phits = tmask ? ((p1 >= 0 && ((tmask >> p1) & 1)) + (p2 >= 0 && ((tmask >> p2) & 1))) : 0;
            for (int base = 0; base < range; base += 32) {
                const int li = base + lane;
                bool pass = false;
                uint32_t code = 0;
                if (li < range) {
                    const int il = S.ll[li];
                    int l1 = -1, l2 = -1;            // l1 = lowest column, l2 = highest (or == l1 for mult 2)
                    if (il) { int q1 = s_ltab[2 * il], q2 = s_ltab[2 * il + 1];
                              l1 = n - 1 - q1; if (q2 != AMBI_NONE) l2 = n - 1 - q2; }
                    bool feas = true;
                    if (li >= nl0) {                 // class 1: parents must come from pattern ends to the right
                        int hi = l2 >= 0 ? l2 : l1;
                        int need = (l2 == l1) ? 2 : 1;
                        int par = (int)(S.col[hi] >> 24) + (p1 > hi) + (p2 > hi);
                        feas = par >= need;
                    }
                    // rows ②④⑤ inside the row: a pattern end strictly right of both loop entries
                    const int nl = (l1 >= 0) + (l2 >= 0);
                    const int lhi = l2 >= 0 ? l2 : l1;
                    if (nl == 2 && pmax > lhi) feas = false;
                    if (feas && tmask) {
                        int hits = phits + (l1 >= 0 && ((tmask >> l1) & 1)) + (l2 >= 0 && ((tmask >> l2) & 1));
                        if (ptgs + hits > 5) feas = false;
                    }
                    if (feas) {
                        ++ncand;
                        int w = np + 2 * nl;
                        int err = abs(2 * (covC + w) - S.c2[a]) + abs(fbC + w - S.j2[a]);
                        pass = err <= slack;
                        code = ((uint32_t)ip << 10) | (uint32_t)il;
                    }
                }
                uint32_t bal = __ballot_sync(0xffffffffu, pass);
                if (pass) S.queue[qn + __popc(bal & lt_mask)] = code;
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) {
                    uint32_t c = S.queue[lane];
                    __syncwarp();
                    heavy(c, true);
                    qn -= 32;
                    uint32_t mv = (lane < qn) ? S.queue[32 + lane] : 0u;
                    __syncwarp();
                    if (lane < qn) S.queue[lane] = mv;
                    __syncwarp();
                }
            }
        }
        if (qn > 0) {
            uint32_t c = lane < qn ? S.queue[lane] : 0u;
            __syncwarp();
            heavy(c, lane < qn);
        }
        // ---- statistics
#pragma unroll
        for (int o = 16; o; o >>= 1) ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
        if (lane == 0) {
            atomicAdd(&A.cand_count[prob], ncand);
            atomicAdd(&A.node_count[prob], 1ull);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// K3a: merge records that reached the same column state.
// ---------------------------------------------------------------------------
struct DedupArgs {
    uint8_t *recs;
    const uint32_t *count;
    uint32_t cap;
    unsigned long long *table;
    uint32_t tmask;
    uint32_t gen;
    uint8_t *dead;
    int stride;
    int level;          // rows 0..level are fixed in these records
};

__device__ __forceinline__ bool same_state(const NodeHdr *x, const NodeHdr *y) {
    if (x->prob != y->prob || x->tgs != y->tgs) return false;
    const uint4 *a = reinterpret_cast<const uint4 *>(x->col), *b = reinterpret_cast<const uint4 *>(y->col);
    uint4 a0 = a[0], b0 = b[0], a1 = a[1], b1 = b[1];
    return a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w &&
           a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
}

// strict order: lower score, then smaller prefix in search order
__device__ __forceinline__ bool rec_better(const uint8_t *x, const uint8_t *y, int nrows) {
    const NodeHdr *hx = reinterpret_cast<const NodeHdr *>(x), *hy = reinterpret_cast<const NodeHdr *>(y);
    if (hx->score != hy->score) return hx->score < hy->score;
    const uint32_t *rx = reinterpret_cast<const uint32_t *>(x + sizeof(NodeHdr));
    const uint32_t *ry = reinterpret_cast<const uint32_t *>(y + sizeof(NodeHdr));
    for (int i = 0; i < nrows; ++i) if (rx[i] != ry[i]) return rx[i] < ry[i];
    return false;
}

__global__ void k_dedup(DedupArgs A)
{
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint8_t *rec = A.recs + (size_t)i * A.stride;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(rec);
        const uint32_t *cw = reinterpret_cast<const uint32_t *>(h->col);
        uint32_t hv = mix32((uint32_t)h->prob * 0x9e3779b9u + (uint32_t)h->tgs);
#pragma unroll
        for (int k = 0; k < 8; ++k) hv = mix32(hv ^ cw[k]) + 0x85ebca6bu * (k + 1);
        uint32_t slot = hv & A.tmask;
        const unsigned long long mine = ((unsigned long long)A.gen << 32) | i;
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&A.table[slot]);
            if ((uint32_t)(cur >> 32) != A.gen) {
                unsigned long long old = atomicCAS(&A.table[slot], cur, mine);
                if (old == cur) break;
                continue;                              // re-read the same slot
            }
            uint32_t o = (uint32_t)cur;
            const uint8_t *orec = A.recs + (size_t)o * A.stride;
            if (same_state(h, reinterpret_cast<const NodeHdr *>(orec))) {
                if (rec_better(rec, orec, A.level + 1)) {
                    unsigned long long old = atomicCAS(&A.table[slot], cur, mine);
                    if (old == cur) { A.dead[o] = 1; break; }
                    continue;
                } else { A.dead[i] = 1; break; }
            }
            slot = (slot + 1) & A.tmask;
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
    const uint8_t *dead;
    uint32_t *next_list;
    uint32_t *next_count;
    int32_t *best_score;        // [prob]
    uint32_t *best_rows;        // [prob][AMBI_MAXN]
    int32_t *fmin;              // beam: per-problem min f (or nullptr)
    int stride;
    int level;
};

__global__ void k_finalize(FinalizeArgs A)
{
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        if (A.dead[i]) continue;
        const uint8_t *rec = A.recs + (size_t)i * A.stride;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(rec);
        const int n = A.probs[h->prob].n;
        if (n == A.level + 1) {
            // all completed assignments of a problem share the empty column state, so the
            // merge above left exactly one alive: the arg-min in (score, search order)
            A.best_score[h->prob] = h->score;
            const uint32_t *r = reinterpret_cast<const uint32_t *>(rec + sizeof(NodeHdr));
            uint32_t *o = A.best_rows + (size_t)h->prob * AMBI_MAXN;
            for (int k = 0; k < n; ++k) o[k] = r[k];
        } else {
            if (A.fmin) atomicMin(&A.fmin[h->prob], h->f);
            else { uint32_t s = atomicAdd(A.next_count, 1u); A.next_list[s] = i; }
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
    const uint8_t *dead;
    const int32_t *fmin;
    uint32_t *hist;             // [prob][BEAM_BINS]
    uint32_t *ticket;           // [prob]
    uint32_t *next_list;
    uint32_t *next_count;
    int width;
    int stride;
    int level;
};

__global__ void k_beam_hist(BeamArgs A)
{
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        if (A.dead[i]) continue;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(A.recs + (size_t)i * A.stride);
        if (A.probs[h->prob].n == A.level + 1) continue;
        int bin = min(h->f - A.fmin[h->prob], BEAM_BINS - 1);
        atomicAdd(&A.hist[(size_t)h->prob * BEAM_BINS + bin], 1u);
    }
}

__global__ void k_beam_select(BeamArgs A)
{
    uint32_t cnt = min(*A.count, A.cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        if (A.dead[i]) continue;
        const NodeHdr *h = reinterpret_cast<const NodeHdr *>(A.recs + (size_t)i * A.stride);
        if (A.probs[h->prob].n == A.level + 1) continue;
        int bin = min(h->f - A.fmin[h->prob], BEAM_BINS - 1);
        const uint32_t *hs = A.hist + (size_t)h->prob * BEAM_BINS;
        uint32_t below = 0;
        for (int k = 0; k < bin; ++k) below += hs[k];
        bool keep;
        if (below + hs[bin] <= (uint32_t)A.width) keep = true;
        else if (below >= (uint32_t)A.width) keep = false;
        else keep = below + atomicAdd(&A.ticket[h->prob], 1u) < (uint32_t)A.width;
        if (keep) { uint32_t s = atomicAdd(A.next_count, 1u); A.next_list[s] = i; }
    }
}

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__global__ void k_init_roots(const DevProblem *probs, int count, uint8_t *recs, int stride, int nmax,
                             uint32_t *list, uint32_t *lcount, int32_t *limit, int32_t *best_score, int first)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    const DevProblem &P = probs[p];
    if (first) { limit[p] = P.zero_score; best_score[p] = 0x7fffffff; }
    if (P.n < 2) return;                  // n == 1 handled by k_solve_n1
    uint32_t s = atomicAdd(lcount, 1u);
    uint8_t *rec = recs + (size_t)s * stride;
    NodeHdr *h = reinterpret_cast<NodeHdr *>(rec);
    for (int b = 0; b < AMBI_MAXN; ++b) h->col[b] = b < P.n ? (uint8_t)((AMBI_BUD_INF << 4) | (AMBI_BUD_INF << 6)) : 0;
    h->score = 0; h->f = 0; h->prob = p; h->tgs = 0;
    uint32_t *r = reinterpret_cast<uint32_t *>(rec + sizeof(NodeHdr));
    for (int k = 0; k < nmax; ++k) r[k] = 0;
    list[s] = s;
}

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
            int e = abs(2 * (pp + 2 * l) - P.c2[0]) + abs(2 * l - P.j2[0]);
            if (e < best) { best = e; bp = pp; bl = l; }
        }
    best_score[p] = best;
    best_rows[(size_t)p * AMBI_MAXN] = 0x80000000u | ((uint32_t)bp << 20) | (uint32_t)bl;   // raw (P,L) for n==1
    cand_count[p] += 2ull * (unsigned long long)(P.loop_ub + 1);
}

__global__ void k_set_limit(const DevProblem *probs, int count, int32_t *limit, const int32_t *best_score,
                            int32_t *incumbent)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    int z = probs[p].zero_score, b = best_score[p];
    int l = b < z ? b : z;
    limit[p] = l; incumbent[p] = l;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    snprintf(err_, sizeof err_, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return -4; } } while (0)

static uint32_t next_pow2(uint64_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

StageA::StageA() { err_[0] = 0; }
StageA::~StageA() { release(); }

void StageA::release()
{
    cudaFree(buf_[0]); cudaFree(buf_[1]); cudaFree(list_[0]); cudaFree(list_[1]); cudaFree(dead_);
    cudaFree(table_); cudaFree(counters_); cudaFree(ptab_); cudaFree(ltab_);
    buf_[0] = buf_[1] = nullptr; list_[0] = list_[1] = nullptr; dead_ = nullptr; table_ = nullptr;
    counters_ = nullptr; ptab_ = ltab_ = nullptr;
    free_problem_arrays();
    for (int i = 0; i < 4; ++i) if (ev_[i]) { cudaEventDestroy(ev_[i]); ev_[i] = nullptr; }
    for (cudaEvent_t e : prof_ev_) cudaEventDestroy(e);
    prof_ev_.clear();
}

void StageA::free_problem_arrays()
{
    cudaFree(d_probs_); cudaFree(d_limit_); cudaFree(d_best_score_); cudaFree(d_best_rows_);
    cudaFree(d_beam_score_); cudaFree(d_beam_rows_); cudaFree(d_cand_); cudaFree(d_node_);
    cudaFree(d_fmin_); cudaFree(d_hist_); cudaFree(d_ticket_); cudaFree(d_incumbent_);
    d_probs_ = nullptr; d_limit_ = d_best_score_ = d_beam_score_ = d_fmin_ = d_incumbent_ = nullptr;
    d_best_rows_ = d_beam_rows_ = d_hist_ = d_ticket_ = nullptr; d_cand_ = d_node_ = nullptr;
    prob_cap_ = 0;
}

int StageA::init(cudaStream_t stream, int64_t node_capacity, int beam_width, int rank, int world, int shard_level)
{
    stream_ = stream;
    cap_ = node_capacity > 0 ? (uint32_t)std::min<int64_t>(node_capacity, 1ll << 30) : (4u << 20);
    beam_ = beam_width > 0 ? beam_width : 64;
    rank_ = rank; world_ = world > 0 ? world : 1; shard_level_ = shard_level;
    stride_ = (int)ambi_node_stride(AMBI_MAXN);
    CK(cudaMalloc(&buf_[0], (size_t)cap_ * stride_));
    CK(cudaMalloc(&buf_[1], (size_t)cap_ * stride_));
    CK(cudaMalloc(&list_[0], (size_t)cap_ * 4));
    CK(cudaMalloc(&list_[1], (size_t)cap_ * 4));
    CK(cudaMalloc(&dead_, (size_t)cap_));
    tsize_ = next_pow2((uint64_t)cap_ * 2);
    CK(cudaMalloc(&table_, (size_t)tsize_ * 8));
    CK(cudaMemsetAsync(table_, 0, (size_t)tsize_ * 8, stream_));
    CK(cudaMalloc(&counters_, 64 * sizeof(uint32_t)));
    // decode tables
    std::vector<uint8_t> pt(AMBI_NP_CODES * 2, AMBI_NONE), lt(AMBI_NL_CODES * 2, AMBI_NONE);
    for (int q1 = 0; q1 < AMBI_MAXN; ++q1) {
        int ip = ambi_ip_code(q1, -1); pt[2 * ip] = q1; pt[2 * ip + 1] = AMBI_NONE;
        for (int q2 = 0; q2 < q1; ++q2) { ip = ambi_ip_code(q1, q2); pt[2 * ip] = q1; pt[2 * ip + 1] = q2; }
        int il = ambi_il_code(q1, -1); lt[2 * il] = q1; lt[2 * il + 1] = AMBI_NONE;
        for (int q2 = 0; q2 <= q1; ++q2) { il = ambi_il_code(q1, q2); lt[2 * il] = q1; lt[2 * il + 1] = q2; }
    }
    CK(cudaMalloc(&ptab_, pt.size())); CK(cudaMalloc(&ltab_, lt.size()));
    CK(cudaMemcpyAsync(ptab_, pt.data(), pt.size(), cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(ltab_, lt.data(), lt.size(), cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ev_[i]));
    int dev = 0; CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev));
    gen_ = 1;
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
    CK(cudaMalloc(&d_fmin_, (size_t)c * 4)); CK(cudaMalloc(&d_hist_, (size_t)c * BEAM_BINS * 4));
    CK(cudaMalloc(&d_ticket_, (size_t)c * 4));
    prob_cap_ = c;
    return 0;
}

int StageA::upload(const DevProblem *host_probs, int count)
{
    if (int rc = ensure_problem_arrays(count)) return rc;
    count_ = count; nmax_ = 0;
    for (int i = 0; i < count; ++i) nmax_ = std::max(nmax_, (int)host_probs[i].n);
    CK(cudaMemcpyAsync(d_probs_, host_probs, (size_t)count * sizeof(DevProblem), cudaMemcpyHostToDevice, stream_));
    return 0;
}

// One level: expand -> merge -> finalize (-> beam selection).
int StageA::run_level(int level, int cur, bool beam, int shard_world)
{
    uint32_t *c_work = counters_ + 0, *c_child = counters_ + 1, *c_next = counters_ + 2, *c_over = counters_ + 3;
    uint32_t *c_cur = counters_ + 4 + (level & 1);      // frontier count of this level
    uint32_t *c_nxt = counters_ + 4 + ((level + 1) & 1);
    (void)c_next;
    CK(cudaMemsetAsync(c_work, 0, 8, stream_));         // work + child counters
    CK(cudaMemsetAsync(c_nxt, 0, 4, stream_));
    CK(cudaMemsetAsync(dead_, 0, cap_, stream_));
    LevelArgs L;
    L.probs = d_probs_; L.parents = buf_[cur]; L.plist = list_[cur]; L.pcount = c_cur;
    L.children = buf_[cur ^ 1]; L.ccount = c_child; L.cap = cap_; L.limit = d_limit_;
    L.cand_count = d_cand_; L.node_count = d_node_; L.work_counter = c_work; L.overflow = c_over;
    L.ptab = ptab_; L.ltab = ltab_; L.level = level; L.stride = stride_; L.nmax = AMBI_MAXN;
    L.shard_world = shard_world; L.shard_rank = rank_; L.shard_level = shard_level_eff_;
    if (profile_) {
        if (prof_used_ + 2 > prof_ev_.size()) { cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); prof_ev_.push_back(e0); prof_ev_.push_back(e1); }
        CK(cudaEventRecord(prof_ev_[prof_used_], stream_));
    }
    k_expand<<<sms_ * 4, WPB * 32, 0, stream_>>>(L); ++launches_;
    if (profile_) { CK(cudaEventRecord(prof_ev_[prof_used_ + 1], stream_)); prof_used_ += 2; }
    DedupArgs D;
    D.recs = buf_[cur ^ 1]; D.count = c_child; D.cap = cap_; D.table = table_; D.tmask = tsize_ - 1;
    D.gen = ++gen_; D.dead = dead_; D.stride = stride_; D.level = level;
    k_dedup<<<sms_ * 8, 256, 0, stream_>>>(D); ++launches_;
    FinalizeArgs F;
    F.probs = d_probs_; F.recs = buf_[cur ^ 1]; F.count = c_child; F.cap = cap_; F.dead = dead_;
    F.next_list = list_[cur ^ 1]; F.next_count = c_nxt;
    F.best_score = beam ? d_beam_score_ : d_best_score_; F.best_rows = beam ? d_beam_rows_ : d_best_rows_;
    F.fmin = nullptr; F.stride = stride_; F.level = level;
    if (beam) {
        CK(cudaMemsetAsync(d_fmin_, 0x7f, (size_t)count_ * 4, stream_));
        CK(cudaMemsetAsync(d_hist_, 0, (size_t)count_ * BEAM_BINS * 4, stream_));
        CK(cudaMemsetAsync(d_ticket_, 0, (size_t)count_ * 4, stream_));
        F.fmin = d_fmin_;
    }
    k_finalize<<<sms_ * 8, 256, 0, stream_>>>(F); ++launches_;
    if (beam) {
        BeamArgs B;
        B.probs = d_probs_; B.recs = buf_[cur ^ 1]; B.count = c_child; B.cap = cap_; B.dead = dead_;
        B.fmin = d_fmin_; B.hist = d_hist_; B.ticket = d_ticket_; B.next_list = list_[cur ^ 1];
        B.next_count = c_nxt; B.width = beam_; B.stride = stride_; B.level = level;
        k_beam_hist<<<sms_ * 8, 256, 0, stream_>>>(B); ++launches_;
        k_beam_select<<<sms_ * 8, 256, 0, stream_>>>(B); ++launches_;
    }
    return 0;
}

int StageA::run(bool sharded)
{
    const int tpb = 128, blocks = (count_ + tpb - 1) / tpb;
    uint32_t *c_over = counters_ + 3;
    prof_used_ = 0;
    CK(cudaEventRecord(ev_[0], stream_));
    CK(cudaMemsetAsync(counters_, 0, 64 * sizeof(uint32_t), stream_));
    CK(cudaMemsetAsync(d_cand_, 0, (size_t)count_ * 8, stream_));
    CK(cudaMemsetAsync(d_node_, 0, (size_t)count_ * 8, stream_));
    CK(cudaMemsetAsync(d_beam_score_, 0x7f, (size_t)count_ * 4, stream_));
    // ---- pass 1: incumbent (beam) ------------------------------------------
    k_init_roots<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, buf_[0], stride_, AMBI_MAXN, list_[0],
                                              counters_ + 4, d_limit_, d_best_score_, 1); ++launches_;
    k_solve_n1<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, d_best_score_, d_best_rows_, d_cand_); ++launches_;
    shard_level_eff_ = -1;
    int cur = 0;
    for (int a = 0; a < nmax_; ++a) { if (int rc = run_level(a, cur, true, 1)) return rc; cur ^= 1; }
    k_set_limit<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, d_limit_, d_beam_score_, d_incumbent_); ++launches_;
    CK(cudaEventRecord(ev_[1], stream_));
    // ---- pass 2: exact ------------------------------------------------------
    // roots again (buffer 0 head was overwritten by children of later levels)
    CK(cudaMemsetAsync(counters_ + 4, 0, 8, stream_));
    k_init_roots<<<blocks, tpb, 0, stream_>>>(d_probs_, count_, buf_[0], stride_, AMBI_MAXN, list_[0],
                                              counters_ + 4, d_limit_, d_best_score_, 0); ++launches_;
    int sw = sharded ? world_ : 1;
    shard_level_eff_ = -1;
    if (sw > 1) {
        int sl = shard_level_ > 0 ? shard_level_ : 2;
        shard_level_eff_ = std::min(sl, std::max(0, nmax_ - 2));
    }
    cur = 0;
    for (int a = 0; a < nmax_; ++a) { if (int rc = run_level(a, cur, false, sw)) return rc; cur ^= 1; }
    CK(cudaEventRecord(ev_[2], stream_));
    (void)c_over;
    return 0;
}

int StageA::fetch(HostResult *out)
{
    std::vector<int32_t> score(count_), inc(count_);
    std::vector<uint32_t> rows((size_t)count_ * AMBI_MAXN);
    std::vector<unsigned long long> cand(count_), node(count_);
    uint32_t ctr[8];
    CK(cudaMemcpyAsync(score.data(), d_best_score_, (size_t)count_ * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(inc.data(), d_incumbent_, (size_t)count_ * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(rows.data(), d_best_rows_, rows.size() * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(cand.data(), d_cand_, (size_t)count_ * 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(node.data(), d_node_, (size_t)count_ * 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(ctr, counters_, sizeof ctr, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    float ms_beam = 0, ms_all = 0;
    cudaEventElapsedTime(&ms_beam, ev_[0], ev_[1]);
    cudaEventElapsedTime(&ms_all, ev_[0], ev_[2]);
    prof_total_ms_ = ms_all; prof_beam_ms_ = ms_beam;
    if (profile_) {
        prof_expand_ms_ = 0; prof_expand_launches_ = prof_used_ / 2;
        for (size_t i = 0; i + 1 < prof_used_; i += 2) { float m = 0; cudaEventElapsedTime(&m, prof_ev_[i], prof_ev_[i + 1]); prof_expand_ms_ += m; }
    }
    const bool overflow = ctr[3] != 0;
    for (int p = 0; p < count_; ++p) {
        out[p].score2 = score[p]; out[p].incumbent = inc[p];
        out[p].candidates = cand[p]; out[p].nodes = node[p];
        memcpy(out[p].rows, &rows[(size_t)p * AMBI_MAXN], AMBI_MAXN * 4);
        out[p].seconds = ms_all * 1e-3; out[p].seconds_to_best = (inc[p] == score[p] ? ms_beam : ms_all) * 1e-3;
        out[p].status = overflow ? -5 : (score[p] == 0x7fffffff ? -4 : 0);
    }
    if (overflow) snprintf(err_, sizeof err_, "frontier overflow: more than %u candidate records in one level", cap_);
    return overflow ? -5 : 0;
}

}  // namespace ambi
