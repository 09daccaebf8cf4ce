// stage_b.cu -- Stage B of the BFB search on sm_100a (kernel K4).
//
// Replaces allTopologicalOrders + the order loop of getBFB (reference
// src/LocalGenomicMap.cpp:3380-3409, 3514-3697).  The reference materialises
// every topological order of the component DAG and assembles each one from
// scratch; the first order (lexicographic in node numbering, lowest free index
// first, LGM:3389-3391) whose components can all be placed wins.  Placing
// order[i] only depends on order[0..i-1], so a failed placement prunes the
// whole subtree: the device walks the tree of *assemblable prefixes* instead.
//
// One warp per DAG.  The (first, second) prefixes are numbered t = 0,1,... in
// lexicographic order and dealt round-robin to the 32 lanes; each lane runs a
// depth-first search below its prefixes with the breakpoint path kept in local
// memory as signed int8 segment offsets and an O(1) undo record per depth.
// The winner is the solution with the smallest t (shared-memory atomicMin),
// which is the lexicographically first valid order.
#include "stage_b.h"
#include "path_surgery.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

namespace ambi {

constexpr int MAXK = AMBI_MAX_COMPONENTS;
constexpr int MAXPATH = 1024;

struct DevDag {
    int32_t k, base, reversed, pad;
    int8_t kind[MAXK];
    int8_t a[MAXK], b[MAXK];               // relative ids (id - base + 1 >= 1)
    int16_t cn[MAXK];                      // a one-segment partition may hold a loop of multiplicity up to loop_ub
    uint64_t adj[MAXK];
};

struct DevOrderOut {                       // compact: 1.1 KB per DAG crosses PCIe / NVLink-C2C
    int32_t found, used_flipped, bkp_len, pad;
    unsigned long long tested;
    uint8_t order[MAXK];
    int8_t bkp[MAXPATH];                   // relative signed ids, see DevDag
};

struct Undo { int16_t f; int16_t cnt; int8_t old; int8_t had; };

// try to place node v on the breakpoint path (getBFB body, LGM:3572-3644)
__device__ __forceinline__ bool place(const DevDag &D, int v, int8_t *B, int &len, Undo &u)
{
    const int a = D.a[v], b = D.b[v];
    if (D.kind[v] == 0) {                              // pattern: must extend the tail
        const int tail = B[len - 1];
        if (len + 2 > MAXPATH) return false;
        if (tail == -a) { B[len] = (int8_t)a; B[len + 1] = (int8_t)b; }
        else if (tail == b) { B[len] = (int8_t)-b; B[len + 1] = (int8_t)-a; }
        else return false;
        len += 2; u.f = -1; u.cnt = 2; u.had = 0; u.old = 0;
        return true;
    }
    // loop: find a run end (odd index) equal to a- (then b+) with the nesting test
    int f = -1; bool first = true;
    for (int i = len - 1; i >= 1; i -= 2)
        if (B[i] == -a && (i == len - 1 || abs((int)B[i - 1]) >= abs((int)B[i + 2]))) { f = i; break; }
    if (f < 0) {
        first = false;
        for (int i = len - 1; i >= 1; i -= 2)
            if (B[i] == b && (i == len - 1 || abs((int)B[i - 1]) <= abs((int)B[i + 2]))) { f = i; break; }
    }
    if (f < 0) return false;
    const int cn = D.cn[v], add = 4 * cn;
    if (len + add > MAXPATH) return false;
    u.f = (int16_t)f; u.cnt = (int16_t)add; u.had = 0; u.old = 0;
    if (f + 1 < len) { u.had = 1; u.old = B[f + 1]; B[f + 1] = (int8_t)(first ? a : -b); }
    for (int i = len - 1; i > f; --i) B[i + add] = B[i];
    for (int c = 0; c < cn; ++c) {
        int8_t *q = B + f + 1 + 4 * c;
        if (first) { q[0] = (int8_t)a; q[1] = (int8_t)b; q[2] = (int8_t)-b; q[3] = (int8_t)-a; }
        else       { q[0] = (int8_t)-b; q[1] = (int8_t)-a; q[2] = (int8_t)a; q[3] = (int8_t)b; }
    }
    len += add;
    return true;
}

__device__ __forceinline__ void unplace(int8_t *B, int &len, const Undo &u)
{
    if (u.f < 0) { len -= u.cnt; return; }
    const int f = u.f, add = u.cnt;
    for (int i = f + 1 + add; i < len; ++i) B[i - add] = B[i];
    len -= add;
    if (u.had) B[f + 1] = u.old;
}

// seed the path with the first node (LGM:3524-3568)
__device__ __forceinline__ void seed(const DevDag &D, int v, bool forward, int8_t *B, int &len)
{
    const int a = D.a[v], b = D.b[v];
    len = 0;
    if (D.kind[v] == 0) {
        if (forward) { B[0] = (int8_t)a; B[1] = (int8_t)b; } else { B[0] = (int8_t)-b; B[1] = (int8_t)-a; }
        len = 2;
    } else {
        for (int c = 0; c < D.cn[v] && len + 4 <= MAXPATH; ++c) {   // the host refuses 4*cn > MAXPATH (to_devdag)
            if (forward) { B[len] = (int8_t)a; B[len + 1] = (int8_t)b; B[len + 2] = (int8_t)-b; B[len + 3] = (int8_t)-a; }
            else         { B[len] = (int8_t)-b; B[len + 1] = (int8_t)-a; B[len + 2] = (int8_t)a; B[len + 3] = (int8_t)b; }
            len += 4;
        }
    }
}

__global__ void __launch_bounds__(128)
k_first_valid_order(const DevDag *dags, DevOrderOut *outs, int count, uint32_t *work)
{
    __shared__ DevDag s_dag[4];
    __shared__ int s_best[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    DevDag &D = s_dag[wib];

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(work, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= (uint32_t)count) break;
        __syncwarp();
        {   // stage the DAG in shared memory
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&dags[wi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&D);
            for (int i = lane; i < (int)(sizeof(DevDag) / 4); i += 32) dst[i] = src[i];
        }
        __syncwarp();
        const int k = D.k;
        DevOrderOut &O = outs[wi];
        unsigned long long tested = 0;
        int found_flip = -1;

        int8_t B[MAXPATH];
        uint8_t order[MAXK], indeg[MAXK], nexti[MAXK];
        Undo undo[MAXK];

        for (int attempt = 0; attempt < 2 && found_flip < 0 && k > 0; ++attempt) {
            const bool forward = (attempt == 0) ? !D.reversed : (D.reversed != 0);   // LGM:3516, 3691-3695
            if (lane == 0) s_best[wib] = 0x7fffffff;
            __syncwarp();
            // indegrees (localhap.cpp:244-250)
            for (int i = 0; i < k; ++i) indeg[i] = 0;
            for (int i = 0; i < k; ++i) { uint64_t m = D.adj[i]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
            uint64_t visited = 0;
            int len = 0, my_t = -1;
            bool have = false;
            int t = 0;
            // enumerate (first, second) prefixes in lexicographic order
            for (int v0 = 0; v0 < k && !have; ++v0) {
                if (indeg[v0] != 0) continue;
                if (k == 1) {
                    if (lane == 0) { seed(D, v0, forward, B, len); order[0] = (uint8_t)v0; ++tested; have = true; my_t = 0; atomicMin(&s_best[wib], 0); }
                    break;
                }
                // remove v0
                { uint64_t m = D.adj[v0]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
                visited |= 1ull << v0;
                for (int v1 = 0; v1 < k && !have; ++v1) {
                    if (indeg[v1] != 0 || ((visited >> v1) & 1)) continue;
                    const int tt = t++;
                    if ((tt & 31) != lane) continue;
                    if (tt > *(volatile int *)&s_best[wib]) continue;
                    // assemble the two-node prefix
                    seed(D, v0, forward, B, len);
                    order[0] = (uint8_t)v0;
                    ++tested;
                    if (!place(D, v1, B, len, undo[1])) continue;
                    order[1] = (uint8_t)v1;
                    if (k == 2) { have = true; my_t = tt; atomicMin(&s_best[wib], tt); break; }
                    { uint64_t m = D.adj[v1]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
                    visited |= 1ull << v1;
                    // depth-first search below, lowest free index first (LGM:3389-3391)
                    int d = 2; nexti[2] = 0;
                    unsigned steps = 0;
                    bool aborted = false;
                    while (d >= 2) {
                        if ((++steps & 63) == 0 && tt > *(volatile int *)&s_best[wib]) { aborted = true; break; }
                        int v = nexti[d];
                        while (v < k && (indeg[v] != 0 || ((visited >> v) & 1))) ++v;
                        if (v >= k) {                          // exhausted: backtrack
                            --d;
                            if (d < 2) break;
                            int pv = order[d];
                            unplace(B, len, undo[d]);
                            visited &= ~(1ull << pv);
                            { uint64_t m = D.adj[pv]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
                            continue;
                        }
                        nexti[d] = (uint8_t)(v + 1);
                        ++tested;
                        if (!place(D, v, B, len, undo[d])) continue;
                        order[d] = (uint8_t)v;
                        if (d == k - 1) { have = true; my_t = tt; atomicMin(&s_best[wib], tt); break; }
                        visited |= 1ull << v;
                        { uint64_t m = D.adj[v]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
                        ++d; nexti[d] = 0;
                    }
                    if (have) break;
                    if (aborted) {
                        // unwind bookkeeping to the two-node prefix
                        for (int dd = d - 1; dd >= 2; --dd) {
                            int pv = order[dd];
                            visited &= ~(1ull << pv);
                            uint64_t m = D.adj[pv]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; }
                        }
                    }
                    // restore v1
                    visited &= ~(1ull << v1);
                    { uint64_t m = D.adj[v1]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
                }
                if (have) break;
                visited &= ~(1ull << v0);
                { uint64_t m = D.adj[v0]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
            }
            __syncwarp();
            const int best = s_best[wib];
            if (best != 0x7fffffff) {
                found_flip = attempt;
                if (have && my_t == best) {
                    O.found = 1; O.used_flipped = attempt; O.bkp_len = len;
                    for (int i = 0; i < k; ++i) O.order[i] = order[i];
                    for (int i = 0; i < len; ++i) O.bkp[i] = B[i];
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) tested += __shfl_xor_sync(0xffffffffu, tested, o);
        if (lane == 0) {
            O.tested = tested;
            if (found_flip < 0) { O.found = 0; O.used_flipped = 0; O.bkp_len = 0; }
        }
        __syncwarp();
    }
}


// --all (printAll, LGM:3672-3685): every order from which the path assembles, in one orientation.
// Same walk as above without the early exit; each complete order is appended to `orders`
// (unordered: the host sorts them lexicographically, which is the order the reference prints).
__global__ void __launch_bounds__(128)
k_all_valid_orders(const DevDag *dag, int forward, int32_t *orders, int8_t *bkps, int32_t *bkp_lens, uint32_t max_orders,
                   uint32_t *n_found, unsigned long long *n_tested)
{
    __shared__ DevDag s_dag;
    const int lane = threadIdx.x & 31;
    const int worker = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp id in the grid
    const int nworkers = gridDim.x * (blockDim.x >> 5) * 32;                    // lanes in the grid
    const int me = worker * 32 + lane;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(dag);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_dag);
        for (int i = threadIdx.x; i < (int)(sizeof(DevDag) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevDag &D = s_dag;
    const int k = D.k;
    if (k <= 0) return;
    int8_t B[MAXPATH];
    uint8_t order[MAXK], indeg[MAXK], nexti[MAXK];
    Undo undo[MAXK];
    unsigned long long tested = 0;
    for (int i = 0; i < k; ++i) indeg[i] = 0;
    for (int i = 0; i < k; ++i) { uint64_t m = D.adj[i]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
    uint64_t visited = 0;
    int len = 0, t = 0;
    auto emit = [&]() {
        uint32_t s = atomicAdd(n_found, 1u);
        if (s < max_orders) {
            for (int i = 0; i < k; ++i) orders[(size_t)s * k + i] = order[i];
            bkp_lens[s] = len;
            for (int i = 0; i < len; ++i) bkps[(size_t)s * MAXPATH + i] = B[i];
        }
    };
    for (int v0 = 0; v0 < k; ++v0) {
        if (indeg[v0] != 0) continue;
        if (k == 1) { if (me == 0) { seed(D, v0, forward != 0, B, len); order[0] = (uint8_t)v0; ++tested; emit(); } break; }
        { uint64_t m = D.adj[v0]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
        visited |= 1ull << v0;
        for (int v1 = 0; v1 < k; ++v1) {
            if (indeg[v1] != 0 || ((visited >> v1) & 1)) continue;
            const int tt = t++;
            if (tt % nworkers != me) continue;
            seed(D, v0, forward != 0, B, len);
            order[0] = (uint8_t)v0;
            ++tested;
            if (!place(D, v1, B, len, undo[1])) continue;
            order[1] = (uint8_t)v1;
            if (k == 2) { emit(); continue; }
            { uint64_t m = D.adj[v1]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
            visited |= 1ull << v1;
            int d = 2; nexti[2] = 0;
            while (d >= 2) {
                int v = nexti[d];
                while (v < k && (indeg[v] != 0 || ((visited >> v) & 1))) ++v;
                if (v >= k) {
                    --d;
                    if (d < 2) break;
                    int pv = order[d];
                    unplace(B, len, undo[d]);
                    visited &= ~(1ull << pv);
                    { uint64_t m = D.adj[pv]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
                    continue;
                }
                nexti[d] = (uint8_t)(v + 1);
                ++tested;
                if (!place(D, v, B, len, undo[d])) continue;
                order[d] = (uint8_t)v;
                if (d == k - 1) { emit(); unplace(B, len, undo[d]); continue; }
                visited |= 1ull << v;
                { uint64_t m = D.adj[v]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]--; } }
                ++d; nexti[d] = 0;
            }
            visited &= ~(1ull << v1);
            { uint64_t m = D.adj[v1]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
        }
        visited &= ~(1ull << v0);
        { uint64_t m = D.adj[v0]; while (m) { int j = __ffsll((long long)m) - 1; m &= m - 1; indeg[j]++; } }
    }
    if (tested) atomicAdd(n_tested, tested);
}

// --all post-processing on the device (getBFB after an order has assembled, LGM:3656-3670): one thread per valid order
// turns its breakpoint path (relative int8 ids) into absolute signed ids, applies imperfectFBI and expands the pairs
// to the vertex path the reference prints.
__global__ void k_path_surgery(const int8_t *bkps, const int32_t *bkp_lens, uint32_t n_orders, int base,
                               const int16_t *inv_src, const int16_t *inv_tgt, int max_id,
                               int16_t *scratch /*[n][MAXPATH]*/, int16_t *paths /*[n][stride]*/, int32_t *path_lens, int stride)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_orders) return;
    const int len = bkp_lens[i];
    int16_t *b = scratch + (size_t)i * MAXPATH;
    for (int t = 0; t < len; ++t) { const int v = bkps[(size_t)i * MAXPATH + t]; b[t] = (int16_t)(v > 0 ? v - 1 + base : -(-v - 1 + base)); }
    imperfect_fbi(b, len, inv_src, inv_tgt, max_id);
    path_lens[i] = expand_path(b, len, paths + (size_t)i * stride, stride);
}

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    snprintf(err_, sizeof err_, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return -4; } } while (0)

StageB::StageB() { err_[0] = 0; }
StageB::~StageB() { release(); }
void StageB::release()
{
    cudaFree(d_dags_); cudaFree(d_out_); cudaFree(d_work_); cudaFree(d_all_); cudaFreeHost(h_dags_); cudaFreeHost(h_out_);
    for (int i = 0; i < 2; ++i) if (ev_[i]) { cudaEventDestroy(ev_[i]); ev_[i] = nullptr; }
    d_dags_ = d_out_ = d_work_ = d_all_ = h_dags_ = h_out_ = nullptr; cap_ = 0; all_bytes_ = 0;
}

int StageB::init(cudaStream_t stream)
{
    stream_ = stream;
    int dev = 0; CKB(cudaGetDevice(&dev));
    CKB(cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev));
    CKB(cudaMalloc(&d_work_, 64));
    return 0;
}

int StageB::ensure(int count)
{
    if (count <= cap_) return 0;
    cudaFree(d_dags_); cudaFree(d_out_); cudaFreeHost(h_dags_); cudaFreeHost(h_out_);
    d_dags_ = d_out_ = h_dags_ = h_out_ = nullptr;
    int c = count < 64 ? 64 : count;
    CKB(cudaMalloc(&d_dags_, (size_t)c * sizeof(DevDag)));
    CKB(cudaMalloc(&d_out_, (size_t)c * sizeof(DevOrderOut)));
    CKB(cudaMallocHost(&h_dags_, (size_t)c * sizeof(DevDag)));          // pinned staging
    CKB(cudaMallocHost(&h_out_, (size_t)c * sizeof(DevOrderOut)));
    cap_ = c;
    return 0;
}

// Host-side narrowing of one ambi_dag into the device layout, with every range the kernels rely on checked:
// -2 bad argument, -3 too large (span of more than 120 segments, a loop whose 4*cn vertices exceed the path buffer).
int StageB::to_devdag(const ambi_dag &g, int reversed, DevDag &d, int index)
{
    memset(&d, 0, sizeof d);
    if (g.k < 0 || g.k > MAXK) { snprintf(err_, sizeof err_, "dag %d: k=%d out of range", index, g.k); return -2; }
    int base = 0x7fffffff, top = 0;
    for (int v = 0; v < g.k; ++v) { if (g.a[v] < base) base = g.a[v]; if (g.b[v] > top) top = g.b[v]; }
    if (g.k == 0) base = 1;
    if (g.k > 0 && top - base + 1 > 120) { snprintf(err_, sizeof err_, "dag %d: segment span too large", index); return -3; }
    d.k = g.k; d.base = base; d.reversed = reversed;
    for (int v = 0; v < g.k; ++v) {
        if (g.cn[v] < 1 || g.a[v] > g.b[v]) { snprintf(err_, sizeof err_, "dag %d: node %d has cn %d, span %d..%d", index, v, g.cn[v], g.a[v], g.b[v]); return -2; }
        if (g.kind[v] != 0 && 4 * (int)g.cn[v] > MAXPATH) { snprintf(err_, sizeof err_, "dag %d: loop multiplicity %d exceeds the path buffer", index, g.cn[v]); return -3; }
        d.kind[v] = g.kind[v]; d.a[v] = (int8_t)(g.a[v] - base + 1); d.b[v] = (int8_t)(g.b[v] - base + 1);
        d.cn[v] = g.cn[v]; d.adj[v] = g.adj[v];
    }
    return 0;
}

int StageB::run(int count, const ambi_dag *dags, const int32_t *reversed, ambi_order_result *results)
{
    if (count <= 0) return 0;
    if (int rc = ensure(count)) return rc;
    DevDag *h = reinterpret_cast<DevDag *>(h_dags_);
    for (int i = 0; i < count; ++i)
        if (int rc = to_devdag(dags[i], reversed ? reversed[i] : 0, h[i], i)) return rc;
    CKB(cudaMemcpyAsync(d_dags_, h, (size_t)count * sizeof(DevDag), cudaMemcpyHostToDevice, stream_));
    CKB(cudaMemsetAsync(d_work_, 0, 64, stream_));
    int blocks = (count + 3) / 4; if (blocks > sms_ * 8) blocks = sms_ * 8;
    if (!ev_[0]) { CKB(cudaEventCreate(&ev_[0])); CKB(cudaEventCreate(&ev_[1])); }
    CKB(cudaEventRecord(ev_[0], stream_));
    k_first_valid_order<<<blocks, 128, 0, stream_>>>((const DevDag *)d_dags_, (DevOrderOut *)d_out_, count, (uint32_t *)d_work_);
    CKB(cudaEventRecord(ev_[1], stream_));
    ++launches_;
    CKB(cudaGetLastError());
    DevOrderOut *ho = reinterpret_cast<DevOrderOut *>(h_out_);
    CKB(cudaMemcpyAsync(ho, d_out_, (size_t)count * sizeof(DevOrderOut), cudaMemcpyDeviceToHost, stream_));
    CKB(cudaStreamSynchronize(stream_));
    { float ms = 0; if (cudaEventElapsedTime(&ms, ev_[0], ev_[1]) == cudaSuccess) kernel_ms_ = ms; }
    for (int i = 0; i < count; ++i) {
        ambi_order_result &r = results[i];
        const DevOrderOut &o = ho[i];
        r.found = o.found; r.used_flipped = o.used_flipped; r.bkp_len = o.found ? o.bkp_len : 0;
        r.orders_tested = o.tested; r.status = 0; r.reserved = 0;
        const int base = h[i].base;
        for (int v = 0; v < dags[i].k; ++v) r.order[v] = o.found ? o.order[v] : 0;
        for (int v = 0; v < r.bkp_len && v < 1024; ++v) { int q = o.bkp[v]; r.bkp[v] = q > 0 ? q - 1 + base : -(-q - 1 + base); }
    }
    return 0;
}

// every valid order of one orientation, sorted lexicographically (the reference's print order)
int StageB::run_all(const ambi_dag &dag, int forward, int32_t max_orders, std::vector<std::vector<int32_t>> &out,
                    std::vector<std::vector<int32_t>> &bkp_out, unsigned long long *tested, bool *truncated,
                    const int32_t *inv_src, const int32_t *inv_tgt, int32_t max_id, int32_t path_stride,
                    std::vector<std::vector<int32_t>> *paths_out)
{
    out.clear(); bkp_out.clear(); *truncated = false;
    if (dag.k <= 0) return 0;
    if (dag.k > MAXK) return -2;
    if (int rc = ensure(1)) return rc;
    DevDag h;
    if (int rc = to_devdag(dag, 0, h, 0)) return rc;
    // grow-only scratch kept in the object (no per-call cudaMalloc, nothing to leak on the error paths)
    const bool surgery = paths_out != nullptr && inv_src && inv_tgt && path_stride > 0;
    if (paths_out) paths_out->clear();
    const size_t need = (size_t)max_orders * ((size_t)dag.k * sizeof(int32_t) + MAXPATH + sizeof(int32_t)) + 64 +
                        (surgery ? (size_t)max_orders * ((size_t)MAXPATH * 2 + (size_t)path_stride * 2 + 4) + (size_t)(max_id + 2) * 4 + 64 : 0);
    if (need > all_bytes_) {
        cudaFree(d_all_); d_all_ = nullptr; all_bytes_ = 0;
        CKB(cudaMalloc(&d_all_, need));
        all_bytes_ = need;
    }
    uint8_t *base = static_cast<uint8_t *>(d_all_);
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(base);
    int32_t *d_orders = reinterpret_cast<int32_t *>(base + 64);
    int32_t *d_lens = d_orders + (size_t)max_orders * dag.k;
    int8_t *d_bkps = reinterpret_cast<int8_t *>(d_lens + max_orders);
    CKB(cudaMemsetAsync(d_cnt, 0, 16, stream_));
    CKB(cudaMemcpyAsync(d_dags_, &h, sizeof h, cudaMemcpyHostToDevice, stream_));
    k_all_valid_orders<<<sms_, 128, 0, stream_>>>((const DevDag *)d_dags_, forward, d_orders, d_bkps, d_lens, (uint32_t)max_orders, d_cnt,
                                                   (unsigned long long *)(d_cnt + 2));
    ++launches_;
    CKB(cudaGetLastError());
    uint32_t cnt[4] = {0, 0, 0, 0};
    CKB(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, stream_));
    CKB(cudaStreamSynchronize(stream_));
    uint32_t n = cnt[0] < (uint32_t)max_orders ? cnt[0] : (uint32_t)max_orders;
    *truncated = cnt[0] > (uint32_t)max_orders;
    std::vector<int32_t> flat((size_t)n * dag.k), lens(n);
    std::vector<int8_t> paths((size_t)n * MAXPATH);
    std::vector<int16_t> hp; std::vector<int32_t> hpl;
    if (n && surgery) {
        // imperfectFBI + expansion of every valid order on the device
        uint8_t *q = reinterpret_cast<uint8_t *>(d_bkps) + (size_t)max_orders * MAXPATH;
        q = reinterpret_cast<uint8_t *>(((uintptr_t)q + 15) & ~(uintptr_t)15);
        int16_t *d_inv_s = reinterpret_cast<int16_t *>(q); int16_t *d_inv_t = d_inv_s + (max_id + 2);
        q += (size_t)(max_id + 2) * 4; q = reinterpret_cast<uint8_t *>(((uintptr_t)q + 15) & ~(uintptr_t)15);
        int16_t *d_scratch = reinterpret_cast<int16_t *>(q); q += (size_t)max_orders * MAXPATH * 2;
        int16_t *d_paths = reinterpret_cast<int16_t *>(q); q += (size_t)max_orders * path_stride * 2;
        int32_t *d_plens = reinterpret_cast<int32_t *>(q);
        std::vector<int16_t> hs((size_t)max_id + 2, 0), ht((size_t)max_id + 2, 0);
        for (int i = 0; i <= max_id; ++i) { hs[i] = (int16_t)inv_src[i]; ht[i] = (int16_t)inv_tgt[i]; }
        CKB(cudaMemcpyAsync(d_inv_s, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice, stream_));
        CKB(cudaMemcpyAsync(d_inv_t, ht.data(), ht.size() * 2, cudaMemcpyHostToDevice, stream_));
        k_path_surgery<<<(n + 127) / 128, 128, 0, stream_>>>(d_bkps, d_lens, n, h.base, d_inv_s, d_inv_t, max_id, d_scratch, d_paths, d_plens, path_stride);
        ++launches_;
        CKB(cudaGetLastError());
        hp.resize((size_t)n * path_stride); hpl.resize(n);
        CKB(cudaMemcpyAsync(hp.data(), d_paths, hp.size() * 2, cudaMemcpyDeviceToHost, stream_));
        CKB(cudaMemcpyAsync(hpl.data(), d_plens, hpl.size() * 4, cudaMemcpyDeviceToHost, stream_));
        CKB(cudaStreamSynchronize(stream_));               // hs / ht go out of scope
    }
    if (n) {
        CKB(cudaMemcpyAsync(flat.data(), d_orders, flat.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        CKB(cudaMemcpyAsync(lens.data(), d_lens, lens.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        CKB(cudaMemcpyAsync(paths.data(), d_bkps, paths.size(), cudaMemcpyDeviceToHost, stream_));
        CKB(cudaStreamSynchronize(stream_));
    }
    // the lanes append in any order: sort lexicographically (the order the reference prints)
    std::vector<uint32_t> perm(n);
    for (uint32_t i = 0; i < n; ++i) perm[i] = i;
    std::sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) {
        return std::lexicographical_compare(flat.begin() + (size_t)x * dag.k, flat.begin() + (size_t)(x + 1) * dag.k,
                                            flat.begin() + (size_t)y * dag.k, flat.begin() + (size_t)(y + 1) * dag.k); });
    out.resize(n); bkp_out.resize(n);
    if (surgery) paths_out->resize(n);
    for (uint32_t q = 0; q < n; ++q) {
        const uint32_t i = perm[q];
        if (surgery) {
            if (hpl[i] > path_stride) { snprintf(err_, sizeof err_, "expanded path of %d vertices exceeds path_stride %d", hpl[i], path_stride); return -3; }
            (*paths_out)[q].assign(hp.begin() + (size_t)i * path_stride, hp.begin() + (size_t)i * path_stride + hpl[i]);
        }
        out[q].assign(flat.begin() + (size_t)i * dag.k, flat.begin() + (size_t)(i + 1) * dag.k);
        bkp_out[q].resize(lens[i]);
        for (int t = 0; t < lens[i]; ++t) { int v = paths[(size_t)i * MAXPATH + t]; bkp_out[q][t] = v > 0 ? v - 1 + h.base : -(-v - 1 + h.base); }
    }
    if (tested) *tested += (unsigned long long)cnt[2] | ((unsigned long long)cnt[3] << 32);
    return 0;
}

}  // namespace ambi
