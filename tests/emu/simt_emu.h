// simt_emu.h -- TEST INFRASTRUCTURE ONLY: a 32-lane warp emulator for debugging warp-synchronous CUDA
// device code on a box without a GPU.  The 32 lanes are ucontext fibers scheduled round-robin on one OS
// thread; a lane only yields inside a warp collective, so everything between two collectives runs atomically
// (shared-memory "atomics" are plain read-modify-writes here).  Never part of the product build: the product
// compiles ambigram_b200/csrc/small_core.cuh with nvcc only.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <ucontext.h>
#include <algorithm>
#include <time.h>
#include <math.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__
#define __launch_bounds__(x)

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
struct int4 { int32_t x, y, z, w; };
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { int4 r = {x, y, z, w}; return r; }
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
template <class T> static inline T __ldg(const T *p) { return *p; }

namespace emu {
constexpr int W = 32;
static ucontext_t g_main, g_ctx[W];
static int g_cur = 0, g_count = 0, g_done = 0;
static unsigned g_gen = 0;
static uint64_t g_x[W];
static int g_tag[W];
static int g_line[W];
static bool g_check = false;
static long g_ncoll[W];
static int g_hist[W][8];
static bool g_finished[W];
static void (*g_entry)(int lane, void *arg);
static void *g_arg;

static inline void yield_next()
{
    int me = g_cur, nx = me;
    do { nx = (nx + 1) % W; } while (g_finished[nx] && nx != me);
    if (nx == me) return;
    g_cur = nx;
    swapcontext(&g_ctx[me], &g_ctx[nx]);
}
static inline void check_tag(int tag)
{
    g_tag[g_cur] = tag; ++g_ncoll[g_cur]; g_check = true;
    for (int i = 7; i > 0; --i) g_hist[g_cur][i] = g_hist[g_cur][i - 1];
    g_hist[g_cur][0] = g_line[g_cur];
}
static inline void verify_tags()
{
    for (int i = 1; i < W; ++i) if (!g_finished[i] && !g_finished[0] && g_tag[i] != g_tag[0]) { fprintf(stderr, "simt_emu: lanes diverged at a collective (lane %d op %d line %d, lane 0 op %d line %d)\n", i, g_tag[i], g_line[i], g_tag[0], g_line[0]); fprintf(stderr, "counts: lane %d %ld, lane 0 %ld\n hist lane %d:", i, g_ncoll[i], g_ncoll[0], i); for (int k = 0; k < 8; ++k) fprintf(stderr, " %d", g_hist[i][k]); fprintf(stderr, "\n hist lane 0:"); for (int k = 0; k < 8; ++k) fprintf(stderr, " %d", g_hist[0][k]); fprintf(stderr, "\n"); abort(); }
}
static inline void barrier()
{
    unsigned gen = g_gen;
    if (++g_count == W - g_done) { if (g_check) verify_tags(); g_check = false; g_count = 0; ++g_gen; return; }
    while (g_gen == gen) yield_next();
}
static void tramp(int lane)
{
    g_entry(lane, g_arg);
    g_finished[lane] = true; ++g_done;
    if (g_count == W - g_done && g_count > 0) { g_count = 0; ++g_gen; }     // others wait in a barrier we never reach
    for (int i = 1; i <= W; ++i) { int nx = (lane + i) % W; if (!g_finished[nx]) { g_cur = nx; setcontext(&g_ctx[nx]); } }
    setcontext(&g_main);
}
static inline void run_warp(void (*entry)(int, void *), void *arg)
{
    static char *stacks = nullptr;
    const size_t SS = 1 << 18;
    if (!stacks) stacks = (char *)malloc(SS * W);
    g_entry = entry; g_arg = arg; g_count = 0; g_done = 0; g_gen = 0;
    for (int i = 0; i < W; ++i) {
        g_finished[i] = false;
        getcontext(&g_ctx[i]);
        g_ctx[i].uc_stack.ss_sp = stacks + SS * i; g_ctx[i].uc_stack.ss_size = SS; g_ctx[i].uc_link = &g_main;
        makecontext(&g_ctx[i], (void (*)())tramp, 1, i);
    }
    g_cur = 0;
    swapcontext(&g_main, &g_ctx[0]);
}
static inline int lane() { return g_cur; }
}  // namespace emu

static inline uint32_t emu__ballot_sync(uint32_t, bool p)
{
    emu::g_x[emu::lane()] = p; emu::check_tag(1); emu::barrier();
    uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= (uint32_t)(emu::g_x[i] & 1) << i;
    emu::barrier(); return r;
}
template <class T> static inline T emu__shfl_sync(uint32_t, T v, int src)
{
    uint64_t b = 0; memcpy(&b, &v, sizeof v); emu::g_x[emu::lane()] = b; emu::check_tag(2); emu::barrier();
    T r; memcpy(&r, &emu::g_x[src & 31], sizeof r); emu::barrier(); return r;
}
template <class T> static inline T emu__shfl_up_sync(uint32_t, T v, int d)
{
    uint64_t b = 0; memcpy(&b, &v, sizeof v); emu::g_x[emu::lane()] = b; emu::check_tag(3); emu::barrier();
    int s = emu::lane() - d; T r = v; if (s >= 0) memcpy(&r, &emu::g_x[s], sizeof r); emu::barrier(); return r;
}
template <class T> static inline T emu__shfl_xor_sync(uint32_t, T v, int m)
{
    uint64_t b = 0; memcpy(&b, &v, sizeof v); emu::g_x[emu::lane()] = b; emu::check_tag(4); emu::barrier();
    T r; memcpy(&r, &emu::g_x[(emu::lane() ^ m) & 31], sizeof r); emu::barrier(); return r;
}
static inline void emu__syncwarp(uint32_t = 0xffffffffu) { emu::check_tag(99); emu::barrier(); }
static inline uint32_t emu__match_any_sync(uint32_t, unsigned long long v)
{
    emu::g_x[emu::lane()] = v; emu::check_tag(5); emu::barrier();
    uint32_t r = 0; for (int i = 0; i < 32; ++i) if (emu::g_x[i] == v) r |= 1u << i;
    emu::barrier(); return r;
}
static inline uint32_t emu__reduce_or_sync(uint32_t, uint32_t v)
{
    emu::g_x[emu::lane()] = v; emu::check_tag(6); emu::barrier();
    uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= (uint32_t)emu::g_x[i];
    emu::barrier(); return r;
}
static inline uint32_t emu__reduce_add_sync(uint32_t, uint32_t v)
{
    emu::g_x[emu::lane()] = v; emu::check_tag(7); emu::barrier();
    uint32_t r = 0; for (int i = 0; i < 32; ++i) r += (uint32_t)emu::g_x[i];
    emu::barrier(); return r;
}
static inline int emu__reduce_min_sync(uint32_t, int v)
{
    emu::g_x[emu::lane()] = (uint64_t)(int64_t)v; emu::check_tag(8); emu::barrier();
    int r = 0x7fffffff; for (int i = 0; i < 32; ++i) r = std::min(r, (int)(int64_t)emu::g_x[i]);
    emu::barrier(); return r;
}
static inline int emu__reduce_max_sync(uint32_t, int v)
{
    emu::g_x[emu::lane()] = (uint64_t)(int64_t)v; emu::check_tag(9); emu::barrier();
    int r = -0x7fffffff - 1; for (int i = 0; i < 32; ++i) r = std::max(r, (int)(int64_t)emu::g_x[i]);
    emu::barrier(); return r;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline uint32_t __float2uint_rz(float f) { return (uint32_t)f; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float exp2f_(float x) { return exp2f(x); }
template <class T> static inline T atomicCAS(T *p, T cmp, T val) { T o = *p; if (o == cmp) *p = val; return o; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
using std::max;
using std::min;
static inline unsigned long long emu_globaltimer()
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}

// grid coordinates of the emulated warp: the harness sets them before it runs a warp (one block = its warps in turn)
namespace emu {
struct Dim3 { unsigned x, y, z; };
static unsigned g_block_idx = 0, g_grid_dim = 1, g_block_dim = 32, g_warp_in_block = 0;
static inline Dim3 thread_idx() { Dim3 d = {g_warp_in_block * 32u + (unsigned)lane(), 0, 0}; return d; }
static inline Dim3 block_idx() { Dim3 d = {g_block_idx, 0, 0}; return d; }
static inline Dim3 block_dim() { Dim3 d = {g_block_dim, 1, 1}; return d; }
static inline Dim3 grid_dim() { Dim3 d = {g_grid_dim, 1, 1}; return d; }
}
#define threadIdx (emu::thread_idx())
#define blockIdx (emu::block_idx())
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())

#define __ballot_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__ballot_sync(__VA_ARGS__))
#define __shfl_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__shfl_sync(__VA_ARGS__))
#define __shfl_up_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__shfl_up_sync(__VA_ARGS__))
#define __shfl_xor_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__shfl_xor_sync(__VA_ARGS__))
#define __syncwarp(...) (emu::g_line[emu::lane()] = __LINE__, emu__syncwarp(__VA_ARGS__))
#define __match_any_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__match_any_sync(__VA_ARGS__))
#define __reduce_or_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__reduce_or_sync(__VA_ARGS__))
#define __reduce_add_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__reduce_add_sync(__VA_ARGS__))
#define __reduce_min_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__reduce_min_sync(__VA_ARGS__))
#define __reduce_max_sync(...) (emu::g_line[emu::lane()] = __LINE__, emu__reduce_max_sync(__VA_ARGS__))
