// capi.cu -- C ABI of libambigram_b200.so (see include/ambigram_b200.h).
//
// Host side of the boundary: flattens the caller's problems into DevProblem
// records, drives the Stage-A / Stage-B device searches, and maps results back
// to the reference's own numbering (element_cn[x] in variableIdx order,
// localhap.cpp:122-133, 204-211).  No search arithmetic happens on the host.
#include "../../include/ambigram_b200.h"
#include "ambi_common.cuh"
#include "stage_a.h"
#include "stage_b.h"
#include "host_flatten.h"
#include "path_surgery.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#include <dlfcn.h>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>
#include <chrono>

using namespace ambi;

// ---- minimal NCCL surface, resolved at run time (libnccl.so.2) -------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllGather && AllReduce;
    }
};
static NcclApi g_nccl;
enum { kNcclUint32 = 3, kNcclInt32 = 2, kNcclMin = 3 };   // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct ambi_ctx {
    ambi_config cfg;
    cudaStream_t stream = nullptr;
    StageA sa;
    StageB sb;
    ncclComm_t comm = nullptr;
    uint32_t *d_gather = nullptr;      // world * 40 words
    std::string err;
    ambi_batch *resident = nullptr;
    uint64_t extra_launches = 0;
};

struct ambi_batch {
    int count = 0;
    std::vector<int32_t> n, bias;
    std::vector<int8_t> ush;
    bool ran = false;
};

// Host-side flattening / result mapping of a large batch is per-problem independent work that would otherwise sit
// between the H2D copy and the launch: it runs on a small pool of worker threads that lives as long as the process
// (starting threads per call cost as much as the work itself at 32768 problems).  Chunks are handed out through an
// atomic cursor; the calling thread works too.  A second caller arriving while the pool is busy does its work inline.
namespace {
struct HostPool {
    std::mutex m, run_m;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int, int)> *fn = nullptr;
    int count = 0, chunk = 1, pending = 0;
    unsigned long long epoch = 0;
    std::atomic<int> next{0};

    explicit HostPool(int nworkers)
    {
        for (int i = 0; i < nworkers; ++i) { workers.emplace_back([this] { work(); }); workers.back().detach(); }
    }
    void drain(const std::function<void(int, int)> &f, int cnt, int ch)
    {
        for (;;) {
            const int lo = next.fetch_add(ch);
            if (lo >= cnt) break;
            f(lo, std::min(cnt, lo + ch));
        }
    }
    void work()
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv_work.wait(lk, [&] { return epoch != seen; });
            seen = epoch;
            const std::function<void(int, int)> *f = fn;
            const int cnt = count, ch = chunk;
            lk.unlock();
            drain(*f, cnt, ch);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    void run(int cnt, int ch, const std::function<void(int, int)> &f)
    {
        std::unique_lock<std::mutex> only(run_m, std::try_to_lock);
        if (!only.owns_lock()) { f(0, cnt); return; }
        {
            std::lock_guard<std::mutex> lk(m);
            fn = &f; count = cnt; chunk = ch; next.store(0); pending = (int)workers.size(); ++epoch;
        }
        cv_work.notify_all();
        drain(f, cnt, ch);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });     // every worker has left drain(): `f` may go out of scope
    }
};
}  // namespace

template <class F> static void parallel_for(int count, F fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
    if (const char *e = getenv("AMBI_HOST_THREADS")) nt = std::max(1, std::min(64, atoi(e)));
    if (nt <= 1 || count < 4096) { fn(0, count); return; }
    static HostPool *pool = new HostPool(nt - 1);           // never destroyed: its threads sleep until the process ends
    const std::function<void(int, int)> f = fn;
    pool->run(count, 512, f);
}

static thread_local std::string g_err;

static int fail(ambi_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    g_err = msg;
    return code;
}

extern "C" int ambi_abi_version(void) { return AMBI_ABI_VERSION; }

extern "C" const char *ambi_status_string(int s)
{
    switch (s) {
    case AMBI_OK: return "ok";
    case AMBI_INFEASIBLE: return "infeasible";
    case AMBI_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case AMBI_ERR_BAD_ARG: return "bad argument";
    case AMBI_ERR_TOO_LARGE: return "partition has more than AMBI_MAX_SEGMENTS segments";
    case AMBI_ERR_CUDA: return "CUDA error";
    case AMBI_ERR_FRONTIER_OVERFLOW: return "candidate buffers exhausted (raise node_capacity)";
    case AMBI_ERR_NOT_HALF_INTEGRAL: return "copy numbers are not half-integral";
    case AMBI_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
    }
}

extern "C" const char *ambi_last_error(const ambi_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int ambi_init(ambi_ctx **out, const ambi_config *cfg)
{
    if (!out) return AMBI_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, AMBI_ERR_NO_DEVICE, std::string("ambi_init: no CUDA device: ") + cudaGetErrorString(e));
    }
    ambi_ctx *ctx = new ambi_ctx();
    memset(&ctx->cfg, 0, sizeof ctx->cfg);
    if (cfg) ctx->cfg = *cfg;
    if (ctx->cfg.world <= 0) ctx->cfg.world = 1;
    if (ctx->cfg.device < 0 || ctx->cfg.device >= ndev) { delete ctx; return fail(nullptr, AMBI_ERR_BAD_ARG, "ambi_init: bad device ordinal"); }
    if ((e = cudaSetDevice(ctx->cfg.device)) != cudaSuccess) { delete ctx; return fail(nullptr, AMBI_ERR_CUDA, cudaGetErrorString(e)); }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return fail(nullptr, AMBI_ERR_CUDA, cudaGetErrorString(e)); }
    if (ctx->sa.init(ctx->stream, ctx->cfg.node_capacity, ctx->cfg.beam_width, ctx->cfg.rank, ctx->cfg.world, ctx->cfg.shard_level, (ctx->cfg.flags & AMBI_CFG_SMALL_KERNEL) != 0)) {
        std::string m = ctx->sa.error(); ctx->sa.release(); cudaStreamDestroy(ctx->stream); delete ctx; return fail(nullptr, AMBI_ERR_CUDA, m);
    }
    if (ctx->sb.init(ctx->stream)) {
        std::string m = ctx->sb.error(); ctx->sa.release(); ctx->sb.release(); cudaStreamDestroy(ctx->stream); delete ctx; return fail(nullptr, AMBI_ERR_CUDA, m);
    }
    *out = ctx;
    return AMBI_OK;
}

extern "C" void ambi_destroy(ambi_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    cudaFree(ctx->d_gather);
    ctx->sa.release(); ctx->sb.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void *ambi_ctx_stream(ambi_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int ambi_ctx_sync(ambi_ctx *ctx)
{
    if (!ctx) return AMBI_ERR_BAD_ARG;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    return e == cudaSuccess ? AMBI_OK : fail(ctx, AMBI_ERR_CUDA, cudaGetErrorString(e));
}
extern "C" int ambi_ctx_set_profile(ambi_ctx *ctx, int enable)
{
    if (!ctx) return AMBI_ERR_BAD_ARG;
    ctx->sa.set_profile(enable != 0);
    return AMBI_OK;
}
extern "C" int ambi_ctx_get_profile(ambi_ctx *ctx, ambi_profile *out)
{
    if (!ctx || !out) return AMBI_ERR_BAD_ARG;
    unsigned long long l = 0;
    ctx->sa.profile(&out->expand_ms, &l, &out->total_ms, &out->incumbent_pass_ms);
    out->expand_launches = l;
    out->record_stride = (uint64_t)ctx->sa.record_stride();
    const unsigned long long *f = ctx->sa.funnel();
    out->configs_enumerated = f[0]; out->bounds_evaluated = f[1]; out->children_written = f[2]; out->bound_batches = f[3];
    out->exchange_rounds = ctx->sa.exchange_rounds(); out->chunked_runs = ctx->sa.chunked_runs();
    out->small_ms = ctx->sa.small_ms(); out->small_solved = (uint64_t)ctx->sa.small_solved(); out->global_solved = (uint64_t)ctx->sa.global_solved();
    out->small_capacity = (uint64_t)ctx->sa.small_capacity(); out->stage_b_kernel_ms = ctx->sb.kernel_ms();
    return AMBI_OK;
}
namespace ambi { int measure_peaks(cudaStream_t st, ambi_peaks *out); }
extern "C" int ambi_measure_peaks(ambi_ctx *ctx, ambi_peaks *out)
{
    if (!ctx || !out) return AMBI_ERR_BAD_ARG;
    cudaSetDevice(ctx->cfg.device);
    memset(out, 0, sizeof *out);
    if (ambi::measure_peaks(ctx->stream, out)) return fail(ctx, AMBI_ERR_CUDA, "ambi_measure_peaks: kernel failed");
    ctx->extra_launches += 18;
    return AMBI_OK;
}
extern "C" uint64_t ambi_ctx_launch_count(const ambi_ctx *ctx) { return ctx ? ctx->sa.launches() + ctx->sb.launches() + ctx->extra_launches : 0; }

extern "C" int ambi_nccl_unique_id(void *id128)
{
    if (!id128) return AMBI_ERR_BAD_ARG;
    if (!g_nccl.load()) return fail(nullptr, AMBI_ERR_NCCL, "cannot load libnccl.so.2");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, AMBI_ERR_NCCL, "ncclGetUniqueId failed");
    memcpy(id128, &id, 128);
    return AMBI_OK;
}

extern "C" int ambi_comm_init(ambi_ctx *ctx, const void *id128)
{
    if (!ctx || !id128) return AMBI_ERR_BAD_ARG;
    if (!g_nccl.load()) return fail(ctx, AMBI_ERR_NCCL, "cannot load libnccl.so.2");
    cudaSetDevice(ctx->cfg.device);
    if (ctx->comm) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_comm_init: this context already has a communicator");
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, ctx->cfg.world, id, ctx->cfg.rank);
    if (r != 0) return fail(ctx, AMBI_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    if (cudaMalloc(&ctx->d_gather, (size_t)(ctx->cfg.world + 1) * 40 * 4) != cudaSuccess) return fail(ctx, AMBI_ERR_CUDA, "cudaMalloc gather");
    // periodic exchange of pruning limits between the ranks of a prefix-sharded depth-first search
    ctx->sa.set_allreduce([](void *user, int32_t *d_buf, int count, int op) -> int {
        ambi_ctx *c = static_cast<ambi_ctx *>(user);
        static const int nccl_op[3] = {3 /* ncclMin */, 2 /* ncclMax */, 0 /* ncclSum */};
        return g_nccl.AllReduce(d_buf, d_buf, (size_t)count, kNcclInt32, nccl_op[op], c->comm, c->stream) != 0;
    }, ctx);
    return AMBI_OK;
}

extern "C" int ambi_bfb_better(int32_t n, int64_t sa, const int32_t *xa, int64_t sb, const int32_t *xb)
{
    if (sa != sb) return sa < sb;
    const int H = n * (n + 1) / 2;
    for (int a = 0; a < n; ++a) {
        for (int b = a; b < n; ++b) { int i = pair_index(n, a, b); if (xa[i] != xb[i]) return xa[i] < xb[i]; }
        for (int b = a; b < n; ++b) { int i = H + pair_index(n, a, b); if (xa[i] != xb[i]) return xa[i] < xb[i]; }
    }
    return 0;
}

// ---- device-resident batch -------------------------------------------------
extern "C" int ambi_batch_upload(ambi_ctx *ctx, int32_t count, const ambi_bfb_problem *problems, ambi_batch **out)
{
    if (!ctx || !out || count <= 0 || !problems) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_batch_upload: bad argument");
    cudaSetDevice(ctx->cfg.device);
    static const bool timing = getenv("AMBI_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    // what travels over PCIe is a compact form (StageA::compact_words): header + the segments' 2c|2J words, sized by
    // the longest partition of the batch; the device expands it (k_unpack).  Pieces are copied by the threads that
    // flattened them, while the others are still working.
    int nmax = 1;
    bool any_tgs = false;
    for (int i = 0; i < count; ++i) {
        nmax = std::max(nmax, (int)problems[i].n);
        any_tgs = any_tgs || (problems[i].n_tgs > 0 && (problems[i].flags & AMBI_FLAG_JUNC_INFO) && problems[i].tgs_pairs);
    }
    nmax = std::min(nmax, (int)AMBI_MAXN);
    const int words = StageA::compact_words(nmax, any_tgs);
    uint32_t *hc = ctx->sa.compact_staging(count);
    if (!hc) return fail(ctx, AMBI_ERR_CUDA, ctx->sa.error());
    uint32_t *dc = ctx->sa.device_compact();
    const int device = ctx->cfg.device;
    cudaStream_t stream = ctx->stream;
    std::atomic<int> copy_err{0};
    ambi_batch *b = new ambi_batch();
    b->count = count; b->n.resize(count); b->bias.resize(count); b->ush.resize(count);
    std::atomic<int> bad_rc{0}, bad_i{-1};
    std::string bad_why;
    parallel_for(count, [&](int lo, int hi) {
        DevProblem d;
        for (int i = lo; i < hi; ++i) {
            std::string why;
            int rc = to_dev(problems[i], d, why);
            if (rc) { int expect = 0; if (bad_rc.compare_exchange_strong(expect, rc)) { bad_i = i; bad_why = why; } return; }
            b->n[i] = problems[i].n; b->bias[i] = problems[i].bias; b->ush[i] = (int8_t)problems[i].unit_shift;
            uint32_t *c = hc + (size_t)i * words;
            c[0] = (uint32_t)d.n | ((uint32_t)d.lmax << 8) | ((uint32_t)d.ush << 16) | (d.has_tgs ? 1u << 24 : 0u);
            c[1] = (uint32_t)d.loop_ub; c[2] = (uint32_t)d.zero_score; c[3] = (uint32_t)d.root_lb;
            for (int j = 0; j < nmax; ++j) c[4 + j] = d.cj[j];
            if (any_tgs) for (int j = 0; j < nmax; ++j) c[4 + nmax + j] = d.tgs_mask[j];
        }
        if (cudaSetDevice(device) != cudaSuccess ||
            cudaMemcpyAsync(dc + (size_t)lo * words, hc + (size_t)lo * words, (size_t)(hi - lo) * words * 4, cudaMemcpyHostToDevice, stream) != cudaSuccess) copy_err = 1;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (bad_rc) { delete b; return fail(ctx, bad_rc, "problem " + std::to_string(bad_i.load()) + ": " + bad_why); }
    if (copy_err) { delete b; (void)cudaGetLastError(); return fail(ctx, AMBI_ERR_CUDA, "ambi_batch_upload: host to device copy failed"); }
    if (ctx->sa.unpack_compact(count, nmax, any_tgs)) { delete b; return fail(ctx, AMBI_ERR_CUDA, ctx->sa.error()); }
    if (int urc = ctx->sa.upload(nullptr, count, true, b->n.data(), b->ush.data())) { delete b; return fail(ctx, urc == -5 ? AMBI_ERR_FRONTIER_OVERFLOW : AMBI_ERR_CUDA, ctx->sa.error()); }
    cudaStreamSynchronize(ctx->stream);
    if (timing) {
        auto t2 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
        fprintf(stderr, "[ambi] upload %d problems: flatten (+ copies enqueued) %.3f ms, bookkeeping + copies done %.3f ms\n", count, ms(t0, t1), ms(t1, t2));
    }
    ctx->resident = b;
    *out = b;
    return AMBI_OK;
}

extern "C" int ambi_batch_run(ambi_ctx *ctx, ambi_batch *b)
{
    if (!ctx || !b || ctx->resident != b) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_batch_run: batch is not resident");
    cudaSetDevice(ctx->cfg.device);
    if (ctx->sa.run(false)) return fail(ctx, AMBI_ERR_CUDA, ctx->sa.error());
    b->ran = true;
    return AMBI_OK;
}

static int fetch_into(ambi_ctx *ctx, ambi_batch *b, ambi_bfb_solution *sol, bool shard)
{
    StageA::Fetched F;
    int rc = ctx->sa.fetch_raw(&F);
    if (rc) return fail(ctx, AMBI_ERR_CUDA, ctx->sa.error());
    std::atomic<int> worst{AMBI_OK};
    parallel_for(b->count, [&](int lo, int hi) {
        int w = AMBI_OK;
        for (int i = lo; i < hi; ++i) {
            const int n = b->n[i];
            ambi_bfb_solution &s = sol[i];
            const int32_t sc = F.score[i];
            s.status = F.overflow ? AMBI_ERR_FRONTIER_OVERFLOW : (sc == 0x7fffffff ? AMBI_ERR_CUDA : AMBI_OK);
            if (shard && s.status == AMBI_ERR_CUDA) s.status = AMBI_OK;      // this shard saw no completed assignment
            s.score2 = sc;
            s.objective2 = sc - ((2ll * b->bias[i]) << b->ush[i]);
            s.candidates_scored = F.cand[i]; s.nodes_expanded = F.node[i];
            s.incumbent_score2 = F.incumbent[i];
            s.seconds = F.t_proof[i]; s.seconds_to_best = F.t_best[i];
            if (s.element_cn) {
                if (sc != 0x7fffffff) rows_to_element_cn(n, F.rows + (size_t)i * F.rows_stride, s.element_cn);
                else memset(s.element_cn, 0, sizeof(int32_t) * n * (n + 1));
            }
            if (s.status < w) w = s.status;
        }
        int cur = worst.load();
        while (w < cur && !worst.compare_exchange_weak(cur, w)) {}
    });
    if (worst == AMBI_ERR_FRONTIER_OVERFLOW) return fail(ctx, worst, ctx->sa.error());
    if (worst) return fail(ctx, worst, "search returned no assignment");
    return AMBI_OK;
}

extern "C" int ambi_batch_fetch(ambi_ctx *ctx, ambi_batch *b, ambi_bfb_solution *solutions)
{
    if (!ctx || !b || !solutions || ctx->resident != b || !b->ran) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_batch_fetch: nothing to fetch");
    cudaSetDevice(ctx->cfg.device);
    return fetch_into(ctx, b, solutions, false);
}

extern "C" void ambi_batch_free(ambi_ctx *ctx, ambi_batch *b)
{
    if (ctx && ctx->resident == b) ctx->resident = nullptr;
    delete b;
}

// ---- Seam A ---------------------------------------------------------------
// rough number of frontier records a partition of n segments needs at its widest level; only used to
// cut a large mixed batch into sub-batches that fit the candidate buffers (an underestimate is
// harmless: the level that does not fit is redone by the chunked depth-first pass)
static double frontier_estimate(int n) { double e = 64.0; for (int i = 12; i < n; ++i) e *= 1.6; return e; }

extern "C" int ambi_bfb_solve_batch(ambi_ctx *ctx, int32_t count, const ambi_bfb_problem *problems,
                                    ambi_bfb_solution *solutions)
{
    if (!ctx || !solutions || !problems || count <= 0) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_solve_batch: bad argument");
    const double budget = 0.4 * (double)ctx->sa.capacity_base();
    int worst = AMBI_OK;
    int32_t lo = 0;
    while (lo < count) {
        int32_t hi = lo; double w = 0;
        while (hi < count && (hi == lo || w + frontier_estimate(problems[hi].n) <= budget)) { w += frontier_estimate(problems[hi].n); ++hi; }
        ambi_batch *b = nullptr;
        static const bool timing = getenv("AMBI_TIMING") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        int rc = ambi_batch_upload(ctx, hi - lo, problems + lo, &b);
        if (rc) return rc;
        auto t1 = std::chrono::steady_clock::now();
        rc = ambi_batch_run(ctx, b);
        if (timing) cudaStreamSynchronize(ctx->stream);
        auto t2 = std::chrono::steady_clock::now();
        if (!rc) rc = ambi_batch_fetch(ctx, b, solutions + lo);
        auto t3 = std::chrono::steady_clock::now();
        ambi_batch_free(ctx, b);
        if (timing) {
            auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
            fprintf(stderr, "[ambi] solve_batch %d problems: flatten+H2D %.3f ms, search %.3f ms, D2H+map %.3f ms\n", hi - lo, ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
        if (rc && rc != AMBI_ERR_FRONTIER_OVERFLOW) return rc;
        if (rc) worst = rc;
        lo = hi;
    }
    return worst;
}

extern "C" int ambi_bfb_solve_shard(ambi_ctx *ctx, const ambi_bfb_problem *problem, ambi_bfb_solution *solution)
{
    if (!ctx || !problem || !solution) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_solve_shard: bad argument");
    ambi_batch *b = nullptr;
    int rc = ambi_batch_upload(ctx, 1, problem, &b);
    if (rc) return rc;
    if (ctx->sa.run(ctx->cfg.world > 1)) { ambi_batch_free(ctx, b); return fail(ctx, AMBI_ERR_CUDA, ctx->sa.error()); }
    b->ran = true;
    rc = fetch_into(ctx, b, solution, ctx->cfg.world > 1);
    ambi_batch_free(ctx, b);
    return rc;
}

extern "C" int ambi_bfb_solve(ambi_ctx *ctx, const ambi_bfb_problem *problem, ambi_bfb_solution *solution)
{
    if (!ctx || !problem || !solution) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_solve: bad argument");
    if (ctx->cfg.world <= 1 || !ctx->comm) return ambi_bfb_solve_batch(ctx, 1, problem, solution);
    // prefix-sharded: every rank explores the subtrees hashed to it, then one all-gather of the
    // (score, row codes) records over NVLink picks the global arg-min in (score, search order)
    const int n = std::max(1, std::min((int)problem->n, (int)AMBI_MAX_SEGMENTS));   // an out-of-range n fails in the shard call below
    std::vector<int32_t> x((size_t)n * (n + 1));
    ambi_bfb_solution local = *solution;
    local.element_cn = x.data();
    // The call is collective: a rank whose shard failed (frontier overflow, bad argument, ...) still takes part in
    // the all-gather and carries its status in word 38 of its record, so that no peer is left waiting in the
    // collective and every rank returns the same error.
    const int rc_local = ambi_bfb_solve_shard(ctx, problem, &local);
    const std::string err_local = rc_local ? ctx->err : std::string();
    const int W = ctx->cfg.world;
    // pack: [score, n rows re-encoded from element_cn in search order]
    std::vector<uint32_t> rec(40, 0), all((size_t)W * 40);
    rec[0] = rc_local ? 0x7fffffffu : (uint32_t)local.score2;
    rec[38] = (uint32_t)rc_local;
    // per-row keys comparable as integers: recompute the row codes from x
    const int H = n * (n + 1) / 2;
    for (int a = 0; a < n && n > 1 && !rc_local; ++a) {
        int p[2] = {-1, -1}, l[2] = {-1, -1}, np = 0, nl = 0;
        for (int b2 = a; b2 < n; ++b2) {
            if (x[pair_index(n, a, b2)] && np < 2) p[np++] = b2;
            for (int c = 0; c < x[H + pair_index(n, a, b2)] && nl < 2; ++c) l[nl++] = b2;
        }
        int ip = np == 0 ? 0 : (np == 1 ? ambi_ip_code(n - 1 - p[0], -1) : ambi_ip_code(n - 1 - p[0], n - 1 - p[1]));
        int il = nl == 0 ? 0 : (nl == 1 ? ambi_il_code(n - 1 - l[0], -1) : ambi_il_code(n - 1 - l[0], n - 1 - l[1]));
        rec[1 + a] = ((uint32_t)ip << 10) | (uint32_t)il;
    }
    if (n == 1 && !rc_local) { rec[1] = 0x80000000u | ((uint32_t)x[0] << 20) | (uint32_t)x[1]; }
    rec[34] = (uint32_t)(local.candidates_scored & 0xffffffffu); rec[35] = (uint32_t)(local.candidates_scored >> 32);
    rec[36] = (uint32_t)(local.nodes_expanded & 0xffffffffu); rec[37] = (uint32_t)(local.nodes_expanded >> 32);
    cudaSetDevice(ctx->cfg.device);
    uint32_t *d_send = ctx->d_gather, *d_recv = ctx->d_gather + 40;
    if (cudaMemcpyAsync(d_send, rec.data(), 160, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return fail(ctx, AMBI_ERR_CUDA, "all-gather: H2D failed");
    ncclResult_t r = g_nccl.AllGather(d_send, d_recv, 40, kNcclUint32, ctx->comm, ctx->stream);
    if (r != 0) return fail(ctx, AMBI_ERR_NCCL, "ncclAllGather failed");
    if (cudaMemcpyAsync(all.data(), d_recv, (size_t)W * 160, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return fail(ctx, AMBI_ERR_CUDA, "all-gather: D2H failed");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(ctx, AMBI_ERR_CUDA, "all-gather sync failed");
    if (rc_local) return fail(ctx, rc_local, err_local);
    for (int w = 0; w < W; ++w) {
        const int rc_w = (int)(int32_t)all[(size_t)w * 40 + 38];
        if (rc_w) return fail(ctx, rc_w, "prefix-sharded solve: rank " + std::to_string(w) + " failed: " + ambi_status_string(rc_w));
    }
    int best = 0; uint64_t cands = 0, nodes = 0;
    for (int w = 0; w < W; ++w) {
        const uint32_t *q = &all[(size_t)w * 40], *bq = &all[(size_t)best * 40];
        cands += (uint64_t)q[34] | ((uint64_t)q[35] << 32); nodes += (uint64_t)q[36] | ((uint64_t)q[37] << 32);
        bool better = false;
        if ((int32_t)q[0] != (int32_t)bq[0]) better = (int32_t)q[0] < (int32_t)bq[0];
        else for (int a = 0; a < n; ++a) if (q[1 + a] != bq[1 + a]) { better = q[1 + a] < bq[1 + a]; break; }
        if (better) best = w;
    }
    const uint32_t *bq = &all[(size_t)best * 40];
    solution->score2 = (int32_t)bq[0];
    solution->objective2 = solution->score2 - ((2ll * problem->bias) << problem->unit_shift);
    solution->candidates_scored = cands; solution->nodes_expanded = nodes;
    solution->incumbent_score2 = local.incumbent_score2;
    solution->seconds = local.seconds; solution->seconds_to_best = local.seconds_to_best;
    solution->status = solution->score2 == 0x7fffffff ? AMBI_ERR_CUDA : AMBI_OK;
    if (solution->element_cn) rows_to_element_cn(n, bq + 1, solution->element_cn);
    return solution->status;
}

// ---- Seam B ---------------------------------------------------------------
// constructDAG (LGM:3276-3378) on the host side of the boundary.
extern "C" int ambi_bfb_construct_dag(int32_t n, int32_t start_id, const int32_t *x, ambi_dag *dag)
{
    if (!x || !dag || n < 1 || n > AMBI_MAX_SEGMENTS) return AMBI_ERR_BAD_ARG;
    memset(dag, 0, sizeof *dag);
    const int H = n * (n + 1) / 2;
    struct Node { int kind, a, b, cn; std::string key; };
    std::vector<Node> loops, pats;
    for (int a = 0; a < n; ++a)
        for (int b = a; b < n; ++b) {
            int i = pair_index(n, a, b);
            if (x[i] > 0) pats.push_back({0, a + start_id, b + start_id, x[i], std::to_string(a + start_id) + "," + std::to_string(b + start_id)});
            if (x[H + i] > 0) loops.push_back({1, a + start_id, b + start_id, x[H + i], std::to_string(a + start_id) + "," + std::to_string(b + start_id)});
        }
    // std::map<string,int> iteration order: "l:..." keys before "p:...", each in string order (LGM:3279)
    auto bykey = [](const Node &u, const Node &v) { return u.key < v.key; };
    std::sort(loops.begin(), loops.end(), bykey);
    std::sort(pats.begin(), pats.end(), bykey);
    // loops by decreasing length, stable (LGM:3303 as intended; see DESIGN.md on k > 16)
    std::stable_sort(loops.begin(), loops.end(), [](const Node &u, const Node &v) { return (u.b - u.a) > (v.b - v.a); });
    const int k = (int)(loops.size() + pats.size());
    if (k > AMBI_MAX_COMPONENTS) return AMBI_ERR_TOO_LARGE;
    std::vector<Node> nodes(loops); nodes.insert(nodes.end(), pats.begin(), pats.end());
    dag->k = k;
    for (int i = 0; i < k; ++i) { dag->kind[i] = (int8_t)nodes[i].kind; dag->a[i] = (int16_t)nodes[i].a; dag->b[i] = (int16_t)nodes[i].b; dag->cn[i] = (int16_t)nodes[i].cn; }
    std::vector<std::vector<int>> adj(k), parents(k);
    auto share = [&](int i, int j) { return nodes[i].a == nodes[j].a || nodes[i].b == nodes[j].b; };
    auto len = [&](int i) { return nodes[i].b - nodes[i].a; };
    for (int i = 0; i < k; ++i) if (nodes[i].kind == 0) {
        for (int j = 0; j < k; ++j) if (nodes[j].kind == 0 && share(i, j) && len(i) > len(j)) { adj[i].push_back(j); parents[j].push_back(i); }   // p -> p
        for (int j = 0; j < k; ++j) if (nodes[j].kind == 1 && share(i, j) && len(i) > len(j)) { adj[i].push_back(j); parents[j].push_back(i); }   // p -> l
    }
    for (int i = 0; i < k; ++i) if (nodes[i].kind == 1) {
        for (int j = 0; j < k; ++j) {                                                                  // l -> p
            if (std::find(parents[i].begin(), parents[i].end(), j) != parents[i].end()) continue;
            if (nodes[j].kind == 0 && share(i, j)) {
                if (len(i) > len(j)) { adj[i].push_back(j); parents[j].push_back(i); }
                else for (int par : parents[i])
                    if (std::find(adj[par].begin(), adj[par].end(), j) != adj[par].end()) { adj[i].push_back(j); parents[j].push_back(i); break; }
            }
        }
        for (int j = 0; j < k; ++j) if (nodes[j].kind == 1 && share(i, j) && len(i) > len(j)) { adj[i].push_back(j); parents[j].push_back(i); }   // l -> l
    }
    for (int i = 0; i < k; ++i) for (int j : adj[i]) dag->adj[i] |= 1ull << j;
    return AMBI_OK;
}

extern "C" int ambi_bfb_first_valid_order_batch(ambi_ctx *ctx, int32_t count, const ambi_dag *dags,
                                                const int32_t *reversed, ambi_order_result *results)
{
    if (!ctx || !dags || !results || count <= 0) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_first_valid_order_batch: bad argument");
    cudaSetDevice(ctx->cfg.device);
    int rc = ctx->sb.run(count, dags, reversed, results);
    if (rc == -2) return fail(ctx, AMBI_ERR_BAD_ARG, ctx->sb.error());
    if (rc == -3) return fail(ctx, AMBI_ERR_TOO_LARGE, ctx->sb.error());
    if (rc) return fail(ctx, AMBI_ERR_CUDA, ctx->sb.error());
    return AMBI_OK;
}

extern "C" int ambi_bfb_first_valid_order(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, ambi_order_result *result)
{
    return ambi_bfb_first_valid_order_batch(ctx, 1, dag, &reversed, result);
}

// --all (printAll): every order the reference would print, in its print order.
// getBFB visits the orders of the initial orientation lexicographically and prints each valid one; if
// the lexicographically LAST topological order fails there, it flips the orientation and goes through
// all orders again (LGM:3691-3695), printing the valid ones of the flipped orientation as well.
// shared body of ambi_bfb_all_valid_orders / ambi_bfb_all_valid_paths
static int all_valid_impl(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, int32_t max_orders,
                          int32_t *orders_out, int32_t *bkp_out, int32_t bkp_stride, int32_t *bkp_len_out,
                          const int32_t *inv_src, const int32_t *inv_tgt, int32_t max_id,
                          int32_t *path_out, int32_t path_stride, int32_t *path_len_out,
                          int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested)
{
    cudaSetDevice(ctx->cfg.device);
    *n_initial = *n_flipped = 0;
    unsigned long long tested = 0;
    const int k = dag->k;
    if (k <= 0) { if (n_tested) *n_tested = 0; return AMBI_OK; }
    const bool surgery = path_out != nullptr;
    std::vector<std::vector<int32_t>> first, second, bfirst, bsecond, pfirst, psecond;
    bool trunc = false;
    const int fwd0 = reversed ? 0 : 1;
    auto sb_status = [](int r) { return r == -2 ? AMBI_ERR_BAD_ARG : (r == -3 ? AMBI_ERR_TOO_LARGE : AMBI_ERR_CUDA); };
    int rc = ctx->sb.run_all(*dag, fwd0, max_orders, first, bfirst, &tested, &trunc, inv_src, inv_tgt, max_id, path_stride, surgery ? &pfirst : nullptr);
    if (rc) return fail(ctx, sb_status(rc), ctx->sb.error());
    if (trunc) return fail(ctx, AMBI_ERR_TOO_LARGE, "more valid orders than max_orders");
    // lexicographically last topological order: always take the highest free index
    std::vector<int32_t> last;
    {
        std::vector<int> indeg(k, 0); std::vector<char> used(k, 0);
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) if ((dag->adj[i] >> j) & 1) indeg[j]++;
        for (int step = 0; step < k; ++step) {
            int pick = -1;
            for (int i = k - 1; i >= 0; --i) if (!used[i] && indeg[i] == 0) { pick = i; break; }
            if (pick < 0) break;                       // cyclic: no complete order exists
            used[pick] = 1; last.push_back(pick);
            for (int j = 0; j < k; ++j) if ((dag->adj[pick] >> j) & 1) indeg[j]--;
        }
    }
    const bool have_orders = (int)last.size() == k;
    const bool last_valid = have_orders && std::binary_search(first.begin(), first.end(), last);
    if (have_orders && !last_valid) {
        rc = ctx->sb.run_all(*dag, 1 - fwd0, max_orders - (int32_t)first.size() > 0 ? max_orders - (int32_t)first.size() : 1, second, bsecond, &tested, &trunc,
                             inv_src, inv_tgt, max_id, path_stride, surgery ? &psecond : nullptr);
        if (rc) return fail(ctx, sb_status(rc), ctx->sb.error());
        if (trunc) return fail(ctx, AMBI_ERR_TOO_LARGE, "more valid orders than max_orders");
    }
    if ((int64_t)first.size() + (int64_t)second.size() > max_orders) return fail(ctx, AMBI_ERR_TOO_LARGE, "more valid orders than max_orders");
    size_t o = 0, row = 0;
    for (auto *lst : {&first, &second}) {
        const auto &bl = lst == &first ? bfirst : bsecond;
        const auto &pl = lst == &first ? pfirst : psecond;
        for (size_t i = 0; i < lst->size(); ++i, ++row) {
            if (orders_out) for (int32_t x : (*lst)[i]) orders_out[o++] = x;
            if (bkp_out && bkp_len_out && bkp_stride > 0) {
                const int len = (int)std::min<size_t>(bl[i].size(), (size_t)bkp_stride);
                bkp_len_out[row] = len;
                for (int t = 0; t < len; ++t) bkp_out[row * (size_t)bkp_stride + t] = bl[i][t];
            }
            if (surgery) {
                path_len_out[row] = (int32_t)pl[i].size();
                for (size_t t = 0; t < pl[i].size(); ++t) path_out[row * (size_t)path_stride + t] = pl[i][t];
            }
        }
    }
    *n_initial = (int32_t)first.size(); *n_flipped = (int32_t)second.size();
    if (n_tested) *n_tested = tested;
    return AMBI_OK;
}

extern "C" int ambi_bfb_all_valid_orders(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, int32_t max_orders,
                                         int32_t *orders_out, int32_t *bkp_out, int32_t bkp_stride, int32_t *bkp_len_out,
                                         int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested)
{
    if (!ctx || !dag || !orders_out || !n_initial || !n_flipped || max_orders <= 0) return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_all_valid_orders: bad argument");
    return all_valid_impl(ctx, dag, reversed, max_orders, orders_out, bkp_out, bkp_stride, bkp_len_out, nullptr, nullptr, 0, nullptr, 0, nullptr,
                          n_initial, n_flipped, n_tested);
}

// --all with the post-search path surgery on the device: every path getBFB would print, as
// signed vertex ids, after imperfectFBI (LGM:3431-3512) and the expansion of the breakpoint pairs (LGM:3661-3670).
extern "C" int ambi_bfb_all_valid_paths(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, const int32_t *inv_src, const int32_t *inv_tgt,
                                        int32_t max_id, int32_t max_orders, int32_t *orders_out, int32_t *path_out, int32_t path_stride,
                                        int32_t *path_len_out, int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested)
{
    if (!ctx || !dag || !inv_src || !inv_tgt || max_id < 0 || !path_out || !path_len_out || path_stride <= 0 || !n_initial || !n_flipped || max_orders <= 0)
        return fail(ctx, AMBI_ERR_BAD_ARG, "ambi_bfb_all_valid_paths: bad argument");
    return all_valid_impl(ctx, dag, reversed, max_orders, orders_out, nullptr, 0, nullptr, inv_src, inv_tgt, max_id, path_out, path_stride, path_len_out,
                          n_initial, n_flipped, n_tested);
}

extern "C" int ambi_bfb_target_cn(int32_t n, const int32_t *x, int32_t *target)
{
    if (!x || !target || n < 1 || n > AMBI_MAX_SEGMENTS) return AMBI_ERR_BAD_ARG;
    const int H = n * (n + 1) / 2;
    for (int i = 0; i < n; ++i) target[i] = 0;
    for (int a = 0; a < n; ++a)
        for (int b = a; b < n; ++b) {
            const int w = x[pair_index(n, a, b)] + 2 * x[H + pair_index(n, a, b)];
            for (int i = a; i <= b && w; ++i) target[i] += w;
        }
    return AMBI_OK;
}

extern "C" int ambi_bfb_expand_path(const int32_t *bkp, int32_t bkp_len, int32_t *path_out, int32_t cap)
{
    return ambi::expand_path(bkp, bkp_len, path_out, cap);      // LGM:3661-3670 (path_surgery.h)
}

// imperfectFBI (LGM:3431-3512), host side of Seam B: rewrite fold vertices whose fold-back junction
// joins two different segments.  inv_src/inv_tgt are indexed by absolute segment id (0 = none) and
// carry the endpoints of inversions[id] as getJuncCN recorded them (LGM:4014-4048).
extern "C" int ambi_bfb_imperfect_fbi(int32_t *bkp, int32_t len, const int32_t *inv_src, const int32_t *inv_tgt, int32_t max_id)
{
    if (!bkp || len < 0 || (len & 1) || !inv_src || !inv_tgt) return AMBI_ERR_BAD_ARG;
    ambi::imperfect_fbi(bkp, len, inv_src, inv_tgt, max_id);     // one implementation for host and device (path_surgery.h)
    return AMBI_OK;
}
