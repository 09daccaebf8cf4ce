// stage_a.h -- host driver of the Stage-A device search (see stage_a.cu).
#pragma once
#include "ambi_common.cuh"
#include "stage_small.h"
#include <vector>

namespace ambi {

struct LevelArgs;
struct Op;

struct HostResult {
    int64_t score2;
    int32_t incumbent;
    int32_t status;
    unsigned long long candidates, nodes;
    uint32_t rows[AMBI_MAXN];
    double seconds, seconds_to_best;
};

class StageA {
public:
    StageA();
    ~StageA();
    int init(cudaStream_t stream, int64_t node_capacity, int beam_width, int rank, int world, int shard_level, bool use_small = false);
    void release();
    // pinned staging for `count` problem descriptors; the caller fills it and then calls upload(ptr, count)
    DevProblem *staging(int count);
    // copy `count` problem descriptors to the device (async on the ctx stream)
    // `copied`: the caller has already enqueued the H2D copies of the problems on the stream (in pieces, from the threads
    // that flattened them: device_problems())
    // `n_arr` / `ush_arr`: the problems' lengths and units as compact arrays (a pass over the 416-byte staged structs that
    // other threads have just written costs milliseconds at 32768 problems)
    int upload(const DevProblem *host_probs, int count, bool copied = false, const int32_t *n_arr = nullptr, const int8_t *ush_arr = nullptr);
    DevProblem *device_problems() { return d_probs_; }
    // Compact upload of a large host batch (capi.cu ambi_batch_upload): a problem travels as 4 header words + the 2c|2J
    // words of its segments (+ the third-generation masks when any partition has some), `words` 32-bit words per problem,
    // and is expanded into DevProblem on the device -- 64 instead of 544 bytes per 12-segment partition over PCIe.
    static int compact_words(int nmax, bool any_tgs) { return 4 + nmax * (any_tgs ? 2 : 1); }
    uint32_t *compact_staging(int count);                  // pinned, room for compact_words(32, true) per problem
    uint32_t *device_compact() { return d_compact_; }
    int unpack_compact(int count, int nmax, bool any_tgs); // after the pieces were copied to device_compact()
    // enqueue the whole search (incumbent pass + exact pass); no host sync
    int run(bool sharded);
    // wait and read back per-problem results (pinned staging; pointers stay valid until the next fetch)
    struct Fetched {
        const int32_t *score, *incumbent; const uint32_t *rows; int rows_stride;   // rows_stride words per problem
        const unsigned long long *cand, *node; bool overflow; double seconds, seconds_beam; bool used_graph;
        // per partition, seconds since the search was enqueued (device %globaltimer): when the returned optimum was first
        // seen as a completed assignment (time to best) and when it was proven (time to proof)
        const double *t_best, *t_proof;
    };
    int fetch_raw(Fetched *out);
    int fetch(HostResult *out);
    const char *error() const { return err_; }
    unsigned long long launches() const { return launches_; }
    int count() const { return count_; }
    uint32_t capacity_base() const { return cap_base_; }
    unsigned long long chunked_runs() const { return chunked_runs_; }
    // partitions of the last run solved by the warp-per-partition kernel / handed on to the global-memory path
    int small_solved() const { return small_solved_; }
    int global_solved() const { return global_count_; }
    double small_ms() const { return prof_small_ms_; }
    int small_capacity() const { return small_cfg_.C; }
    unsigned long long fused_stops() const { return fused_stops_; }
    // per-kernel timing (CUDA events around every k_expand launch; adds launch gaps, so only
    // enabled for the bench's roofline pass)
    void set_profile(bool on) { profile_ = on; }
    void profile(double *expand_ms, unsigned long long *expand_launches, double *total_ms, double *beam_ms) const {
        *expand_ms = prof_expand_ms_; *expand_launches = prof_expand_launches_; *total_ms = prof_total_ms_; *beam_ms = prof_beam_ms_;
    }
    size_t record_stride() const { return (size_t)stride_; }
    // Prefix-sharded search over several GPUs: in-place all-reduce of `count` int32 on the ctx stream
    // (op 0 = min, 1 = max, 2 = sum), supplied by the C-ABI layer when an NCCL communicator exists.  Used for
    // the periodic exchange of pruning limits while ranks search depth-first (run_chunked).
    typedef int (*AllReduceFn)(void *user, int32_t *d_buf, int count, int op);
    void set_allreduce(AllReduceFn fn, void *user) { allreduce_ = fn; allreduce_user_ = user; }
    unsigned long long exchange_rounds() const { return exchange_rounds_; }
    // whole-run funnel of the last fetched search: configurations enumerated, bounds evaluated, children written
    const unsigned long long *funnel() const { return funnel_; }

private:
    int ensure_problem_arrays(int count);
    void free_problem_arrays();
    int run_level(int level, int cur, int mode, int shard_world);
    void launch_expand(const LevelArgs &L);
    int op(const Op &o);
    int launch_op(const Op &o);
    int op_memset(void *ptr, int byte, size_t bytes, bool sync_after = true);
    int run_fused(int first, int last);
    int run_pass(int mode, int shard_world, bool first);
    int enqueue_incumbent(bool coop);
    int enqueue_exact(bool sharded);
    int capture(cudaGraphExec_t *exec, unsigned long long *nlaunch, int part, bool flag);
    void drop_graphs();
    int run_chunked(bool sharded, int start_level, bool exchange);
    int run_chunked_core(bool sharded, int start_level, bool exchange);
    int exchange_round(int active, int *total_active);

    cudaStream_t stream_ = nullptr;
    uint32_t cap_ = 0, cap_base_ = 0, cap_max_ = 0, tsize_ = 0, gen_ = 1;
    size_t buf_bytes_ = 0, pool_bytes_ = 0;
    uint8_t *pool_ = nullptr;
    bool last_sharded_ = false, chunked_failed_ = false, scaled_ = false;
    unsigned long long chunked_runs_ = 0;
    int beam_shift_ = 2;
    int beam_floor_max_ = 16;     // cap of the batch-scaled beam width `16*4096/count` (measured: with the width growing
                                  // in n a wide beam for small batches only costs time; AMBI_BEAM_FLOOR_MAX restores it)
    float beam_growth_ = 0.5f;    // beam width doubles every 1/growth segments beyond 12 (see k_beam_select)
    int gpw_ = 2;                 // k_expand: record groups per resident warp (load balance vs lane utilisation)
    int beam_delta_ = 8;          // heuristic band of the incumbent beam (see k_expand)
    // ... and for partitions of up to beam_short_n_ segments (measured on the B200, DESIGN.md section 5: the band is a
    // trade between the beam's own cost and the incumbent it leaves the exact pass; results do not depend on it)
    int beam_delta_short_ = 16, beam_shift_short_ = 3, beam_short_n_ = 20;
    int beam_ = 64, rank_ = 0, world_ = 1, shard_level_ = 0, shard_level_eff_ = -1;
    int stride_ = 0, sms_ = 148;
    uint32_t leaf_levels_ = 0;      // bit a: some problem of the batch has n == a+1
    uint8_t *buf_[2] = {nullptr, nullptr};      // full records of the current / next frontier
    uint8_t *kbuf_ = nullptr;                   // key records of a level's children
    uint2 *lbuf_ = nullptr;                     // their (parent, row code) links
    uint8_t *dbuf_ = nullptr;                   // their dropped-by-the-merge flags
    int dedup_shift_ = 1;                       // merge table: at least 1 << shift slots per record (measured: 1 beats 2 and 0)
    int kstride_ = 48;
    uint32_t *list_[2] = {nullptr, nullptr};
    unsigned long long *table_ = nullptr;
    uint32_t *counters_ = nullptr;
    size_t expand_smem_[2] = {0, 0};
    int expand_blocks_[2] = {148, 148};
    // per-problem arrays
    int prob_cap_ = 0, count_ = 0, nmax_ = 0;
    DevProblem *d_probs_ = nullptr, *h_probs_ = nullptr;   // device array, pinned host staging
    uint32_t *d_compact_ = nullptr, *h_compact_ = nullptr;  // compact form of a host batch (see compact_words)
    uint8_t *h_res_ = nullptr;                              // pinned result staging
    int h_cap_ = 0;
    int32_t *d_limit_ = nullptr, *d_incumbent_ = nullptr, *d_best_score_ = nullptr, *d_beam_score_ = nullptr, *d_fmin_ = nullptr;
    uint32_t *d_best_rows_ = nullptr, *d_beam_rows_ = nullptr, *d_hist_ = nullptr, *d_ticket_ = nullptr, *d_cutq_ = nullptr;
    unsigned long long *d_cand_ = nullptr, *d_node_ = nullptr, *d_funnel_ = nullptr;
    unsigned long long funnel_[4] = {0, 0, 0, 0};
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long launches_ = 0;
    bool profile_ = false, capturing_ = false, used_graph_ = false;
    cudaGraphExec_t graph_exec_ = nullptr, graph_exec2_ = nullptr;
    int graph_mode_ = 0, beam_div_ = 1;
    unsigned long long graph_launches2_ = 0;
    unsigned long long graph_launches_ = 0;
    std::vector<cudaEvent_t> prof_ev_;
    size_t prof_used_ = 0;
    double prof_expand_ms_ = 0, prof_total_ms_ = 0, prof_beam_ms_ = 0;
    unsigned long long prof_expand_launches_ = 0;
    AllReduceFn allreduce_ = nullptr;
    void *allreduce_user_ = nullptr;
    int32_t *d_flag_ = nullptr;
    // warp-per-partition kernel (stage_small.cu): partitions of 2..16 segments
    bool small_on_ = false;
    SmallConfig small_cfg_;
    uint32_t *d_small_list_ = nullptr, *d_arena_ = nullptr;
    uint8_t *d_active_ = nullptr, *d_active_base_ = nullptr;     // active: still to be solved by the global-memory path
    int32_t *d_seen_ = nullptr;
    unsigned long long *d_stamps_ = nullptr, *d_stamp0_ = nullptr;
    uint32_t *h_flag_ = nullptr;                                 // pinned: partitions the small kernel handed on
    std::vector<double> h_tbest_, h_tproof_;
    int n_small_ = 0, n_large_ = 0, small_solved_ = 0, global_count_ = 0;
    bool ran_global_ = false;
    double prof_small_ms_ = 0;
    cudaEvent_t ev_small_[2] = {nullptr, nullptr};
    int prepare();
    // the launch sequence as a program (stage_a.cu "The launch sequence of a search as a PROGRAM")
    std::vector<Op> *rec_ = nullptr;
    std::vector<Op> prog_;
    Op *d_prog_ = nullptr;
    size_t d_prog_cap_ = 0;
    int prog_split_ = 0, prog_mode_ = -1;
    bool prog_valid_ = false, fused_on_ = false;
    int fused_max_count_ = 0, fused_blocks_[2] = {148, 148};      // 0: off (measured no faster than graph replay; AMBI_FUSED_MAX=64 turns it on)
    uint32_t fused_big_ = 32768;
    unsigned long long fused_stops_ = 0;
    unsigned long long exchange_rounds_ = 0;
    char err_[256];
};

}  // namespace ambi
