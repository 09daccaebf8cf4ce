/*
 * ambigram_b200.h -- C ABI of libambigram_b200.so: the B200 (sm_100a) drop-in
 * for Ambigram's BFB fold-pattern search (`Ambigram --op bfb`).
 *
 * The reference has no plugin/FFI boundary for this path.  It sits behind
 *   (i)  LocalGenomicMap member functions called only from main()
 *        (include/LocalGenomicMap.hpp:191-208), and
 *   (ii) a subprocess plus two text files: `<prefix>.lp` written by BFB_ILP
 *        (src/LocalGenomicMap.cpp:4397-4752), solved by an external `cbc`
 *        (localhap.cpp:179-181), `<prefix>.sol` parsed back
 *        (localhap.cpp:184-212).
 * The entry points below are what a C++ maintainer binds at those two seams
 * (INTEGRATION.md shows the edit to localhap.cpp):
 *
 *   Seam A  ambi_bfb_solve / ambi_bfb_solve_batch
 *           replaces  BFB_ILP + system("cbc ...") + .sol parse
 *                     (localhap.cpp:173-212)
 *   Seam B  ambi_bfb_construct_dag + ambi_bfb_first_valid_order
 *           replaces  constructDAG + allTopologicalOrders + the order loop of
 *                     getBFB (localhap.cpp:235-261,
 *                     src/LocalGenomicMap.cpp:3276-3409, 3514-3697)
 *
 * Plain pointers and sizes only; all buffers are HOST memory owned by the
 * caller (the reference owns everything with raw new[], localhap.cpp:155-156).
 * The library owns device memory, streams and (optionally) the NCCL
 * communicator inside ambi_ctx.  Calls are blocking; one thread per ctx.
 * There is no CPU fallback: without a CUDA device ambi_init fails with
 * AMBI_ERR_NO_DEVICE.
 */
#ifndef AMBIGRAM_B200_H
#define AMBIGRAM_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMBI_ABI_VERSION 1
#define AMBI_MAX_SEGMENTS 32      /* n per partition (BASELINE configs: 6..30) */
#define AMBI_MAX_COMPONENTS 64    /* k <= 2n active patterns/loops            */

/* status codes (0 = ok; the reference's conventions are exit(1) on I/O error,
 * localhap.cpp:187-190, and the "Infeasible" fallback, localhap.cpp:213-220) */
enum {
    AMBI_OK                    = 0,
    AMBI_INFEASIBLE            = 1,   /* never produced: all-zero is feasible  */
    AMBI_ERR_NO_DEVICE         = -1,
    AMBI_ERR_BAD_ARG           = -2,
    AMBI_ERR_TOO_LARGE         = -3,  /* n > AMBI_MAX_SEGMENTS                 */
    AMBI_ERR_CUDA              = -4,
    AMBI_ERR_FRONTIER_OVERFLOW = -5,  /* candidate buffers exhausted           */
    AMBI_ERR_NOT_HALF_INTEGRAL = -6,  /* unit_shift out of range (hosts: copy numbers not on the grid) */
    AMBI_ERR_NCCL              = -7
};

typedef struct ambi_ctx ambi_ctx;

typedef struct {
    int32_t device;            /* CUDA device ordinal                           */
    int32_t rank;              /* this process' shard (0 if single GPU)         */
    int32_t world;             /* number of shards (1 if single GPU)            */
    int32_t beam_width;        /* incumbent pass width per problem (0 = 16)     */
    int64_t node_capacity;     /* candidate records per frontier buffer (0 = auto) */
    int32_t shard_level;       /* rows expanded redundantly before the frontier
                                  is split by prefix across shards (0 = auto)  */
    uint32_t flags;            /* AMBI_CFG_*                                    */
} ambi_config;

/* ambi_config.flags */
#define AMBI_CFG_SMALL_KERNEL 2u /* partitions of <= 16 segments are first tried by the warp-per-partition kernel (whole search
                                  in shared memory, one launch per batch; stage_small.cu); what does not fit its record
                                  capacity joins the others on the level-synchronous global-memory path.  Off by default:
                                  measured slower than the global-memory path on B200. */

/* One partition's Stage-A problem: what BFB_ILP reads from the graph.
 *   seg_cn2[i] = 2*c_i  (segment copy number after getIndelBias,
 *                        LGM:3699-3744, rows LGM:4442-4449)
 *   fbi_cn2[i] = 2*J_i  (juncCN[i][1] from getJuncCN, LGM:3989-4050, rows
 *                        LGM:4485-4492)
 *   loop_ub    = floor(sum of CN over ALL graph segments) (LGM:4708-4723)
 *   bias       = localhap.cpp:141-146 (constant in the objective)
 *   tgs_pairs  = (min,max) absolute segment ids of each readComponents()
 *                component kept for this partition (localhap.cpp:158-162,
 *                LGM:4687-4689); the row is active iff n_tgs>0 && junc_info */
typedef struct {
    int32_t n;                 /* endID - startID + 1                           */
    int32_t start_id;          /* absolute id of the first segment              */
    const int32_t *seg_cn2;    /* [n] */
    const int32_t *fbi_cn2;    /* [n] */
    int32_t loop_ub;
    int32_t bias;
    int32_t n_tgs;
    const int32_t *tgs_pairs;  /* [2*n_tgs] */
    uint32_t flags;            /* bit0 = junc_info (--junc_info true)           */
    uint32_t unit_shift;       /* copy numbers in units of 1/(2U) copies, U = 1 << unit_shift (0..6): seg_cn2[i] = 2U*c_i,
                                  fbi_cn2[i] = 2U*J_i.  0 = half copies.  The reference's model takes any real copy number
                                  (Weight.hpp:11, Graph.cpp:369-405); a host passes the finest grid its inputs need, or
                                  rounds to 1/128 copy (unit_shift 6).  score2 / objective2 come back in the same unit. */
} ambi_bfb_problem;

#define AMBI_FLAG_JUNC_INFO 1u

/* Result of Stage A, in the reference's own terms:
 *   element_cn[x] for x < n(n+1) in variableIdx numbering -- pattern block
 *   then loop block, each in combinations() order (localhap.cpp:113-133,
 *   155-156, 204-211).  objective2 = 2*(sum of eps) - 2*bias, i.e. twice what
 *   cbc prints as "objective value" (localhap.cpp:199-203).                  */
typedef struct {
    int32_t *element_cn;       /* [n(n+1)] caller-allocated                     */
    int64_t score2;            /* 2U * sum(eps)  (integer; U = 1 << unit_shift)  */
    int64_t objective2;        /* score2 - 2U*bias: divide by 2U for what cbc prints as "objective value" */
    uint64_t candidates_scored;/* row configurations of the expanded records the device disposed of (by the
                                  column-mask bound, the O(1) prefilter or the O(n) bound)               */
    uint64_t nodes_expanded;   /* partial assignments expanded                  */
    int32_t status;
    int32_t incumbent_score2;  /* score2 found by the beam (incumbent) pass     */
    double seconds;            /* time to proof: device time (%globaltimer) from the start of the search of the batch until
                                  this partition's optimum was proven                                          */
    double seconds_to_best;    /* time to best: ... until the returned optimum was first seen as a completed assignment */
} ambi_bfb_solution;

/* Stage-B inputs: the DAG of active components (constructDAG, LGM:3276-3378). */
typedef struct {
    int32_t k;                              /* nodes                            */
    int8_t  kind[AMBI_MAX_COMPONENTS];      /* 0 pattern, 1 loop                */
    int16_t a[AMBI_MAX_COMPONENTS];         /* absolute segment ids, a<=b       */
    int16_t b[AMBI_MAX_COMPONENTS];
    int16_t cn[AMBI_MAX_COMPONENTS];        /* multiplicity (node2loop[i][2])   */
    uint64_t adj[AMBI_MAX_COMPONENTS];      /* bit j: edge i -> j               */
} ambi_dag;

typedef struct {
    int32_t found;                          /* 1 if some order assembles        */
    int32_t used_flipped;                   /* orientation flipped, LGM:3691-3695 */
    int32_t order[AMBI_MAX_COMPONENTS];     /* winning order (node indices)     */
    int32_t bkp_len;                        /* breakpoint path length           */
    int32_t bkp[1024];                      /* signed ids: +id '+', -id '-'     */
    uint64_t orders_tested;                 /* (partial) orders assembled on device */
    int32_t status;
    int32_t reserved;
} ambi_order_result;

/* ---- lifecycle ---------------------------------------------------------- */
int  ambi_abi_version(void);
int  ambi_init(ambi_ctx **ctx, const ambi_config *cfg);
void ambi_destroy(ambi_ctx *ctx);
const char *ambi_last_error(const ambi_ctx *ctx);
const char *ambi_status_string(int status);

/* Multi-GPU: one process per GPU.  The 128-byte NCCL unique id is created on
 * rank 0 (ambi_nccl_unique_id) and distributed by the host launcher; every
 * rank then calls ambi_comm_init.  Without it, world>1 contexts still shard
 * work but the final combine must be done by the caller (see
 * ambi_bfb_solve_shard).                                                     */
int ambi_nccl_unique_id(void *id128);
int ambi_comm_init(ambi_ctx *ctx, const void *id128);

/* ---- Seam A -------------------------------------------------------------- */
/* Solve one partition (all GPUs of the ctx's world cooperate by prefix when
 * world>1 and a communicator is present).                                    */
int ambi_bfb_solve(ambi_ctx *ctx, const ambi_bfb_problem *problem, ambi_bfb_solution *solution);

/* Solve `count` independent partitions in one device pass (frontiers of all
 * problems share the candidate buffers).  With world>1 the caller shards the
 * batch by sample across ranks; no collective is involved.                   */
int ambi_bfb_solve_batch(ambi_ctx *ctx, int32_t count, const ambi_bfb_problem *problems,
                         ambi_bfb_solution *solutions);

/* Prefix-sharded solve without a communicator: this rank explores only the
 * subtrees whose prefix hashes to ctx.rank and reports its local optimum
 * (score2 and element_cn in search order are comparable across ranks with
 * ambi_bfb_better).  Used by the gloo/CPU-side combine and by tests.         */
int ambi_bfb_solve_shard(ambi_ctx *ctx, const ambi_bfb_problem *problem, ambi_bfb_solution *solution);

/* Tie-break order of two solutions of the same problem: returns 1 if A is
 * strictly better than B (lower score2, or equal score2 and smaller in search
 * order: rows a=0..n-1, P[a,a..] then L[a,a..]).                             */
int ambi_bfb_better(int32_t n, int64_t score2_a, const int32_t *element_cn_a,
                    int64_t score2_b, const int32_t *element_cn_b);

/* ---- Seam B -------------------------------------------------------------- */
/* constructDAG on the host of the boundary (O(k^2), LGM:3276-3378).          */
int ambi_bfb_construct_dag(int32_t n, int32_t start_id, const int32_t *element_cn, ambi_dag *dag);

/* First topological order (lexicographic in node numbering) from which a BFB
 * path assembles; orientation per --reversed, flipped if none (LGM:3691-3695). */
int ambi_bfb_first_valid_order(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed,
                               ambi_order_result *result);
int ambi_bfb_first_valid_order_batch(ambi_ctx *ctx, int32_t count, const ambi_dag *dags,
                                     const int32_t *reversed, ambi_order_result *results);

/* --all (printAll, LGM:3672-3685): every order getBFB would print, in print order:
 * the valid orders of the initial orientation in lexicographic order, followed --
 * only if the lexicographically last topological order fails there (LGM:3691-3695)
 * -- by the valid orders of the flipped orientation.  orders_out holds
 * (*n_initial + *n_flipped) rows of dag->k node indices; AMBI_ERR_TOO_LARGE if more
 * than max_orders exist.  Row r's breakpoint path (before imperfectFBI) is
 * bkp_out[r*bkp_stride .. + bkp_len_out[r]) (both may be NULL).                 */
int ambi_bfb_all_valid_orders(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, int32_t max_orders,
                              int32_t *orders_out, int32_t *bkp_out, int32_t bkp_stride, int32_t *bkp_len_out,
                              int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested);

/* --all with the post-search path surgery done on the device: row r of path_out
 * (path_stride entries per row) holds the path_len_out[r] signed vertex ids (+id '+', -id '-') the reference prints
 * for the r-th valid order -- breakpoint path, imperfectFBI (LGM:3431-3512; inv_src/inv_tgt as for
 * ambi_bfb_imperfect_fbi), expansion of the pairs (LGM:3661-3670) -- in the print order of
 * ambi_bfb_all_valid_orders.  orders_out may be NULL.  AMBI_ERR_TOO_LARGE if a path has more than path_stride vertices. */
int ambi_bfb_all_valid_paths(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, const int32_t *inv_src,
                             const int32_t *inv_tgt, int32_t max_id, int32_t max_orders, int32_t *orders_out,
                             int32_t *path_out, int32_t path_stride, int32_t *path_len_out,
                             int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested);

/* imperfectFBI (LGM:3431-3512), host bookkeeping for the winning order only: fold
 * vertices whose junction joins two different segments are rewritten in place.
 * inv_src/inv_tgt[id] = endpoints of inversions[id] (0 = none), id <= max_id.  */
int ambi_bfb_imperfect_fbi(int32_t *bkp, int32_t bkp_len, const int32_t *inv_src,
                           const int32_t *inv_tgt, int32_t max_id);

/* Expand breakpoint pairs to the vertex path (LGM:3661-3670); returns length. */
int ambi_bfb_expand_path(const int32_t *bkp, int32_t bkp_len, int32_t *path_out, int32_t cap);

/* targetCN (localhap.cpp:222-232): copy number of every segment implied by element_cn,
 * target_cn[i] = sum over components covering segment i of cn (patterns) / 2*cn (loops). */
int ambi_bfb_target_cn(int32_t n, const int32_t *element_cn, int32_t *target_cn);

/* ---- device-resident batch (bench: inputs already in HBM) ---------------- */
/* Upload a batch once, then run the search repeatedly without host<->device
 * copies of the inputs; results stay on the device until fetched.            */
typedef struct ambi_batch ambi_batch;
int  ambi_batch_upload(ambi_ctx *ctx, int32_t count, const ambi_bfb_problem *problems, ambi_batch **batch);
int  ambi_batch_run(ambi_ctx *ctx, ambi_batch *batch);                 /* async on ctx stream */
int  ambi_batch_fetch(ambi_ctx *ctx, ambi_batch *batch, ambi_bfb_solution *solutions);
void ambi_batch_free(ambi_ctx *ctx, ambi_batch *batch);
void *ambi_ctx_stream(ambi_ctx *ctx);                                  /* cudaStream_t */
int  ambi_ctx_sync(ambi_ctx *ctx);
/* Per-kernel timing of the last ambi_batch_run/solve (CUDA events around every
 * launch of the expansion+scoring kernel; only when enabled: it serialises
 * launches and is meant for the bench's roofline pass).                      */
typedef struct {
    double expand_ms;          /* sum of k_expand launch durations              */
    uint64_t expand_launches;
    double total_ms;           /* whole search (incumbent + exact pass)         */
    double incumbent_pass_ms;
    uint64_t record_stride;    /* bytes per candidate record in HBM             */
    uint64_t configs_enumerated; /* row configurations enumerated one by one (light test); the others were
                                    dropped with their columns before the enumeration                  */
    uint64_t bounds_evaluated;   /* ... that reached the O(n) admissible bound  */
    uint64_t children_written;   /* ... that were materialised as records       */
    uint64_t bound_batches;      /* 32-lane batches the bound ran in (lane use = bounds_evaluated / (32*bound_batches)) */
    uint64_t exchange_rounds;    /* bound-exchange rounds (NCCL all-reduce of the pruning limits) since ambi_init */
    uint64_t chunked_runs;       /* searches that continued depth-first because a level did not fit, since ambi_init */
    double small_ms;             /* duration of the warp-per-partition kernel (partitions of <= 16 segments) of the last run */
    uint64_t small_solved;       /* partitions it solved in the last run ...                                   */
    uint64_t global_solved;      /* ... and partitions that took the global-memory path (longer, or frontier too wide) */
    uint64_t small_capacity;     /* records per level and warp of the warp-per-partition kernel               */
    double stage_b_kernel_ms;    /* device time of the last ambi_bfb_first_valid_order[_batch] kernel (Stage B)      */
} ambi_profile;
int  ambi_ctx_set_profile(ambi_ctx *ctx, int enable);
int  ambi_ctx_get_profile(ambi_ctx *ctx, ambi_profile *out);
/* On-box roofline denominators that MEASURED_PEAKS.json does not carry: integer issue
 * rate of the alu pipe alone (IADD3/LOP3 chains), of the alu and fma pipes together (IADD3 + IMAD),
 * and conflict-free shared-memory load bandwidth.  A few milliseconds of micro-benchmark kernels. */
typedef struct {
    double alu_tops;           /* 1e12 thread-level INT32 operations / s, alu pipe only      */
    double alu_fma_tops;       /* ... with every other operation an IMAD (both pipes)       */
    double smem_tbs;           /* 1e12 bytes / s of LDS.128                                  */
    int32_t sms;
    int32_t reserved;
} ambi_peaks;
int  ambi_measure_peaks(ambi_ctx *ctx, ambi_peaks *out);
/* kernels launched by the ctx since creation (bench "gpu_launches") */
uint64_t ambi_ctx_launch_count(const ambi_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
