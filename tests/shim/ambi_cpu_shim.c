/*
 * ambi_cpu_shim.c -- TEST INFRASTRUCTURE ONLY (never shipped, never loaded by the product path).
 *
 * Lets the C++ `--op bfb` host driver (ambigram_b200/host/ambigram_bfb.cpp) run on a box without a GPU: preloaded
 * in front of libambigram_b200.so (LD_PRELOAD), it answers the few entry points that need the device --
 * ambi_init / ambi_destroy / ambi_bfb_solve[_batch] / ambi_bfb_first_valid_order[_batch] / ambi_bfb_all_valid_orders -- with the CPU oracle
 * (oracle/ambi_oracle.c).  Everything else the driver calls (ambi_bfb_construct_dag, ambi_bfb_imperfect_fbi,
 * ambi_bfb_expand_path: host code of the library) is the library's own.  What is under test is the driver: its
 * .lh / .juncs readers, depth -> copy number, getJuncCN / getIndelBias flattening, indelBFB and printing, against
 * the reference binary's stdout (tests/test_host_driver_cpu.py).
 */
#include "../../include/ambigram_b200.h"
#include "../../oracle/ambi_oracle.h"
#include <stdlib.h>
#include <string.h>

struct ambi_ctx { int dummy; };
static const char *g_err = "";

int ambi_init(ambi_ctx **ctx, const ambi_config *cfg) { (void)cfg; *ctx = (ambi_ctx *)calloc(1, sizeof(struct ambi_ctx)); return *ctx ? AMBI_OK : AMBI_ERR_CUDA; }
void ambi_destroy(ambi_ctx *ctx) { free(ctx); }
const char *ambi_last_error(const ambi_ctx *ctx) { (void)ctx; return g_err; }

int ambi_bfb_solve_batch(ambi_ctx *ctx, int32_t count, const ambi_bfb_problem *problems, ambi_bfb_solution *solutions)
{
    (void)ctx;
    for (int i = 0; i < count; i++) {
        const ambi_bfb_problem *p = &problems[i];
        int32_t tg[128];
        int nt = p->n_tgs < 64 ? p->n_tgs : 64;
        for (int t = 0; t < 2 * nt; t++) tg[t] = p->tgs_pairs[t] - p->start_id;     /* the oracle takes 0-based offsets */
        orc_solve_stats st;
        int rc = orc_solve_units(p->n, p->seg_cn2, p->fbi_cn2, 1 << p->unit_shift, p->loop_ub, nt, tg, (p->flags & AMBI_FLAG_JUNC_INFO) != 0, 0,
                                 1, 0, -1, solutions[i].element_cn, &st);
        if (rc) { g_err = "shim: orc_solve failed"; return AMBI_ERR_CUDA; }
        solutions[i].score2 = st.score2; solutions[i].objective2 = st.score2 - ((2 * (int64_t)p->bias) << p->unit_shift);
        solutions[i].candidates_scored = (uint64_t)st.candidates; solutions[i].nodes_expanded = (uint64_t)st.nodes;
        solutions[i].status = AMBI_OK; solutions[i].incumbent_score2 = (int32_t)st.score2;
        solutions[i].seconds = solutions[i].seconds_to_best = 0;
    }
    return AMBI_OK;
}

int ambi_bfb_first_valid_order_batch(ambi_ctx *ctx, int32_t count, const ambi_dag *dags, const int32_t *reversed,
                                     ambi_order_result *results)
{
    (void)ctx;
    for (int i = 0; i < count; i++) {
        orc_dag d;                                  /* same layout as ambi_dag, copied field by field anyway */
        memset(&d, 0, sizeof d);
        d.k = dags[i].k;
        for (int v = 0; v < d.k; v++) { d.kind[v] = dags[i].kind[v]; d.a[v] = dags[i].a[v]; d.b[v] = dags[i].b[v]; d.cn[v] = dags[i].cn[v]; d.adj[v] = dags[i].adj[v]; }
        ambi_order_result *r = &results[i];
        memset(r, 0, sizeof *r);
        int64_t tested = 0;
        r->found = d.k > 0 ? orc_first_valid_order(&d, reversed[i], r->order, &r->used_flipped, r->bkp, &r->bkp_len, &tested) : 0;
        if (!r->found) r->bkp_len = 0;
        r->orders_tested = (uint64_t)tested;
    }
    return AMBI_OK;
}

/* the single-call seams the reference's own main() binds (tools/patches/localhap_seams.patch) */
int ambi_bfb_solve(ambi_ctx *ctx, const ambi_bfb_problem *problem, ambi_bfb_solution *solution)
{
    return ambi_bfb_solve_batch(ctx, 1, problem, solution);
}

int ambi_bfb_first_valid_order(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, ambi_order_result *result)
{
    return ambi_bfb_first_valid_order_batch(ctx, 1, dag, &reversed, result);
}

/* --all: the valid orders of the initial orientation, then -- if the lexicographically last topological order
 * fails there (LGM:3691-3695) -- those of the flipped orientation (orders only; the host replays them) */
int ambi_bfb_all_valid_orders(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, int32_t max_orders,
                              int32_t *orders_out, int32_t *bkp_out, int32_t bkp_stride, int32_t *bkp_len_out,
                              int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested)
{
    (void)ctx; (void)bkp_out; (void)bkp_stride; (void)bkp_len_out;
    orc_dag d;
    memset(&d, 0, sizeof d);
    d.k = dag->k;
    for (int v = 0; v < d.k; v++) { d.kind[v] = dag->kind[v]; d.a[v] = dag->a[v]; d.b[v] = dag->b[v]; d.cn[v] = dag->cn[v]; d.adj[v] = dag->adj[v]; }
    *n_initial = *n_flipped = 0;
    if (n_tested) *n_tested = 0;
    if (d.k <= 0) return AMBI_OK;
    const int fwd0 = reversed ? 0 : 1;
    int64_t n1 = orc_all_valid_orders(&d, fwd0, orders_out, max_orders);
    if (n1 > max_orders) { g_err = "shim: more valid orders than max_orders"; return AMBI_ERR_TOO_LARGE; }
    /* lexicographically last topological order and whether it assembles in the initial orientation */
    int32_t last[AMBI_MAX_COMPONENTS]; int indeg[AMBI_MAX_COMPONENTS] = {0}; char used[AMBI_MAX_COMPONENTS] = {0};
    for (int i = 0; i < d.k; i++) for (int j = 0; j < d.k; j++) if ((d.adj[i] >> j) & 1) indeg[j]++;
    int len = 0;
    for (int step = 0; step < d.k; step++) {
        int pick = -1;
        for (int i = d.k - 1; i >= 0; i--) if (!used[i] && indeg[i] == 0) { pick = i; break; }
        if (pick < 0) break;
        used[pick] = 1; last[len++] = pick;
        for (int j = 0; j < d.k; j++) if ((d.adj[pick] >> j) & 1) indeg[j]--;
    }
    int64_t n2 = 0;
    if (len == d.k) {
        int32_t bkp[4096]; int32_t bl = 0;
        const int placed = orc_assemble_order(&d, last, fwd0, bkp, &bl);
        if (placed < d.k) {
            n2 = orc_all_valid_orders(&d, 1 - fwd0, orders_out + n1 * d.k, max_orders - n1);
            if (n1 + n2 > max_orders) { g_err = "shim: more valid orders than max_orders"; return AMBI_ERR_TOO_LARGE; }
        }
    }
    *n_initial = (int32_t)n1; *n_flipped = (int32_t)n2;
    return AMBI_OK;
}

/* --all with the path surgery: orders from above, then imperfectFBI and expansion per order with the oracle */
int ambi_bfb_all_valid_paths(ambi_ctx *ctx, const ambi_dag *dag, int32_t reversed, const int32_t *inv_src, const int32_t *inv_tgt,
                             int32_t max_id, int32_t max_orders, int32_t *orders_out, int32_t *path_out, int32_t path_stride,
                             int32_t *path_len_out, int32_t *n_initial, int32_t *n_flipped, uint64_t *n_tested)
{
    const int k = dag->k > 0 ? dag->k : 1;
    int32_t *ords = orders_out ? orders_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)max_orders * k);
    int rc = ambi_bfb_all_valid_orders(ctx, dag, reversed, max_orders, ords, NULL, 0, NULL, n_initial, n_flipped, n_tested);
    if (rc == AMBI_OK) {
        orc_dag d;
        memset(&d, 0, sizeof d);
        d.k = dag->k;
        for (int v = 0; v < d.k; v++) { d.kind[v] = dag->kind[v]; d.a[v] = dag->a[v]; d.b[v] = dag->b[v]; d.cn[v] = dag->cn[v]; d.adj[v] = dag->adj[v]; }
        for (int r = 0; r < *n_initial + *n_flipped; r++) {
            int32_t bkp[4096]; int32_t bl = 0;
            const int fwd = (r < *n_initial) ? (reversed ? 0 : 1) : (reversed ? 1 : 0);
            orc_assemble_order(&d, ords + (size_t)r * d.k, fwd, bkp, &bl);
            orc_imperfect_fbi(bkp, bl, inv_src, inv_tgt, max_id);
            path_len_out[r] = orc_expand_path(bkp, bl, path_out + (size_t)r * path_stride, path_stride);
        }
    }
    if (!orders_out) free(ords);
    return rc;
}
