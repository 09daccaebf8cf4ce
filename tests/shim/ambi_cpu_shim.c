/*
 * ambi_cpu_shim.c -- TEST INFRASTRUCTURE ONLY (never shipped, never loaded by the product path).
 *
 * Lets the C++ `--op bfb` host driver (ambigram_b200/host/ambigram_bfb.cpp) run on a box without a GPU: preloaded
 * in front of libambigram_b200.so (LD_PRELOAD), it answers the few entry points that need the device --
 * ambi_init / ambi_destroy / ambi_bfb_solve_batch / ambi_bfb_first_valid_order_batch -- with the CPU oracle
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
        int rc = orc_solve(p->n, p->seg_cn2, p->fbi_cn2, p->loop_ub, nt, tg, (p->flags & AMBI_FLAG_JUNC_INFO) != 0, 0,
                           solutions[i].element_cn, &st);
        if (rc) { g_err = "shim: orc_solve failed"; return AMBI_ERR_CUDA; }
        solutions[i].score2 = st.score2; solutions[i].objective2 = st.score2 - 2 * (int64_t)p->bias;
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
