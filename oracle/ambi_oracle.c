/*
 * ambi_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see ambi_oracle.h).
 *
 * Plain-C restatement of the reference's BFB fold-pattern search.  Every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference; LGM = src/LocalGenomicMap.cpp, main = localhap.cpp).
 *
 * Nothing here is copied from the reference: the ILP is restated as direct
 * integer row evaluations, the solver (external Cbc in the reference) is an
 * independent exact branch and bound, Stage B is re-coded on signed-id arrays.
 */
#include "ambi_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* numbering                                                           */
/* ------------------------------------------------------------------ */

int orc_num_vars(int n) { return n * (n + 1); }

/* combinations(start,end,2) emits (a,b), a<=b, a ascending then b ascending
 * (LGM:3254-3264); variableIdx maps "p:a,b" to that position and "l:a,b" to
 * position + n(n+1)/2 (main:122-133).                                        */
int orc_pair_index(int n, int a, int b) { return a * n - a * (a - 1) / 2 + (b - a); }

#define PX(a, b) x[orc_pair_index(n, (a), (b))]
#define LX(a, b) x[H + orc_pair_index(n, (a), (b))]

/* ------------------------------------------------------------------ */
/* Stage A model: row-for-row evaluation of BFB_ILP                    */
/* ------------------------------------------------------------------ */

int orc_model_eval(int n, const int32_t *c2, const int32_t *j2, int32_t loop_ub,
                   int n_tgs, const int32_t *tgs_pairs, int junc_info,
                   const int32_t *x, int64_t *score2)
{
    return orc_model_eval_units(n, c2, j2, 1, loop_ub, n_tgs, tgs_pairs, junc_info, x, score2);
}

/* The same rows with the right-hand sides in units of 1/(2*unit) copies: c2[i] = 2*unit*c_i,
 * j2[i] = 2*unit*J_i (any copy number that is a multiple of 1/(2*unit)); *score2 = 2*unit*sum(eps). */
int orc_model_eval_units(int n, const int32_t *c2, const int32_t *j2, int unit, int32_t loop_ub,
                         int n_tgs, const int32_t *tgs_pairs, int junc_info,
                         const int32_t *x, int64_t *score2)
{
    const int H = n * (n + 1) / 2;
    int64_t tot = 0;
    int ok = 1;
    /* variable bounds, LGM:4712-4724 */
    for (int a = 0; a < n; a++)
        for (int b = a; b < n; b++) {
            if (PX(a, b) < 0 || PX(a, b) > 1) ok = 0;
            if (LX(a, b) < 0 || LX(a, b) > loop_ub) ok = 0;
        }
    /* segment rows (LGM:4426-4451) and fold-back rows (LGM:4455-4494) */
    for (int i = 0; i < n; i++) {
        int64_t cov = 0;
        for (int a = 0; a <= i; a++)
            for (int b = i; b < n; b++) cov += PX(a, b) + 2 * LX(a, b);
        int64_t d = 2 * (int64_t)unit * cov - c2[i];
        tot += d < 0 ? -d : d;
        /* coef[] of LGM:4455-4477, doubled: loops 2 (coef 1), patterns 1 (0.5) */
        int64_t fb2 = 0;
        for (int a = 0; a < n; a++)
            for (int b = a; b < n; b++)
                if (a == i || b == i) fb2 += 2 * LX(a, b);
        for (int a = 0; a < n; a++)
            for (int b = a; b < n; b++) {
                /* pattern (a,b) receives 0.5 iff some other pattern shares the
                 * endpoint i on the same side and has a different length
                 * (pair loop LGM:4464-4476 sets both members of the pair). */
                int half = 0;
                for (int a2 = 0; a2 < n && !half; a2++)
                    for (int b2 = a2; b2 < n; b2++) {
                        if ((a == i && a2 == i) || (b == i && b2 == i)) {
                            int l1 = b - a, l2 = b2 - a2;
                            if (l1 != l2) { half = 1; break; }
                        }
                    }
                if (half) fb2 += PX(a, b);
            }
        int64_t f = (int64_t)unit * fb2 - j2[i];
        tot += f < 0 ? -f : f;
    }
    /* pattern parent / children rows, LGM:4544-4583 */
    for (int a = 0; a < n; a++)
        for (int b = a; b < n; b++) {
            if (a > 0 || b < n - 1) {
                int64_t s = -PX(a, b);
                for (int j = 0; j < a; j++) s += PX(j, b);
                for (int j = b + 1; j < n; j++) s += PX(a, j);
                if (s < 0) ok = 0;
            }
            if (a < b) {
                int64_t s = PX(a, b);
                for (int j = a; j < b; j++) s += PX(a, j);
                for (int j = a + 1; j <= b; j++) s += PX(j, b);
                if (s > 2 || s < 0) ok = 0;
            }
        }
    /* loop parent rows, LGM:4587-4612 */
    for (int a = 0; a < n; a++)
        for (int b = a; b < n; b++)
            if (a > 0 || b < n - 1) {
                int64_t s = -LX(a, b);
                for (int j = 0; j < a; j++) s += PX(j, b) + LX(j, b);
                for (int j = b + 1; j < n; j++) s += PX(a, j) + LX(a, j);
                if (s < 0) ok = 0;
            }
    /* loop children rows, LGM:4615-4646 */
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) {
            int64_t c = 0;
            for (int j = a; j < b; j++) c += LX(a, j);
            for (int j = a + 1; j <= b; j++) c += LX(j, b);
            if (c + LX(a, b) > 2 || c + LX(a, b) < 0) ok = 0;
            if (c + PX(a, b) > 2 || c + PX(a, b) < 0) ok = 0;
        }
    /* mixed children rows, LGM:4649-4681 */
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) {
            int64_t s1 = PX(a, b), s2 = PX(a, b);
            for (int j = a; j < b; j++) { s1 += LX(a, j); s2 += PX(a, j); }
            for (int j = a + 1; j <= b; j++) { s1 += PX(j, b); s2 += LX(j, b); }
            if (s1 > 2 || s1 < 0 || s2 > 2 || s2 < 0) ok = 0;
        }
    /* third-generation row, LGM:4684-4703 */
    if (n_tgs > 0 && junc_info) {
        int64_t s = 0;
        for (int t = 0; t < n_tgs; t++) {
            int sa = tgs_pairs[2 * t], se = tgs_pairs[2 * t + 1];
            if (sa == 0 && se == n - 1) continue;
            int dup = 0;
            for (int u = 0; u < t; u++)
                if (tgs_pairs[2 * u] == sa && tgs_pairs[2 * u + 1] == se) dup = 1;
            if (dup) continue;
            s += LX(sa, se) + PX(sa, se);
        }
        if (s > 5 || s < 0) ok = 0;
    }
    if (score2) *score2 = tot;
    return ok;
}

/* search-order comparison of two assignments: rows a=0..n-1, P[a,a..] then
 * L[a,a..]; returns <0 if x is smaller. */
static int cmp_search_order(int n, const int32_t *x, const int32_t *y)
{
    const int H = n * (n + 1) / 2;
    for (int a = 0; a < n; a++) {
        for (int b = a; b < n; b++) {
            int i = orc_pair_index(n, a, b);
            if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
        }
        for (int b = a; b < n; b++) {
            int i = H + orc_pair_index(n, a, b);
            if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
        }
    }
    return 0;
}

int64_t orc_bruteforce(int n, const int32_t *c2, const int32_t *j2, int32_t loop_ub,
                       int n_tgs, const int32_t *tgs_pairs, int junc_info,
                       int lmax, int32_t *x_out, int64_t *n_opt)
{
    if (n < 1 || n > 4) return -1;
    const int H = n * (n + 1) / 2, N = 2 * H;
    int lcap = loop_ub < lmax ? loop_ub : lmax;
    int32_t x[64];
    memset(x, 0, sizeof x);
    int64_t best = -1, cnt = 0;
    for (;;) {
        int64_t s;
        if (orc_model_eval(n, c2, j2, loop_ub, n_tgs, tgs_pairs, junc_info, x, &s)) {
            if (best < 0 || s < best) {
                best = s; cnt = 1; memcpy(x_out, x, N * sizeof(int32_t));
            } else if (s == best) {
                cnt++;
                if (cmp_search_order(n, x, x_out) < 0) memcpy(x_out, x, N * sizeof(int32_t));
            }
        }
        int i = 0;
        for (; i < N; i++) {
            int cap = i < H ? 1 : lcap;
            if (x[i] < cap) { x[i]++; break; }
            x[i] = 0;
        }
        if (i == N) break;
    }
    if (n_opt) *n_opt = cnt;
    return best;
}

/* ------------------------------------------------------------------ */
/* Stage A exact search                                                */
/* ------------------------------------------------------------------ */
/*
 * Rows of the upper-triangular variable grid are fixed in order a = 0..n-1
 * (a = start segment offset).  After row a every variable touching segment a
 * is fixed, so its two deviation terms are exact:
 *     cov_a  = sum_{a'<=a<=b} (P+2L)[a',b]                    (LGM:4427-4440)
 *     fb2_a  = sum_{a'<a} (P+2L)[a',a] + sum_b (P+2L)[a,b]    (LGM:4458-4483)
 * What later rows need from earlier ones is, per column b > a:
 *     nP[b], nL[b]  entries already in column b (parents for rows ①③,
 *                   committed coverage / fold-back for the bound)
 *     bP[b], bL[b]  how many more P / L entries column b may still take,
 *                   = min over earlier rows a' of the slack of rows ②④⑤ at
 *                   (a',b) (LGM:4557-4582, 4615-4681).
 * plus the running count of the third-generation row ⑥ (LGM:4684-4703).
 */

typedef struct {
    int8_t nP[ORC_MAXN], nL[ORC_MAXN], bP[ORC_MAXN], bL[ORC_MAXN];
    int32_t tgs;
} st_t;

typedef struct {
    int32_t tot;      /* err + future lower bound                       */
    int32_t err;
    int8_t p1, p2;    /* pattern end columns (p1 < p2), -1 if absent    */
    int8_t l1, l2;    /* loop end columns (l1 <= l2); l1==l2 => mult 2  */
    st_t st;
} child_t;

typedef struct { uint8_t key[ORC_MAXN + 2]; int32_t score; uint8_t used; } memo_ent;

typedef struct {
    int n, loop_ub, has_tgs;
    int unit;                   /* scores and right-hand sides are in units of 1/(2*unit) copies */
    int c2[ORC_MAXN], j2[ORC_MAXN];
    uint8_t tgs[ORC_MAXN][ORC_MAXN];
    /* search */
    int64_t limit;              /* children with score+tot > limit are cut   */
    int mode;                   /* 0 best-first improve, 1 first in lex order,
                                   2 first in reverse lex order              */
    int found;
    int8_t cur[ORC_MAXN][4];    /* current row choices p1,p2,l1,l2           */
    int8_t best[ORC_MAXN][4];
    int64_t best_score;
    int64_t cands, nodes;
    memo_ent *memo; size_t memo_cap, memo_cnt;
    int shard_world, shard_rank, shard_level;   /* prefix sharding (multi-GPU emulation) */
} srch_t;

/* row-configuration codes, identical to ambigram_b200/csrc/ambi_common.cuh: ends numbered from the
 * right (q = n-1-b) so that integer order of (ip<<10|il) is the search order */
static int ip_code(int q1, int q2) { if (q1 < 0) return 0; int base = 1 + q1 * (q1 + 1) / 2; return q2 < 0 ? base : base + 1 + q2; }
static int il_code(int q1, int q2)
{
    if (q1 < 0) return 0;
    int base = 1 + 2 * q1 + q1 * (q1 - 1) / 2;
    if (q2 < 0) return base;
    return q2 == q1 ? base + 1 + q1 : base + 1 + q2;
}
static uint32_t row_code(int n, const int8_t r[4])
{
    int ip = r[0] < 0 ? 0 : ip_code(n - 1 - r[0], r[1] < 0 ? -1 : n - 1 - r[1]);
    int il = r[2] < 0 ? 0 : il_code(n - 1 - r[2], r[3] < 0 ? -1 : n - 1 - r[3]);
    return ((uint32_t)ip << 10) | (uint32_t)il;
}
static uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
/* which shard owns the prefix rows[0..level] (same hash as k_expand) */
static int prefix_shard(int n, int8_t rows[][4], int level, int world)
{
    uint32_t h = 0;
    for (int lane = 0; lane < 32; lane++) {
        uint32_t v = lane <= level ? row_code(n, rows[lane]) : 0u;
        h ^= mix32(v + 0x9e3779b9u * (uint32_t)(lane + 1));
    }
    return (int)(mix32(h) % (uint32_t)world);
}

#define BUD_INF 3

static uint64_t hash_key(const uint8_t *k, int len)
{
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < len; i++) { h ^= k[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h;
}

static void memo_reset(srch_t *S)
{
    if (!S->memo) { S->memo_cap = 1 << 16; S->memo = calloc(S->memo_cap, sizeof(memo_ent)); }
    else memset(S->memo, 0, S->memo_cap * sizeof(memo_ent));
    S->memo_cnt = 0;
}

static memo_ent *memo_find(srch_t *S, const uint8_t *key)
{
    size_t mask = S->memo_cap - 1;
    size_t i = hash_key(key, ORC_MAXN + 2) & mask;
    while (S->memo[i].used && memcmp(S->memo[i].key, key, ORC_MAXN + 2) != 0) i = (i + 1) & mask;
    return &S->memo[i];
}

static void memo_grow(srch_t *S)
{
    memo_ent *old = S->memo; size_t oc = S->memo_cap;
    S->memo_cap *= 2; S->memo = calloc(S->memo_cap, sizeof(memo_ent));
    for (size_t i = 0; i < oc; i++)
        if (old[i].used) { memo_ent *e = memo_find(S, old[i].key); *e = old[i]; }
    free(old);
}

/* returns 1 if the state should be skipped (already expanded at <= score) */
static int memo_visit(srch_t *S, int a, const st_t *st, int score)
{
    uint8_t key[ORC_MAXN + 2];
    memset(key, 0, sizeof key);
    for (int b = a; b < S->n; b++)
        key[b] = (uint8_t)(st->nP[b] | (st->nL[b] << 2) | (st->bP[b] << 4) | (st->bL[b] << 6));
    key[ORC_MAXN] = (uint8_t)a; key[ORC_MAXN + 1] = (uint8_t)st->tgs;
    memo_ent *e = memo_find(S, key);
    if (e->used) {
        if (e->score <= score) return 1;
        e->score = score; return 0;
    }
    memcpy(e->key, key, sizeof key); e->score = score; e->used = 1;
    if (++S->memo_cnt * 2 > S->memo_cap) memo_grow(S);
    return 0;
}

/* Admissible bound on the deviation of segments j > a given the column
 * contents after row a: coverage / fold-back already committed can only grow
 * (over-shoot term); what can still be added is capped by the column budgets
 * and by 6 per remaining row (under-shoot term).                            */
static int future_bound(const srch_t *S, int a, const st_t *t)
{
    int n = S->n, lb = 0, cov = 0, capsum = 0;
    const int U = S->unit;
    for (int j = n - 1; j > a; j--) {
        int w = t->nP[j] + 2 * t->nL[j];
        cov += w;
        int cp = t->bP[j] < 2 ? t->bP[j] : 2, cl = t->bL[j] < 2 ? t->bL[j] : 2;
        int cap = cp + 2 * cl;
        capsum += cap;
        int fmax = 6 * (j - a) < capsum ? 6 * (j - a) : capsum;
        int d = 2 * U * cov - S->c2[j];
        if (d > 0) lb += d; else if (-d - 2 * U * fmax > 0) lb += -d - 2 * U * fmax;
        int gmax = cap + (capsum < 6 ? capsum : 6);
        int f = U * w - S->j2[j];
        if (f > 0) lb += f; else if (-f - U * gmax > 0) lb += -f - U * gmax;
    }
    return lb;
}

/* Try one row configuration; on success append to out[]. */
static void try_config(srch_t *S, int a, const st_t *st, int covC, int fbC, int slack,
                       int p1, int p2, int l1, int l2, child_t **out, size_t *cnt, size_t *cap)
{
    const int n = S->n;
    S->cands++;
    int np = (p1 >= 0) + (p2 >= 0), nl = (l1 >= 0) + (l2 >= 0);
    int w = np + 2 * nl;
    int fbw = (n == 1) ? 2 * nl : w;      /* pattern coefficient is 0 when n==1 */
    int e1 = 2 * S->unit * (covC + w) - S->c2[a]; if (e1 < 0) e1 = -e1;
    int e2 = S->unit * (fbC + fbw) - S->j2[a]; if (e2 < 0) e2 = -e2;
    int err = e1 + e2;
    if (err > slack) return;
    st_t t = *st;
    int rp = 0, rl = 0;
    for (int b = a; b < n; b++) {
        int pab = (b == p1) + (b == p2), lab = (b == l1) + (b == l2);
        t.nP[b] += pab; t.nL[b] += lab;
        if (t.bP[b] != BUD_INF || b == 0) t.bP[b] -= pab;   /* column 0 holds one cell */
        if (t.bL[b] != BUD_INF || b == 0) t.bL[b] -= lab;
        if (b > a) {
            int m = rp > rl ? rp : rl;
            int kP = 2 - pab - m;                               /* rows ② ⑤a at (a,b) */
            int k1 = lab + rl, k2 = pab + rl, k3 = pab + rp;    /* rows ④a ④b ⑤b     */
            int kL = 2 - (k1 > k2 ? (k1 > k3 ? k1 : k3) : (k2 > k3 ? k2 : k3));
            if (kP < 0 || kL < 0) return;
            if (kP < t.bP[b]) t.bP[b] = kP;
            if (kL < t.bL[b]) t.bL[b] = kL;
        }
        rp += pab; rl += lab;
        if (S->has_tgs && S->tgs[a][b]) t.tgs += pab + lab;
    }
    if (t.tgs > 5) return;
    int lb = future_bound(S, a, &t);
    if (err + lb > slack) return;
    if (*cnt == *cap) { *cap = *cap ? *cap * 2 : 64; *out = realloc(*out, *cap * sizeof(child_t)); }
    child_t *c = &(*out)[(*cnt)++];
    c->tot = err + lb; c->err = err; c->p1 = p1; c->p2 = p2; c->l1 = l1; c->l2 = l2; c->st = t;
}

/* Enumerate the feasible configurations of row a in search order
 * (P part then L part, each as a digit string with column a most
 * significant): {} < {n-1} < {n-2} < {n-2,n-1} < ...                        */
static void expand_row(srch_t *S, int a, const st_t *st, int slack,
                       child_t **out, size_t *cnt, size_t *cap)
{
    const int n = S->n;
    int covC = 0;
    for (int b = a; b < n; b++) covC += st->nP[b] + 2 * st->nL[b];
    int fbC = st->nP[a] + 2 * st->nL[a];
    int lmax = S->loop_ub < 2 ? S->loop_ub : 2;
    /* pattern sets: pa = lowest column (or -1), pb = higher column (or -1) */
    for (int pa = n; pa >= a; pa--) {
        int pa_ = (pa == n) ? -1 : pa;            /* first iteration: empty set */
        for (int pb = n; pb > (pa_ < 0 ? n - 1 : pa_); pb--) {
            int pb_ = (pb == n) ? -1 : pb;
            if (pa_ < 0 && pb_ >= 0) continue;
            /* feasibility of the pattern part: column budget and row ① */
            if (pa_ >= 0) {
                if (st->bP[pa_] < 1) continue;
                if (pb_ >= 0) {
                    if (st->bP[pb_] < 1) continue;
                    if (!(a == 0 && pb_ == n - 1) && st->nP[pb_] < 1) continue;
                } else {
                    if (!(a == 0 && pa_ == n - 1) && st->nP[pa_] < 1) continue;
                }
            }
            int npar_hi = 0; (void)npar_hi;
            /* loop multisets in search order */
            for (int la = n; la >= a; la--) {
                int la_ = (la == n) ? -1 : la;
                /* order within fixed lowest column la: {la}, {la,lb} lb desc, {la,la} */
                for (int step = 0;; step++) {
                    int l1, l2;
                    if (la_ < 0) { if (step > 0) break; l1 = l2 = -1; }
                    else if (step == 0) { l1 = la_; l2 = -1; }
                    else if (step <= n - 1 - la_) { l1 = la_; l2 = n - step; }
                    else if (step == n - la_) { l1 = l2 = la_; }
                    else break;
                    if (l1 >= 0) {
                        int mult_lo = 1 + (l2 == l1);
                        if (mult_lo > lmax) continue;
                        if (st->bL[l1] < mult_lo) continue;
                        if (l2 >= 0 && l2 != l1 && (lmax < 1 || st->bL[l2] < 1)) continue;
                        /* row ③: parents = column entries above + row entries to the right */
                        int hi = (l2 > l1) ? l2 : l1;
                        {   /* highest loop end */
                            int par = st->nP[hi] + st->nL[hi] + (pa_ > hi) + (pb_ > hi);
                            int need = (hi == l1 && l2 == l1) ? 2 : 1;
                            if (!(a == 0 && hi == n - 1) && par < need) continue;
                        }
                        if (l2 > l1) { /* lower end has the higher loop as a row parent */ }
                    }
                    /* ordering used by child_t: l1 <= l2 with -1 meaning absent */
                    int c1 = l1, c2_ = l2;
                    if (c2_ >= 0 && c1 > c2_) { int t = c1; c1 = c2_; c2_ = t; }
                    try_config(S, a, st, covC, fbC, slack, pa_, pb_, c1, c2_, out, cnt, cap);
                }
            }
        }
    }
}

static int cmp_child_tot(const void *x, const void *y)
{
    const child_t *a = x, *b = y;
    return a->tot - b->tot;
}

static void dfs(srch_t *S, int a, const st_t *st, int score)
{
    const int n = S->n;
    if (S->mode != 0 && S->found) return;
    if (a == n) {
        if (S->mode == 0) {
            if (score < S->best_score) {
                S->best_score = score; memcpy(S->best, S->cur, sizeof S->cur);
                S->limit = score - 1;
            }
        } else {
            S->found = 1; S->best_score = score; memcpy(S->best, S->cur, sizeof S->cur);
        }
        return;
    }
    if (S->shard_world > 1) {
        if (a == S->shard_level + 1 && prefix_shard(n, S->cur, S->shard_level, S->shard_world) != S->shard_rank) return;
        /* merging states across prefixes is only valid once the prefix is owned by one shard */
        if (a > S->shard_level && memo_visit(S, a, st, score)) return;
    } else if (memo_visit(S, a, st, score)) return;
    S->nodes++;
    child_t *ch = NULL; size_t cnt = 0, cap = 0;
    expand_row(S, a, st, (int)(S->limit - score), &ch, &cnt, &cap);
    if (S->mode == 0) qsort(ch, cnt, sizeof(child_t), cmp_child_tot);
    for (size_t q = 0; q < cnt; q++) {
        size_t i = (S->mode == 2) ? cnt - 1 - q : q;
        if (score + ch[i].tot > S->limit) { if (S->mode == 0) break; else continue; }
        S->cur[a][0] = ch[i].p1; S->cur[a][1] = ch[i].p2; S->cur[a][2] = ch[i].l1; S->cur[a][3] = ch[i].l2;
        dfs(S, a + 1, &ch[i].st, score + ch[i].err);
        if (S->mode != 0 && S->found) break;
    }
    free(ch);
}

static void rows_to_x(int n, int8_t rows[][4], int32_t *x)
{
    const int H = n * (n + 1) / 2;
    memset(x, 0, 2 * H * sizeof(int32_t));
    for (int a = 0; a < n; a++) {
        if (rows[a][0] >= 0) PX(a, rows[a][0]) += 1;
        if (rows[a][1] >= 0) PX(a, rows[a][1]) += 1;
        if (rows[a][2] >= 0) LX(a, rows[a][2]) += 1;
        if (rows[a][3] >= 0) LX(a, rows[a][3]) += 1;
    }
}

int orc_solve(int n, const int32_t *c2, const int32_t *j2, int32_t loop_ub,
              int n_tgs, const int32_t *tgs_pairs, int junc_info,
              int want_unique, int32_t *x_out, orc_solve_stats *stats)
{
    return orc_solve_shard(n, c2, j2, loop_ub, n_tgs, tgs_pairs, junc_info, want_unique, 1, 0, -1, x_out, stats);
}

int orc_solve_shard(int n, const int32_t *c2, const int32_t *j2, int32_t loop_ub,
                    int n_tgs, const int32_t *tgs_pairs, int junc_info, int want_unique,
                    int world, int rank, int shard_level, int32_t *x_out, orc_solve_stats *stats)
{
    return orc_solve_units(n, c2, j2, 1, loop_ub, n_tgs, tgs_pairs, junc_info, want_unique, world, rank, shard_level, x_out, stats);
}

/* Exact search with right-hand sides in units of 1/(2*unit) copies (see orc_model_eval_units);
 * stats->score2 = 2*unit*sum(eps).  unit = 1 is the half-integral case every other entry point uses. */
int orc_solve_units(int n, const int32_t *c2, const int32_t *j2, int unit, int32_t loop_ub,
                    int n_tgs, const int32_t *tgs_pairs, int junc_info, int want_unique,
                    int world, int rank, int shard_level, int32_t *x_out, orc_solve_stats *stats)
{
    orc_solve_stats dummy;
    if (!stats) stats = &dummy;
    memset(stats, 0, sizeof *stats);
    stats->unique = -1;
    if (n < 1 || n > ORC_MAXN || unit < 1) { stats->status = -1; return -1; }
    const int H = n * (n + 1) / 2;
    if (n == 1) {
        /* single segment: no a<b rows exist; P[s,s] in {0,1}, L[s,s] in
         * [0,loop_ub] (LGM:4708-4724); pattern fold-back coefficient is 0
         * because the pair loop LGM:4464-4476 never fires.                */
        int64_t best = -1; int bp = 0, bl = 0; int64_t nopt = 0;
        for (int p = 0; p <= 1; p++)
            for (int l = 0; l <= loop_ub; l++) {
                int64_t e1 = 2 * (int64_t)unit * (p + 2 * l) - c2[0]; if (e1 < 0) e1 = -e1;
                int64_t e2 = 2 * (int64_t)unit * l - j2[0]; if (e2 < 0) e2 = -e2;
                stats->candidates++;
                if (best < 0 || e1 + e2 < best) { best = e1 + e2; bp = p; bl = l; nopt = 1; }
                else if (e1 + e2 == best) nopt++;
            }
        x_out[0] = bp; x_out[1] = bl;
        stats->score2 = best; stats->unique = nopt == 1; stats->nodes = 1;
        return 0;
    }
    srch_t *S = calloc(1, sizeof *S);
    S->n = n; S->loop_ub = loop_ub; S->unit = unit;
    S->shard_world = world > 1 ? world : 1; S->shard_rank = rank;
    S->shard_level = shard_level < 0 ? 0 : (shard_level > n - 2 ? n - 2 : shard_level);
    for (int i = 0; i < n; i++) { S->c2[i] = c2[i]; S->j2[i] = j2[i]; }
    if (n_tgs > 0 && junc_info)
        for (int t = 0; t < n_tgs; t++) {
            int sa = tgs_pairs[2 * t], se = tgs_pairs[2 * t + 1];
            if (sa == 0 && se == n - 1) continue;
            if (sa < 0 || se >= n || sa > se) continue;
            S->tgs[sa][se] = 1; S->has_tgs = 1;
        }
    st_t root; memset(&root, 0, sizeof root);
    for (int b = 0; b < n; b++) { root.bP[b] = BUD_INF; root.bL[b] = BUD_INF; }
    /* phase 1: optimum value (best-first depth-first, strict improvement).
     * All-zero is always feasible: its score seeds the incumbent.           */
    int64_t zero = 0;
    for (int i = 0; i < n; i++) zero += c2[i] + j2[i];
    S->mode = 0; S->best_score = zero + 1; S->limit = zero;
    memo_reset(S);
    dfs(S, 0, &root, 0);
    int64_t vstar = S->best_score;
    if (vstar > zero) vstar = zero;      /* shard without an improving assignment: bound by all-zero */
    /* phase 2: first optimum in search order */
    S->mode = 1; S->found = 0; S->limit = vstar;
    memo_reset(S);
    dfs(S, 0, &root, 0);
    if (!S->found) {                 /* only possible for a shard that owns no completed assignment */
        stats->status = S->shard_world > 1 ? 1 : -2; stats->score2 = 0x7fffffff;
        memset(x_out, 0, 2 * H * sizeof(int32_t));
        int rc = S->shard_world > 1 ? 0 : -2;
        free(S->memo); free(S); return rc;
    }
    rows_to_x(n, S->best, x_out);
    stats->score2 = vstar;
    if (want_unique) {
        int32_t *y = malloc(2 * H * sizeof(int32_t));
        S->mode = 2; S->found = 0; S->limit = vstar;
        memo_reset(S);
        dfs(S, 0, &root, 0);
        rows_to_x(n, S->best, y);
        stats->unique = memcmp(y, x_out, 2 * H * sizeof(int32_t)) == 0;
        free(y);
    }
    stats->candidates = S->cands; stats->nodes = S->nodes;
    free(S->memo); free(S);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Stage B                                                             */
/* ------------------------------------------------------------------ */

typedef struct { int kind, a, b, cn; char key[24]; } bnode_t;

static int cmp_key(const void *x, const void *y) { return strcmp(((const bnode_t *)x)->key, ((const bnode_t *)y)->key); }

/* constructDAG, LGM:3276-3378.  Nodes are the components with elementCN>0 in
 * std::map<string,int> key order (LGM:3279: "l:a,b" keys sort before "p:a,b",
 * each group in string order); node2loop is then sorted by decreasing length
 * (LGM:3303).  For <= 16 nodes libstdc++'s std::sort is a plain insertion sort,
 * which with compareLoops (LGM:3267-3274) keeps the loops in front, stably
 * ordered by decreasing b-a; that is the behaviour restated here for every k
 * (the reference's comparator is not a strict weak order and scrambles the
 * table for k > 16 -- see DESIGN.md).                                      */
int orc_construct_dag(int n, int start_id, const int32_t *x, orc_dag *dag)
{
    const int H = n * (n + 1) / 2;
    bnode_t loops[ORC_MAXK * 8], pats[ORC_MAXK * 8];
    int nl = 0, np = 0;
    memset(dag, 0, sizeof *dag);
    for (int a = 0; a < n; a++)
        for (int b = a; b < n; b++) {
            if (PX(a, b) > 0 && np < ORC_MAXK * 8) {
                bnode_t *q = &pats[np++]; q->kind = 0; q->a = a + start_id; q->b = b + start_id; q->cn = PX(a, b);
                snprintf(q->key, sizeof q->key, "%d,%d", q->a, q->b);
            }
            if (LX(a, b) > 0 && nl < ORC_MAXK * 8) {
                bnode_t *q = &loops[nl++]; q->kind = 1; q->a = a + start_id; q->b = b + start_id; q->cn = LX(a, b);
                snprintf(q->key, sizeof q->key, "%d,%d", q->a, q->b);
            }
        }
    if (nl + np > ORC_MAXK) return -1;
    qsort(loops, nl, sizeof(bnode_t), cmp_key);
    qsort(pats, np, sizeof(bnode_t), cmp_key);
    /* stable insertion sort by decreasing length */
    for (int i = 1; i < nl; i++) {
        bnode_t v = loops[i]; int j = i - 1;
        while (j >= 0 && (v.b - v.a) > (loops[j].b - loops[j].a)) { loops[j + 1] = loops[j]; j--; }
        loops[j + 1] = v;
    }
    bnode_t nd[ORC_MAXK];
    int k = 0;
    for (int i = 0; i < nl; i++) nd[k++] = loops[i];
    for (int i = 0; i < np; i++) nd[k++] = pats[i];
    dag->k = k;
    for (int i = 0; i < k; i++) { dag->kind[i] = (int8_t)nd[i].kind; dag->a[i] = (int16_t)nd[i].a; dag->b[i] = (int16_t)nd[i].b; dag->cn[i] = (int16_t)nd[i].cn; }
    /* adjacency in the reference's emission order; parents[] mirrors LGM:3278 */
    static __thread int adjl[ORC_MAXK][ORC_MAXK], adjn[ORC_MAXK], parl[ORC_MAXK][ORC_MAXK], parn[ORC_MAXK];
    memset(adjn, 0, sizeof adjn); memset(parn, 0, sizeof parn);
#define SHARE(i, j) (nd[i].a == nd[j].a || nd[i].b == nd[j].b)
#define LEN(i) (nd[i].b - nd[i].a)
#define ADD(i, j) do { adjl[i][adjn[i]++] = (j); parl[j][parn[j]++] = (i); } while (0)
    for (int i = 0; i < k; i++) {
        if (nd[i].kind != 0) continue;
        for (int j = 0; j < k; j++) if (nd[j].kind == 0 && SHARE(i, j) && LEN(i) > LEN(j)) ADD(i, j);   /* LGM:3314-3323 */
        for (int j = 0; j < k; j++) if (nd[j].kind == 1 && SHARE(i, j) && LEN(i) > LEN(j)) ADD(i, j);   /* LGM:3326-3336 */
    }
    for (int i = 0; i < k; i++) {
        if (nd[i].kind != 1) continue;
        for (int j = 0; j < k; j++) {                                                                   /* LGM:3342-3363 */
            int isparent = 0;
            for (int q = 0; q < parn[i]; q++) if (parl[i][q] == j) isparent = 1;
            if (isparent) continue;
            if (nd[j].kind == 0 && SHARE(i, j)) {
                if (LEN(i) > LEN(j)) ADD(i, j);
                else {
                    for (int q = 0; q < parn[i]; q++) {
                        int par = parl[i][q], hit = 0;
                        for (int r = 0; r < adjn[par]; r++) if (adjl[par][r] == j) hit = 1;
                        if (hit) { ADD(i, j); break; }
                    }
                }
            }
        }
        for (int j = 0; j < k; j++) if (nd[j].kind == 1 && SHARE(i, j) && LEN(i) > LEN(j)) ADD(i, j);   /* LGM:3365-3375 */
    }
    for (int i = 0; i < k; i++) for (int r = 0; r < adjn[i]; r++) dag->adj[i] |= 1ull << adjl[i][r];
    return 0;
}

#define BKP_MAX 4096

/* getBFB body for one order (LGM:3519-3646): returns how many components were placed. */
int orc_assemble_order(const orc_dag *dag, const int32_t *order, int forward, int32_t *B, int32_t *len_out)
{
    int len = 0;
    const int k = dag->k;
    if (k == 0) { *len_out = 0; return 0; }
    {   /* seed, LGM:3524-3568 */
        int v = order[0], a = dag->a[v], b = dag->b[v];
        if (dag->kind[v] == 0) {
            if (forward) { B[0] = a; B[1] = b; } else { B[0] = -b; B[1] = -a; }
            len = 2;
        } else {
            for (int c = 0; c < dag->cn[v]; c++) {
                if (forward) { B[len] = a; B[len + 1] = b; B[len + 2] = -b; B[len + 3] = -a; }
                else { B[len] = -b; B[len + 1] = -a; B[len + 2] = a; B[len + 3] = b; }
                len += 4;
            }
        }
    }
    int i;
    for (i = 1; i < k; i++) {
        int v = order[i], a = dag->a[v], b = dag->b[v];
        if (dag->kind[v] == 0) {                                   /* LGM:3572-3585 */
            if (B[len - 1] == -a) { B[len] = a; B[len + 1] = b; len += 2; }
            else if (B[len - 1] == b) { B[len] = -b; B[len + 1] = -a; len += 2; }
            else break;
        } else {                                                   /* LGM:3586-3644 */
            /* reverse find of a- : forward index f must be odd, and either the last
             * element or id(B[f-1]) >= id(B[f+2]) (LGM:3591-3595) */
            int f = -1, isv1 = 1;
            for (int p = len - 1; p >= 0; p--) {
                if (B[p] != -a) continue;
                if (p % 2 == 0) continue;
                if (p < len - 2 && abs(B[p - 1]) < abs(B[p + 2])) continue;
                f = p; break;
            }
            if (f < 0) {                                           /* LGM:3596-3602 */
                isv1 = 0;
                for (int p = len - 1; p >= 0; p--) {
                    if (B[p] != b) continue;
                    if (p % 2 == 0) continue;
                    if (p < len - 2 && abs(B[p - 1]) > abs(B[p + 2])) continue;
                    f = p; break;
                }
            }
            if (f < 0) break;
            int add = 4 * dag->cn[v];
            if (len + add > BKP_MAX) break;
            if (f + 1 < len) B[f + 1] = isv1 ? a : -b;             /* LGM:3626-3628, 3639-3641 */
            memmove(B + f + 1 + add, B + f + 1, (size_t)(len - f - 1) * sizeof(int32_t));
            for (int c = 0; c < dag->cn[v]; c++) {
                int32_t *q = B + f + 1 + 4 * c;
                if (isv1) { q[0] = a; q[1] = b; q[2] = -b; q[3] = -a; }
                else { q[0] = -b; q[1] = -a; q[2] = a; q[3] = b; }
            }
            len += add;
        }
    }
    *len_out = len;
    return i;
}

typedef struct {
    const orc_dag *dag; int forward;
    int indeg[ORC_MAXK]; int visited[ORC_MAXK];
    int32_t res[ORC_MAXK]; int depth;
    int found; int32_t order[ORC_MAXK]; int32_t bkp[BKP_MAX]; int32_t bkp_len;
    int64_t tested, valid, cap; int count_mode; int naive;
    int32_t *collect; int64_t collect_cap;      /* count_mode: every valid order, in visiting (= lexicographic) order */
} topo_t;

/* allTopologicalOrders (LGM:3380-3409): lowest free index first.  In pruned mode a
 * prefix that cannot be assembled is abandoned (every completion would fail in
 * getBFB); naive mode visits complete orders only, as the reference does.        */
static void topo_rec(topo_t *T)
{
    const orc_dag *d = T->dag;
    if (T->found && !T->count_mode) return;
    if (T->cap > 0 && T->tested >= T->cap) return;
    if (T->depth == d->k) {
        int32_t B[BKP_MAX], len;
        T->tested++;
        int placed = orc_assemble_order(d, T->res, T->forward, B, &len);
        if (placed == d->k) {
            if (T->collect && T->valid < T->collect_cap) memcpy(T->collect + T->valid * d->k, T->res, sizeof(int32_t) * d->k);
            T->valid++;
            if (!T->found) { T->found = 1; memcpy(T->order, T->res, sizeof(int32_t) * d->k); memcpy(T->bkp, B, sizeof(int32_t) * len); T->bkp_len = len; }
        }
        return;
    }
    for (int i = 0; i < d->k; i++) {
        if (T->indeg[i] != 0 || T->visited[i]) continue;
        if (!T->naive && T->depth > 0) {
            /* prefix test: assemble res[0..depth-1] + i as a (depth+1)-node order */
            orc_dag sub = *d; int32_t B[BKP_MAX], len;
            T->res[T->depth] = i;
            sub.k = T->depth + 1;
            if (orc_assemble_order(&sub, T->res, T->forward, B, &len) != sub.k) continue;
        }
        for (int j = 0; j < d->k; j++) if ((d->adj[i] >> j) & 1) T->indeg[j]--;
        T->res[T->depth++] = i; T->visited[i] = 1;
        topo_rec(T);
        T->visited[i] = 0; T->depth--;
        for (int j = 0; j < d->k; j++) if ((d->adj[i] >> j) & 1) T->indeg[j]++;
        if (T->found && !T->count_mode) return;
    }
}

static void topo_init(topo_t *T, const orc_dag *dag, int forward)
{
    memset(T, 0, sizeof *T);
    T->dag = dag; T->forward = forward;
    for (int i = 0; i < dag->k; i++)
        for (int j = 0; j < dag->k; j++) if ((dag->adj[i] >> j) & 1) T->indeg[j]++;     /* main:244-250 */
}

static int first_valid(const orc_dag *dag, int reversed, int naive, int64_t cap, int32_t *order_out,
                       int32_t *used_flipped, int32_t *bkp_out, int32_t *bkp_len, int64_t *n_tested)
{
    topo_t *T = malloc(sizeof *T);
    int64_t tested = 0;
    int ok = 0;
    for (int attempt = 0; attempt < 2 && !ok && dag->k > 0; attempt++) {
        int forward = attempt == 0 ? !reversed : (reversed != 0);     /* LGM:3516, 3691-3695 */
        topo_init(T, dag, forward); T->naive = naive; T->cap = cap;
        topo_rec(T);
        tested += T->tested;
        if (T->found) {
            ok = 1; *used_flipped = attempt;
            memcpy(order_out, T->order, sizeof(int32_t) * dag->k);
            memcpy(bkp_out, T->bkp, sizeof(int32_t) * T->bkp_len); *bkp_len = T->bkp_len;
        }
    }
    if (n_tested) *n_tested = tested;
    free(T);
    return ok;
}

int orc_first_valid_order(const orc_dag *dag, int reversed, int32_t *order_out, int32_t *used_flipped,
                          int32_t *bkp_out, int32_t *bkp_len, int64_t *n_tested)
{
    return first_valid(dag, reversed, 0, 0, order_out, used_flipped, bkp_out, bkp_len, n_tested);
}

/* reference-faithful variant: complete orders only, capped */
int orc_first_valid_order_naive(const orc_dag *dag, int reversed, int64_t cap, int32_t *order_out,
                                int32_t *used_flipped, int32_t *bkp_out, int32_t *bkp_len, int64_t *n_tested)
{
    return first_valid(dag, reversed, 1, cap, order_out, used_flipped, bkp_out, bkp_len, n_tested);
}

int64_t orc_count_valid_orders(const orc_dag *dag, int forward, int64_t cap, int64_t *n_orders)
{
    topo_t *T = malloc(sizeof *T);
    topo_init(T, dag, forward); T->count_mode = 1; T->cap = cap;
    topo_rec(T);
    int64_t v = T->valid;
    if (n_orders) *n_orders = T->tested;
    free(T);
    return v;
}

/* printAll (LGM:3672-3685): every valid order of one orientation, lexicographic; returns the count
 * (orders beyond cap are counted but not stored). */
int64_t orc_all_valid_orders(const orc_dag *dag, int forward, int32_t *orders_out, int64_t cap)
{
    topo_t *T = malloc(sizeof *T);
    topo_init(T, dag, forward); T->count_mode = 1; T->collect = orders_out; T->collect_cap = cap;
    topo_rec(T);
    int64_t v = T->valid;
    free(T);
    return v;
}

int orc_expand_path(const int32_t *bkp, int32_t bkp_len, int32_t *path_out, int32_t cap)
{
    int len = 0;                                                   /* LGM:3661-3670 */
    for (int j = 1; j < bkp_len; j += 2) {
        int u = bkp[j - 1], v = bkp[j];
        if (u > 0) { for (int k = u; k <= v; k++) { if (len < cap) path_out[len] = k; len++; } }
        else { for (int k = -u; k >= -v; k--) { if (len < cap) path_out[len] = -k; len++; } }
    }
    return len;
}

/* ------------------------------------------------------------------ */
/* batch driver for the CPU baseline (bench.py): independent problems  */
/* are solved in parallel with OpenMP, one problem per thread at a time */
/* ------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* problems are packed: n[i], then c2/j2 at offsets off[i] (each n[i] ints);
 * x_out at xoff[i]; no third-generation rows in the packed form. */
int orc_solve_batch(int count, const int32_t *n, const int32_t *off, const int32_t *c2, const int32_t *j2,
                    const int32_t *loop_ub, const int64_t *xoff, int32_t *x_out, int64_t *score2,
                    int64_t *candidates, int threads)
{
    int bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int i = 0; i < count; i++) {
        orc_solve_stats st;
        int rc = orc_solve(n[i], c2 + off[i], j2 + off[i], loop_ub[i], 0, NULL, 0, 0, x_out + xoff[i], &st);
        if (rc) bad++;
        score2[i] = st.score2; candidates[i] = st.candidates;
    }
    return bad;
}

/* ------------------------------------------------------------------ */
/* imperfectFBI (LGM:3431-3512)                                        */
/* ------------------------------------------------------------------ */
/* A fold-back junction may join two different (adjacent) segments; after the
 * breakpoint path is assembled the reference rewrites the fold vertices to the
 * junction's real endpoints.  bkp holds signed ids; inversions[id] is given as
 * inv_src[id] / inv_tgt[id] (source / target segment id of that junction, 0 if
 * the segment has none), as recorded by getJuncCN (LGM:4014-4048).             */
void orc_imperfect_fbi(int32_t *B, int32_t len, const int32_t *inv_src, const int32_t *inv_tgt, int32_t max_id)
{
#define HAS(id) ((id) >= 1 && (id) <= max_id && inv_src[(id)] != 0)
    int pos = 0;
    while (pos < len) {
        /* r = find(pos+3, end, complement(*pos)), LGM:3439 (guarded: pos+3 may pass the end) */
        int r = len;
        for (int q = pos + 3; q < len; q++) if (B[q] == -B[pos]) { r = q; break; }
        int l = r - 1;
        if (r == len || B[l] != -B[pos + 1]) {                     /* LGM:3442-3472 */
            int id = abs(B[pos + 1]);
            if (HAS(id)) {
                int s = inv_src[id], t = inv_tgt[id];
                if (B[pos + 1] > 0) B[pos + 1] = s < t ? s : t;     /* smaller id, '+' */
                else B[pos + 1] = -(s < t ? t : s);                 /* larger id, '-'  */
            }
            if (pos > 0) {
                id = abs(B[pos]);
                if (HAS(id) && abs(B[pos - 1]) == id) {
                    int other = inv_src[id] == id ? inv_tgt[id] : inv_src[id];
                    B[pos] = B[pos] > 0 ? other : -other;
                }
            }
            if (B[pos] > 0 && abs(B[pos]) > abs(B[pos + 1])) B[pos + 1] = B[pos];   /* adjustment */
            if (B[pos] < 0 && abs(B[pos]) < abs(B[pos + 1])) B[pos + 1] = B[pos];
            pos += 2;
        } else {                                                   /* LGM:3473-3509 */
            int p1 = pos + (l - pos) / 2, p2 = p1 + 1;
            while (p1 >= pos - 1 && p1 > 0) {
                int id = abs(B[p1]);
                if (HAS(id)) {
                    int s = inv_src[id], t = inv_tgt[id];
                    if (B[p1] > 0) {
                        if (s < t) { B[p1] = s; B[p1 + 1] = -t; } else { B[p1] = t; B[p1 + 1] = -s; }
                    } else {
                        if (s < t) { B[p1] = -t; B[p1 + 1] = s; } else { B[p1] = -s; B[p1 + 1] = t; }
                    }
                    if (p2 != p1 + 1) {
                        if (p1 > pos - 1) B[p2] = -B[p1];
                        B[p2 - 1] = -B[p1 + 1];
                    }
                }
                p1 -= 2; p2 += 2;
            }
            pos = r + 1;
        }
    }
#undef HAS
}
