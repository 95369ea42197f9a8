/*
 * swd_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see swd_oracle.h).
 *
 * Plain-C restatement of the reference hot path on flat CSR/CSC arrays.
 * Citations are file:line into /root/reference/src.
 *
 * Deliberate, documented definitions where the reference is undefined:
 *  - posterior-history rings start at 0.0 for every decode call (the reference
 *    never clears them: pyx:39-42, bpgd.cpp:357-358 uses uninitialised new[]);
 *  - the multi-thread branch tree is executed serially and pm ties go to the
 *    first branch in the order main, tree id 1.. (primary, then backup), side 0..
 *    (the reference resolves ties by thread timing, bpgd.cpp:454-458);
 *  - if the main branch's reset fails the correction on the kept columns is 0
 *    (the reference keeps the previous shot's buffer, bpgd.cpp:617-622);
 *  - select_vn returning no candidate (guess_vn == -1) in the single-thread
 *    schedule is treated as a failed branch (the reference indexes vn_mask[-1]).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#include "swd_oracle.h"

#define MAX_PM 10000.0
#define CLIP 50.0
#define BIG 1e308

/* ---------------------------------------------------------------- graph -- */
typedef struct {
    int m, n, nnz;
    int *rp, *rc;   /* CSR: row pointer, column of each entry (ascending)     */
    int *cp, *cr;   /* CSC: col pointer, row of each entry (ascending)        */
    int *c2r;       /* CSC entry -> CSR position (messages live in CSR order) */
} graph;

static void graph_free(graph *g) {
    if (!g) return;
    free(g->rp); free(g->rc); free(g->cp); free(g->cr); free(g->c2r); free(g);
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* columns sel[0..nn) of the CSC matrix (cp, cr) become columns 0..nn) of the
 * new graph (mod2sparse_copycols, mod2sparse.c:239-272; insert keeps rows
 * sorted by column and columns sorted by row, mod2sparse.c:358-432). */
static graph *graph_build(int m, int nn, const int *cp, const int *cr, const int *sel) {
    graph *g = (graph *)calloc(1, sizeof(graph));
    g->m = m; g->n = nn;
    g->cp = (int *)calloc(nn + 1, sizeof(int));
    for (int j = 0; j < nn; j++) {
        int c = sel ? sel[j] : j;
        g->cp[j + 1] = g->cp[j] + (cp[c + 1] - cp[c]);
    }
    g->nnz = g->cp[nn];
    g->cr = (int *)malloc(sizeof(int) * (g->nnz + 1));
    g->c2r = (int *)malloc(sizeof(int) * (g->nnz + 1));
    g->rc = (int *)malloc(sizeof(int) * (g->nnz + 1));
    g->rp = (int *)calloc(m + 2, sizeof(int));
    for (int j = 0; j < nn; j++) {
        int c = sel ? sel[j] : j;
        int d = cp[c + 1] - cp[c];
        memcpy(g->cr + g->cp[j], cr + cp[c], sizeof(int) * d);
        qsort(g->cr + g->cp[j], d, sizeof(int), cmp_int);
        for (int e = g->cp[j]; e < g->cp[j + 1]; e++) g->rp[g->cr[e] + 1]++;
    }
    for (int r = 0; r < m; r++) g->rp[r + 1] += g->rp[r];
    int *fill = (int *)malloc(sizeof(int) * (m + 1));
    memcpy(fill, g->rp, sizeof(int) * m);
    for (int j = 0; j < nn; j++)
        for (int e = g->cp[j]; e < g->cp[j + 1]; e++) {
            int p = fill[g->cr[e]]++;
            g->rc[p] = j; g->c2r[e] = p;
        }
    free(fill);
    return g;
}

/* --------------------------------------------------------------- sorting -- */
typedef struct { double v; int i; } kv;
static int cmp_kv(const void *a, const void *b) {
    const kv *x = (const kv *)a, *y = (const kv *)b;
    if (x->v < y->v) return -1;
    if (y->v < x->v) return 1;
    return x->i - y->i;   /* stability: equal keys keep index order */
}
void orc_index_sort(const double *v, int n, int *cols) {
    kv *t = (kv *)malloc(sizeof(kv) * (n + 1));
    for (int i = 0; i < n; i++) { t[i].v = v[i]; t[i].i = i; }
    qsort(t, n, sizeof(kv), cmp_kv);
    for (int i = 0; i < n; i++) cols[i] = t[i].i;
    free(t);
}

/* ------------------------------------------------- one min-sum iteration -- */
/* Check update in min1/min2/argmin/parity form; identical in value to the
 * reference's exclusive prefix/suffix min + sign count (bpgd.cpp:103-148):
 *   c2b[k] = (min over j != k of |clip(b2c[j])|) * (alpha * (-1)^(s + sum_{j!=k} [b2c[j] <= 0]))
 * with 1e308 standing in for the empty min (degree-1 check).
 * vmask == NULL / cmask == NULL -> everything active (pre-BP). s_c[] is the
 * sign seed per check (original syndrome for pre-BP, residual cn_mask else). */
static int64_t ms_iteration(const graph *g, const int8_t *vmask, const int8_t *cmask,
                            const int8_t *sseed, const double *prior, double alpha,
                            double *b2c, double *c2b, double *hist, int slot, int8_t *dec) {
    int64_t edges = 0;
    for (int c = 0; c < g->m; c++) {
        if (cmask && cmask[c] == -1) continue;
        double m1 = BIG, m2 = BIG; int arg = -1; int par = (sseed[c] == 1);
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) {
            if (vmask && vmask[g->rc[p]] != -1) continue;
            double b = b2c[p];
            if (b > CLIP) b = CLIP; else if (b < -CLIP) b = -CLIP;
            b2c[p] = b;
            double a = fabs(b);
            if (a < m1) { m2 = m1; m1 = a; arg = p; } else if (a < m2) m2 = a;
            if (b <= 0) par ^= 1;
        }
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) {
            if (vmask && vmask[g->rc[p]] != -1) continue;
            double mag = (p == arg) ? m2 : m1;
            int sg = par ^ (b2c[p] <= 0);
            c2b[p] = mag * (alpha * (sg ? -1.0 : 1.0));
        }
    }
    for (int v = 0; v < g->n; v++) {
        if (vmask && vmask[v] != -1) continue;
        double t = prior[v];
        for (int e = g->cp[v]; e < g->cp[v + 1]; e++) {
            if (cmask && cmask[g->cr[e]] == -1) continue;
            int p = g->c2r[e];
            b2c[p] = t; t += c2b[p]; edges++;
        }
        hist[4 * v + slot] = t;
        dec[v] = (t <= 0) ? 1 : 0;
        t = 0.0;
        for (int e = g->cp[v + 1] - 1; e >= g->cp[v]; e--) {
            if (cmask && cmask[g->cr[e]] == -1) continue;
            int p = g->c2r[e];
            b2c[p] += t; t += c2b[p];
        }
    }
    return edges;
}

/* Product-sum check update for the unmasked pre-BP (NOT part of the reference's own sources: it exists only in the
 * third-party `ldpc` package the drivers call as BpOsdDecoder(bp_method="product_sum"), un-vendored and unpinned;
 * this restates its published forward / backward tanh-product form):
 *   c2b[k] = s_c * log((1 + P_k) / (1 - P_k)),  P_k = prod_{j != k} tanh(b2c[j] / 2),  s_c = -1 if the syndrome bit is set,
 * with P_k saturated to +-(1 - 2^-52) so that the logarithm stays finite (|c2b| <= ~36.7).  The variable update is the
 * reference's prefix / suffix form.  PARITY UNPINNED. */
static int g_bp_method = 0;       /* 0 = min-sum (reference), 1 = product-sum */
void orc_set_bp_method(int method) { g_bp_method = method; }
static int64_t ps_iteration(const graph *g, const int8_t *sseed, const double *prior,
                            double *b2c, double *c2b, double *hist, int slot, int8_t *dec) {
    const double PMAX = 1.0 - 2.220446049250313e-16;
    int64_t edges = 0;
    for (int c = 0; c < g->m; c++) {
        double tmp = 1.0;
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) { c2b[p] = tmp; tmp *= tanh(b2c[p] * 0.5); }
        tmp = 1.0;
        const double sg = (sseed[c] == 1) ? -1.0 : 1.0;
        for (int p = g->rp[c + 1] - 1; p >= g->rp[c]; p--) {
            double P = c2b[p] * tmp;
            if (P > PMAX) P = PMAX; else if (P < -PMAX) P = -PMAX;
            c2b[p] = sg * log((1.0 + P) / (1.0 - P));
            tmp *= tanh(b2c[p] * 0.5);
        }
    }
    for (int v = 0; v < g->n; v++) {
        double t = prior[v];
        for (int e = g->cp[v]; e < g->cp[v + 1]; e++) { int p = g->c2r[e]; b2c[p] = t; t += c2b[p]; edges++; }
        hist[4 * v + slot] = t;
        dec[v] = (t <= 0) ? 1 : 0;
        t = 0.0;
        for (int e = g->cp[v + 1] - 1; e >= g->cp[v]; e--) { int p = g->c2r[e]; b2c[p] += t; t += c2b[p]; }
    }
    return edges;
}

/* H*dec == synd over all rows / all columns (mod2sparse_mulvec + compare,
 * bpgd.cpp:185-194; pyx:129-137). tsynd receives H*dec. */
static int synd_match(const graph *g, const int8_t *dec, const int8_t *synd, int8_t *tsynd) {
    int eq = 1;
    for (int c = 0; c < g->m; c++) {
        int s = 0;
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) s ^= (dec[g->rc[p]] != 0);   /* mulvec treats any non-zero as 1 */
        tsynd[c] = (int8_t)s;
        if (s != synd[c]) eq = 0;
    }
    return eq;
}

/* ------------------------------------------------------ unmasked pre-BP -- */
static int bp_pre(const graph *g, const double *llr, const int8_t *synd, int max_iter, double alpha,
                  double *hist, int8_t *dec, int *iters, int64_t *edge_iters) {
    double *b2c = (double *)malloc(sizeof(double) * (g->nnz + 1));
    double *c2b = (double *)malloc(sizeof(double) * (g->nnz + 1));
    int8_t *ts = (int8_t *)malloc(g->m + 1);
    for (int v = 0; v < g->n; v++)
        for (int e = g->cp[v]; e < g->cp[v + 1]; e++) b2c[g->c2r[e]] = llr[v];   /* pyx:55-60 */
    int conv = 0, it;
    for (it = 0; it < max_iter; it++) {
        int64_t ed = g_bp_method ? ps_iteration(g, synd, llr, b2c, c2b, hist, it % 4, dec)
                                 : ms_iteration(g, NULL, NULL, synd, llr, alpha, b2c, c2b, hist, it % 4, dec);
        if (edge_iters) *edge_iters += ed;
        if (synd_match(g, dec, synd, ts)) { conv = 1; it++; break; }
    }
    if (iters) *iters = it;
    free(b2c); free(c2b); free(ts);
    return conv;
}

int orc_bp_decode(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                  int max_iter, double alpha, double *hist, int8_t *dec, int *iters_out) {
    graph *g = graph_build(m, n, cp, cr, NULL);
    int r = bp_pre(g, llr, synd, max_iter, alpha, hist, dec, iters_out, NULL);
    graph_free(g);
    return r;
}

/* ------------------------------------------------------------- BPGD state -- */
typedef struct {
    graph *g;                 /* sub-matrix m x n (n = new_n)               */
    int m, n;
    double *prior, *hist, *b2c, *c2b;
    int8_t *vn_mask, *cn_mask, *error, *synd, *tsynd;
    int *cn_deg, *vn_deg;
    int A, A_sum, C, D;       /* ints, bpgd.hpp:15                          */
    int num_iter, low_error_mode;
    double factor;
    orc_stats *st;
} bpgd;

static bpgd *bpgd_new(graph *g, int num_iter, int low_error_mode, double factor, orc_stats *st) {
    bpgd *s = (bpgd *)calloc(1, sizeof(bpgd));
    s->g = g; s->m = g->m; s->n = g->n;
    s->prior = (double *)calloc(s->n + 1, sizeof(double));
    s->hist = (double *)calloc(4 * s->n + 4, sizeof(double));
    s->b2c = (double *)calloc(g->nnz + 1, sizeof(double));
    s->c2b = (double *)calloc(g->nnz + 1, sizeof(double));
    s->vn_mask = (int8_t *)calloc(s->n + 1, 1);
    s->error = (int8_t *)calloc(s->n + 1, 1);
    s->cn_mask = (int8_t *)calloc(s->m + 1, 1);
    s->synd = (int8_t *)calloc(s->m + 1, 1);
    s->tsynd = (int8_t *)calloc(s->m + 1, 1);
    s->cn_deg = (int *)calloc(s->m + 1, sizeof(int));
    s->vn_deg = (int *)calloc(s->n + 1, sizeof(int));
    s->num_iter = num_iter; s->low_error_mode = low_error_mode; s->factor = factor; s->st = st;
    s->A = -3; s->A_sum = -12; s->C = 30; s->D = 3;
    return s;
}
static void bpgd_free(bpgd *s) {
    if (!s) return;
    free(s->prior); free(s->hist); free(s->b2c); free(s->c2b); free(s->vn_mask); free(s->error);
    free(s->cn_mask); free(s->synd); free(s->tsynd); free(s->cn_deg); free(s->vn_deg); free(s);
}

/* bpgd.cpp:51-80 */
static int bpgd_set(bpgd *s, int vn, int value) {
    if (s->vn_mask[vn] != -1) return (s->vn_mask[vn] == value) ? 0 : -1;
    s->vn_mask[vn] = (int8_t)value; s->error[vn] = (int8_t)value;
    const graph *g = s->g;
    for (int e = g->cp[vn]; e < g->cp[vn + 1]; e++) {
        int cn = g->cr[e];
        if (s->cn_mask[cn] == -1 || s->cn_deg[cn] == 0) return -1;
        int deg = s->cn_deg[cn] - 1;
        if (value) s->cn_mask[cn] = 1 - s->cn_mask[cn];
        s->cn_deg[cn] = deg;
        if (deg == 0) {
            if (s->cn_mask[cn] != 0) return -1;
            s->cn_mask[cn] = -1;
        }
    }
    return 0;
}

/* bpgd.cpp:13-49 */
static int bpgd_peel(bpgd *s) {
    const graph *g = s->g;
    for (;;) {
        int clean = 1;
        for (int cn = 0; cn < s->m; cn++) {
            if (s->cn_mask[cn] == -1) continue;
            if (s->cn_deg[cn] >= 2) continue;
            if (s->cn_deg[cn] <= 0) { s->cn_mask[cn] = -1; continue; }
            clean = 0;
            int vn = -1;
            for (int p = g->rp[cn]; p < g->rp[cn + 1]; p++)
                if (s->vn_mask[g->rc[p]] == -1) { vn = g->rc[p]; break; }
            if (vn == -1) return -1;
            if (bpgd_set(s, vn, s->cn_mask[cn]) == -1) return -1;
        }
        if (clean) return 0;
    }
}

/* bpgd.cpp:82-95 */
static void bpgd_init(bpgd *s) {
    const graph *g = s->g;
    for (int v = 0; v < s->n; v++) {
        if (s->vn_mask[v] != -1) continue;
        for (int e = g->cp[v]; e < g->cp[v + 1]; e++) s->b2c[g->c2r[e]] = s->prior[v];
    }
}

/* bpgd.cpp:199-239.  The sub-matrix graph is built by the caller (shared by
 * all branches of a shot); this restores the per-branch state. */
static int bpgd_reset(bpgd *s, const int *cols, const double *llr, const int8_t *synd) {
    const graph *g = s->g;
    for (int v = 0; v < s->n; v++) s->prior[v] = llr[cols[v]];
    memset(s->vn_mask, -1, s->n);
    for (int c = 0; c < s->m; c++) {
        s->cn_mask[c] = synd[c]; s->synd[c] = synd[c];
        s->cn_deg[c] = g->rp[c + 1] - g->rp[c];
        if (s->cn_deg[c] == 0) s->cn_mask[c] = -1;
    }
    for (int v = 0; v < s->n; v++) s->vn_deg[v] = g->cp[v + 1] - g->cp[v];
    memset(s->error, 0, s->n);
    if (bpgd_peel(s) == -1) return -1;
    bpgd_init(s);
    return 0;
}

/* bpgd.cpp:241-248 */
static void bpgd_set_masks(bpgd *s, const int8_t *vm, const int8_t *cm, const int *cd) {
    memcpy(s->vn_mask, vm, s->n); memcpy(s->error, vm, s->n);
    memcpy(s->cn_mask, cm, s->m); memcpy(s->cn_deg, cd, sizeof(int) * s->m);
    bpgd_init(s);
}

/* bpgd.cpp:250-256 */
static double bpgd_pm(const bpgd *s) {
    double pm = 0;
    for (int v = 0; v < s->n; v++) if (s->error[v]) pm += s->prior[v];
    return pm;
}

/* bpgd.cpp:97-197 */
static int bpgd_bp(bpgd *s) {
    if (s->st) s->st->bp_calls++;
    for (int it = 0; it < s->num_iter; it++) {
        int64_t ed = ms_iteration(s->g, s->vn_mask, s->cn_mask, s->cn_mask, s->prior, s->factor,
                                  s->b2c, s->c2b, s->hist, it % 4, s->error);
        if (s->st) s->st->edge_iters += ed;
        if (synd_match(s->g, s->error, s->synd, s->tsynd)) return 1;
    }
    return 0;
}

/* bpgd.cpp:288-351.  Returns favor (0/1) or -1; *guess_vn = -1 if no candidate. */
static int bpgd_select(bpgd *s, int depth, int *guess_vn) {
    const graph *g = s->g;
    int best = -1, best_neg = -1;
    double sum_best = MAX_PM, sum_best_neg = MAX_PM;
    for (int v = 0; v < s->n; v++) {
        if (s->vn_mask[v] != -1) continue;
        if (s->vn_deg[v] <= 2) continue;
        int num_flip = 0;
        for (int e = g->cp[v]; e < g->cp[v + 1]; e++) {
            int c = g->cr[e];
            if (s->cn_mask[c] == -1) continue;
            if (s->synd[c] != s->tsynd[c]) num_flip++;
        }
        const double *h = s->hist + 4 * v;
        int leA = 1, neg = 1, geC = 1, geD = 1; double sum = 0.0;
        for (int i = 0; i < 4; i++) {
            double l = h[i]; sum += l;
            if (l < s->C) geC = 0;
            if (l < s->D) geD = 0;
            if (l > s->A) leA = 0;
            if (l > 0) neg = 0;
        }
        if (!s->low_error_mode && geC && depth < 4) { if (bpgd_set(s, v, 0) == -1) return -1; }
        else if (!s->low_error_mode && num_flip >= 3 && geD) { if (bpgd_set(s, v, 0) == -1) return -1; }
        else if (!s->low_error_mode && leA && sum < s->A_sum) { if (bpgd_set(s, v, 1) == -1) return -1; }
        else {
            if (sum < sum_best) { sum_best = sum; best = v; }
            if (neg && sum < sum_best_neg) { sum_best_neg = sum; best_neg = v; }
        }
    }
    if (bpgd_peel(s) == -1) return -1;
    if (best_neg != -1) { *guess_vn = best_neg; return 1; }
    *guess_vn = best;
    return (sum_best > 0) ? 0 : 1;
}

/* bpgd.cpp:258-286 */
static int bpgd_decimate_reliable(bpgd *s) {
    int best = -1, sign = 0; double largest = 0.0;
    for (int v = 0; v < s->n; v++) {
        if (s->vn_mask[v] != -1) continue;
        double h = s->hist[4 * v + 3];
        if (fabs(h) > largest) { largest = fabs(h); best = v; sign = (h > 0) ? 0 : 1; }
    }
    if (best == -1) return -1;   /* reference would index vn_mask[-1] */
    if (bpgd_set(s, best, sign) == -1) return -1;
    if (bpgd_peel(s) == -1) return -1;
    return 0;
}

/* ------------------------------------------------ multi-thread GDG tree -- */
typedef struct {
    int valid; int8_t *vn_mask, *cn_mask; int *cn_deg; int vn, value, depth;
} snapshot;

typedef struct { double pm; int8_t *err; int n; } best_t;
static void best_offer(best_t *b, double pm, const int8_t *err) {
    if (getenv("ORC_TRACE")) {
        int wt = 0; unsigned hsh = 0;
        for (int i = 0; i < b->n; i++) if (err[i]) { wt++; hsh = hsh * 31u + (unsigned)i; }
        fprintf(stderr, "[orc] offer pm=%.17g wt=%d hash=%08x %s\n", pm, wt, hsh, pm < b->pm ? "TAKEN" : "");
    }
    if (pm < b->pm) { b->pm = pm; memcpy(b->err, err, b->n); }
}

/* BPGD_main_thread::do_work, bpgd.cpp:591-688 (main branch part) */
static int run_main(bpgd *s, const orc_gdg_params *P, snapshot *side, best_t *best, int *converged) {
    int T = P->max_tree_depth, S = P->max_side_depth;
    int conv = 0;
    s->A = -3; s->A_sum = -12; s->C = 30; s->D = 3;
    for (int depth = 0; depth < P->max_step; depth++) {
        conv = bpgd_bp(s);
        int guess = -1;
        s->A_sum = (depth == 0) ? -16 : -12;
        int favor = bpgd_select(s, depth, &guess);
        if (conv || favor == -1 || guess == -1) {
            if (conv) best_offer(best, bpgd_pm(s), s->error);
            break;
        }
        if (depth >= T && depth < S) {
            snapshot *q = &side[depth - T];
            memcpy(q->vn_mask, s->vn_mask, s->n); memcpy(q->cn_mask, s->cn_mask, s->m);
            memcpy(q->cn_deg, s->cn_deg, sizeof(int) * s->m);
            q->vn = guess; q->value = 1 - favor; q->depth = depth + 1; q->valid = 1;
        }
        if (bpgd_set(s, guess, favor) != -1 && bpgd_peel(s) != -1) continue;
        break;
    }
    *converged = conv;
    return 0;
}

/* BPGD_tree_thread::do_work, bpgd.cpp:435-525 */
static void run_tree(bpgd *s, int id, const orc_gdg_params *P, best_t *best) {
    int T = P->max_tree_depth, steps = P->max_tree_branch_step;
    double own_pm = MAX_PM;
    int on_side = 0, saved = 0, depth;
    int8_t *bvn = (int8_t *)malloc(s->n + 1), *bcn = (int8_t *)malloc(s->m + 1);
    int *bdeg = (int *)malloc(sizeof(int) * (s->m + 1));
    int bvar = -1, bval = 0;
    s->A = -3; s->A_sum = -16; s->C = 30; s->D = 3;
    for (depth = 0; depth < steps + T + 1; depth++) {
        if (depth > 0 && !on_side) s->A_sum = -12;
        if (bpgd_bp(s)) {
            if (s->st) s->st->paths_converged++;
            best_offer(best, bpgd_pm(s), s->error);
            goto done;
        }
        int guess = -1;
        int favor = bpgd_select(s, depth, &guess);
        if (favor == -1 || guess == -1) break;
        if (depth < T) {
            int dir = (id >> (T - 1 - depth)) & 1;
            if (dir) { on_side = 1; s->A = 0; s->A_sum = -10; favor = 1 - favor; }
        } else if (depth == T) {
            memcpy(bvn, s->vn_mask, s->n); memcpy(bcn, s->cn_mask, s->m);
            memcpy(bdeg, s->cn_deg, sizeof(int) * s->m);
            bvar = guess; bval = 1 - favor; saved = 1;
        }
        if (bpgd_set(s, guess, favor) == -1) break;
        if (bpgd_peel(s) == -1) break;
    }
    if (!saved) goto done;
    bpgd_set_masks(s, bvn, bcn, bdeg);            /* :492-496 */
    if (bpgd_set(s, bvar, bval) == -1) goto done;
    if (bpgd_peel(s) == -1) goto done;
    depth = T + 1;
    for (int i = 0; i < steps; i++) {
        if (bpgd_bp(s)) {
            double pm = bpgd_pm(s);
            if (pm > own_pm) goto done;
            if (s->st) s->st->paths_converged++;
            best_offer(best, pm, s->error);
            goto done;
        }
        int guess = -1;
        int favor = bpgd_select(s, depth, &guess);
        if (favor == -1 || guess == -1) goto done;
        if (bpgd_set(s, guess, favor) == -1) goto done;
        if (bpgd_peel(s) == -1) goto done;
        depth++;
    }
done:
    free(bvn); free(bcn); free(bdeg);
}

/* BPGD_side_thread::do_work, bpgd.cpp:527-570 */
static void run_side(bpgd *s, const snapshot *q, const orc_gdg_params *P, best_t *best) {
    s->A = 0; s->A_sum = -10; s->C = 30; s->D = 3;      /* bpgd.hpp:111 */
    /* masks copied over the freshly reset state; messages keep reset's init */
    memcpy(s->vn_mask, q->vn_mask, s->n); memcpy(s->cn_mask, q->cn_mask, s->m);
    memcpy(s->cn_deg, q->cn_deg, sizeof(int) * s->m);
    memcpy(s->error, s->vn_mask, s->n);
    int depth = q->depth;
    if (bpgd_set(s, q->vn, q->value) == -1) return;
    if (bpgd_peel(s) == -1) return;
    for (int i = 0; i < P->max_side_branch_step; i++) {
        if (bpgd_bp(s)) {
            if (s->st) s->st->paths_converged++;
            best_offer(best, bpgd_pm(s), s->error);
            return;
        }
        int guess = -1;
        int favor = bpgd_select(s, depth, &guess);
        if (favor == -1 || guess == -1) return;
        if (bpgd_set(s, guess, favor) == -1) return;
        if (bpgd_peel(s) == -1) return;
        depth++;
    }
}

/* gdg_multi_thread, pyx:238-251 + bpgd.cpp:591-688.  err_out [new_n]. */
static int gdg_multi(graph *sub, const int *cols, const double *llr, const int8_t *synd,
                     const orc_gdg_params *P, int8_t *err_out, double *min_pm, orc_stats *st) {
    int m = sub->m, nn = sub->n;
    int T = P->max_tree_depth, S = P->max_side_depth;
    int n_tree = (1 << T) - 1, n_side = S - T; if (n_side < 0) n_side = 0;
    best_t best; best.pm = MAX_PM; best.err = err_out; best.n = nn;
    memset(err_out, 0, nn);
    snapshot *side = (snapshot *)calloc(n_side + 1, sizeof(snapshot));
    for (int j = 0; j < n_side; j++) {
        side[j].vn_mask = (int8_t *)malloc(nn + 1); side[j].cn_mask = (int8_t *)malloc(m + 1);
        side[j].cn_deg = (int *)malloc(sizeof(int) * (m + 1));
    }
    bpgd *s = bpgd_new(sub, P->max_iter_per_step, P->low_error_mode, P->gdg_factor, st);
    int ok = bpgd_reset(s, cols, llr, synd);
    if (ok == 0) {
        int conv = 0;
        if (st) st->paths_run++;
        run_main(s, P, side, &best, &conv);
        if (conv && st) st->paths_converged++;
        /* keep main's final error for the fallback (:678-683) */
        int8_t *main_err = (int8_t *)malloc(nn + 1);
        memcpy(main_err, s->error, nn);
        for (int id = 1; id <= n_tree; id++) {
            bpgd *t = bpgd_new(sub, P->max_iter_per_step, P->low_error_mode, P->gdg_factor, st);
            if (bpgd_reset(t, cols, llr, synd) == 0) { if (st) st->paths_run++; run_tree(t, id, P, &best); }
            bpgd_free(t);
        }
        for (int j = 0; j < n_side; j++) {
            if (!side[j].valid) continue;
            bpgd *t = bpgd_new(sub, P->max_iter_per_step, P->low_error_mode, P->gdg_factor, st);
            if (bpgd_reset(t, cols, llr, synd) == 0) { if (st) st->paths_run++; run_side(t, &side[j], P, &best); }
            bpgd_free(t);
        }
        if (!conv && best.pm > MAX_PM - 1.0) memcpy(err_out, main_err, nn);
        free(main_err);
    }
    bpgd_free(s);
    for (int j = 0; j < n_side; j++) { free(side[j].vn_mask); free(side[j].cn_mask); free(side[j].cn_deg); }
    free(side);
    *min_pm = best.pm;
    return best.pm < 9999.0;
}

/* ------------------------------------- single-thread GDG (pyx:254-442) -- */
typedef struct {
    int max_guess, used; int min_conv_depth;
    int8_t **vn, **cn; int **deg; int *dvn, *dval, *ddepth;
} gstack;

/* bpgdg_decoder.select_vn, pyx:340-442 */
static int st_select(bpgd *s, gstack *G, const orc_gdg_params *P, int side_branch, int depth) {
    s->A = side_branch ? 0 : -3;
    s->A_sum = side_branch ? -10 : -12;
    if (depth == 0) s->A_sum = -16;
    s->C = 30; s->D = 3;
    int guess = -1;
    int favor = bpgd_select(s, depth, &guess);   /* same scan + peel + choice (pyx:357-412) */
    if (favor == -1) return -1;
    if (guess == -1) return -1;                  /* reference: UB */
    int do_guess = 1;
    if (depth > G->min_conv_depth) do_guess = 0;
    if (!side_branch && depth >= P->max_side_depth) do_guess = 0;
    if (side_branch && depth > P->max_tree_depth) do_guess = 0;
    if (do_guess && G->used < G->max_guess) {
        int u = G->used;
        G->dval[u] = 1 - favor; G->dvn[u] = guess; G->ddepth[u] = depth + 1;
        memcpy(G->vn[u], s->vn_mask, s->n); memcpy(G->cn[u], s->cn_mask, s->m);
        memcpy(G->deg[u], s->cn_deg, sizeof(int) * s->m);
        G->used = u + 1;
    }
    if (bpgd_set(s, guess, favor) == -1) return -1;
    if (bpgd_peel(s) == -1) return -1;
    return 0;
}

/* bpgdg_decoder.gdg, pyx:254-338. Returns converge; have_err=0 if reset failed. */
static int gdg_single(graph *sub, const int *cols, const double *llr, const int8_t *synd,
                      const orc_gdg_params *P, int8_t *err_out, double *min_pm_out, int *have_err, orc_stats *st) {
    int nn = sub->n, m = sub->m, converge = 0;
    bpgd *s = bpgd_new(sub, P->max_iter_per_step, P->low_error_mode, P->gdg_factor, st);
    *have_err = 0;
    if (bpgd_reset(s, cols, llr, synd) == -1) { bpgd_free(s); *min_pm_out = MAX_PM; return 0; }
    *have_err = 1;
    if (st) st->paths_run++;
    gstack G;
    G.max_guess = ((1 << P->max_tree_depth) - 1) * 2 + P->max_side_depth - P->max_tree_depth;
    if (G.max_guess < 0) G.max_guess = 0;
    G.used = 0; G.min_conv_depth = P->max_step;
    G.vn = (int8_t **)malloc(sizeof(void *) * (G.max_guess + 1));
    G.cn = (int8_t **)malloc(sizeof(void *) * (G.max_guess + 1));
    G.deg = (int **)malloc(sizeof(void *) * (G.max_guess + 1));
    G.dvn = (int *)malloc(sizeof(int) * (G.max_guess + 1));
    G.dval = (int *)malloc(sizeof(int) * (G.max_guess + 1));
    G.ddepth = (int *)malloc(sizeof(int) * (G.max_guess + 1));
    for (int i = 0; i < G.max_guess; i++) {
        G.vn[i] = (int8_t *)malloc(nn + 1); G.cn[i] = (int8_t *)malloc(m + 1);
        G.deg[i] = (int *)malloc(sizeof(int) * (m + 1));
    }
    double min_pm = MAX_PM;
    for (int depth = 0; depth < P->max_step; depth++) {
        if (bpgd_bp(s)) {
            converge = 1; G.min_conv_depth = depth;
            min_pm = bpgd_pm(s); memcpy(err_out, s->error, nn);
            if (st) st->paths_converged++;
            break;
        }
        if (st_select(s, &G, P, 0, depth) == -1) break;
    }
    if (!converge) memcpy(err_out, s->error, nn);
    for (int i = 0; i < G.used; i++) {
        int depth = G.ddepth[i];
        if (depth > G.min_conv_depth) continue;
        bpgd_set_masks(s, G.vn[i], G.cn[i], G.deg[i]);
        if (bpgd_set(s, G.dvn[i], G.dval[i]) == -1) continue;
        if (bpgd_peel(s) == -1) continue;
        if (st) st->paths_run++;
        for (int j = 0; j < P->max_side_branch_step; j++) {
            depth = G.ddepth[i] + j;
            if (bpgd_bp(s)) {
                converge = 1;
                if (st) st->paths_converged++;
                double pm = bpgd_pm(s);
                if (pm < min_pm) {
                    if (depth < G.min_conv_depth) G.min_conv_depth = depth;
                    memcpy(err_out, s->error, nn); min_pm = pm;
                }
                break;
            }
            if (depth > G.min_conv_depth + 2) break;
            if (st_select(s, &G, P, 1, depth) == -1) break;
        }
    }
    for (int i = 0; i < G.max_guess; i++) { free(G.vn[i]); free(G.cn[i]); free(G.deg[i]); }
    free(G.vn); free(G.cn); free(G.deg); free(G.dvn); free(G.dval); free(G.ddepth);
    bpgd_free(s);
    *min_pm_out = min_pm;
    return converge;
}

static int eff_new_n(int m, int n, int new_n) {
    if (new_n <= 0) { new_n = 2 * m; }          /* pyx:187-190 */
    return new_n < n ? new_n : n;
}

/* bpgdg_decoder.decode, pyx:221-236 (g: the window graph, built once per decoder / batch) */
static int bpgdg_decode_g(graph *g, int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                          const orc_gdg_params *P, int8_t *dec, double *min_pm_out, orc_stats *st) {
    double *hist = (double *)calloc(4 * n + 4, sizeof(double));
    int iters = 0, conv;
    int64_t ei = 0;
    conv = bp_pre(g, llr, synd, P->max_iter, P->ms_scaling_factor, hist, dec, &iters, &ei);
    if (st) { st->pre_iters += iters; st->edge_iters += ei; st->stage = 0; }
    if (min_pm_out) *min_pm_out = MAX_PM;
    if (!conv) {
        if (st) st->stage = 1;
        int nn = eff_new_n(m, n, P->new_n);
        double *sum = (double *)malloc(sizeof(double) * (n + 1));
        int *cols = (int *)malloc(sizeof(int) * (n + 1));
        for (int v = 0; v < n; v++) sum[v] = hist[4 * v] + hist[4 * v + 1] + hist[4 * v + 2] + hist[4 * v + 3];
        orc_index_sort(sum, n, cols);
        graph *sub = graph_build(m, nn, cp, cr, cols);
        int8_t *err = (int8_t *)calloc(nn + 1, 1);
        double pm = MAX_PM;
        if (P->multi_thread) {
            conv = gdg_multi(sub, cols, llr, synd, P, err, &pm, st);
            for (int v = 0; v < nn; v++) dec[cols[v]] = err[v];
            for (int v = nn; v < n; v++) dec[cols[v]] = 0;
        } else {
            int have = 0;
            for (int v = nn; v < n; v++) dec[cols[v]] = 0;
            conv = gdg_single(sub, cols, llr, synd, P, err, &pm, &have, st);
            if (have) for (int v = 0; v < nn; v++) dec[cols[v]] = err[v];
        }
        if (min_pm_out) *min_pm_out = pm;
        free(err); graph_free(sub); free(sum); free(cols);
    }
    free(hist);
    return conv;
}

int orc_bpgdg_decode(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                     const orc_gdg_params *P, int8_t *dec, double *min_pm_out, orc_stats *st) {
    graph *g = graph_build(m, n, cp, cr, NULL);
    int r = bpgdg_decode_g(g, m, n, cp, cr, llr, synd, P, dec, min_pm_out, st);
    graph_free(g);
    return r;
}

/* bpgd_decoder.decode / gd, pyx:501-560 */
int orc_bpgd_decode(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                    const orc_gdg_params *P, int8_t *dec, double *min_pm_out, orc_stats *st) {
    memset(dec, 0, n);
    if (min_pm_out) *min_pm_out = MAX_PM;
    if (P->max_iter <= -1) return 0;             /* pyx:506 */
    graph *g = graph_build(m, n, cp, cr, NULL);
    double *hist = (double *)calloc(4 * n + 4, sizeof(double));
    int iters = 0; int64_t ei = 0;
    int conv = bp_pre(g, llr, synd, P->max_iter, P->ms_scaling_factor, hist, dec, &iters, &ei);
    if (st) { st->pre_iters += iters; st->edge_iters += ei; st->stage = 0; }
    if (!conv) {
        if (st) st->stage = 1;
        int nn = eff_new_n(m, n, P->new_n);
        double *sum = (double *)malloc(sizeof(double) * (n + 1));
        int *cols = (int *)malloc(sizeof(int) * (n + 1));
        for (int v = 0; v < n; v++) sum[v] = hist[4 * v] + hist[4 * v + 1] + hist[4 * v + 2] + hist[4 * v + 3];
        orc_index_sort(sum, n, cols);
        for (int v = nn; v < n; v++) dec[cols[v]] = 0;
        graph *sub = graph_build(m, nn, cp, cr, cols);
        bpgd *s = bpgd_new(sub, P->max_iter_per_step, 0, P->gdg_factor, st);
        if (bpgd_reset(s, cols, llr, synd) != -1) {
            if (st) st->paths_run++;
            for (int depth = 0; depth < P->max_step; depth++) {
                if (bpgd_bp(s)) { conv = 1; if (min_pm_out) *min_pm_out = bpgd_pm(s); break; }
                if (bpgd_decimate_reliable(s) == -1) break;
            }
            for (int v = 0; v < nn; v++) dec[cols[v]] = s->error[v];
        }
        bpgd_free(s); graph_free(sub); free(sum); free(cols);
    }
    free(hist); graph_free(g);
    return conv;
}

/* ------------------------------------------------------------- GF(2) OSD -- */
typedef struct {
    int m, W, R;              /* rows, words per row-vector, words per comb */
    int rank;
    uint64_t *vec;            /* [rank][W] reduced basis                    */
    uint64_t *comb;           /* [rank][R] combination of pivot columns     */
    int *prow;                /* pivot row of basis i                       */
    int *pcol;                /* original column of basis i                 */
    int cap;
} gf2_basis;

static gf2_basis *basis_new(int m, int cap) {
    gf2_basis *b = (gf2_basis *)calloc(1, sizeof(gf2_basis));
    b->m = m; b->W = (m + 63) / 64; b->R = (cap + 63) / 64; b->cap = cap;
    b->vec = (uint64_t *)calloc((size_t)(cap + 1) * b->W, 8);
    b->comb = (uint64_t *)calloc((size_t)(cap + 1) * b->R, 8);
    b->prow = (int *)calloc(cap + 1, sizeof(int)); b->pcol = (int *)calloc(cap + 1, sizeof(int));
    return b;
}
static void basis_free(gf2_basis *b) { free(b->vec); free(b->comb); free(b->prow); free(b->pcol); free(b); }

/* reduce v (W words) against the basis in insertion order; accumulate the
 * used basis combination into cb (R words). */
static void basis_reduce(const gf2_basis *b, uint64_t *v, uint64_t *cb) {
    for (int i = 0; i < b->rank; i++) {
        int r = b->prow[i];
        if ((v[r >> 6] >> (r & 63)) & 1) {
            for (int w = 0; w < b->W; w++) v[w] ^= b->vec[(size_t)i * b->W + w];
            for (int w = 0; w < b->R; w++) cb[w] ^= b->comb[(size_t)i * b->R + w];
        }
    }
}
/* try to add column (rows list) -> 1 if independent */
static int basis_add(gf2_basis *b, const int *rows, int d, int col, uint64_t *tv, uint64_t *tc) {
    memset(tv, 0, 8 * b->W); memset(tc, 0, 8 * b->R);
    for (int k = 0; k < d; k++) tv[rows[k] >> 6] ^= 1ull << (rows[k] & 63);
    basis_reduce(b, tv, tc);
    int pr = -1;
    for (int w = 0; w < b->W && pr < 0; w++) if (tv[w]) pr = w * 64 + __builtin_ctzll(tv[w]);
    if (pr < 0) return 0;
    int i = b->rank++;
    memcpy(b->vec + (size_t)i * b->W, tv, 8 * b->W);
    tc[i >> 6] ^= 1ull << (i & 63);
    memcpy(b->comb + (size_t)i * b->R, tc, 8 * b->R);
    b->prow[i] = pr; b->pcol[i] = col;
    return 1;
}

int orc_gf2_rank(int m, int n, const int *cp, const int *cr) {
    int cap = m < n ? m : n;
    gf2_basis *b = basis_new(m, cap);
    uint64_t *tv = (uint64_t *)malloc(8 * (b->W + 1)), *tc = (uint64_t *)malloc(8 * (b->R + 1));
    for (int c = 0; c < n && b->rank < cap; c++) basis_add(b, cr + cp[c], cp[c + 1] - cp[c], c, tv, tc);
    int r = b->rank;
    free(tv); free(tc); basis_free(b);
    return r;
}

/* solve H_P x_P = g for the pivot set held in b; x (n chars) gets x_P scattered
 * to original columns (other entries untouched). */
static void basis_solve(const gf2_basis *b, const uint64_t *g, int8_t *x, uint64_t *tv, uint64_t *tc) {
    memcpy(tv, g, 8 * b->W); memset(tc, 0, 8 * b->R);
    basis_reduce(b, tv, tc);
    for (int i = 0; i < b->rank; i++) x[b->pcol[i]] = (int8_t)((tc[i >> 6] >> (i & 63)) & 1);
}

/* osd_window.osd, osd_window.pyx:201-284 (+ mod2sparse_extra.cpp:78-376).
 * cur_vn: -1 undecided / 0 / 1.  Writes osd0 and osdw, returns min_pm. */
static double osd_run_keys(const graph *g, const double *llr, const int8_t *synd, double *key,
                           const orc_osd_params *P, int rank, int nn, int8_t *osd0, int8_t *osdw);
static double osd_run(const graph *g, const double *llr, const int8_t *synd, const int8_t *cur_vn,
                      const double *hist, const orc_osd_params *P, int rank, int nn,
                      int8_t *osd0, int8_t *osdw) {
    int n = g->n;
    double *key = (double *)malloc(sizeof(double) * (n + 1));
    for (int v = 0; v < n; v++) {
        if (cur_vn[v] == 1) key[v] = -1000;
        else if (cur_vn[v] == 0) key[v] = 1000;
        else key[v] = hist[4 * v] + hist[4 * v + 1] + hist[4 * v + 2] + hist[4 * v + 3];
    }
    return osd_run_keys(g, llr, synd, key, P, rank, nn, osd0, osdw);     /* frees key */
}
/* the OSD proper for a given column ranking key (ascending = most likely in error first); also what bp4_osd.osd
 * does (bp4_osd.pyx:261-368) with key = llr_post */
static double osd_run_keys(const graph *g, const double *llr, const int8_t *synd, double *key,
                           const orc_osd_params *P, int rank, int nn, int8_t *osd0, int8_t *osdw) {
    int m = g->m, n = g->n;
    int *order = (int *)malloc(sizeof(int) * (n + 1));
    orc_index_sort(key, n, order);
    gf2_basis *b = basis_new(m, rank);
    uint64_t *tv = (uint64_t *)malloc(8 * (b->W + 1)), *tc = (uint64_t *)malloc(8 * (b->R + 1));
    uint64_t *sv = (uint64_t *)calloc(b->W + 1, 8), *gv = (uint64_t *)calloc(b->W + 1, 8);
    int8_t *is_piv = (int8_t *)calloc(n + 1, 1);
    for (int i = 0; i < n && b->rank < rank; i++) {
        int c = order[i];
        if (basis_add(b, g->cr + g->cp[c], g->cp[c + 1] - g->cp[c], c, tv, tc)) is_piv[c] = 1;
    }
    for (int c = 0; c < m; c++) if (synd[c]) sv[c >> 6] |= 1ull << (c & 63);
    memset(osd0, 0, (size_t)(n > 0 ? n : 0));
    basis_solve(b, sv, osd0, tv, tc);
    double min_pm = 0.0;
    for (int v = 0; v < n; v++) { if (osd0[v]) min_pm += llr[v]; osdw[v] = osd0[v]; }
    if (P->osd_order > 0 && P->osd_method != 0) {
        int k = nn - rank, cnt = 0;
        int *T = (int *)malloc(sizeof(int) * (k + 1));
        for (int i = 0; i < nn && cnt < k; i++) if (!is_piv[order[i]]) T[cnt++] = order[i];
        int w = P->osd_order;
        long ncand = (P->osd_method == 2) ? (long)k + (long)w * (w - 1) / 2 : (1L << w);
        int8_t *y = (int8_t *)malloc(n + 1);
        int8_t *x = (int8_t *)calloc(k + 1, 1);
        long pair_i = 0, pair_j = 1;
        for (long l = 0; l < ncand; l++) {
            memset(x, 0, k);
            if (P->osd_method == 2) {
                if (l < k) x[l] = 1;
                else {
                    x[pair_i] = 1; x[pair_j] = 1;
                    if (++pair_j >= w) { pair_i++; pair_j = pair_i + 1; }
                }
            } else {
                long d = l;                      /* decimal_to_binary_reverse */
                for (int i = 0; i < k; i++) { x[i] = (int8_t)(d & 1); d >>= 1; if (d == 0) break; }
            }
            memcpy(gv, sv, 8 * b->W);
            for (int i = 0; i < k; i++) if (x[i]) {
                int c = T[i];
                for (int e = g->cp[c]; e < g->cp[c + 1]; e++) gv[g->cr[e] >> 6] ^= 1ull << (g->cr[e] & 63);
            }
            memset(y, 0, n);
            basis_solve(b, gv, y, tv, tc);
            for (int i = 0; i < k; i++) y[T[i]] = x[i];
            double pm = 0.0;
            for (int v = 0; v < n; v++) if (y[v]) pm += llr[v];
            if (pm < min_pm) { min_pm = pm; memcpy(osdw, y, n); }
        }
        free(T); free(y); free(x);
    }
    free(key); free(order); free(tv); free(tc); free(sv); free(gv); free(is_piv); basis_free(b);
    return min_pm;
}

/* osd_window.vn_set_value, osd_window.pyx:340-368 (skips inactive checks) */
static int ow_set(const graph *g, int8_t *cur_vn, int8_t *cur_cn, int *deg, int8_t *dec, int vn, int value) {
    if (cur_vn[vn] != -1) return (cur_vn[vn] == value) ? 0 : -1;
    cur_vn[vn] = (int8_t)value; dec[vn] = (int8_t)value;
    for (int e = g->cp[vn]; e < g->cp[vn + 1]; e++) {
        int cn = g->cr[e];
        if (cur_cn[cn] == -1) continue;
        int d = deg[cn] - 1;
        if (value) cur_cn[cn] = 1 - cur_cn[cn];
        if (d == 0) { if (cur_cn[cn] != 0) return -1; cur_cn[cn] = -1; }
        deg[cn] = d;
    }
    return 0;
}
/* osd_window.peel, osd_window.pyx:306-338 */
static int ow_peel(const graph *g, int8_t *cur_vn, int8_t *cur_cn, int *deg, int8_t *dec) {
    for (;;) {
        int clean = 1;
        for (int cn = 0; cn < g->m; cn++) {
            if (cur_cn[cn] == -1) continue;
            if (deg[cn] >= 2) continue;
            clean = 0;
            int vn = -1;
            for (int p = g->rp[cn]; p < g->rp[cn + 1]; p++)
                if (cur_vn[g->rc[p]] == -1) { vn = g->rc[p]; break; }
            if (vn == -1) return -1;             /* reference: out-of-bounds access */
            if (ow_set(g, cur_vn, cur_cn, deg, dec, vn, cur_cn[cn]) == -1) return -1;
        }
        if (clean) return 0;
    }
}

int orc_osd_window_decode(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                          const orc_osd_params *P, int rank, int8_t *dec, int8_t *bp_dec, int8_t *osd0_dec,
                          int8_t *osdw_dec, double *hist_out, double *min_pm_out, int *bp_iter_out, orc_stats *st) {
    graph *g = graph_build(m, n, cp, cr, NULL);
    double *hist = (double *)calloc(4 * n + 4, sizeof(double));
    double *b2c = (double *)malloc(sizeof(double) * (g->nnz + 1));
    double *c2b = (double *)malloc(sizeof(double) * (g->nnz + 1));
    int8_t *cur_vn = (int8_t *)malloc(n + 1), *cur_cn = (int8_t *)malloc(m + 1), *ts = (int8_t *)malloc(m + 1);
    int *deg = (int *)malloc(sizeof(int) * (m + 1));
    int8_t *bpd = (int8_t *)calloc(n + 1, 1);
    int8_t *o0 = (int8_t *)calloc(n + 1, 1), *ow = (int8_t *)calloc(n + 1, 1);
    int nn = eff_new_n(m, n, P->new_n);
    int conv = 0, bp_iter = 0, used_osd = 0;
    double min_pm = 0.0;
    /* reset :288-303 */
    for (int c = 0; c < m; c++) { deg[c] = g->rp[c + 1] - g->rp[c]; cur_cn[c] = synd[c]; }
    memset(cur_vn, -1, n);
    /* bp_init :370-379 + pre-BP :381-485 */
    for (int v = 0; v < n; v++) for (int e = g->cp[v]; e < g->cp[v + 1]; e++) b2c[g->c2r[e]] = llr[v];
    if (st) st->stage = 0;
    for (int it = 0; it < P->pre_max_iter; it++) {
        bp_iter++;
        int64_t ed = g_bp_method ? ps_iteration(g, synd, llr, b2c, c2b, hist, it % 4, bpd)       /* nothing is masked before the shortening */
                                 : ms_iteration(g, cur_vn, cur_cn, cur_cn, llr, P->ms_scaling_factor, b2c, c2b, hist, it % 4, bpd);
        if (st) { st->edge_iters += ed; st->pre_iters++; }
        if (synd_match(g, bpd, synd, ts)) { conv = 1; break; }
    }
    if (!conv) {
        if (st) st->stage = 1;
        double *sum = (double *)malloc(sizeof(double) * (n + 1));
        int *cols = (int *)malloc(sizeof(int) * (n + 1));
        for (int v = 0; v < n; v++) sum[v] = hist[4 * v] + hist[4 * v + 1] + hist[4 * v + 2] + hist[4 * v + 3];
        orc_index_sort(sum, n, cols);
        int failed = 0;
        for (int v = nn; v < n; v++)
            if (ow_set(g, cur_vn, cur_cn, deg, bpd, cols[v], 0) == -1) { failed = 1; break; }   /* :178-181 */
        if (!failed) {
            for (int v = nn; v < n; v++) bpd[cols[v]] = 0;
            if (ow_peel(g, cur_vn, cur_cn, deg, bpd) == -1) failed = 1;                          /* :184-186 */
        }
        if (!failed) {
            for (int v = 0; v < n; v++) {                                                        /* bp_init */
                if (cur_vn[v] != -1) continue;
                for (int e = g->cp[v]; e < g->cp[v + 1]; e++) b2c[g->c2r[e]] = llr[v];
            }
            if (st) st->bp_calls++;
            for (int it = 0; it < P->post_max_iter; it++) {
                bp_iter++;
                int64_t ed = ms_iteration(g, cur_vn, cur_cn, cur_cn, llr, P->ms_scaling_factor, b2c, c2b, hist, it % 4, bpd);
                if (st) st->edge_iters += ed;
                if (synd_match(g, bpd, synd, ts)) { conv = 1; break; }
            }
            if (!conv && P->osd_order > -1) {
                if (st) st->stage = 2;
                min_pm = osd_run(g, llr, synd, cur_vn, hist, P, rank, nn, o0, ow);
                used_osd = 1;
            }
        }
        free(sum); free(cols);
    }
    if (conv) for (int v = 0; v < n; v++) if (bpd[v]) min_pm += llr[v];       /* :168-170, :190-191 */
    memcpy(dec, used_osd ? ow : bpd, n);
    if (bp_dec) memcpy(bp_dec, bpd, n);
    if (osd0_dec) memcpy(osd0_dec, o0, n);
    if (osdw_dec) memcpy(osdw_dec, ow, n);
    if (hist_out) memcpy(hist_out, hist, sizeof(double) * 4 * n);
    if (min_pm_out) *min_pm_out = min_pm;
    if (bp_iter_out) *bp_iter_out = bp_iter;
    free(hist); free(b2c); free(c2b); free(cur_vn); free(cur_cn); free(ts); free(deg); free(bpd); free(o0); free(ow);
    graph_free(g);
    return conv;
}

/* ----------------------------------------------------------------- batch -- */
static void stats_add(orc_stats *sum, const orc_stats *s) {
    sum->pre_iters += s->pre_iters; sum->edge_iters += s->edge_iters; sum->bp_calls += s->bp_calls;
    sum->paths_run += s->paths_run; sum->paths_converged += s->paths_converged;
}
void orc_bpgdg_decode_batch(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                            int64_t B, const orc_gdg_params *P, int8_t *dec, int8_t *conv, double *pm, orc_stats *sum) {
    graph *g = graph_build(m, n, cp, cr, NULL);
    for (int64_t b = 0; b < B; b++) {
        orc_stats s; memset(&s, 0, sizeof s);
        double p = 0;
        int c = bpgdg_decode_g(g, m, n, cp, cr, llr, synd + b * m, P, dec + b * n, &p, &s);
        if (conv) conv[b] = (int8_t)c;
        if (pm) pm[b] = p;
        if (sum) stats_add(sum, &s);
    }
    graph_free(g);
}
void orc_osd_window_decode_batch(int m, int n, const int *cp, const int *cr, const double *llr, const int8_t *synd,
                                 int64_t B, const orc_osd_params *P, int rank, int8_t *dec, int8_t *conv, double *pm,
                                 orc_stats *sum) {
    for (int64_t b = 0; b < B; b++) {
        orc_stats s; memset(&s, 0, sizeof s);
        double p = 0;
        int c = orc_osd_window_decode(m, n, cp, cr, llr, synd + b * m, P, rank, dec + b * n, NULL, NULL, NULL, NULL,
                                      &p, NULL, &s);
        if (conv) conv[b] = (int8_t)c;
        if (pm) pm[b] = p;
        if (sum) stats_add(sum, &s);
    }
}

/* ------------------------------------------------------------ bp4_osd -- */
/* bpgd.cpp:399-416 */
static double b4_log1pexp(double x) {
    if (x > -log(2.220446049250313e-16)) return x + log1p(exp(-x));
    return log1p(exp(x));
}
static double b4_logaddexp(double x, double y) {
    const double tmp = x - y;
    if (x == y) return x + 0.69314718055994530942;      /* M_LN2 */
    if (tmp > 0) return x + b4_log1pexp(-tmp);
    else if (tmp <= 0) return y + b4_log1pexp(tmp);
    return tmp;
}
/* bp4_osd.cn_update_all (bp4_osd.pyx:483-529): unmasked normalised min-sum on one of the two graphs */
static void b4_cn_update(const graph *g, const int8_t *synd, double alpha, double *b2c, double *c2b) {
    for (int c = 0; c < g->m; c++) {
        double m1 = BIG, m2 = BIG; int arg = -1; int par = (synd[c] == 1);
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) {
            double b = b2c[p];
            if (b > CLIP) b = CLIP; else if (b < -CLIP) b = -CLIP;
            b2c[p] = b;
            double a = fabs(b);
            if (a < m1) { m2 = m1; m1 = a; arg = p; } else if (a < m2) m2 = a;
            if (b <= 0) par ^= 1;
        }
        for (int p = g->rp[c]; p < g->rp[c + 1]; p++) {
            double mag = (p == arg) ? m2 : m1;
            int sg = par ^ (b2c[p] <= 0);
            c2b[p] = mag * ((sg ? -1.0 : 1.0) * alpha);
        }
    }
}
/* one BP4 run (bp4_decode_llr, pyx:444-481) from initialised messages; fix_vn >= 0: that qubit is decided (pyx:389-423)
 * and skipped by the variable pass, its messages stay at their initial value; seeds = current_cn (syndrome with the
 * decided qubit's contribution toggled).  Returns converge. */
static int b4_run(const graph *gx, const graph *gz, const double *llrx, const double *llry, const double *llrz,
                  const int8_t *synd_x, const int8_t *synd_z, const int8_t *seed_x, const int8_t *seed_z, int fix_vn,
                  int max_iter, double alpha, double *b2cx, double *c2bx, double *b2cz, double *c2bz,
                  int8_t *bx, int8_t *bz, double *lpr, int8_t *tsx, int8_t *tsz, int *iters) {
    int n = gx->n, it = 0, conv = 0;
    for (int iter = 0; iter < max_iter; iter++) {
        it++;
        b4_cn_update(gx, seed_x, alpha, b2cx, c2bx);
        b4_cn_update(gz, seed_z, alpha, b2cz, c2bz);
        for (int v = 0; v < n; v++) {
            if (v == fix_vn) continue;
            double llrx_hx = 0.0, llrz_hz = 0.0;
            for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) llrx_hx += c2bz[gz->c2r[e]];
            for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) llrz_hz += c2bx[gx->c2r[e]];
            double llry_all = llrx_hx + llrz_hz + llry[v];
            llrx_hx = llrx_hx + llrx[v];
            llrz_hz = llrz_hz + llrz[v];
            lpr[3 * v] = llrx_hx; lpr[3 * v + 1] = llry_all; lpr[3 * v + 2] = llrz_hz;
            int idx;
            if (0 < llrx_hx && 0 < llry_all && 0 < llrz_hz) idx = 0;
            else if (llrx_hx < llry_all && llrx_hx < llrz_hz) idx = 1;
            else if (llry_all > llrz_hz) idx = 2;
            else idx = 3;
            bx[v] = (int8_t)(idx % 2); bz[v] = (int8_t)(idx / 2);
            double num_hx = b4_log1pexp(-1. * llrx_hx);
            for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) {
                int p = gx->c2r[e];
                double msg = c2bx[p];
                b2cx[p] = num_hx - b4_logaddexp(-1. * (llrz_hz - msg), -1. * (llry_all - msg));
            }
            double num_hz = b4_log1pexp(-1. * llrz_hz);
            for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) {
                int p = gz->c2r[e];
                double msg = c2bz[p];
                b2cz[p] = num_hz - b4_logaddexp(-1. * (llrx_hx - msg), -1. * (llry_all - msg));
            }
        }
        if (synd_match(gx, bz, synd_x, tsx) && synd_match(gz, bx, synd_z, tsz)) { conv = 1; break; }
    }
    *iters = it;
    return conv;
}
static void b4_init(const graph *gx, const graph *gz, const double *llrx, const double *llry, const double *llrz, double *b2cx, double *b2cz) {
    for (int v = 0; v < gx->n; v++) {                               /* bp_init, pyx:425-442 */
        double msg_x = b4_log1pexp(-1. * llrx[v]) - b4_logaddexp(-1. * llry[v], -1. * llrz[v]);
        double msg_z = b4_log1pexp(-1. * llrz[v]) - b4_logaddexp(-1. * llry[v], -1. * llrz[v]);   /* sic (pyx:438) */
        for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) b2cx[gx->c2r[e]] = msg_x;
        for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) b2cz[gz->c2r[e]] = msg_z;
    }
}
/* bp4_osd.camel_decode (pyx:223-248): four BP runs with the last qubit pinned to I / X / Z / Y, keep the converged run
 * with the smallest path metric (cal_pm, pyx:250-259; strict <, value order 0..3).  dec [2n] (zeros if no run converged;
 * the reference leaves the previous call's buffer), lpr / bp_iter of the last run.  Returns converge. */
int orc_bp4_camel_decode(int mx, int mz, int n, const int *cpx, const int *crx, const int *cpz, const int *crz,
                         const double *llrx, const double *llry, const double *llrz,
                         const int8_t *synd_x, const int8_t *synd_z, int max_iter, double alpha,
                         int8_t *dec, double *lpr, double *min_pm_out, int *bp_iter_out) {
    graph *gx = graph_build(mx, n, cpx, crx, NULL), *gz = graph_build(mz, n, cpz, crz, NULL);
    double *b2cx = (double *)calloc(gx->nnz + 1, 8), *c2bx = (double *)calloc(gx->nnz + 1, 8);
    double *b2cz = (double *)calloc(gz->nnz + 1, 8), *c2bz = (double *)calloc(gz->nnz + 1, 8);
    int8_t *bp = (int8_t *)calloc(2 * (size_t)n + 1, 1), *bx = bp, *bz = bp + n;
    int8_t *tsx = (int8_t *)malloc(mx + 1), *tsz = (int8_t *)malloc(mz + 1);
    int8_t *cx = (int8_t *)malloc(mx + 1), *cz = (int8_t *)malloc(mz + 1);
    double min_pm = 10000.0; int it = 0;
    memset(dec, 0, 2 * (size_t)n); memset(lpr, 0, sizeof(double) * 3 * (size_t)n);
    for (int value = 0; value < 4; value++) {
        memset(bp, 0, 2 * (size_t)n);
        memcpy(cx, synd_x, mx); memcpy(cz, synd_z, mz);
        b4_init(gx, gz, llrx, llry, llrz, b2cx, b2cz);
        int vn = n - 1, x = value % 2, z = value / 2;                /* vn_set_value, pyx:389-423 */
        bx[vn] = (int8_t)x; bz[vn] = (int8_t)z;
        if (z) for (int e = gx->cp[vn]; e < gx->cp[vn + 1]; e++) cx[gx->cr[e]] = 1 - cx[gx->cr[e]];
        if (x) for (int e = gz->cp[vn]; e < gz->cp[vn + 1]; e++) cz[gz->cr[e]] = 1 - cz[gz->cr[e]];
        if (b4_run(gx, gz, llrx, llry, llrz, synd_x, synd_z, cx, cz, vn, max_iter, alpha, b2cx, c2bx, b2cz, c2bz, bx, bz, lpr, tsx, tsz, &it)) {
            double pm = 0.0;                                            /* cal_pm */
            for (int v = 0; v < n; v++) {
                if (bx[v] && bz[v]) pm += llry[v];
                else if (bx[v]) pm += llrx[v];
                else if (bz[v]) pm += llrz[v];
            }
            if (pm < min_pm) { min_pm = pm; memcpy(dec, bp, 2 * (size_t)n); }
        }
    }
    if (min_pm_out) *min_pm_out = min_pm;
    if (bp_iter_out) *bp_iter_out = it;
    free(b2cx); free(c2bx); free(b2cz); free(c2bz); free(bp); free(tsx); free(tsz); free(cx); free(cz);
    graph_free(gx); graph_free(gz);
    return min_pm < 9999.0;
}

/* bp4_osd.decode (bp4_osd.pyx:197-221): quaternary BP over (Hx, Hz) + one OSD per basis.
 * llr*: channel LLRs log((1-px-py-pz)/p*), prior_llr_x/z: log((1-(px+py))/(px+py)) resp. (pz+py) (pyx:123-133),
 * computed by the caller with libm.  dec / bp_dec / osd0: [2n] = x part then z part.  lpr: [n][3] (x, y, z).
 * Returns converge.  camel_decode is not restated. */
int orc_bp4_osd_decode(int mx, int mz, int n, const int *cpx, const int *crx, const int *cpz, const int *crz,
                       const double *llrx, const double *llry, const double *llrz,
                       const double *prior_llr_x, const double *prior_llr_z,
                       const int8_t *synd_x, const int8_t *synd_z, int max_iter, double alpha,
                       int osd_method, int osd_order, int rank_x, int rank_z,
                       int8_t *dec, int8_t *bp_dec, int8_t *osd0, double *lpr, int *bp_iter_out) {
    graph *gx = graph_build(mx, n, cpx, crx, NULL), *gz = graph_build(mz, n, cpz, crz, NULL);
    double *b2cx = (double *)calloc(gx->nnz + 1, 8), *c2bx = (double *)calloc(gx->nnz + 1, 8);
    double *b2cz = (double *)calloc(gz->nnz + 1, 8), *c2bz = (double *)calloc(gz->nnz + 1, 8);
    int8_t *bx = bp_dec, *bz = bp_dec + n;
    int8_t *tsx = (int8_t *)malloc(mx + 1), *tsz = (int8_t *)malloc(mz + 1);
    memset(bp_dec, 0, 2 * (size_t)n); memset(osd0, 0, 2 * (size_t)n);
    for (int v = 0; v < n; v++) {                                   /* bp_init, pyx:425-442 */
        double msg_x = b4_log1pexp(-1. * llrx[v]) - b4_logaddexp(-1. * llry[v], -1. * llrz[v]);
        double msg_z = b4_log1pexp(-1. * llrz[v]) - b4_logaddexp(-1. * llry[v], -1. * llrz[v]);   /* sic (pyx:438) */
        for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) b2cx[gx->c2r[e]] = msg_x;
        for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) b2cz[gz->c2r[e]] = msg_z;
    }
    int conv = 0, it = 0;
    for (int iter = 0; iter < max_iter; iter++) {
        it++;
        b4_cn_update(gx, synd_x, alpha, b2cx, c2bx);
        b4_cn_update(gz, synd_z, alpha, b2cz, c2bz);
        for (int v = 0; v < n; v++) {                               /* vn_update, pyx:533-591 */
            double llrx_hx = 0.0, llrz_hz = 0.0;
            for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) llrx_hx += c2bz[gz->c2r[e]];
            for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) llrz_hz += c2bx[gx->c2r[e]];
            double llry_all = llrx_hx + llrz_hz + llry[v];
            llrx_hx = llrx_hx + llrx[v];
            llrz_hz = llrz_hz + llrz[v];
            lpr[3 * v] = llrx_hx; lpr[3 * v + 1] = llry_all; lpr[3 * v + 2] = llrz_hz;
            int idx;
            if (0 < llrx_hx && 0 < llry_all && 0 < llrz_hz) idx = 0;
            else if (llrx_hx < llry_all && llrx_hx < llrz_hz) idx = 1;
            else if (llry_all > llrz_hz) idx = 2;
            else idx = 3;
            bx[v] = (int8_t)(idx % 2); bz[v] = (int8_t)(idx / 2);
            double num_hx = b4_log1pexp(-1. * llrx_hx);
            for (int e = gx->cp[v]; e < gx->cp[v + 1]; e++) {
                int p = gx->c2r[e];
                double msg = c2bx[p];
                b2cx[p] = num_hx - b4_logaddexp(-1. * (llrz_hz - msg), -1. * (llry_all - msg));
            }
            double num_hz = b4_log1pexp(-1. * llrz_hz);
            for (int e = gz->cp[v]; e < gz->cp[v + 1]; e++) {
                int p = gz->c2r[e];
                double msg = c2bz[p];
                b2cz[p] = num_hz - b4_logaddexp(-1. * (llrx_hx - msg), -1. * (llry_all - msg));
            }
        }
        if (synd_match(gx, bz, synd_x, tsx) && synd_match(gz, bx, synd_z, tsz)) { conv = 1; break; }
    }
    if (bp_iter_out) *bp_iter_out = it;
    memcpy(dec, bp_dec, 2 * (size_t)n);
    if (conv) memcpy(osd0, bp_dec, 2 * (size_t)n);                  /* pyx:210-212 */
    else if (osd_order > -1) {
        orc_osd_params P; memset(&P, 0, sizeof(P));
        P.osd_method = osd_method; P.osd_order = (osd_method == 0) ? 0 : osd_order;
        double *key = (double *)malloc(sizeof(double) * (n + 1));
        /* osd('x'): Hx, synd_x -> z part (pyx:264-280) */
        for (int v = 0; v < n; v++) key[v] = b4_log1pexp(-1. * lpr[3 * v]) - b4_logaddexp(-1. * lpr[3 * v + 1], -1. * lpr[3 * v + 2]);
        osd_run_keys(gx, prior_llr_x, synd_x, key, &P, rank_x, n, osd0 + n, dec + n);
        key = (double *)malloc(sizeof(double) * (n + 1));
        /* osd('z'): Hz, synd_z -> x part (pyx:281-297); k = n - rank_x there (pyx:284) */
        for (int v = 0; v < n; v++) key[v] = b4_log1pexp(-1. * lpr[3 * v + 2]) - b4_logaddexp(-1. * lpr[3 * v + 1], -1. * lpr[3 * v]);
        osd_run_keys(gz, prior_llr_z, synd_z, key, &P, rank_z, (rank_x == rank_z) ? n : n, osd0, dec);   /* equal ranks for CSS codes with hx ~ hz */
    }
    free(b2cx); free(c2bx); free(b2cz); free(c2bz); free(tsx); free(tsz);
    graph_free(gx); graph_free(gz);
    return conv;
}
