/*
 * oracle/lsap.c -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * CPU restatement of the rectangular linear-sum-assignment solver that the
 * reference calls at detr/matcher.py:94 (`scipy.optimize.linear_sum_assignment`,
 * imported at detr/matcher.py:8).  The solver itself is a third-party dependency
 * that is NOT under /root/reference: SciPy (unpinned in the reference's
 * setup.py:11-28; 1.18.1 in this image, compiled `_lsap*.so`, no source on disk).
 * Its published algorithm is D. F. Crouse, "On implementing 2D rectangular
 * assignment algorithms", IEEE T-AES 52(4), 2016 (shortest augmenting path with
 * dual variables), and the behavioural contract restated here is SURVEY.md
 * section 8(c):
 *
 *   - arithmetic in float64 whatever the input type,
 *   - tall matrices (more rows than columns) are solved on the transpose,
 *   - the set of unvisited columns is an array filled in REVERSE order and
 *     shrunk by swap-with-last,
 *   - among equal minimum tentative distances an unassigned column wins
 *     (the LAST one in array order), otherwise the FIRST minimum in array order,
 *   - output rows ascending (after un-transposing: sorted by original row).
 *
 * Parity pinning: tests/test_oracle_lsap.py checks this file against SciPy itself
 * (present in the image on both the CPU container and the GPU box) on random,
 * tie-heavy, constant and DETR-shaped matrices, and against the known-answer
 * vectors of SURVEY.md section 8(c).
 *
 * Return codes: 0 ok, 1 invalid entry (NaN or -inf), 2 infeasible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LSAP_OK 0
#define LSAP_INVALID 1
#define LSAP_INFEASIBLE 2

typedef struct {
    double *u, *v, *dist;
    int64_t *pred_row, *col_of_row, *row_of_col, *todo;
    unsigned char *row_seen, *col_seen;
} lsap_ws;

static int ws_alloc(lsap_ws *w, int64_t nr, int64_t nc) {
    memset(w, 0, sizeof(*w));
    w->u = (double *)calloc((size_t)nr, sizeof(double));
    w->v = (double *)calloc((size_t)nc, sizeof(double));
    w->dist = (double *)malloc((size_t)nc * sizeof(double));
    w->pred_row = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->col_of_row = (int64_t *)malloc((size_t)nr * sizeof(int64_t));
    w->row_of_col = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->todo = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->row_seen = (unsigned char *)malloc((size_t)nr);
    w->col_seen = (unsigned char *)malloc((size_t)nc);
    return w->u && w->v && w->dist && w->pred_row && w->col_of_row && w->row_of_col && w->todo &&
           w->row_seen && w->col_seen;
}

static void ws_free(lsap_ws *w) {
    free(w->u); free(w->v); free(w->dist); free(w->pred_row); free(w->col_of_row);
    free(w->row_of_col); free(w->todo); free(w->row_seen); free(w->col_seen);
}

/* One shortest-augmenting-path search started at free row `start`.
 * Returns the sink column (>=0) or -1 when no finite path exists. */
static int64_t grow_path(const double *c, int64_t nr, int64_t nc, int64_t start, lsap_ws *w,
                         double *out_min) {
    double reach = 0.0;
    int64_t n_todo = nc;
    for (int64_t t = 0; t < nc; ++t) w->todo[t] = nc - 1 - t; /* reverse fill */
    memset(w->row_seen, 0, (size_t)nr);
    memset(w->col_seen, 0, (size_t)nc);
    for (int64_t j = 0; j < nc; ++j) w->dist[j] = INFINITY;

    int64_t i = start, sink = -1;
    while (sink < 0) {
        int64_t best_t = -1;
        double best = INFINITY;
        w->row_seen[i] = 1;
        for (int64_t t = 0; t < n_todo; ++t) {
            int64_t j = w->todo[t];
            /* operation order matters for bit-exact duals: ((reach + c) - u) - v */
            double r = reach + c[i * nc + j] - w->u[i] - w->v[j];
            if (r < w->dist[j]) {
                w->dist[j] = r;
                w->pred_row[j] = i;
            }
            if (w->dist[j] < best || (w->dist[j] == best && w->row_of_col[j] < 0)) {
                best = w->dist[j];
                best_t = t;
            }
        }
        reach = best;
        if (reach == INFINITY) return -1;
        int64_t j = w->todo[best_t];
        if (w->row_of_col[j] < 0) sink = j;
        else i = w->row_of_col[j];
        w->col_seen[j] = 1;
        w->todo[best_t] = w->todo[--n_todo];
    }
    *out_min = reach;
    return sink;
}

/* cost: row-major (nr x nc) float64.  rows_out/cols_out: min(nr,nc) entries. */
int lsap_oracle_f64(const double *cost, int64_t nr, int64_t nc, int64_t *rows_out,
                    int64_t *cols_out) {
    if (nr < 0 || nc < 0 || nr > (1 << 20) || nc > (1 << 20)) return -1;
    if (nr == 0 || nc == 0) return LSAP_OK;
    const int flip = nc < nr;
    double *ct = NULL;
    const double *c = cost;
    if (flip) {
        ct = (double *)malloc((size_t)(nr * nc) * sizeof(double));
        if (!ct) return -1;
        for (int64_t a = 0; a < nr; ++a)
            for (int64_t b = 0; b < nc; ++b) ct[b * nr + a] = cost[a * nc + b];
        int64_t tmp = nr; nr = nc; nc = tmp;
        c = ct;
    }
    for (int64_t k = 0; k < nr * nc; ++k)
        if (c[k] != c[k] || c[k] == -INFINITY) { free(ct); return LSAP_INVALID; }

    lsap_ws w;
    if (!ws_alloc(&w, nr, nc)) { ws_free(&w); free(ct); return -1; }
    for (int64_t a = 0; a < nr; ++a) w.col_of_row[a] = -1;
    for (int64_t b = 0; b < nc; ++b) { w.row_of_col[b] = -1; w.pred_row[b] = -1; }

    int rc = LSAP_OK;
    for (int64_t cur = 0; cur < nr; ++cur) {
        double m = 0.0;
        int64_t sink = grow_path(c, nr, nc, cur, &w, &m);
        if (sink < 0) { rc = LSAP_INFEASIBLE; break; }
        /* dual update */
        w.u[cur] += m;
        for (int64_t a = 0; a < nr; ++a)
            if (w.row_seen[a] && a != cur) w.u[a] += m - w.dist[w.col_of_row[a]];
        for (int64_t b = 0; b < nc; ++b)
            if (w.col_seen[b]) w.v[b] -= m - w.dist[b];
        /* flip the matching along the path sink -> cur */
        int64_t j = sink;
        for (;;) {
            int64_t a = w.pred_row[j];
            w.row_of_col[j] = a;
            int64_t prev = w.col_of_row[a];
            w.col_of_row[a] = j;
            j = prev;
            if (a == cur) break;
        }
    }
    if (rc == LSAP_OK) {
        if (flip) {
            /* working rows are original columns: emit sorted by original row.
             * row_of_col[] over working columns (= original rows) is already that order. */
            int64_t k = 0;
            for (int64_t b = 0; b < nc; ++b)
                if (w.row_of_col[b] >= 0) { rows_out[k] = b; cols_out[k] = w.row_of_col[b]; ++k; }
        } else {
            for (int64_t a = 0; a < nr; ++a) { rows_out[a] = a; cols_out[a] = w.col_of_row[a]; }
        }
    }
    ws_free(&w);
    free(ct);
    return rc;
}

/* float32 entry point: the reference hands SciPy a float32 tensor (detr/matcher.py:93-94);
 * SciPy widens to float64, which is exact. */
int lsap_oracle_f32(const float *cost, int64_t nr, int64_t nc, int64_t *rows_out,
                    int64_t *cols_out) {
    if (nr == 0 || nc == 0) return LSAP_OK;
    double *wide = (double *)malloc((size_t)(nr * nc) * sizeof(double));
    if (!wide) return -1;
    for (int64_t k = 0; k < nr * nc; ++k) wide[k] = (double)cost[k];
    int rc = lsap_oracle_f64(wide, nr, nc, rows_out, cols_out);
    free(wide);
    return rc;
}

/* Batched convenience for the CPU-baseline timing in bench.py: `n` ragged problems,
 * problem p is (nr[p] x nc[p]) float32 at cost + off[p]; outputs at out_off[p]. */
int lsap_oracle_batch_f32(const float *cost, const int64_t *off, const int64_t *nr,
                          const int64_t *nc, int64_t n, const int64_t *out_off,
                          int64_t *rows_out, int64_t *cols_out) {
    for (int64_t p = 0; p < n; ++p) {
        int rc = lsap_oracle_f32(cost + off[p], nr[p], nc[p], rows_out + out_off[p],
                                 cols_out + out_off[p]);
        if (rc) return rc;
    }
    return LSAP_OK;
}
