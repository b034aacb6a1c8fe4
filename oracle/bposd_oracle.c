/*
 * bposd_oracle.c -- CPU restatement of the BP+OSD decode path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (bp_osd_b200/) never links, imports or calls it.
 *
 * What it restates: the arithmetic behind `bposd_decoder` / `BpOsdDecoder`, which the
 * reference imports from the third-party package `ldpc` (constraint `ldpc>=2.0.0`,
 * /root/reference/setup.py:30; re-exported at src/bposd/__init__.py:1; constructed at
 * src/bposd/css_decode_sim.py:444-463; called at css_decode_sim.py:174-202 and
 * README.md:176-197).  `ldpc` is NOT vendored in /root/reference and cannot be installed
 * offline, so this file restates the published algorithm (arXiv:2005.07016, cited at
 * README.md:3) with the iteration orders and tie conventions of ldpc v2 as listed in
 * SURVEY.md section 8(a), rows a1-a19.  Each function names the row it follows.
 *
 * PARITY STATUS: "parity unpinned" at the ldpc boundary -- the reference's own tests never
 * call the decoder (tests/test_css.py, test_hgp.py, test_stab.py only build codes).  The
 * one decoder known-answer in the reference, README.md:194-216 (vector G1), is reproduced
 * by tests/test_oracle.py; a second, independently written literal restatement
 * (oracle/slow_ref.py) is cross-checked against this file on random inputs.
 *
 * Arithmetic: IEEE double, no FMA contraction (build with -ffp-contract=off), sequential
 * accumulation in the documented edge orders.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* tanh / log of the product-sum update: the portable, FMA-free implementations shared with the CUDA
 * kernels (include/bposd_math.h) so that fp64 product-sum can be compared bit for bit, or the host libm
 * (what ldpc itself calls; its last bits depend on the glibc build and CPU). */
#include "../include/bposd_math.h"
#define ORACLE_MATH_SHARED 0
#define ORACLE_MATH_LIBM 1

#define BP_PRODUCT_SUM 0
#define BP_MINIMUM_SUM 1
#define OSD_0 0
#define OSD_E 1
#define OSD_CS 2

typedef struct {
    int m, n, nnz;
    int *row_ptr, *col_idx;          /* CSR, ascending column inside a row   */
    int *col_ptr, *row_idx, *to_csr; /* CSC, ascending row inside a column; to_csr = CSR edge id */
    double *probs;
    int max_iter, bp_method, osd_method, osd_order;
    double ms_scaling_factor;
    int rank, k;
    /* per-decode state (result attributes, row a15) */
    double *b2c, *c2b, *prior, *llr;
    uint8_t *bp_dec, *cand, *osd0, *osdw;
    int converge, iter, osd_ran;
    /* OSD scratch */
    int *order, *piv_col, *nonpiv;
    uint64_t *work;
    long stat_elim_wordxors;
    long total_elim_wordxors, total_osd; /* accumulated over decodes (algorithmic-op count of SURVEY.md 8d) */
    int math_mode;                       /* ORACLE_MATH_SHARED (default) or ORACLE_MATH_LIBM */
    int schedule;                        /* 0 parallel (flooding), 1 serial (row f4) */
    int *sched_order;                    /* serial schedule: bit order, NULL = 0 .. n-1 */
} oracle_t;

static double o_tanh(const oracle_t *o, double x) { return o->math_mode == ORACLE_MATH_LIBM ? tanh(x) : bpm_tanh(x); }
static double o_log(const oracle_t *o, double x) { return o->math_mode == ORACLE_MATH_LIBM ? log(x) : bpm_log(x); }
void oracle_set_math(oracle_t *o, int mode) { o->math_mode = mode; }
double oracle_math_tanh(double x) { return bpm_tanh(x); }
double oracle_math_log(double x) { return bpm_log(x); }
double oracle_math_expm1(double x) { return bpm_expm1(x); }
/* element-wise forms for the device-vs-host comparison of the shared header (fn: 0 a / b, 1 tanh, 2 log, 3 (1 + a) / (1 - a)) */
void oracle_math_map(int fn, const double *a, const double *b, double *out, long long count) {
    for (long long i = 0; i < count; i++)
        out[i] = fn == 0 ? bpm_div(a[i], b[i]) : fn == 1 ? bpm_tanh(a[i]) : fn == 2 ? bpm_log(a[i]) : (1 + a[i]) / (1 - a[i]);
}

/* ------------------------------------------------------------------ helpers */

static int gf2_rank_dense(const oracle_t *o) {
    /* plain row-major elimination in natural column order, used once at setup (row a1) */
    int m = o->m, n = o->n, W = (n + 63) / 64, r = 0;
    uint64_t *a = (uint64_t *)calloc((size_t)m * W, 8);
    for (int i = 0; i < m; i++)
        for (int e = o->row_ptr[i]; e < o->row_ptr[i + 1]; e++)
            a[(size_t)i * W + o->col_idx[e] / 64] ^= 1ull << (o->col_idx[e] % 64);
    for (int c = 0; c < n && r < m; c++) {
        int w = c / 64;
        uint64_t bit = 1ull << (c % 64);
        int p = -1;
        for (int i = r; i < m; i++)
            if (a[(size_t)i * W + w] & bit) { p = i; break; }
        if (p < 0) continue;
        if (p != r)
            for (int x = 0; x < W; x++) {
                uint64_t t = a[(size_t)p * W + x];
                a[(size_t)p * W + x] = a[(size_t)r * W + x];
                a[(size_t)r * W + x] = t;
            }
        for (int i = r + 1; i < m; i++)
            if (a[(size_t)i * W + w] & bit)
                for (int x = w; x < W; x++) a[(size_t)i * W + x] ^= a[(size_t)r * W + x];
        r++;
    }
    free(a);
    return r;
}

/* ------------------------------------------------------------------ lifecycle */

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    free(o->row_ptr); free(o->col_idx); free(o->col_ptr); free(o->row_idx); free(o->to_csr);
    free(o->probs); free(o->b2c); free(o->c2b); free(o->prior); free(o->llr);
    free(o->bp_dec); free(o->cand); free(o->osd0); free(o->osdw);
    free(o->order); free(o->piv_col); free(o->nonpiv); free(o->work); free(o->sched_order);
    free(o);
}

/* row a1: constructor.  CSR input must have ascending column indices inside each row. */
oracle_t *oracle_create(const int *row_ptr, const int *col_idx, int m, int n, const double *probs,
                        int max_iter, int bp_method, double ms_scaling_factor, int osd_method,
                        int osd_order) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    int nnz = row_ptr[m];
    o->m = m; o->n = n; o->nnz = nnz;
    o->row_ptr = (int *)malloc(sizeof(int) * (m + 1));
    o->col_idx = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    memcpy(o->row_ptr, row_ptr, sizeof(int) * (m + 1));
    memcpy(o->col_idx, col_idx, sizeof(int) * nnz);
    o->col_ptr = (int *)calloc(n + 1, sizeof(int));
    o->row_idx = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    o->to_csr = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    for (int e = 0; e < nnz; e++) o->col_ptr[col_idx[e] + 1]++;
    for (int j = 0; j < n; j++) o->col_ptr[j + 1] += o->col_ptr[j];
    int *fill = (int *)calloc(n, sizeof(int));
    for (int i = 0; i < m; i++) /* ascending i => ascending row inside each column */
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            int j = col_idx[e], p = o->col_ptr[j] + fill[j]++;
            o->row_idx[p] = i;
            o->to_csr[p] = e;
        }
    free(fill);
    o->probs = (double *)malloc(sizeof(double) * n);
    memcpy(o->probs, probs, sizeof(double) * n);
    o->max_iter = max_iter > 0 ? max_iter : n; /* max_iter=0 => n (css_decode_sim.py:72) */
    o->bp_method = bp_method;
    o->ms_scaling_factor = ms_scaling_factor;
    o->osd_method = osd_method;
    o->osd_order = (osd_method == OSD_0) ? 0 : osd_order;
    o->b2c = (double *)calloc(nnz > 0 ? nnz : 1, 8);
    o->c2b = (double *)calloc(nnz > 0 ? nnz : 1, 8);
    o->prior = (double *)calloc(n, 8);
    o->llr = (double *)calloc(n, 8);
    o->bp_dec = (uint8_t *)calloc(n, 1);
    o->cand = (uint8_t *)calloc(m > 0 ? m : 1, 1);
    o->osd0 = (uint8_t *)calloc(n, 1);
    o->osdw = (uint8_t *)calloc(n, 1);
    o->order = (int *)malloc(sizeof(int) * n);
    o->piv_col = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
    o->nonpiv = (int *)malloc(sizeof(int) * n);
    o->rank = gf2_rank_dense(o);
    o->k = n - o->rank;
    return o;
}

int oracle_rank(const oracle_t *o) { return o->rank; }
int oracle_k(const oracle_t *o) { return o->k; }
int oracle_converge(const oracle_t *o) { return o->converge; }
int oracle_iter(const oracle_t *o) { return o->iter; }
int oracle_osd_ran(const oracle_t *o) { return o->osd_ran; }
long oracle_stat_elim_wordxors(const oracle_t *o) { return o->stat_elim_wordxors; }
long oracle_total_elim_wordxors(const oracle_t *o) { return o->total_elim_wordxors; }
long oracle_total_osd(const oracle_t *o) { return o->total_osd; }
const double *oracle_llr(const oracle_t *o) { return o->llr; }
const uint8_t *oracle_bp_decoding(const oracle_t *o) { return o->bp_dec; }
const uint8_t *oracle_osd0_decoding(const oracle_t *o) { return o->osd0; }
const uint8_t *oracle_osdw_decoding(const oracle_t *o) { return o->osdw; }

/* row a16: update_channel_probs (css_decode_sim.py:229,248) */
void oracle_update_channel_probs(oracle_t *o, const double *probs) {
    memcpy(o->probs, probs, sizeof(double) * o->n);
}

/* ------------------------------------------------------------------ BP (rows a3-a8) */

/* row f4: ldpc's BpOsdDecoder(schedule="serial", serial_schedule_order=...), an option the reference never passes.
 * order: a permutation of 0 .. n-1 (copied), or NULL for the natural order. */
void oracle_set_schedule(oracle_t *o, int schedule, const int *order) {
    free(o->sched_order);
    o->sched_order = NULL;
    o->schedule = schedule;
    if (schedule == 1 && order) {
        o->sched_order = (int *)malloc(sizeof(int) * (size_t)(o->n > 0 ? o->n : 1));
        memcpy(o->sched_order, order, sizeof(int) * (size_t)o->n);
    }
}

/* Serial schedule, restated from upstream ldpc's bp_decode_serial (src_cpp/bp.hpp of ldpc >= 2.0; absent from
 * /root/reference, "parity unpinned" like the rest of this file).  Inside an iteration the bits are visited one after the
 * other; bit j, for each of its edges in ascending check order: the check-to-bit message from the CURRENT bit-to-check
 * messages of the check's other edges (min-sum: smallest magnitude, sign = syndrome + count of messages <= 0, scaled by
 * alpha as in the parallel schedule; product-sum: product of tanh(b2c / 2) over the others in ascending column order,
 * +-log((1 + x) / (1 - x))); the edge's bit-to-check message becomes the running sum "prior + messages of the edges
 * before it", the message is added to the sum.  The sum is the bit's log-probability ratio (hard decision: 1 iff <= 0);
 * then, from the last edge backwards, every bit-to-check message also receives the messages of the edges after it.  The
 * candidate syndrome is tested at the end of every iteration. */
static void bp_decode_serial(oracle_t *o, const uint8_t *synd) {
    const int m = o->m, n = o->n;
    double *b2c = o->b2c, *c2b = o->c2b;
    o->converge = 0;
    o->iter = 0;
    for (int j = 0; j < n; j++) {
        o->prior[j] = log((1.0 - o->probs[j]) / o->probs[j]);
        o->llr[j] = o->prior[j];
        o->bp_dec[j] = 0;
        for (int p = o->col_ptr[j]; p < o->col_ptr[j + 1]; p++) b2c[o->to_csr[p]] = o->prior[j];
    }
    for (int it = 1; it <= o->max_iter; it++) {
        const double alpha = (o->ms_scaling_factor == 0.0) ? 1.0 - pow(2.0, -1.0 * it) : o->ms_scaling_factor;
        for (int k = 0; k < n; k++) {
            const int j = o->sched_order ? o->sched_order[k] : k;
            double llr = o->prior[j];
            for (int p = o->col_ptr[j]; p < o->col_ptr[j + 1]; p++) {
                const int e = o->to_csr[p], i = o->row_idx[p];
                double c;
                if (o->bp_method == BP_PRODUCT_SUM) {
                    double x = 1.0;
                    for (int g = o->row_ptr[i]; g < o->row_ptr[i + 1]; g++)
                        if (g != e) x *= o_tanh(o, b2c[g] / 2);
                    c = (synd[i] ? -1.0 : 1.0) * o_log(o, (1 + x) / (1 - x));
                } else {
                    int sgn = synd[i];
                    double t = DBL_MAX;
                    for (int g = o->row_ptr[i]; g < o->row_ptr[i + 1]; g++) {
                        if (g == e) continue;
                        const double a = fabs(b2c[g]);
                        if (a < t) t = a;
                        if (b2c[g] <= 0) sgn += 1;
                    }
                    c = ((sgn % 2 == 0) ? 1.0 : -1.0) * alpha * t;
                }
                c2b[e] = c;
                b2c[e] = llr;
                llr += c;
            }
            o->llr[j] = llr;
            o->bp_dec[j] = (llr <= 0) ? 1 : 0;
            double t = 0;
            for (int p = o->col_ptr[j + 1] - 1; p >= o->col_ptr[j]; p--) {
                const int e = o->to_csr[p];
                b2c[e] += t;
                t += c2b[e];
            }
        }
        int same = 1;
        for (int i = 0; i < m && same; i++) {
            int par = 0;
            for (int e = o->row_ptr[i]; e < o->row_ptr[i + 1]; e++) par ^= o->bp_dec[o->col_idx[e]];
            if (par != (synd[i] & 1)) same = 0;
        }
        o->iter = it;
        if (same) { o->converge = 1; return; }
    }
}

static void bp_decode(oracle_t *o, const uint8_t *synd) {
    if (o->schedule == 1) { bp_decode_serial(o, synd); return; }
    const int m = o->m, n = o->n;
    double *b2c = o->b2c, *c2b = o->c2b;
    o->converge = 0;
    o->iter = 0;
    /* a3: priors and initial bit->check messages */
    for (int j = 0; j < n; j++) {
        o->prior[j] = log((1.0 - o->probs[j]) / o->probs[j]);
        o->llr[j] = o->prior[j];
        o->bp_dec[j] = 0;
        for (int p = o->col_ptr[j]; p < o->col_ptr[j + 1]; p++) b2c[o->to_csr[p]] = o->prior[j];
    }
    for (int it = 1; it <= o->max_iter; it++) {
        if (o->bp_method == BP_PRODUCT_SUM) {
            /* a5: forward products left to right, reverse products right to left */
            for (int i = 0; i < m; i++) {
                o->cand[i] = 0;
                double t = 1.0;
                for (int e = o->row_ptr[i]; e < o->row_ptr[i + 1]; e++) {
                    c2b[e] = t;
                    t *= o_tanh(o, b2c[e] / 2);
                }
                t = 1.0;
                for (int e = o->row_ptr[i + 1] - 1; e >= o->row_ptr[i]; e--) {
                    c2b[e] *= t;
                    double sgn = synd[i] ? -1.0 : 1.0;
                    c2b[e] = sgn * o_log(o, (1 + c2b[e]) / (1 - c2b[e]));
                    t *= o_tanh(o, b2c[e] / 2);
                }
            }
        } else {
            /* a4: min-sum with running minima; alpha = 1 - 2^-it when the factor is 0 */
            double alpha = (o->ms_scaling_factor == 0.0) ? 1.0 - pow(2.0, -1.0 * it)
                                                         : o->ms_scaling_factor;
            for (int i = 0; i < m; i++) {
                o->cand[i] = 0;
                int tot = synd[i];
                double t = DBL_MAX;
                for (int e = o->row_ptr[i]; e < o->row_ptr[i + 1]; e++) {
                    if (b2c[e] <= 0) tot += 1;
                    c2b[e] = t;
                    double a = fabs(b2c[e]);
                    if (a < t) t = a;
                }
                t = DBL_MAX;
                for (int e = o->row_ptr[i + 1] - 1; e >= o->row_ptr[i]; e--) {
                    int sgn = tot;
                    if (b2c[e] <= 0) sgn += 1;
                    if (t < c2b[e]) c2b[e] = t;
                    double ms = (sgn % 2 == 0) ? 1.0 : -1.0;
                    c2b[e] *= ms * alpha;
                    double a = fabs(b2c[e]);
                    if (a < t) t = a;
                }
            }
        }
        /* a6: bit update, hard decision, candidate syndrome */
        for (int j = 0; j < n; j++) {
            double t = o->prior[j];
            for (int p = o->col_ptr[j]; p < o->col_ptr[j + 1]; p++) {
                int e = o->to_csr[p];
                b2c[e] = t;
                t += c2b[e];
            }
            o->llr[j] = t;
            if (t <= 0) {
                o->bp_dec[j] = 1;
                for (int p = o->col_ptr[j]; p < o->col_ptr[j + 1]; p++) o->cand[o->row_idx[p]] ^= 1;
            } else {
                o->bp_dec[j] = 0;
            }
        }
        /* a7: convergence test, every iteration */
        int same = 1;
        for (int i = 0; i < m; i++)
            if ((o->cand[i] & 1) != (synd[i] & 1)) { same = 0; break; }
        o->iter = it;
        if (same) { o->converge = 1; return; }
        /* a8: bit->check, descending row order */
        for (int j = 0; j < n; j++) {
            double t = 0;
            for (int p = o->col_ptr[j + 1] - 1; p >= o->col_ptr[j]; p--) {
                int e = o->to_csr[p];
                b2c[e] += t;
                t += c2b[e];
            }
        }
    }
}

/* ------------------------------------------------------------------ OSD (rows a9-a14) */

typedef struct { double v; int idx; } key_t_;

static void merge_sort_keys(key_t_ *a, key_t_ *tmp, int lo, int hi) {
    /* stable ascending merge sort on the LLR value: ties keep ascending index (row a9) */
    if (hi - lo < 2) return;
    int mid = lo + (hi - lo) / 2;
    merge_sort_keys(a, tmp, lo, mid);
    merge_sort_keys(a, tmp, mid, hi);
    int i = lo, j = mid, k = lo;
    while (i < mid && j < hi) tmp[k++] = (a[j].v < a[i].v) ? a[j++] : a[i++];
    while (i < mid) tmp[k++] = a[i++];
    while (j < hi) tmp[k++] = a[j++];
    memcpy(a + lo, tmp + lo, sizeof(key_t_) * (hi - lo));
}

static double soft_weight(const oracle_t *o, const uint8_t *x) {
    /* a14: W(x) = sum over set bits, ascending j, of log(1/p_j) */
    double w = 0;
    for (int j = 0; j < o->n; j++)
        if (x[j]) w += log(1 / o->probs[j]);
    return w;
}

static void osd_decode(oracle_t *o, const uint8_t *synd) {
    const int m = o->m, n = o->n;
    o->osd_ran = 1;
    /* a9: column order, least reliable first */
    key_t_ *keys = (key_t_ *)malloc(sizeof(key_t_) * 2 * n);
    for (int j = 0; j < n; j++) { keys[j].v = o->llr[j]; keys[j].idx = j; }
    merge_sort_keys(keys, keys + n, 0, n);
    for (int t = 0; t < n; t++) o->order[t] = keys[t].idx;
    free(keys);

    /* a10: Gauss-Jordan on [H(:,order) | I_m], row-major bit-packed, columns scanned in order */
    const int WH = (n + 63) / 64, WT = (m + 63) / 64, W = WH + WT;
    uint64_t *a = (uint64_t *)calloc((size_t)m * W, 8);
    int *inv = (int *)malloc(sizeof(int) * n);
    for (int t = 0; t < n; t++) inv[o->order[t]] = t;
    for (int i = 0; i < m; i++) {
        for (int e = o->row_ptr[i]; e < o->row_ptr[i + 1]; e++) {
            int t = inv[o->col_idx[e]];
            a[(size_t)i * W + t / 64] ^= 1ull << (t % 64);
        }
        a[(size_t)i * W + WH + i / 64] ^= 1ull << (i % 64);
    }
    int rank = 0, nnp = 0, max_rank = m < n ? m : n;
    uint8_t *is_piv = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    long xors = 0;
    for (int t = 0; t < n; t++) {
        if (rank == max_rank) break;
        int w = t / 64;
        uint64_t bit = 1ull << (t % 64);
        int p = -1;
        for (int i = rank; i < m; i++)
            if (a[(size_t)i * W + w] & bit) { p = i; break; }
        if (p < 0) continue;
        if (p != rank)
            for (int x = 0; x < W; x++) {
                uint64_t tmp = a[(size_t)p * W + x];
                a[(size_t)p * W + x] = a[(size_t)rank * W + x];
                a[(size_t)rank * W + x] = tmp;
            }
        const uint64_t *pr = a + (size_t)rank * W;
        for (int i = 0; i < m; i++) {
            if (i == rank || !(a[(size_t)i * W + w] & bit)) continue;
            uint64_t *ri = a + (size_t)i * W;
            for (int x = w; x < W; x++) ri[x] ^= pr[x]; /* pivot row is zero left of word w */
            xors += (W - w) * 2;                        /* counted as 32-bit word XORs */
        }
        o->piv_col[rank] = t; /* position in the sorted order */
        is_piv[t] = 1;
        rank++;
    }
    o->stat_elim_wordxors = xors;
    o->total_elim_wordxors += xors;
    o->total_osd += 1;
    for (int t = 0; t < n; t++)
        if (!is_piv[t]) o->nonpiv[nnp++] = t; /* non-pivots keep sorted order */

    /* a11: OSD-0.  s' = T s; x[pivot col of row r] = s'[r], x_T = 0 */
    uint8_t *sp = (uint8_t *)calloc(m > 0 ? m : 1, 1);
    for (int r = 0; r < m; r++) {
        const uint64_t *tr = a + (size_t)r * W + WH;
        int acc = 0;
        for (int i = 0; i < m; i++)
            if (synd[i] & 1) acc ^= (int)((tr[i / 64] >> (i % 64)) & 1);
        sp[r] = (uint8_t)acc;
    }
    memset(o->osd0, 0, n);
    for (int r = 0; r < rank; r++) o->osd0[o->order[o->piv_col[r]]] = sp[r];
    memcpy(o->osdw, o->osd0, n);

    int order_w = o->osd_order;
    if (o->osd_method != OSD_0 && order_w > 0) {
        double best = soft_weight(o, o->osd0);
        uint8_t *x = (uint8_t *)malloc(n);
        int kk = n - rank;
        long ncand;
        if (o->osd_method == OSD_E) ncand = (1L << order_w) - 1;
        else ncand = kk + (long)order_w * (order_w - 1) / 2;
        int pi = 0, pj = 0; /* pair cursor for the weight-2 sweep */
        for (long c = 0; c < ncand; c++) {
            int sel[64], ns = 0;
            if (o->osd_method == OSD_E) {
                /* a12: bit b of (c+1) <-> b-th non-pivot */
                long v = c + 1;
                for (int b = 0; b < order_w; b++)
                    if ((v >> b) & 1) sel[ns++] = b;
            } else if (c < kk) {
                sel[ns++] = (int)c; /* a13: all k weight-1 strings, in order */
            } else {
                if (c == kk) { pi = 0; pj = 1; }
                sel[ns++] = pi; sel[ns++] = pj; /* a13: pairs i<j<w, i outer */
                if (++pj >= order_w) { pi++; pj = pi + 1; }
            }
            /* a14: x_S from s' + reduced columns of the selected non-pivots */
            memset(x, 0, n);
            for (int r = 0; r < rank; r++) {
                int v = sp[r];
                for (int q = 0; q < ns; q++) {
                    int t = o->nonpiv[sel[q]];
                    v ^= (int)((a[(size_t)r * W + t / 64] >> (t % 64)) & 1);
                }
                x[o->order[o->piv_col[r]]] = (uint8_t)v;
            }
            for (int q = 0; q < ns; q++) x[o->order[o->nonpiv[sel[q]]]] = 1;
            double wgt = soft_weight(o, x);
            if (wgt < best) { /* strict: earliest candidate wins ties, OSD-0 wins all ties */
                best = wgt;
                memcpy(o->osdw, x, n);
            }
        }
        free(x);
    }
    free(sp); free(is_piv); free(inv); free(a);
}

/* row a2: decode(syndrome) */
void oracle_decode(oracle_t *o, const uint8_t *synd) {
    o->osd_ran = 0;
    o->stat_elim_wordxors = 0;
    bp_decode(o, synd);
    if (o->converge) {
        memcpy(o->osd0, o->bp_dec, o->n);
        memcpy(o->osdw, o->bp_dec, o->n);
    } else {
        osd_decode(o, synd);
    }
}

/* batch driver: a plain loop over decode(), used by tests and by the CPU-baseline timer */
void oracle_decode_batch(oracle_t *o, const uint8_t *synd, long B, uint8_t *osdw, uint8_t *osd0,
                         uint8_t *bp, double *llr, uint8_t *converge, int32_t *iter) {
    for (long b = 0; b < B; b++) {
        oracle_decode(o, synd + b * o->m);
        if (osdw) memcpy(osdw + b * o->n, o->osdw, o->n);
        if (osd0) memcpy(osd0 + b * o->n, o->osd0, o->n);
        if (bp) memcpy(bp + b * o->n, o->bp_dec, o->n);
        if (llr) memcpy(llr + b * o->n, o->llr, sizeof(double) * o->n);
        if (converge) converge[b] = (uint8_t)o->converge;
        if (iter) iter[b] = o->iter;
    }
}

/* ------------------------------------------------------------------ harness step (rows a17-a19) */

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    memcpy(out, ctr, 16);
    philox4x32_10(out, key[0], key[1]);
}

/* a17 (device-side variant): one 32-bit uniform per qubit from Philox4x32-10 with
 * counter = (j/4, 0, shot_lo, shot_hi), key = seed; lane j%4.  Pauli split as
 * css_decode_sim.py:471-496 with integer thresholds t1<=t2<=t3 per qubit:
 * r<t1 -> Z; t1<=r<t2 -> X; t2<=r<t3 -> Y. */
void oracle_sample_errors(uint64_t seed, uint64_t shot0, long B, int n, const uint32_t *t1,
                          const uint32_t *t2, const uint32_t *t3, uint8_t *ex, uint8_t *ez) {
    for (long b = 0; b < B; b++) {
        uint64_t g = shot0 + (uint64_t)b;
        for (int q = 0; q * 4 < n; q++) {
            uint32_t c[4] = {(uint32_t)q, 0u, (uint32_t)g, (uint32_t)(g >> 32)};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            for (int l = 0; l < 4 && q * 4 + l < n; l++) {
                int j = q * 4 + l;
                uint32_t r = c[l];
                int z = (r < t1[j]) || (r >= t2[j] && r < t3[j]);
                int x = (r >= t1[j] && r < t3[j]);
                if (ez) ez[b * n + j] = (uint8_t)z;
                if (ex) ex[b * n + j] = (uint8_t)x;
            }
        }
    }
}

/* a18: syndrome = H e mod 2 (css_decode_sim.py:173; README.md:196) */
void oracle_syndrome(const oracle_t *o, const uint8_t *e, long B, uint8_t *synd) {
    for (long b = 0; b < B; b++)
        for (int i = 0; i < o->m; i++) {
            int acc = 0;
            for (int p = o->row_ptr[i]; p < o->row_ptr[i + 1]; p++) acc ^= e[b * o->n + o->col_idx[p]] & 1;
            synd[b * o->m + i] = (uint8_t)acc;
        }
}

/* a19: logical failure of a residual, (L @ (e ^ d) % 2).any() (css_decode_sim.py:257-261) */
void oracle_logical_fail(const int *l_row_ptr, const int *l_col_idx, int K, int n, const uint8_t *e,
                         const uint8_t *d, long B, uint8_t *fail) {
    for (long b = 0; b < B; b++) {
        int any = 0;
        for (int r = 0; r < K && !any; r++) {
            int acc = 0;
            for (int p = l_row_ptr[r]; p < l_row_ptr[r + 1]; p++) {
                int j = l_col_idx[p];
                acc ^= (e[b * n + j] ^ d[b * n + j]) & 1;
            }
            any |= acc;
        }
        fail[b] = (uint8_t)any;
    }
}
