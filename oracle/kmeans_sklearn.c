/* CPU oracle: scikit-learn's KMeans(k, random_state=42, n_init='auto').fit_predict on uint8 colours,
 * restated operation by operation in IEEE double.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): compiled by oracle/Makefile into
 * oracle/_build/libkmeans_sklearn.so, loaded only by tests/, smoke() and bench.py's CPU legs.
 *
 * The reference calls it at /root/reference/encoder/compression/clustering.py:751-752.  scikit-learn is
 * an un-vendored dependency (requirements.txt:6, unpinned; 1.9.0 with OpenBLAS 0.3.30/0.3.31 "SkylakeX"
 * kernels in the build container, which is where tests/golden/ was recorded).  The labels depend on the
 * rounding of every floating-point step, so this file follows the arithmetic of that stack exactly:
 *
 *   sklearn/cluster/_kmeans.py:1464-1490   float64 copy, tolerance, centring  (X -= X.mean(axis=0))
 *   sklearn/cluster/_kmeans.py:216-282     k-means++ seeding
 *   sklearn/metrics/pairwise.py            _euclidean_distances: -2 X.Y^T + |x|^2 + |y|^2, clipped at 0
 *   sklearn/utils/extmath.py:89            row_norms = einsum('ij,ij->i')
 *   sklearn/cluster/_k_means_lloyd.pyx     E step in chunks of 256 samples: |c|^2 - 2 x.c through dgemm,
 *                                          first minimum; M step: sums in sample order
 *   sklearn/cluster/_k_means_common.pyx    _average_centers, _center_shift, empty-cluster relocation
 *   sklearn/cluster/_kmeans.py:703-757     Lloyd driver: strict convergence, tolerance, final E step
 *
 * and the summation orders of the library kernels underneath (established by bit-comparison against the
 * libraries in the build container; tools/kmeans_replay.py re-checks them):
 *
 *   einsum 'ij,ij->i', 3 columns    (x0*x0 + x2*x2) + x1*x1              (2-lane SIMD, horizontal add)
 *   dgemm, inner dimension 3        fma(x2,c2, fma(x1,c1, x0*c0)); C = fma(alpha, acc, C)
 *       except, in Lloyd's dgemm (M = k > 192 is cut into blocks; OpenBLAS level3.c, DGEMM_P = 192): the
 *       4-row edge of the last block against sample groups of 12 uses fma(x2,c2, x0*c0 + x1*c1)
 *   dgemv (1 x 3 times 3 x n)       fma(x2,c2, fma(x0,c0, x1*c1))       (first centre's distances)
 *   ddot(v, ones)                   32 strided partial sums (4 x 8 lanes), folded 8->4 lanes, 16-element
 *                                   steps on 4 x 4 lanes, ((a0+a1)+a2)+a3, (l0+l2)+(l1+l3), scalar tail
 *   dgemv_t(D[t,n], ones)           per column, row blocks of 2048: 4 lanes ((l0+l2)+(l1+l3)) for column
 *                                   groups of 4 and single columns, 2 lanes (l0+l1) for a column pair;
 *                                   blocks added in order; 1-3 tail rows ((a0+a1)+a2)
 *   np.sum (contiguous)             numpy's pairwise summation (8 accumulators, blocks of 128)
 *   np.cumsum, reductions over axis 0: sequential
 *
 * The OpenMP reduction of the M step is taken in its single-thread order (sample order); scikit-learn
 * with 1 and with 8 threads reproduces every recorded call of tests/golden/, so the goldens do not
 * depend on it.  Empty-cluster relocation with more than one empty cluster uses numpy's argpartition in
 * scikit-learn, whose order among the selected elements is an implementation detail: here the farthest
 * points are taken in descending distance (ties: higher index first); no recorded call relocates.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC   (fma() must be the fused instruction and no
 * other contraction may happen).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 256          /* sklearn/cluster/_k_means_common.pyx:13 */
#define GEMM_P 192         /* OpenBLAS DGEMM_DEFAULT_P for SKYLAKEX */
#define GEMM_UNROLL_M 16
#define NBMAX_T 2048       /* OpenBLAS kernel/x86_64/dgemv_t_4.c */

static double norm3(const double* c) { return (c[0] * c[0] + c[2] * c[2]) + c[1] * c[1]; }
static double dot_gemm(const double* x, const double* c) { return fma(x[2], c[2], fma(x[1], c[1], x[0] * c[0])); }
static double dot_gemm_edge(const double* x, const double* c) { return fma(x[2], c[2], x[0] * c[0] + x[1] * c[1]); }
static double dot_gemv(const double* x, const double* c) { return fma(x[2], c[2], fma(x[0], c[0], x[1] * c[1])); }

/* numpy pairwise summation of a contiguous double vector */
static double np_pairwise(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}

/* OpenBLAS SkylakeX ddot(v, ones) */
double km_ddot_ones(const double* x, int n) {
    double dot = 0.0;
    int n1 = n & -16, i = 0;
    if (n1) {
        double a[4][8];
        memset(a, 0, sizeof a);
        int n32 = n1 & ~31;
        for (; i < n32; i += 32)
            for (int q = 0; q < 4; q++)
                for (int l = 0; l < 8; l++) a[q][l] += x[i + 8 * q + l];
        double b[4][4];
        for (int q = 0; q < 4; q++)
            for (int l = 0; l < 4; l++) b[q][l] = a[q][l] + a[q][l + 4];
        for (; i < n1; i += 16)
            for (int q = 0; q < 4; q++)
                for (int l = 0; l < 4; l++) b[q][l] += x[i + 4 * q + l];
        double c[4];
        for (int l = 0; l < 4; l++) c[l] = ((b[0][l] + b[1][l]) + b[2][l]) + b[3][l];
        dot = (c[0] + c[2]) + (c[1] + c[3]);
    }
    for (; i < n; i++) dot += x[i];
    return dot;
}

/* OpenBLAS dgemv_t(D, ones): sum of one column of length m; lanes = 4 (column groups of 4, single
 * columns) or 2 (column pairs) */
static double gemv_block(const double* a, int nb, int lanes) {
    if (lanes == 4) {
        double l[4] = {0, 0, 0, 0};
        for (int i = 0; i < nb; i += 4)
            for (int q = 0; q < 4; q++) l[q] += a[i + q];
        return (l[0] + l[2]) + (l[1] + l[3]);
    }
    double l0 = 0, l1 = 0;
    for (int i = 0; i < nb; i += 2) { l0 += a[i]; l1 += a[i + 1]; }
    return l0 + l1;
}
double km_dgemv_t_ones(const double* a, int m, int lanes) {
    double y = 0.0;
    int m3 = m & 3, m1 = m & -4, m2 = (m & (NBMAX_T - 1)) - m3;
    int nb = NBMAX_T;
    const double* ap = a;
    while (nb == NBMAX_T) {
        m1 -= nb;
        if (m1 < 0) {
            if (m2 == 0) break;
            nb = m2;
        }
        y += gemv_block(ap, nb, lanes);
        ap += nb;
    }
    if (m3 == 3) y += (ap[0] + ap[1]) + ap[2];
    else if (m3 == 2) y += ap[0] + ap[1];
    else if (m3 == 1) y += ap[0];
    return y;
}
/* lanes used by column j of t columns: groups of 4, then a pair if (t & 2), then a single */
static int gemv_lanes(int j, int t) {
    int g = (t >> 2) * 4;
    if (j < g) return 4;
    return ((t & 2) && j - g < 2) ? 2 : 4;
}

/* numpy RandomState.choice(n, p=uniform): first index whose normalised cumulative probability
 * exceeds u (numpy/random/mtrand.pyx: cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(side='right')) */
int km_first_seed(int n, double u) {
    double p = 1.0 / (double)n, last = 0.0;
    for (int i = 0; i < n; i++) last += p;
    double c = 0.0;
    for (int i = 0; i < n; i++) {
        c += p;
        if (c / last > u) return i;
    }
    return n - 1;
}

/* Row range [lo, hi) of the dgemm result (cluster index) computed by the 4-row edge kernel of a block
 * after the first (OpenBLAS driver/level3/level3.c); lo = hi = 0 when there is none. */
void km_gemm_edge_rows(int k, int* lo, int* hi) {
    *lo = *hi = 0;
    if (k <= GEMM_P) return;
    int is = 0, min_i = k;
    for (;;) {
        min_i = k - is;
        if (min_i >= 2 * GEMM_P) min_i = GEMM_P;
        else if (min_i > GEMM_P) min_i = ((min_i / 2 + GEMM_UNROLL_M - 1) / GEMM_UNROLL_M) * GEMM_UNROLL_M;
        if (is + min_i >= k) break;
        is += min_i;
    }
    if (is > 0 && (min_i & 4)) {
        *lo = is + (min_i & ~15) + (min_i & 8);
        *hi = *lo + 4;
    }
}

typedef struct {
    int n, k;
    double* xc;      /* [n,3] centred */
    double* xsn;     /* [n] */
    double tol;
} km_problem;

static void e_step(const km_problem* p, const double* centers, int32_t* labels) {
    int n = p->n, k = p->k;
    double* csn = (double*)malloc(sizeof(double) * k);
    for (int j = 0; j < k; j++) csn[j] = norm3(centers + 3 * j);
    int elo, ehi;
    km_gemm_edge_rows(k, &elo, &ehi);
    for (int s = 0; s < n; s += CHUNK) {
        int m = n - s < CHUNK ? n - s : CHUNK;
        int lim12 = (m / 12) * 12;
        for (int r = 0; r < m; r++) {
            const double* x = p->xc + 3 * (s + r);
            double best = 0;
            int lab = 0;
            for (int j = 0; j < k; j++) {
                double acc = (j >= elo && j < ehi && r < lim12) ? dot_gemm_edge(x, centers + 3 * j)
                                                                 : dot_gemm(x, centers + 3 * j);
                double d = fma(-2.0, acc, csn[j]);
                if (j == 0 || d < best) { best = d; lab = j; }
            }
            labels[s + r] = lab;
        }
    }
    free(csn);
}

/* Labels of n points against k centres as sklearn's _labels_inertia computes them (the same chunked E step
 * with update_centers == False): used by oracle/minibatch_restated.py, whose batches and final prediction
 * go through it.  x[n,3] and centers[k,3] as given (MiniBatchKMeans does not centre the data). */
void km_estep_labels(const double* x, int n, const double* centers, int k, int32_t* labels) {
    km_problem P;
    P.n = n; P.k = k; P.xc = (double*)x; P.xsn = 0; P.tol = 0.0;
    e_step(&P, centers, labels);
}

static int cmp_far(const void* a, const void* b, void* ctx) {
    const double* d = (const double*)ctx;
    int ia = *(const int*)a, ib = *(const int*)b;
    if (d[ia] != d[ib]) return d[ia] > d[ib] ? -1 : 1;
    return ib - ia;
}

/* seeds_out[k]; labels_out[n]; r = RandomState(42).random_sample stream, at least 1 + (k-1)*t values.
 * info_out[0] = Lloyd iterations, [1] = strict convergence flag, [2] = relocations performed. */
int km_sklearn_labels(const uint8_t* colors, int n, int k, const double* r, int32_t* labels_out,
                      int64_t* seeds_out, int32_t* info_out) {
    if (k < 1 || k > n) return -1;
    km_problem P;
    P.n = n; P.k = k;
    double* x = (double*)malloc(sizeof(double) * 3 * n);
    P.xc = x;
    P.xsn = (double*)malloc(sizeof(double) * n);
    double mean[3], var[3];
    for (int d = 0; d < 3; d++) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += (double)colors[3 * i + d];
        mean[d] = s / (double)n;
    }
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) x[3 * i + d] = (double)colors[3 * i + d] - mean[d];
    for (int d = 0; d < 3; d++) {                       /* np.var(X, axis=0) */
        double s = 0.0;
        for (int i = 0; i < n; i++) s += x[3 * i + d] * x[3 * i + d];
        var[d] = s / (double)n;
    }
    P.tol = (((var[0] + var[1]) + var[2]) / 3.0) * 1e-4;
    for (int i = 0; i < n; i++) P.xsn[i] = norm3(x + 3 * i);

    /* ---- k-means++ (sklearn/cluster/_kmeans.py:216-282) */
    int t = 2 + (int)log((double)k);
    int ri = 0;
    int first = km_first_seed(n, r[ri++]);
    int64_t* seeds = seeds_out;
    seeds[0] = first;
    double* closest = (double*)malloc(sizeof(double) * n);
    double* cum = (double*)malloc(sizeof(double) * n);
    double* dist = (double*)malloc(sizeof(double) * n * t);
    {
        const double* c = x + 3 * first;
        double xx = norm3(c);
        for (int i = 0; i < n; i++) {
            double d = -2.0 * dot_gemv(x + 3 * i, c);
            d += xx;
            d += P.xsn[i];
            closest[i] = d > 0 ? d : 0.0;
        }
    }
    double pot = km_ddot_ones(closest, n);
    int cand[64];
    for (int c = 1; c < k; c++) {
        double s = 0.0;
        for (int i = 0; i < n; i++) { s += closest[i]; cum[i] = s; }
        for (int j = 0; j < t; j++) {
            double rv = r[ri++] * pot;
            int lo = 0, hi = n;                       /* searchsorted side='left' */
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cum[mid] < rv) lo = mid + 1; else hi = mid;
            }
            cand[j] = lo > n - 1 ? n - 1 : lo;
        }
        int best = 0;
        double best_pot = 0;
        for (int j = 0; j < t; j++) {
            const double* cc = x + 3 * cand[j];
            double xx = P.xsn[cand[j]];
            double* dj = dist + (size_t)j * n;
            for (int i = 0; i < n; i++) {
                double d = -2.0 * dot_gemm(cc, x + 3 * i);
                d += xx;
                d += P.xsn[i];
                d = d > 0 ? d : 0.0;
                dj[i] = closest[i] < d ? closest[i] : d;   /* np.minimum(closest, d) */
            }
            double pj = t == 1 ? km_ddot_ones(dj, n) : km_dgemv_t_ones(dj, n, gemv_lanes(j, t));
            if (j == 0 || pj < best_pot) { best_pot = pj; best = j; }
        }
        pot = best_pot;
        memcpy(closest, dist + (size_t)best * n, sizeof(double) * n);
        seeds[c] = cand[best];
    }

    /* ---- Lloyd (sklearn/cluster/_kmeans.py:703-757) */
    double* centers = (double*)malloc(sizeof(double) * 3 * k);
    double* cnew = (double*)malloc(sizeof(double) * 3 * k);
    double* w = (double*)malloc(sizeof(double) * k);
    double* shift2 = (double*)malloc(sizeof(double) * k);
    int32_t* labels = labels_out;
    int32_t* labels_old = (int32_t*)malloc(sizeof(int32_t) * n);
    for (int j = 0; j < k; j++) memcpy(centers + 3 * j, x + 3 * seeds[j], 3 * sizeof(double));
    for (int i = 0; i < n; i++) labels_old[i] = -1;
    int strict = 0, it, relocations = 0;
    for (it = 0; it < 300; it++) {
        e_step(&P, centers, labels);
        memset(cnew, 0, sizeof(double) * 3 * k);
        memset(w, 0, sizeof(double) * k);
        for (int i = 0; i < n; i++) {
            int l = labels[i];
            w[l] += 1.0;
            for (int d = 0; d < 3; d++) cnew[3 * l + d] += x[3 * i + d];
        }
        int n_empty = 0;
        for (int j = 0; j < k; j++) n_empty += w[j] == 0;
        if (n_empty) {                                   /* _k_means_common.pyx:167-211 */
            double* dd = (double*)malloc(sizeof(double) * n);
            int* order = (int*)malloc(sizeof(int) * n);
            double mx = 0;
            for (int i = 0; i < n; i++) {
                const double* c = centers + 3 * labels[i];
                double a0 = x[3 * i] - c[0], a1 = x[3 * i + 1] - c[1], a2 = x[3 * i + 2] - c[2];
                dd[i] = (a0 * a0 + a1 * a1) + a2 * a2;
                if (dd[i] > mx) mx = dd[i];
                order[i] = i;
            }
            if (mx != 0) {
                qsort_r(order, n, sizeof(int), cmp_far, dd);
                int e = 0;
                for (int j = 0; j < k; j++) {
                    if (w[j] != 0) continue;
                    int f = order[e++], o = labels[f];
                    for (int d = 0; d < 3; d++) {
                        cnew[3 * o + d] -= x[3 * f + d] * 1.0;
                        cnew[3 * j + d] = x[3 * f + d] * 1.0;
                    }
                    w[j] = 1.0;
                    w[o] -= 1.0;
                    relocations++;
                }
            }
            free(dd);
            free(order);
        }
        int am = 0;
        for (int j = 1; j < k; j++) if (w[j] > w[am]) am = j;
        for (int j = 0; j < k; j++) {                    /* _average_centers */
            if (w[j] > 0) {
                double alpha = 1.0 / w[j];
                for (int d = 0; d < 3; d++) cnew[3 * j + d] *= alpha;
            } else {
                for (int d = 0; d < 3; d++) cnew[3 * j + d] = cnew[3 * am + d];
            }
        }
        for (int j = 0; j < k; j++) {                    /* _center_shift, then center_shift**2 */
            double res = 0.0;
            for (int d = 0; d < 3; d++) {
                double a = cnew[3 * j + d] - centers[3 * j + d];
                res += a * a;
            }
            double sh = sqrt(res);
            shift2[j] = sh * sh;
        }
        double* tmp = centers; centers = cnew; cnew = tmp;
        int same = 1;
        for (int i = 0; i < n; i++) if (labels[i] != labels_old[i]) { same = 0; break; }
        if (same) { strict = 1; break; }
        if (np_pairwise(shift2, k) <= P.tol) break;
        memcpy(labels_old, labels, sizeof(int32_t) * n);
    }
    int n_iter = it < 300 ? it + 1 : 300;
    if (!strict) e_step(&P, centers, labels);
    if (info_out) { info_out[0] = n_iter; info_out[1] = strict; info_out[2] = relocations; }
    free(x); free(P.xsn); free(closest); free(cum); free(dist);
    free(centers); free(cnew); free(w); free(shift2); free(labels_old);
    return 0;
}
