// Floating-point forms of scikit-learn's K-Means distance evaluations (float64, operation for operation),
// shared by the K-Means kernel (rhccq_split.cu) and the MiniBatchKMeans kernel (rhccq_minibatch.cu).
// oracle/kmeans_sklearn.c states where each form comes from and how it was established.
#pragma once
#include "rhccq_common.cuh"

#define RHCCQ_GEMM_P 192               // OpenBLAS DGEMM_DEFAULT_P (SkylakeX): the E step's dgemm is cut into blocks of k
#define RHCCQ_GEMM_UNROLL_M 16
#define RHCCQ_SK_CHUNK 256             // samples per E-step chunk (sklearn/cluster/_k_means_common.pyx:13)

__device__ __forceinline__ double rhccq_sk_norm3(double c0, double c1, double c2) {     // (c0^2 + c2^2) + c1^2
    return __dadd_rn(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c2, c2)), __dmul_rn(c1, c1));
}
__device__ __forceinline__ double rhccq_sk_dot_gemm(double x0, double x1, double x2, double c0, double c1, double c2) {
    return __fma_rn(x2, c2, __fma_rn(x1, c1, __dmul_rn(x0, c0)));
}
__device__ __forceinline__ double rhccq_sk_dot_edge(double x0, double x1, double x2, double c0, double c1, double c2) {
    return __fma_rn(x2, c2, __dadd_rn(__dmul_rn(x0, c0), __dmul_rn(x1, c1)));
}
__device__ __forceinline__ double rhccq_sk_dot_gemv(double x0, double x1, double x2, double c0, double c1, double c2) {
    return __fma_rn(x2, c2, __fma_rn(x0, c0, __dmul_rn(x1, c1)));
}
struct rhccq_sk_pt { double x0, x1, x2; };
__device__ __forceinline__ rhccq_sk_pt rhccq_sk_centred(uint32_t c, const double (&mean)[3]) {
    rhccq_sk_pt p;
    p.x0 = __dsub_rn((double)rhccq_key_r(c), mean[0]);
    p.x1 = __dsub_rn((double)rhccq_key_g(c), mean[1]);
    p.x2 = __dsub_rn((double)rhccq_key_b(c), mean[2]);
    return p;
}
// _euclidean_distances(centre, X, squared=True) of the seeding: max(0, ((-2 x.c) + |c|^2) + |x|^2);
// the first centre's product goes through dgemv, the candidates' through dgemm
__device__ __forceinline__ double rhccq_sk_seed_dist(bool first, const rhccq_sk_pt& p, const double* c) {
    const double dot = first ? rhccq_sk_dot_gemv(p.x0, p.x1, p.x2, c[0], c[1], c[2])
                             : rhccq_sk_dot_gemm(c[0], c[1], c[2], p.x0, p.x1, p.x2);
    double d = __dmul_rn(-2.0, dot);
    d = __dadd_rn(d, rhccq_sk_norm3(c[0], c[1], c[2]));
    d = __dadd_rn(d, rhccq_sk_norm3(p.x0, p.x1, p.x2));
    return d > 0.0 ? d : 0.0;
}
// rows [lo, hi) of Lloyd's dgemm result (cluster index) that the 4-row edge kernel of a block after the
// first computes (OpenBLAS level3.c blocking of M = k); empty for k <= RHCCQ_GEMM_P
__device__ __forceinline__ void rhccq_sk_edge_rows(int k, int& lo, int& hi) {
    lo = hi = 0;
    if (k <= RHCCQ_GEMM_P) return;
    int is = 0, min_i = k;
    for (;;) {
        min_i = k - is;
        if (min_i >= 2 * RHCCQ_GEMM_P) min_i = RHCCQ_GEMM_P;
        else if (min_i > RHCCQ_GEMM_P) min_i = ((min_i / 2 + RHCCQ_GEMM_UNROLL_M - 1) / RHCCQ_GEMM_UNROLL_M) * RHCCQ_GEMM_UNROLL_M;
        if (is + min_i >= k) break;
        is += min_i;
    }
    if (is > 0 && (min_i & 4)) { lo = is + (min_i & ~15) + (min_i & 8); hi = lo + 4; }
}
// Score of Lloyd's E step, |c|^2 - 2 x.c, of sample i (position inside the call) against centre j
__device__ __forceinline__ double rhccq_sk_score(const rhccq_sk_pt& p, const double* c, double csn, bool edge) {
    const double acc = edge ? rhccq_sk_dot_edge(p.x0, p.x1, p.x2, c[0], c[1], c[2])
                            : rhccq_sk_dot_gemm(p.x0, p.x1, p.x2, c[0], c[1], c[2]);
    return __fma_rn(-2.0, acc, csn);
}
__device__ __forceinline__ bool rhccq_sk_edge_sample(int i, int n) {   // inside a group of 12 of its chunk of 256
    const int s = i & ~(RHCCQ_SK_CHUNK - 1), r = i - s;
    const int m = n - s < RHCCQ_SK_CHUNK ? n - s : RHCCQ_SK_CHUNK;
    return r < (m / 12) * 12;
}

