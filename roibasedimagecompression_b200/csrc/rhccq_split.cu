// Cluster -> palette-entry assignment: one CTA per palette ("problem").
//
//   rhccq_k_palette_split    small clusters -> one entry each, large clusters ->
//                            recursive K-Means split
//                            (/root/reference/encoder/compression/clustering.py:253-355, :720-775),
//                            K-Means with the decisions of scikit-learn's float64 arithmetic as
//                            restated in oracle/kmeans_sklearn.c (which cites the lines it follows).
//
// Shape of the work.  A stage-1 problem is ~3 000 colours: one top-level
// K-Means (k = 12..25, ~25 Lloyd iterations) followed by ~10 small K-Means
// calls (k = 2..4 on 100..400 colours) on the leaves that are still larger than
// max_colors_per_cluster, recursively.  The recursion of the reference is
// depth-first, but its RESULT only depends on the tree: a cluster's members sit
// in one contiguous range of a permutation array, a split partitions the range
// stably by K-Means label, and the depth-first leaf order is the order of the
// leaves' ranges.  So the kernel works level by level on a queue of ranges:
// ranges that are large (or need many centres) are split by the whole CTA one
// after another, the many small ones by one warp each, concurrently; leaves are
// flagged at their first position and numbered by one scan at the end.
//
// The working set (20 bytes per colour) lives in shared memory when it fits
// (two CTAs per SM for palettes of up to ~4 000 colours), otherwise in the
// caller's global workspace.  Nothing here touches HBM beyond reading the
// palette and writing one int per row (and, for the rare decisions that must be
// re-evaluated in sequential float64, a scratch area in the global workspace):
// the kernel is bound by FP64 issue and by barrier latency, not by memory bandwidth.
#include "rhccq_common.cuh"
#include "rhccq_sklearn.cuh"
#include "rhccq_kernels.h"

// Optional phase timers (cycles summed over all CTAs into rhccq_split_prof[8]); compiled in with
// -DRHCCQ_SPLIT_PROFILE by tools/split_phases.py only.
#if defined(RHCCQ_SPLIT_PROFILE) && !defined(RHCCQ_HOST_EMU)
__device__ unsigned long long rhccq_split_prof[16];
#define RHCCQ_COUNT(slot, v) atomicAdd(&rhccq_split_prof[slot], (unsigned long long)(v))
#define RHCCQ_PROF_T0() long long prof_t_ = clock64()
#define RHCCQ_PROF(slot) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&rhccq_split_prof[slot], (unsigned long long)(n_ - prof_t_)); prof_t_ = n_; } } while (0)
#else
#define RHCCQ_PROF_T0() do {} while (0)
#define RHCCQ_PROF(slot) do {} while (0)
#define RHCCQ_COUNT(slot, v) do {} while (0)
#endif

#ifndef RHCCQ_SPLIT_THREADS
#define RHCCQ_SPLIT_THREADS 256
#endif
#define RHCCQ_KM_MAXT 12               // 2 + int(log(k)) for k < 22027
#ifndef RHCCQ_PRUNE_MIN_K
#define RHCCQ_PRUNE_MIN_K 10            // centres from which the E step prunes by the triangle inequality
#endif
#define RHCCQ_KM_CAND 16               // ints per candidate table: RHCCQ_KM_MAXT candidates + 2 flags
#define RHCCQ_KC 128                   // centres of a CTA-level K-Means kept in shared memory
#define RHCCQ_KW 8                     // centres of a warp-level K-Means
#ifndef RHCCQ_WARP_RANGE
#define RHCCQ_WARP_RANGE 1024          // largest range a single warp splits
#endif
#define RHCCQ_KPRIV 32                 // k up to which the CTA-level M step uses per-warp histograms
#define RHCCQ_SPLIT_THREADS_BIG 512     // few, large palettes (stage 2): one CTA per SM, twice the warps
#define RHCCQ_SPLIT_MAX_WARPS (RHCCQ_SPLIT_THREADS_BIG / 32)

// ---------------------------------------------------------------- index types
// Small: palettes of up to 32 767 rows (every palette the DBSCAN branch can see:
// the reference switches to MiniBatchKMeans at 10 000 colours).  Bit 15 of a
// permutation entry flags a leaf start; bit 15 of a label marks a point taken
// by the empty-cluster relocation.  Cumulative sums of squared distances fit in
// 32 bits up to 11 008 points (195 075 per point).
struct rhccq_cfg_small {
    typedef uint16_t idx_t;
    typedef uint32_t cum_t;
    typedef uint32_t q_t;
    typedef uint32_t key_t;
    static const uint32_t FLAG = 0x8000u;
    static const int MAX_ROWS = 11008;
    __device__ static __forceinline__ q_t q_pack(int lo, int hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
    __device__ static __forceinline__ int q_lo(q_t v) { return (int)(v & 0xffffu); }
    __device__ static __forceinline__ int q_hi(q_t v) { return (int)(v >> 16); }
    __device__ static __forceinline__ key_t key(int label, int j) { return ((uint32_t)label << 16) | (uint32_t)j; }
    __device__ static __forceinline__ int key_j(key_t v) { return (int)(v & 0xffffu); }
};
struct rhccq_cfg_large {
    typedef uint32_t idx_t;
    typedef unsigned long long cum_t;
    typedef unsigned long long q_t;
    typedef unsigned long long key_t;
    static const uint32_t FLAG = 0x80000000u;
    static const int MAX_ROWS = 0x7fffffff;
    __device__ static __forceinline__ q_t q_pack(int lo, int hi) { return (unsigned long long)(uint32_t)lo | ((unsigned long long)(uint32_t)hi << 32); }
    __device__ static __forceinline__ int q_lo(q_t v) { return (int)(v & 0xffffffffull); }
    __device__ static __forceinline__ int q_hi(q_t v) { return (int)(v >> 32); }
    __device__ static __forceinline__ key_t key(int label, int j) { return ((unsigned long long)(uint32_t)label << 32) | (uint32_t)j; }
    __device__ static __forceinline__ int key_j(key_t v) { return (int)(v & 0xffffffffull); }
};

// ---------------------------------------------------------------- thread groups
// The K-Means below is written once against a "group": the whole CTA or one
// warp.  Every member of the group must call the collective functions.
struct rhccq_grp_cta {
    static const bool kCta = true;
    long long* sll;                    // RHCCQ_MAX_WARPS * RHCCQ_KM_MAXT + 2 elements of shared scratch
    long long* svb;                    // 2 x RHCCQ_MAX_WARPS * RHCCQ_KM_MAXT: sum_vec alternates between the halves
    mutable int svp;                   // which half the next sum_vec uses (same in every thread)
    unsigned long long* csum;          // one slot per thread: chunk sums of the seeding
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int size() const { return (int)blockDim.x; }
    __device__ __forceinline__ int sub() const { return RHCCQ_WARP; }          // warp of the caller inside the group
    __device__ __forceinline__ int nsub() const { return RHCCQ_NWARPS; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int any(int v) const { return rhccq_block_or(v, (int*)sll); }
    // number of true flags over the group (every thread holds 0 or 1): one barrier
    __device__ __forceinline__ int count(int v) const {
#ifdef RHCCQ_HOST_EMU
        return v;
#else
        return __syncthreads_count(v);
#endif
    }
    template <class T> __device__ __forceinline__ T excl_scan(T v, T* total) const {
        return rhccq_block_excl_scan<T>(v, total, reinterpret_cast<T*>(sll));
    }
    __device__ __forceinline__ double max_d(double v) const { return rhccq_block_max<double>(v, reinterpret_cast<double*>(sll)); }
    __device__ __forceinline__ int min_i(int v) const { return rhccq_block_min<int>(v, reinterpret_cast<int*>(sll)); }
    __device__ __forceinline__ int sum_i(int v) const { return rhccq_block_sum<int>(v, reinterpret_cast<int*>(sll)); }
    // v[0..cnt) <- sums over the group (integers: the order of additions is immaterial)
    __device__ __forceinline__ void sum_vec(long long* v, int cnt) const {
#ifndef RHCCQ_HOST_EMU
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
        // one barrier: consecutive calls use different halves, and a half is rewritten only two calls later,
        // after every thread has passed the barrier of the call in between
        long long* buf = svb + (size_t)svp * (RHCCQ_SPLIT_MAX_WARPS * RHCCQ_KM_MAXT);
        svp ^= 1;
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {                   // unrolled: v stays in registers
            if (t < cnt) {
                long long x = v[t];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                if (lane == 0) buf[warp * RHCCQ_KM_MAXT + t] = x;
            }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
            if (t < cnt) {
                long long x = 0;
                for (int w = 0; w < nwarp; ++w) x += buf[w * RHCCQ_KM_MAXT + t];
                v[t] = x;
            }
        }
#endif
    }
    __device__ __forceinline__ double bcast_d(double v) const {   // value of thread 0
#ifndef RHCCQ_HOST_EMU
        double* sd = reinterpret_cast<double*>(sll);
        __syncthreads();
        if (threadIdx.x == 0) sd[0] = v;
        __syncthreads();
        v = sd[0];
#endif
        return v;
    }
};

struct rhccq_grp_warp {
    static const bool kCta = false;
    unsigned long long* csum;          // unused
    __device__ __forceinline__ int tid() const { return RHCCQ_LANE; }
    __device__ __forceinline__ int size() const { return RHCCQ_WARP_SIZE; }
    __device__ __forceinline__ int sub() const { return 0; }
    __device__ __forceinline__ int nsub() const { return 1; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
#ifdef RHCCQ_HOST_EMU
    __device__ __forceinline__ int any(int v) const { return v; }
    template <class T> __device__ __forceinline__ T excl_scan(T v, T* total) const { *total = v; return (T)0; }
    __device__ __forceinline__ double max_d(double v) const { return v; }
    __device__ __forceinline__ int min_i(int v) const { return v; }
    __device__ __forceinline__ int sum_i(int v) const { return v; }
    __device__ __forceinline__ int count(int v) const { return v; }
    __device__ __forceinline__ void sum_vec(long long*, int) const {}
    __device__ __forceinline__ double bcast_d(double v) const { return v; }
#else
    // every collective doubles as a memory barrier among the lanes (like its CTA counterpart)
    __device__ __forceinline__ int any(int v) const { __syncwarp(); return __any_sync(0xffffffffu, v); }
    template <class T> __device__ __forceinline__ T excl_scan(T v, T* total) const {
        const int lane = threadIdx.x & 31;
        __syncwarp();
        T incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        *total = __shfl_sync(0xffffffffu, incl, 31);
        return incl - v;
    }
    __device__ __forceinline__ double max_d(double v) const {
        __syncwarp();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
        return v;
    }
    __device__ __forceinline__ int min_i(int v) const {
        __syncwarp();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, d); v = o < v ? o : v; }
        return v;
    }
    __device__ __forceinline__ int sum_i(int v) const {
        __syncwarp();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        return v;
    }
    __device__ __forceinline__ int count(int v) const { return sum_i(v); }
    __device__ __forceinline__ void sum_vec(long long* v, int cnt) const {
        __syncwarp();
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
            if (t < cnt) {
                long long x = v[t];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                v[t] = x;
            }
        }
    }
    __device__ __forceinline__ double bcast_d(double v) const { __syncwarp(); return __shfl_sync(0xffffffffu, v, 0); }
#endif
};

__device__ __forceinline__ unsigned long long rhccq_warp_incl_scan_u64(unsigned long long v) {
#ifndef RHCCQ_HOST_EMU
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
#endif
    return v;
}

// ---------------------------------------------------------------- K-Means in scikit-learn's arithmetic
//
// The labels must be those of KMeans(k, random_state=42, n_init='auto').fit_predict(colours) of
// scikit-learn 1.9.0 as restated operation by operation in oracle/kmeans_sklearn.c (which cites the
// scikit-learn / numpy / OpenBLAS lines and summation orders it follows).  Every DECISION of the
// algorithm — which candidate a random draw selects, which candidate seeds a centre, which centre a
// colour is nearest to, whether the centre shift is within the tolerance — is taken exactly as that
// arithmetic takes it.  The values the decisions are taken on are sequential float64 sums
// (np.cumsum, the M step in sample order, ddot / dgemv of the seeding), which a GPU cannot evaluate
// quickly; so every decision is first evaluated on an exact-integer / well-conditioned form whose
// distance from the float64 value is bounded, and only a decision that falls inside that bound is
// re-evaluated in the sequential float64 form (by one thread, on scratch in global memory):
//
//   seeding   squared distances of uint8 colours are integers; the float64 distances of the centred
//             data differ from them by < 1e-9 each.  Candidate = first index whose cumulative sum
//             reaches r * pot: decided on the integer sums unless r * pot lies within
//             2 n (1e-9 + ulp(pot)) of a cumulative sum.  Best candidate = smallest potential: decided
//             on the integer potentials unless two different colours tie.
//   E step    centres from exact integer sums, c = S / n - mean, are within eps_c of the float64
//             centres (sequential sum of centred colours times 1/n); a colour whose two best scores
//             are closer than the margin that eps_c allows is re-evaluated on the float64 centres.
//   tolerance decided on the integer-derived values unless |shift - tol| is inside their error band.
//
// The bounds are derived next to each use; RHCCQ_KM_FORCE_EXACT (test builds) makes every decision
// take the float64 path, and tests/ runs both builds against the oracle.
#define RHCCQ_NBMAX_T 2048             // row block of OpenBLAS dgemv_t
#define RHCCQ_EX_PER_ROW 10            // doubles of global scratch per palette row (see rhccq_km_arrays::ex)

__device__ __forceinline__ int rhccq_kmeans_local_trials(int k) {   // 2 + int(log(k)), sklearn/_kmeans.py:226
    const int e[] = {3, 8, 21, 55, 149, 404, 1097, 2981, 8104, 22027, 59875, 162755, 442414, 1202605};
    int t = 2;
    for (int i = 0; i < 14; ++i) if (k >= e[i]) ++t;
    return t;
}

// Per-position arrays of a problem (a K-Means call on the range [lo, lo + n) only touches
// positions of its range, so concurrent calls on disjoint ranges do not interfere) ...
template <class Cfg> struct rhccq_km_arrays {
    uint32_t* x;                        // colour at every position
    uint32_t* closest;                  // seeding: squared distance to the nearest chosen centre
    typename Cfg::cum_t* cum;           // seeding: inclusive cumulative sum of closest over the range;
                                        // Lloyd: second label array and the list of undecided positions
    typename Cfg::idx_t* label;         // Lloyd: cluster of every position
    double* ex;                         // global scratch of the float64 re-evaluations, RHCCQ_EX_PER_ROW per position
};
// ... and the per-call centre tables (k entries, private to the calling group).
struct rhccq_km_centers {
    double* center;                     // [3k]  centred coordinates
    double* center_new;                 // [3k]
    double* term;                       // [k]
    double* csn;                        // [k]   |c|^2 as einsum('ij,ij->i') evaluates it
    float* cf;                          // [4k]  (-2 c0, -2 c1, -2 c2, |c|^2) in float32: first level of the E step
    int* hb;                            // [k] float32 pruning threshold of every centre; nullptr: every point is evaluated
    int* sums;                          // [3k]
    int* cnt;                           // [k]
    int* hist;                          // unused
    int* cand;                          // [RHCCQ_KM_MAXT + 2]: candidates, then two flag / counter slots
    int* poff;                          // [k * nsub] offsets of the counting partition, or nullptr (sort instead)
    int* wl;                            // worklist counter of the pruned E step, or nullptr (evaluate every centre)
};

// ---- single-thread float64 kernels of the libraries underneath scikit-learn (oracle/kmeans_sklearn.c)
// OpenBLAS ddot(v, ones)
__device__ __noinline__ double rhccq_sk_ddot_ones(const double* x, int n) {
    double dot = 0.0;
    const int n1 = n & -16;
    int i = 0;
    if (n1) {
        double a[4][8];
        for (int q = 0; q < 4; ++q) for (int l = 0; l < 8; ++l) a[q][l] = 0.0;
        const int n32 = n1 & ~31;
        for (; i < n32; i += 32)
            for (int q = 0; q < 4; ++q) for (int l = 0; l < 8; ++l) a[q][l] = __dadd_rn(a[q][l], x[i + 8 * q + l]);
        double b[4][4];
        for (int q = 0; q < 4; ++q) for (int l = 0; l < 4; ++l) b[q][l] = __dadd_rn(a[q][l], a[q][l + 4]);
        for (; i < n1; i += 16)
            for (int q = 0; q < 4; ++q) for (int l = 0; l < 4; ++l) b[q][l] = __dadd_rn(b[q][l], x[i + 4 * q + l]);
        double c[4];
        for (int l = 0; l < 4; ++l) c[l] = __dadd_rn(__dadd_rn(__dadd_rn(b[0][l], b[1][l]), b[2][l]), b[3][l]);
        dot = __dadd_rn(__dadd_rn(c[0], c[2]), __dadd_rn(c[1], c[3]));
    }
    for (; i < n; ++i) dot = __dadd_rn(dot, x[i]);
    return dot;
}
// OpenBLAS dgemv_t(D, ones), one column of length m: element i is min(f[i], distance of i to `cand`)
// (cand == nullptr: f[i] itself); lanes = 4 or 2 (see rhccq_sk_gemv_lanes)
struct rhccq_sk_col {
    const double* f; const uint32_t* x; const double* cand; const double* mean;
    __device__ __forceinline__ double operator()(int i) const {
        const double a = f[i];
        if (cand == nullptr) return a;
        const double m3[3] = {mean[0], mean[1], mean[2]};
        const double d = rhccq_sk_seed_dist(false, rhccq_sk_centred(x[i], m3), cand);
        return a < d ? a : d;
    }
};
__device__ __noinline__ double rhccq_sk_dgemv_t_ones(const rhccq_sk_col& v, int m, int lanes) {
    double y = 0.0;
    const int m3 = m & 3, m2 = (m & (RHCCQ_NBMAX_T - 1)) - m3;
    int m1 = m & -4, nb = RHCCQ_NBMAX_T, at = 0;
    while (nb == RHCCQ_NBMAX_T) {
        m1 -= nb;
        if (m1 < 0) { if (m2 == 0) break; nb = m2; }
        double blk;
        if (lanes == 4) {
            double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
            for (int i = 0; i < nb; i += 4) {
                l0 = __dadd_rn(l0, v(at + i)); l1 = __dadd_rn(l1, v(at + i + 1));
                l2 = __dadd_rn(l2, v(at + i + 2)); l3 = __dadd_rn(l3, v(at + i + 3));
            }
            blk = __dadd_rn(__dadd_rn(l0, l2), __dadd_rn(l1, l3));
        } else {
            double l0 = 0, l1 = 0;
            for (int i = 0; i < nb; i += 2) { l0 = __dadd_rn(l0, v(at + i)); l1 = __dadd_rn(l1, v(at + i + 1)); }
            blk = __dadd_rn(l0, l1);
        }
        y = __dadd_rn(y, blk);
        at += nb;
    }
    if (m3 == 3) y = __dadd_rn(y, __dadd_rn(__dadd_rn(v(at), v(at + 1)), v(at + 2)));
    else if (m3 == 2) y = __dadd_rn(y, __dadd_rn(v(at), v(at + 1)));
    else if (m3 == 1) y = __dadd_rn(y, v(at));
    return y;
}
__device__ __forceinline__ int rhccq_sk_gemv_lanes(int j, int t) {   // column groups of 4, a pair if (t & 2), a single
    const int g = (t >> 2) * 4;
    if (j < g) return 4;
    return ((t & 2) && j - g < 2) ? 2 : 4;
}
// numpy's pairwise summation of a contiguous vector
__device__ __noinline__ double rhccq_sk_np_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(rhccq_sk_np_sum(a, n2), rhccq_sk_np_sum(a + n2, n - n2));
}
// Float64 centres of the clusters of `lab` (M step in sample order, _average_centers): E[4j .. 4j+2]
// centre, E[4j+3] = |c|^2.  One thread.
template <class Cfg>
__device__ __noinline__ void rhccq_sk_all_centers(const uint32_t* x, const typename Cfg::idx_t* lab, int n, int k,
                                                  const double* mean, double* E) {
    const double m3[3] = {mean[0], mean[1], mean[2]};
    for (int q = 0; q < 4 * k; ++q) E[q] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int l = (int)(lab[i] & (typename Cfg::idx_t)~Cfg::FLAG);
        const rhccq_sk_pt p = rhccq_sk_centred(x[i], m3);
        E[4 * l] = __dadd_rn(E[4 * l], p.x0); E[4 * l + 1] = __dadd_rn(E[4 * l + 1], p.x1);
        E[4 * l + 2] = __dadd_rn(E[4 * l + 2], p.x2); E[4 * l + 3] = __dadd_rn(E[4 * l + 3], 1.0);
    }
}
// _average_centers on a table of (sum0, sum1, sum2, weight) rows, in place; then weight <- |c|^2
__device__ __noinline__ void rhccq_sk_average(double* E, int k) {
    int am = 0;
    for (int j = 1; j < k; ++j) if (E[4 * j + 3] > E[4 * am + 3]) am = j;
    for (int j = 0; j < k; ++j) {
        if (E[4 * j + 3] > 0.0) {
            const double alpha = __ddiv_rn(1.0, E[4 * j + 3]);
            for (int d = 0; d < 3; ++d) E[4 * j + d] = __dmul_rn(E[4 * j + d], alpha);
        } else {
            for (int d = 0; d < 3; ++d) E[4 * j + d] = E[4 * am + d];
        }
    }
    for (int j = 0; j < k; ++j) E[4 * j + 3] = rhccq_sk_norm3(E[4 * j], E[4 * j + 1], E[4 * j + 2]);
}

// First level of the E step: the scores |c|^2 - 2 x.c of RHCCQ_EB points against every centre in float32
// (every centre is one 16-byte load per RHCCQ_EB points; the chains are independent and branch-free).
// Error of one float32 score against the float64 score on the same centres: inputs x - mean and -2 c
// rounded to float32 (<= 3.1e-5 each, |x| <= 255, |2c| <= 510), |c|^2 (<= ulp(2e5)/2 = 0.008) and three
// fused steps (<= 0.016 each): <= 0.13.  und[u] is set when the runner-up is within RHCCQ_F32_MARGIN = 0.5
// of the best; such a point goes to the second level: the float64 scores on the integer-derived centres, one warp per
// point, and from there — when two scores are within `margin` — to the third: the float64 centres themselves.
#define RHCCQ_EB 4
#define RHCCQ_F32_MARGIN 0.5f
template <int NB>
__device__ __forceinline__ void rhccq_nearest_centers_f32(const uint32_t (&c)[NB], const float (&meanf)[3],
                                                          const float* cf, int k, int (&bi)[NB], bool (&und)[NB]) {
    float x0[NB], x1[NB], x2[NB], best[NB], sec[NB];
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    {
        const float4 q0 = c4[0];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            x0[u] = __fsub_rn((float)rhccq_key_r(c[u]), meanf[0]);
            x1[u] = __fsub_rn((float)rhccq_key_g(c[u]), meanf[1]);
            x2[u] = __fsub_rn((float)rhccq_key_b(c[u]), meanf[2]);
            best[u] = fmaf(x0[u], q0.x, fmaf(x1[u], q0.y, fmaf(x2[u], q0.z, q0.w)));
            sec[u] = 3.0e38f;
            bi[u] = 0;
        }
    }
#pragma unroll 2
    for (int q = 1; q < k; ++q) {
        const float4 cq = c4[q];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const float d = fmaf(x0[u], cq.x, fmaf(x1[u], cq.y, fmaf(x2[u], cq.z, cq.w)));
            sec[u] = fminf(sec[u], fmaxf(d, best[u]));              // runner-up so far
            const bool lt = d < best[u];
            best[u] = lt ? d : best[u];
            bi[u] = lt ? q : bi[u];
        }
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) und[u] = !(sec[u] - best[u] > RHCCQ_F32_MARGIN);
}
// the same decision on the float64 centres E (one point, every centre, the dgemm edge rows included)
__device__ __noinline__ int rhccq_nearest_exact(uint32_t c, const double (&mean)[3], const double* E, int k,
                                                   int i, int n, int e_lo, int e_hi) {
    const rhccq_sk_pt p = rhccq_sk_centred(c, mean);
    const bool es = e_hi > e_lo && rhccq_sk_edge_sample(i, n);
    double best = rhccq_sk_score(p, E, E[3], es && 0 >= e_lo && 0 < e_hi);
    int bi = 0;
    for (int q = 1; q < k; ++q) {
        const double d = rhccq_sk_score(p, E + 4 * q, E[4 * q + 3], es && q >= e_lo && q < e_hi);
        if (d < best) { best = d; bi = q; }
    }
    return bi;
}

#ifdef RHCCQ_KM_FORCE_EXACT
#define RHCCQ_KM_FORCED 1
#else
#define RHCCQ_KM_FORCED 0
#endif

// ---- the float64 re-evaluations: rare, kept out of line (and unrolled nowhere) so that they cost no
// instruction-cache space next to the hot loops
// float64 closest distances of positions [j0, j1) to the c centres chosen so far (every thread, its chunk)
__device__ __noinline__ void rhccq_sk_closest_f(const uint32_t* x, int j0, int j1, const double* mean, const double* cen,
                                                int c, double* f) {
    const double m3[3] = {mean[0], mean[1], mean[2]};
#pragma unroll 1
    for (int j = j0; j < j1; ++j) {
        const rhccq_sk_pt p = rhccq_sk_centred(x[j], m3);
        double v = rhccq_sk_seed_dist(true, p, cen);
#pragma unroll 1
        for (int s = 1; s < c; ++s) { const double d = rhccq_sk_seed_dist(false, p, cen + 3 * s); v = d < v ? d : v; }
        f[j] = v;
    }
}
// candidates of one seeding step on the float64 values: the potential as the previous step's BLAS call summed
// it, np.cumsum, np.searchsorted(side='left'), clipped (one thread)
__device__ __noinline__ void rhccq_sk_draw(const uint32_t* x, int n, const double* mean, const double* f, int prev_slot,
                                           int T, const double* rng_step, int* cand) {
    rhccq_sk_col col; col.f = f; col.x = x; col.cand = nullptr; col.mean = mean;
    const double pot_f = prev_slot < 0 ? rhccq_sk_ddot_ones(f, n)
                                       : rhccq_sk_dgemv_t_ones(col, n, rhccq_sk_gemv_lanes(prev_slot, T));
    double rv[RHCCQ_KM_MAXT];
    int fnd[RHCCQ_KM_MAXT];
#pragma unroll 1
    for (int t = 0; t < T; ++t) { rv[t] = __dmul_rn(rng_step[t], pot_f); fnd[t] = -1; }
    double s = 0.0;
    int open = T;
#pragma unroll 1
    for (int i = 0; i < n && open > 0; ++i) {
        s = __dadd_rn(s, f[i]);
#pragma unroll 1
        for (int t = 0; t < T; ++t) if (fnd[t] < 0 && !(s < rv[t])) { fnd[t] = i; --open; }
    }
#pragma unroll 1
    for (int t = 0; t < T; ++t) cand[t] = fnd[t] < 0 ? n - 1 : fnd[t];
}
// np.argmin of the float64 potentials over the candidate slots of `mask` (one thread)
__device__ __noinline__ int rhccq_sk_best(const uint32_t* x, int n, const double* mean, const double* f, int T,
                                          const int* cand, unsigned mask) {
    const double m3[3] = {mean[0], mean[1], mean[2]};
    double bp = 0.0;
    int bs = -1;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const rhccq_sk_pt pc = rhccq_sk_centred(x[cand[t]], m3);
        const double cc[3] = {pc.x0, pc.x1, pc.x2};
        rhccq_sk_col col; col.f = f; col.x = x; col.cand = cc; col.mean = mean;
        const double pf = rhccq_sk_dgemv_t_ones(col, n, rhccq_sk_gemv_lanes(t, T));
        if (bs < 0 || pf < bp) { bp = pf; bs = t; }
    }
    return bs;
}
// float64 centres of an E step (one thread): the table itself when it holds them, else the M step of `lab`
template <class Cfg>
__device__ __noinline__ void rhccq_sk_e_old(const uint32_t* x, const typename Cfg::idx_t* lab, int n, int k, const double* mean,
                                            const double* cen, const double* csn, bool cen_exact, double* E) {
    if (cen_exact) {
#pragma unroll 1
        for (int q = 0; q < k; ++q) { E[4 * q] = cen[3 * q]; E[4 * q + 1] = cen[3 * q + 1]; E[4 * q + 2] = cen[3 * q + 2]; E[4 * q + 3] = csn[q]; }
    } else {
        rhccq_sk_all_centers<Cfg>(x, lab, n, k, mean, E);
        rhccq_sk_average(E, k);
    }
}
// center_shift_tot <= tol in float64: _center_shift, (shift ** 2).sum(), np.var (one thread)
__device__ __noinline__ int rhccq_sk_converged(const uint32_t* x, int n, int k, const double* mean, const double* E_old,
                                               const double* E_new, double* s2) {
#pragma unroll 1
    for (int q = 0; q < k; ++q) {
        double res = 0.0;
#pragma unroll 1
        for (int d = 0; d < 3; ++d) { const double a = __dsub_rn(E_new[4 * q + d], E_old[4 * q + d]); res = __dadd_rn(res, __dmul_rn(a, a)); }
        const double sh = __dsqrt_rn(res);
        s2[q] = __dmul_rn(sh, sh);
    }
    double var[3];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        double s = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const double xv = __dsub_rn((double)((x[i] >> (16 - 8 * d)) & 255u), mean[d]);
            s = __dadd_rn(s, __dmul_rn(xv, xv));
        }
        var[d] = __ddiv_rn(s, (double)n);
    }
    const double tol_e = __dmul_rn(__ddiv_rn(__dadd_rn(__dadd_rn(var[0], var[1]), var[2]), 3.0), 1e-4);
    return rhccq_sk_np_sum(s2, k) <= tol_e ? 1 : 0;
}
// Empty clusters (_k_means_common.pyx:167-211), one thread.  E_new <- float64 M step of `lab`; the points
// farthest from their (old) centres, farthest first, ties to the higher index, give their colour to the empty
// clusters in ascending order and the donor's sum loses it by one float64 subtraction; then _average_centers.
template <class Cfg>
__device__ __noinline__ void rhccq_sk_relocate(const uint32_t* x, typename Cfg::idx_t* lab, int n, int k, int n_empty,
                                               const double* mean, const double* E_old, double* E_new, double* s2) {
    typedef typename Cfg::idx_t idx_t;
    const double m3[3] = {mean[0], mean[1], mean[2]};
    rhccq_sk_all_centers<Cfg>(x, lab, n, k, mean, E_new);
#pragma unroll 1
    for (int q = 0; q < k; ++q) s2[q] = E_new[4 * q + 3] == 0.0 ? 1.0 : 0.0;       // the empty set is fixed first
    int e = 0;
#pragma unroll 1
    for (int done = 0; done < n_empty; ++done) {
        while (s2[e] == 0.0) ++e;
        double bm = -1.0;
        int bj = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            if (lab[i] & Cfg::FLAG) continue;                                  // already taken
            const rhccq_sk_pt p = rhccq_sk_centred(x[i], m3);
            const double* ce = E_old + 4 * (int)lab[i];
            const double a0 = __dsub_rn(p.x0, ce[0]), a1 = __dsub_rn(p.x1, ce[1]), a2 = __dsub_rn(p.x2, ce[2]);
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
            if (d >= bm) { bm = d; bj = i; }                                   // ascending i: the last maximum stays
        }
        if (done == 0 && bm == 0.0) break;                                     // np.max(distances) == 0: nothing to relocate
        const int o = (int)lab[bj];
        const rhccq_sk_pt p = rhccq_sk_centred(x[bj], m3);
        E_new[4 * o] = __dsub_rn(E_new[4 * o], p.x0); E_new[4 * o + 1] = __dsub_rn(E_new[4 * o + 1], p.x1);
        E_new[4 * o + 2] = __dsub_rn(E_new[4 * o + 2], p.x2);
        E_new[4 * o + 3] = __dsub_rn(E_new[4 * o + 3], 1.0);
        E_new[4 * e] = p.x0; E_new[4 * e + 1] = p.x1; E_new[4 * e + 2] = p.x2; E_new[4 * e + 3] = 1.0;
        lab[bj] = (idx_t)(o | Cfg::FLAG);
        ++e;
    }
#pragma unroll 1
    for (int i = 0; i < n; ++i) lab[i] = (idx_t)(lab[i] & (idx_t)~Cfg::FLAG);
    rhccq_sk_average(E_new, k);
}

// Labels of KMeans(k, random_state=42, n_init='auto').fit_predict on the n colours at positions
// [lo, lo + n).  On return A.label holds the labels and C.cnt the cluster sizes.  Group-uniform control
// flow; every thread of the group must call.
template <class G, class Cfg>
__device__ __forceinline__ void rhccq_kmeans(const G& g, const rhccq_km_arrays<Cfg>& A, const rhccq_km_centers& C, int lo, int n, int k,
                             const double* __restrict__ rng) {
    typedef typename Cfg::cum_t cum_t;
    typedef typename Cfg::idx_t idx_t;
    const int tid = g.tid(), gsz = g.size();
    const uint32_t* x = A.x + lo;
    uint32_t* closest = A.closest + lo;
    cum_t* cum = A.cum + lo;
    double* ex = A.ex + (size_t)RHCCQ_EX_PER_ROW * lo;              // float64 scratch of this call
    int* flag = C.cand + RHCCQ_KM_MAXT;                             // two group-shared ints
    // contiguous chunk of the caller (scans need a fixed element order)
    const int per = (n + gsz - 1) / gsz;
    const int c_lo = tid * per < n ? tid * per : n;
    const int c_hi = c_lo + per < n ? c_lo + per : n;

    RHCCQ_PROF_T0();
    // ---- k-means++ seeding (sklearn/cluster/_kmeans.py:216-282)
    const int T = rhccq_kmeans_local_trials(k);
    // RandomState.choice(n, p=uniform) walks the normalised cumulative sum of 1/n: floor(u n) for this u
    // and every n up to 20 000 (tests/test_oracle_golden.py checks the whole range against numpy's form)
    int first = (int)__dmul_rn(rng[0], (double)n);
    if (first > n - 1) first = n - 1;
    long long acc[RHCCQ_KM_MAXT];
    cum_t chunk_sum = 0;
    {
        const uint32_t cf = x[first];
        long long p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
        for (int j = c_lo; j < c_hi; ++j) {
            const uint32_t c = x[j];
            const uint32_t d = (uint32_t)rhccq_d2(c, cf);
            closest[j] = d;
            chunk_sum += d;
            const long long r = rhccq_key_r(c), gg = rhccq_key_g(c), b = rhccq_key_b(c);
            p1[0] += r; p1[1] += gg; p1[2] += b;
            p2[0] += r * r; p2[1] += gg * gg; p2[2] += b * b;
        }
        acc[0] = (long long)chunk_sum;
        for (int d = 0; d < 3; ++d) { acc[1 + d] = p1[d]; acc[4 + d] = p2[d]; }
        g.sum_vec(acc, 7);
    }
    long long pot = acc[0];
    const long long S1[3] = {acc[1], acc[2], acc[3]}, S2[3] = {acc[4], acc[5], acc[6]};
    // X.mean(axis=0): the sums are exact integers, one rounding in the division
    const double mean[3] = {__ddiv_rn((double)S1[0], (double)n), __ddiv_rn((double)S1[1], (double)n),
                            __ddiv_rn((double)S1[2], (double)n)};
    const float meanf[3] = {(float)mean[0], (float)mean[1], (float)mean[2]};
    for (int q = tid; q < 3; q += gsz)
        C.center[q] = __dsub_rn((double)((x[first] >> (16 - 8 * q)) & 255u), mean[q]);
    int ri = 1, prev_slot = -1;
    // The distances to the centre chosen last are folded into `closest` lazily — by the next pass over the points
    // (the potentials below; the cumulative sum of the warp-level form) and by the candidate search on the
    // elements it looks at — instead of by a pass of their own: `pend` is that centre, pending the fold.
    bool pending = false;
    uint32_t pend = 0u;
    for (int c = 1; c < k; ++c) {
        // ---- candidates: searchsorted(cumsum(closest), r * pot), clipped (:247-252)
        // |float64 cumulative sum - integer one| <= n (1e-9 + ulp(pot)/2) and the same for r * pot, so the
        // integer search decides unless r * pot is within und_a of one of the two sums it falls between
        const double und_a = RHCCQ_KM_FORCED ? 1.0e300
                                             : __dmul_rn(2.0 * (double)n, __dadd_rn(1.0e-9, __dmul_rn((double)pot, 2.3e-16)));
        if (tid == 0) flag[0] = 0;
        if (G::kCta) {
            // Without materialising the cumulative sum: every thread publishes the sum of its chunk; one warp
            // per trial scans the chunk sums (shuffles), finds the chunk in which the inclusive sum first
            // reaches r * pot, then the element inside it.
            g.csum[tid] = (unsigned long long)chunk_sum;
            g.sync();
            const int E = (gsz + RHCCQ_WARP_SIZE - 1) / RHCCQ_WARP_SIZE;     // chunk sums per lane
            for (int t = RHCCQ_WARP; t < T; t += RHCCQ_NWARPS) {
                const double rv = __dmul_rn(rng[ri + t], (double)pot);
                unsigned long long loc = 0;
                for (int e = 0; e < E; ++e) { const int ix = RHCCQ_LANE * E + e; if (ix < gsz) loc += g.csum[ix]; }
                const unsigned long long incl = rhccq_warp_incl_scan_u64(loc);
                const unsigned m = rhccq_ballot((double)incl >= rv);
                int found = n - 1;                                  // beyond the total: clipped
                int und = 1;
                if (m != 0u) {
                    const int L = __ffs((int)m) - 1;
                    unsigned long long pre = rhccq_shfl(incl - loc, L);
                    int cix = L * E;
                    for (int e = 0; e < E; ++e) {                   // the chunk inside lane L's share (warp-uniform walk)
                        const int ix = L * E + e;
                        if (ix >= gsz) break;
                        const unsigned long long nx = pre + g.csum[ix];
                        cix = ix;
                        if ((double)nx >= rv) break;
                        pre = nx;
                    }
                    const int j0 = cix * per < n ? cix * per : n, j1 = j0 + per < n ? j0 + per : n;
                    for (int base = j0; base < j1; base += RHCCQ_WARP_SIZE) {
                        const int j = base + RHCCQ_LANE;
                        unsigned long long own = 0ull;
                        if (j < j1) {
                            uint32_t o = closest[j];
                            if (pending) { const uint32_t d = (uint32_t)rhccq_d2(x[j], pend); o = d < o ? d : o; }
                            own = o;
                        }
                        const unsigned long long inc = rhccq_warp_incl_scan_u64(own);
                        const unsigned mm = rhccq_ballot(j < j1 && (double)(pre + inc) >= rv);
                        if (mm != 0u) {
                            const int w = __ffs((int)mm) - 1;
                            found = base + w;
                            const unsigned long long at = pre + rhccq_shfl(inc, w), below = at - rhccq_shfl(own, w);
                            und = ((double)at - rv < und_a) || (rv - (double)below <= und_a);
                            break;
                        }
                        pre += rhccq_shfl(inc, RHCCQ_WARP_SIZE - 1);
                    }
                }
                if (RHCCQ_LANE == 0) {
                    C.cand[t] = found < n - 1 ? found : n - 1;
                    if (und) atomicOr(&flag[0], 1);
                }
            }
            g.sync();
        } else {
            // inclusive cumulative sum of closest (exact integers)
            cum_t total;
            cum_t run = g.template excl_scan<cum_t>(chunk_sum, &total);
            for (int j = c_lo; j < c_hi; ++j) {
                uint32_t o = closest[j];
                if (pending) { const uint32_t d = (uint32_t)rhccq_d2(x[j], pend); o = d < o ? d : o; closest[j] = o; }
                run += o; cum[j] = run;
            }
            g.sync();
            for (int t = tid; t < T; t += gsz) {
                const double rv = __dmul_rn(rng[ri + t], (double)pot);
                int l = 0, h = n;                                   // first j with cum[j] >= rv
                while (l < h) {
                    const int mid = (l + h) >> 1;
                    if ((double)cum[mid] < rv) l = mid + 1; else h = mid;
                }
                int und = 1;
                if (l < n) {
                    const double at = (double)cum[l], below = l > 0 ? (double)cum[l - 1] : 0.0;
                    und = (at - rv < und_a) || (rv - below <= und_a);
                }
                C.cand[t] = l < n - 1 ? l : n - 1;
                if (und) atomicOr(&flag[0], 1);
            }
            g.sync();
        }
        bool have_f = false;
        if (flag[0]) {
            // float64 re-evaluation of the draws
            if (tid == 0) RHCCQ_COUNT(12, 1);
            rhccq_sk_closest_f(x, c_lo, c_hi, mean, C.center, c, ex);
            have_f = true;
            g.sync();
            if (tid == 0) rhccq_sk_draw(x, n, mean, ex, prev_slot, T, rng + ri, C.cand);
            g.sync();
        }
        ri += T;
        // ---- potentials of the candidates (:255-261), integer form
        uint32_t xc[RHCCQ_KM_MAXT];
        int cj[RHCCQ_KM_MAXT];
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) { acc[t] = 0; cj[t] = t < T ? C.cand[t] : 0; xc[t] = t < T ? x[cj[t]] : 0u; }
        cum_t part[RHCCQ_KM_MAXT];                                  // this thread's share of every candidate's potential
        if (per <= 20000) {                                         // 20 000 x 195 075 < 2^32: 32-bit partial sums
            uint32_t a32[RHCCQ_KM_MAXT];
#pragma unroll
            for (int t = 0; t < RHCCQ_KM_MAXT; ++t) a32[t] = 0u;
            const bool fold = pending && G::kCta;
            for (int j = c_lo; j < c_hi; ++j) {
                const uint32_t cjx = x[j];
                uint32_t o = closest[j];
                if (fold) { const uint32_t d = (uint32_t)rhccq_d2(cjx, pend); o = d < o ? d : o; closest[j] = o; }
#pragma unroll
                for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
                    if (t >= T) break;                              // uniform: no predicated-off slots are issued
                    const uint32_t d = (uint32_t)rhccq_d2(cjx, xc[t]);
                    a32[t] += d < o ? d : o;
                }
            }
#pragma unroll
            for (int t = 0; t < RHCCQ_KM_MAXT; ++t) { acc[t] = (long long)a32[t]; part[t] = (cum_t)a32[t]; }
        } else {
            const bool fold = pending && G::kCta;
            for (int j = c_lo; j < c_hi; ++j) {
                const uint32_t cjx = x[j];
                uint32_t o = closest[j];
                if (fold) { const uint32_t d = (uint32_t)rhccq_d2(cjx, pend); o = d < o ? d : o; closest[j] = o; }
#pragma unroll
                for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
                    if (t >= T) break;
                    const uint32_t d = (uint32_t)rhccq_d2(cjx, xc[t]);
                    acc[t] += d < o ? d : o;
                }
            }
#pragma unroll
            for (int t = 0; t < RHCCQ_KM_MAXT; ++t) part[t] = (cum_t)acc[t];
        }
        g.sum_vec(acc, T);
        int best = 0;
        unsigned tied = 0u;
        long long best_pot = acc[0];
#pragma unroll
        for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t < T && acc[t] < best_pot) { best_pot = acc[t]; best = t; }
        // np.argmin of the float64 potentials (:264): they are within n (1e-9 + ulp) << 1/2 of the integers, so
        // only candidates that tie in the integers — and are different colours — need the float64 sums
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t)
            if (t < T && (acc[t] == best_pot || RHCCQ_KM_FORCED)) tied |= 1u << t;
        bool tie = false;
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) if (((tied >> t) & 1u) && cj[t] != cj[best]) tie = true;
        if (tie) {
            if (tid == 0) RHCCQ_COUNT(13, 1);
            if (!have_f) rhccq_sk_closest_f(x, c_lo, c_hi, mean, C.center, c, ex);
            g.sync();
            if (tid == 0) flag[1] = rhccq_sk_best(x, n, mean, ex, T, C.cand, tied);
            g.sync();
            best = flag[1];
            best_pot = acc[0];
#pragma unroll
            for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t == best) best_pot = acc[t];
            g.sync();                                               // flag[1] is rewritten in a later step
        }
        uint32_t cs = xc[0];
#pragma unroll
        for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t == best) cs = xc[t];
        chunk_sum = part[0];                                        // sum over the chunk of min(closest, distance to the new centre)
#pragma unroll
        for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t == best) chunk_sum = part[t];
        pending = true;
        pend = cs;
        for (int q = tid; q < 3; q += gsz) C.center[3 * c + q] = __dsub_rn((double)((cs >> (16 - 8 * q)) & 255u), mean[q]);
        pot = best_pot;
        prev_slot = best;
        // the next pass writes cum / cand / flag only after a collective, which orders it after the reads above
    }

    if (g.size() > RHCCQ_WARP_SIZE) RHCCQ_PROF(1);
    // ---- tolerance: mean(var(X, axis=0)) * 1e-4 (sklearn/_kmeans.py:285-293).  The exact variances differ
    // from numpy's sequential float64 ones by < n 2^-52 relative; the decision below re-evaluates inside a band.
    double tol;
    {
        const double nn = __dmul_rn((double)n, (double)n);
        double v[3];
        for (int d = 0; d < 3; ++d) v[d] = __ddiv_rn((double)((long long)n * S2[d] - S1[d] * S1[d]), nn);
        tol = __dmul_rn(__ddiv_rn(__dadd_rn(__dadd_rn(v[0], v[1]), v[2]), 3.0), 1e-4);
    }
    // centre error of the integer form against the float64 M step: the sequential sum of n_j centred colours
    // (|x| <= 255, partial sums <= 255 n_j) rounds by <= n_j ulp(255 n_j) / 2 <= 2.9e-14 n_j^2, divided by n_j
    const double eps_c = __dmul_rn(4.0e-14, (double)(n + 8));
    // two scores (|c|^2 - 2 x.c, gradient 2 |c - x|_1 <= 1530) on centres eps_c off, plus their own rounding
    const double margin_full = RHCCQ_KM_FORCED ? 1.0e300 : __dadd_rn(__dmul_rn(4000.0, eps_c), 1.0e-9);
    int e_lo, e_hi;
    rhccq_sk_edge_rows(k, e_lo, e_hi);

    // ---- Lloyd (sklearn/_kmeans.py:703-757)
    //
    // The cluster sums are kept incrementally as exact integers: a point only touches them when its label
    // changes.  Two label arrays alternate so that the labels the current centres were averaged from stay
    // readable: the float64 centres are recomputed from them when a decision needs them.
    const idx_t NOLABEL = (idx_t)(Cfg::FLAG - 1u);                  // never a real label (k < FLAG - 1)
    idx_t* lab_prev = A.label + lo;                                 // labels the centres `cen` come from
    idx_t* lab_cur = reinterpret_cast<idx_t*>(cum);                 // labels the E step writes
    idx_t* wl = reinterpret_cast<idx_t*>(cum) + n;                  // positions the E step evaluates; Cfg::FLAG: decision open
    bool prune_valid = false;                                       // C.hb holds the pruning thresholds of the centres `cen`
    for (int j = tid; j < n; j += gsz) lab_prev[j] = NOLABEL;
    for (int q = tid; q < k; q += gsz) {
        C.cnt[q] = 0; C.sums[3 * q] = 0; C.sums[3 * q + 1] = 0; C.sums[3 * q + 2] = 0;
        C.csn[q] = rhccq_sk_norm3(C.center[3 * q], C.center[3 * q + 1], C.center[3 * q + 2]);
        C.cf[4 * q] = (float)__dmul_rn(-2.0, C.center[3 * q]); C.cf[4 * q + 1] = (float)__dmul_rn(-2.0, C.center[3 * q + 1]);
        C.cf[4 * q + 2] = (float)__dmul_rn(-2.0, C.center[3 * q + 2]); C.cf[4 * q + 3] = (float)C.csn[q];
    }
    double* cen = C.center;                                         // current / next centre tables, swapped per iteration
    double* cen_new = C.center_new;
    double* E_old = ex;                                             // float64 centres of the E step [4k]
    double* E_new = ex + 4 * (size_t)k;                             // float64 centres after the M step [4k]
    double* s2 = ex + 8 * (size_t)k;                                // center_shift ** 2 [k]
    bool recount = true;                                            // add every point to the sums, subtract none
    bool cen_exact = true;                                          // `cen` holds the float64 centres themselves (seeds; after a relocation)
    bool final_pass = false;                                        // the closing E step (:737-749): labels only
    int changed = 0;

    // one point's new label: label arrays, change flag, integer sums
    auto commit = [&](int j, uint32_t c, int bi) {
        const int old = (int)lab_prev[j];
        lab_cur[j] = (idx_t)bi;
        if (final_pass) return;
        if (old != bi) changed = 1;
        if (old != bi || recount) {
            const int r = rhccq_key_r(c), gg = rhccq_key_g(c), b = rhccq_key_b(c);
            if (!recount && old != (int)NOLABEL) {
                atomicAdd(&C.sums[3 * old], -r); atomicAdd(&C.sums[3 * old + 1], -gg); atomicAdd(&C.sums[3 * old + 2], -b);
                atomicAdd(&C.cnt[old], -1);
            }
            atomicAdd(&C.sums[3 * bi], r); atomicAdd(&C.sums[3 * bi + 1], gg); atomicAdd(&C.sums[3 * bi + 2], b);
            atomicAdd(&C.cnt[bi], 1);
        }
    };

    g.sync();
    for (int it = 0; ; ++it) {
        // ---- E step: lab_cur <- nearest centre of every point (first minimum of the float64 scores)
        //
        // Pruning: a point whose distance to the centre of its previous cluster is safely less than half the
        // distance from that centre to the nearest other centre keeps its cluster by the triangle inequality.
        // In float32: u = |x - c_a|^2 < 0.24975 m - 0.5 with m the squared distance from c_a to the nearest other
        // centre; the float32 errors of u and m are below 0.15 and 0.3, so the true u is below 0.24975 m - 0.27 and
        // every other centre is farther by m (1 - 2 sqrt(0.24975)) >= 5e-4 m >= 1e-3 in the squared distance (m > 2
        // for the test to pass at all) — far above every margin of the levels below, so none of them is consulted.
        const double margin = cen_exact ? (RHCCQ_KM_FORCED ? 1.0e300 : (e_hi > e_lo ? 1.0e-9 : 0.0)) : margin_full;
        const bool prune = C.hb != nullptr && prune_valid && !RHCCQ_KM_FORCED;
        bool e_old_ready = false;                                   // E_old holds the float64 centres of this E step
        changed = 0;
        if (tid == 0) { flag[0] = 0; flag[1] = 0; }                 // open decisions; length of the worklist
        g.sync();
        if (prune) {
            const float* thr = reinterpret_cast<const float*>(C.hb);
            for (int j = tid; j - RHCCQ_LANE < n; j += gsz) {          // warp-uniform trip count
                bool push = false;
                if (j < n) {
                    const uint32_t c = x[j];
                    const int a = (int)lab_prev[j];
                    const float d0 = fmaf(0.5f, C.cf[4 * a], __fsub_rn((float)rhccq_key_r(c), meanf[0]));
                    const float d1 = fmaf(0.5f, C.cf[4 * a + 1], __fsub_rn((float)rhccq_key_g(c), meanf[1]));
                    const float d2 = fmaf(0.5f, C.cf[4 * a + 2], __fsub_rn((float)rhccq_key_b(c), meanf[2]));
                    push = !(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)) < thr[a]);
                    if (!push) lab_cur[j] = (idx_t)a;
                }
                const unsigned m = rhccq_ballot(push);                 // one counter bump per warp
                int base = 0;
                if (RHCCQ_LANE == 0 && m) base = atomicAdd(&flag[1], __popc(m));
                base = rhccq_shfl(base, 0);
                if (push) wl[base + __popc(m & rhccq_lanemask_lt())] = (idx_t)j;
            }
        }
        g.sync();
        // without pruning every position is evaluated in place and the open ones are collected at the head of wl
        const int n_work = prune ? flag[1] : n;
        // few points: one per thread (short dependent chains, every warp busy); many: RHCCQ_EB consecutive entries per
        // thread (every centre is loaded once per RHCCQ_EB points)
        if (n_work <= 2 * gsz) {
            for (int i = tid; i < n_work; i += gsz) {
                const int j = prune ? (int)wl[i] : i;
                const uint32_t cb[1] = {x[j]};
                int bb[1];
                bool und[1];
                rhccq_nearest_centers_f32<1>(cb, meanf, C.cf, k, bb, und);
                if (und[0] || RHCCQ_KM_FORCED) {                        // decided below
                    const int slot = atomicAdd(&flag[0], 1);
                    wl[prune ? i : slot] = (idx_t)(j | Cfg::FLAG);
                } else commit(j, cb[0], bb[0]);
            }
        } else {
            for (int i0 = RHCCQ_EB * tid; i0 < n_work; i0 += RHCCQ_EB * gsz) {
                uint32_t cb[RHCCQ_EB];
                int bb[RHCCQ_EB], jj[RHCCQ_EB];
                bool und[RHCCQ_EB];
#pragma unroll
                for (int u = 0; u < RHCCQ_EB; ++u) {
                    const int i = i0 + u;
                    const int is = i < n_work ? i : i0;                 // slots past the end redo an entry and drop it
                    jj[u] = prune ? (int)wl[is] : is;
                    cb[u] = x[jj[u]];
                }
                rhccq_nearest_centers_f32<RHCCQ_EB>(cb, meanf, C.cf, k, bb, und);
#pragma unroll
                for (int u = 0; u < RHCCQ_EB; ++u) {
                    const int i = i0 + u;
                    if (i >= n_work) continue;
                    if (und[u] || RHCCQ_KM_FORCED) {
                        const int slot = atomicAdd(&flag[0], 1);
                        wl[prune ? i : slot] = (idx_t)(jj[u] | Cfg::FLAG);
                    } else commit(jj[u], cb[u], bb[u]);
                }
            }
        }
        g.sync();
        const int n_und1 = flag[0];
        const int n_scan = prune ? n_work : n_und1;                 // entries of wl that may carry Cfg::FLAG
        if (tid == 0) { RHCCQ_COUNT(8, 1); RHCCQ_COUNT(14, n_und1); }
        if (n_und1 > 0) {
            // second level: float64 scores on the integer-derived centres
            g.sync();                                               // every thread has read the count
            if (tid == 0) flag[0] = 0;
            g.sync();
#pragma unroll 1
            for (int base = g.sub() * RHCCQ_WARP_SIZE; base < n_scan; base += g.nsub() * RHCCQ_WARP_SIZE) {
                // the lanes look at 32 entries of the worklist; every open one is then decided by the whole warp,
                // lanes over the centres
                const int il = base + RHCCQ_LANE;
                unsigned open = rhccq_ballot(il < n_scan && (wl[il] & Cfg::FLAG) != 0);
#pragma unroll 1
                while (open) {
                    const int i = base + __ffs((int)open) - 1;
                    open &= open - 1u;
                    const int j = (int)(wl[i] & (idx_t)~Cfg::FLAG);
                    const rhccq_sk_pt p = rhccq_sk_centred(x[j], mean);
                    double best = 1.0e300, sec = 1.0e300;
                    int bi = 0x7fffffff;
#pragma unroll 1
                    for (int q = RHCCQ_LANE; q < k; q += RHCCQ_WARP_SIZE) {
                        const double d = rhccq_sk_score(p, cen + 3 * q, C.csn[q], false);
                        if (d < best) { sec = best; best = d; bi = q; }
                        else if (d < sec) sec = d;
                    }
#pragma unroll
                    for (int o = RHCCQ_WARP_SIZE >> 1; o > 0; o >>= 1) {
                        const double ob = rhccq_shfl_xor(best, o), os = rhccq_shfl_xor(sec, o);
                        const int oi = rhccq_shfl_xor(bi, o);
                        const double hi = ob > best ? ob : best, lo2 = os < sec ? os : sec;
                        sec = hi < lo2 ? hi : lo2;                      // runner-up of the union
                        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    __syncwarp();
                    if (RHCCQ_LANE == 0) {
                        if (sec < __dadd_rn(best, margin)) atomicAdd(&flag[0], 1);
                        else { commit(j, x[j], bi); wl[i] = (idx_t)j; }
                    }
                    __syncwarp();
                }
            }
            g.sync();
            if (flag[0] > 0) {
                // third level: the float64 centres of scikit-learn's M step, one thread
                if (tid == 0) { RHCCQ_COUNT(15, 1); }
                if (tid == 0) rhccq_sk_e_old<Cfg>(x, lab_prev, n, k, mean, cen, C.csn, cen_exact, E_old);
                e_old_ready = true;
                g.sync();
#pragma unroll 1
                for (int i = tid; i < n_scan; i += gsz) {
                    if (!(wl[i] & Cfg::FLAG)) continue;
                    const int j = (int)(wl[i] & (idx_t)~Cfg::FLAG);
                    commit(j, x[j], rhccq_nearest_exact(x[j], mean, E_old, k, j, n, e_lo, e_hi));
                }
            }
        }
        if (final_pass) { idx_t* t = lab_prev; lab_prev = lab_cur; lab_cur = t; break; }
        recount = false;
        changed = g.any(changed);
        int empty = 0;
        for (int q = tid; q < k; q += gsz) if (C.cnt[q] == 0) ++empty;
        const int n_empty = k <= gsz ? g.count(empty) : g.sum_i(empty);    // k <= gsz: at most one cluster per thread
        int decided = 0;
        bool next_exact = false;
        if (n_empty > 0) {
            // everything in float64, one thread; the new centres are the float64 ones themselves
            if (tid == 0) {
                if (!e_old_ready) rhccq_sk_e_old<Cfg>(x, lab_prev, n, k, mean, cen, C.csn, cen_exact, E_old);
                rhccq_sk_relocate<Cfg>(x, lab_cur, n, k, n_empty, mean, E_old, E_new, s2);
                flag[1] = rhccq_sk_converged(x, n, k, mean, E_old, E_new, s2);
#pragma unroll 1
                for (int q = 0; q < k; ++q) {
                    cen_new[3 * q] = E_new[4 * q]; cen_new[3 * q + 1] = E_new[4 * q + 1]; cen_new[3 * q + 2] = E_new[4 * q + 2];
                    C.csn[q] = E_new[4 * q + 3];
                }
            }
            g.sync();
            decided = flag[1];
            next_exact = true;
            prune_valid = false;                                    // C.hb was not recomputed for the relocated centres
            recount = true;                                         // the integer sums no longer follow the labels
            for (int q = tid; q < k; q += gsz) {
                C.cnt[q] = 0; C.sums[3 * q] = 0; C.sums[3 * q + 1] = 0; C.sums[3 * q + 2] = 0;
                C.cf[4 * q] = (float)__dmul_rn(-2.0, cen_new[3 * q]); C.cf[4 * q + 1] = (float)__dmul_rn(-2.0, cen_new[3 * q + 1]);
                C.cf[4 * q + 2] = (float)__dmul_rn(-2.0, cen_new[3 * q + 2]); C.cf[4 * q + 3] = (float)C.csn[q];
            }
        } else {
            for (int q = tid; q < k; q += gsz) {
                const double cn = (double)C.cnt[q];
                const double c0 = __dsub_rn(__ddiv_rn((double)C.sums[3 * q], cn), mean[0]);
                const double c1 = __dsub_rn(__ddiv_rn((double)C.sums[3 * q + 1], cn), mean[1]);
                const double c2 = __dsub_rn(__ddiv_rn((double)C.sums[3 * q + 2], cn), mean[2]);
                cen_new[3 * q] = c0; cen_new[3 * q + 1] = c1; cen_new[3 * q + 2] = c2;
                const double a0 = __dsub_rn(c0, cen[3 * q]), a1 = __dsub_rn(c1, cen[3 * q + 1]), a2 = __dsub_rn(c2, cen[3 * q + 2]);
                C.term[q] = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
            }
            g.sync();
            double shift = 0.0;
            if (k <= 64) {
                for (int q = 0; q < k; ++q) shift = __dadd_rn(shift, C.term[q]);         // every thread, same order
            } else {
                if (tid == 0) for (int q = 0; q < k; ++q) shift = __dadd_rn(shift, C.term[q]);
                shift = g.bcast_d(shift);
            }
            if (changed) {
                // center_shift_tot <= tol (:724-733).  Each centre coordinate is within eps_c of its float64 value,
                // so the shift is within 4 sqrt(3) eps_c sqrt(k shift) of the float64 one, the tolerance within
                // n 2^-52 relative.  Inside that band: float64 centres, shifts, numpy's pairwise sum, np.var.
                const double big = shift > tol ? shift : tol;
                const double band = __dadd_rn(__dmul_rn(__dmul_rn(16.0, eps_c), sqrt(__dmul_rn((double)k, big))), __dmul_rn(1.0e-9, tol));
                const double diff = shift > tol ? __dsub_rn(shift, tol) : __dsub_rn(tol, shift);
                if (diff <= band || RHCCQ_KM_FORCED) {
                    g.sync();
                    if (tid == 0) {
                        RHCCQ_COUNT(11, 1);
                        if (!e_old_ready) rhccq_sk_e_old<Cfg>(x, lab_prev, n, k, mean, cen, C.csn, cen_exact, E_old);
                        rhccq_sk_all_centers<Cfg>(x, lab_cur, n, k, mean, E_new);
                        rhccq_sk_average(E_new, k);
                        flag[1] = rhccq_sk_converged(x, n, k, mean, E_old, E_new, s2);
                    }
                    g.sync();
                    decided = flag[1];
                } else {
                    decided = shift <= tol ? 1 : 0;
                }
            }
            g.sync();                                               // term / csn are rewritten next
            for (int q = tid; q < k; q += gsz) {
                C.csn[q] = rhccq_sk_norm3(cen_new[3 * q], cen_new[3 * q + 1], cen_new[3 * q + 2]);
                C.cf[4 * q] = (float)__dmul_rn(-2.0, cen_new[3 * q]); C.cf[4 * q + 1] = (float)__dmul_rn(-2.0, cen_new[3 * q + 1]);
                C.cf[4 * q + 2] = (float)__dmul_rn(-2.0, cen_new[3 * q + 2]); C.cf[4 * q + 3] = (float)C.csn[q];
            }
            if (C.hb != nullptr) {
                // pruning thresholds of the new centres: 0.24975 m - 0.5, m = squared distance to the nearest other centre
                g.sync();
                float* thr = reinterpret_cast<float*>(C.hb);
                for (int q = g.sub(); q < k; q += g.nsub()) {           // one warp per centre, lanes over the others
                    float m = 3.0e38f;
                    const float q0 = C.cf[4 * q], q1 = C.cf[4 * q + 1], q2 = C.cf[4 * q + 2];
                    for (int r = RHCCQ_LANE; r < k; r += RHCCQ_WARP_SIZE) {
                        const float d0 = q0 - C.cf[4 * r], d1 = q1 - C.cf[4 * r + 1], d2 = q2 - C.cf[4 * r + 2];
                        const float d = 0.25f * fmaf(d0, d0, fmaf(d1, d1, d2 * d2));      // cf holds -2 c
                        m = (r != q && d < m) ? d : m;
                    }
#pragma unroll
                    for (int o = RHCCQ_WARP_SIZE >> 1; o > 0; o >>= 1) m = fminf(m, rhccq_shfl_xor(m, o));
                    if (RHCCQ_LANE == 0) thr[q] = fmaf(0.24975f, m, -0.5f);
                }
                prune_valid = true;
            }
        }
        { double* t = cen; cen = cen_new; cen_new = t; }
        { idx_t* t = lab_prev; lab_prev = lab_cur; lab_cur = t; }
        cen_exact = next_exact;
        if (!changed) break;                                        // strict convergence (:717-722): the labels stand
        if (decided || it == 299) final_pass = true;                // :724-733 / max_iter: one E step with the final centres
        g.sync();                                                   // the tables above are read by the next E step
    }
    // the result belongs in A.label (the partition that follows overwrites `cum`)
    g.sync();
    if (lab_prev != A.label + lo) {
        idx_t* dst = A.label + lo;
        for (int j = tid; j < n; j += gsz) dst[j] = lab_prev[j];
    }
    for (int q = tid; q < k; q += gsz) C.cnt[q] = 0;
    g.sync();
    for (int j = tid; j < n; j += gsz) atomicAdd(&C.cnt[(int)A.label[lo + j]], 1);
    g.sync();
    if (g.size() > RHCCQ_WARP_SIZE) RHCCQ_PROF(2);
}

// Ascending sort of a[0..n) for any n: the bitonic network in its "flip" form, in which every
// compare-exchange orders (lower index, higher index) ascending; padding the array to a power of two
// with +infinity would leave the padding in place, so pairs that reach beyond n are skipped.
template <class G, class T>
__device__ __forceinline__ void rhccq_group_sort(const G& g, T* a, int n) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int t = g.tid(); t < (np2 >> 1); t += g.size()) {      // flip step: i ^ (k - 1)
            const int blk = t / (k >> 1), off = t % (k >> 1);
            const int i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < n) { const T xa = a[i], xb = a[l]; if (xb < xa) { a[i] = xb; a[l] = xa; } }
        }
        g.sync();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = g.tid(); t < (np2 >> 1); t += g.size()) {
                const int i = 2 * t - (t & (j - 1)), l = i + j;
                if (l < n) { const T xa = a[i], xb = a[l]; if (xb < xa) { a[i] = xb; a[l] = xa; } }
            }
            g.sync();
        }
    }
}

// ---------------------------------------------------------------- split driver
template <class Cfg> struct rhccq_split_ws {
    rhccq_km_arrays<Cfg> A;
    typename Cfg::idx_t* perm;          // row of every position; Cfg::FLAG marks the first position of a leaf
    typename Cfg::q_t* queue;           // ranges still to split, level by level
};

__device__ __forceinline__ void rhccq_carve_centers(rhccq_km_centers& C, unsigned char* base, size_t kc) {
    rhccq_carver cv(base);
    C.center = cv.take<double>(3 * kc);
    C.center_new = cv.take<double>(3 * kc);
    C.term = cv.take<double>(kc);
    C.csn = cv.take<double>(kc);
    C.cf = cv.take<float>(4 * kc);
    C.hb = cv.take<int>(kc);
    C.sums = cv.take<int>(3 * kc);
    C.cnt = cv.take<int>(kc);
}

template <class Cfg>
__host__ __device__ static inline size_t rhccq_split_row_bytes(size_t rows) {
    typedef typename Cfg::idx_t idx_t;
    return rhccq_carve_bytes(rows, 4) * 2 + rhccq_carve_bytes(rows, sizeof(typename Cfg::cum_t))
           + rhccq_carve_bytes(rows, sizeof(idx_t)) * 2 + rhccq_carve_bytes(rows, sizeof(typename Cfg::q_t));
}
__host__ __device__ static inline size_t rhccq_split_center_bytes(size_t kc) {
    return rhccq_carve_bytes(3 * kc, 8) * 2 + rhccq_carve_bytes(kc, 8) * 2 + rhccq_carve_bytes(4 * kc, 4)
           + rhccq_carve_bytes(kc, 4) + rhccq_carve_bytes(3 * kc, 4) + rhccq_carve_bytes(kc, 4);
}

size_t rhccq_palette_split_ws_bytes(int max_rows) {
    // one slice of the global workspace: the per-row arrays and centre tables for k up to max_rows
    const size_t r = (size_t)(max_rows > 1 ? max_rows : 1);
    const size_t rows = max_rows <= rhccq_cfg_small::MAX_ROWS ? rhccq_split_row_bytes<rhccq_cfg_small>(r)
                                                              : rhccq_split_row_bytes<rhccq_cfg_large>(r);
    // never 0 and never within the shared-memory budget: the caller must always pass a workspace (the
    // centre tables of a K-Means with more than RHCCQ_KC centres live there)
    // ... and the float64 scratch of the decisions that are re-evaluated in scikit-learn's own arithmetic
    const size_t need = rows + rhccq_split_center_bytes(r) + rhccq_carve_bytes((size_t)RHCCQ_EX_PER_ROW * r, 8);
    return need > RHCCQ_SMEM_BUDGET ? need : (size_t)RHCCQ_SMEM_BUDGET + 16;
}

// Split the range [lo, hi) by K-Means and partition it stably by label; children that are still too
// large are queued, the others are flagged as leaves.  Every thread of the group must call.
template <class G, class Cfg>
__device__ __forceinline__ void rhccq_split_range(const G& g, const rhccq_split_ws<Cfg>& W, const rhccq_km_centers& C, int lo, int hi,
                                  int k, int mcpc, const double* __restrict__ rng, int* q_tail, int q_cap, int* err) {
    typedef typename Cfg::idx_t idx_t;
    typedef typename Cfg::key_t key_t;
    const int cnt = hi - lo;
    rhccq_kmeans<G, Cfg>(g, W.A, C, lo, cnt, k, rng);
    uint32_t* tx = W.A.closest + lo;                                 // dead after the seeding
    if (C.poff != nullptr) {
        // stable partition by label, counting form: every warp of the group owns a contiguous segment of the
        // range; offsets per (label, warp) in label-major order; inside a warp the elements of a tile of 32
        // consecutive positions are ranked among equal labels with a match, tiles in order
        const int nsub = g.nsub(), sub = g.sub();
        const int seg = (cnt + nsub - 1) / nsub;
        const int s_lo = sub * seg < cnt ? sub * seg : cnt;
        const int s_hi = s_lo + seg < cnt ? s_lo + seg : cnt;
        idx_t* tp = reinterpret_cast<idx_t*>(W.A.cum + lo);          // dead as well; labels stay readable
        for (int q = g.tid(); q < k * nsub; q += g.size()) C.poff[q] = 0;
        g.sync();
        for (int j = s_lo + RHCCQ_LANE; j < s_hi; j += RHCCQ_WARP_SIZE)
            atomicAdd(&C.poff[(int)W.A.label[lo + j] * nsub + sub], 1);
        g.sync();
        {   // exclusive scan of poff[0 .. k * nsub)
            const int N = k * nsub, gsz = g.size(), tid = g.tid();
            const int per = (N + gsz - 1) / gsz;
            const int a = tid * per < N ? tid * per : N, b = a + per < N ? a + per : N;
            int sum = 0;
            for (int i = a; i < b; ++i) sum += C.poff[i];
            int total;
            int base = g.template excl_scan<int>(sum, &total);
            for (int i = a; i < b; ++i) { const int v = C.poff[i]; C.poff[i] = base; base += v; }
        }
        g.sync();
        for (int q = g.tid(); q < k; q += g.size()) C.sums[q] = C.poff[q * nsub];      // first position of child q
        g.sync();
        for (int t0 = s_lo; t0 < s_hi; t0 += RHCCQ_WARP_SIZE) {
            const int j = t0 + RHCCQ_LANE;
            const bool valid = j < s_hi;
            const int L = valid ? (int)W.A.label[lo + j] : -1;
#ifdef RHCCQ_HOST_EMU
            const int rank = 0, same = 1;
            const bool leader = true;
#else
            const unsigned m = __match_any_sync(0xffffffffu, L);
            const int rank = __popc(m & ((1u << (threadIdx.x & 31)) - 1u)), same = __popc(m);
            const bool leader = rank == same - 1;
#endif
            int base = 0;
            if (valid) base = C.poff[L * nsub + sub];
            __syncwarp();
            if (valid) {
                tx[base + rank] = W.A.x[lo + j];
                tp[base + rank] = W.perm[lo + j];
                if (leader) C.poff[L * nsub + sub] = base + same;
            }
            __syncwarp();
        }
        g.sync();
        for (int j = g.tid(); j < cnt; j += g.size()) { W.A.x[lo + j] = tx[j]; W.perm[lo + j] = tp[j]; }
    } else {
        // many centres: sort (label, position) keys, then move colours and rows
        key_t* keys = reinterpret_cast<key_t*>(W.A.cum + lo);
        for (int j = g.tid(); j < cnt; j += g.size()) keys[j] = Cfg::key((int)W.A.label[lo + j], j);
        g.sync();
        rhccq_group_sort<G, key_t>(g, keys, cnt);
        idx_t* tp = W.A.label + lo;                                  // labels are in the keys now
        for (int j = g.tid(); j < cnt; j += g.size()) {
            const int src = lo + Cfg::key_j(keys[j]);
            tx[j] = W.A.x[src];
            tp[j] = W.perm[src];
        }
        g.sync();
        for (int j = g.tid(); j < cnt; j += g.size()) { W.A.x[lo + j] = tx[j]; W.perm[lo + j] = tp[j]; }
        // first position of every child
        for (int q = g.tid(); q < k; q += g.size()) C.sums[q] = C.cnt[q];
        g.sync();
        if (g.tid() == 0) {
            int run = 0;
            for (int q = 0; q < k; ++q) { const int c = C.sums[q]; C.sums[q] = run; run += c; }
        }
    }
    g.sync();
    // children in label order (clustering.py:755-767): one decision per child
    for (int q = g.tid(); q < k; q += g.size()) {
        const int c = C.cnt[q];
        if (c == 0) continue;                                        // clustering.py:757
        const int s = lo + C.sums[q];
        if (c > mcpc && c < cnt && c > 2) {                          // :763-767, :745
            const int slot = atomicAdd(q_tail, 1);
            if (slot < q_cap) W.queue[slot] = Cfg::q_pack(s, s + c); else *err = 1;
        } else {
            W.perm[s] = (idx_t)(W.perm[s] | Cfg::FLAG);
        }
    }
    g.sync();
}

template <class Cfg>
__device__ __forceinline__ void rhccq_palette_split_problem(const rhccq_palette_batch& B, int p, const int* __restrict__ labels,
                                            const int* __restrict__ status_in, const int* __restrict__ max_cpc,
                                            const double* __restrict__ rng, int rng_len, int* __restrict__ leaf,
                                            int* __restrict__ n_leaves, int max_rows, unsigned char* row_base,
                                            unsigned char* small_base, unsigned char* cent_s, int kc_s,
                                            unsigned char* cent_g, int kc_g, double* ex) {
    typedef typename Cfg::idx_t idx_t;
    typedef typename Cfg::q_t q_t;
    __shared__ long long s_ll[RHCCQ_MAX_WARPS * RHCCQ_KM_MAXT + 2];
    __shared__ long long s_sv[2 * RHCCQ_SPLIT_MAX_WARPS * RHCCQ_KM_MAXT];
    __shared__ int s_scan[RHCCQ_MAX_WARPS + 2];
    __shared__ int s_tail, s_err, s_base, s_claim, s_wl;
    __shared__ int wlc[RHCCQ_MAX_WARPS];
    RHCCQ_PROF_T0();
    const int n = B.pal_cnt[p];
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    const int* lab = labels + B.pal_off[p];
    int* lf = leaf + B.pal_off[p];
    // max_cpc[p] = -k: one KMeans(k) over the whole palette and nothing else (the third-party operator of
    // clustering.py:751-752 by itself): every row is a member of one root, the k children are the leaves
    const int kforce = max_cpc[p] < 0 ? -max_cpc[p] : 0;
    const int mcpc = kforce ? 0 : max_cpc[p];
    if (n < 0 || (status_in != nullptr && status_in[p] < 0)) {     // upstream error: pass it on
        if (threadIdx.x == 0) n_leaves[p] = n < 0 ? -2 : status_in[p];
        return;
    }
    if (n > max_rows || n > Cfg::MAX_ROWS) {
        if (threadIdx.x == 0) n_leaves[p] = -1;
        return;
    }
    rhccq_split_ws<Cfg> W;
    {
        rhccq_carver cv(row_base);
        W.A.x = cv.take<uint32_t>(max_rows);
        W.A.closest = cv.take<uint32_t>(max_rows);
        W.A.cum = cv.take<typename Cfg::cum_t>(max_rows);
        W.A.label = cv.take<idx_t>(max_rows);
        W.perm = cv.take<idx_t>(max_rows);
        W.queue = cv.take<q_t>(max_rows);
        W.A.ex = ex;
    }
    // CTA-level centre tables: in shared memory for k <= kc_s, else in the global workspace (k <= kc_g)
    rhccq_km_centers CS, CG;
    rhccq_carve_centers(CS, cent_s, (size_t)kc_s);
    rhccq_carve_centers(CG, cent_g, (size_t)kc_g);
    // small shared tables: per-warp centre sets, per-warp M-step histograms, candidate slots
    rhccq_carver sv(small_base);
    const size_t nw = (size_t)RHCCQ_NWARPS;                        // the tables are sized for the launch's warps
    int* cand = sv.take<int>((nw + 1) * RHCCQ_KM_CAND);
    double* wcent = sv.take<double>(nw * 8 * RHCCQ_KW);
    int* wint = sv.take<int>(nw * 5 * RHCCQ_KW);
    float* wcf = sv.take<float>(nw * 4 * RHCCQ_KW);
    int* poff = sv.take<int>(nw * RHCCQ_KC);
    CS.cand = CG.cand = cand + nw * RHCCQ_KM_CAND;

    // ---- entries that need no K-Means.  csize / crank live in the (still unused) seeding arrays.
    int* csize = reinterpret_cast<int*>(W.A.x);
    int* crank = reinterpret_cast<int*>(W.A.closest);
    // One look at the rows: is there a black row, a noise row, and how far do the cluster labels reach.  The
    // usual palette has neither black nor noise and a single cluster (the eps radii of the reference chain
    // whole palettes), so the scans below shrink to nothing.
    int f_black = 0, f_noise = 0, l_max = -1;
    RHCCQ_PAR_FOR(i, n) {
        const int l = lab[i];
        if (keys[i] == 0u) f_black = 1; else if (l == -1) f_noise = 1;
        l_max = l > l_max ? l : l_max;
    }
    f_black = rhccq_block_or(f_black, s_scan);
    f_noise = rhccq_block_or(f_noise, s_scan);
    const int n_lab = rhccq_block_max<int>(l_max, s_scan) + 1;      // labels are 0 .. n_lab - 1 (n_lab <= n)
    RHCCQ_PAR_FOR(l, n_lab) csize[l] = 0;
    __syncthreads();
    // black rows first, one entry each, in row order (clustering.py:253-255)
    int n_black = 0;
    if (f_black) {
        RHCCQ_PAR_FOR(i, n) crank[i] = (keys[i] == 0u) ? 1 : 0;
        __syncthreads();
        n_black = rhccq_block_excl_scan_array<int>(crank, n, s_scan);
        RHCCQ_PAR_FOR(i, n) if (keys[i] == 0u) lf[i] = crank[i];
        __syncthreads();
    }
    // cluster sizes; labels are dense non-negative (a noise row would carry -1: one entry each, :258-264)
    RHCCQ_PAR_FOR(i, n) if (lab[i] >= 0) atomicAdd(&csize[lab[i]], 1);
    __syncthreads();
    int n_noise = 0;
    if (f_noise) {
        RHCCQ_PAR_FOR(i, n) crank[i] = (keys[i] != 0u && lab[i] == -1) ? 1 : 0;
        __syncthreads();
        n_noise = rhccq_block_excl_scan_array<int>(crank, n, s_scan);
        RHCCQ_PAR_FOR(i, n) if (keys[i] != 0u && lab[i] == -1) lf[i] = n_black + crank[i];
        __syncthreads();
    }
    // small clusters in ascending label order (:273-310)
    RHCCQ_PAR_FOR(l, n_lab) crank[l] = (csize[l] > 0 && csize[l] <= mcpc) ? 1 : 0;
    __syncthreads();
    const int n_small = rhccq_block_excl_scan_array<int>(crank, n_lab, s_scan);
    RHCCQ_PAR_FOR(i, n) {
        const int l = lab[i];
        if (l >= 0 && csize[l] <= mcpc) lf[i] = n_black + n_noise + crank[l];
    }
    __syncthreads();
    // large clusters in ascending label order (:315-355): their members, in ascending row order, fill
    // consecutive ranges of the permutation; crank[l] <- first position of cluster l
    RHCCQ_PAR_FOR(l, n_lab) crank[l] = csize[l] > mcpc ? csize[l] : 0;
    __syncthreads();
    const int n_members = rhccq_block_excl_scan_array<int>(crank, n_lab, s_scan);
    if (threadIdx.x == 0) { s_tail = 0; s_err = 0; s_base = n_black + n_noise + n_small; }
    __syncthreads();
    // roots: a cluster of more than two colours is split (:745), a larger-than-allowed pair stays one entry
    RHCCQ_PAR_FOR(l, n_lab) {
        const int c = csize[l];
        if (c > mcpc) {
            const int slot = atomicAdd(&s_tail, 1);
            W.queue[slot] = Cfg::q_pack(crank[l], crank[l] + c);     // at most n / 2 roots: fits
        }
    }
    __syncthreads();
    const int n_roots = s_tail;
    if (n_roots > 0) {
        // rows of the large clusters in (cluster, row) order: rank of a row inside its cluster by a scan per
        // root (there is rarely more than one root: the eps radii of the reference chain whole palettes)
        for (int ri = 0; ri < n_roots; ++ri) {
            const int r_lo = Cfg::q_lo(W.queue[ri]);
            int L = -1;
            // the cluster whose range starts at r_lo
            __shared__ int s_L;
            RHCCQ_PAR_FOR(l, n_lab) if (csize[l] > mcpc && crank[l] == r_lo) s_L = l;
            __syncthreads();
            L = s_L;
            int* flag = reinterpret_cast<int*>(W.A.cum);             // n ints fit: cum_t is at least 4 bytes
            RHCCQ_PAR_FOR(i, n) flag[i] = (lab[i] == L) ? 1 : 0;
            __syncthreads();
            rhccq_block_excl_scan_array<int>(flag, n, s_scan);
            RHCCQ_PAR_FOR(i, n) if (lab[i] == L) W.perm[r_lo + flag[i]] = (idx_t)i;
            __syncthreads();
        }
    }
    // csize / crank are dead from here on: x takes the colours of the permuted rows
    __syncthreads();
    {
        // a root of one or two colours cannot be split (:745): it is a leaf; mark and drop it from the queue
        __shared__ int s_keep;
        if (threadIdx.x == 0) {
            int keep = 0;
            for (int ri = 0; ri < n_roots; ++ri) {
                const q_t e = W.queue[ri];
                if (Cfg::q_hi(e) - Cfg::q_lo(e) <= 2) W.perm[Cfg::q_lo(e)] = (idx_t)(W.perm[Cfg::q_lo(e)] | Cfg::FLAG);
                else W.queue[keep++] = e;
            }
            s_keep = keep;
            s_tail = keep;
        }
        __syncthreads();
        (void)s_keep;
    }
    RHCCQ_PAR_FOR(j, n_members) W.A.x[j] = keys[(int)(W.perm[j] & (idx_t)~Cfg::FLAG)];
    __syncthreads();

    RHCCQ_PROF(0);                                                 // prologue
    // ---- level-synchronous splitting
    rhccq_grp_cta gc; gc.sll = s_ll; gc.svb = s_sv; gc.svp = 0;
    gc.csum = reinterpret_cast<unsigned long long*>(poff);         // the partition table is idle during a seeding (8 B x threads fit)
    rhccq_grp_warp gw; gw.csum = nullptr;
    int head = 0;
    while (true) {
        __syncthreads();
        const int tail = s_tail;
        if (head >= tail || s_err) break;
        // ranges for the whole CTA
        for (int e = head; e < tail; ++e) {
            const q_t qe = W.queue[e];
            const int lo = Cfg::q_lo(qe), hi = Cfg::q_hi(qe), cnt = hi - lo;
            int k = kforce ? kforce : (cnt + mcpc - 1) / mcpc;      // clustering.py:739-742
            if (k < 2) k = 2;
            if (k > cnt) k = cnt;
            if (cnt <= RHCCQ_WARP_RANGE && k <= RHCCQ_KW) continue;  // a warp's job
            const bool in_smem = k <= kc_s;
            if (1 + (k - 1) * rhccq_kmeans_local_trials(k) > rng_len || (!in_smem && (cent_g == nullptr || k > kc_g))) {
                // random table too short, or more centres than the caller's workspace holds: report, do not guess
                if (threadIdx.x == 0) s_err = (1 + (k - 1) * rhccq_kmeans_local_trials(k) > rng_len) ? 2 : 3;
                break;
            }
            if (in_smem) {
                rhccq_km_centers C = CS;
                if (k < RHCCQ_PRUNE_MIN_K) C.hb = nullptr;
                C.hist = nullptr;
                C.poff = poff;
                C.wl = &s_wl;
                rhccq_split_range<rhccq_grp_cta, Cfg>(gc, W, C, lo, hi, k, kforce ? 0x7fffffff : mcpc, rng, &s_tail, max_rows, &s_err);
            } else {
                rhccq_km_centers C = CG;
                C.hist = nullptr;
                C.poff = nullptr;
                C.wl = nullptr;
                C.hb = nullptr;                                      // k * k centre pairs per iteration would not pay
                rhccq_split_range<rhccq_grp_cta, Cfg>(gc, W, C, lo, hi, k, kforce ? 0x7fffffff : mcpc, rng, &s_tail, max_rows, &s_err);
            }
        }
        __syncthreads();
        RHCCQ_PROF(3);                                             // CTA-level splits (K-Means + partition)
        if (s_err) break;
        // ranges for single warps, concurrently; warps claim the next range when they are done with one
        if (threadIdx.x == 0) s_claim = head;
        __syncthreads();
        while (true) {
            int e = 0;
            if (RHCCQ_LANE == 0) e = atomicAdd(&s_claim, 1);
            e = rhccq_shfl(e, 0);
            if (e >= tail) break;
            const q_t qe = W.queue[e];
            const int lo = Cfg::q_lo(qe), hi = Cfg::q_hi(qe), cnt = hi - lo;
            int k = kforce ? kforce : (cnt + mcpc - 1) / mcpc;
            if (k < 2) k = 2;
            if (k > cnt) k = cnt;
            if (!(cnt <= RHCCQ_WARP_RANGE && k <= RHCCQ_KW)) continue;
            if (1 + (k - 1) * rhccq_kmeans_local_trials(k) > rng_len) { s_err = 2; continue; }
            rhccq_km_centers C;
            double* wc = wcent + (size_t)RHCCQ_WARP * 8 * RHCCQ_KW;
            int* wi = wint + (size_t)RHCCQ_WARP * 5 * RHCCQ_KW;
            C.center = wc; C.center_new = wc + 3 * RHCCQ_KW; C.term = wc + 6 * RHCCQ_KW; C.csn = wc + 7 * RHCCQ_KW;
            C.sums = wi; C.cnt = wi + 3 * RHCCQ_KW; C.poff = wi + 4 * RHCCQ_KW;
            C.cf = wcf + (size_t)RHCCQ_WARP * 4 * RHCCQ_KW;
            C.hb = nullptr;                                          // k <= RHCCQ_KW centres: evaluating them all is cheaper than pruning
            C.hist = nullptr;
            C.cand = cand + RHCCQ_WARP * RHCCQ_KM_CAND;
            C.wl = wlc + RHCCQ_WARP;
            rhccq_split_range<rhccq_grp_warp, Cfg>(gw, W, C, lo, hi, k, kforce ? 0x7fffffff : mcpc, rng, &s_tail, max_rows, &s_err);
        }
        head = tail;
        __syncthreads();
        RHCCQ_PROF(4);                                             // warp-level splits of this level
    }
    __syncthreads();
    if (s_err) {
        if (threadIdx.x == 0) n_leaves[p] = s_err == 2 ? -2 : -1;
        return;
    }
    // ---- leaves in range order == depth-first K-Means label order
    int* lrank = reinterpret_cast<int*>(W.A.closest);
    RHCCQ_PAR_FOR(j, n_members) lrank[j] = (W.perm[j] & Cfg::FLAG) ? 1 : 0;
    __syncthreads();
    const int n_split_leaves = rhccq_block_excl_scan_array<int>(lrank, n_members, s_scan);
    {
        // inclusive rank - 1 == leaf number of the position; every thread walks a contiguous chunk
        const int nt = (int)blockDim.x, t = (int)threadIdx.x;
        const int per = (n_members + nt - 1) / nt;
        const int lo = t * per < n_members ? t * per : n_members;
        const int hi = lo + per < n_members ? lo + per : n_members;
        const int base = s_base;
        for (int j = lo; j < hi; ++j) {
            const uint32_t pe = W.perm[j];
            const int is_start = (pe & Cfg::FLAG) ? 1 : 0;
            lf[(int)(pe & ~Cfg::FLAG)] = base + lrank[j] + is_start - 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) n_leaves[p] = s_base + n_split_leaves;
    RHCCQ_PROF(5);                                                 // leaf numbering
}

template <class Cfg, bool ROWS_SMEM, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == RHCCQ_SPLIT_THREADS ? 2 : 1)
rhccq_k_palette_split(rhccq_palette_batch B, const int* __restrict__ labels, const int* __restrict__ status_in,
                      const int* __restrict__ max_cpc, const double* __restrict__ rng, int rng_len,
                      int* __restrict__ leaf, int* __restrict__ n_leaves, int max_rows, int kc_s,
                      unsigned char* gws, size_t gws_stride, size_t small_bytes, size_t row_bytes) {
    RHCCQ_DYN_SMEM(dyn);
    // shared memory: [small tables][per-row arrays, if they fit][centre tables for kc_s centres]
    // global slice:  [per-row arrays][centre tables for max_rows centres]
    unsigned char* slice = gws ? gws + (size_t)blockIdx.x * gws_stride : nullptr;
    unsigned char* row_base = ROWS_SMEM ? dyn + small_bytes : slice;    // compile-time: keeps the address space known
    unsigned char* cent_s = dyn + small_bytes + (ROWS_SMEM ? row_bytes : 0);
    unsigned char* cent_g = slice ? slice + row_bytes : nullptr;
    double* ex = slice ? reinterpret_cast<double*>(slice + row_bytes + rhccq_split_center_bytes((size_t)max_rows)) : nullptr;
    // problems are handed out by a grid-wide cursor when there is one (palettes take unequal time); it lives in
    // the row area of workspace slice 0, which is unused while the rows are in shared memory
    int* cursor = (ROWS_SMEM && gws) ? reinterpret_cast<int*>(gws) : nullptr;
    __shared__ int s_next;
    for (int p = blockIdx.x; ; p += gridDim.x) {
        if (cursor) {
            if (threadIdx.x == 0) s_next = atomicAdd(cursor, 1);
            __syncthreads();
            p = s_next;
            __syncthreads();
        }
        if (p >= B.n_problems) break;
        rhccq_palette_split_problem<Cfg>(B, p, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves, max_rows,
                                         row_base, dyn, cent_s, kc_s, cent_g, max_rows, ex);
        __syncthreads();
    }
}

static size_t rhccq_split_small_bytes(int threads) {
#ifdef RHCCQ_HOST_EMU
    const size_t nw = 1; (void)threads;
#else
    const size_t nw = (size_t)threads / 32;
#endif
    return rhccq_carve_bytes((nw + 1) * RHCCQ_KM_CAND, 4)
           + rhccq_carve_bytes(nw * 8 * RHCCQ_KW, 8) + rhccq_carve_bytes(nw * 5 * RHCCQ_KW, 4)
           + rhccq_carve_bytes(nw * 4 * RHCCQ_KW, 4)
           + rhccq_carve_bytes(nw * RHCCQ_KC, 4);
}

template <class Cfg>
static int rhccq_launch_split_cfg(const rhccq_palette_batch& B, const int* labels, const int* status_in,
                                  const int* max_cpc, const double* rng, int rng_len, int* leaf, int* n_leaves,
                                  int max_rows, rhccq_launch_ws ws, void* stream) {
    const size_t rows = (size_t)(max_rows > 1 ? max_rows : 1);
    // few palettes (less than one per SM slot): a CTA of 512 threads each, they are latency-bound
    const bool big = B.n_problems <= rhccq_sm_count() * 2 && max_rows > 4200;
    const size_t small = rhccq_split_small_bytes(big ? RHCCQ_SPLIT_THREADS_BIG : RHCCQ_SPLIT_THREADS);
    const size_t row_bytes = rhccq_split_row_bytes<Cfg>(rows);
    const size_t kc_s = rows < RHCCQ_KC ? rows : RHCCQ_KC;
    const size_t cent_s = rhccq_split_center_bytes(kc_s);
    const size_t slice = rhccq_palette_split_ws_bytes(max_rows);
    const size_t slices = ws.ws ? ws.ws_bytes / slice : 0;
    // everything the kernel needs next to its ~6.4 KB of static shared memory, within the 227 KB of an SM
    const int rows_in_smem = small + row_bytes + cent_s + 7168 <= 227 * 1024;
    if (slices == 0) {
        rhccq_set_error("rhccq_palette_split: the workspace (%zu bytes) holds no slice of %zu bytes (float64 scratch, "
                        "centre tables%s)", ws.ws_bytes, slice, rows_in_smem ? "" : ", per-row working set");
        return -1;
    }
    // one CTA per problem, at most one per workspace slice (a CTA without a slice could not run a K-Means
    // with more than RHCCQ_KC centres); CTAs walk the problems with a grid stride
    int grid = B.n_problems;
    if (slices > 0 && (size_t)grid > slices) grid = (int)slices;
    if (!rows_in_smem) { const int cap = rhccq_sm_count() * 2; if (grid > cap) grid = cap; }
    const size_t smem = small + (rows_in_smem ? row_bytes : 0) + cent_s;
    unsigned char* gws = slices > 0 ? ws.ws : nullptr;
    if (rows_in_smem && gws) {                                     // the problem cursor (see the kernel)
#ifdef RHCCQ_HOST_EMU
        memset(gws, 0, 4);
#else
        cudaMemsetAsync(gws, 0, 4, (cudaStream_t)stream);
#endif
    }
#define RHCCQ_SPLIT_GO(ROWS, THREADS)                                                                                  \
    do {                                                                                                               \
        if (rhccq_smem_optin((const void*)rhccq_k_palette_split<Cfg, ROWS, THREADS>, smem) != 0) return -1;            \
        RHCCQ_LAUNCH((rhccq_k_palette_split<Cfg, ROWS, THREADS>), grid, THREADS, smem, (cudaStream_t)stream,           \
                     B, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves, max_rows, (int)kc_s,                 \
                     gws, slice, small, row_bytes);                                                                    \
    } while (0)
    if (rows_in_smem) { if (big) RHCCQ_SPLIT_GO(true, RHCCQ_SPLIT_THREADS_BIG); else RHCCQ_SPLIT_GO(true, RHCCQ_SPLIT_THREADS); }
    else { if (big) RHCCQ_SPLIT_GO(false, RHCCQ_SPLIT_THREADS_BIG); else RHCCQ_SPLIT_GO(false, RHCCQ_SPLIT_THREADS); }
#undef RHCCQ_SPLIT_GO
    return 0;
}

int rhccq_launch_palette_split(const rhccq_palette_batch& B, const int* labels, const int* status_in, const int* max_cpc,
                               const double* rng, int rng_len, int* leaf, int* n_leaves, int max_rows,
                               rhccq_launch_ws ws, void* stream) {
    if (B.n_problems <= 0) return 0;
    if (max_rows <= rhccq_cfg_small::MAX_ROWS)
        return rhccq_launch_split_cfg<rhccq_cfg_small>(B, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves,
                                                       max_rows, ws, stream);
    return rhccq_launch_split_cfg<rhccq_cfg_large>(B, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves,
                                                   max_rows, ws, stream);
}

#if defined(RHCCQ_SPLIT_PROFILE) && !defined(RHCCQ_HOST_EMU)
// tools/split_phases.py: copy out (reset != 0: clear) the phase counters
extern "C" int rhccq_split_prof_read(unsigned long long* host_out, int reset) {
    if (reset) { unsigned long long z[16] = {0}; return (int)cudaMemcpyToSymbol(rhccq_split_prof, z, sizeof z); }
    return (int)cudaMemcpyFromSymbol(host_out, rhccq_split_prof, 16 * sizeof(unsigned long long));
}
#endif
