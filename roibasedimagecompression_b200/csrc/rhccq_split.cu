// Cluster -> palette-entry assignment: one CTA per palette ("problem").
//
//   rhccq_k_palette_split    small clusters -> one entry each, large clusters ->
//                            recursive K-Means split
//                            (/root/reference/encoder/compression/clustering.py:253-355, :720-775),
//                            K-Means in the exact arithmetic of oracle/kmeans_restated.py
//                            (which cites the scikit-learn lines it restates).
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
// palette and writing one int per row: the kernel is bound by FP64 issue
// (distances are evaluated in IEEE double without contraction, like the oracle)
// and by barrier latency, not by memory bandwidth.
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"

// Optional phase timers (cycles summed over all CTAs into rhccq_split_prof[8]); compiled in with
// -DRHCCQ_SPLIT_PROFILE by tools/split_phases.py only.
#if defined(RHCCQ_SPLIT_PROFILE) && !defined(RHCCQ_HOST_EMU)
__device__ unsigned long long rhccq_split_prof[8];
#define RHCCQ_PROF_T0() long long prof_t_ = clock64()
#define RHCCQ_PROF(slot) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&rhccq_split_prof[slot], (unsigned long long)(n_ - prof_t_)); prof_t_ = n_; } } while (0)
#else
#define RHCCQ_PROF_T0() do {} while (0)
#define RHCCQ_PROF(slot) do {} while (0)
#endif

#ifndef RHCCQ_PRUNE_MIN_WORK
#define RHCCQ_PRUNE_MIN_WORK 20000     // n * k from which the pruned E step pays for its extra pass
#endif
#define RHCCQ_SPLIT_THREADS 256
#define RHCCQ_KM_MAXT 12               // 2 + int(log(k)) for k < 22027
#define RHCCQ_KC 128                   // centres of a CTA-level K-Means kept in shared memory
#define RHCCQ_KW 8                     // centres of a warp-level K-Means
#define RHCCQ_WARP_RANGE 1024          // largest range a single warp splits
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
        __syncthreads();                                          // scratch may still be read by a previous call
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {                   // unrolled: v stays in registers
            if (t < cnt) {
                long long x = v[t];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                if (lane == 0) sll[warp * RHCCQ_KM_MAXT + t] = x;
            }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
            if (t < cnt) {
                long long x = 0;
                for (int w = 0; w < nwarp; ++w) x += sll[w * RHCCQ_KM_MAXT + t];
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

// ---------------------------------------------------------------- K-Means (exact arithmetic)
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
    typename Cfg::cum_t* cum;           // seeding: inclusive cumulative sum of closest over the range
    typename Cfg::idx_t* label;         // Lloyd: cluster of every position
};
// ... and the per-call centre tables (k entries, private to the calling group).
struct rhccq_km_centers {
    double* center;                     // [3k]
    double* center_new;                 // [3k]
    double* term;                       // [k]
    int* sums;                          // [3k]
    int* cnt;                           // [k]
    int* hist;                          // [nsub * 4k] per-warp (r, g, b, count) accumulators, or nullptr
    int* cand;                          // [RHCCQ_KM_MAXT]
    int* poff;                          // [k * nsub] offsets of the counting partition, or nullptr (sort instead)
    int* wl;                            // worklist counter of the pruned E step, or nullptr (evaluate every centre)
};

__device__ __forceinline__ double rhccq_dist3(double x0, double x1, double x2, const double* c) {
    const double d0 = __dsub_rn(x0, c[0]), d1 = __dsub_rn(x1, c[1]), d2 = __dsub_rn(x2, c[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}

// First minimum over centres of ((d0^2 + d1^2) + d2^2) in IEEE double, for NB points at once: every
// centre is loaded once per NB points, and the NB distance chains are independent (the FP64 pipe is
// the bound of this kernel).
#define RHCCQ_EB 4
__device__ __forceinline__ void rhccq_nearest_centers(const uint32_t (&c)[RHCCQ_EB], const double* center, int k,
                                                      int (&bi)[RHCCQ_EB]) {
    double x0[RHCCQ_EB], x1[RHCCQ_EB], x2[RHCCQ_EB], best[RHCCQ_EB];
#pragma unroll
    for (int u = 0; u < RHCCQ_EB; ++u) {
        x0[u] = (double)rhccq_key_r(c[u]); x1[u] = (double)rhccq_key_g(c[u]); x2[u] = (double)rhccq_key_b(c[u]);
        best[u] = rhccq_dist3(x0[u], x1[u], x2[u], center);
        bi[u] = 0;
    }
    for (int q = 1; q < k; ++q) {
        const double c0 = center[3 * q], c1 = center[3 * q + 1], c2 = center[3 * q + 2];
#pragma unroll
        for (int u = 0; u < RHCCQ_EB; ++u) {
            const double d0 = __dsub_rn(x0[u], c0), d1 = __dsub_rn(x1[u], c1), d2 = __dsub_rn(x2[u], c2);
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            if (d < best[u]) { best[u] = d; bi[u] = q; }
        }
    }
}

// M-step accumulation of one point into a table of (r, g, b, count) rows / into the sums and counts.
// Plain shared-memory atomics: summing equal labels inside the warp first (match_any + reduce_add) was
// measured twice as slow on B200 as letting the atomic unit serialise the conflicts.
__device__ __forceinline__ void rhccq_acc_rows(int* rows, int stride, int bi, uint32_t c, bool valid) {
    if (valid) {
        int* h = rows + stride * bi;
        atomicAdd(h, rhccq_key_r(c)); atomicAdd(h + 1, rhccq_key_g(c)); atomicAdd(h + 2, rhccq_key_b(c)); atomicAdd(h + 3, 1);
    }
}
__device__ __forceinline__ void rhccq_acc_split(int* sums, int* cnt, int bi, uint32_t c, bool valid) {
    if (valid) {
        atomicAdd(&sums[3 * bi], rhccq_key_r(c)); atomicAdd(&sums[3 * bi + 1], rhccq_key_g(c));
        atomicAdd(&sums[3 * bi + 2], rhccq_key_b(c)); atomicAdd(&cnt[bi], 1);
    }
}

// Labels of KMeans(k, random_state=42, n_init='auto').fit_predict on the n colours at positions
// [lo, lo + n), as restated in oracle/kmeans_restated.py.  On return A.label holds the labels and
// C.cnt the cluster sizes.  Group-uniform control flow; every thread of the group must call.
template <class G, class Cfg>
__device__ __forceinline__ void rhccq_kmeans(const G& g, const rhccq_km_arrays<Cfg>& A, const rhccq_km_centers& C, int lo, int n, int k,
                             const double* __restrict__ rng) {
    typedef typename Cfg::cum_t cum_t;
    typedef typename Cfg::idx_t idx_t;
    const int tid = g.tid(), gsz = g.size();
    const uint32_t* x = A.x + lo;
    uint32_t* closest = A.closest + lo;
    cum_t* cum = A.cum + lo;
    idx_t* label = A.label + lo;
    // contiguous chunk of the caller (scans need a fixed element order)
    const int per = (n + gsz - 1) / gsz;
    const int c_lo = tid * per < n ? tid * per : n;
    const int c_hi = c_lo + per < n ? c_lo + per : n;

    RHCCQ_PROF_T0();
    // ---- k-means++ seeding (kmeans_restated.kmeans_pp_seeds)
    const int T = rhccq_kmeans_local_trials(k);
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
        for (int q = tid; q < 3; q += gsz) C.center[q] = (double)((cf >> (16 - 8 * q)) & 255u);
    }
    long long pot = acc[0];
    const long long S1[3] = {acc[1], acc[2], acc[3]}, S2[3] = {acc[4], acc[5], acc[6]};
    int ri = 1;
    for (int c = 1; c < k; ++c) {
        if (G::kCta) {
            // Candidates without materialising the cumulative sum: every thread publishes the sum of its chunk;
            // one warp per trial scans the chunk sums (shuffles), finds the chunk in which the inclusive sum
            // first reaches r * pot, then the element inside it — the same element a search of the cumulative
            // array gives (searchsorted(cum, r * pot, side='left'), clipped).  One barrier instead of five.
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
                        const unsigned long long inc = rhccq_warp_incl_scan_u64(j < j1 ? (unsigned long long)closest[j] : 0ull);
                        const unsigned mm = rhccq_ballot(j < j1 && (double)(pre + inc) >= rv);
                        if (mm != 0u) { found = base + __ffs((int)mm) - 1; break; }
                        pre += rhccq_shfl(inc, RHCCQ_WARP_SIZE - 1);
                    }
                }
                if (RHCCQ_LANE == 0) C.cand[t] = found < n - 1 ? found : n - 1;
            }
            g.sync();
        } else {
            // inclusive cumulative sum of closest (exact integers)
            cum_t total;
            cum_t run = g.template excl_scan<cum_t>(chunk_sum, &total);
            for (int j = c_lo; j < c_hi; ++j) { run += closest[j]; cum[j] = run; }
            g.sync();
            // candidates: searchsorted(cum, r * pot, side='left'), clipped
            for (int t = tid; t < T; t += gsz) {
                const double rv = __dmul_rn(rng[ri + t], (double)pot);
                int l = 0, h = n;                                   // first j with cum[j] >= rv
                while (l < h) {
                    const int mid = (l + h) >> 1;
                    if ((double)cum[mid] < rv) l = mid + 1; else h = mid;
                }
                C.cand[t] = l < n - 1 ? l : n - 1;
            }
            g.sync();
        }
        ri += T;
        uint32_t xc[RHCCQ_KM_MAXT];
#pragma unroll
        for (int t = 0; t < RHCCQ_KM_MAXT; ++t) { acc[t] = 0; xc[t] = t < T ? x[C.cand[t]] : 0u; }
        for (int j = c_lo; j < c_hi; ++j) {
            const uint32_t cj = x[j], o = closest[j];
#pragma unroll
            for (int t = 0; t < RHCCQ_KM_MAXT; ++t) {
                if (t < T) {
                    const uint32_t d = (uint32_t)rhccq_d2(cj, xc[t]);
                    acc[t] += d < o ? d : o;
                }
            }
        }
        g.sum_vec(acc, T);
        int best = 0;
        long long best_pot = acc[0];
#pragma unroll
        for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t < T && acc[t] < best_pot) { best_pot = acc[t]; best = t; }
        uint32_t cs = xc[0];
#pragma unroll
        for (int t = 1; t < RHCCQ_KM_MAXT; ++t) if (t == best) cs = xc[t];
        chunk_sum = 0;
        for (int j = c_lo; j < c_hi; ++j) {
            const uint32_t d = (uint32_t)rhccq_d2(x[j], cs);
            const uint32_t o = closest[j];
            const uint32_t m = d < o ? d : o;
            closest[j] = m;
            chunk_sum += m;
        }
        for (int q = tid; q < 3; q += gsz) C.center[3 * c + q] = (double)((cs >> (16 - 8 * q)) & 255u);
        pot = best_pot;
        // the next pass writes cum / cand only after a collective, which orders it after the reads above
    }

    if (g.size() > RHCCQ_WARP_SIZE) RHCCQ_PROF(1);
    // ---- tolerance (kmeans_restated.tolerance)
    double tol;
    {
        const double nn = __dmul_rn((double)n, (double)n);
        double v[3];
        for (int d = 0; d < 3; ++d) v[d] = __ddiv_rn((double)((long long)n * S2[d] - S1[d] * S1[d]), nn);
        tol = __dmul_rn(__ddiv_rn(__dadd_rn(__dadd_rn(v[0], v[1]), v[2]), 3.0), 1e-4);
    }

    // ---- Lloyd (kmeans_restated.kmeans_labels)
    //
    // The cluster sums of the M step are kept incrementally: a point only touches them when its label
    // changes (- old cluster, + new cluster; integers, so the result is the sum over the members exactly).
    // After the first few iterations a few percent of the points move, which takes the shared-memory
    // atomics — the main cost next to the distances — off the critical path.  An empty-cluster relocation
    // edits the sums without changing labels, so the iteration after one recounts from scratch.
    const idx_t NOLABEL = (idx_t)(Cfg::FLAG - 1u);                  // never a real label (k < FLAG - 1)
    for (int j = tid; j < n; j += gsz) label[j] = NOLABEL;
    for (int q = tid; q < k; q += gsz) { C.cnt[q] = 0; C.sums[3 * q] = 0; C.sums[3 * q + 1] = 0; C.sums[3 * q + 2] = 0; }
    double* cen = C.center;                                         // current / next centre tables, swapped per iteration
    double* cen_new = C.center_new;
    g.sync();
    bool strict = false, recount = true;                            // recount: add every point, subtract none
    for (int it = 0; it < 300; ++it) {
        // E step.  A point whose distance to the centre of its previous cluster is (safely) less than half
        // the distance from that centre to the nearest other centre keeps its cluster by the triangle
        // inequality — every other centre is strictly farther by a margin far above the rounding of the
        // distance evaluation — so it skips the loop over the centres.  The result is identical to
        // evaluating every centre; the other points (about 40 %) are compacted into a worklist so that the
        // warps of the full loop stay full.  Worth its extra pass only for large problems.
        int changed = 0;
        const bool prune = it > 0 && !recount && C.wl != nullptr && (long long)n * k >= RHCCQ_PRUNE_MIN_WORK;
        uint32_t* wl = reinterpret_cast<uint32_t*>(closest);       // dead after the seeding; positions of this range
        if (prune) {
            g.sync();                                               // slower threads may still be summing the last shift from term
            for (int q = tid; q < k; q += gsz) {
                const double* cq = cen + 3 * q;
                double m = 1.0e300;
                for (int r = 0; r < k; ++r) {
                    if (r == q) continue;
                    const double d = rhccq_dist3(cq[0], cq[1], cq[2], cen + 3 * r);
                    m = d < m ? d : m;
                }
                // usable only when the centres are at least 1 apart: then the margin below (1e-6 relative)
                // is thousands of times the rounding error of a distance (< 1e-9 absolute)
                C.term[q] = m >= 1.0 ? __dmul_rn(__dmul_rn(0.25, m), 0.999999) : 0.0;
            }
            if (tid == 0) *C.wl = 0;
            g.sync();
            for (int j = tid; j - RHCCQ_LANE < n; j += gsz) {          // warp-uniform trip count
                bool push = false;
                if (j < n) {
                    const uint32_t c = x[j];
                    const int a = (int)label[j];
                    const double u = rhccq_dist3((double)rhccq_key_r(c), (double)rhccq_key_g(c), (double)rhccq_key_b(c), cen + 3 * a);
                    push = !(u < C.term[a]);
                }
                const unsigned m = rhccq_ballot(push);                 // one counter bump per warp
                int base = 0;
                if (RHCCQ_LANE == 0 && m) base = atomicAdd(C.wl, __popc(m));
                base = rhccq_shfl(base, 0);
                if (push) wl[base + __popc(m & rhccq_lanemask_lt())] = (uint32_t)j;
            }
            g.sync();
        }
        const int n_work = prune ? *C.wl : n;
        for (int i0 = tid; i0 < n_work; i0 += RHCCQ_EB * gsz) {
            uint32_t cb[RHCCQ_EB];
            int bb[RHCCQ_EB], jj[RHCCQ_EB];
#pragma unroll
            for (int u = 0; u < RHCCQ_EB; ++u) {
                const int i = i0 + u * gsz;
                const int is = i < n_work ? i : i0;                     // lanes past the end redo an entry and drop it
                jj[u] = prune ? (int)wl[is] : is;
                cb[u] = x[jj[u]];
            }
            rhccq_nearest_centers(cb, cen, k, bb);
#pragma unroll
            for (int u = 0; u < RHCCQ_EB; ++u) {
                if (i0 + u * gsz >= n_work) continue;
                const int j = jj[u], bi = bb[u], old = (int)label[j];
                if (old != bi) { changed = 1; label[j] = (idx_t)bi; }
                if (old != bi || recount) {
                    const uint32_t c = cb[u];
                    const int r = rhccq_key_r(c), gg = rhccq_key_g(c), b = rhccq_key_b(c);
                    if (!recount && old != (int)NOLABEL) {
                        atomicAdd(&C.sums[3 * old], -r); atomicAdd(&C.sums[3 * old + 1], -gg); atomicAdd(&C.sums[3 * old + 2], -b);
                        atomicAdd(&C.cnt[old], -1);
                    }
                    atomicAdd(&C.sums[3 * bi], r); atomicAdd(&C.sums[3 * bi + 1], gg); atomicAdd(&C.sums[3 * bi + 2], b);
                    atomicAdd(&C.cnt[bi], 1);
                }
            }
        }
        recount = false;
        changed = g.any(changed);
        int empty = 0;
        for (int q = tid; q < k; q += gsz) if (C.cnt[q] == 0) ++empty;
        const int n_empty = k <= gsz ? g.count(empty) : g.sum_i(empty);    // k <= gsz: at most one cluster per thread
        if (n_empty > 0) {
            // relocate empty clusters to the points farthest from their centre
            // (_k_means_common.pyx:177-211): farthest first, ties to the lower index.
            double mx = 0.0;
            for (int j = tid; j < n; j += gsz) {
                const uint32_t c = x[j];
                const double d = rhccq_dist3((double)rhccq_key_r(c), (double)rhccq_key_g(c), (double)rhccq_key_b(c),
                                             cen + 3 * (int)label[j]);
                if (d > mx) mx = d;
            }
            mx = g.max_d(mx);
            if (mx != 0.0) {
                // the empty set is fixed before any relocation (a donor cluster may drop to zero later)
                for (int q = tid; q < k; q += gsz) C.term[q] = C.cnt[q] == 0 ? 1.0 : 0.0;
                g.sync();
                int e = 0;
                for (int done = 0; done < n_empty; ++done) {
                    while (C.term[e] == 0.0) ++e;                   // next empty cluster, ascending (group-uniform)
                    double bm = -1.0;
                    int bj = 0x7fffffff;
                    for (int j = tid; j < n; j += gsz) {
                        const uint32_t lj = label[j];
                        if (lj & Cfg::FLAG) continue;               // already taken
                        const uint32_t c = x[j];
                        const double d = rhccq_dist3((double)rhccq_key_r(c), (double)rhccq_key_g(c),
                                                     (double)rhccq_key_b(c), cen + 3 * (int)lj);
                        if (d > bm) { bm = d; bj = j; }             // ascending j: the first maximum stays
                    }
                    const double gm = g.max_d(bm);
                    bj = g.min_i(bm == gm ? bj : 0x7fffffff);
                    if (tid == 0) {
                        const int old = (int)label[bj];
                        const uint32_t c = x[bj];
                        C.sums[3 * old] -= rhccq_key_r(c); C.sums[3 * old + 1] -= rhccq_key_g(c); C.sums[3 * old + 2] -= rhccq_key_b(c);
                        C.sums[3 * e] = rhccq_key_r(c); C.sums[3 * e + 1] = rhccq_key_g(c); C.sums[3 * e + 2] = rhccq_key_b(c);
                        C.cnt[e] = 1;
                        C.cnt[old] -= 1;
                        label[bj] = (idx_t)(old | Cfg::FLAG);
                    }
                    g.sync();
                    ++e;
                }
                for (int j = tid; j < n; j += gsz) label[j] = (idx_t)(label[j] & ~Cfg::FLAG);
                recount = true;                                     // the sums no longer follow the labels
                g.sync();
            }
        }
        for (int q = tid; q < k; q += gsz) {
            double c0, c1, c2;
            if (C.cnt[q] > 0) {
                const double cn = (double)C.cnt[q];
                c0 = __ddiv_rn((double)C.sums[3 * q], cn);
                c1 = __ddiv_rn((double)C.sums[3 * q + 1], cn);
                c2 = __ddiv_rn((double)C.sums[3 * q + 2], cn);
            } else {
                c0 = __ddiv_rn((double)S1[0], (double)n);
                c1 = __ddiv_rn((double)S1[1], (double)n);
                c2 = __ddiv_rn((double)S1[2], (double)n);
            }
            cen_new[3 * q] = c0; cen_new[3 * q + 1] = c1; cen_new[3 * q + 2] = c2;
            const double a0 = __dsub_rn(c0, cen[3 * q]), a1 = __dsub_rn(c1, cen[3 * q + 1]), a2 = __dsub_rn(c2, cen[3 * q + 2]);
            C.term[q] = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
            if (recount) { C.cnt[q] = 0; C.sums[3 * q] = 0; C.sums[3 * q + 1] = 0; C.sums[3 * q + 2] = 0; }
        }
        g.sync();
        double shift = 0.0;
        if (k <= 64) {
            for (int q = 0; q < k; ++q) shift = __dadd_rn(shift, C.term[q]);             // fixed order, every thread
        } else {
            if (tid == 0) for (int q = 0; q < k; ++q) shift = __dadd_rn(shift, C.term[q]);
            shift = g.bcast_d(shift);
        }
        { double* t = cen; cen = cen_new; cen_new = t; }
        if (!changed) { strict = true; break; }
        if (shift <= tol) break;
    }
    if (!strict) {
        for (int j0 = tid; j0 < n; j0 += RHCCQ_EB * gsz) {
            uint32_t cb[RHCCQ_EB];
            int bb[RHCCQ_EB];
#pragma unroll
            for (int u = 0; u < RHCCQ_EB; ++u) { const int j = j0 + u * gsz; cb[u] = x[j < n ? j : j0]; }
            rhccq_nearest_centers(cb, cen, k, bb);
#pragma unroll
            for (int u = 0; u < RHCCQ_EB; ++u) { const int j = j0 + u * gsz; if (j < n) label[j] = (idx_t)bb[u]; }
        }
    }
    for (int q = tid; q < k; q += gsz) C.cnt[q] = 0;
    g.sync();
    for (int j = tid; j < n; j += gsz) atomicAdd(&C.cnt[(int)label[j]], 1);
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
    return rhccq_carve_bytes(3 * kc, 8) * 2 + rhccq_carve_bytes(kc, 8) + rhccq_carve_bytes(3 * kc, 4)
           + rhccq_carve_bytes(kc, 4);
}

size_t rhccq_palette_split_ws_bytes(int max_rows) {
    // one slice of the global workspace: the per-row arrays and centre tables for k up to max_rows
    const size_t r = (size_t)(max_rows > 1 ? max_rows : 1);
    const size_t rows = max_rows <= rhccq_cfg_small::MAX_ROWS ? rhccq_split_row_bytes<rhccq_cfg_small>(r)
                                                              : rhccq_split_row_bytes<rhccq_cfg_large>(r);
    // never 0 and never within the shared-memory budget: the caller must always pass a workspace (the
    // centre tables of a K-Means with more than RHCCQ_KC centres live there)
    const size_t need = rows + rhccq_split_center_bytes(r);
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
                                            unsigned char* cent_g, int kc_g) {
    typedef typename Cfg::idx_t idx_t;
    typedef typename Cfg::q_t q_t;
    __shared__ long long s_ll[RHCCQ_MAX_WARPS * RHCCQ_KM_MAXT + 2];
    __shared__ int s_scan[RHCCQ_MAX_WARPS + 2];
    __shared__ int s_tail, s_err, s_base, s_claim, s_wl;
    __shared__ int wlc[RHCCQ_MAX_WARPS];
    RHCCQ_PROF_T0();
    const int n = B.pal_cnt[p];
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    const int* lab = labels + B.pal_off[p];
    int* lf = leaf + B.pal_off[p];
    const int mcpc = max_cpc[p];
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
    }
    // CTA-level centre tables: in shared memory for k <= kc_s, else in the global workspace (k <= kc_g)
    rhccq_km_centers CS, CG;
    rhccq_carve_centers(CS, cent_s, (size_t)kc_s);
    rhccq_carve_centers(CG, cent_g, (size_t)kc_g);
    // small shared tables: per-warp centre sets, per-warp M-step histograms, candidate slots
    rhccq_carver sv(small_base);
    const size_t nw = (size_t)RHCCQ_NWARPS;                        // the tables are sized for the launch's warps
    int* cand = sv.take<int>((nw + 1) * RHCCQ_KM_MAXT);
    double* wcent = sv.take<double>(nw * 7 * RHCCQ_KW);
    int* wint = sv.take<int>(nw * 5 * RHCCQ_KW);
    int* poff = sv.take<int>(nw * RHCCQ_KC);
    CS.cand = CG.cand = cand + nw * RHCCQ_KM_MAXT;

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
    rhccq_grp_cta gc; gc.sll = s_ll;
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
            int k = (cnt + mcpc - 1) / mcpc;                        // clustering.py:739-742
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
                C.hist = nullptr;
                C.poff = poff;
                C.wl = &s_wl;
                rhccq_split_range<rhccq_grp_cta, Cfg>(gc, W, C, lo, hi, k, mcpc, rng, &s_tail, max_rows, &s_err);
            } else {
                rhccq_km_centers C = CG;
                C.hist = nullptr;
                C.poff = nullptr;
                C.wl = nullptr;
                rhccq_split_range<rhccq_grp_cta, Cfg>(gc, W, C, lo, hi, k, mcpc, rng, &s_tail, max_rows, &s_err);
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
            int k = (cnt + mcpc - 1) / mcpc;
            if (k < 2) k = 2;
            if (k > cnt) k = cnt;
            if (!(cnt <= RHCCQ_WARP_RANGE && k <= RHCCQ_KW)) continue;
            if (1 + (k - 1) * rhccq_kmeans_local_trials(k) > rng_len) { s_err = 2; continue; }
            rhccq_km_centers C;
            double* wc = wcent + (size_t)RHCCQ_WARP * 7 * RHCCQ_KW;
            int* wi = wint + (size_t)RHCCQ_WARP * 5 * RHCCQ_KW;
            C.center = wc; C.center_new = wc + 3 * RHCCQ_KW; C.term = wc + 6 * RHCCQ_KW;
            C.sums = wi; C.cnt = wi + 3 * RHCCQ_KW; C.poff = wi + 4 * RHCCQ_KW;
            C.hist = nullptr;
            C.cand = cand + RHCCQ_WARP * RHCCQ_KM_MAXT;
            C.wl = wlc + RHCCQ_WARP;
            rhccq_split_range<rhccq_grp_warp, Cfg>(gw, W, C, lo, hi, k, mcpc, rng, &s_tail, max_rows, &s_err);
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
                                         row_base, dyn, cent_s, kc_s, cent_g, max_rows);
        __syncthreads();
    }
}

static size_t rhccq_split_small_bytes(int threads) {
#ifdef RHCCQ_HOST_EMU
    const size_t nw = 1; (void)threads;
#else
    const size_t nw = (size_t)threads / 32;
#endif
    return rhccq_carve_bytes((nw + 1) * RHCCQ_KM_MAXT, 4)
           + rhccq_carve_bytes(nw * 7 * RHCCQ_KW, 8) + rhccq_carve_bytes(nw * 5 * RHCCQ_KW, 4)
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
    // everything the kernel needs next to its ~3.3 KB of static shared memory, within the 227 KB of an SM
    const int rows_in_smem = small + row_bytes + cent_s + 4096 <= 227 * 1024;
    if (!rows_in_smem && slices == 0) {
        rhccq_set_error("rhccq_palette_split: the per-row working set (%zu bytes) exceeds shared memory and the "
                        "workspace (%zu bytes) holds no slice of %zu bytes", row_bytes, ws.ws_bytes, slice);
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
    if (reset) { unsigned long long z[8] = {0}; return (int)cudaMemcpyToSymbol(rhccq_split_prof, z, sizeof z); }
    return (int)cudaMemcpyFromSymbol(host_out, rhccq_split_prof, 8 * sizeof(unsigned long long));
}
#endif
