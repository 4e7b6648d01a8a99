// The >= 10 000-colour branch of the palette clustering:
//   MiniBatchKMeans(n_clusters=ceil(n*q/100/10), batch_size=1000, random_state=42, n_init='auto')
//       .fit_predict(non_black_colours.astype(float))
// (/root/reference/encoder/compression/clustering.py:207-218), in the exact arithmetic and with the exact
// consumption of the RandomState(42) stream of oracle/minibatch_restated.py (which cites the scikit-learn
// lines: _kmeans.py:2056-2229, :1566-1684, :1974-2054, _k_means_minibatch.pyx:68-118).
//
// Shape of the work: a handful of palettes per batch (stage 2 of large images), each a strictly sequential
// chain of ~50-100 mini-batch steps on 1000 points, preceded by a k-means++ seeding on 3000 points and
// followed by one assignment of all n colours.  One CTA walks the chain of a palette (the Mersenne Twister,
// numpy's randint / choice / shuffle and the inertia sum are sequential by definition and run on thread 0);
// the final assignment, which is the only part proportional to n*k, is a separate kernel over all SMs.
#include "rhccq_common.cuh"
#include "rhccq_sklearn.cuh"
#include "rhccq_kernels.h"

// Optional phase timers (cycles of rank 0 / thread 0 of every cluster, summed into rhccq_mb_prof[16]); compiled in
// with -DRHCCQ_MB_PROFILE by tools/minibatch_phases.py only.
#if defined(RHCCQ_MB_PROFILE) && !defined(RHCCQ_HOST_EMU)
__device__ unsigned long long rhccq_mb_prof[40];
#define RHCCQ_MBP_T0() long long mbp_t_ = clock64(); const long long mbp_start_ = mbp_t_
#define RHCCQ_MBP_END() do { if (rank == 0 && threadIdx.x == 0) atomicMax(&rhccq_mb_prof[15], (unsigned long long)(clock64() - mbp_start_)); } while (0)
#define RHCCQ_MBP(slot) do { if (rank == 0 && threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&rhccq_mb_prof[slot], (unsigned long long)(n_ - mbp_t_)); mbp_t_ = n_; } } while (0)
#define RHCCQ_MBP_ADD(slot, v) do { if (rank == 0 && threadIdx.x == 0) atomicAdd(&rhccq_mb_prof[slot], (unsigned long long)(v)); } while (0)
#define RHCCQ_MBP_ANY(slot, v) atomicAdd(&rhccq_mb_prof[slot], (unsigned long long)(v))
#define RHCCQ_MBP_RANK_T0() const long long mbr_t_ = clock64()
#define RHCCQ_MBP_RANK(slot0) do { if (threadIdx.x == 0) atomicAdd(&rhccq_mb_prof[(slot0) + rank], (unsigned long long)(clock64() - mbr_t_)); } while (0)
#else
#define RHCCQ_MBP_T0() do {} while (0)
#define RHCCQ_MBP_END() do {} while (0)
#define RHCCQ_MBP(slot) do {} while (0)
#define RHCCQ_MBP_ADD(slot, v) do {} while (0)
#define RHCCQ_MBP_ANY(slot, v) do {} while (0)
#define RHCCQ_MBP_RANK_T0() do {} while (0)
#define RHCCQ_MBP_RANK(slot0) do {} while (0)
#endif
#define RHCCQ_MB_THREADS 512
#define RHCCQ_MB_BATCH 1000
#define RHCCQ_MB_MAXT 12
#define RHCCQ_MB_SEED_GROUP 32       // seeding steps whose random numbers are drawn in one go
#define RHCCQ_MB_SEED_CAP 3072      // subset sizes up to this are seeded out of shared memory (3 * batch = 3 000 is the usual one)
#define RHCCQ_MB_GROUP_CAP 1024       // clusters per CTA up to which the centre update groups the batch in shared memory
#define RHCCQ_MB_CLUSTER 8           // CTAs (SMs) that walk one palette together

// ---------------------------------------------------------------- MT19937 as numpy.random.RandomState(42)
struct rhccq_mt { uint32_t* s; int* pos; };
__device__ __forceinline__ void rhccq_mt_seed(const rhccq_mt& m, uint32_t seed) {     // init_genrand
    m.s[0] = seed;
    for (int i = 1; i < 624; ++i) m.s[i] = 1812433253u * (m.s[i - 1] ^ (m.s[i - 1] >> 30)) + (uint32_t)i;
    *m.pos = 624;
}
__device__ __forceinline__ uint32_t rhccq_mt_next(const rhccq_mt& m) {
    if (*m.pos >= 624) {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (m.s[k] & 0x80000000u) | (m.s[(k + 1) % 624] & 0x7fffffffu);
            m.s[k] = m.s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        *m.pos = 0;
    }
    uint32_t y = m.s[(*m.pos)++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
__device__ __forceinline__ uint32_t rhccq_mt_temper(uint32_t y) {
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
// The next `count` outputs of the generator into raw[], by the whole CTA: tempering is independent per
// element; the state refill ("twist") has a dependency structure that one warp resolves in four phases
// (elements 0..226 read only old values, 227..453 and 454..622 read values renewed one phase earlier,
// element 623 reads the renewed elements 0 and 396).  Every thread must call.
__device__ __forceinline__ void rhccq_mt_twist_warp0(const rhccq_mt& m) {
    if (RHCCQ_WARP == 0) {
        const int lo[4] = {0, 227, 454, 623}, hi[4] = {227, 454, 623, 624};
        for (int ph = 0; ph < 4; ++ph) {
            for (int k0 = lo[ph]; k0 < hi[ph]; k0 += RHCCQ_WARP_SIZE) {
                const int k = k0 + RHCCQ_LANE;
                uint32_t v = 0;
                if (k < hi[ph]) {
                    const uint32_t y = (m.s[k] & 0x80000000u) | (m.s[(k + 1) % 624] & 0x7fffffffu);
                    v = m.s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
                }
                __syncwarp();                              // all reads of this round before its writes
                if (k < hi[ph]) m.s[k] = v;
                __syncwarp();
            }
        }
    }
}
__device__ void rhccq_mt_fill_raw(const rhccq_mt& m, uint32_t* raw, int count) {
    int done = 0;
    while (done < count) {
        __syncthreads();
        int pos = *m.pos;
        if (pos >= 624) {
            rhccq_mt_twist_warp0(m);
            __syncthreads();
            pos = 0;
        }
        const int take = 624 - pos < count - done ? 624 - pos : count - done;
        RHCCQ_PAR_FOR(i, take) raw[done + i] = rhccq_mt_temper(m.s[pos + i]);
        __syncthreads();
        if (threadIdx.x == 0) *m.pos = pos + take;
        done += take;
    }
    __syncthreads();
}
// For consumers whose number of draws depends on the values (masked rejection): every thread calls; the unread
// outputs of the generator's current block (after a refill if it is exhausted) are tempered into buf[0..return)
// and the position is left where it was — the one consuming thread adds what it used.
__device__ int rhccq_mt_peek_block(const rhccq_mt& m, uint32_t* buf) {
    __syncthreads();
    int pos = *m.pos;
    if (pos >= 624) {
        rhccq_mt_twist_warp0(m);
        __syncthreads();
        if (threadIdx.x == 0) *m.pos = 0;
        pos = 0;
    }
    RHCCQ_PAR_FOR(i, 624 - pos) buf[i] = rhccq_mt_temper(m.s[pos + i]);
    __syncthreads();
    return 624 - pos;
}
// `count` draws of randint(0, max + 1) (masked rejection, as rhccq_mt_below), in order, into out[] (or dropped)
__device__ void rhccq_mt_below_seq(const rhccq_mt& m, uint32_t max, int count, int* out, uint32_t* buf, int* s_prog) {
    if (max == 0) { if (out) RHCCQ_PAR_FOR(i, count) out[i] = 0; __syncthreads(); return; }
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    if (threadIdx.x == 0) *s_prog = 0;
    while (true) {
        const int avail = rhccq_mt_peek_block(m, buf);
        if (threadIdx.x == 0) {
            int i = *s_prog, r = 0;
            while (i < count && r < avail) {
                const uint32_t v = buf[r++] & mask;
                if (v <= max) { if (out) out[i] = (int)v; ++i; }
            }
            *m.pos += r;
            *s_prog = i;
        }
        __syncthreads();
        if (*s_prog >= count) break;
    }
    __syncthreads();
}
__device__ __forceinline__ double rhccq_mt_double(const rhccq_mt& m) {                // random_sample
    const uint32_t a = rhccq_mt_next(m) >> 5, b = rhccq_mt_next(m) >> 6;
    return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
}
__device__ __forceinline__ uint32_t rhccq_mt_below(const rhccq_mt& m, uint32_t max) {  // uniform in [0, max], masked rejection
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (rhccq_mt_next(m) & mask)) > max) {}
    return v;
}

__device__ __forceinline__ double rhccq_mb_dist(double x0, double x1, double x2, const double* c) {
    const double d0 = __dsub_rn(x0, c[0]), d1 = __dsub_rn(x1, c[1]), d2 = __dsub_rn(x2, c[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}

// inclusive running sum along the lanes of a warp
__device__ __forceinline__ unsigned long long rhccq_mb_warp_incl_scan(unsigned long long v) {
#ifndef RHCCQ_HOST_EMU
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
#endif
    return v;
}
// sum over the lanes of a warp, in every lane
__device__ __forceinline__ uint32_t rhccq_mb_warp_sum(uint32_t v) {
#ifdef RHCCQ_HOST_EMU
    return v;
#else
    return __reduce_add_sync(0xffffffffu, v);
#endif
}
__device__ __forceinline__ unsigned long long rhccq_mb_warp_sum(unsigned long long v) {
#ifndef RHCCQ_HOST_EMU
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
#endif
    return v;
}

// number of clusters: ceil(n * (q / 100) / 10) as Python evaluates it (clustering.py:210)
__host__ __device__ static inline int rhccq_mb_k(int n, double q) {
    return (int)ceil(((double)n * (q / 100.0)) / 10.0);
}
__host__ __device__ static inline int rhccq_mb_init_size(int n, int k) {
    int batch = n < RHCCQ_MB_BATCH ? n : RHCCQ_MB_BATCH;
    long long s = 3LL * batch;
    if (s < k) s = 3LL * k;
    return (int)(s < n ? s : n);
}

struct rhccq_mb_ws {
    int* nb;                 // [n] rows of the non-black colours
    double* cdf;             // [n]
    int* sub;                // [init] rows (in nb order) of the init subset; later the batch
    uint32_t* closest;       // [init]
    unsigned long long* cum; // [init]
    double* center;          // [3k]
    double* center_new;      // [3k]
    double* counts;          // [k]
    unsigned long long* skey;// [pow2(k)]
    int* flag;               // [k] to_reassign, then its exclusive scan
    int* perm;               // [batch]
    uint32_t* perm_xs;       // [init] colours of the init subset
    float* cf;               // [cluster size][4k] per CTA: (-2 (c - 128), |c - 128|^2) in float32, first level of the batch labels
};
__host__ __device__ static inline size_t rhccq_mb_bytes(size_t n, size_t kmax, size_t init_max) {
    size_t k2 = 1;
    while (k2 < kmax) k2 <<= 1;
    return rhccq_carve_bytes(n, 4) + rhccq_carve_bytes(n, 8) + rhccq_carve_bytes(init_max, 4) * 2 + rhccq_carve_bytes(init_max, 8)
           + rhccq_carve_bytes(3 * kmax, 8) * 2 + rhccq_carve_bytes(kmax, 8) + rhccq_carve_bytes(k2, 8)
           + rhccq_carve_bytes(kmax + 1, 4) + rhccq_carve_bytes(RHCCQ_MB_BATCH, 4) + rhccq_carve_bytes(init_max, 4)
           + rhccq_carve_bytes(16, 4) + rhccq_carve_bytes(RHCCQ_MB_BATCH, 4) + rhccq_carve_bytes(RHCCQ_MB_BATCH, 8)
           + rhccq_carve_bytes((size_t)RHCCQ_MB_CLUSTER * 4 * kmax, 4);
}
size_t rhccq_palette_minibatch_ws_bytes(int max_rows) {
    // k <= n / 10 (q <= 100); init subset <= max(3000, 3k)
    const size_t n = (size_t)(max_rows > 1 ? max_rows : 1), kmax = n / 10 + 2;
    const size_t init = 3 * kmax > 3 * RHCCQ_MB_BATCH ? 3 * kmax : 3 * RHCCQ_MB_BATCH;
    return rhccq_mb_bytes(n, kmax, init < n ? init : n);
}

// Barrier over the CTAs of a thread-block cluster (release / acquire at cluster scope: global-memory writes
// before it are visible to every CTA of the cluster after it).  One CTA per "cluster" in the emulation build.
__device__ __forceinline__ void rhccq_cluster_sync() {
#ifdef RHCCQ_HOST_EMU
    __syncthreads();
#else
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void rhccq_cluster_rank(int* rank, int* size, int* id) {
#ifdef RHCCQ_HOST_EMU
    *rank = 0; *size = 1; *id = (int)blockIdx.x;
#else
    unsigned r, n, c;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(c));
    *rank = (int)r; *size = (int)n; *id = (int)c;
#endif
}

// One palette, walked by the CTAs of one cluster.  The seeding and everything that consumes the random stream
// or is sequential by definition (batch selection, inertia, reassignment, convergence) runs on the CTA of rank
// 0; the two phases proportional to batch x k — the labels of the batch and the centre update — are spread
// over all CTAs, exchanging through the global workspace between cluster barriers.
__device__ __forceinline__ void rhccq_minibatch_problem(const rhccq_palette_batch& B, int p, const double* __restrict__ quality,
                                        int* __restrict__ n_clusters, int max_rows, unsigned char* wsbase,
                                        double* __restrict__ centers_out, int* __restrict__ labels, int rank, int csize) {
    __shared__ uint32_t s_mt[624];
    __shared__ int s_mtpos;
    __shared__ int s_i[RHCCQ_MAX_WARPS + 2 + 16];
    __shared__ long long s_ll[RHCCQ_MAX_WARPS * RHCCQ_MB_MAXT + 2];
    __shared__ double s_d[RHCCQ_MAX_WARPS + 2];
    // one buffer, two tenants: the seeding keeps the colours and closest distances of the (usual) 3 000-point
    // subset and the cumulative sums of the threads' chunks here; the mini-batch steps their labels / values / colours
    __shared__ unsigned long long s_raw8[RHCCQ_MB_SEED_CAP + RHCCQ_MB_THREADS];
    unsigned char* s_raw = reinterpret_cast<unsigned char*>(s_raw8);
    double* s_own = reinterpret_cast<double*>(s_raw);                                   // [RHCCQ_MB_BATCH]
    int* s_lab = reinterpret_cast<int*>(s_raw + RHCCQ_MB_BATCH * 8);                    // [RHCCQ_MB_BATCH]
    uint32_t* s_col = reinterpret_cast<uint32_t*>(s_raw + RHCCQ_MB_BATCH * 12);         // [RHCCQ_MB_BATCH]
    unsigned long long* s_chunk = reinterpret_cast<unsigned long long*>(s_raw);         // [RHCCQ_MB_THREADS]
    uint32_t* s_xs = reinterpret_cast<uint32_t*>(s_raw + RHCCQ_MB_THREADS * 8);         // [RHCCQ_MB_SEED_CAP]
    uint32_t* s_closest = s_xs + RHCCQ_MB_SEED_CAP;                                     // [RHCCQ_MB_SEED_CAP]
    uint32_t* s_rbuf = reinterpret_cast<uint32_t*>(s_raw) + 1024;       // 624 raw generator outputs, beside a 1000-int array (never while the seeding arrays live)
    __shared__ int s_prog;
    __shared__ uint32_t s_rvraw[2 * RHCCQ_MB_MAXT * RHCCQ_MB_SEED_GROUP];
    __shared__ int s_cand[RHCCQ_MB_MAXT];
    __shared__ double s_val;
    __shared__ unsigned long long s_wtot[RHCCQ_MAX_WARPS];
    const int n_all = B.pal_cnt[p];
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    const size_t kmax = (size_t)max_rows / 10 + 2;
    size_t init_max = 3 * kmax > 3 * RHCCQ_MB_BATCH ? 3 * kmax : 3 * RHCCQ_MB_BATCH;
    if (init_max > (size_t)max_rows) init_max = (size_t)max_rows;
    rhccq_mb_ws W;
    int* hdr; int* lab_g; double* own_g;
    {
        rhccq_carver cv(wsbase);
        size_t k2 = 1;
        while (k2 < kmax) k2 <<= 1;
        W.nb = cv.take<int>(max_rows);
        W.cdf = cv.take<double>(max_rows);
        W.sub = cv.take<int>(init_max);
        W.closest = cv.take<uint32_t>(init_max);
        W.cum = cv.take<unsigned long long>(init_max);
        W.center = cv.take<double>(3 * kmax);
        W.center_new = cv.take<double>(3 * kmax);
        W.counts = cv.take<double>(kmax);
        W.skey = cv.take<unsigned long long>(k2);
        W.flag = cv.take<int>(kmax + 1);
        W.perm = cv.take<int>(RHCCQ_MB_BATCH);
        W.perm_xs = cv.take<uint32_t>(init_max);
        hdr = cv.take<int>(16);
        lab_g = cv.take<int>(RHCCQ_MB_BATCH);
        own_g = cv.take<double>(RHCCQ_MB_BATCH);
        W.cf = cv.take<float>((size_t)RHCCQ_MB_CLUSTER * 4 * kmax);
    }
    const rhccq_mt mt = {s_mt, &s_mtpos};
    int status = 0, n = 0, k = 0, batch = 0;
    RHCCQ_MBP_T0();
    if (rank == 0) {
        do {
            if (n_all > max_rows) { status = -1; break; }
    // non-black rows in row order (clustering.py:185-192)
    int* mark = reinterpret_cast<int*>(W.cdf);                      // n ints fit into n doubles
    RHCCQ_PAR_FOR(i, n_all) mark[i] = keys[i] != 0u ? 1 : 0;
    __syncthreads();
    n = rhccq_block_excl_scan_array<int>(mark, n_all, s_i);
    // (the label slot of a row carries its position among the non-black rows until rhccq_k_minibatch_assign
    // replaces it: scikit-learn's E step depends on where in its chunk of 256 a sample sits)
    RHCCQ_PAR_FOR(i, n_all) if (keys[i] != 0u) { W.nb[mark[i]] = i; labels[B.pal_off[p] + i] = mark[i]; }
    __syncthreads();
    k = rhccq_mb_k(n, quality[p]);
    if (k < 1 || k > n || (size_t)k > kmax - 1) { status = -1; break; }
    batch = n < RHCCQ_MB_BATCH ? n : RHCCQ_MB_BATCH;
    const int init_size = rhccq_mb_init_size(n, k);
    // ---- RandomState(42); validation indices are drawn and dropped (n_init == 1); init subset
    if (threadIdx.x == 0) rhccq_mt_seed(mt, 42u);
    rhccq_mt_below_seq(mt, (uint32_t)(n - 1), init_size, nullptr, s_rbuf, &s_prog);    // _kmeans.py:2110
    if (init_size < n) rhccq_mt_below_seq(mt, (uint32_t)(n - 1), init_size, W.sub, s_rbuf, &s_prog);
    else RHCCQ_PAR_FOR(i, n) W.sub[i] = i;
    __syncthreads();
    const int ns = init_size;
    // colours of the subset, contiguous (the k-means++ passes below read them k * (T + 2) times)
    const bool seed_smem = ns <= RHCCQ_MB_SEED_CAP;
    uint32_t* xs = seed_smem ? s_xs : reinterpret_cast<uint32_t*>(W.perm_xs);
    uint32_t* closest = seed_smem ? s_closest : W.closest;
    unsigned long long* chunk_incl = seed_smem ? s_chunk : W.cum;      // inclusive sums of the threads' chunks (blockDim entries)
    RHCCQ_PAR_FOR(j, ns) xs[j] = keys[W.nb[W.sub[j]]];
    __syncthreads();
    // ---- k-means++ on the subset (_kmeans.py:216-282): first centre by choice(ns, p=uniform)
    if (threadIdx.x == 0) {
        // cdf = cumsum(ones / ns) / last;  searchsorted(cdf, u, side='right')
        const double pi = __ddiv_rn(1.0, (double)ns);
        double run = 0.0;
        for (int i = 0; i < ns; ++i) { run = __dadd_rn(run, pi); W.cdf[i] = run; }
        const double last = W.cdf[ns - 1];
        const double u = rhccq_mt_double(mt);
        int lo = 0, hi = ns;                                       // first i with cdf[i] / last > u
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ddiv_rn(W.cdf[mid], last) <= u) lo = mid + 1; else hi = mid; }
        s_cand[0] = lo < ns - 1 ? lo : ns - 1;
    }
    __syncthreads();
    RHCCQ_MBP(0);                                                   // setup: rows, random subset
    const int T = 2 + (k >= 3) + (k >= 8) + (k >= 21) + (k >= 55) + (k >= 149) + (k >= 404) + (k >= 1097) + (k >= 2981)
                  + (k >= 8104) + (k >= 22027);
    // every thread owns a contiguous chunk of the subset (the cumulative sum needs a fixed order)
    const int per = (ns + (int)blockDim.x - 1) / (int)blockDim.x;
    const int c_lo = (int)threadIdx.x * per < ns ? (int)threadIdx.x * per : ns;
    const int c_hi = c_lo + per < ns ? c_lo + per : ns;
    unsigned long long chunk_sum = 0;
    {
        const uint32_t cf = xs[s_cand[0]];
        for (int j = c_lo; j < c_hi; ++j) { const uint32_t d = (uint32_t)rhccq_d2(xs[j], cf); closest[j] = d; chunk_sum += d; }
        RHCCQ_PAR_FOR(q, 3) W.center[q] = (double)((cf >> (16 - 8 * q)) & 255u);
    }
    const unsigned long long pot0 = (unsigned long long)rhccq_block_sum<long long>((long long)chunk_sum, s_ll);
    // Three block barriers per centre: [scan of the threads' chunk sums inside every warp] | [T threads search their
    // candidate: warp totals, lanes of one warp, elements of one chunk] | [every thread's share of the T potentials,
    // added up inside the warp] | [every warp adds the warps' shares and knows the winner].  S is the type the sums
    // are carried in: 32 bits (one REDUX per warp sum) when the whole subset's potential fits, else 64.
    auto seed_loop = [&](auto s_zero) {
        using S = decltype(s_zero);
        S* s_sv = reinterpret_cast<S*>(s_ll);                       // [warps][RHCCQ_MB_MAXT]
        const int lane = RHCCQ_LANE, warp = RHCCQ_WARP, nw = RHCCQ_NWARPS;
        unsigned long long pot = pot0;
        // the distances to the centre chosen last are folded into `closest` by the next pass over the points (and by
        // the search on the elements it looks at) instead of by a pass of their own
        bool pending = false;
        uint32_t pend = 0u;
        for (int c = 1; c < k; ++c) {
            const unsigned long long incl = rhccq_mb_warp_incl_scan(chunk_sum);
            chunk_incl[threadIdx.x] = incl;                         // inclusive inside the warp
            if (lane == RHCCQ_WARP_SIZE - 1) s_wtot[warp] = incl;
            // uniform(size=T) of this step: the raw outputs of RHCCQ_MB_SEED_GROUP steps are drawn together by the
            // whole CTA (nothing else consumes the stream during the seeding), two per double, in order
            const int gi = (c - 1) % RHCCQ_MB_SEED_GROUP;
            if (gi == 0) {
                const int steps = k - c < RHCCQ_MB_SEED_GROUP ? k - c : RHCCQ_MB_SEED_GROUP;
                rhccq_mt_fill_raw(mt, s_rvraw, 2 * T * steps);
            }
            __syncthreads();
            for (int t = warp; t < T; t += nw) {                    // the T searches side by side, a warp each
                const uint32_t a = s_rvraw[2 * (gi * T + t)] >> 5, b = s_rvraw[2 * (gi * T + t) + 1] >> 6;
                const double u = __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
                const double rv = __dmul_rn(u, (double)pot);
                // first j with cum[j] >= rv, cum = running sum of `closest`, in three warp-wide steps (every running sum
                // below is non-decreasing along the lanes, so "how many are below rv" is the index looked for):
                // the warp whose running total gets there, the chunk inside it, the element inside the chunk
                const unsigned long long run = rhccq_mb_warp_incl_scan(lane < nw ? s_wtot[lane] : 0ull);
                const int w = __popc(rhccq_ballot(lane < nw && (double)run < rv));
                unsigned long long base = rhccq_shfl(run, w > 0 ? w - 1 : 0);
                if (w == 0) base = 0ull;
                int j = ns;
                if (w < nw) {                                       // (warp-uniform; always, while pot > 0)
                    const unsigned long long ci = base + chunk_incl[w * RHCCQ_WARP_SIZE + lane];
                    const int lo = __popc(rhccq_ballot((double)ci < rv));
                    if (lo < RHCCQ_WARP_SIZE) {
                        unsigned long long run0 = rhccq_shfl(ci, lo > 0 ? lo - 1 : 0);
                        if (lo == 0) run0 = base;
                        const int chunk = w * RHCCQ_WARP_SIZE + lo;
                        const int j_lo = chunk * per < ns ? chunk * per : ns, j_hi = j_lo + per < ns ? j_lo + per : ns;
                        for (int j0 = j_lo; j0 < j_hi; j0 += RHCCQ_WARP_SIZE) {
                            const int jj = j0 + lane;
                            unsigned long long o = 0ull;
                            if (jj < j_hi) {
                                uint32_t cl = closest[jj];
                                if (pending) { const uint32_t d = (uint32_t)rhccq_d2(xs[jj], pend); cl = d < cl ? d : cl; }
                                o = cl;
                            }
                            const unsigned long long e = rhccq_mb_warp_incl_scan(o);
                            const unsigned reach = rhccq_ballot(jj < j_hi && !((double)(run0 + e) < rv));
                            if (reach) { j = j0 + __ffs((int)reach) - 1; break; }
                            run0 += rhccq_shfl(e, RHCCQ_WARP_SIZE - 1);
                        }
                    }
                }
                if (lane == 0) s_cand[t] = j < ns - 1 ? j : ns - 1;
            }
            __syncthreads();
            uint32_t xc[RHCCQ_MB_MAXT];
            S part[RHCCQ_MB_MAXT];                                  // this thread's share of every candidate's potential
#pragma unroll
            for (int t = 0; t < RHCCQ_MB_MAXT; ++t) { part[t] = 0; xc[t] = t < T ? xs[s_cand[t]] : 0u; }
            for (int j = c_lo; j < c_hi; ++j) {
                const uint32_t cj = xs[j];
                uint32_t o = closest[j];
                if (pending) { const uint32_t d = (uint32_t)rhccq_d2(cj, pend); o = d < o ? d : o; closest[j] = o; }
#pragma unroll
                for (int t = 0; t < RHCCQ_MB_MAXT; ++t) {
                    if (t >= T) break;                              // uniform: no predicated-off slots are issued
                    const uint32_t d = (uint32_t)rhccq_d2(cj, xc[t]);
                    part[t] += (S)(d < o ? d : o);
                }
            }
#pragma unroll
            for (int t = 0; t < RHCCQ_MB_MAXT; ++t) {
                if (t >= T) break;
                const S x = rhccq_mb_warp_sum(part[t]);
                if (lane == 0) s_sv[warp * RHCCQ_MB_MAXT + t] = x;
            }
            __syncthreads();
            int best = 0;
            S best_pot = 0;
#pragma unroll
            for (int t = 0; t < RHCCQ_MB_MAXT; ++t) {
                if (t >= T) break;
                const S x = rhccq_mb_warp_sum(lane < nw ? s_sv[lane * RHCCQ_MB_MAXT + t] : (S)0);
                if (t == 0 || x < best_pot) { best_pot = x; best = t; }
            }
            uint32_t cs = xc[0];
            chunk_sum = part[0];                                    // sum over the chunk of min(closest, distance to the new centre)
#pragma unroll
            for (int t = 1; t < RHCCQ_MB_MAXT; ++t) if (t == best) { cs = xc[t]; chunk_sum = part[t]; }
            pending = true;
            pend = cs;
            RHCCQ_PAR_FOR(q, 3) W.center[3 * c + q] = (double)((cs >> (16 - 8 * q)) & 255u);
            pot = best_pot;
        }
    };
    if ((unsigned long long)ns * 195075ull < 4294967296ull) seed_loop((uint32_t)0);
    else seed_loop((unsigned long long)0);
    __syncthreads();
    RHCCQ_MBP(1);                                                   // seeding
            RHCCQ_PAR_FOR(q, k) W.counts[q] = 0.0;
            if (threadIdx.x == 0) {                                 // cdf of choice(n, batch, p=ones/n)
                const double pi = __ddiv_rn(1.0, (double)n);
                double run = 0.0;
                for (int i = 0; i < n; ++i) { run = __dadd_rn(run, pi); W.cdf[i] = run; }
                s_val = run;
            }
            __syncthreads();
            {                                                       // cdf /= cdf[-1], once instead of in every search step
                const double last = s_val;
                RHCCQ_PAR_FOR(i, n) W.cdf[i] = __ddiv_rn(W.cdf[i], last);
            }
        } while (0);
        __syncthreads();
        if (threadIdx.x == 0) { hdr[0] = status; hdr[1] = n; hdr[2] = k; hdr[3] = batch; hdr[4] = 0; }
    }
    RHCCQ_MBP(2);                                                   // cumulative probabilities
    rhccq_cluster_sync();
    RHCCQ_MBP_ADD(12, hdr[1]); RHCCQ_MBP_ADD(13, hdr[2]);
    status = hdr[0]; n = hdr[1]; k = hdr[2]; batch = hdr[3];
    if (status < 0) {                                               // cluster-uniform
        if (rank == 0 && threadIdx.x == 0) n_clusters[p] = status;
        return;
    }
    // ---- mini-batch steps (_kmeans.py:2160-2215)
    const long long n_steps = (100LL * n) / batch;
    int n_since = 0, no_improvement = 0;                            // rank 0's bookkeeping
    bool have_ewa = false, have_min = false;
    double ewa = 0.0, ewa_min = 0.0;
    int* bidx = W.sub;                                              // the batch's rows (nb order); the subset is dead
    int* bidx_next = W.perm;                                        // ... of the step after, when they are drawn ahead
    double* cen = W.center;
    double* cen_new = W.center_new;
    // With more than one CTA, rank 0 leaves the centre update to the others: while they update, it adds up the batch
    // inertia, takes the stopping decision and draws the next batch's rows.  A step that may reassign centres (its
    // random draws come between the two batches', and it edits the centres after the update) keeps the serial order.
    const bool overlap = csize > 1;
    const int workers = overlap ? csize - 1 : 1, wrank = overlap ? rank - 1 : 0;      // of the centre update
    const bool worker = !overlap || rank > 0;
    const int pts_per_cta = (batch + csize - 1) / csize;
    const int k_per_cta = (k + workers - 1) / workers;
    bool have_next = false;                                         // rank 0: `bidx` already holds this step's rows
    auto draw_batch = [&](int* dst) {
        // uniform_samples = random_sample(batch): two raw outputs per double, converted in place
        rhccq_mt_fill_raw(mt, reinterpret_cast<uint32_t*>(s_own), 2 * batch);
        RHCCQ_PAR_FOR(i, batch) {
            const uint32_t a = reinterpret_cast<uint32_t*>(s_own)[2 * i] >> 5, b = reinterpret_cast<uint32_t*>(s_own)[2 * i + 1] >> 6;
            s_own[i] = __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
        }
        __syncthreads();
        RHCCQ_PAR_FOR(i, batch) {
            // searchsorted(cdf / last, u, side='right'): first g with cdf[g] > u. The cdf is that of n equal
            // weights, so u * n is at most a step or two off; the walk makes the answer exact wherever it starts
            const double u = s_own[i];
            int g = (int)(u * (double)n);
            g = g < 0 ? 0 : (g > n - 1 ? n - 1 : g);
            while (g > 0 && W.cdf[g - 1] > u) --g;
            while (g < n && W.cdf[g] <= u) ++g;
            dst[i] = g < n ? g : n - 1;
        }
        __syncthreads();
    };
    int e_lo, e_hi;
    rhccq_sk_edge_rows(k, e_lo, e_hi);
    for (long long step = 0; step < n_steps; ++step) {
        bool reassign = false;
        if (rank == 0) {
            if (!have_next) draw_batch(bidx);
            have_next = false;
            // _random_reassign (:2039-2054)
            n_since += batch;
            int zero = 0;
            RHCCQ_PAR_FOR(q, k) if (W.counts[q] == 0.0) zero = 1;
            zero = rhccq_block_or(zero, s_i);
            reassign = zero || n_since >= 10 * k;
            if (reassign) n_since = 0;
            if (threadIdx.x == 0) hdr[5] = reassign ? 1 : 0;
        }
        RHCCQ_MBP(3);                                               // the batch's rows (rank 0)
        rhccq_cluster_sync();
        RHCCQ_MBP(4);
        const bool serial = !overlap || hdr[5] != 0;                // (cluster-uniform)
        // labels and distances of this CTA's share of the batch: four threads per point, each a quarter of the
        // centres (interleaved).  First level in float32 on this CTA's own table of (-2 (c - 128), |c - 128|^2): the
        // score |c'|^2 - 2 x'.c' with x' = x - 128 differs from the float64 squared distance minus |x'|^2 by < 0.13
        // (inputs exact or rounded by < 3.1e-5, |x'| <= 128, three fused steps at |score| < 1.4e5), so a point whose
        // runner-up is more than 0.5 behind keeps the float32 winner; the others — ties of integer colours among
        // them — take the float64 loop, whose first minimum by (distance, index) is the result either way.
        {
            RHCCQ_MBP_RANK_T0();
            float* cf = W.cf + (size_t)rank * 4 * k;
            RHCCQ_PAR_FOR(q, k) {
                const double c0 = cen[3 * q] - 128.0, c1 = cen[3 * q + 1] - 128.0, c2 = cen[3 * q + 2] - 128.0;
                cf[4 * q] = (float)(-2.0 * c0); cf[4 * q + 1] = (float)(-2.0 * c1); cf[4 * q + 2] = (float)(-2.0 * c2);
                cf[4 * q + 3] = (float)(c0 * c0 + c1 * c1 + c2 * c2);
            }
            __syncthreads();
            const float4* c4 = reinterpret_cast<const float4*>(cf);
            const int p_lo = rank * pts_per_cta < batch ? rank * pts_per_cta : batch;
            const int p_hi = p_lo + pts_per_cta < batch ? p_lo + pts_per_cta : batch;
            const int tpp = RHCCQ_WARP_SIZE >= 4 ? 4 : 1;           // threads per point
            for (int i0 = p_lo; i0 < p_hi; i0 += (int)blockDim.x / tpp) {
                const int i = i0 + (int)threadIdx.x / tpp, part = (int)threadIdx.x % tpp;
                const bool have = i < p_hi;
                const uint32_t c = keys[W.nb[bidx[have ? i : p_lo]]];
                const double x0 = (double)rhccq_key_r(c), x1 = (double)rhccq_key_g(c), x2 = (double)rhccq_key_b(c);
                const float f0 = (float)(rhccq_key_r(c) - 128), f1 = (float)(rhccq_key_g(c) - 128), f2 = (float)(rhccq_key_b(c) - 128);
                float fb = 3.0e38f, fs = 3.0e38f;
                int fq = 0x7fffffff;
                for (int q = part; q < k; q += tpp) {
                    const float4 cq = c4[q];
                    const float d = fmaf(f0, cq.x, fmaf(f1, cq.y, fmaf(f2, cq.z, cq.w)));
                    fs = fminf(fs, fmaxf(d, fb));
                    const bool lt = d < fb;
                    fb = lt ? d : fb;
                    fq = lt ? q : fq;
                }
                __syncwarp();                                       // the four parts ran different trip counts
                for (int m = 1; m < tpp; m <<= 1) {
                    const float ob = rhccq_shfl_xor(fb, m), os = rhccq_shfl_xor(fs, m);
                    const int oq = rhccq_shfl_xor(fq, m);
                    fs = fminf(fminf(fs, os), fmaxf(fb, ob));       // runner-up of the union
                    if (ob < fb || (ob == fb && oq < fq)) { fb = ob; fq = oq; }
                }
                // A point whose float32 runner-up is within 0.5 of its best ("open": 0.6 % of them) is decided by
                // scikit-learn's own evaluation (_labels_inertia -> the chunked E step): |c|^2 - 2 x.c through dgemm in
                // float64, first minimum, the sample's place in its chunk of 256 selecting the dgemm kernel — taken
                // over the centres whose float32 score is within the margin of the best one (any other centre's
                // float64 score is above the best centre's by more than 0.5 - 2 * 0.13).  The whole warp serves one
                // open point at a time, a lane every 32nd centre.
                const bool open_pt = have && !(fs - fb > 0.5f);     // (the same in the four lanes of a point)
                int bq = fq;
                unsigned open_mask = rhccq_ballot(open_pt && part == 0);
                while (open_mask) {                                 // (warp-uniform)
                    const int src = __ffs((int)open_mask) - 1;
                    open_mask &= open_mask - 1u;
                    const float g0 = rhccq_shfl(f0, src), g1 = rhccq_shfl(f1, src), g2 = rhccq_shfl(f2, src);
                    const float lim = rhccq_shfl(fb, src) + 0.5f;
                    const rhccq_sk_pt pt = {rhccq_shfl(x0, src), rhccq_shfl(x1, src), rhccq_shfl(x2, src)};
                    const bool es = rhccq_sk_edge_sample(rhccq_shfl(i, src), batch);
                    double cd = 1.0e300;
                    int cq = 0x7fffffff;
                    for (int q = RHCCQ_LANE; q < k; q += RHCCQ_WARP_SIZE) {
                        const float4 c4q = c4[q];
                        if (fmaf(g0, c4q.x, fmaf(g1, c4q.y, fmaf(g2, c4q.z, c4q.w))) > lim) continue;
                        const double* cq3 = cen + 3 * q;
                        const double d = rhccq_sk_score(pt, cq3, rhccq_sk_norm3(cq3[0], cq3[1], cq3[2]), es && q >= e_lo && q < e_hi);
                        if (d < cd) { cd = d; cq = q; }
                        RHCCQ_MBP_ANY(17, 1);
                    }
                    for (int m = RHCCQ_WARP_SIZE >> 1; m > 0; m >>= 1) {
                        const double od = rhccq_shfl_xor(cd, m);
                        const int oq = rhccq_shfl_xor(cq, m);
                        if (od < cd || (od == cd && oq < cq)) { cd = od; cq = oq; }
                    }
                    if (RHCCQ_LANE / tpp == src / tpp) bq = cq;     // the lanes of that point
                    if (RHCCQ_LANE == src) RHCCQ_MBP_ANY(16, 1);
                }
                // the distance that enters the batch inertia is the direct one (_inertia_dense)
                if (have && part == 0) { lab_g[i] = bq; own_g[i] = rhccq_mb_dist(x0, x1, x2, cen + 3 * bq); }
            }
            __syncthreads();
            RHCCQ_MBP_RANK(18);
        }
        RHCCQ_MBP(5);                                               // labels of the batch
        rhccq_cluster_sync();
        RHCCQ_MBP(6);
        // centre update of this CTA's share of the clusters, members in batch order (_k_means_minibatch.pyx:68-118)
        if (worker) {
        RHCCQ_PAR_FOR(i, batch) { s_lab[i] = lab_g[i]; s_col[i] = keys[W.nb[bidx[i]]]; }
        __syncthreads();
        {
            const int q_lo = wrank * k_per_cta < k ? wrank * k_per_cta : k, q_hi = q_lo + k_per_cta < k ? q_lo + k_per_cta : k;
            const int kpc = q_hi > q_lo ? q_hi - q_lo : 0;
            // the batch grouped by cluster, batch order kept inside a group (the sums below are float64 additions in
            // that order): counts, exclusive scan, then a stable scatter 32 elements at a time (rank among the equal
            // labels of a tile by warp match).  Lives behind the batch's arrays in the shared buffer.
            int* g_start = reinterpret_cast<int*>(s_raw + RHCCQ_MB_BATCH * 16);                // [kpc + 1]
            int* g_cur = g_start + RHCCQ_MB_GROUP_CAP + 1;                                     // [kpc]
            int* g_order = g_cur + RHCCQ_MB_GROUP_CAP;                                         // [batch]
            const bool grouped = kpc <= RHCCQ_MB_GROUP_CAP;
            if (grouped) {
                RHCCQ_PAR_FOR(q, kpc + 1) g_start[q] = 0;
                __syncthreads();
                RHCCQ_PAR_FOR(i, batch) { const int L = s_lab[i] - q_lo; if (L >= 0 && L < kpc) atomicAdd(&g_start[L], 1); }
                __syncthreads();
                rhccq_block_excl_scan_array<int>(g_start, kpc + 1, s_i);
                RHCCQ_PAR_FOR(q, kpc) g_cur[q] = g_start[q];
                __syncthreads();
                if (threadIdx.x < RHCCQ_WARP_SIZE) {
                    for (int base = 0; base < batch; base += RHCCQ_WARP_SIZE) {
                        const int i = base + RHCCQ_LANE;
                        const int L = i < batch ? s_lab[i] - q_lo : -1;
                        const bool mine = L >= 0 && L < kpc;
#ifdef RHCCQ_HOST_EMU
                        if (mine) g_order[g_cur[L]++] = i;
#else
                        const unsigned m = __match_any_sync(0xffffffffu, mine ? L : 0x40000000 + RHCCQ_LANE);
                        const int rk = __popc(m & rhccq_lanemask_lt());
                        int at = 0;
                        if (mine) at = g_cur[L];
                        __syncwarp();
                        if (mine) {
                            g_order[at + rk] = i;
                            if (rk == __popc(m) - 1) g_cur[L] = at + __popc(m);
                        }
                        __syncwarp();
#endif
                    }
                }
                __syncthreads();
            }
            for (int q = q_lo + (int)threadIdx.x; q < q_hi; q += (int)blockDim.x) {
                int members = 0;
                if (grouped) members = g_start[q - q_lo + 1] - g_start[q - q_lo];
                else for (int i = 0; i < batch; ++i) members += s_lab[i] == q;
                if (members > 0) {
                    const double w = W.counts[q];
                    double c0 = __dmul_rn(cen[3 * q], w), c1 = __dmul_rn(cen[3 * q + 1], w), c2 = __dmul_rn(cen[3 * q + 2], w);
                    if (grouped) {
                        for (int t = g_start[q - q_lo]; t < g_start[q - q_lo + 1]; ++t) {
                            const uint32_t c = s_col[g_order[t]];
                            c0 = __dadd_rn(c0, (double)rhccq_key_r(c)); c1 = __dadd_rn(c1, (double)rhccq_key_g(c)); c2 = __dadd_rn(c2, (double)rhccq_key_b(c));
                        }
                    } else {
                        for (int i = 0; i < batch; ++i) if (s_lab[i] == q) {
                            const uint32_t c = s_col[i];
                            c0 = __dadd_rn(c0, (double)rhccq_key_r(c)); c1 = __dadd_rn(c1, (double)rhccq_key_g(c)); c2 = __dadd_rn(c2, (double)rhccq_key_b(c));
                        }
                    }
                    const double wn = __dadd_rn(w, (double)members);
                    const double alpha = __ddiv_rn(1.0, wn);
                    W.counts[q] = wn;
                    cen_new[3 * q] = __dmul_rn(c0, alpha); cen_new[3 * q + 1] = __dmul_rn(c1, alpha); cen_new[3 * q + 2] = __dmul_rn(c2, alpha);
                } else {
                    cen_new[3 * q] = cen[3 * q]; cen_new[3 * q + 1] = cen[3 * q + 1]; cen_new[3 * q + 2] = cen[3 * q + 2];
                }
            }
        }
        }
        // inertia of the batch, reassignment of starved centres, stopping rule: rank 0, after the update — or beside it
        // when the step cannot reassign (nothing it does then touches what the update reads or writes)
        auto after_update = [&]() {
            RHCCQ_PAR_FOR(i, batch) s_own[i] = own_g[i];
            __syncthreads();
            if (threadIdx.x == 0) {                                 // inertia in batch order
                double in = 0.0;
                for (int i = 0; i < batch; ++i) in = __dadd_rn(in, s_own[i]);
                s_val = in;
            }
            __syncthreads();
            const double inertia = s_val;
            if (reassign) {                                             // :1652-1682
                double mx = 0.0;
                RHCCQ_PAR_FOR(q, k) if (W.counts[q] > mx) mx = W.counts[q];
                mx = rhccq_block_max<double>(mx, s_d);
                const double lim = __dmul_rn(0.01, mx);
                int cnt = 0;
                RHCCQ_PAR_FOR(q, k) { const int f = W.counts[q] < lim ? 1 : 0; W.flag[q] = f; cnt += f; }
                cnt = rhccq_block_sum<int>(cnt, s_i);
                if ((double)cnt > 0.5 * (double)batch) {
                    // keep all but the int(0.5 * batch) smallest counts (stable order: ties by index).  The counts are
                    // small integers while many clusters are still empty, which is when this branch is taken step after
                    // step: then the cut falls inside the run of one small count value v, and the kept set is "count < v,
                    // plus the first few clusters with count == v by index" — one scan instead of a sort of all k keys.
                    const int first_kept = (int)(0.5 * (double)batch);
                    int* rank_eq = reinterpret_cast<int*>(W.skey);
                    bool selected = false;
                    int below = 0;                                      // clusters with count < v
                    for (int v = 0; v < 16 && !selected; ++v) {
                        int eq = 0;
                        RHCCQ_PAR_FOR(q, k) eq += W.counts[q] == (double)v ? 1 : 0;
                        eq = rhccq_block_sum<int>(eq, s_i);
                        if (below + eq >= first_kept) {
                            // the cut is inside the clusters with count == v: the first (first_kept - below) of them stay flagged
                            RHCCQ_PAR_FOR(q, k) rank_eq[q] = W.counts[q] == (double)v ? 1 : 0;
                            __syncthreads();
                            rhccq_block_excl_scan_array<int>(rank_eq, k, s_i);
                            const int take = first_kept - below;
                            RHCCQ_PAR_FOR(q, k) {
                                const double c = W.counts[q];
                                if (c > (double)v || (c == (double)v && rank_eq[q] >= take)) W.flag[q] = 0;
                            }
                            selected = true;
                        }
                        below += eq;
                    }
                    if (!selected) {
                        int k2 = 1;
                        while (k2 < k) k2 <<= 1;
                        for (int q = threadIdx.x; q < k2; q += blockDim.x)
                            W.skey[q] = q < k ? (((unsigned long long)(long long)W.counts[q] << 24) | (unsigned)q) : ~0ull;
                        __syncthreads();
                        rhccq_block_bitonic_sort<unsigned long long>(W.skey, k2);
                        RHCCQ_PAR_FOR(r, k) if (r >= first_kept) W.flag[(int)(W.skey[r] & 0xffffffu)] = 0;
                    }
                    __syncthreads();
                    cnt = 0;
                    RHCCQ_PAR_FOR(q, k) cnt += W.flag[q];
                    cnt = rhccq_block_sum<int>(cnt, s_i);
                }
                __syncthreads();
                int* s_perm = reinterpret_cast<int*>(s_raw);            // (the batch's values are consumed: the buffer is free)
                if (cnt > 0) {
                    // choice(batch, replace=False, size=cnt) == permutation(batch)[:cnt]: Fisher-Yates from the top, one
                    // thread, in shared memory; the raw outputs it consumes are produced a generator block at a time
                    // by the whole CTA
                    RHCCQ_PAR_FOR(i, batch) s_perm[i] = i;
                    if (threadIdx.x == 0) s_prog = batch - 1;
                    while (true) {
                        const int avail = rhccq_mt_peek_block(mt, s_rbuf);
                        if (threadIdx.x == 0) {
                            int i = s_prog, r = 0;
                            const uint32_t* __restrict__ rb = s_rbuf;       // (disjoint from the permutation: lets the raw
                            int* __restrict__ pm = s_perm;                  //  values be fetched eight at a time, ahead of the swaps)
                            while (i >= 1 && r < avail) {
                                uint32_t raw8[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) raw8[j] = r + j < avail ? rb[r + j] : 0u;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    if (i >= 1 && r < avail) {
                                        uint32_t mask = (uint32_t)i;
                                        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                                        const uint32_t v = raw8[j] & mask;
                                        ++r;
                                        if (v <= (uint32_t)i) { const int t = pm[i]; pm[i] = pm[v]; pm[v] = t; --i; }
                                    }
                                }
                            }
                            *mt.pos += r;
                            s_prog = i;
                        }
                        __syncthreads();
                        if (s_prog < 1) break;
                    }
                    // rank of every flagged centre among the flagged ones, ascending
                    int* rank = reinterpret_cast<int*>(W.skey);
                    RHCCQ_PAR_FOR(q, k) rank[q] = W.flag[q];
                    __syncthreads();
                    rhccq_block_excl_scan_array<int>(rank, k, s_i);
                    RHCCQ_PAR_FOR(q, k) if (W.flag[q]) {
                        const uint32_t c = keys[W.nb[bidx[s_perm[rank[q]]]]];
                        cen_new[3 * q] = (double)rhccq_key_r(c); cen_new[3 * q + 1] = (double)rhccq_key_g(c); cen_new[3 * q + 2] = (double)rhccq_key_b(c);
                    }
                }
                double mn = 1.0e300;
                RHCCQ_PAR_FOR(q, k) if (!W.flag[q] && W.counts[q] < mn) mn = W.counts[q];
                mn = rhccq_block_min<double>(mn, s_d);
                if (mn < 1.0e300) RHCCQ_PAR_FOR(q, k) if (W.flag[q]) W.counts[q] = mn;
                __syncthreads();
            }
            int done = 0;
            do {
                // _mini_batch_convergence (:1974-2037), tol == 0
                const double bi = __ddiv_rn(inertia, (double)batch);
                if (step == 0) break;
                if (!have_ewa) { ewa = bi; have_ewa = true; }
                else {
                    double alpha = __ddiv_rn(__dmul_rn((double)batch, 2.0), (double)(n + 1));
                    if (alpha > 1.0) alpha = 1.0;
                    ewa = __dadd_rn(__dmul_rn(ewa, __dsub_rn(1.0, alpha)), __dmul_rn(bi, alpha));
                }
                if (!have_min || ewa < ewa_min) { no_improvement = 0; ewa_min = ewa; have_min = true; }
                else ++no_improvement;
                if (no_improvement >= 10) done = 1;
            } while (0);
            __syncthreads();
            if (threadIdx.x == 0) hdr[4] = done;
        };
        if (rank == 0 && !serial) { after_update(); draw_batch(bidx_next); have_next = true; }
        RHCCQ_MBP(7);                                               // centre update
        rhccq_cluster_sync();
        RHCCQ_MBP(8);
        if (serial) {
            if (rank == 0) after_update();
            RHCCQ_MBP(9);                                           // inertia, reassignment, convergence (rank 0)
            rhccq_cluster_sync();
            RHCCQ_MBP(10);
        }
        RHCCQ_MBP_ADD(11, 1);
        if (reassign) RHCCQ_MBP_ADD(14, 1);
        { double* t = cen; cen = cen_new; cen_new = t; }
        if (!serial) { int* t = bidx; bidx = bidx_next; bidx_next = t; }     // the rows drawn ahead are the next step's
        if (hdr[4]) break;                                          // cluster-uniform
    }
    RHCCQ_MBP_END();
    if (rank == 0) {
        RHCCQ_PAR_FOR(q, 3 * k) centers_out[q] = cen[q];
        if (threadIdx.x == 0) { n_clusters[p] = k; centers_out[3 * k] = (double)n; }   // (k <= kmax - 1: the slot exists)
    }
}

#if defined(RHCCQ_MB_PROFILE) && !defined(RHCCQ_HOST_EMU)
extern "C" int rhccq_mb_prof_read(unsigned long long* host_out, int reset) {
    if (reset) { unsigned long long z[40] = {0}; return (int)cudaMemcpyToSymbol(rhccq_mb_prof, z, sizeof z); }
    return (int)cudaMemcpyFromSymbol(host_out, rhccq_mb_prof, 40 * sizeof(unsigned long long));
}
#endif

// status -4 (set by rhccq_k_palette_dbscan) selects the palettes of this branch; one cluster of CTAs per palette
__global__ void __launch_bounds__(RHCCQ_MB_THREADS)
rhccq_k_palette_minibatch(rhccq_palette_batch B, const double* __restrict__ quality, int* __restrict__ n_clusters,
                          int max_rows, unsigned char* gws, size_t gws_stride, double* __restrict__ centers,
                          size_t centers_stride, int* __restrict__ todo, int* __restrict__ claim, int* __restrict__ labels) {
    int rank, csize, cid;
    rhccq_cluster_rank(&rank, &csize, &cid);
    while (true) {
        if (rank == 0 && threadIdx.x == 0) {
            // next palette that needs the branch: a grid-wide cursor keeps one workspace slice per cluster enough
            int p;
            while ((p = atomicAdd(todo, 1)) < B.n_problems && n_clusters[p] != -4) {}
            claim[cid] = p;
        }
        rhccq_cluster_sync();
        const int p = claim[cid];
        rhccq_cluster_sync();
        if (p >= B.n_problems) return;
        rhccq_minibatch_problem(B, p, quality, n_clusters, max_rows, gws + (size_t)cid * gws_stride,
                                centers + (size_t)p * centers_stride, labels, rank, csize);
        rhccq_cluster_sync();
    }
}

// labels of all non-black rows: scikit-learn's final _labels_inertia over the whole palette (chunked E step as in
// the batches: |c|^2 - 2 x.c, first minimum, dgemm kernel by the row's place in its chunk)
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_minibatch_assign(rhccq_palette_batch B, const int* __restrict__ n_clusters_before, const int* __restrict__ n_clusters,
                         const double* __restrict__ centers, size_t centers_stride, int* __restrict__ labels, int chunks) {
    for (int w = blockIdx.x; w < B.n_problems * chunks; w += gridDim.x) {
        const int p = w / chunks, ch = w % chunks;
        if (n_clusters_before[p] != -4 || n_clusters[p] <= 0) continue;
        const int n = B.pal_cnt[p], k = n_clusters[p];
        const uint32_t* keys = B.pal_keys + B.pal_off[p];
        int* lab = labels + B.pal_off[p];
        const double* C = centers + (size_t)p * centers_stride;
        const int n_rows = (int)C[3 * k];                           // non-black rows
        int e_lo, e_hi;
        rhccq_sk_edge_rows(k, e_lo, e_hi);
        const int per = (n + chunks - 1) / chunks;
        const int lo = ch * per, hi = lo + per < n ? lo + per : n;
        for (int i = lo + (int)threadIdx.x; i < hi; i += (int)blockDim.x) {
            const uint32_t c = keys[i];
            if (c == 0u) { lab[i] = -2; continue; }
            const rhccq_sk_pt pt = {(double)rhccq_key_r(c), (double)rhccq_key_g(c), (double)rhccq_key_b(c)};
            const bool es = rhccq_sk_edge_sample(lab[i], n_rows);   // (the slot holds the row's position, see above)
            double bd = 1.0e300;
            int bi = 0;
            for (int q = 0; q < k; ++q) {
                const double* cq = C + 3 * q;
                const double d = rhccq_sk_score(pt, cq, rhccq_sk_norm3(cq[0], cq[1], cq[2]), es && q >= e_lo && q < e_hi);
                if (d < bd) { bd = d; bi = q; }
            }
            lab[i] = bi;
        }
    }
}

int rhccq_launch_palette_minibatch(const rhccq_palette_batch& B, const double* quality, int* labels, int* n_clusters,
                                   int max_rows, rhccq_launch_ws ws, void* stream) {
    if (B.n_problems <= 0) return 0;
    const size_t slice = rhccq_palette_minibatch_ws_bytes(max_rows);
    const size_t kmax = (size_t)max_rows / 10 + 2;
    const size_t cstride = 3 * kmax;
    // workspace: [todo cursor, claim slots, copy of the status vector][centres of every palette][one slice per cluster]
    const size_t head = rhccq_carve_bytes((size_t)B.n_problems + 4 + 256, 4);
    const size_t cbytes = rhccq_carve_bytes((size_t)B.n_problems * cstride, 8);
    if (!ws.ws || ws.ws_bytes < head + cbytes + slice) {
        rhccq_set_error("rhccq_palette_minibatch: workspace of %zu bytes is smaller than %zu", ws.ws_bytes, head + cbytes + slice);
        return -1;
    }
    // one cluster of RHCCQ_MB_CLUSTER CTAs per palette in flight; every cluster owns one workspace slice
    int workers = (int)((ws.ws_bytes - head - cbytes) / slice);
    if (workers > B.n_problems) workers = B.n_problems;
    const int cap = rhccq_sm_count() / RHCCQ_MB_CLUSTER;
    if (workers > cap) workers = cap;
    if (workers < 1) workers = 1;
    int* todo = (int*)ws.ws;
    int* claim = todo + 4;                                          // one slot per cluster, then the status copy
    int* before = claim + workers;
    double* centers = (double*)(ws.ws + head);
    unsigned char* slices = ws.ws + head + cbytes;
#ifdef RHCCQ_HOST_EMU
    memset(todo, 0, 16);
    memcpy(before, n_clusters, (size_t)B.n_problems * 4);
    RHCCQ_LAUNCH(rhccq_k_palette_minibatch, workers, RHCCQ_MB_THREADS, 0, (cudaStream_t)stream, B, quality, n_clusters, max_rows,
                 slices, slice, centers, cstride, todo, claim, labels);
#else
    cudaMemsetAsync(todo, 0, 16, (cudaStream_t)stream);
    cudaMemcpyAsync(before, n_clusters, (size_t)B.n_problems * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(workers * RHCCQ_MB_CLUSTER), 1, 1);
        cfg.blockDim = dim3(RHCCQ_MB_THREADS, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = RHCCQ_MB_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, rhccq_k_palette_minibatch, B, quality, n_clusters, max_rows, slices, slice,
                                           centers, cstride, todo, claim, labels);
        if (e != cudaSuccess) { rhccq_set_error("rhccq_palette_minibatch: cluster launch failed: %s", cudaGetErrorString(e)); return -1; }
    }
#endif
    const int chunks = 64;
    int g2 = B.n_problems * chunks;
    const int cap2 = rhccq_sm_count() * 8;
    if (g2 > cap2) g2 = cap2;
    RHCCQ_LAUNCH(rhccq_k_minibatch_assign, g2, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, B, before, n_clusters, centers, cstride,
                 labels, chunks);
    return 0;
}
