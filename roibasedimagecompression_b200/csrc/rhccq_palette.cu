// Palette clustering kernels: one CTA per palette ("problem").
//
//   rhccq_k_palette_dbscan   connected components of the eps-graph over the
//                            non-black palette rows == DBSCAN(min_samples=1)
//                            as the reference calls it
//                            (/root/reference/encoder/compression/clustering.py:185-235;
//                            sklearn/cluster/_dbscan.py:397-470 for the semantics)
//   rhccq_k_palette_split    small clusters -> one entry, large clusters ->
//                            recursive K-Means split (clustering.py:273-355,
//                            :720-775), K-Means in the exact arithmetic of
//                            oracle/kmeans_restated.py
//   rhccq_k_palette_finish   new colour = truncated mean of the members
//                            (clustering.py:305,347)
//
// A batch of problems is described by offsets into flat arrays: problem p owns
// rows pal_off[p] .. pal_off[p] + pal_cnt[p].
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"

#define RHCCQ_EMPTY_KEY 0xFFFFFFFFu

// ---------------------------------------------------------------- eps-graph components
//
// Cells of side s = 1 + floor(sqrt(thr/3)) are cliques (3 (s-1)^2 <= thr), so
// the union-find runs over cells; a pair of cells is joined by the first pair
// of points found within eps.  Cells further apart than `reach` cells in any
// channel cannot be joined directly.
struct rhccq_cellgrid {
    int s, G, reach, direct, hcap, hshift;
};

__host__ __device__ static inline void rhccq_cellgrid_init(rhccq_cellgrid& g, int thr, int n_rows) {
    int s = 1;
    while ((long long)3 * s * s <= (long long)thr) ++s;            // s - 1 = floor(sqrt(thr / 3))
    g.s = s;
    g.G = (255 / s) + 1;
    int reach = 0;                                                 // largest D with ((D-1) s + 1)^2 <= thr
    while (true) {
        long long gap = (long long)reach * s + 1;                  // candidate D = reach + 1
        if (gap * gap <= (long long)thr) ++reach; else break;
    }
    g.reach = reach;
    long long cells = (long long)g.G * g.G * g.G;
    g.direct = cells <= 4096;
    if (g.direct) {
        g.hcap = (int)cells;
        g.hshift = 0;
    } else {
        int want = n_rows + n_rows / 2 + 1, cap = 64, sh = 6;
        while (cap < want) { cap <<= 1; ++sh; }
        g.hcap = cap;
        g.hshift = 32 - sh;
    }
}

size_t rhccq_palette_dbscan_ws_bytes(int max_rows, int max_slots) {
    return rhccq_carve_bytes(max_rows, 4) * 3 + rhccq_carve_bytes(max_slots, 4) * 4;
}

extern "C" int rhccq_palette_dbscan_slots(int thr, int n_rows) {
    rhccq_cellgrid g;
    rhccq_cellgrid_init(g, thr, n_rows);
    return g.hcap;
}

__device__ __forceinline__ int rhccq_uf_find(volatile int* parent, int x) {
    while (true) {
        int p = parent[x];
        if (p == x) return x;
        int gp = parent[p];
        if (gp != p) parent[x] = gp;                               // path halving; racing writers only shorten paths
        x = p;
    }
}

__device__ __forceinline__ void rhccq_uf_union(int* parent, int a, int b) {
    while (true) {
        a = rhccq_uf_find(parent, a);
        b = rhccq_uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }                    // link the larger root under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

// float64 evaluation of a pair at d2 == eps^2, as the KD-tree leaf does it
// (sklearn/neighbors/_binary_tree.pxi.tp:1952-1957 on X = palette / 255.0).
__device__ __forceinline__ bool rhccq_tie_accept(uint32_t a, uint32_t b, double eps) {
    const double r = __ddiv_rn(eps, 255.0);
    const double r2 = __dmul_rn(r, r);
    double s = 0.0;
#pragma unroll
    for (int sh = 16; sh >= 0; sh -= 8) {
        const double x = __ddiv_rn((double)((a >> sh) & 255u), 255.0);
        const double y = __ddiv_rn((double)((b >> sh) & 255u), 255.0);
        const double d = __dsub_rn(x, y);
        const double q = __dmul_rn(d, d);
        s = (sh == 16) ? q : __dadd_rn(s, q);
    }
    return s <= r2;
}

__device__ __forceinline__ int rhccq_slot_lookup(const rhccq_cellgrid& g, const uint32_t* hkey,
                                                 const int* head, uint32_t cell) {
    if (g.direct) return head[cell] >= 0 ? (int)cell : -1;
    uint32_t h = (cell * 2654435761u) >> g.hshift;
    while (true) {
        uint32_t k = hkey[h];
        if (k == cell) return (int)h;
        if (k == RHCCQ_EMPTY_KEY) return -1;
        h = (h + 1) & (uint32_t)(g.hcap - 1);
    }
}

__device__ void rhccq_palette_dbscan_problem(const rhccq_palette_batch& B, int p, int* __restrict__ labels,
                                             int* __restrict__ n_clusters, int max_rows, int max_slots,
                                             unsigned char* wsbase) {
    __shared__ int s_scratch[RHCCQ_MAX_WARPS + 2];
    const int n = B.pal_cnt[p];
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    int* lab = labels + B.pal_off[p];
    const int thr = B.thr[p];
    const int tie = B.tie[p];
    const double eps = B.eps[p];
    if (n < 0) {                                                   // upstream error
        if (threadIdx.x == 0) n_clusters[p] = -2;
        return;
    }
    if (n > max_rows) {                                            // caller sized the workspace too small
        if (threadIdx.x == 0) n_clusters[p] = -1;
        return;
    }
    {
        int nb = 0;
        RHCCQ_PAR_FOR(i, n) if (keys[i] != 0u) ++nb;
        nb = rhccq_block_sum<int>(nb, s_scratch);
        if (nb >= 10000) {                                         // clustering.py:207-218: MiniBatchKMeans branch
            if (threadIdx.x == 0) n_clusters[p] = -4;
            return;
        }
    }
    rhccq_cellgrid g;
    rhccq_cellgrid_init(g, thr, n);
    if (g.hcap > max_slots) {
        if (threadIdx.x == 0) n_clusters[p] = -1;
        return;
    }
    rhccq_carver cv(wsbase);
    uint32_t* col = cv.take<uint32_t>(max_rows);
    int* next = cv.take<int>(max_rows);                            // cell lists; reused as the label scan
    int* slot_of = cv.take<int>(max_rows);
    uint32_t* hkey = cv.take<uint32_t>(max_slots);
    int* head = cv.take<int>(max_slots);
    int* parent = cv.take<int>(max_slots);
    int* minidx = cv.take<int>(max_slots);

    RHCCQ_PAR_FOR(s, g.hcap) { hkey[s] = RHCCQ_EMPTY_KEY; head[s] = -1; parent[s] = s; minidx[s] = 0x7fffffff; }
    RHCCQ_PAR_FOR(i, n) col[i] = keys[i];
    __syncthreads();

    // cell lists
    RHCCQ_PAR_FOR(i, n) {
        const uint32_t c = col[i];
        if (c == 0u) { slot_of[i] = -1; next[i] = -1; continue; }  // black rows take no part (clustering.py:185-192)
        const uint32_t cell = ((uint32_t)(rhccq_key_r(c) / g.s) * g.G + (uint32_t)(rhccq_key_g(c) / g.s)) * g.G
                              + (uint32_t)(rhccq_key_b(c) / g.s);
        int slot;
        if (g.direct) {
            slot = (int)cell;
        } else {
            uint32_t h = (cell * 2654435761u) >> g.hshift;
            while (true) {
                uint32_t k = atomicCAS(&hkey[h], RHCCQ_EMPTY_KEY, cell);
                if (k == RHCCQ_EMPTY_KEY || k == cell) break;
                h = (h + 1) & (uint32_t)(g.hcap - 1);
            }
            slot = (int)h;
        }
        slot_of[i] = slot;
        next[i] = atomicExch(&head[slot], i);
    }
    __syncthreads();

    // join cells
    const int side = 2 * g.reach + 1;
    const int half = (side * side * side - 1) / 2;                 // offsets after the centre in raster order
    const long long work = (long long)n * half;
    for (long long w = threadIdx.x; w < work; w += blockDim.x) {
        const int i = (int)(w / half);
        const int a = slot_of[i];
        if (a < 0) continue;
        const int L = half + 1 + (int)(w % half);
        const int dr = L / (side * side) - g.reach;
        const int dg = (L / side) % side - g.reach;
        const int db = L % side - g.reach;
        const uint32_t c = col[i];
        const int cr = rhccq_key_r(c) / g.s + dr, cg = rhccq_key_g(c) / g.s + dg, cb = rhccq_key_b(c) / g.s + db;
        if (cr < 0 || cg < 0 || cb < 0 || cr >= g.G || cg >= g.G || cb >= g.G) continue;
        const int b = rhccq_slot_lookup(g, hkey, head, ((uint32_t)cr * g.G + (uint32_t)cg) * g.G + (uint32_t)cb);
        if (b < 0) continue;
        if (rhccq_uf_find(parent, a) == rhccq_uf_find(parent, b)) continue;
        for (int j = ((volatile int*)head)[b]; j >= 0; j = next[j]) {
            const int d2 = rhccq_d2(c, col[j]);
            bool hit = tie ? (d2 < thr) : (d2 <= thr);
            if (!hit && tie && d2 == thr) hit = rhccq_tie_accept(c, col[j], eps);
            if (hit) { rhccq_uf_union(parent, a, b); break; }
        }
    }
    __syncthreads();

    // canonical numbering: clusters in order of their lowest row (sklearn's DFS seeds ascend)
    RHCCQ_PAR_FOR(i, n) if (slot_of[i] >= 0) atomicMin(&minidx[rhccq_uf_find(parent, slot_of[i])], i);
    __syncthreads();
    RHCCQ_PAR_FOR(i, n) next[i] = 0;
    __syncthreads();
    RHCCQ_PAR_FOR(s, g.hcap) if (head[s] >= 0 && parent[s] == s) next[minidx[s]] = 1;
    __syncthreads();
    const int total = rhccq_block_excl_scan_array<int>(next, n, s_scratch);
    RHCCQ_PAR_FOR(i, n) lab[i] = slot_of[i] >= 0 ? next[minidx[rhccq_uf_find(parent, slot_of[i])]] : -2;
    if (threadIdx.x == 0) n_clusters[p] = total;
}

__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_palette_dbscan(rhccq_palette_batch B, int* __restrict__ labels, int* __restrict__ n_clusters,
                       int max_rows, int max_slots, unsigned char* gws, size_t gws_stride) {
    RHCCQ_DYN_SMEM(dyn);
    unsigned char* wsbase = gws ? gws + (size_t)blockIdx.x * gws_stride : dyn;
    for (int p = blockIdx.x; p < B.n_problems; p += gridDim.x) {
        rhccq_palette_dbscan_problem(B, p, labels, n_clusters, max_rows, max_slots, wsbase);
        __syncthreads();
    }
}

// Grid and workspace choice shared by the palette launchers: the per-CTA working
// set goes to shared memory when it fits, otherwise to a slice of the caller's
// global workspace (then at most ws_bytes / need CTAs walk the problems).
int rhccq_pick_grid(const void* kernel, size_t need, int n_problems, rhccq_launch_ws ws,
                           size_t* smem, unsigned char** gws, const char* what) {
    if (need <= RHCCQ_SMEM_BUDGET) {
        if (rhccq_smem_optin(kernel, need) != 0) return -1;
        *smem = need;
        *gws = nullptr;
        return n_problems;                                         // the hardware scheduler balances uneven problems
    }
    const size_t slices = ws.ws ? ws.ws_bytes / need : 0;
    if (slices == 0) {
        rhccq_set_error("%s: working set of %zu bytes per problem exceeds shared memory and the workspace "
                        "(%zu bytes) holds no slice", what, need, ws.ws_bytes);
        return -1;
    }
    *smem = 0;
    *gws = ws.ws;
    int grid = n_problems;
    if ((size_t)grid > slices) grid = (int)slices;
    const int cap = rhccq_sm_count() * 4;
    return grid < cap ? grid : cap;
}

int rhccq_launch_palette_dbscan(const rhccq_palette_batch& B, int* labels, int* n_clusters, int max_rows,
                                int max_slots, rhccq_launch_ws ws, void* stream) {
    if (B.n_problems <= 0) return 0;
    const size_t need = rhccq_palette_dbscan_ws_bytes(max_rows, max_slots);
    size_t smem; unsigned char* gws;
    const int grid = rhccq_pick_grid((const void*)rhccq_k_palette_dbscan, need, B.n_problems, ws, &smem, &gws,
                                     "rhccq_palette_dbscan");
    if (grid < 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_palette_dbscan, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                 B, labels, n_clusters, max_rows, max_slots, gws, need);
    return 0;
}

// ---------------------------------------------------------------- K-Means (exact arithmetic)
__device__ __forceinline__ int rhccq_kmeans_local_trials(int k) {   // 2 + int(log(k)), sklearn/_kmeans.py:226
    const int e[] = {3, 8, 21, 55, 149, 404, 1097, 2981, 8104, 22027, 59875, 162755, 442414, 1202605};
    int t = 2;
    for (int i = 0; i < 14; ++i) if (k >= e[i]) ++t;
    return t;
}

struct rhccq_kmeans_ws {
    uint32_t* x;          // [n] member colours in member order
    int* closest;         // [n]
    int* label;           // [n]
    int* label_old;       // [n]
    long long* cum;       // [n]
    double* center;       // [3k]
    double* center_new;   // [3k]
    double* term;         // [k]
    int* sums;            // [3k]
    int* cnt;             // [k]
};

__device__ __forceinline__ double rhccq_dist3(double x0, double x1, double x2, const double* c) {
    const double d0 = __dsub_rn(x0, c[0]), d1 = __dsub_rn(x1, c[1]), d2 = __dsub_rn(x2, c[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}

// E step: label = first minimum over centres of ((d0^2 + d1^2) + d2^2) in IEEE double.
__device__ __forceinline__ void rhccq_kmeans_estep(const rhccq_kmeans_ws& W, int n, int k) {
    RHCCQ_PAR_FOR(j, n) {
        const uint32_t c = W.x[j];
        const double x0 = (double)rhccq_key_r(c), x1 = (double)rhccq_key_g(c), x2 = (double)rhccq_key_b(c);
        double best = rhccq_dist3(x0, x1, x2, W.center);
        int bi = 0;
        for (int q = 1; q < k; ++q) {
            const double d = rhccq_dist3(x0, x1, x2, W.center + 3 * q);
            if (d < best) { best = d; bi = q; }
        }
        W.label[j] = bi;
    }
    __syncthreads();
}

// Labels of KMeans(k, random_state=42, n_init='auto').fit_predict on the n
// colours W.x[0..n), as restated in oracle/kmeans_restated.py (which cites the
// scikit-learn lines).  On return W.label holds the labels and W.cnt the
// cluster sizes.  Block-uniform control flow; every thread must call.
__device__ void rhccq_kmeans(const rhccq_kmeans_ws& W, int n, int k, const double* __restrict__ rng,
                             int* s_i, long long* s_ll, double* s_d) {
    // ---- k-means++ seeding (kmeans_restated.kmeans_pp_seeds)
    const int T = rhccq_kmeans_local_trials(k);
    int first = (int)__dmul_rn(rng[0], (double)n);
    if (first > n - 1) first = n - 1;
    long long part = 0, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
    {
        const uint32_t cf = W.x[first];
        RHCCQ_PAR_FOR(j, n) {
            const uint32_t c = W.x[j];
            const int d = rhccq_d2(c, cf);
            W.closest[j] = d;
            part += d;
            const long long r = rhccq_key_r(c), g = rhccq_key_g(c), b = rhccq_key_b(c);
            p1[0] += r; p1[1] += g; p1[2] += b;
            p2[0] += r * r; p2[1] += g * g; p2[2] += b * b;
        }
    }
    long long pot = rhccq_block_sum<long long>(part, s_ll);
    long long S1[3], S2[3];
    for (int d = 0; d < 3; ++d) { S1[d] = rhccq_block_sum<long long>(p1[d], s_ll); S2[d] = rhccq_block_sum<long long>(p2[d], s_ll); }
    RHCCQ_PAR_FOR(q, 3) W.center[q] = (double)((W.x[first] >> (16 - 8 * q)) & 255u);
    int ri = 1;
    for (int c = 1; c < k; ++c) {
        // inclusive cumulative sum of closest (exact integers)
        RHCCQ_PAR_FOR(j, n) W.cum[j] = W.closest[j];
        __syncthreads();
        rhccq_block_excl_scan_array<long long>(W.cum, n, s_ll);
        // candidates: searchsorted(cum, r * pot, side='left'), clipped
        RHCCQ_PAR_FOR(t, T) {
            const double rv = __dmul_rn(rng[ri + t], (double)pot);
            int lo = 0, hi = n;                                     // first j with incl[j] >= rv
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((double)(W.cum[mid] + W.closest[mid]) < rv) lo = mid + 1; else hi = mid;
            }
            s_i[t] = lo < n - 1 ? lo : n - 1;
        }
        __syncthreads();
        ri += T;
        int best = 0;
        long long best_pot = 0;
        for (int t = 0; t < T; ++t) {
            const uint32_t cc = W.x[s_i[t]];
            long long ps = 0;
            RHCCQ_PAR_FOR(j, n) {
                const int d = rhccq_d2(W.x[j], cc);
                const int o = W.closest[j];
                ps += d < o ? d : o;
            }
            const long long tot = rhccq_block_sum<long long>(ps, s_ll);
            if (t == 0 || tot < best_pot) { best_pot = tot; best = t; }
        }
        const int seed = s_i[best];
        const uint32_t cs = W.x[seed];
        __syncthreads();                                            // everyone has read s_i
        RHCCQ_PAR_FOR(j, n) {
            const int d = rhccq_d2(W.x[j], cs);
            if (d < W.closest[j]) W.closest[j] = d;
        }
        RHCCQ_PAR_FOR(q, 3) W.center[3 * c + q] = (double)((cs >> (16 - 8 * q)) & 255u);
        pot = best_pot;
        __syncthreads();
    }

    // ---- tolerance (kmeans_restated.tolerance)
    double tol;
    {
        const double nn = __dmul_rn((double)n, (double)n);
        double v[3];
        for (int d = 0; d < 3; ++d) v[d] = __ddiv_rn((double)((long long)n * S2[d] - S1[d] * S1[d]), nn);
        tol = __dmul_rn(__ddiv_rn(__dadd_rn(__dadd_rn(v[0], v[1]), v[2]), 3.0), 1e-4);
    }

    // ---- Lloyd (kmeans_restated.kmeans_labels)
    RHCCQ_PAR_FOR(j, n) W.label_old[j] = -1;
    __syncthreads();
    bool strict = false;
    for (int it = 0; it < 300; ++it) {
        rhccq_kmeans_estep(W, n, k);
        RHCCQ_PAR_FOR(q, k) { W.cnt[q] = 0; W.sums[3 * q] = 0; W.sums[3 * q + 1] = 0; W.sums[3 * q + 2] = 0; }
        __syncthreads();
        RHCCQ_PAR_FOR(j, n) {
            const int l = W.label[j];
            const uint32_t c = W.x[j];
            atomicAdd(&W.cnt[l], 1);
            atomicAdd(&W.sums[3 * l], rhccq_key_r(c));
            atomicAdd(&W.sums[3 * l + 1], rhccq_key_g(c));
            atomicAdd(&W.sums[3 * l + 2], rhccq_key_b(c));
        }
        __syncthreads();
        int n_empty = 0;
        RHCCQ_PAR_FOR(q, k) if (W.cnt[q] == 0) ++n_empty;
        n_empty = rhccq_block_sum<int>(n_empty, s_i + 16);
        if (n_empty > 0) {
            // relocate empty clusters to the points farthest from their centre
            // (_k_means_common.pyx:177-211): farthest first, ties to the lower index.
            // `cum` is free here; it holds the bit pattern of each point's own distance.
            double* own = reinterpret_cast<double*>(W.cum);
            double mx = 0.0;
            RHCCQ_PAR_FOR(j, n) {
                const uint32_t c = W.x[j];
                const double d = rhccq_dist3((double)rhccq_key_r(c), (double)rhccq_key_g(c), (double)rhccq_key_b(c),
                                             W.center + 3 * W.label[j]);
                own[j] = d;
                if (d > mx) mx = d;
            }
            mx = rhccq_block_max<double>(mx, s_d);
            if (mx != 0.0) {
                // the empty set is fixed before any relocation (a donor cluster may drop to zero later)
                RHCCQ_PAR_FOR(q, k) W.term[q] = W.cnt[q] == 0 ? 1.0 : 0.0;
                __syncthreads();
                int e = 0;
                for (int done = 0; done < n_empty; ++done) {
                    while (W.term[e] == 0.0) ++e;                   // next empty cluster, ascending (block-uniform)
                    // farthest remaining point: max own, then min index
                    double bm = -1.0;
                    RHCCQ_PAR_FOR(j, n) if (own[j] > bm) bm = own[j];
                    bm = rhccq_block_max<double>(bm, s_d);
                    int bj = 0x7fffffff;
                    RHCCQ_PAR_FOR(j, n) if (own[j] == bm && j < bj) bj = j;
                    bj = rhccq_block_min<int>(bj, s_i + 16);
                    if (threadIdx.x == 0) {
                        const int old = W.label[bj];
                        const uint32_t c = W.x[bj];
                        W.sums[3 * old] -= rhccq_key_r(c); W.sums[3 * old + 1] -= rhccq_key_g(c); W.sums[3 * old + 2] -= rhccq_key_b(c);
                        W.sums[3 * e] = rhccq_key_r(c); W.sums[3 * e + 1] = rhccq_key_g(c); W.sums[3 * e + 2] = rhccq_key_b(c);
                        W.cnt[e] = 1;
                        W.cnt[old] -= 1;
                        own[bj] = -2.0;                             // taken
                    }
                    __syncthreads();
                    ++e;
                }
            }
        }
        RHCCQ_PAR_FOR(q, k) {
            double c0, c1, c2;
            if (W.cnt[q] > 0) {
                const double cn = (double)W.cnt[q];
                c0 = __ddiv_rn((double)W.sums[3 * q], cn);
                c1 = __ddiv_rn((double)W.sums[3 * q + 1], cn);
                c2 = __ddiv_rn((double)W.sums[3 * q + 2], cn);
            } else {
                c0 = __ddiv_rn((double)S1[0], (double)n);
                c1 = __ddiv_rn((double)S1[1], (double)n);
                c2 = __ddiv_rn((double)S1[2], (double)n);
            }
            W.center_new[3 * q] = c0; W.center_new[3 * q + 1] = c1; W.center_new[3 * q + 2] = c2;
            const double a0 = __dsub_rn(c0, W.center[3 * q]), a1 = __dsub_rn(c1, W.center[3 * q + 1]),
                         a2 = __dsub_rn(c2, W.center[3 * q + 2]);
            W.term[q] = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double shift = 0.0;
            for (int q = 0; q < k; ++q) shift = __dadd_rn(shift, W.term[q]);     // fixed order
            s_d[0] = shift;
        }
        int changed = 0;
        RHCCQ_PAR_FOR(j, n) if (W.label[j] != W.label_old[j]) changed = 1;
        RHCCQ_PAR_FOR(q, 3 * k) W.center[q] = W.center_new[q];
        changed = rhccq_block_or(changed, s_i + 16);                // also orders s_d[0] and the centre copy
        __syncthreads();
        const double shift = s_d[0];
        __syncthreads();
        if (!changed) { strict = true; break; }
        if (shift <= tol) break;
        RHCCQ_PAR_FOR(j, n) W.label_old[j] = W.label[j];
        __syncthreads();
    }
    if (!strict) rhccq_kmeans_estep(W, n, k);
    RHCCQ_PAR_FOR(q, k) W.cnt[q] = 0;
    __syncthreads();
    RHCCQ_PAR_FOR(j, n) atomicAdd(&W.cnt[W.label[j]], 1);
    __syncthreads();
}

// ---------------------------------------------------------------- split driver
size_t rhccq_palette_split_ws_bytes(int max_rows) {
    size_t r = (size_t)max_rows, r2 = 1;
    while (r2 < r) r2 <<= 1;
    // x, closest, label, label_old, perm, perm2, csize, crank, llist, stack (3 ints per entry), cum, sort keys,
    // center, center_new, term, sums, cnt
    return rhccq_carve_bytes(r, 4) * 9 + rhccq_carve_bytes(3 * (r + 1), 4) + rhccq_carve_bytes(r, 8)
           + rhccq_carve_bytes(r2, 8) + rhccq_carve_bytes(3 * r, 8) * 2 + rhccq_carve_bytes(r, 8)
           + rhccq_carve_bytes(3 * r, 4) + rhccq_carve_bytes(r, 4);
}

__device__ void rhccq_palette_split_problem(const rhccq_palette_batch& B, int p, const int* __restrict__ labels,
                                            const int* __restrict__ status_in, const int* __restrict__ max_cpc, const double* __restrict__ rng,
                                            int rng_len, int* __restrict__ leaf, int* __restrict__ n_leaves,
                                            int max_rows, unsigned char* wsbase) {
    __shared__ int s_i[RHCCQ_MAX_WARPS + 2 + 16];
    __shared__ long long s_ll[RHCCQ_MAX_WARPS + 2];
    __shared__ double s_d[RHCCQ_MAX_WARPS + 2];
    __shared__ int s_sp, s_next_leaf;
    const int n = B.pal_cnt[p];
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    const int* lab = labels + B.pal_off[p];
    int* lf = leaf + B.pal_off[p];
    const int mcpc = max_cpc[p];
    if (n < 0 || (status_in != nullptr && status_in[p] < 0)) {     // upstream error: pass it on
        if (threadIdx.x == 0) n_leaves[p] = n < 0 ? -2 : status_in[p];
        return;
    }
    if (n > max_rows) {
        if (threadIdx.x == 0) n_leaves[p] = -1;
        return;
    }
    rhccq_carver cv(wsbase);
    rhccq_kmeans_ws W;
    W.x = cv.take<uint32_t>(max_rows);
    W.closest = cv.take<int>(max_rows);
    W.label = cv.take<int>(max_rows);
    W.label_old = cv.take<int>(max_rows);
    int* perm = cv.take<int>(max_rows);
    int* perm2 = cv.take<int>(max_rows);
    int* csize = cv.take<int>(max_rows);
    int* crank = cv.take<int>(max_rows);
    int* llist = cv.take<int>(max_rows);
    int* stack = cv.take<int>(3 * ((size_t)max_rows + 1));
    W.cum = cv.take<long long>(max_rows);
    unsigned long long* skey = cv.take<unsigned long long>(rhccq_next_pow2(max_rows));   // sort keys, power-of-two padded
    W.center = cv.take<double>(3 * (size_t)max_rows);
    W.center_new = cv.take<double>(3 * (size_t)max_rows);
    W.term = cv.take<double>(max_rows);
    W.sums = cv.take<int>(3 * (size_t)max_rows);
    W.cnt = cv.take<int>(max_rows);

    // black rows first, one entry each, in row order (clustering.py:253-255)
    RHCCQ_PAR_FOR(i, n) { perm[i] = (keys[i] == 0u) ? 1 : 0; csize[i] = 0; }
    __syncthreads();
    const int n_black = rhccq_block_excl_scan_array<int>(perm, n, s_i);
    RHCCQ_PAR_FOR(i, n) if (keys[i] == 0u) lf[i] = perm[i];
    // cluster sizes; labels are dense non-negative (a noise row would carry -1: one entry each, :258-264)
    RHCCQ_PAR_FOR(i, n) if (lab[i] >= 0) atomicAdd(&csize[lab[i]], 1);
    __syncthreads();
    RHCCQ_PAR_FOR(i, n) perm[i] = (keys[i] != 0u && lab[i] == -1) ? 1 : 0;
    __syncthreads();
    const int n_noise = rhccq_block_excl_scan_array<int>(perm, n, s_i);
    RHCCQ_PAR_FOR(i, n) if (keys[i] != 0u && lab[i] == -1) lf[i] = n_black + perm[i];
    // small clusters in ascending label order (:273-310)
    RHCCQ_PAR_FOR(l, n) crank[l] = (csize[l] > 0 && csize[l] <= mcpc) ? 1 : 0;
    __syncthreads();
    const int n_small = rhccq_block_excl_scan_array<int>(crank, n, s_i);
    RHCCQ_PAR_FOR(i, n) {
        const int l = lab[i];
        if (l >= 0 && csize[l] <= mcpc) lf[i] = n_black + n_noise + crank[l];
    }
    if (threadIdx.x == 0) s_next_leaf = n_black + n_noise + n_small;
    __syncthreads();

    // large clusters in ascending label order (:315-355), leaves in depth-first K-Means label order
    RHCCQ_PAR_FOR(l, n) crank[l] = csize[l] > mcpc ? 1 : 0;
    __syncthreads();
    const int n_large = rhccq_block_excl_scan_array<int>(crank, n, s_i);
    RHCCQ_PAR_FOR(l, n) if (csize[l] > mcpc) llist[crank[l]] = l;
    __syncthreads();
    for (int li = 0; li < n_large; ++li) {
        const int L = llist[li];
        const int size = csize[L];
        // members in ascending row order
        RHCCQ_PAR_FOR(i, n) perm2[i] = (lab[i] == L) ? 1 : 0;
        __syncthreads();
        rhccq_block_excl_scan_array<int>(perm2, n, s_i);
        RHCCQ_PAR_FOR(i, n) if (lab[i] == L) perm[perm2[i]] = i;
        if (threadIdx.x == 0) { stack[0] = 0; stack[1] = size; stack[2] = 0; s_sp = 1; }
        __syncthreads();
        while (true) {
            __syncthreads();
            const int sp = s_sp;
            if (sp == 0) break;
            const int lo = stack[3 * (sp - 1)], hi = stack[3 * (sp - 1) + 1], force = stack[3 * (sp - 1) + 2];
            const int cnt = hi - lo;
            __syncthreads();
            int k = (cnt + mcpc - 1) / mcpc;                        // clustering.py:739-742
            if (k < 2) k = 2;
            if (k > cnt) k = cnt;
            const bool is_leaf = force || cnt <= mcpc || cnt <= 2 || k < 2;
            if (is_leaf) {
                const int id = s_next_leaf;
                RHCCQ_PAR_FOR(j, cnt) lf[perm[lo + j]] = id;
                __syncthreads();
                if (threadIdx.x == 0) { s_next_leaf = id + 1; s_sp = sp - 1; }
                continue;
            }
            const int need = 1 + (k - 1) * rhccq_kmeans_local_trials(k);
            if (need > rng_len) {                                   // random table too short: report, do not guess
                if (threadIdx.x == 0) n_leaves[p] = -2;
                return;
            }
            RHCCQ_PAR_FOR(j, cnt) W.x[j] = keys[perm[lo + j]];
            __syncthreads();
            rhccq_kmeans(W, cnt, k, rng, s_i, s_ll, s_d);
            // stable partition by K-Means label
            const int np2 = rhccq_next_pow2(cnt);
            RHCCQ_PAR_FOR(j, np2)
                skey[j] = j < cnt ? (((unsigned long long)W.label[j] << 32) | (unsigned)j) : ~0ull;
            __syncthreads();
            rhccq_block_bitonic_sort<unsigned long long>(skey, np2);
            RHCCQ_PAR_FOR(j, cnt) perm2[j] = perm[lo + (int)(skey[j] & 0xffffffffu)];
            __syncthreads();
            RHCCQ_PAR_FOR(j, cnt) perm[lo + j] = perm2[j];
            // children replace the parent on the stack, last label deepest
            if (threadIdx.x == 0) {
                int top = sp - 1;
                int end = hi;
                for (int q = k - 1; q >= 0; --q) {
                    const int c = W.cnt[q];
                    if (c == 0) continue;                           // clustering.py:757
                    stack[3 * top] = end - c;
                    stack[3 * top + 1] = end;
                    stack[3 * top + 2] = (c > mcpc && c < cnt) ? 0 : 1;   // :763-767
                    end -= c;
                    ++top;
                }
                s_sp = top;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) n_leaves[p] = s_next_leaf;
}

__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_palette_split(rhccq_palette_batch B, const int* __restrict__ labels, const int* __restrict__ status_in,
                      const int* __restrict__ max_cpc,
                      const double* __restrict__ rng, int rng_len, int* __restrict__ leaf, int* __restrict__ n_leaves,
                      int max_rows, unsigned char* gws, size_t gws_stride) {
    RHCCQ_DYN_SMEM(dyn);
    unsigned char* wsbase = gws ? gws + (size_t)blockIdx.x * gws_stride : dyn;
    for (int p = blockIdx.x; p < B.n_problems; p += gridDim.x) {
        rhccq_palette_split_problem(B, p, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves, max_rows, wsbase);
        __syncthreads();
    }
}

int rhccq_launch_palette_split(const rhccq_palette_batch& B, const int* labels, const int* status_in, const int* max_cpc,
                               const double* rng, int rng_len, int* leaf, int* n_leaves, int max_rows,
                               rhccq_launch_ws ws, void* stream) {
    if (B.n_problems <= 0) return 0;
    const size_t need = rhccq_palette_split_ws_bytes(max_rows);
    size_t smem; unsigned char* gws;
    const int grid = rhccq_pick_grid((const void*)rhccq_k_palette_split, need, B.n_problems, ws, &smem, &gws,
                                     "rhccq_palette_split");
    if (grid < 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_palette_split, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                 B, labels, status_in, max_cpc, rng, rng_len, leaf, n_leaves, max_rows, gws, need);
    return 0;
}

// ---------------------------------------------------------------- finish: truncated means
__device__ void rhccq_palette_finish_problem(const rhccq_palette_batch& B, int p, const int* __restrict__ leaf,
                                             const int* __restrict__ n_leaves, uint32_t* __restrict__ new_keys,
                                             int max_rows, unsigned char* wsbase) {
    const int n = B.pal_cnt[p];
    const int m = n_leaves[p];
    if (n < 0 || n > max_rows || m < 0 || m > n) return;
    const uint32_t* keys = B.pal_keys + B.pal_off[p];
    const int* lf = leaf + B.pal_off[p];
    uint32_t* out = new_keys + B.pal_off[p];
    rhccq_carver cv(wsbase);
    int* sums = cv.take<int>(4 * (size_t)max_rows);
    RHCCQ_PAR_FOR(q, 4 * m) sums[q] = 0;
    __syncthreads();
    RHCCQ_PAR_FOR(i, n) {
        const int l = lf[i];
        const uint32_t c = keys[i];
        atomicAdd(&sums[4 * l], rhccq_key_r(c));
        atomicAdd(&sums[4 * l + 1], rhccq_key_g(c));
        atomicAdd(&sums[4 * l + 2], rhccq_key_b(c));
        atomicAdd(&sums[4 * l + 3], 1);
    }
    __syncthreads();
    RHCCQ_PAR_FOR(q, m) {
        const int c = sums[4 * q + 3];
        out[q] = c > 0 ? rhccq_pack_rgb(sums[4 * q] / c, sums[4 * q + 1] / c, sums[4 * q + 2] / c) : 0u;
    }
}

size_t rhccq_palette_finish_ws_bytes(int max_rows) { return rhccq_carve_bytes(4 * (size_t)max_rows, 4); }

__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_palette_finish(rhccq_palette_batch B, const int* __restrict__ leaf, const int* __restrict__ n_leaves,
                       uint32_t* __restrict__ new_keys, int max_rows, unsigned char* gws, size_t gws_stride) {
    RHCCQ_DYN_SMEM(dyn);
    unsigned char* wsbase = gws ? gws + (size_t)blockIdx.x * gws_stride : dyn;
    for (int p = blockIdx.x; p < B.n_problems; p += gridDim.x) {
        rhccq_palette_finish_problem(B, p, leaf, n_leaves, new_keys, max_rows, wsbase);
        __syncthreads();
    }
}

int rhccq_launch_palette_finish(const rhccq_palette_batch& B, const int* leaf, const int* n_leaves,
                                uint32_t* new_keys, int max_rows, rhccq_launch_ws ws, void* stream) {
    if (B.n_problems <= 0) return 0;
    const size_t need = rhccq_palette_finish_ws_bytes(max_rows);
    size_t smem; unsigned char* gws;
    const int grid = rhccq_pick_grid((const void*)rhccq_k_palette_finish, need, B.n_problems, ws, &smem, &gws,
                                     "rhccq_palette_finish");
    if (grid < 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_palette_finish, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                 B, leaf, n_leaves, new_keys, max_rows, gws, need);
    return 0;
}
