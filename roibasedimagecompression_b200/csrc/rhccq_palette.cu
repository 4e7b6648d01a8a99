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
