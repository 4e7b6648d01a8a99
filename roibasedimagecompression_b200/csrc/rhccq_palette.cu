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
    // col, slot_of, order per row; hkey, cstart, cend, parent, minidx, box_lo, box_hi per slot; offset table
    return rhccq_carve_bytes(max_rows, 4) * 3 + rhccq_carve_bytes((size_t)max_slots + 1, 4) * 7
           + rhccq_carve_bytes(128, 4);
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

__device__ __forceinline__ int rhccq_slot_lookup(const rhccq_cellgrid& g, const uint32_t* hkey, uint32_t cell) {
    // cell = packed (r, g, b) cell coordinates; returns its slot or -1
    if (g.direct) return (int)((((cell >> 16) & 255u) * g.G + ((cell >> 8) & 255u)) * g.G + (cell & 255u));
    uint32_t h = (cell * 2654435761u) >> g.hshift;
    while (true) {
        uint32_t k = hkey[h];
        if (k == cell) return (int)h;
        if (k == RHCCQ_EMPTY_KEY) return -1;
        h = (h + 1) & (uint32_t)(g.hcap - 1);
    }
}

// squared distance between a colour and a box / between two boxes (packed bytes, exact)
__device__ __forceinline__ int rhccq_d2_point_box(uint32_t c, uint32_t lo, uint32_t hi) {
    const unsigned gap = __vadd4(__vsubus4(lo, c), __vsubus4(c, hi));   // one of the two terms is 0 per channel
    return (int)__dp4a(gap, gap, 0u);
}
__device__ __forceinline__ int rhccq_d2_box_box_min(uint32_t loa, uint32_t hia, uint32_t lob, uint32_t hib) {
    const unsigned gap = __vadd4(__vsubus4(loa, hib), __vsubus4(lob, hia));
    return (int)__dp4a(gap, gap, 0u);
}
__device__ __forceinline__ int rhccq_d2_box_box_max(uint32_t loa, uint32_t hia, uint32_t lob, uint32_t hib) {
    const unsigned far = __vmaxu4(__vabsdiffu4(hia, lob), __vabsdiffu4(hib, loa));
    return (int)__dp4a(far, far, 0u);
}

// The join phase is organised by cells, one warp per occupied cell: its lanes look up the forward
// neighbour cells (offsets decoded from a small table, not by division), and every occupied neighbour
// that is not yet in the same set is tested by the whole warp: bounding boxes first (certainly apart /
// certainly all within eps), then lanes over the points of one cell against the points of the other,
// leaving at the first pair within eps.
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
    int* slot_of = cv.take<int>(max_rows);
    int* order = cv.take<int>(max_rows);                           // rows grouped by cell; later the label scan
    uint32_t* hkey = cv.take<uint32_t>((size_t)max_slots + 1);
    int* cstart = cv.take<int>((size_t)max_slots + 1);
    int* cend = cv.take<int>((size_t)max_slots + 1);
    int* parent = cv.take<int>((size_t)max_slots + 1);
    int* minidx = cv.take<int>((size_t)max_slots + 1);
    uint32_t* box_lo = cv.take<uint32_t>((size_t)max_slots + 1);
    uint32_t* box_hi = cv.take<uint32_t>((size_t)max_slots + 1);
    int* offtab = cv.take<int>(128);                               // forward neighbour offsets, 3 signed bytes each

    const int side = 2 * g.reach + 1;                              // reach <= 2: at most 5^3 cells around a cell
    const int half = (side * side * side - 1) / 2;                 // offsets after the centre in raster order
    RHCCQ_PAR_FOR(s, g.hcap) { hkey[s] = RHCCQ_EMPTY_KEY; cstart[s] = 0; parent[s] = s; minidx[s] = 0x7fffffff; }
    RHCCQ_PAR_FOR(o, half) {
        const int L = half + 1 + o;
        const int dr = L / (side * side) - g.reach, dg = (L / side) % side - g.reach, db = L % side - g.reach;
        offtab[o] = ((dr & 255) << 16) | ((dg & 255) << 8) | (db & 255);
    }
    RHCCQ_PAR_FOR(i, n) col[i] = keys[i];
    __syncthreads();

    // cell of every row; black rows take no part (clustering.py:185-192)
    RHCCQ_PAR_FOR(i, n) {
        const uint32_t c = col[i];
        if (c == 0u) { slot_of[i] = -1; continue; }
        const uint32_t cell = ((uint32_t)(rhccq_key_r(c) / g.s) << 16) | ((uint32_t)(rhccq_key_g(c) / g.s) << 8)
                              | (uint32_t)(rhccq_key_b(c) / g.s);
        int slot;
        if (g.direct) {
            slot = rhccq_slot_lookup(g, hkey, cell);
            hkey[slot] = cell;                                     // every writer stores the same value
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
        atomicAdd(&cstart[slot], 1);
        atomicMin(&minidx[slot], i);
    }
    __syncthreads();
    rhccq_block_excl_scan_array<int>(cstart, g.hcap, s_scratch);
    RHCCQ_PAR_FOR(s, g.hcap) cend[s] = cstart[s];
    __syncthreads();
    RHCCQ_PAR_FOR(i, n) if (slot_of[i] >= 0) order[atomicAdd(&cend[slot_of[i]], 1)] = i;
    __syncthreads();
    // bounding box of every occupied cell
    for (int s = RHCCQ_WARP; s < g.hcap; s += RHCCQ_NWARPS) {
        const int a0 = cstart[s], a1 = cend[s];
        if (a1 <= a0) continue;
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int j = a0 + RHCCQ_LANE; j < a1; j += RHCCQ_WARP_SIZE) {
            const uint32_t c = col[order[j]];
            lo = __vminu4(lo, c); hi = __vmaxu4(hi, c);
        }
        for (int d = RHCCQ_WARP_SIZE >> 1; d > 0; d >>= 1) {
            lo = __vminu4(lo, rhccq_shfl_xor(lo, d)); hi = __vmaxu4(hi, rhccq_shfl_xor(hi, d));
        }
        if (RHCCQ_LANE == 0) { box_lo[s] = lo & 0x00ffffffu; box_hi[s] = hi; }
    }
    __syncthreads();

    // join cells
    for (int s = RHCCQ_WARP; s < g.hcap; s += RHCCQ_NWARPS) {
        const int a0 = cstart[s], a1 = cend[s];
        if (a1 <= a0) continue;
        const uint32_t cell = hkey[s];
        const int cr = (int)((cell >> 16) & 255u), cg = (int)((cell >> 8) & 255u), cb = (int)(cell & 255u);
        const uint32_t alo = box_lo[s], ahi = box_hi[s];
        for (int o0 = 0; o0 < half; o0 += RHCCQ_WARP_SIZE) {
            const int o = o0 + RHCCQ_LANE;
            int b = -1;
            if (o < half) {
                const int e = offtab[o];
                const int nr = cr + (int)(signed char)(e >> 16), ng = cg + (int)(signed char)(e >> 8),
                          nb = cb + (int)(signed char)e;
                if (nr >= 0 && ng >= 0 && nb >= 0 && nr < g.G && ng < g.G && nb < g.G) {
                    b = rhccq_slot_lookup(g, hkey, ((uint32_t)nr << 16) | ((uint32_t)ng << 8) | (uint32_t)nb);
                    if (b >= 0 && cend[b] <= cstart[b]) b = -1;
                }
            }
            unsigned found = rhccq_ballot(b >= 0);
            while (found) {
                const int src = __ffs((int)found) - 1;
                found &= found - 1;
                const int bb = rhccq_shfl(b, src);
                // other warps move parents concurrently: one lane decides, so that the warp stays converged
                int same = 0;
                if (RHCCQ_LANE == 0) same = rhccq_uf_find(parent, s) == rhccq_uf_find(parent, bb);
                if (rhccq_shfl(same, 0)) continue;
                const uint32_t blo = box_lo[bb], bhi = box_hi[bb];
                if (rhccq_d2_box_box_min(alo, ahi, blo, bhi) > thr) continue;        // certainly apart
                const int far = rhccq_d2_box_box_max(alo, ahi, blo, bhi);
                bool joined = tie ? far < thr : far <= thr;                           // certainly all within eps
                if (!joined) {
                    const int b0 = cstart[bb], b1 = cend[bb];
                    for (int ja = a0; ja < a1 && !joined; ja += RHCCQ_WARP_SIZE) {
                        const int j = ja + RHCCQ_LANE;
                        const uint32_t c = j < a1 ? col[order[j]] : 0u;
                        const bool live = j < a1 && rhccq_d2_point_box(c, blo, bhi) <= thr;
                        if (!rhccq_any(live)) continue;
                        for (int jb = b0; jb < b1; ++jb) {
                            const uint32_t cj = col[order[jb]];
                            bool hit = false;
                            if (live) {
                                const int d2 = rhccq_d2(c, cj);
                                hit = tie ? (d2 < thr) : (d2 <= thr);
                                if (!hit && tie && d2 == thr) hit = rhccq_tie_accept(c, cj, eps);
                            }
                            if (rhccq_any(hit)) { joined = true; break; }
                        }
                    }
                }
                if (joined && RHCCQ_LANE == 0) rhccq_uf_union(parent, s, bb);
                __syncwarp();
            }
        }
    }
    __syncthreads();

    // canonical numbering: clusters in order of their lowest row (sklearn's DFS seeds ascend)
    RHCCQ_PAR_FOR(s, g.hcap) {
        if (cend[s] > cstart[s]) {
            const int r = rhccq_uf_find(parent, s);
            if (r != s) atomicMin(&minidx[r], minidx[s]);          // a concurrently lowered value is a member too
        }
    }
    RHCCQ_PAR_FOR(i, n) order[i] = 0;
    __syncthreads();
    RHCCQ_PAR_FOR(s, g.hcap) if (cend[s] > cstart[s] && parent[s] == s) order[minidx[s]] = 1;
    __syncthreads();
    const int total = rhccq_block_excl_scan_array<int>(order, n, s_scratch);
    RHCCQ_PAR_FOR(i, n) lab[i] = slot_of[i] >= 0 ? order[minidx[rhccq_uf_find(parent, slot_of[i])]] : -2;
    if (threadIdx.x == 0) n_clusters[p] = total;
}

#define RHCCQ_DBSCAN_THREADS 256

__global__ void __launch_bounds__(RHCCQ_DBSCAN_THREADS)
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
    RHCCQ_LAUNCH(rhccq_k_palette_dbscan, grid, RHCCQ_DBSCAN_THREADS, smem, (cudaStream_t)stream,
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
