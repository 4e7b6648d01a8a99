// DBSCAN of large point clouds (x, y, colour ...) — the operator the reference calls as
// sklearn.cluster.DBSCAN(eps, min_samples).fit_predict(X)
// (/root/reference/encoder/compression/clustering.py:233-235; semantics:
// sklearn/cluster/_dbscan.py:397-470, _dbscan_inner.pyx), for N up to tens of millions of points.
//
//   bin      points -> cells of side eps over the first 2 or 3 coordinates; counting sort into 32-byte
//            records (coordinates, neighbour count, original index), cells in raster order
//   count    number of points within eps of every point (itself included); core <=> count >= min_samples
//   union    lock-free union-find over core points that are within eps of each other; a set's root is its
//            lowest original index, which is also the order sklearn numbers clusters in
//   border   pointer-jumping flatten; a non-core point takes the lowest-rooted cluster among the core
//            points within eps (the cluster sklearn's depth-first labelling reaches it from first)
//   relabel  cluster number = rank of the root among roots
//
// count / union / border share one traversal.  A CTA owns a run of consecutive cells of one cell row;
// the points of those cells and of the neighbouring cells (3^(grid dims) cells around each) are
// contiguous spans of the sorted record array, one span per neighbouring row, so they are staged into
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier, double buffered) and every thread
// walks only the records of the 3 cells around its own point in each row.
//
// Distances: scikit-learn's KD-tree evaluates sum_d (x_d - y_d)^2 <= eps^2 in float64 on the float32
// inputs.  The kernel evaluates the sum in float32 and, when it lands within a relative 2^-18 of eps^2
// (well above the float32 evaluation error of (dims + 3) * 2^-24), decides in float64 without contraction.
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"
#include "../../include/rhccq.h"

#define RHCCQ_PT_THREADS 256
#define RHCCQ_PT_CAP 1024                 // records per staging buffer (32 KiB)
#define RHCCQ_REC_FLOATS 8
#define RHCCQ_REC_COUNT 6                 // slot of the neighbour count (int bits)
#define RHCCQ_REC_ID 7                    // slot of the original index (int bits)

struct rhccq_pt_grid {
    int n, dims, gd, min_pts;
    int nc0, nc1, nc2;                    // cells per grid dimension (nc2 = 1 for a 2-D grid)
    int cw;                               // cells per tile along dimension 0
    int tiles_per_row, rows;
    double o0, o1, o2, inv_side;
    float r2f, r2lo, r2hi;
    double r2;
};

__host__ __device__ static inline int rhccq_pt_cell(double x, double o, double inv, int nc) {
    int c = (int)floor((x - o) * inv);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

// ---------------------------------------------------------------- bounds of the grid coordinates
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_bounds(const float* __restrict__ pts, int n, int dims, int gd, float* __restrict__ partial) {
    __shared__ float s_f[RHCCQ_MAX_WARPS + 2];
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        for (int d = 0; d < gd; ++d) {
            const float v = pts[i * dims + d];
            lo[d] = v < lo[d] ? v : lo[d];
            hi[d] = v > hi[d] ? v : hi[d];
        }
    for (int d = 0; d < 3; ++d) {
        const float a = rhccq_block_min<float>(lo[d], s_f);
        const float b = rhccq_block_max<float>(hi[d], s_f);
        if (threadIdx.x == 0) { partial[blockIdx.x * 6 + d] = a; partial[blockIdx.x * 6 + 3 + d] = b; }
    }
}
__global__ void rhccq_k_pt_bounds_final(const float* __restrict__ partial, int nblocks, double* __restrict__ out) {
    // one small block: out[0..3) = min, out[3..6) = max
    RHCCQ_PAR_FOR(d, 6) {
        float v = partial[d];
        for (int b = 1; b < nblocks; ++b) {
            const float w = partial[b * 6 + d];
            v = d < 3 ? (w < v ? w : v) : (w > v ? w : v);
        }
        out[d] = (double)v;
    }
}

// ---------------------------------------------------------------- multi-block exclusive scan (int32)
#define RHCCQ_SCAN_ITEMS 8
#define RHCCQ_SCAN_TILE (RHCCQ_PT_THREADS * RHCCQ_SCAN_ITEMS)
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_scan_partial(const int* __restrict__ in, long long n, int* __restrict__ block_sums) {
    __shared__ int s_i[RHCCQ_MAX_WARPS + 2];
    const long long base = (long long)blockIdx.x * RHCCQ_SCAN_TILE;
    int s = 0;
    for (int k = (int)threadIdx.x; k < RHCCQ_SCAN_TILE; k += (int)blockDim.x)
        if (base + k < n) s += in[base + k];
    s = rhccq_block_sum<int>(s, s_i);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_scan_apply(const int* __restrict__ in, long long n, const int* __restrict__ block_offsets, int* __restrict__ out) {
    __shared__ int s_i[RHCCQ_MAX_WARPS + 2];
    // every thread owns a contiguous chunk of the tile: sum it, scan the sums, then write (reading each
    // input again just before its slot is written, so out may alias in)
    const int per = RHCCQ_SCAN_TILE / (int)blockDim.x;
    const long long base = (long long)blockIdx.x * RHCCQ_SCAN_TILE + (long long)threadIdx.x * per;
    int s = 0;
    for (int k = 0; k < per; ++k) if (base + k < n) s += in[base + k];
    int total;
    int run = rhccq_block_excl_scan<int>(s, &total, s_i) + block_offsets[blockIdx.x];
    for (int k = 0; k < per; ++k) if (base + k < n) { const int v = in[base + k]; out[base + k] = run; run += v; }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_scan_small(int* __restrict__ a, int n, int* __restrict__ total_out) {     // one block, in place
    __shared__ int s_i[RHCCQ_MAX_WARPS + 2];
    const int t = rhccq_block_excl_scan_array<int>(a, n, s_i);
    if (threadIdx.x == 0 && total_out) *total_out = t;
}

// out[i] = sum_{j<i} in[i] for i < n (out may alias in); scratch: ints for the block sums of every level
static size_t rhccq_scan_scratch_ints(long long n) {
    size_t tot = 0;
    while (n > RHCCQ_SCAN_TILE) { n = (n + RHCCQ_SCAN_TILE - 1) / RHCCQ_SCAN_TILE; tot += (size_t)n + 1; }
    return tot + 1;
}
static int rhccq_scan_i32(const int* in, long long n, int* out, int* scratch, void* stream) {
    if (n <= 0) return 0;
    if (n <= RHCCQ_SCAN_TILE) {
        if (in != out) { rhccq_set_error("rhccq_scan_i32: small scans run in place"); return -1; }
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, out, (int)n, (int*)nullptr);
        return 0;
    }
    const long long nb = (n + RHCCQ_SCAN_TILE - 1) / RHCCQ_SCAN_TILE;
    RHCCQ_LAUNCH(rhccq_k_scan_partial, (unsigned)nb, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, in, n, scratch);
    if (rhccq_scan_i32(scratch, nb, scratch, scratch + nb + 1, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_scan_apply, (unsigned)nb, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, in, n, scratch, out);
    return 0;
}

// ---------------------------------------------------------------- bin
// a record is 32 bytes, 32-byte aligned: two 16-byte accesses
__device__ __forceinline__ void rhccq_load_rec(const float* r, float* v) {
#ifdef RHCCQ_HOST_EMU
    for (int d = 0; d < RHCCQ_REC_FLOATS; ++d) v[d] = r[d];
#else
    const float4 u = reinterpret_cast<const float4*>(r)[0], w = reinterpret_cast<const float4*>(r)[1];
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; v[4] = w.x; v[5] = w.y; v[6] = w.z; v[7] = w.w;
#endif
}
__device__ __forceinline__ void rhccq_store_rec(float* r, const float* v) {
#ifdef RHCCQ_HOST_EMU
    for (int d = 0; d < RHCCQ_REC_FLOATS; ++d) r[d] = v[d];
#else
    reinterpret_cast<float4*>(r)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(r)[1] = make_float4(v[4], v[5], v[6], v[7]);
#endif
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_cellid(const float* __restrict__ pts, rhccq_pt_grid G, int* __restrict__ cell_id, int* __restrict__ cell_cnt) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.n; i += (long long)gridDim.x * blockDim.x) {
        const float* p = pts + i * G.dims;
        const int c0 = rhccq_pt_cell((double)p[0], G.o0, G.inv_side, G.nc0);
        const int c1 = rhccq_pt_cell((double)p[1], G.o1, G.inv_side, G.nc1);
        const int c2 = G.gd > 2 ? rhccq_pt_cell((double)p[2], G.o2, G.inv_side, G.nc2) : 0;
        const int key = (c2 * G.nc1 + c1) * G.nc0 + c0;
        cell_id[i] = key;
        atomicAdd(&cell_cnt[key], 1);
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_scatter(const float* __restrict__ pts, rhccq_pt_grid G, const int* __restrict__ cell_id,
                   const int* __restrict__ cell_start, int* __restrict__ cell_fill, float* __restrict__ rec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.n; i += (long long)gridDim.x * blockDim.x) {
        const int key = cell_id[i];
        const int pos = cell_start[key] + atomicAdd(&cell_fill[key], 1);
        const float* p = pts + i * G.dims;
        float v[RHCCQ_REC_FLOATS];
        for (int d = 0; d < RHCCQ_REC_FLOATS; ++d) v[d] = d < G.dims ? p[d] : 0.0f;
        int id = (int)i;
        memcpy(&v[RHCCQ_REC_ID], &id, 4);
        rhccq_store_rec(rec + (size_t)pos * RHCCQ_REC_FLOATS, v);
    }
}

// ---------------------------------------------------------------- staging: TMA bulk copy + mbarrier
#ifndef RHCCQ_HOST_EMU
__device__ __forceinline__ uint32_t rhccq_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rhccq_mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rhccq_smem_addr(bar)));
}
__device__ __forceinline__ void rhccq_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rhccq_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rhccq_smem_addr(dst)), "l"(src), "r"(bytes), "r"(rhccq_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rhccq_mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        ::"r"(rhccq_smem_addr(bar)), "r"(phase) : "memory");
}
#endif

// float32 evaluation with a float64 decision near the radius (see the header)
__device__ __forceinline__ bool rhccq_pt_within(const float* a, const float* b, const rhccq_pt_grid& G) {
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < 6; ++d) { const float t = a[d] - b[d]; s = fmaf(t, t, s); }     // slots past dims hold 0 in both
    if (s <= G.r2lo) return true;
    if (s > G.r2hi) return false;
    double sd = 0.0;
    for (int d = 0; d < 6; ++d) {
        const double t = __dsub_rn((double)a[d], (double)b[d]);
        sd = __dadd_rn(sd, __dmul_rn(t, t));
    }
    return sd <= G.r2;
}

__device__ __forceinline__ int rhccq_pt_find(int* parent, int x) {
    while (true) {
        const int p = ((volatile int*)parent)[x];
        if (p == x) return x;
        const int gp = ((volatile int*)parent)[p];
        if (gp != p) parent[x] = gp;                               // path halving; racing writers only shorten paths
        x = p;
    }
}
__device__ __forceinline__ int rhccq_pt_find_ro(const int* parent, int x) {     // no path compression: no writes
    while (true) {
        const int p = parent[x];
        if (p == x) return x;
        x = p;
    }
}
__device__ __forceinline__ void rhccq_pt_union(int* parent, int a, int b) {
    while (true) {
        a = rhccq_pt_find(parent, a);
        b = rhccq_pt_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }              // the larger root goes under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

// MODE 0: count + core flag, 1: union of core pairs, 2: labels of roots (border attachment)
template <int MODE>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_sweep(rhccq_pt_grid G, float* __restrict__ rec, const int* __restrict__ cell_start,
                 uint8_t* __restrict__ core, int* __restrict__ parent, int* __restrict__ rootlab) {
    RHCCQ_DYN_SMEM(dyn);
    float* buf0 = reinterpret_cast<float*>(dyn);
    float* buf1 = buf0 + (size_t)RHCCQ_PT_CAP * RHCCQ_REC_FLOATS;
#ifndef RHCCQ_HOST_EMU
    __shared__ __align__(8) uint64_t s_bar[2];
    if (threadIdx.x == 0) {
        rhccq_mbar_init(&s_bar[0]); rhccq_mbar_init(&s_bar[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase[2] = {0u, 0u};
#endif
    const int nrows_nb = G.gd > 2 ? 9 : 3;
    for (int tile = blockIdx.x; tile < G.rows * G.tiles_per_row; tile += gridDim.x) {
        const int row = tile / G.tiles_per_row, bx = tile % G.tiles_per_row;
        const int cy = row % G.nc1, cz = row / G.nc1;
        const int cx0 = bx * G.cw, cx1 = (cx0 + G.cw < G.nc0 ? cx0 + G.cw : G.nc0) - 1;
        const int rowkey = row * G.nc0;
        const int t0 = cell_start[rowkey + cx0], t1 = cell_start[rowkey + cx1 + 1];
        if (t1 <= t0) continue;                                    // no point in the tile (block-uniform)
        const int sx0 = cx0 > 0 ? cx0 - 1 : 0, sx1 = cx1 + 1 < G.nc0 ? cx1 + 1 : G.nc0 - 1;
        for (int c0 = t0; c0 < t1; c0 += (int)blockDim.x) {        // centre points, one per thread
            const int me = c0 + (int)threadIdx.x;
            const bool have = me < t1;
            float a[RHCCQ_REC_FLOATS];
            int my_id = -1, my_cnt = 0, mycx = cx0;
            bool active = false;
            if (have) {
                rhccq_load_rec(rec + (size_t)me * RHCCQ_REC_FLOATS, a);
                memcpy(&my_id, &a[RHCCQ_REC_ID], 4);
                memcpy(&my_cnt, &a[RHCCQ_REC_COUNT], 4);
                mycx = rhccq_pt_cell((double)a[0], G.o0, G.inv_side, G.nc0);
                active = MODE == 0 ? true : (MODE == 1 ? my_cnt >= G.min_pts : my_cnt < G.min_pts);
            }
            int acc = MODE == 2 ? 0x7fffffff : 0;
            int my_root = my_id;                                   // MODE 1: last known root of my set
            const int mx0 = mycx > 0 ? mycx - 1 : 0, mx1 = mycx + 1 < G.nc0 ? mycx + 1 : G.nc0 - 1;
            // stages: (neighbour row, chunk of its span); the next stage is in flight while this one is read
            int ri = 0, off = 0;                                   // current stage
            int s_lo = 0, s_hi = 0, s_key = 0;
            auto row_span = [&](int r, int& lo, int& hi, int& key) -> bool {
                const int dy = r % 3 - 1, dz = r / 3 - (G.gd > 2 ? 1 : 0);
                const int ny = cy + dy, nz = cz + dz;
                if (ny < 0 || ny >= G.nc1 || nz < 0 || nz >= G.nc2) { lo = hi = 0; key = 0; return false; }
                key = (nz * G.nc1 + ny) * G.nc0;
                lo = cell_start[key + sx0]; hi = cell_start[key + sx1 + 1];
                return hi > lo;
            };
            auto next_stage = [&](int& r, int& o, int& lo, int& hi, int& key) -> bool {
                // advance (r, o) to the next non-empty chunk; r == nrows_nb means done
                while (r < nrows_nb) {
                    if (o == 0) { if (!row_span(r, lo, hi, key)) { ++r; continue; } }
                    if (lo + o < hi) return true;
                    ++r; o = 0;
                }
                return false;
            };
            int stage = 0;
            bool more = next_stage(ri, off, s_lo, s_hi, s_key);
#ifndef RHCCQ_HOST_EMU
            if (more && threadIdx.x == 0) {
                const int cnt = s_hi - (s_lo + off) < RHCCQ_PT_CAP ? s_hi - (s_lo + off) : RHCCQ_PT_CAP;
                rhccq_bulk_load(buf0, rec + (size_t)(s_lo + off) * RHCCQ_REC_FLOATS, (uint32_t)cnt * 32u, &s_bar[0]);
            }
#endif
            while (more) {
                const int cur_lo = s_lo + off;
                const int cur_n = s_hi - cur_lo < RHCCQ_PT_CAP ? s_hi - cur_lo : RHCCQ_PT_CAP;
                const int cur_key = s_key;
                float* cur = (stage & 1) ? buf1 : buf0;
                // look ahead
                int nri = ri, noff = off + cur_n, nlo = s_lo, nhi = s_hi, nkey = s_key;
                const bool nmore = next_stage(nri, noff, nlo, nhi, nkey);
#ifdef RHCCQ_HOST_EMU
                memcpy(cur, rec + (size_t)cur_lo * RHCCQ_REC_FLOATS, (size_t)cur_n * 32);
#else
                if (nmore && threadIdx.x == 0) {
                    const int cnt = nhi - (nlo + noff) < RHCCQ_PT_CAP ? nhi - (nlo + noff) : RHCCQ_PT_CAP;
                    rhccq_bulk_load((stage & 1) ? buf0 : buf1, rec + (size_t)(nlo + noff) * RHCCQ_REC_FLOATS,
                                    (uint32_t)cnt * 32u, &s_bar[(stage + 1) & 1]);
                }
                rhccq_mbar_wait(&s_bar[stage & 1], phase[stage & 1]);
                phase[stage & 1] ^= 1u;
#endif
                if (active) {
                    // records of the 3 cells around my cell in this row, clipped to the staged chunk
                    int lo = cell_start[cur_key + mx0], hi = cell_start[cur_key + mx1 + 1];
                    lo = lo > cur_lo ? lo : cur_lo;
                    hi = hi < cur_lo + cur_n ? hi : cur_lo + cur_n;
                    for (int j = lo; j < hi; ++j) {
                        float b[RHCCQ_REC_FLOATS];
                        rhccq_load_rec(cur + (size_t)(j - cur_lo) * RHCCQ_REC_FLOATS, b);
                        if (MODE != 0) {
                            int bc; memcpy(&bc, &b[RHCCQ_REC_COUNT], 4);
                            if (bc < G.min_pts) continue;          // only core candidates matter
                        }
                        if (!rhccq_pt_within(a, b, G)) continue;
                        if (MODE == 0) ++acc;
                        else {
                            int bid; memcpy(&bid, &b[RHCCQ_REC_ID], 4);
                            if (MODE == 1) {
                                // one hop is enough to see that a neighbour already hangs under my root (the
                                // common case once a dense region is linked); only otherwise walk and link
                                if (bid < my_id && ((volatile int*)parent)[bid] != my_root) {
                                    rhccq_pt_union(parent, my_id, bid);
                                    my_root = rhccq_pt_find(parent, my_id);
                                }
                            } else { const int r = rootlab[bid]; acc = r < acc ? r : acc; }
                        }
                    }
                }
                __syncthreads();                                   // everyone is done with `cur` before it is refilled
                ri = nri; off = noff; s_lo = nlo; s_hi = nhi; s_key = nkey; more = nmore;
                ++stage;
            }
            if (have) {
                if (MODE == 0) {
                    memcpy(&rec[(size_t)me * RHCCQ_REC_FLOATS + RHCCQ_REC_COUNT], &acc, 4);
                    core[my_id] = acc >= G.min_pts ? 1 : 0;
                } else if (MODE == 2) {
                    // core points keep the root the flatten pass stored; others take the lowest adjacent root
                    if (my_cnt < G.min_pts) rootlab[my_id] = acc == 0x7fffffff ? -1 : acc;
                }
            }
            __syncthreads();                                       // (barrier parities are tracked per buffer)
        }
    }
}

__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_init_parent(int n, int* __restrict__ parent) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        parent[i] = (int)i;
}
// Root of every core point into rootlab (read-only walk: a compressing walk could overwrite a finished
// entry with a stale ancestor), -1 for the others; is_root marks the roots.
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_flatten(int n, const int* __restrict__ parent, const uint8_t* __restrict__ core, int* __restrict__ rootlab,
                   int* __restrict__ is_root) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = core[i] ? rhccq_pt_find_ro(parent, (int)i) : -1;
        rootlab[i] = r;
        is_root[i] = r == (int)i ? 1 : 0;
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_pt_labels(int n, const int* __restrict__ rootlab, const int* __restrict__ root_rank, int* __restrict__ labels) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        labels[i] = r < 0 ? -1 : root_rank[r];
    }
}

// ---------------------------------------------------------------- strips: boundary edges, merge, lookup
// A very large point set is split into strips of consecutive points (rows of an image), one per GPU, each
// with a halo of 2 eps: neighbour counts are then exact for the strip's own points and for the halo points
// within eps of it, so every edge of the global core graph is seen by at least one rank.  Local components
// are glued at the points two ranks share: for every core point p of the boundary zone a rank emits the edge
// (p, local root of p) in global indices; all ranks gather all edges (NCCL) and run the same union-find over
// them (roots = lowest global index), which maps local roots to global roots.
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_emit(const int* __restrict__ rootlab, int lo, int hi, int g0, int* __restrict__ edges, int* __restrict__ counter,
                int capacity) {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        if (r < 0 || r == (int)i) continue;                        // not core, or its own root: nothing to glue
        const int slot = atomicAdd(counter, 1);
        if (slot < capacity) { edges[2 * slot] = g0 + (int)i; edges[2 * slot + 1] = g0 + r; }
    }
}

__device__ __forceinline__ int rhccq_tab_insert(int* keys, int cap_mask, int id) {
    uint32_t h = ((uint32_t)id * 2654435761u) & (uint32_t)cap_mask;
    while (true) {
        const int k = atomicCAS(&keys[h], -1, id);
        if (k == -1 || k == id) return (int)h;
        h = (h + 1) & (uint32_t)cap_mask;
    }
}
__device__ __forceinline__ int rhccq_tab_find(const int* keys, int cap_mask, int id) {
    uint32_t h = ((uint32_t)id * 2654435761u) & (uint32_t)cap_mask;
    while (true) {
        const int k = keys[h];
        if (k == id) return (int)h;
        if (k == -1) return -1;
        h = (h + 1) & (uint32_t)cap_mask;
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_tab_init(int* __restrict__ keys, int* __restrict__ parent, int cap) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x) {
        keys[i] = -1; parent[i] = (int)i;
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_tab_insert(const int* __restrict__ edges, int n_edges, int* __restrict__ keys, int cap_mask) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < 2LL * n_edges; e += (long long)gridDim.x * blockDim.x)
        rhccq_tab_insert(keys, cap_mask, edges[e]);
}
// union by key: the slot whose key (global index) is larger goes under the other
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_tab_union(const int* __restrict__ edges, int n_edges, const int* __restrict__ keys, int* __restrict__ parent,
                     int cap_mask) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
        int a = rhccq_tab_find(keys, cap_mask, edges[2 * e]), b = rhccq_tab_find(keys, cap_mask, edges[2 * e + 1]);
        while (true) {
            a = rhccq_pt_find(parent, a);
            b = rhccq_pt_find(parent, b);
            if (a == b) break;
            if (keys[a] < keys[b]) { const int t = a; a = b; b = t; }      // a: larger key
            if (atomicCAS(&parent[a], a, b) == a) break;
        }
    }
}
// rootlab (local roots, local indices; -1 = not core) -> global roots (global indices) through the table
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_lookup(int* __restrict__ rootlab, int n, int g0, const int* __restrict__ keys, const int* __restrict__ parent,
                  int cap_mask) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        if (r < 0) continue;
        int gr = g0 + r;
        if (cap_mask >= 0) {
            const int s = rhccq_tab_find(keys, cap_mask, gr);
            if (s >= 0) gr = keys[rhccq_pt_find_ro(parent, s)];
        }
        rootlab[i] = gr;
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_root_flags(const int* __restrict__ rootlab, int n, int lo, int hi, int g0, int* __restrict__ flags) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        flags[i] = (i >= lo && i < hi && rootlab[i] == g0 + (int)i) ? 1 : 0;
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_root_scatter(const int* __restrict__ rootlab, const int* __restrict__ rank, int lo, int hi, int g0,
                        int* __restrict__ out_ids, int* __restrict__ out_count, int n, int out_cap) {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x)
        if (rootlab[i] == g0 + (int)i && rank[i] < out_cap) out_ids[rank[i]] = g0 + (int)i;
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = hi < n ? rank[hi] : rank[n - 1] + (rootlab[n - 1] == g0 + n - 1 && n - 1 >= lo ? 1 : 0);
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_rank(const int* __restrict__ sorted_roots, int n_roots, const int* __restrict__ rootlab, int lo, int hi,
                int* __restrict__ labels) {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        int a = 0, b = n_roots;
        while (a < b) { const int m = (a + b) >> 1; if (sorted_roots[m] < r) a = m + 1; else b = m; }
        labels[i - lo] = (r < 0 || a >= n_roots || sorted_roots[a] != r) ? -1 : a;
    }
}

// The same on gathered buffers, without the host in between: every rank contributes one fixed-capacity block of
// `stride` ints — [0] = number of payload rows, payload from int 2 — so that one all-gather of equal blocks
// carries counts and rows together and nothing has to be sized on the host.
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_tab_insert_g(const int* __restrict__ g, int world, int stride, int cap_rows, int* __restrict__ keys, int cap_mask) {
    const long long per = 2LL * cap_rows, total = per * world;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / per);
        const long long i = e - r * per;
        const int* blk = g + (long long)r * stride;
        if (i < 2LL * blk[0]) rhccq_tab_insert(keys, cap_mask, blk[2 + i]);
    }
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_tab_union_g(const int* __restrict__ g, int world, int stride, int cap_rows, const int* __restrict__ keys,
                       int* __restrict__ parent, int cap_mask) {
    const long long total = (long long)cap_rows * world;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / cap_rows);
        const long long i = e - (long long)r * cap_rows;
        const int* blk = g + (long long)r * stride;
        if (i >= blk[0]) continue;
        int a = rhccq_tab_find(keys, cap_mask, blk[2 + 2 * i]), b = rhccq_tab_find(keys, cap_mask, blk[2 + 2 * i + 1]);
        while (true) {
            a = rhccq_pt_find(parent, a);
            b = rhccq_pt_find(parent, b);
            if (a == b) break;
            if (keys[a] < keys[b]) { const int t = a; a = b; b = t; }      // a: larger key
            if (atomicCAS(&parent[a], a, b) == a) break;
        }
    }
}
// label = rank of the point's global root among the roots of all ranks; the blocks hold ascending ids and the
// ranks own ascending index ranges, so the rank is (rows of the lower blocks) + (position inside its block).
// A block whose count exceeds its capacity (more roots than the caller provided for) marks every label -2.
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_uf_rank_g(const int* __restrict__ g, int world, int stride, int cap_rows, const int* __restrict__ rootlab, int lo, int hi,
                  int* __restrict__ labels) {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        int lab = -1, before = 0;
        bool over = false;
        for (int w = 0; w < world; ++w) {
            const int* blk = g + (long long)w * stride;
            const int c = blk[0];
            if (c > cap_rows) over = true;
            if (r >= 0 && lab < 0 && c > 0 && c <= cap_rows && blk[2] <= r && r <= blk[2 + c - 1]) {
                int a = 0, b = c;
                while (a < b) { const int m = (a + b) >> 1; if (blk[2 + m] < r) a = m + 1; else b = m; }
                if (a < c && blk[2 + a] == r) lab = before + a;
            }
            before += c;
        }
        labels[i - lo] = over ? -2 : lab;
    }
}

// ---------------------------------------------------------------- host side
struct rhccq_pt_ws {
    float* rec; int* cell_id; int* cell_start; int* cell_fill; int* parent; int* rootlab; int* is_root; int* scan;
    float* partial;
};
static size_t rhccq_al(size_t b) { return (b + 255) & ~(size_t)255; }
static size_t rhccq_pt_carve(const rhccq_dbscan_plan* P, unsigned char* base, rhccq_pt_ws* W) {
    size_t o = 0;
    const size_t n = (size_t)P->n, m = (size_t)P->n_cells + 2;
    auto take = [&](size_t bytes) { unsigned char* p = base ? base + o : nullptr; o += rhccq_al(bytes); return p; };
    float* rec = (float*)take(n * 32);
    int* cell_id = (int*)take(n * 4);
    int* cell_start = (int*)take(m * 4);
    int* cell_fill = (int*)take(m * 4);
    int* parent = (int*)take(n * 4);
    int* rootlab = (int*)take(n * 4);
    int* is_root = (int*)take(n * 4);
    const size_t big = m > n ? m : n;
    int* scan = (int*)take(rhccq_scan_scratch_ints((long long)big) * 4 + 64);
    float* partial = (float*)take(1024 * 6 * 4);
    if (W) { W->rec = rec; W->cell_id = cell_id; W->cell_start = cell_start; W->cell_fill = cell_fill; W->parent = parent;
             W->rootlab = rootlab; W->is_root = is_root; W->scan = scan; W->partial = partial; }
    return o;
}

static rhccq_pt_grid rhccq_pt_make_grid(const rhccq_dbscan_plan* P) {
    rhccq_pt_grid G;
    G.n = P->n; G.dims = P->dims; G.gd = P->grid_dims; G.min_pts = P->min_pts;
    G.nc0 = P->ncell[0]; G.nc1 = P->ncell[1]; G.nc2 = P->grid_dims > 2 ? P->ncell[2] : 1;
    G.cw = P->cells_per_tile; G.tiles_per_row = (G.nc0 + G.cw - 1) / G.cw; G.rows = G.nc1 * G.nc2;
    G.o0 = P->origin[0]; G.o1 = P->origin[1]; G.o2 = P->origin[2]; G.inv_side = 1.0 / P->side;
    G.r2 = P->eps * P->eps;
    G.r2f = (float)G.r2;
    G.r2lo = (float)(G.r2 * (1.0 - 1.0 / 262144.0));
    G.r2hi = (float)(G.r2 * (1.0 + 1.0 / 262144.0));
    return G;
}

static int rhccq_pt_blocks(long long n) {
    long long b = (n + RHCCQ_PT_THREADS - 1) / RHCCQ_PT_THREADS;
    const long long cap = (long long)rhccq_sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

int rhccq_dbscan_plan_make(int n, int dims, int grid_dims, double eps, int min_pts, const double* lo,
                           const double* hi, rhccq_dbscan_plan* P) {
    if (!P || !lo || !hi || n < 0 || dims < 2 || dims > 6 || grid_dims < 2 || grid_dims > 3 || grid_dims > dims ||
        !(eps > 0.0) || min_pts < 1) {
        rhccq_set_error("rhccq_dbscan_plan_make: need n >= 0, 2 <= dims <= 6, grid_dims 2 or 3 (<= dims), eps > 0, "
                        "min_pts >= 1");
        return -1;
    }
    memset(P, 0, sizeof *P);
    P->n = n; P->dims = dims; P->grid_dims = grid_dims; P->min_pts = min_pts; P->eps = eps;
    // cells a hair wider than eps: two points within eps are then at most one cell apart even after the
    // rounding of the cell computation (done in float64)
    P->side = eps * (1.0 + 1.0 / 1048576.0);
    long long cells = 1;
    for (int d = 0; d < 3; ++d) {
        P->origin[d] = d < grid_dims ? lo[d] : 0.0;
        long long nc = 1;
        if (d < grid_dims) {
            const double span = hi[d] - lo[d];
            if (!(span >= 0.0)) { rhccq_set_error("rhccq_dbscan_plan_make: empty or NaN bounds in dimension %d", d); return -1; }
            nc = (long long)floor(span / P->side) + 1;
        }
        if (nc > 2000000000LL) { rhccq_set_error("rhccq_dbscan_plan_make: too many cells in dimension %d", d); return -1; }
        P->ncell[d] = (int)nc;
        cells *= nc;
        if (cells > 1500000000LL) {
            rhccq_set_error("rhccq_dbscan_plan_make: %lld cells exceed the int32 cell index; use a larger eps or fewer grid dims", cells);
            return -1;
        }
    }
    P->n_cells = cells;
    const double ppc = cells > 0 ? (double)n / (double)cells : 1.0;
    long long cw = (long long)(RHCCQ_PT_THREADS / (ppc > 0.0625 ? ppc : 0.0625));
    if (cw < 1) cw = 1;
    if (cw > P->ncell[0]) cw = P->ncell[0];
    P->cells_per_tile = (int)cw;
    P->n_tiles = (long long)((P->ncell[0] + cw - 1) / cw) * P->ncell[1] * (grid_dims > 2 ? P->ncell[2] : 1);
    return 0;
}

size_t rhccq_dbscan_workspace_bytes(const rhccq_dbscan_plan* P) {
    return P ? rhccq_pt_carve(P, nullptr, nullptr) : 0;
}

int rhccq_dbscan_bounds(const float* pts, int n, int dims, int grid_dims, double* out6, void* ws, size_t ws_bytes,
                        void* stream) {
    if (n <= 0 || !pts || !out6 || !ws || ws_bytes < 1024 * 6 * 4) { rhccq_set_error("rhccq_dbscan_bounds: bad arguments"); return -1; }
    int nb = rhccq_pt_blocks(n);
    if (nb > 1024) nb = 1024;
    RHCCQ_LAUNCH(rhccq_k_pt_bounds, nb, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, pts, n, dims, grid_dims, (float*)ws);
    RHCCQ_LAUNCH(rhccq_k_pt_bounds_final, 1, 32, 0, (cudaStream_t)stream, (const float*)ws, nb, out6);
    return 0;
}

#define RHCCQ_PT_ARGS(name)                                                                        \
    if (!P || !ws || ws_bytes < rhccq_pt_carve(P, nullptr, nullptr)) {                             \
        rhccq_set_error(name ": plan or workspace missing / too small"); return -1; }              \
    rhccq_pt_ws W; rhccq_pt_carve(P, (unsigned char*)ws, &W);                                      \
    const rhccq_pt_grid G = rhccq_pt_make_grid(P);                                                 \
    if (P->n == 0) return 0;

int rhccq_dbscan_bin(const rhccq_dbscan_plan* P, const float* pts, void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_bin")
    const long long m = P->n_cells + 1;
#ifdef RHCCQ_HOST_EMU
    memset(W.cell_start, 0, (size_t)(m + 1) * 4); memset(W.cell_fill, 0, (size_t)(m + 1) * 4);
#else
    cudaMemsetAsync(W.cell_start, 0, (size_t)(m + 1) * 4, (cudaStream_t)stream);
    cudaMemsetAsync(W.cell_fill, 0, (size_t)(m + 1) * 4, (cudaStream_t)stream);
#endif
    RHCCQ_LAUNCH(rhccq_k_pt_cellid, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, pts, G, W.cell_id, W.cell_start);
    if (m <= RHCCQ_SCAN_TILE) {
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, W.cell_start, (int)m, (int*)nullptr);
    } else if (rhccq_scan_i32(W.cell_start, m, W.cell_start, W.scan, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_pt_scatter, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, pts, G, W.cell_id,
                 W.cell_start, W.cell_fill, W.rec);
    return 0;
}

static int rhccq_pt_sweep_launch(int mode, const rhccq_pt_grid& G, const rhccq_dbscan_plan* P, rhccq_pt_ws& W, uint8_t* core,
                                 void* stream) {
    const size_t smem = (size_t)2 * RHCCQ_PT_CAP * 32;
    long long tiles = P->n_tiles;
    const long long cap = (long long)rhccq_sm_count() * 3 * 8;      // 3 resident CTAs per SM, several waves
    const int grid = (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
    if (mode == 0) {
        if (rhccq_smem_optin((const void*)rhccq_k_pt_sweep<0>, smem) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_pt_sweep<0>, grid, RHCCQ_PT_THREADS, smem, (cudaStream_t)stream, G, W.rec, W.cell_start, core, W.parent, W.rootlab);
    } else if (mode == 1) {
        if (rhccq_smem_optin((const void*)rhccq_k_pt_sweep<1>, smem) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_pt_sweep<1>, grid, RHCCQ_PT_THREADS, smem, (cudaStream_t)stream, G, W.rec, W.cell_start, core, W.parent, W.rootlab);
    } else {
        if (rhccq_smem_optin((const void*)rhccq_k_pt_sweep<2>, smem) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_pt_sweep<2>, grid, RHCCQ_PT_THREADS, smem, (cudaStream_t)stream, G, W.rec, W.cell_start, core, W.parent, W.rootlab);
    }
    return 0;
}

int rhccq_dbscan_count(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, uint8_t* core, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_count")
    return rhccq_pt_sweep_launch(0, G, P, W, core, stream);
}

int rhccq_dbscan_union(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, uint8_t* core, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_union")
    RHCCQ_LAUNCH(rhccq_k_pt_init_parent, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, P->n, W.parent);
    return rhccq_pt_sweep_launch(1, G, P, W, core, stream);
}

int rhccq_dbscan_flatten(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, uint8_t* core, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_flatten")
    RHCCQ_LAUNCH(rhccq_k_pt_flatten, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, P->n, W.parent, core, W.rootlab, W.is_root);
    return 0;
}

int rhccq_dbscan_attach(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, uint8_t* core, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_attach")
    return rhccq_pt_sweep_launch(2, G, P, W, core, stream);
}

int rhccq_dbscan_border(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, uint8_t* core, void* stream) {
    if (rhccq_dbscan_flatten(P, ws, ws_bytes, core, stream) != 0) return -1;
    return rhccq_dbscan_attach(P, ws, ws_bytes, core, stream);
}

size_t rhccq_dbscan_ws_offset(const rhccq_dbscan_plan* P, int which) {
    if (!P) return 0;
    rhccq_pt_ws W;
    rhccq_pt_carve(P, (unsigned char*)256, &W);                   // offsets relative to a fake non-null base
    const unsigned char* p = which == 0 ? (unsigned char*)W.rootlab : which == 1 ? (unsigned char*)W.is_root
                           : which == 2 ? (unsigned char*)W.parent : (unsigned char*)W.rec;
    return (size_t)(p - (unsigned char*)256);
}

int rhccq_uf_emit_edges(const int32_t* rootlab, int lo, int hi, int g0, int32_t* edges, int32_t* counter, int capacity,
                        void* stream) {
    if (hi <= lo) return 0;
    RHCCQ_LAUNCH(rhccq_k_uf_emit, rhccq_pt_blocks(hi - lo), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, rootlab, lo, hi, g0,
                 edges, counter, capacity);
    return 0;
}

int rhccq_uf_merge_edges(const int32_t* edges, int n_edges, int32_t* table_keys, int32_t* table_parent, int table_cap,
                         void* stream) {
    if (table_cap < 2 || (table_cap & (table_cap - 1)) != 0 || (long long)table_cap < 4LL * n_edges) {
        rhccq_set_error("rhccq_uf_merge_edges: table capacity must be a power of two >= 4 * n_edges"); return -1;
    }
    RHCCQ_LAUNCH(rhccq_k_uf_tab_init, rhccq_pt_blocks(table_cap), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, table_keys, table_parent, table_cap);
    if (n_edges <= 0) return 0;
    RHCCQ_LAUNCH(rhccq_k_uf_tab_insert, rhccq_pt_blocks(2LL * n_edges), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, edges, n_edges, table_keys, table_cap - 1);
    RHCCQ_LAUNCH(rhccq_k_uf_tab_union, rhccq_pt_blocks(n_edges), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, edges, n_edges, table_keys, table_parent, table_cap - 1);
    return 0;
}

int rhccq_uf_lookup_roots(int32_t* rootlab, int n, int g0, const int32_t* table_keys, const int32_t* table_parent,
                          int table_cap, void* stream) {
    if (n <= 0) return 0;
    RHCCQ_LAUNCH(rhccq_k_uf_lookup, rhccq_pt_blocks(n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, rootlab, n, g0, table_keys,
                 table_parent, table_keys ? table_cap - 1 : -1);
    return 0;
}

int rhccq_uf_merge_edges_gathered(const int32_t* gathered, int world, int stride_ints, int cap_rows, int32_t* table_keys,
                                  int32_t* table_parent, int table_cap, void* stream) {
    if (!gathered || world < 1 || cap_rows < 1 || stride_ints < 2 + 2 * cap_rows || table_cap < 2 ||
        (table_cap & (table_cap - 1)) != 0 || (long long)table_cap < 4LL * world * cap_rows) {
        rhccq_set_error("rhccq_uf_merge_edges_gathered: need blocks of 2 + 2 cap_rows ints and a power-of-two table >= 4 world cap_rows");
        return -1;
    }
    RHCCQ_LAUNCH(rhccq_k_uf_tab_init, rhccq_pt_blocks(table_cap), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, table_keys, table_parent, table_cap);
    RHCCQ_LAUNCH(rhccq_k_uf_tab_insert_g, rhccq_pt_blocks(2LL * world * cap_rows), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream,
                 gathered, world, stride_ints, cap_rows, table_keys, table_cap - 1);
    RHCCQ_LAUNCH(rhccq_k_uf_tab_union_g, rhccq_pt_blocks((long long)world * cap_rows), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream,
                 gathered, world, stride_ints, cap_rows, table_keys, table_parent, table_cap - 1);
    return 0;
}

int rhccq_uf_rank_labels_gathered(const int32_t* gathered, int world, int stride_ints, int cap_rows, const int32_t* rootlab,
                                  int lo, int hi, int32_t* labels, void* stream) {
    if (!gathered || world < 1 || cap_rows < 1 || stride_ints < 2 + cap_rows) {
        rhccq_set_error("rhccq_uf_rank_labels_gathered: need blocks of 2 + cap_rows ints"); return -1;
    }
    if (hi <= lo) return 0;
    RHCCQ_LAUNCH(rhccq_k_uf_rank_g, rhccq_pt_blocks(hi - lo), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, gathered, world,
                 stride_ints, cap_rows, rootlab, lo, hi, labels);
    return 0;
}

size_t rhccq_uf_own_roots_scratch_ints(int n) { return (size_t)(n > 0 ? n : 1) + rhccq_scan_scratch_ints(n) + 16; }

int rhccq_uf_own_roots(const int32_t* rootlab, int n, int own_lo, int own_hi, int g0, int32_t* scratch, int32_t* out_ids,
                       int32_t* out_count, int out_capacity, void* stream) {
    if (!rootlab || !scratch || !out_ids || !out_count || n <= 0 || own_lo < 0 || own_hi > n || own_hi < own_lo) {
        rhccq_set_error("rhccq_uf_own_roots: bad arguments"); return -1;
    }
    int* flags = scratch;
    int* scan = scratch + n;
    RHCCQ_LAUNCH(rhccq_k_uf_root_flags, rhccq_pt_blocks(n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, rootlab, n, own_lo, own_hi, g0, flags);
    if (n <= RHCCQ_SCAN_TILE) {
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, flags, n, (int*)nullptr);
    } else if (rhccq_scan_i32(flags, n, flags, scan, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_uf_root_scatter, rhccq_pt_blocks(own_hi - own_lo), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, rootlab,
                 flags, own_lo, own_hi, g0, out_ids, out_count, n, out_capacity > 0 ? out_capacity : 0x7fffffff);
    return 0;
}

int rhccq_dbscan_own_roots(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, int own_lo, int own_hi, int g0,
                           int32_t* out_ids, int32_t* out_count, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_own_roots")
    RHCCQ_LAUNCH(rhccq_k_uf_root_flags, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, W.rootlab, P->n, own_lo, own_hi, g0, W.is_root);
    if (P->n <= RHCCQ_SCAN_TILE) {
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, W.is_root, P->n, (int*)nullptr);
    } else if (rhccq_scan_i32(W.is_root, P->n, W.is_root, W.scan, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_uf_root_scatter, rhccq_pt_blocks(own_hi - own_lo), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, W.rootlab,
                 W.is_root, own_lo, own_hi, g0, out_ids, out_count, P->n, 0x7fffffff);
    return 0;
}

int rhccq_uf_rank_labels(const int32_t* sorted_roots, int n_roots, const int32_t* rootlab, int lo, int hi, int32_t* labels,
                         void* stream) {
    if (hi <= lo) return 0;
    RHCCQ_LAUNCH(rhccq_k_uf_rank, rhccq_pt_blocks(hi - lo), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, sorted_roots, n_roots,
                 rootlab, lo, hi, labels);
    return 0;
}

int rhccq_dbscan_relabel(const rhccq_dbscan_plan* P, void* ws, size_t ws_bytes, int32_t* labels, void* stream) {
    RHCCQ_PT_ARGS("rhccq_dbscan_relabel")
    if (P->n <= RHCCQ_SCAN_TILE) {
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, W.is_root, P->n, (int*)nullptr);
    } else if (rhccq_scan_i32(W.is_root, P->n, W.is_root, W.scan, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_pt_labels, rhccq_pt_blocks(P->n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, P->n, W.rootlab, W.is_root, labels);
    return 0;
}

}  // extern "C"

// ================================================================= image lattice
// The same operator when the points ARE an image: X[i] = (x, y, R, G, B) of pixel i in raster order (the
// reference's pixel features; BASELINE.json configs 3-5).  Then the cell grid is the pixel lattice itself:
// two pixels are within eps iff (dx^2 + dy^2) + |dRGB|^2 <= eps^2, all integers, so a pixel's neighbours are
// found with a (2 floor(eps) + 1)^2 stencil and a colour budget per offset on packed 8-bit colours
// (two instructions per candidate), from a shared-memory tile with halo.  The count pass reads the 20-byte
// points once, coalesced, checks that they really are a lattice of 8-bit colours (status flag otherwise: the
// caller then takes the generic path), and writes the count and a packed colour (+ core bit) per pixel — its
// HBM traffic is the algorithmic 24 bytes per point plus 4.
#define RHCCQ_LT_W 64
#define RHCCQ_LT_MAXR 16
#define RHCCQ_LT_INVALID 0xFF000000u
#define RHCCQ_LT_CORE 0x01000000u

struct rhccq_lt_args {
    int H, W, R, thr, min_pts, n_off;
};

// stencil entries, raster order: x = (dy + 64) << 8 | (dx + 64), y = colour budget thr - (dy^2 + dx^2).
// Built by one thread (a few hundred entries at most).
__device__ __forceinline__ int rhccq_lt_build_offsets(int2* offs, int R, int thr, bool forward_only) {
    int n = 0;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
            const int s = dy * dy + dx * dx;
            if (s > thr) continue;
            if (forward_only && !(dy > 0 || (dy == 0 && dx > 0))) continue;
            offs[n].x = ((dy + 64) << 8) | (dx + 64);
            offs[n].y = thr - s;
            ++n;
        }
    return n;
}

// one float32 point -> packed colour, checking that it is pixel (x, y) of an 8-bit image
__device__ __forceinline__ uint32_t rhccq_lt_pack_point(float fx, float fy, float r, float g, float b, int x, int y, int* bad) {
    const int ir = (int)r, ig = (int)g, ib = (int)b;
    if (fx != (float)x || fy != (float)y || (float)ir != r || (float)ig != g || (float)ib != b ||
        (unsigned)ir > 255u || (unsigned)ig > 255u || (unsigned)ib > 255u) *bad = 1;
    return rhccq_pack_rgb((unsigned)ir & 255u, (unsigned)ig & 255u, (unsigned)ib & 255u);
}

// tile of packed colours (top byte: 0xFF outside the image, bit 24 = core) with a halo of R pixels
template <int SRC>   // 0: float32 points [H*W,5], 1: uint8 image [H,W,3], 2: packed uint32 [H*W]
__device__ __forceinline__ void rhccq_lt_load_tile(const void* src, const rhccq_lt_args& A, int ty0, int tx0, uint32_t* tile,
                                                   int tw, int th, int* bad, int pad = -1) {
    if (pad < 0) pad = A.R;                                        // columns of halo on each side (rows: always A.R)
#ifndef RHCCQ_HOST_EMU
    // float points, interior of the tile: four pixels = 80 bytes = five 16-byte loads, contiguous across the lanes
    const bool vec = SRC == 0 && (A.W & 3) == 0;
    if (vec) {
        const int own_h = th - 2 * A.R, quads = RHCCQ_LT_W / 4;
        RHCCQ_PAR_FOR(t, own_h * quads) {
            const int ly = t / quads, lx = (t % quads) * 4;
            const int y = ty0 + ly, x = tx0 + lx;
            if (y >= A.H || x >= A.W) continue;                        // W % 4 == 0: a quad is inside or outside as a whole
            const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + ((size_t)y * A.W + x) * 5);
            const float4 a = p[0], b = p[1], c = p[2], d = p[3], e = p[4];
            uint32_t* o = tile + (ly + A.R) * tw + lx + pad;
            o[0] = rhccq_lt_pack_point(a.x, a.y, a.z, a.w, b.x, x, y, bad);
            o[1] = rhccq_lt_pack_point(b.y, b.z, b.w, c.x, c.y, x + 1, y, bad);
            o[2] = rhccq_lt_pack_point(c.z, c.w, d.x, d.y, d.z, x + 2, y, bad);
            o[3] = rhccq_lt_pack_point(d.w, e.x, e.y, e.z, e.w, x + 3, y, bad);
        }
    }
#else
    const bool vec = false;
#endif
    RHCCQ_PAR_FOR(t, tw * th) {
        const int ty = t / tw, tx = t % tw;
        const int y = ty0 - A.R + ty, x = tx0 - pad + tx;
        const bool inside = y >= 0 && y < A.H && x >= 0 && x < A.W;
        if (vec && inside && ty >= A.R && ty < th - A.R && tx >= pad && tx < pad + RHCCQ_LT_W) continue;   // loaded above
        uint32_t v = RHCCQ_LT_INVALID;
        if (inside) {
            const size_t i = (size_t)y * A.W + x;
            if (SRC == 0) {
                const float* p = reinterpret_cast<const float*>(src) + i * 5;
                v = rhccq_lt_pack_point(p[0], p[1], p[2], p[3], p[4], x, y, bad);
            } else if (SRC == 1) {
                const uint8_t* p = reinterpret_cast<const uint8_t*>(src) + i * 3;
                v = rhccq_pack_rgb(p[0], p[1], p[2]);
            } else {
                v = reinterpret_cast<const uint32_t*>(src)[i];
            }
        }
        tile[t] = v;
    }
}

#define RHCCQ_LT_PPT 8        // pixels per thread of the count pass: rows ly, ly + TH/8, ...

// small shared-memory set of (set a, set b) pairs: true the first time a pair is offered (or when the
// neighbourhood of its slot is full: a repeated union is harmless)
#define RHCCQ_LT_PAIRS 1024
__device__ __forceinline__ bool rhccq_lt_pair_is_new(unsigned long long* pairs, int a, int b) {
    const unsigned long long key = ((unsigned long long)(unsigned)a << 32) | (unsigned)b;
    unsigned slot = ((unsigned)a * 2654435761u ^ (unsigned)b * 2246822519u) >> 22;       // 10 bits
    for (int probe = 0; probe < 16; ++probe) {
        const unsigned long long old = atomicCAS(&pairs[slot], ~0ull, key);
        if (old == ~0ull) return true;
        if (old == key) return false;
        slot = (slot + 1u) & (RHCCQ_LT_PAIRS - 1u);
    }
    return true;
}

// MODE 0: count (+ packed colours with the core bit)
//      3: union across tile borders (global union-find), after rhccq_k_lt_union_tile has finished every tile
//      2: border attachment
template <int MODE, int SRC, int TH>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_sweep(const void* __restrict__ src, rhccq_lt_args A, int* __restrict__ count, uint32_t* __restrict__ packed,
                 uint8_t* __restrict__ core, int* __restrict__ parent, int* __restrict__ rootlab, int* __restrict__ status) {
    RHCCQ_DYN_SMEM(dyn);
    __shared__ int s_noff, s_bad, s_cnt;
    const int tw = RHCCQ_LT_W + 2 * A.R, th = TH + 2 * A.R;
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn);
    int2* offs = reinterpret_cast<int2*>(tile + (((size_t)tw * th + 1) & ~(size_t)1));
    int* lpar = reinterpret_cast<int*>(offs + (2 * A.R + 1) * (2 * A.R + 1));      // RHCCQ_LT_W * TH ints: work list
    if (threadIdx.x == 0) { s_noff = rhccq_lt_build_offsets(offs, A.R, A.thr, MODE == 3); s_bad = 0; }
    __syncthreads();
    const int n_off = s_noff;
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH;
    for (int tI = blockIdx.x; tI < tiles_x * tiles_y; tI += gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        rhccq_lt_load_tile<SRC>(src, A, ty0, tx0, tile, tw, th, &s_bad);
        __syncthreads();
        if (MODE == 0) {
            // every thread owns RHCCQ_LT_PPT pixels of one column: a stencil entry is decoded once for all of them
            RHCCQ_PAR_FOR(t, RHCCQ_LT_W * (TH / RHCCQ_LT_PPT)) {
                const int ly = t / RHCCQ_LT_W, lx = t % RHCCQ_LT_W;
                int ci[RHCCQ_LT_PPT], acc[RHCCQ_LT_PPT];
                uint32_t c[RHCCQ_LT_PPT];
#pragma unroll
                for (int u = 0; u < RHCCQ_LT_PPT; ++u) {
                    ci[u] = (ly + u * (TH / RHCCQ_LT_PPT) + A.R) * tw + lx + A.R;
                    c[u] = tile[ci[u]] & 0x00ffffffu;
                    acc[u] = 0;
                }
                const uint32_t* base = tile + ci[0];
                const int rstep = (TH / RHCCQ_LT_PPT) * tw;
                for (int o = 0; o < n_off; ++o) {
                    const int2 e = offs[o];
                    const uint32_t* nbp = base + ((e.x >> 8) - 64) * tw + ((e.x & 255) - 64);
#pragma unroll
                    for (int u = 0; u < RHCCQ_LT_PPT; ++u)                 // 0xFF top byte (outside): never within
                        acc[u] += (int)((unsigned)rhccq_d2(c[u], nbp[u * rstep]) <= (unsigned)e.y);
                }
#pragma unroll
                for (int u = 0; u < RHCCQ_LT_PPT; ++u) {
                    const int y = ty0 + ly + u * (TH / RHCCQ_LT_PPT), x = tx0 + lx;
                    if (y >= A.H || x >= A.W) continue;
                    const int id = y * A.W + x;
                    const int is_core = acc[u] >= A.min_pts;
                    count[id] = acc[u];
                    core[id] = (uint8_t)is_core;
                    packed[id] = c[u] | (is_core ? RHCCQ_LT_CORE : 0u);
                }
            }
        } else {
            // the pixels with work to do, compacted so that every lane of the pass below is busy:
            // attachment (2): the non-core pixels; cross-border links (3): core pixels within R of the tile's
            // left, right or bottom edge (forward offsets never leave through the top)
            int* list = lpar;
            unsigned long long* pairs = reinterpret_cast<unsigned long long*>(lpar + RHCCQ_LT_W * TH);
            if (threadIdx.x == 0) s_cnt = 0;
            if (MODE == 3) RHCCQ_PAR_FOR(k, RHCCQ_LT_PAIRS) pairs[k] = ~0ull;
            __syncthreads();
            RHCCQ_PAR_FOR(t, RHCCQ_LT_W * TH) {                         // (trip count uniform: votes are safe)
                const int ly = t / RHCCQ_LT_W, lx = t % RHCCQ_LT_W;
                bool want = ty0 + ly < A.H && tx0 + lx < A.W;
                if (want) {
                    const bool is_core = (tile[(ly + A.R) * tw + lx + A.R] >> 24) == 1u;
                    want = MODE == 2 ? !is_core : (is_core && (lx < A.R || lx >= RHCCQ_LT_W - A.R || ly >= TH - A.R));
                }
                const unsigned m = rhccq_ballot(want);
                int base = 0;
                if (RHCCQ_LANE == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
                base = rhccq_shfl(base, 0);
                if (want) list[base + __popc(m & rhccq_lanemask_lt())] = t;
            }
            __syncthreads();
            const int n_list = s_cnt;
            RHCCQ_PAR_FOR(li, n_list) {
                const int t = list[li];
                const int ly = t / RHCCQ_LT_W, lx = t % RHCCQ_LT_W;
                const int ci = (ly + A.R) * tw + lx + A.R;
                const uint32_t c = tile[ci] & 0x00ffffffu;
                const int id = (ty0 + ly) * A.W + tx0 + lx;
                int best = 0x7fffffff;
                const int my_root = MODE == 3 ? ((volatile int*)parent)[id] : id;     // (3) a star from the tile pass: its set
                for (int o = 0; o < n_off; ++o) {
                    const int2 e = offs[o];
                    const int dy = (e.x >> 8) - 64, dx = (e.x & 255) - 64;
                    if (MODE == 3 && (unsigned)(ly + dy) < (unsigned)TH && (unsigned)(lx + dx) < (unsigned)RHCCQ_LT_W) continue;   // linked inside the tile
                    const uint32_t nb = tile[ci + dy * tw + dx];
                    if ((nb >> 24) != 1u) continue;                     // outside the image, or not core
                    if ((unsigned)rhccq_d2(c, nb & 0x00ffffffu) > (unsigned)e.y) continue;
                    if (MODE == 3) {
                        // many pixel pairs bridge the same two tile-local sets: only the first of a
                        // (my set, neighbour's set) pair seen by this block goes to the global forest
                        const int rb = ((volatile int*)parent)[id + dy * A.W + dx];
                        if (rb != my_root && rhccq_lt_pair_is_new(pairs, my_root, rb)) rhccq_pt_union(parent, my_root, rb);
                    } else {
                        const int r = rootlab[id + dy * A.W + dx];
                        best = r < best ? r : best;
                    }
                }
                if (MODE == 2) rootlab[id] = best == 0x7fffffff ? -1 : best;
            }
        }
        __syncthreads();
    }
    if (MODE == 0 && SRC == 0 && threadIdx.x == 0 && s_bad) *status = 1;
}

// Union inside a tile (what MODE 1 above did, restructured so that the shared-memory forest stays shallow):
//  1. horizontal runs — core pixels linked to their left neighbour — are found with warp votes and start as
//     stars under the run's first pixel, with no union at all;
//  2. the remaining forward offsets are taken one at a time, all pixels in step. Every tree is a star when a
//     pass starts, so "already linked" is one comparison of two shared-memory loads, and only pairs that really
//     bridge two sets walk and link trees; the trees are flattened to stars again after each pass.
// The result goes to `parent` as stars in global indices (a set's root is its lowest index, as before).
template <int TH>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_union_tile(const uint32_t* __restrict__ packed, rhccq_lt_args A, int* __restrict__ parent) {
    RHCCQ_DYN_SMEM(dyn);
    __shared__ int s_noff, s_bad;
    const int tw = RHCCQ_LT_W + 2 * A.R, th = TH + 2 * A.R;
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn);
    int2* offs = reinterpret_cast<int2*>(tile + (((size_t)tw * th + 1) & ~(size_t)1));
    int* lpar = reinterpret_cast<int*>(offs + (2 * A.R + 1) * (2 * A.R + 1));
    if (threadIdx.x == 0) { s_noff = rhccq_lt_build_offsets(offs, A.R, A.thr, true); s_bad = 0; }
    __syncthreads();
    const int n_off = s_noff;
    const int run_budget = A.thr - 1;                                   // offset (0, 1)
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH;
    for (int tI = blockIdx.x; tI < tiles_x * tiles_y; tI += gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        rhccq_lt_load_tile<2>(packed, A, ty0, tx0, tile, tw, th, &s_bad);
        __syncthreads();
        // 1. horizontal runs
        for (int row = RHCCQ_WARP; row < TH; row += RHCCQ_NWARPS) {
            const uint32_t* trow = tile + (row + A.R) * tw + A.R;
            unsigned long long linked = 0ull;                           // bit x: pixel x continues the run of x - 1
            for (int h = 0; h < RHCCQ_LT_W / RHCCQ_WARP_SIZE; ++h) {
                const int lx = h * RHCCQ_WARP_SIZE + RHCCQ_LANE;
                bool conn = false;
                if (lx > 0) {
                    const uint32_t me = trow[lx], lf = trow[lx - 1];
                    conn = (me >> 24) == 1u && (lf >> 24) == 1u && rhccq_d2(me & 0x00ffffffu, lf & 0x00ffffffu) <= run_budget;
                }
                linked |= (unsigned long long)rhccq_ballot(conn) << (h * RHCCQ_WARP_SIZE);
            }
            for (int h = 0; h < RHCCQ_LT_W / RHCCQ_WARP_SIZE; ++h) {
                const int lx = h * RHCCQ_WARP_SIZE + RHCCQ_LANE;
                const unsigned long long starts = ~linked & ((2ull << lx) - 1ull);      // bit 0 is always a start
                lpar[row * RHCCQ_LT_W + lx] = row * RHCCQ_LT_W + 63 - __clzll((long long)starts);
            }
        }
        __syncthreads();
        // 2. one forward offset at a time, every tree a star when the pass starts
        for (int o = 0; o < n_off; ++o) {
            const int2 e = offs[o];
            const int dy = (e.x >> 8) - 64, dx = (e.x & 255) - 64;
            if (dy == 0 && dx == 1) continue;                           // the runs
            RHCCQ_PAR_FOR(t, RHCCQ_LT_W * TH) {
                const int ly = t / RHCCQ_LT_W, lx = t % RHCCQ_LT_W;
                if (!((unsigned)(ly + dy) < (unsigned)TH && (unsigned)(lx + dx) < (unsigned)RHCCQ_LT_W)) continue;   // cross-border pass
                const int ci = (ly + A.R) * tw + lx + A.R;
                const uint32_t me = tile[ci], nb = tile[ci + dy * tw + dx];
                if ((me >> 24) != 1u || (nb >> 24) != 1u) continue;     // core pixels of the image only
                if ((unsigned)rhccq_d2(me & 0x00ffffffu, nb & 0x00ffffffu) > (unsigned)e.y) continue;
                const int ra = ((volatile int*)lpar)[t], rb = ((volatile int*)lpar)[t + dy * RHCCQ_LT_W + dx];
                if (ra != rb) rhccq_pt_union(lpar, ra, rb);
            }
            __syncthreads();
            RHCCQ_PAR_FOR(t, RHCCQ_LT_W * TH) {
                const int r = rhccq_pt_find_ro(lpar, t);
                lpar[t] = r;                                            // racing readers see the old parent or the root: both ancestors
            }
            __syncthreads();
        }
        RHCCQ_PAR_FOR(t, RHCCQ_LT_W * TH) {
            const int ly = t / RHCCQ_LT_W, lx = t % RHCCQ_LT_W;
            const int y = ty0 + ly, x = tx0 + lx;
            if (y >= A.H || x >= A.W) continue;
            const int r = lpar[t];
            parent[y * A.W + x] = (ty0 + r / RHCCQ_LT_W) * A.W + tx0 + r % RHCCQ_LT_W;
        }
        __syncthreads();
    }
}

// Count pass for small radii (R <= 4, the usual eps 1..4.9): every thread owns 8 horizontally adjacent pixels.
// The tile has RHCCQ_LTT_PAD halo columns on each side whatever the radius, so that a thread's window of a
// stencil row (its 8 pixels + 4 either side) is four aligned 16-byte shared-memory loads; the dx offsets slide
// over that register window. A candidate costs three instructions: byte-wise |difference|, a dot product that
// starts from -(budget + 1), and the sign bit of the result added to the count.
#define RHCCQ_LTT_PAD 4
#define RHCCQ_LTT_TW (RHCCQ_LT_W + 2 * RHCCQ_LTT_PAD)
#define RHCCQ_LTT_TH 32

template <int RT, int THR = -1>       // THR >= 0: the threshold is a compile-time constant (no per-offset branches)
__device__ __forceinline__ void rhccq_lt_stencil8(const uint32_t* tile, int ly, int lx0, int thr_rt, uint32_t (&c)[8], int (&acc)[8]) {
    const int TW = RHCCQ_LTT_TW;
    const int thr = THR >= 0 ? THR : thr_rt;
    {
        const uint4* ctr = reinterpret_cast<const uint4*>(tile + (ly + RT) * TW + lx0 + RHCCQ_LTT_PAD);
        const uint4 a = ctr[0], b = ctr[1];
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 1;                            // the pixel itself
#pragma unroll
    for (int dy = -RT; dy <= RT; ++dy) {
        if (dy * dy > thr) continue;                                   // (block-uniform)
        uint32_t win[16];
        const uint4* row = reinterpret_cast<const uint4*>(tile + (ly + RT + dy) * TW + lx0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const uint4 v = row[j]; win[4 * j] = v.x; win[4 * j + 1] = v.y; win[4 * j + 2] = v.z; win[4 * j + 3] = v.w; }
#pragma unroll
        for (int dx = -RT; dx <= RT; ++dx) {
            if (dy == 0 && dx == 0) continue;
            const int budget = thr - dy * dy - dx * dx;
            if (budget < 0) continue;
            const unsigned start = (unsigned)(-(budget + 1));
#pragma unroll
            for (int u = 0; u < 8; ++u) {                              // 0xFF top byte (outside the image): never within
                const unsigned d = __vabsdiffu4(c[u], win[u + dx + RHCCQ_LTT_PAD]);
                acc[u] += (int)(__dp4a(d, d, start) >> 31);            // d2 - budget - 1 < 0  <=>  d2 <= budget
            }
        }
    }
}

// results of 8 adjacent pixels: counts, packed colours with the core bit, core flags
__device__ __forceinline__ void rhccq_lt_store8(const rhccq_lt_args& A, int y, int x0, const uint32_t (&c)[8], const int (&acc)[8],
                                                int* __restrict__ count, uint32_t* __restrict__ packed, uint8_t* __restrict__ core) {
    if (y >= A.H || x0 >= A.W) return;
    uint32_t pk[8];
    unsigned long long cbits = 0ull;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int is_core = acc[u] >= A.min_pts;
        pk[u] = c[u] | (is_core ? RHCCQ_LT_CORE : 0u);
        cbits |= (unsigned long long)is_core << (8 * u);
    }
    const size_t id0 = (size_t)y * A.W + x0;
#ifndef RHCCQ_HOST_EMU
    if ((A.W & 7) == 0 && x0 + 8 <= A.W) {                             // 8 adjacent pixels: 32-byte aligned vector stores
        reinterpret_cast<int4*>(count + id0)[0] = make_int4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<int4*>(count + id0)[1] = make_int4(acc[4], acc[5], acc[6], acc[7]);
        reinterpret_cast<uint4*>(packed + id0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(packed + id0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        *reinterpret_cast<unsigned long long*>(core + id0) = cbits;
        return;
    }
#endif
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (x0 + u >= A.W) break;
        count[id0 + u] = acc[u];
        core[id0 + u] = (uint8_t)((cbits >> (8 * u)) & 1ull);
        packed[id0 + u] = pk[u];
    }
}

template <int SRC, int RT>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_count_rows(const void* __restrict__ src, rhccq_lt_args A, int* __restrict__ count, uint32_t* __restrict__ packed,
                      uint8_t* __restrict__ core, int* __restrict__ status) {
    RHCCQ_DYN_SMEM(dyn);
    __shared__ int s_bad;
    const int TH = RHCCQ_LTT_TH, SW = 8;
    const int tw = RHCCQ_LTT_TW, th = TH + 2 * RT;
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn);
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH;
    for (int tI = blockIdx.x; tI < tiles_x * tiles_y; tI += gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        rhccq_lt_load_tile<SRC>(src, A, ty0, tx0, tile, tw, th, &s_bad, RHCCQ_LTT_PAD);
        __syncthreads();
        RHCCQ_PAR_FOR(t, (RHCCQ_LT_W / SW) * TH) {
            const int ly = t / (RHCCQ_LT_W / SW), lx0 = (t % (RHCCQ_LT_W / SW)) * SW;
            uint32_t c[SW];
            int acc[SW];
            rhccq_lt_stencil8<RT>(tile, ly, lx0, A.thr, c, acc);
            rhccq_lt_store8(A, ty0 + ly, tx0 + lx0, c, acc, count, packed, core);
        }
        __syncthreads();
    }
    if (SRC == 0 && threadIdx.x == 0 && s_bad) *status = 1;
}

// The same tile-local union for small radii (R <= 4): a thread owns 8 adjacent pixels and looks UP and LEFT
// (the mirror image of the forward offsets: same edges). Tile words are re-encoded so that one byte-wise distance
// covers the core tests as well: core pixels carry top byte 0x00, everything else 0xFF, and a non-core centre
// 0x7F — any pair that is not core/core is at least 127^2 apart. One offset per pass, all pixels in step: the
// pairs within the budget are collected in a bit mask, and only those compare (and, when different, link) roots.
template <int RT>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_union_rows(const uint32_t* __restrict__ packed, rhccq_lt_args A, int* __restrict__ parent) {
    RHCCQ_DYN_SMEM(dyn);
    __shared__ int s_linked;
    constexpr int TH = RHCCQ_LTT_TH, TW = RHCCQ_LTT_TW, PAD = RHCCQ_LTT_PAD, TR = TH + RT, NTASK = TH * (RHCCQ_LT_W / 8);
    if (threadIdx.x == 0) s_linked = 0;
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn);                  // [TR][TW], RT rows of nothing on top
    int* lpar = reinterpret_cast<int*>(tile + TR * TW);                 // [TR][TW], parents as indices of this layout
    const int thr = A.thr;
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH;
    for (int tI = blockIdx.x; tI < tiles_x * tiles_y; tI += gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        // tile: own pixels re-encoded, the frame invalid; parents: everything its own root
        RHCCQ_PAR_FOR(t, TR * (TW / 4)) {
            const int r = t / (TW / 4), q = t % (TW / 4);
            const int y = ty0 + r - RT, x = tx0 + 4 * q - PAD;
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t v = RHCCQ_LT_INVALID;
                if (r >= RT && q >= 1 && q <= RHCCQ_LT_W / 4 && y < A.H && x + j < A.W) {
                    const uint32_t pk = packed[(size_t)y * A.W + x + j];
                    v = (pk >> 24) == 1u ? (pk & 0x00ffffffu) : (RHCCQ_LT_INVALID | (pk & 0x00ffffffu));
                }
                w[j] = v;
            }
            const int i0 = r * TW + 4 * q;
            *reinterpret_cast<uint4*>(tile + i0) = uint4{w[0], w[1], w[2], w[3]};
            *reinterpret_cast<uint4*>(lpar + i0) = uint4{(unsigned)i0, (unsigned)i0 + 1u, (unsigned)i0 + 2u, (unsigned)i0 + 3u};
        }
        __syncthreads();
        // 1. horizontal runs
        for (int row = RHCCQ_WARP; row < TH; row += RHCCQ_NWARPS) {
            const int rb = (row + RT) * TW + PAD;
            unsigned long long linked = 0ull;                           // bit x: pixel x continues the run of x - 1
            for (int h = 0; h < RHCCQ_LT_W / RHCCQ_WARP_SIZE; ++h) {
                const int lx = h * RHCCQ_WARP_SIZE + RHCCQ_LANE;
                const uint32_t me = tile[rb + lx], lf = tile[rb + lx - 1];              // (lx == 0: the frame, never linked)
                const bool conn = (me >> 24) == 0u && (int)__dp4a(__vabsdiffu4(me, lf), __vabsdiffu4(me, lf), (unsigned)(-thr)) < 0;
                linked |= (unsigned long long)rhccq_ballot(conn) << (h * RHCCQ_WARP_SIZE);
            }
            for (int h = 0; h < RHCCQ_LT_W / RHCCQ_WARP_SIZE; ++h) {
                const int lx = h * RHCCQ_WARP_SIZE + RHCCQ_LANE;
                const unsigned long long starts = ~linked & ((2ull << lx) - 1ull);      // bit 0 is always a start
                lpar[rb + lx] = rb + 63 - __clzll((long long)starts);
            }
        }
        __syncthreads();
        // 2. one offset at a time, all pixels in step; the forest is flattened to stars after a pass that linked
        for (int dy = 0; dy <= RT; ++dy)
            for (int dx = -RT; dx <= RT; ++dx) {
                if (dy == 0 && dx >= -1) continue;                      // same row: the left side only, and (0, -1) are the runs
                const int budget = thr - dy * dy - dx * dx;
                if (budget < 0) continue;                               // (block-uniform)
                const unsigned start = (unsigned)(-(budget + 1));
                RHCCQ_PAR_FOR(task, NTASK) {
                    const int ly = task / (RHCCQ_LT_W / 8), lx0 = (task % (RHCCQ_LT_W / 8)) * 8;
                    const int mi = (ly + RT) * TW + lx0 + PAD, ni = mi - dy * TW + dx;       // my pixels, their neighbours
                    uint32_t c[8];
                    {
                        const uint4 a = *reinterpret_cast<const uint4*>(tile + mi), b = *reinterpret_cast<const uint4*>(tile + mi + 4);
                        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
                    }
                    unsigned linked = 0u;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t cu = (c[u] >> 24) ? (c[u] ^ 0x80000000u) : c[u];      // non-core centre: 0x7F, far from both
                        const unsigned d = __vabsdiffu4(cu, tile[ni + u]);
                        linked |= (__dp4a(d, d, start) >> 31) << u;
                    }
                    bool any = false;
                    while (linked) {
                        const int u = __ffs((int)linked) - 1;
                        linked &= linked - 1u;
                        const int ra = ((volatile int*)lpar)[mi + u], rb = ((volatile int*)lpar)[ni + u];
                        if (ra != rb) { rhccq_pt_union(lpar, ra, rb); any = true; }
                    }
                    if (any) s_linked = 1;
                }
                __syncthreads();
                const int flat = s_linked;
                __syncthreads();
                if (!flat) continue;
                if (threadIdx.x == 0) s_linked = 0;
                RHCCQ_PAR_FOR(task, NTASK) {
                    const int ly = task / (RHCCQ_LT_W / 8), lx0 = (task % (RHCCQ_LT_W / 8)) * 8;
                    int* mine = lpar + (ly + RT) * TW + lx0 + PAD;
                    // (no link is made during this pass: plain loads, the eight walks side by side)
                    int p0[8], g[8];
                    {
                        const uint4 a = *reinterpret_cast<const uint4*>(mine), b = *reinterpret_cast<const uint4*>(mine + 4);
                        p0[0] = (int)a.x; p0[1] = (int)a.y; p0[2] = (int)a.z; p0[3] = (int)a.w;
                        p0[4] = (int)b.x; p0[5] = (int)b.y; p0[6] = (int)b.z; p0[7] = (int)b.w;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) g[u] = lpar[p0[u]];
                    bool deep = false;
#pragma unroll
                    for (int u = 0; u < 8; ++u) deep |= g[u] != p0[u];
                    if (deep) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (g[u] == p0[u]) continue;
                            const int r = rhccq_pt_find_ro(lpar, g[u]);
                            mine[u] = r;                                // (non-roots only, with an ancestor)
                        }
                    }
                }
                __syncthreads();
            }
        RHCCQ_PAR_FOR(task, NTASK) {
            const int ly = task / (RHCCQ_LT_W / 8), lx0 = (task % (RHCCQ_LT_W / 8)) * 8;
            const int y = ty0 + ly, x0 = tx0 + lx0;
            if (y >= A.H) continue;
            const int* mine = lpar + (ly + RT) * TW + lx0 + PAD;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (x0 + u >= A.W) break;
                const int r = rhccq_pt_find_ro(lpar, mine[u]);
                parent[(size_t)y * A.W + x0 + u] = (ty0 + r / TW - RT) * A.W + tx0 + r % TW - PAD;
            }
        }
        __syncthreads();
    }
}

// Border attachment for small radii (R <= 4), in the sliding-window form of the count pass: a thread owns 8 adjacent
// pixels; the tile holds the colours re-encoded as in the union pass (core: top byte 0x00, everything else 0xFF, so
// one byte-wise distance covers "is a core pixel" as well) and, beside it, the root labels of the same pixels.  A
// non-core pixel takes the smallest root label among the core pixels within eps (what rhccq_k_lt_sweep<2> computes
// over a compacted list with one global load per qualifying neighbour).
template <int RT>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_attach_rows(const uint32_t* __restrict__ packed, rhccq_lt_args A, int* __restrict__ rootlab) {
    RHCCQ_DYN_SMEM(dyn);
    constexpr int TH = RHCCQ_LTT_TH, TW = RHCCQ_LTT_TW, PAD = RHCCQ_LTT_PAD, THH = TH + 2 * RT, NTASK = TH * (RHCCQ_LT_W / 8);
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn);                  // [THH][TW]
    int* rl = reinterpret_cast<int*>(tile + THH * TW);                  // [THH][TW] root labels
    const int thr = A.thr;
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH;
    for (int tI = blockIdx.x; tI < tiles_x * tiles_y; tI += gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        RHCCQ_PAR_FOR(t, THH * (TW / 4)) {
            const int r = t / (TW / 4), q = t % (TW / 4);
            const int y = ty0 + r - RT, x = tx0 + 4 * q - PAD;
            uint32_t w[4];
            int l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t v = RHCCQ_LT_INVALID;
                int lab = 0x7fffffff;
                if (y >= 0 && y < A.H && x + j >= 0 && x + j < A.W) {
                    const size_t id = (size_t)y * A.W + x + j;
                    const uint32_t pk = packed[id];
                    const bool is_core = (pk >> 24) == 1u;
                    v = is_core ? (pk & 0x00ffffffu) : (0xFE000000u | (pk & 0x00ffffffu));     // 0xFE: a pixel of the image, not core
                    if (is_core) lab = rootlab[id];
                }
                w[j] = v; l[j] = lab;
            }
            const int i0 = r * TW + 4 * q;
            *reinterpret_cast<uint4*>(tile + i0) = uint4{w[0], w[1], w[2], w[3]};
            *reinterpret_cast<uint4*>(rl + i0) = uint4{(unsigned)l[0], (unsigned)l[1], (unsigned)l[2], (unsigned)l[3]};
        }
        __syncthreads();
        RHCCQ_PAR_FOR(task, NTASK) {
            const int ly = task / (RHCCQ_LT_W / 8), lx0 = (task % (RHCCQ_LT_W / 8)) * 8;
            const int y = ty0 + ly, x0 = tx0 + lx0;
            uint32_t c[8];
            {
                const uint4* ctr = reinterpret_cast<const uint4*>(tile + (ly + RT) * TW + lx0 + PAD);
                const uint4 a = ctr[0], b = ctr[1];
                c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
            }
            unsigned want = 0u;                                         // the non-core pixels of the image among my eight
#pragma unroll
            for (int u = 0; u < 8; ++u) want |= ((c[u] >> 24) == 0xFEu ? 1u : 0u) << u;
            if (!want || y >= A.H) continue;
            int best[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { best[u] = 0x7fffffff; c[u] &= 0x00ffffffu; }   // as a core pixel's word: top byte 0
#pragma unroll
            for (int dy = -RT; dy <= RT; ++dy) {
                if (dy * dy > thr) continue;                            // (block-uniform)
                uint32_t win[16];
                int wl[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 v = *reinterpret_cast<const uint4*>(tile + (ly + RT + dy) * TW + lx0 + 4 * j);
                    win[4 * j] = v.x; win[4 * j + 1] = v.y; win[4 * j + 2] = v.z; win[4 * j + 3] = v.w;
                    const uint4 r = *reinterpret_cast<const uint4*>(rl + (ly + RT + dy) * TW + lx0 + 4 * j);
                    wl[4 * j] = (int)r.x; wl[4 * j + 1] = (int)r.y; wl[4 * j + 2] = (int)r.z; wl[4 * j + 3] = (int)r.w;
                }
#pragma unroll
                for (int dx = -RT; dx <= RT; ++dx) {
                    if (dy == 0 && dx == 0) continue;
                    const int budget = thr - dy * dy - dx * dx;
                    if (budget < 0) continue;
                    const unsigned start = (unsigned)(-(budget + 1));
#pragma unroll
                    for (int u = 0; u < 8; ++u) {                       // a neighbour that is not core is at least 254^2 away
                        const unsigned d = __vabsdiffu4(c[u], win[u + dx + PAD]);
                        const int cand = (int)__dp4a(d, d, start) < 0 ? wl[u + dx + PAD] : 0x7fffffff;
                        best[u] = cand < best[u] ? cand : best[u];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (((want >> u) & 1u) && x0 + u < A.W) rootlab[(size_t)y * A.W + x0 + u] = best[u] == 0x7fffffff ? -1 : best[u];
        }
        __syncthreads();
    }
}

#ifndef RHCCQ_HOST_EMU
// One colour channel of a float32 point -> 0..255 without the conversion unit: v + 2^23 holds round(v) in its low
// mantissa bits. `eor` collects the bits of every "must be exactly zero" difference, `uor` every channel value.
__device__ __forceinline__ uint32_t rhccq_lt_chan(float v, uint32_t& eor, uint32_t& uor) {
    const float s = v + 8388608.0f;
    const uint32_t sb = __float_as_uint(s);
    eor |= __float_as_uint(s - 8388608.0f) ^ __float_as_uint(v);      // v is an integer (bit pattern of its rounding)
    uor |= sb ^ 0x4B000000u;                                           // ... of 0..255 (nothing above the low byte)
    return sb & 255u;
}
__device__ __forceinline__ uint32_t rhccq_lt_pack_fast(float fx, float fy, float r, float g, float b, float xf, float yf,
                                                       uint32_t& eor, uint32_t& uor) {
    eor |= (__float_as_uint(fx) ^ __float_as_uint(xf)) | (__float_as_uint(fy) ^ __float_as_uint(yf));
    const uint32_t ur = rhccq_lt_chan(r, eor, uor), ug = rhccq_lt_chan(g, eor, uor), ub = rhccq_lt_chan(b, eor, uor);
    return (ur << 16) + (ug << 8) + ub;
}

// The same count pass for float32 points, as a persistent kernel fed by TMA: the rows of a tile (halo included)
// are contiguous runs of 20-byte points in global memory, so one cp.async.bulk per row brings the raw floats into
// a staging buffer; the copies of the next tile are in flight while the stencil of the current one runs.
template <int RT, int THR>
__global__ void __launch_bounds__(RHCCQ_PT_THREADS, 3)
rhccq_k_lt_count_tma(const float* __restrict__ pts, rhccq_lt_args A, int* __restrict__ count, uint32_t* __restrict__ packed,
                     uint8_t* __restrict__ core, int* __restrict__ status) {
    RHCCQ_DYN_SMEM(dyn);
    __shared__ uint64_t s_bar;
    __shared__ int s_bad;
    constexpr int TH = RHCCQ_LTT_TH, TW = RHCCQ_LTT_TW, THH = TH + 2 * RT, QUADS = TW / 4;
    float* stage = reinterpret_cast<float*>(dyn);                                   // [THH][TW] points of 5 floats
    uint32_t* tile = reinterpret_cast<uint32_t*>(dyn + (size_t)THH * TW * 20);      // [THH][TW] packed colours
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W, tiles_y = (A.H + TH - 1) / TH, ntiles = tiles_x * tiles_y;
    if (tid == 0) {
        s_bad = 0;
        rhccq_mbar_init(&s_bar);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tI) {                                                      // warp 0: one bulk copy per tile row
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        const int y_lo = max(ty0 - RT, 0), y_hi = min(ty0 + TH + RT, A.H);
        const int x_lo = max(tx0 - RHCCQ_LTT_PAD, 0), x_hi = min(tx0 + RHCCQ_LT_W + RHCCQ_LTT_PAD, A.W);
        const uint32_t row_bytes = (uint32_t)(x_hi - x_lo) * 20u;                   // W % 4 == 0: a multiple of 16
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rhccq_smem_addr(&s_bar)), "r"(row_bytes * (uint32_t)(y_hi - y_lo)) : "memory");
        __syncwarp();
        for (int y = y_lo + lane; y < y_hi; y += 32) {
            float* dst = stage + ((size_t)(y - (ty0 - RT)) * TW + (x_lo - (tx0 - RHCCQ_LTT_PAD))) * 5;
            const float* srcp = pts + ((size_t)y * A.W + x_lo) * 5;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(rhccq_smem_addr(dst)), "l"(srcp), "r"(row_bytes), "r"(rhccq_smem_addr(&s_bar)) : "memory");
        }
    };
    uint32_t phase = 0, eor = 0, uor = 0;
    if (warp == 0 && (int)blockIdx.x < ntiles) issue((int)blockIdx.x);
    for (int tI = (int)blockIdx.x; tI < ntiles; tI += (int)gridDim.x) {
        const int ty0 = (tI / tiles_x) * TH, tx0 = (tI % tiles_x) * RHCCQ_LT_W;
        rhccq_mbar_wait(&s_bar, phase);
        phase ^= 1u;
        for (int t = tid; t < THH * QUADS; t += RHCCQ_PT_THREADS) {                 // raw floats -> packed colours
            const int ty = t / QUADS, q = t % QUADS;
            const int y = ty0 - RT + ty, x = tx0 - RHCCQ_LTT_PAD + 4 * q;
            uint4 o = make_uint4(RHCCQ_LT_INVALID, RHCCQ_LT_INVALID, RHCCQ_LT_INVALID, RHCCQ_LT_INVALID);
            if ((unsigned)y < (unsigned)A.H && (unsigned)x < (unsigned)A.W) {
                const float4* p = reinterpret_cast<const float4*>(stage + ((size_t)ty * TW + 4 * q) * 5);
                const float4 a = p[0], b = p[1], c4 = p[2], d = p[3], e = p[4];
                const float xf = (float)x, yf = (float)y;
                o.x = rhccq_lt_pack_fast(a.x, a.y, a.z, a.w, b.x, xf, yf, eor, uor);
                o.y = rhccq_lt_pack_fast(b.y, b.z, b.w, c4.x, c4.y, xf + 1.0f, yf, eor, uor);
                o.z = rhccq_lt_pack_fast(c4.z, c4.w, d.x, d.y, d.z, xf + 2.0f, yf, eor, uor);
                o.w = rhccq_lt_pack_fast(d.w, e.x, e.y, e.z, e.w, xf + 3.0f, yf, eor, uor);
            }
            *reinterpret_cast<uint4*>(tile + ty * TW + 4 * q) = o;
        }
        __syncthreads();                                                            // tile complete, staging buffer free
        const int next = tI + (int)gridDim.x;
        if (warp == 0 && next < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(next);
        }
        {
            const int ly = tid / (RHCCQ_LT_W / 8), lx0 = (tid % (RHCCQ_LT_W / 8)) * 8;   // 256 threads = 32 rows x 8 octets
            uint32_t c[8];
            int acc[8];
            rhccq_lt_stencil8<RT, THR>(tile, ly, lx0, A.thr, c, acc);
            rhccq_lt_store8(A, ty0 + ly, tx0 + lx0, c, acc, count, packed, core);
        }
        __syncthreads();                                                            // before the next conversion overwrites the tile
    }
    if ((eor | (uor >> 8)) != 0u) s_bad = 1;            // (a -0.0 is flagged too: the caller then takes the generic path, which is always right)
    __syncthreads();
    if (tid == 0 && s_bad) *status = 1;
}
#endif

// relabel without a scan over every pixel: root flags -> bit mask + word populations; label = rank of the root
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_root_bits(const int* __restrict__ is_root, long long n, uint32_t* __restrict__ bits, int* __restrict__ wcount) {
#ifdef RHCCQ_HOST_EMU
    for (long long w = (long long)blockIdx.x; w < (n + 31) / 32; w += (long long)gridDim.x) {
        uint32_t word = 0u;
        for (int b = 0; b < 32; ++b) if (32 * w + b < n && is_root[32 * w + b]) word |= 1u << b;
        bits[w] = word; wcount[w] = __builtin_popcount(word);
    }
#else
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {      // (block-uniform trip count)
        const long long i = i0 + threadIdx.x;
        const unsigned word = __ballot_sync(0xffffffffu, i < n && is_root[i] != 0);
        if ((threadIdx.x & 31) == 0 && i < n) { bits[i >> 5] = word; wcount[i >> 5] = __popc(word); }
    }
#endif
}
__global__ void __launch_bounds__(RHCCQ_PT_THREADS)
rhccq_k_lt_labels_bits(long long n, const int* __restrict__ rootlab, const uint32_t* __restrict__ bits,
                       const int* __restrict__ wprefix, int* __restrict__ labels) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = rootlab[i];
        labels[i] = r < 0 ? -1 : wprefix[r >> 5] + __popc(bits[r >> 5] & ((1u << (r & 31)) - 1u));
    }
}

#define RHCCQ_LT_H 16         // tile height of the count and attachment passes
#define RHCCQ_LT_UH 32        // tile height of the union passes: fewer edges cross tile borders

static int rhccq_lt_args_make(int H, int W, double eps, int min_pts, rhccq_lt_args* A) {
    if (H < 1 || W < 1 || (long long)H * W > 2000000000LL || !(eps > 0.0) || min_pts < 1) {
        rhccq_set_error("rhccq_dbscan_lattice: need 1 <= H*W <= 2e9, eps > 0, min_pts >= 1");
        return -1;
    }
    const double r2 = eps * eps;
    int R = 0;
    while ((double)(R + 1) * (R + 1) <= r2) ++R;
    if (R > RHCCQ_LT_MAXR || r2 >= 65000.0) {
        rhccq_set_error("rhccq_dbscan_lattice: eps %.3f needs a stencil radius above %d; use the generic path", eps, RHCCQ_LT_MAXR);
        return -1;
    }
    A->H = H; A->W = W; A->R = R; A->thr = (int)floor(r2); A->min_pts = min_pts; A->n_off = 0;
    return 0;
}
static size_t rhccq_lt_smem(const rhccq_lt_args& A, int TH) {
    const size_t side = 2 * (size_t)A.R + 1;
    return ((size_t)(RHCCQ_LT_W + 2 * A.R) * (TH + 2 * A.R) + 2) * 4 + side * side * 8 + (size_t)RHCCQ_LT_W * TH * 4 + RHCCQ_LT_PAIRS * 8;
}
static int rhccq_lt_grid(const rhccq_lt_args& A, int TH) {
    const long long tiles = (long long)((A.W + RHCCQ_LT_W - 1) / RHCCQ_LT_W) * ((A.H + TH - 1) / TH);
    const long long cap = (long long)rhccq_sm_count() * 32;
    return (int)(tiles < cap ? tiles : cap);
}

extern "C" {

size_t rhccq_dbscan_lattice_workspace_bytes(int H, int W) {
    const size_t n = (size_t)H * W;
    return rhccq_al(n * 4) * 4 + rhccq_al(rhccq_scan_scratch_ints((long long)n) * 4 + 64) + 256;
}

struct rhccq_lt_ws { uint32_t* packed; int* parent; int* rootlab; int* is_root; int* scan; int* status; };
static void rhccq_lt_carve(int H, int W, void* ws, rhccq_lt_ws* L) {
    const size_t n = (size_t)H * W;
    unsigned char* p = (unsigned char*)ws;
    L->packed = (uint32_t*)p; p += rhccq_al(n * 4);
    L->parent = (int*)p; p += rhccq_al(n * 4);
    L->rootlab = (int*)p; p += rhccq_al(n * 4);
    L->is_root = (int*)p; p += rhccq_al(n * 4);
    L->scan = (int*)p; p += rhccq_al(rhccq_scan_scratch_ints((long long)n) * 4 + 64);
    L->status = (int*)p;
}

#define RHCCQ_LT_PROLOGUE(name)                                                                         \
    rhccq_lt_args A;                                                                                    \
    if (rhccq_lt_args_make(H, W, eps, min_pts, &A) != 0) return -1;                                     \
    if (!ws || ws_bytes < rhccq_dbscan_lattice_workspace_bytes(H, W)) {                                 \
        rhccq_set_error(name ": workspace missing or too small"); return -1; }                          \
    rhccq_lt_ws L; rhccq_lt_carve(H, W, ws, &L);                                                        \
    const size_t smem = rhccq_lt_smem(A, RHCCQ_LT_H); const int grid = rhccq_lt_grid(A, RHCCQ_LT_H);                       \
    (void)smem; (void)grid;

int rhccq_dbscan_lattice_count(const void* src, int src_kind, int H, int W, double eps, int min_pts, int32_t* count,
                               uint8_t* core, int32_t* status, void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_LT_PROLOGUE("rhccq_dbscan_lattice_count")
    if (!src || !count || !core || !status || (src_kind != 0 && src_kind != 1)) { rhccq_set_error("rhccq_dbscan_lattice_count: bad arguments"); return -1; }
#ifdef RHCCQ_HOST_EMU
    *status = 0;
#else
    cudaMemsetAsync(status, 0, 4, (cudaStream_t)stream);
#endif
    const size_t csmem = rhccq_lt_smem(A, RHCCQ_LT_UH);
    const int cgrid = rhccq_lt_grid(A, RHCCQ_LT_UH);
    if (A.R >= 1 && A.R <= 4) {                                    // sliding-window form
        const size_t rsmem = (size_t)RHCCQ_LTT_TW * (RHCCQ_LTT_TH + 2 * A.R) * 4;
#ifndef RHCCQ_HOST_EMU
        if (src_kind == 0 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {     // persistent, TMA-fed
            const size_t tsmem = rsmem * 6;                        // staging (20 B / pixel) + tile (4 B / pixel)
            const long long tiles = (long long)((W + RHCCQ_LT_W - 1) / RHCCQ_LT_W) * ((H + RHCCQ_LTT_TH - 1) / RHCCQ_LTT_TH);
            const long long cap = (long long)rhccq_sm_count() * 3;
            const int tgrid = (int)(tiles < cap ? tiles : cap);
#define RHCCQ_LT_TMA(RT, THR)                                                                                        \
    do { if (rhccq_smem_optin((const void*)rhccq_k_lt_count_tma<RT, THR>, tsmem) != 0) return -1;                    \
         RHCCQ_LAUNCH((rhccq_k_lt_count_tma<RT, THR>), tgrid, RHCCQ_PT_THREADS, tsmem, (cudaStream_t)stream,         \
                      (const float*)src, A, count, L.packed, core, status); } while (0)
            // integer radii (thr = R^2) get a kernel with the stencil resolved at compile time
            if (A.R == 1) { if (A.thr == 1) RHCCQ_LT_TMA(1, 1); else RHCCQ_LT_TMA(1, -1); }
            else if (A.R == 2) { if (A.thr == 4) RHCCQ_LT_TMA(2, 4); else RHCCQ_LT_TMA(2, -1); }
            else if (A.R == 3) { if (A.thr == 9) RHCCQ_LT_TMA(3, 9); else RHCCQ_LT_TMA(3, -1); }
            else { if (A.thr == 16) RHCCQ_LT_TMA(4, 16); else RHCCQ_LT_TMA(4, -1); }
#undef RHCCQ_LT_TMA
            return 0;
        }
#endif
#define RHCCQ_LT_ROWS(SRCK, RT)                                                                                      \
    RHCCQ_LAUNCH((rhccq_k_lt_count_rows<SRCK, RT>), cgrid, RHCCQ_PT_THREADS, rsmem, (cudaStream_t)stream, src, A, count,   \
                 L.packed, core, status)
        if (src_kind == 0) {
            if (A.R == 1) RHCCQ_LT_ROWS(0, 1); else if (A.R == 2) RHCCQ_LT_ROWS(0, 2); else if (A.R == 3) RHCCQ_LT_ROWS(0, 3); else RHCCQ_LT_ROWS(0, 4);
        } else {
            if (A.R == 1) RHCCQ_LT_ROWS(1, 1); else if (A.R == 2) RHCCQ_LT_ROWS(1, 2); else if (A.R == 3) RHCCQ_LT_ROWS(1, 3); else RHCCQ_LT_ROWS(1, 4);
        }
        return 0;
    }
#undef RHCCQ_LT_ROWS
    if (src_kind == 0) {
        if (rhccq_smem_optin((const void*)rhccq_k_lt_sweep<0, 0, RHCCQ_LT_UH>, csmem) != 0) return -1;
        RHCCQ_LAUNCH((rhccq_k_lt_sweep<0, 0, RHCCQ_LT_UH>), cgrid, RHCCQ_PT_THREADS, csmem, (cudaStream_t)stream, src, A, count, L.packed, core, L.parent, L.rootlab, status);
    } else {
        if (rhccq_smem_optin((const void*)rhccq_k_lt_sweep<0, 1, RHCCQ_LT_UH>, csmem) != 0) return -1;
        RHCCQ_LAUNCH((rhccq_k_lt_sweep<0, 1, RHCCQ_LT_UH>), cgrid, RHCCQ_PT_THREADS, csmem, (cudaStream_t)stream, src, A, count, L.packed, core, L.parent, L.rootlab, status);
    }
    return 0;
}

int rhccq_dbscan_lattice_union(int H, int W, double eps, int min_pts, void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_LT_PROLOGUE("rhccq_dbscan_lattice_union")
    // (the tile-local pass writes the parent of every pixel)
    const size_t usmem = rhccq_lt_smem(A, RHCCQ_LT_UH);
    const int ugrid = rhccq_lt_grid(A, RHCCQ_LT_UH);
    if (A.R >= 1 && A.R <= 4) {
        const size_t rsmem = (size_t)RHCCQ_LTT_TW * (RHCCQ_LTT_TH + A.R) * 8;
        const long long tiles = (long long)((W + RHCCQ_LT_W - 1) / RHCCQ_LT_W) * ((H + RHCCQ_LTT_TH - 1) / RHCCQ_LTT_TH);
        const long long cap = (long long)rhccq_sm_count() * 32;
        const int rgrid = (int)(tiles < cap ? tiles : cap);
#define RHCCQ_LT_UROWS(RT) RHCCQ_LAUNCH((rhccq_k_lt_union_rows<RT>), rgrid, RHCCQ_PT_THREADS, rsmem, (cudaStream_t)stream, L.packed, A, L.parent)
        if (A.R == 1) RHCCQ_LT_UROWS(1); else if (A.R == 2) RHCCQ_LT_UROWS(2); else if (A.R == 3) RHCCQ_LT_UROWS(3); else RHCCQ_LT_UROWS(4);
#undef RHCCQ_LT_UROWS
    } else {
        if (rhccq_smem_optin((const void*)rhccq_k_lt_union_tile<RHCCQ_LT_UH>, usmem) != 0) return -1;
        RHCCQ_LAUNCH((rhccq_k_lt_union_tile<RHCCQ_LT_UH>), ugrid, RHCCQ_PT_THREADS, usmem, (cudaStream_t)stream, L.packed, A, L.parent);
    }
    if (rhccq_smem_optin((const void*)rhccq_k_lt_sweep<3, 2, RHCCQ_LT_UH>, usmem) != 0) return -1;
    RHCCQ_LAUNCH((rhccq_k_lt_sweep<3, 2, RHCCQ_LT_UH>), ugrid, RHCCQ_PT_THREADS, usmem, (cudaStream_t)stream, (const void*)L.packed, A,
                 (int*)nullptr, L.packed, (uint8_t*)nullptr, L.parent, L.rootlab, (int*)nullptr);
    return 0;
}

int rhccq_dbscan_lattice_attach(int H, int W, double eps, int min_pts, void* ws, size_t ws_bytes, void* stream);

int rhccq_dbscan_lattice_flatten(int H, int W, const uint8_t* core, void* ws, size_t ws_bytes, void* stream) {
    if (!ws || ws_bytes < rhccq_dbscan_lattice_workspace_bytes(H, W) || !core) { rhccq_set_error("rhccq_dbscan_lattice_flatten: bad arguments"); return -1; }
    rhccq_lt_ws L; rhccq_lt_carve(H, W, ws, &L);
    RHCCQ_LAUNCH(rhccq_k_pt_flatten, rhccq_pt_blocks((long long)H * W), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, H * W, L.parent, core, L.rootlab, L.is_root);
    return 0;
}

size_t rhccq_dbscan_lattice_ws_offset(int H, int W, int which) {
    rhccq_lt_ws L; rhccq_lt_carve(H, W, (void*)256, &L);
    const unsigned char* p = which == 0 ? (unsigned char*)L.rootlab : which == 1 ? (unsigned char*)L.is_root
                           : which == 2 ? (unsigned char*)L.parent : which == 3 ? (unsigned char*)L.packed : (unsigned char*)L.scan;
    return (size_t)(p - (unsigned char*)256);
}

int rhccq_dbscan_lattice_border(int H, int W, double eps, int min_pts, const uint8_t* core, void* ws, size_t ws_bytes, void* stream) {
    if (rhccq_dbscan_lattice_flatten(H, W, core, ws, ws_bytes, stream) != 0) return -1;
    return rhccq_dbscan_lattice_attach(H, W, eps, min_pts, ws, ws_bytes, stream);
}

int rhccq_dbscan_lattice_attach(int H, int W, double eps, int min_pts, void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_LT_PROLOGUE("rhccq_dbscan_lattice_attach")
    if (A.R >= 1 && A.R <= 4) {                                    // sliding-window form
        const size_t rsmem = (size_t)RHCCQ_LTT_TW * (RHCCQ_LTT_TH + 2 * A.R) * 8;
        const long long tiles = (long long)((W + RHCCQ_LT_W - 1) / RHCCQ_LT_W) * ((H + RHCCQ_LTT_TH - 1) / RHCCQ_LTT_TH);
        const long long cap = (long long)rhccq_sm_count() * 32;
        const int rgrid = (int)(tiles < cap ? tiles : cap);
#define RHCCQ_LT_AROWS(RT) RHCCQ_LAUNCH((rhccq_k_lt_attach_rows<RT>), rgrid, RHCCQ_PT_THREADS, rsmem, (cudaStream_t)stream, L.packed, A, L.rootlab)
        if (A.R == 1) RHCCQ_LT_AROWS(1); else if (A.R == 2) RHCCQ_LT_AROWS(2); else if (A.R == 3) RHCCQ_LT_AROWS(3); else RHCCQ_LT_AROWS(4);
#undef RHCCQ_LT_AROWS
        return 0;
    }
    if (rhccq_smem_optin((const void*)rhccq_k_lt_sweep<2, 2, RHCCQ_LT_H>, smem) != 0) return -1;
    RHCCQ_LAUNCH((rhccq_k_lt_sweep<2, 2, RHCCQ_LT_H>), grid, RHCCQ_PT_THREADS, smem, (cudaStream_t)stream, (const void*)L.packed, A, (int*)nullptr, L.packed,
                 (uint8_t*)nullptr, L.parent, L.rootlab, (int*)nullptr);
    return 0;
}

int rhccq_dbscan_lattice_relabel(int H, int W, void* ws, size_t ws_bytes, int32_t* labels, void* stream) {
    if (!ws || ws_bytes < rhccq_dbscan_lattice_workspace_bytes(H, W) || !labels) { rhccq_set_error("rhccq_dbscan_lattice_relabel: bad arguments"); return -1; }
    rhccq_lt_ws L; rhccq_lt_carve(H, W, ws, &L);
    const long long n = (long long)H * W;
    // The rank of a root among the roots, without a scan over all n flags: the flags as a bit mask (one word per 32
    // pixels), an exclusive scan over the n / 32 word populations, and rank(r) = prefix[r / 32] + the bits below r
    // in its word.  Both tables live in the parent array, which is dead once the forest has been flattened.
    const long long m = (n + 31) / 32;
    uint32_t* bits = reinterpret_cast<uint32_t*>(L.parent);
    int* wpre = L.parent + ((m + 63) & ~63LL);
    if (2 * ((m + 63) & ~63LL) > n) {                              // (tiny images: the plain way)
        if (n <= RHCCQ_SCAN_TILE) {
            RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, L.is_root, (int)n, (int*)nullptr);
        } else if (rhccq_scan_i32(L.is_root, n, L.is_root, L.scan, stream) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_pt_labels, rhccq_pt_blocks(n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, (int)n, L.rootlab, L.is_root, labels);
        return 0;
    }
    RHCCQ_LAUNCH(rhccq_k_lt_root_bits, rhccq_pt_blocks(n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, L.is_root, n, bits, wpre);
    if (m <= RHCCQ_SCAN_TILE) {
        RHCCQ_LAUNCH(rhccq_k_scan_small, 1, RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, wpre, (int)m, (int*)nullptr);
    } else if (rhccq_scan_i32(wpre, m, wpre, L.scan, stream) != 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_lt_labels_bits, rhccq_pt_blocks(n), RHCCQ_PT_THREADS, 0, (cudaStream_t)stream, n, L.rootlab, bits, wpre, labels);
    return 0;
}

}  // extern "C"
