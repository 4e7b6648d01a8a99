// Per-pixel kernels of the stage-1 driver: one CTA per segment crop.
//
//   rhccq_k_unique        sorted unique colours of a segment + index per pixel
//                         (/root/reference/encoder/compression/clustering.py:4-103,
//                         with the crop / mask / black-repaint rules of
//                         encoder/compression/subregions.py:315-421)
//   rhccq_k_remap_first   indices <- LUT[indices] (clustering.py:373-377) and the
//                         raster position at which every new colour first
//                         appears (the order encoder/compression/merging.py:52-82
//                         assigns palette slots in)
//   rhccq_k_paint         final index plane through the composed entry table
//
// A crop is eight int32: image b, row r0, column c0, height h, width w, segment
// id (0 = every pixel of the rectangle belongs to the crop), class k, unused.
// Pixels of the rectangle whose label differs are "outside": black in the
// reference's crop.  Label maps and index planes are laid out [K,B,H,W] (one
// plane set per class: the ROI and non-ROI calls of the stage-1 driver have
// their own label maps, encoder/compression/test.py:105-106); the image is
// [B,H,W,3].
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"

#define RHCCQ_PAD_KEY 0xFFFFFFFFu

#define RHCCQ_UQ_BUCKETS 2048           // buckets of the hashed path (key-ordered: by R, then by a slice of G)
#define RHCCQ_UQ_MAXB 256               // largest bucket the hashed path ranks by counting; beyond: the sorting path
#define RHCCQ_UQ_EMPTY 0xFFFFFFFFu

size_t rhccq_unique_ws_bytes(int max_valid) {
    size_t np2 = 1;
    while (np2 < (size_t)(max_valid > 1 ? max_valid : 1)) np2 <<= 1;
    const size_t sorting = rhccq_carve_bytes(np2, 4) * 2;
    // hashed path: table of 2 np2 keys, a 16-bit rank per slot, the slots grouped by bucket, the bucket counters
    const size_t hashed = rhccq_carve_bytes(2 * np2, 4) + rhccq_carve_bytes(2 * np2, 2) + rhccq_carve_bytes(np2, 2)
                          + rhccq_carve_bytes(RHCCQ_UQ_BUCKETS, 4);
    return np2 <= 16384 && hashed > sorting ? hashed : sorting;
}

__device__ __forceinline__ int rhccq_lower_bound_u32(const uint32_t* a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t rhccq_uq_hash(uint32_t key, uint32_t mask) { return (key * 2654435761u >> 7) & mask; }

// Hashed form of the same result for segments of up to 16 384 pixels: the colours go into an open-addressing
// table in shared memory (duplicates vanish there: ~45 % of the pixels), the distinct ones are grouped into
// key-ordered buckets (R, then a slice of G scaled to the segment's range) and ranked inside their bucket by
// counting — np.unique's lexicographic order without sorting every pixel — and a pixel's index is one table
// lookup.  Returns false (nothing written) when a bucket is too large for counting: the caller then sorts.
template <class IdxT>
__device__ bool rhccq_unique_hashed(int p, const uint8_t* __restrict__ img, const int32_t* __restrict__ seg,
                                    int B, int H, int W, const int32_t* __restrict__ crops,
                                    const int* __restrict__ pal_off, uint32_t* __restrict__ pal_keys,
                                    int* __restrict__ pal_cnt, IdxT* __restrict__ index_plane, int repaint_black,
                                    int cap, unsigned char* wsbase) {
    __shared__ int s_uniq, s_valid, s_black, s_rg[4], s_maxb;
    __shared__ unsigned long long s_best;
    __shared__ uint32_t s_repl;
    __shared__ int s_scr[RHCCQ_MAX_WARPS + 2];
    const int32_t* cr = crops + 8 * (size_t)p;
    const int b = cr[0], r0 = cr[1], c0 = cr[2], h = cr[3], w = cr[4], sid = cr[5];
    const int npx = h * w;
    const size_t iplane = (size_t)b * H * W;
    const size_t plane = ((size_t)cr[6] * B + b) * H * W;
    int np2cap = 1;
    while (np2cap < (cap > 1 ? cap : 1)) np2cap <<= 1;
    const int nslots = 2 * np2cap;
    const float inv_w = 1.0f / (float)w;                           // row of pixel q without an integer division (npx <= 16 384)
    const uint32_t mask = (uint32_t)nslots - 1u;
    rhccq_carver cv(wsbase);
    uint32_t* hk = cv.take<uint32_t>(nslots);
    uint16_t* rk = cv.take<uint16_t>(nslots);
    uint16_t* bslot = cv.take<uint16_t>(np2cap);
    int* hist = cv.take<int>(RHCCQ_UQ_BUCKETS);
    if (threadIdx.x == 0) {
        s_uniq = 0; s_valid = 0; s_black = 0; s_best = ~0ull; s_repl = 0u; s_maxb = 0;
        s_rg[0] = 255; s_rg[1] = 0; s_rg[2] = 255; s_rg[3] = 0;
    }
    RHCCQ_PAR_FOR(j, nslots) hk[j] = RHCCQ_UQ_EMPTY;
    RHCCQ_PAR_FOR(j, RHCCQ_UQ_BUCKETS) hist[j] = 0;
    __syncthreads();
    // pass A: the distinct non-black colours of the segment
    unsigned long long best_local = ~0ull;
    int rmin = 255, rmax = 0, gmin = 255, gmax = 0;
    for (int q = (int)threadIdx.x; q - RHCCQ_LANE < npx; q += (int)blockDim.x) {          // warp-uniform trip count
        bool valid = q < npx;
        uint32_t key = 0u;
        if (valid) {
            int rq = (int)(((float)q + 0.5f) * inv_w), cq = q - rq * w;
            if (cq < 0) { --rq; cq += w; } else if (cq >= w) { ++rq; cq -= w; }
            const int r = r0 + rq, c = c0 + cq;
            const size_t pos = plane + (size_t)r * W + c;
            if (seg != nullptr && sid != 0 && seg[pos] != sid) valid = false;
            else {
                const uint8_t* px = img + 3 * (iplane + (size_t)r * W + c);
                key = rhccq_pack_rgb(px[0], px[1], px[2]);
            }
        }
        const bool nonblack = valid && key != 0u;
        bool fresh = false;
        if (nonblack) {
            uint32_t hslot = rhccq_uq_hash(key, mask);
            while (true) {
                const uint32_t k = atomicCAS(&hk[hslot], RHCCQ_UQ_EMPTY, key);
                if (k == RHCCQ_UQ_EMPTY) { fresh = true; break; }
                if (k == key) break;
                hslot = (hslot + 1u) & mask;
            }
            if (fresh) {
                const int rr = rhccq_key_r(key), gg = rhccq_key_g(key);
                rmin = rr < rmin ? rr : rmin; rmax = rr > rmax ? rr : rmax;
                gmin = gg < gmin ? gg : gmin; gmax = gg > gmax ? gg : gmax;
            }
            if (repaint_black) {
                const unsigned long long cand = ((unsigned long long)rhccq_d2(key, 0u) << 32) | (unsigned)q;
                best_local = cand < best_local ? cand : best_local;
            }
        }
        const unsigned mv = rhccq_ballot(valid), mnb = rhccq_ballot(nonblack), mf = rhccq_ballot(fresh);
        if (RHCCQ_LANE == 0) {
            if (mf) atomicAdd(&s_uniq, __popc(mf));
            if (mv) atomicAdd(&s_valid, __popc(mv));
            if (mv & ~mnb) atomicAdd(&s_black, __popc(mv & ~mnb));
        }
    }
    for (int d = RHCCQ_WARP_SIZE >> 1; d > 0; d >>= 1) {
        const unsigned long long o = rhccq_shfl_xor(best_local, d);
        best_local = o < best_local ? o : best_local;
        const int a0 = rhccq_shfl_xor(rmin, d), a1 = rhccq_shfl_xor(rmax, d), a2 = rhccq_shfl_xor(gmin, d), a3 = rhccq_shfl_xor(gmax, d);
        rmin = a0 < rmin ? a0 : rmin; rmax = a1 > rmax ? a1 : rmax; gmin = a2 < gmin ? a2 : gmin; gmax = a3 > gmax ? a3 : gmax;
    }
    if (RHCCQ_LANE == 0) {
        if (repaint_black && best_local != ~0ull) atomicMin(&s_best, best_local);
        atomicMin(&s_rg[0], rmin); atomicMax(&s_rg[1], rmax); atomicMin(&s_rg[2], gmin); atomicMax(&s_rg[3], gmax);
    }
    __syncthreads();
    const int nu = s_uniq, nvalid = s_valid, nblack = s_black;
    if (nvalid > cap) {                                            // more pixels than the table was sized for: report
        if (threadIdx.x == 0) pal_cnt[p] = -1;
        return true;
    }
    const bool repaint = repaint_black && nblack > 0 && nu > 0;
    if (repaint && threadIdx.x == 0) {
        const int q = (int)(s_best & 0xffffffffu);
        const uint8_t* px = img + 3 * (iplane + (size_t)(r0 + q / w) * W + (c0 + q % w));
        s_repl = rhccq_pack_rgb(px[0], px[1], px[2]);
    }
    const int has_black = (nvalid < npx || (nblack > 0 && !repaint)) ? 1 : 0;
    // key-ordered buckets: (R - Rmin) * gb + floor((G - Gmin) * gb / Grange)
    const int Rmin = s_rg[0], Gmin = s_rg[2];
    const int rrange = nu > 0 ? s_rg[1] - Rmin + 1 : 1, grange = nu > 0 ? s_rg[3] - Gmin + 1 : 1;
    int gb = RHCCQ_UQ_BUCKETS / rrange;
    gb = gb > 64 ? 64 : (gb < 1 ? 1 : gb);
    const float gscale = (float)gb / (float)grange;
    auto bucket = [&](uint32_t key) -> int {
        int gs = (int)((float)(rhccq_key_g(key) - Gmin) * gscale);
        gs = gs > gb - 1 ? gb - 1 : gs;
        return (rhccq_key_r(key) - Rmin) * gb + gs;
    };
    RHCCQ_PAR_FOR(j, nslots) {
        const uint32_t key = hk[j];
        if (key != RHCCQ_UQ_EMPTY) atomicAdd(&hist[bucket(key)], 1);
    }
    __syncthreads();
    RHCCQ_PAR_FOR(j, RHCCQ_UQ_BUCKETS) if (hist[j] > RHCCQ_UQ_MAXB) atomicMax(&s_maxb, hist[j]);
    __syncthreads();
    if (s_maxb > RHCCQ_UQ_MAXB) return false;                      // (uniform) a bucket too large to rank by counting
    rhccq_block_excl_scan_array<int>(hist, RHCCQ_UQ_BUCKETS, s_scr);          // hist[b] = first position of bucket b
    RHCCQ_PAR_FOR(j, nslots) {
        const uint32_t key = hk[j];
        if (key != RHCCQ_UQ_EMPTY) bslot[atomicAdd(&hist[bucket(key)], 1)] = (uint16_t)j;   // hist[b] ends as the end of b
    }
    __syncthreads();
    uint32_t* pal = pal_keys + pal_off[p];
    if (threadIdx.x == 0) {
        if (has_black) pal[0] = 0u;
        pal_cnt[p] = ((size_t)(has_black + nu) > (size_t)((IdxT)~(IdxT)0) + 1) ? -3 : has_black + nu;
    }
    RHCCQ_PAR_FOR(q, nu) {
        const int slot = (int)bslot[q];
        const uint32_t key = hk[slot];
        const int bk = bucket(key);
        const int lo = bk > 0 ? hist[bk - 1] : 0, hi = hist[bk];
        int r = 0;
        for (int t = lo; t < hi; ++t) r += hk[bslot[t]] < key ? 1 : 0;
        rk[slot] = (uint16_t)(lo + r);
        pal[has_black + lo + r] = key;
    }
    __syncthreads();
    // pass B: index of every pixel of the segment
    const uint32_t repl = s_repl;
    RHCCQ_PAR_FOR(q, npx) {
        int rq = (int)(((float)q + 0.5f) * inv_w), cq = q - rq * w;
        if (cq < 0) { --rq; cq += w; } else if (cq >= w) { ++rq; cq -= w; }
        const int r = r0 + rq, c = c0 + cq;
        const size_t pos = plane + (size_t)r * W + c;
        if (seg != nullptr && sid != 0 && seg[pos] != sid) continue;
        const uint8_t* px = img + 3 * (iplane + (size_t)r * W + c);
        uint32_t key = rhccq_pack_rgb(px[0], px[1], px[2]);
        if (key == 0u && repaint) key = repl;
        int idx = 0;
        if (key != 0u) {
            uint32_t hslot = rhccq_uq_hash(key, mask);
            while (hk[hslot] != key) hslot = (hslot + 1u) & mask;
            idx = has_black + (int)rk[hslot];
        }
        index_plane[pos] = (IdxT)idx;
    }
    return true;
}

template <class IdxT>
__device__ void rhccq_unique_problem(int p, const uint8_t* __restrict__ img, const int32_t* __restrict__ seg,
                                     int B, int H, int W, const int32_t* __restrict__ crops,
                                     const int* __restrict__ pal_off, uint32_t* __restrict__ pal_keys,
                                     int* __restrict__ pal_cnt, IdxT* __restrict__ index_plane, int repaint_black,
                                     int cap, unsigned char* wsbase) {
    __shared__ int s_nb, s_valid, s_black;
    __shared__ unsigned long long s_best;
    __shared__ uint32_t s_repl;
    __shared__ int s_scr[RHCCQ_MAX_WARPS + 2];
    const int32_t* cr = crops + 8 * (size_t)p;
    const int b = cr[0], r0 = cr[1], c0 = cr[2], h = cr[3], w = cr[4], sid = cr[5];
    const int npx = h * w;
    const size_t iplane = (size_t)b * H * W;                       // image
    const size_t plane = ((size_t)cr[6] * B + b) * H * W;          // label map / index plane of the class
    rhccq_carver cv(wsbase);
    int np2cap = 1;
    while (np2cap < (cap > 1 ? cap : 1)) np2cap <<= 1;
    uint32_t* buf = cv.take<uint32_t>(np2cap);
    int* rank = cv.take<int>(np2cap);
    if (threadIdx.x == 0) { s_nb = 0; s_valid = 0; s_black = 0; s_best = ~0ull; s_repl = 0u; }
    __syncthreads();
    // pass A: collect the non-black colours of the segment.  The three counters are bumped once per warp
    // (ballot + popcount) and the repaint candidate is reduced in registers first: per-pixel atomics on four
    // shared addresses were the bulk of this kernel's time.
    unsigned long long best_local = ~0ull;
    for (int q = (int)threadIdx.x; q - RHCCQ_LANE < npx; q += (int)blockDim.x) {          // warp-uniform trip count
        bool valid = q < npx;
        uint32_t key = 0u;
        if (valid) {
            const int r = r0 + q / w, c = c0 + q % w;
            const size_t pos = plane + (size_t)r * W + c;
            if (seg != nullptr && sid != 0 && seg[pos] != sid) valid = false;
            else {
                const uint8_t* px = img + 3 * (iplane + (size_t)r * W + c);
                key = rhccq_pack_rgb(px[0], px[1], px[2]);
            }
        }
        const bool nonblack = valid && key != 0u;
        const unsigned mv = rhccq_ballot(valid), mnb = rhccq_ballot(nonblack);
        int base = 0;
        if (RHCCQ_LANE == 0) {
            if (mnb) base = atomicAdd(&s_nb, __popc(mnb));
            if (mv) atomicAdd(&s_valid, __popc(mv));
            if (mv & ~mnb) atomicAdd(&s_black, __popc(mv & ~mnb));
        }
        base = rhccq_shfl(base, 0);
        if (nonblack) {
            const int slot = base + __popc(mnb & rhccq_lanemask_lt());
            if (slot < cap) buf[slot] = key;
            if (repaint_black) {
                // "nearest" colour to black = smallest norm, first in raster order (subregions.py:406-416)
                const unsigned long long cand = ((unsigned long long)rhccq_d2(key, 0u) << 32) | (unsigned)q;
                best_local = cand < best_local ? cand : best_local;
            }
        }
    }
    if (repaint_black) {
        for (int d = RHCCQ_WARP_SIZE >> 1; d > 0; d >>= 1) {
            const unsigned long long o = rhccq_shfl_xor(best_local, d);
            best_local = o < best_local ? o : best_local;
        }
        if (RHCCQ_LANE == 0 && best_local != ~0ull) atomicMin(&s_best, best_local);
    }
    __syncthreads();
    const int nb = s_nb, nvalid = s_valid, nblack = s_black;
    if (nb > cap) {                                                // capacity exceeded: report, never truncate
        if (threadIdx.x == 0) pal_cnt[p] = -1;
        return;
    }
    const bool repaint = repaint_black && nblack > 0 && nb > 0;
    if (repaint && threadIdx.x == 0) {
        const int q = (int)(s_best & 0xffffffffu);
        const uint8_t* px = img + 3 * (iplane + (size_t)(r0 + q / w) * W + (c0 + q % w));
        s_repl = rhccq_pack_rgb(px[0], px[1], px[2]);
    }
    const int has_black = (nvalid < npx || (nblack > 0 && !repaint)) ? 1 : 0;
    const int np2 = rhccq_next_pow2(nb > 1 ? nb : 1);
    for (int j = nb + (int)threadIdx.x; j < np2; j += (int)blockDim.x) buf[j] = RHCCQ_PAD_KEY;
    __syncthreads();
    rhccq_block_bitonic_sort<uint32_t>(buf, np2);
    RHCCQ_PAR_FOR(j, np2) rank[j] = (j < nb && (j == 0 || buf[j] != buf[j - 1])) ? 1 : 0;
    __syncthreads();
    const int n_unique = rhccq_block_excl_scan_array<int>(rank, np2, s_scr);
    uint32_t* pal = pal_keys + pal_off[p];
    if (threadIdx.x == 0) {
        if (has_black) pal[0] = 0u;
        pal_cnt[p] = ((size_t)(has_black + n_unique) > (size_t)((IdxT)~(IdxT)0) + 1) ? -3 : has_black + n_unique;
    }
    RHCCQ_PAR_FOR(j, nb) if (j == 0 || buf[j] != buf[j - 1]) pal[has_black + rank[j]] = buf[j];
    // pass B: index of every pixel of the segment
    const uint32_t repl = s_repl;
    RHCCQ_PAR_FOR(q, npx) {
        const int r = r0 + q / w, c = c0 + q % w;
        const size_t pos = plane + (size_t)r * W + c;
        if (seg != nullptr && sid != 0 && seg[pos] != sid) continue;
        const uint8_t* px = img + 3 * (iplane + (size_t)r * W + c);
        uint32_t key = rhccq_pack_rgb(px[0], px[1], px[2]);
        if (key == 0u && repaint) key = repl;
        int idx = 0;
        if (key != 0u) idx = has_black + rank[rhccq_lower_bound_u32(buf, nb, key)];
        index_plane[pos] = (IdxT)idx;
    }
}

template <class IdxT>
__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_unique(const uint8_t* __restrict__ img, const int32_t* __restrict__ seg, int B, int H, int W,
               const int32_t* __restrict__ crops, int n_crops, const int* __restrict__ pal_off,
               uint32_t* __restrict__ pal_keys, int* __restrict__ pal_cnt, IdxT* __restrict__ index_plane,
               int repaint_black, int cap, unsigned char* gws, size_t gws_stride) {
    RHCCQ_DYN_SMEM(dyn);
    unsigned char* wsbase = gws ? gws + (size_t)blockIdx.x * gws_stride : dyn;
    for (int p = blockIdx.x; p < n_crops; p += gridDim.x) {
        bool done = false;
        if (cap <= 16384 && (size_t)((IdxT)~(IdxT)0) >= 0xffffu)   // 16-bit ranks in the table
            done = rhccq_unique_hashed<IdxT>(p, img, seg, B, H, W, crops, pal_off, pal_keys, pal_cnt, index_plane,
                                             repaint_black, cap, wsbase);
        __syncthreads();
        if (!done)
            rhccq_unique_problem<IdxT>(p, img, seg, B, H, W, crops, pal_off, pal_keys, pal_cnt, index_plane,
                                       repaint_black, cap, wsbase);
        __syncthreads();
    }
}

int rhccq_launch_unique(const uint8_t* img, const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                        const int* pal_off, uint32_t* pal_keys, int* pal_cnt, void* index_plane, int idx_bytes,
                        int repaint_black, int max_valid, rhccq_launch_ws ws, void* stream) {
    if (n_crops <= 0) return 0;
    const size_t need = rhccq_unique_ws_bytes(max_valid);
    size_t smem; unsigned char* gws;
    if (idx_bytes == 2) {
        const int grid = rhccq_pick_grid((const void*)rhccq_k_unique<uint16_t>, need, n_crops, ws, &smem, &gws,
                                         "rhccq_unique_index");
        if (grid < 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_unique<uint16_t>, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                     img, seg, B, H, W, crops, n_crops, pal_off, pal_keys, pal_cnt, (uint16_t*)index_plane,
                     repaint_black, max_valid, gws, need);
    } else if (idx_bytes == 4) {
        const int grid = rhccq_pick_grid((const void*)rhccq_k_unique<uint32_t>, need, n_crops, ws, &smem, &gws,
                                         "rhccq_unique_index");
        if (grid < 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_unique<uint32_t>, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                     img, seg, B, H, W, crops, n_crops, pal_off, pal_keys, pal_cnt, (uint32_t*)index_plane,
                     repaint_black, max_valid, gws, need);
    } else {
        rhccq_set_error("rhccq_unique_index: idx_bytes must be 2 or 4, got %d", idx_bytes);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- remap + first appearance
//
// Entries of crop p live at ent_off[p] .. ent_off[p] + n_leaves[p] of the
// compact entry table: colour (the clustered palette row) and the smallest
// raster position r * W + c at which a pixel of the crop carries it
// (0xFFFFFFFF: no pixel does).
template <class IdxT>
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_remap_first(const int32_t* __restrict__ seg, int B, int H, int W, const int32_t* __restrict__ crops, int n_crops,
                    const int* __restrict__ pal_off, const int* __restrict__ leaf, const int* __restrict__ n_leaves,
                    const uint32_t* __restrict__ new_keys, const int* __restrict__ ent_off,
                    IdxT* __restrict__ index_plane, uint32_t* __restrict__ ent_color, uint32_t* __restrict__ ent_fpos,
                    int smem_rows) {
    RHCCQ_DYN_SMEM(dyn);
    uint32_t* s_first = reinterpret_cast<uint32_t*>(dyn);
    for (int p = blockIdx.x; p < n_crops; p += gridDim.x) {
        const int32_t* cr = crops + 8 * (size_t)p;
        const int b = cr[0], r0 = cr[1], c0 = cr[2], h = cr[3], w = cr[4], sid = cr[5];
        const int npx = h * w;
        const int m = n_leaves[p];
        if (m < 0) continue;                                       // an earlier kernel reported an error for p
        const size_t plane = ((size_t)cr[6] * B + b) * H * W;
        const int* lf = leaf + pal_off[p];
        uint32_t* first = m <= smem_rows ? s_first : ent_fpos + ent_off[p];
        RHCCQ_PAR_FOR(v, m) first[v] = 0xFFFFFFFFu;
        __syncthreads();
        RHCCQ_PAR_FOR(q, npx) {
            const int r = r0 + q / w, c = c0 + q % w;
            const size_t pos = plane + (size_t)r * W + c;
            if (seg != nullptr && sid != 0 && seg[pos] != sid) continue;
            const int v = lf[(int)index_plane[pos]] & 0xffff;      // uint16 LUT (clustering.py:373)
            index_plane[pos] = (IdxT)v;
            atomicMin(&first[v], (uint32_t)(r * W + c));
        }
        __syncthreads();
        RHCCQ_PAR_FOR(v, m) {
            if (first == s_first) ent_fpos[ent_off[p] + v] = s_first[v];
            ent_color[ent_off[p] + v] = new_keys[pal_off[p] + v];
        }
        __syncthreads();
    }
}

int rhccq_launch_remap_first(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                             const int* pal_off, const int* leaf, const int* n_leaves, const uint32_t* new_keys,
                             const int* ent_off, void* index_plane, int idx_bytes, uint32_t* ent_color,
                             uint32_t* ent_fpos, int max_leaves, void* stream) {
    if (n_crops <= 0) return 0;
    int smem_rows = max_leaves;
    if ((size_t)smem_rows * 4 > 64 * 1024) smem_rows = 16 * 1024;
    const size_t smem = (size_t)smem_rows * 4;
    if (idx_bytes == 2) {
        if (rhccq_smem_optin((const void*)rhccq_k_remap_first<uint16_t>, smem) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_remap_first<uint16_t>, n_crops, RHCCQ_PIXEL_THREADS, smem, (cudaStream_t)stream,
                     seg, B, H, W, crops, n_crops, pal_off, leaf, n_leaves, new_keys, ent_off,
                     (uint16_t*)index_plane, ent_color, ent_fpos, smem_rows);
    } else if (idx_bytes == 4) {
        if (rhccq_smem_optin((const void*)rhccq_k_remap_first<uint32_t>, smem) != 0) return -1;
        RHCCQ_LAUNCH(rhccq_k_remap_first<uint32_t>, n_crops, RHCCQ_PIXEL_THREADS, smem, (cudaStream_t)stream,
                     seg, B, H, W, crops, n_crops, pal_off, leaf, n_leaves, new_keys, ent_off,
                     (uint32_t*)index_plane, ent_color, ent_fpos, smem_rows);
    } else {
        rhccq_set_error("rhccq_remap_first: idx_bytes must be 2 or 4, got %d", idx_bytes);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- paint
// out[pos] = ent_final[ent_off[p] + index_plane[pos]] for the pixels of crop p
// whose entry paints (>= 0) and whose crop belongs to class `cls` (cls < 0:
// every crop).  out_plane is [B,H,W].  Launch once per class in the reference's paint
// order (last listed component first) so that the first listed wins overlaps
// (merging.py:52).
template <class IdxT>
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_paint(const int32_t* __restrict__ seg, int B, int H, int W, const int32_t* __restrict__ crops, int n_crops,
              const int* __restrict__ ent_off, const int* __restrict__ ent_final,
              int cls, const IdxT* __restrict__ index_plane, uint16_t* __restrict__ out_plane) {
    for (int p = blockIdx.x; p < n_crops; p += gridDim.x) {
        const int32_t* cr = crops + 8 * (size_t)p;
        if (cls >= 0 && cr[6] != cls) continue;
        const int b = cr[0], r0 = cr[1], c0 = cr[2], h = cr[3], w = cr[4], sid = cr[5];
        const int npx = h * w;
        const size_t oplane = (size_t)b * H * W;
        const size_t plane = ((size_t)cr[6] * B + b) * H * W;
        const int* fin = ent_final + ent_off[p];
        RHCCQ_PAR_FOR(q, npx) {
            const int r = r0 + q / w, c = c0 + q % w;
            const size_t pos = plane + (size_t)r * W + c;
            if (seg != nullptr && sid != 0 && seg[pos] != sid) continue;
            const int f = fin[(int)index_plane[pos]];
            if (f >= 0) out_plane[oplane + (size_t)r * W + c] = (uint16_t)f;
        }
    }
}

int rhccq_launch_paint(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops, const int* ent_off,
                       const int* ent_final, int cls, const void* index_plane, int idx_bytes,
                       uint16_t* out_plane, void* stream) {
    if (n_crops <= 0) return 0;
    if (idx_bytes == 2) {
        RHCCQ_LAUNCH(rhccq_k_paint<uint16_t>, n_crops, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream,
                     seg, B, H, W, crops, n_crops, ent_off, ent_final, cls,
                     (const uint16_t*)index_plane, out_plane);
    } else if (idx_bytes == 4) {
        RHCCQ_LAUNCH(rhccq_k_paint<uint32_t>, n_crops, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream,
                     seg, B, H, W, crops, n_crops, ent_off, ent_final, cls,
                     (const uint32_t*)index_plane, out_plane);
    } else {
        rhccq_set_error("rhccq_paint: idx_bytes must be 2 or 4, got %d", idx_bytes);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- decoder gather and quality metrics
// image = palette[indices] (/root/reference/decoder/uncompression/uncompression.py:209) and the sums behind
// MSE / PSNR / MAE (decoder/uncompression/comparison.py:43-44,64-79): streaming, HBM-bound.
template <class IdxT>
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_decode_gather(const IdxT* __restrict__ idx, long long n, const uint8_t* __restrict__ pal, int n_pal,
                      uint8_t* __restrict__ out, int* __restrict__ bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned v = (unsigned)idx[i];
        if (v >= (unsigned)n_pal) { *bad = 1; continue; }
        out[3 * i] = pal[3 * v]; out[3 * i + 1] = pal[3 * v + 1]; out[3 * i + 2] = pal[3 * v + 2];
    }
}
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_sq_abs_err(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n, long long* __restrict__ acc) {
    __shared__ long long s_ll[RHCCQ_MAX_WARPS + 2];
    long long sq = 0, ab = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)a[i] - (int)b[i];
        sq += d * d; ab += d < 0 ? -d : d;
    }
    sq = rhccq_block_sum<long long>(sq, s_ll);
    ab = rhccq_block_sum<long long>(ab, s_ll);
    if (threadIdx.x == 0) { atomicAdd((unsigned long long*)&acc[0], (unsigned long long)sq); atomicAdd((unsigned long long*)&acc[1], (unsigned long long)ab); }
}

int rhccq_launch_decode_gather(const void* idx, int idx_bytes, long long n, const uint8_t* pal, int n_pal, uint8_t* out,
                               int* bad, void* stream) {
    if (n <= 0) return 0;
    long long blocks = (n + RHCCQ_PIXEL_THREADS - 1) / RHCCQ_PIXEL_THREADS;
    const long long cap = (long long)rhccq_sm_count() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (idx_bytes == 1) RHCCQ_LAUNCH(rhccq_k_decode_gather<uint8_t>, grid, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, (const uint8_t*)idx, n, pal, n_pal, out, bad);
    else if (idx_bytes == 2) RHCCQ_LAUNCH(rhccq_k_decode_gather<uint16_t>, grid, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, (const uint16_t*)idx, n, pal, n_pal, out, bad);
    else if (idx_bytes == 4) RHCCQ_LAUNCH(rhccq_k_decode_gather<uint32_t>, grid, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, (const uint32_t*)idx, n, pal, n_pal, out, bad);
    else { rhccq_set_error("rhccq_decode_gather: idx_bytes must be 1, 2 or 4"); return -1; }
    return 0;
}
int rhccq_launch_sq_abs_err(const uint8_t* a, const uint8_t* b, long long n, long long* acc, void* stream) {
    if (n <= 0) return 0;
    long long blocks = (n + RHCCQ_PIXEL_THREADS * 8 - 1) / (RHCCQ_PIXEL_THREADS * 8);
    const long long cap = (long long)rhccq_sm_count() * 8;
    const int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    RHCCQ_LAUNCH(rhccq_k_sq_abs_err, grid, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, a, b, n, acc);
    return 0;
}
