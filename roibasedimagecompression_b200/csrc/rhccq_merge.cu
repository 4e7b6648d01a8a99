// Entry-level merge of component palettes onto one canvas palette:
// /root/reference/encoder/compression/merging.py:8-120 restated on "entries"
// (component, palette row) instead of pixels.
//
// The reference paints components in reversed list order, walks each one in
// raster order, and gives a colour the next free palette slot the first time
// it meets it; black never paints.  A colour's slot is therefore the rank of
//     okey = (reversed rank of the first component that shows it) << 32
//            | (first raster position inside that component)
// among all colours of the group, plus one for the leading black.  The first
// raster position of every entry is produced by rhccq_k_remap_first, so the
// merge touches entries (thousands), not pixels (millions).  A group with a
// single component is passed through unchanged (merging.py:16-21).
//
// Components of one group must not overlap in pixels when the merged result is
// merged again (stage 1 -> stage 2): first positions are propagated as minima
// over the members.  That holds for segments of a region and regions of a
// class (disjoint masks); the last level (ROI over non-ROI) may overlap and is
// resolved per pixel by rhccq_k_paint's launch order.
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"

#define RHCCQ_EMPTY_KEY 0xFFFFFFFFu
#define RHCCQ_MERGE_COUNT_CAP 1024     // colours one component may own for the counting rank (else: sort)

__host__ __device__ static inline size_t rhccq_pow2_sz(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }

__host__ __device__ static inline size_t rhccq_merge_hcap(int max_entries) {
    return rhccq_pow2_sz((size_t)max_entries + (size_t)max_entries / 2 + 2);
}

size_t rhccq_merge_level_ws_bytes(int max_entries, int max_comps) {
    const size_t hcap = rhccq_merge_hcap(max_entries);
    return rhccq_carve_bytes(hcap, 4) * 2 + rhccq_carve_bytes(hcap, 8) + rhccq_carve_bytes(hcap, 4)
           + rhccq_carve_bytes(rhccq_pow2_sz(max_entries > 1 ? max_entries : 1), 8)
           + rhccq_carve_bytes((size_t)max_comps + 1, 4) * 2;
}

__device__ __forceinline__ int rhccq_lower_bound_u64(const unsigned long long* a, int n, unsigned long long key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ void rhccq_merge_level_group(const rhccq_merge_args& M, int g, int max_entries, int max_comps,
                                        unsigned char* wsbase) {
    __shared__ int s_scr[RHCCQ_MAX_WARPS + 2];
    __shared__ int s_cnt, s_single;
    const int c0 = M.grp_comp_off[g], c1 = M.grp_comp_off[g + 1];
    const int ncomp = c1 - c0;
    const int out0 = M.comp_start[c0] + g;                         // capacity = sum of the components' + 1
    if (threadIdx.x == 0) M.out_off[g] = out0;
    if (ncomp > max_comps) {
        if (threadIdx.x == 0) { M.out_cnt[g] = -1; M.out_present[g] = 0; }
        return;
    }
    const size_t hcap = rhccq_merge_hcap(max_entries);
    rhccq_carver cv(wsbase);
    uint32_t* hcol = cv.take<uint32_t>(hcap);
    uint32_t* hpos = cv.take<uint32_t>(hcap);
    unsigned long long* hkey = cv.take<unsigned long long>(hcap);
    uint32_t* hrank = cv.take<uint32_t>(hcap);                       // slot -> rank of its colour
    unsigned long long* sortbuf = cv.take<unsigned long long>(rhccq_pow2_sz(max_entries > 1 ? max_entries : 1));
    int* prank = cv.take<int>((size_t)max_comps + 1);
    int* ownc = cv.take<int>((size_t)max_comps + 1);

    // components that hold at least one entry take part (regions.py:18-29 drops empty ones)
    if (threadIdx.x == 0) s_single = -1;
    RHCCQ_PAR_FOR(i, ncomp) prank[i] = M.comp_cnt[c0 + i] > 0 ? 1 : 0;
    __syncthreads();
    const int n_present = rhccq_block_excl_scan_array<int>(prank, ncomp, s_scr);
    int n_entries = 0, bad = 0;
    RHCCQ_PAR_FOR(i, ncomp) {
        const int c = M.comp_cnt[c0 + i];
        if (c < 0) bad = 1;
        if (c > 0) { n_entries += c; if (n_present == 1) s_single = c0 + i; }
    }
    n_entries = rhccq_block_sum<int>(n_entries, s_scr);
    bad = rhccq_block_or(bad, s_scr);
    if (bad || n_entries > max_entries) {
        if (threadIdx.x == 0) { M.out_cnt[g] = bad ? -2 : -1; M.out_present[g] = 0; }
        return;
    }
    if (threadIdx.x == 0) M.out_present[g] = n_present;
    if (n_present == 0) {
        if (threadIdx.x == 0) M.out_cnt[g] = 0;
        return;
    }
    if (n_present == 1) {                                          // pass-through
        const int c = s_single;
        const int st = M.comp_start[c], cn = M.comp_cnt[c];
        RHCCQ_PAR_FOR(j, cn) {
            M.color_out[out0 + j] = M.color_in[st + j];
            M.fpos_out[out0 + j] = M.fpos_in[st + j];
            M.map[st + j] = j;
        }
        if (threadIdx.x == 0) M.out_cnt[g] = cn;
        return;
    }
    // colour -> (smallest okey, smallest raster position)
    const uint32_t hmask = (uint32_t)(hcap - 1);
    for (size_t s = threadIdx.x; s < hcap; s += blockDim.x) { hcol[s] = RHCCQ_EMPTY_KEY; hpos[s] = 0xFFFFFFFFu; hkey[s] = ~0ull; }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = RHCCQ_WARP; i < ncomp; i += RHCCQ_NWARPS) {
        const int st = M.comp_start[c0 + i], cn = M.comp_cnt[c0 + i];
        const unsigned long long hi = (unsigned long long)(n_present - 1 - prank[i]) << 32;
        for (int j = RHCCQ_LANE; j < cn; j += RHCCQ_WARP_SIZE) {
            const uint32_t col = M.color_in[st + j], fp = M.fpos_in[st + j];
            if (col == 0u || fp == 0xFFFFFFFFu) continue;          // black never paints; unused rows never appear
            uint32_t h = (col * 2654435761u) & hmask;
            while (true) {
                const uint32_t k = atomicCAS(&hcol[h], RHCCQ_EMPTY_KEY, col);
                if (k == RHCCQ_EMPTY_KEY || k == col) break;
                h = (h + 1) & hmask;
            }
            atomicMin(&hkey[h], hi | fp);
            atomicMin(&hpos[h], fp);
        }
    }
    __syncthreads();
    // Rank of every colour's okey among the group's.  The okeys order first by component (last listed first), then by
    // raster position inside it, and every colour has one owner entry — the one whose (component, position) is the
    // colour's okey.  So: owners per component (one warp per component), a scan over the components in paint order,
    // and the rank of an owner among its component's owners by counting.  A component that owns more than
    // RHCCQ_MERGE_COUNT_CAP colours makes the group take the general way instead: sort all okeys.
    int over = 0;
    for (int i = RHCCQ_WARP; i < ncomp; i += RHCCQ_NWARPS) {
        const int st = M.comp_start[c0 + i], cn = M.comp_cnt[c0 + i];
        const unsigned long long hi = (unsigned long long)(n_present - 1 - prank[i]) << 32;
        int owned = 0;
        for (int j0 = 0; j0 < cn; j0 += RHCCQ_WARP_SIZE) {
            const int j = j0 + RHCCQ_LANE;
            bool owner = false;
            if (j < cn) {
                const uint32_t col = M.color_in[st + j], fp = M.fpos_in[st + j];
                if (col != 0u && fp != 0xFFFFFFFFu) {
                    uint32_t h = (col * 2654435761u) & hmask;
                    while (hcol[h] != col) h = (h + 1) & hmask;
                    owner = hkey[h] == (hi | fp);
                }
            }
            owned += __popc(rhccq_ballot(owner));
        }
        if (RHCCQ_LANE == 0) ownc[i] = owned;
        over |= owned > RHCCQ_MERGE_COUNT_CAP;
    }
    over = rhccq_block_or(over, s_scr);
    int U;
    if (!over) {
        U = rhccq_block_excl_scan_array<int>(ownc, ncomp, s_scr);    // ownc[i] <- owners in the components listed before i
        if (threadIdx.x == 0) ownc[ncomp] = U;
        __syncthreads();
        for (int i = RHCCQ_WARP; i < ncomp; i += RHCCQ_NWARPS) {
            const int st = M.comp_start[c0 + i], cn = M.comp_cnt[c0 + i];
            const int owned = ownc[i + 1] - ownc[i];
            if (owned == 0) continue;
            const int base = U - ownc[i + 1];                         // owners in the components painted before this one
            const unsigned long long hi = (unsigned long long)(n_present - 1 - prank[i]) << 32;
            unsigned long long* seg = sortbuf + base;                 // (position << 32) | table slot of this component's owners
            int filled = 0;
            for (int j0 = 0; j0 < cn; j0 += RHCCQ_WARP_SIZE) {
                const int j = j0 + RHCCQ_LANE;
                bool owner = false;
                uint32_t h = 0, fp = 0;
                if (j < cn) {
                    const uint32_t col = M.color_in[st + j];
                    fp = M.fpos_in[st + j];
                    if (col != 0u && fp != 0xFFFFFFFFu) {
                        h = (col * 2654435761u) & hmask;
                        while (hcol[h] != col) h = (h + 1) & hmask;
                        owner = hkey[h] == (hi | fp);
                    }
                }
                const unsigned m = rhccq_ballot(owner);
                if (owner) seg[filled + __popc(m & rhccq_lanemask_lt())] = ((unsigned long long)fp << 32) | h;
                filled += __popc(m);
            }
#ifndef RHCCQ_HOST_EMU
            __syncwarp();
#endif
            for (int e = RHCCQ_LANE; e < owned; e += RHCCQ_WARP_SIZE) {
                const unsigned long long mine = seg[e];
                int rank = base;
                for (int o = 0; o < owned; ++o) rank += seg[o] < mine ? 1 : 0;      // (positions of one component are distinct)
                const uint32_t h = (uint32_t)mine;
                M.color_out[out0 + 1 + rank] = hcol[h];
                M.fpos_out[out0 + 1 + rank] = hpos[h];
                hrank[h] = (uint32_t)rank;
            }
        }
    } else {
        for (size_t s = threadIdx.x; s < hcap; s += blockDim.x)
            if (hcol[s] != RHCCQ_EMPTY_KEY) sortbuf[atomicAdd(&s_cnt, 1)] = hkey[s];
        __syncthreads();
        U = s_cnt;
        const int np2 = rhccq_next_pow2(U > 1 ? U : 1);
        for (int j = U + (int)threadIdx.x; j < np2; j += (int)blockDim.x) sortbuf[j] = ~0ull;
        __syncthreads();
        rhccq_block_bitonic_sort<unsigned long long>(sortbuf, np2);
        for (size_t s = threadIdx.x; s < hcap; s += blockDim.x) {
            if (hcol[s] == RHCCQ_EMPTY_KEY) continue;
            const int rank = rhccq_lower_bound_u64(sortbuf, U, hkey[s]);
            M.color_out[out0 + 1 + rank] = hcol[s];
            M.fpos_out[out0 + 1 + rank] = hpos[s];
            hrank[s] = (uint32_t)rank;
        }
    }
    if (threadIdx.x == 0) {
        M.color_out[out0] = 0u;                                    // merging.py:42-44
        M.fpos_out[out0] = 0xFFFFFFFFu;
        M.out_cnt[g] = U + 1;
    }
    __syncthreads();
    for (int i = RHCCQ_WARP; i < ncomp; i += RHCCQ_NWARPS) {
        const int st = M.comp_start[c0 + i], cn = M.comp_cnt[c0 + i];
        for (int j = RHCCQ_LANE; j < cn; j += RHCCQ_WARP_SIZE) {
            const uint32_t col = M.color_in[st + j], fp = M.fpos_in[st + j];
            int out = 0;
            if (col != 0u && fp != 0xFFFFFFFFu) {
                uint32_t h = (col * 2654435761u) & hmask;
                while (hcol[h] != col) h = (h + 1) & hmask;
                out = 1 + (int)hrank[h];
            }
            M.map[st + j] = out;
        }
    }
}

__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_merge_level(rhccq_merge_args M, int max_entries, int max_comps, unsigned char* gws, size_t gws_stride) {
    RHCCQ_DYN_SMEM(dyn);
    unsigned char* wsbase = gws ? gws + (size_t)blockIdx.x * gws_stride : dyn;
    if (blockIdx.x == 0 && threadIdx.x == 0)                       // sentinel for the next level's comp_start
        M.out_off[M.n_groups] = M.comp_start[M.grp_comp_off[M.n_groups]] + M.n_groups;
    for (int g = blockIdx.x; g < M.n_groups; g += gridDim.x) {
        rhccq_merge_level_group(M, g, max_entries, max_comps, wsbase);
        __syncthreads();
    }
}

int rhccq_launch_merge_level(const rhccq_merge_args& M, int max_entries, int max_comps, rhccq_launch_ws ws,
                             void* stream) {
    if (M.n_groups <= 0) return 0;
    const size_t need = rhccq_merge_level_ws_bytes(max_entries, max_comps);
    size_t smem; unsigned char* gws;
    const int grid = rhccq_pick_grid((const void*)rhccq_k_merge_level, need, M.n_groups, ws, &smem, &gws,
                                     "rhccq_merge_level");
    if (grid < 0) return -1;
    RHCCQ_LAUNCH(rhccq_k_merge_level, grid, RHCCQ_PALETTE_THREADS, smem, (cudaStream_t)stream,
                 M, max_entries, max_comps, gws, need);
    return 0;
}

// ---------------------------------------------------------------- after clustering a merged palette
// Entry j of group g (at off[g] + j, j < cnt[g]) moved to row leaf[off[g] + j]
// of the clustered palette; the clustered row first appears where the earliest
// of its members did.
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_first_min(const int* __restrict__ off, const int* __restrict__ cnt, const int* __restrict__ n_leaves,
                  int n_groups, const int* __restrict__ leaf, const uint32_t* __restrict__ fpos_in,
                  uint32_t* __restrict__ fpos_out) {
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int o = off[g], n = cnt[g], m = n_leaves[g];
        if (n < 0 || m < 0) continue;
        RHCCQ_PAR_FOR(j, m) fpos_out[o + j] = 0xFFFFFFFFu;
        __syncthreads();
        RHCCQ_PAR_FOR(j, n) {
            const uint32_t fp = fpos_in[o + j];
            if (fp != 0xFFFFFFFFu) atomicMin(&fpos_out[o + (leaf[o + j] & 0xffff)], fp);
        }
        __syncthreads();
    }
}

int rhccq_launch_first_min(const int* off, const int* cnt, const int* n_leaves, int n_groups, const int* leaf,
                           const uint32_t* fpos_in, uint32_t* fpos_out, void* stream) {
    if (n_groups <= 0) return 0;
    RHCCQ_LAUNCH(rhccq_k_first_min, n_groups, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream,
                 off, cnt, n_leaves, n_groups, leaf, fpos_in, fpos_out);
    return 0;
}

// ---------------------------------------------------------------- composed entry -> final index
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_compose_final(rhccq_compose C) {
    for (int p = blockIdx.x; p < C.n_segments; p += gridDim.x) {
        const int m = C.n_leaves1[p];
        if (m < 0) continue;
        const int r = C.seg_region[p], g = C.region_group[r], b = C.group_image[g];
        const int e0 = C.ent_off0[p], oA = C.offA[r], oB = C.offB[g], oC = C.offC[b];
        const bool merged_top = C.presentC[b] >= 2;
        RHCCQ_PAR_FOR(v, m) {
            const int a = C.mapA[e0 + v];
            const int bb = C.mapB[oA + a];
            const int j2 = C.leaf2[oB + bb] & 0xffff;
            const int cc = C.mapC[oB + j2];
            const int fin = C.leaf3[oC + cc] & 0xffff;
            const bool paints = !merged_top || C.color2[oB + j2] != 0u;      // merging.py:76
            C.ent_final[e0 + v] = paints ? fin : -1;
        }
    }
}

int rhccq_launch_compose_final(const rhccq_compose& C, void* stream) {
    if (C.n_segments <= 0) return 0;
    RHCCQ_LAUNCH(rhccq_k_compose_final, C.n_segments, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream, C);
    return 0;
}

// ---------------------------------------------------------------- exclusive scan of a count vector
// out[i] = sum_{j<i} max(in[j], 0), out[n] = total.  One CTA; n is the number of
// segments of a batch (tens of thousands), so this is launch-latency sized.
__global__ void __launch_bounds__(RHCCQ_PALETTE_THREADS)
rhccq_k_excl_scan(const int* __restrict__ in, int n, int* __restrict__ out) {
    __shared__ int s_scr[RHCCQ_MAX_WARPS + 2];
    RHCCQ_PAR_FOR(i, n) out[i] = in[i] > 0 ? in[i] : 0;
    __syncthreads();
    const int total = rhccq_block_excl_scan_array<int>(out, n, s_scr);
    if (threadIdx.x == 0) out[n] = total;
}

int rhccq_launch_excl_scan(const int* in, int n, int* out, void* stream) {
    RHCCQ_LAUNCH(rhccq_k_excl_scan, 1, RHCCQ_PALETTE_THREADS, 0, (cudaStream_t)stream, in, n, out);
    return 0;
}

// ---------------------------------------------------------------- operator-level merge of arbitrary components
// merge_region_components_simple on component dicts (palette + flat indices +
// top-left + shape), which may overlap and may stick out of the canvas
// (merging.py:52-82).  comps: int32 [n,8] = (pixel offset, h, w, row0, col0
// relative to the canvas, palette offset, palette rows, list position).
__global__ void __launch_bounds__(RHCCQ_PIXEL_THREADS)
rhccq_k_comp_pass(const int32_t* __restrict__ comps, int n_comps, const int32_t* __restrict__ indices,
                  int Hc, int Wc, int mode, const int* __restrict__ map, uint32_t* __restrict__ fpos,
                  int* __restrict__ prio, int32_t* __restrict__ canvas) {
    for (int c = blockIdx.x; c < n_comps; c += gridDim.x) {
        const int32_t* cm = comps + 8 * (size_t)c;
        const int poff = cm[0], h = cm[1], w = cm[2], r0 = cm[3], c0 = cm[4], pal = cm[5], cnt = cm[6], rank = cm[7];
        RHCCQ_PAR_FOR(q, h * w) {
            const int r = r0 + q / w, cc = c0 + q % w;
            if (r < 0 || r >= Hc || cc < 0 || cc >= Wc) continue;          // merging.py:69
            const int idx = indices[poff + q];
            if (idx < 0 || idx >= cnt) continue;                            // merging.py:71
            const int pos = r * Wc + cc;
            if (mode == 0) {
                atomicMin(&fpos[pal + idx], (uint32_t)pos);
            } else {
                const int m = map[pal + idx];
                if (m == 0) continue;                                       // black does not paint (:76)
                if (mode == 1) atomicMin(&prio[pos], rank);                 // first listed wins (:52)
                else if (prio[pos] == rank) canvas[pos] = m;
            }
        }
    }
}

int rhccq_launch_comp_pass(const int32_t* comps, int n_comps, const int32_t* indices, int Hc, int Wc, int mode,
                           const int* map, uint32_t* fpos, int* prio, int32_t* canvas, void* stream) {
    if (n_comps <= 0) return 0;
    RHCCQ_LAUNCH(rhccq_k_comp_pass, n_comps, RHCCQ_PIXEL_THREADS, 0, (cudaStream_t)stream,
                 comps, n_comps, indices, Hc, Wc, mode, map, fpos, prio, canvas);
    return 0;
}
