// C-ABI glue of librhccq.so (include/rhccq.h): argument checks, error text,
// launch wrappers, and the host-side helpers that need no GPU.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "rhccq_common.cuh"
#include "rhccq_kernels.h"
#include "../../include/rhccq.h"

static thread_local char g_err[512] = "";

void rhccq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#ifdef RHCCQ_HOST_EMU
rhccq_emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
unsigned char* rhccq_emu_dyn_smem = nullptr;
static size_t g_emu_smem_cap = 0;
void rhccq_emu_prepare_smem(size_t bytes) {
    if (bytes > g_emu_smem_cap) {
        free(rhccq_emu_dyn_smem);
        rhccq_emu_dyn_smem = (unsigned char*)malloc(bytes + 64);
        g_emu_smem_cap = bytes;
    }
    if (bytes) memset(rhccq_emu_dyn_smem, 0xCD, bytes);           // uninitialised reads show up as garbage
}
int rhccq_smem_optin(const void*, size_t bytes) {
    if (bytes > 227 * 1024) { rhccq_set_error("shared memory request of %zu bytes exceeds 227 KiB", bytes); return -1; }
    return 0;
}
int rhccq_sm_count(void) { return 1; }
static int rhccq_after_launch(const char*) { return 0; }
extern "C" int rhccq_device_check(void) { return 0; }
#else
int rhccq_smem_optin(const void* kernel, size_t bytes) {
    if (bytes <= 8 * 1024) return 0;                               // static + dynamic must stay below 48 KiB without the opt-in
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        rhccq_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu): %s", bytes, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}
int rhccq_sm_count(void) {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}
static int rhccq_after_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        rhccq_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}
extern "C" int rhccq_device_check(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        rhccq_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return -1;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        rhccq_set_error("device %d has compute capability %d.x; librhccq is built for sm_100a only", dev, major);
        return -1;
    }
    return 0;
}
#endif

extern "C" {

int rhccq_abi_version(void) { return RHCCQ_ABI_VERSION; }
const char* rhccq_last_error(void) { return g_err; }

// ---------------------------------------------------------------- MT19937 as numpy.random.RandomState seeds it
int rhccq_kmeans_rng_fill(double* host_out, int count) {
    if (!host_out || count < 0) { rhccq_set_error("rhccq_kmeans_rng_fill: bad arguments"); return -1; }
    uint32_t mt[624];
    mt[0] = 42u;                                                   // init_genrand(42)
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int pos = 624;
    auto next32 = [&]() -> uint32_t {
        if (pos >= 624) {
            for (int k = 0; k < 624; ++k) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    };
    for (int i = 0; i < count; ++i) {                              // random_sample: 53-bit resolution
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        host_out[i] = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    return 0;
}

int rhccq_kmeans_rng_need(int max_rows) {
    // k <= max_rows centres, 2 + int(log(k)) trials each (sklearn/cluster/_kmeans.py:226)
    if (max_rows < 1) return 1;
    const int e[] = {3, 8, 21, 55, 149, 404, 1097, 2981, 8104, 22027, 59875, 162755, 442414, 1202605};
    int t = 2;
    for (int i = 0; i < 14; ++i) if (max_rows >= e[i]) ++t;
    long long need = 1 + (long long)(max_rows - 1) * t;
    return need > 0x7fffffff ? 0x7fffffff : (int)need;
}

size_t rhccq_workspace_total_bytes(size_t need, int n_problems) {
    // shared memory holds it: no global workspace; else one slice per resident CTA
    if (need <= RHCCQ_SMEM_BUDGET || n_problems <= 0) return 0;
    size_t slices = (size_t)rhccq_sm_count() * 4;
    if ((size_t)n_problems < slices) slices = (size_t)n_problems;
    return need * slices;
}
size_t rhccq_unique_index_workspace_bytes(int max_valid) { return rhccq_unique_ws_bytes(max_valid); }
size_t rhccq_palette_dbscan_workspace_bytes(int max_rows, int max_slots) {
    return rhccq_palette_dbscan_ws_bytes(max_rows, max_slots);
}
size_t rhccq_palette_split_workspace_bytes(int max_rows) { return rhccq_palette_split_ws_bytes(max_rows); }
size_t rhccq_palette_finish_workspace_bytes(int max_rows) { return rhccq_palette_finish_ws_bytes(max_rows); }
size_t rhccq_merge_level_workspace_bytes(int max_entries, int max_comps) {
    return rhccq_merge_level_ws_bytes(max_entries, max_comps);
}

#define RHCCQ_REQUIRE(cond, what)                                              \
    do { if (!(cond)) { rhccq_set_error("%s: %s", what, #cond); return -1; } } while (0)

int rhccq_unique_index(const uint8_t* img, const int32_t* seg, int B, int H, int W, const int32_t* crops,
                       int n_crops, const int32_t* pal_off, uint32_t* pal_keys, int32_t* pal_cnt,
                       void* index_plane, int idx_bytes, int repaint_black, int max_valid, void* ws,
                       size_t ws_bytes, void* stream) {
    RHCCQ_REQUIRE(img && crops && pal_off && pal_keys && pal_cnt && index_plane, "rhccq_unique_index");
    RHCCQ_REQUIRE(B > 0 && H > 0 && W > 0 && n_crops >= 0 && max_valid >= 0, "rhccq_unique_index");
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_unique(img, seg, B, H, W, crops, n_crops, pal_off, pal_keys, pal_cnt, index_plane, idx_bytes,
                            repaint_black, max_valid, lw, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_unique_index");
}

int rhccq_palette_dbscan(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                         const int32_t* thr, const int32_t* tie, const double* eps, int n_problems,
                         int32_t* labels, int32_t* n_clusters, int max_rows, int max_slots, void* ws,
                         size_t ws_bytes, void* stream) {
    RHCCQ_REQUIRE(pal_keys && pal_off && pal_cnt && thr && tie && eps && labels && n_clusters, "rhccq_palette_dbscan");
    RHCCQ_REQUIRE(n_problems >= 0 && max_rows >= 1 && max_slots >= 1, "rhccq_palette_dbscan");
    rhccq_palette_batch Bt = {pal_keys, pal_off, pal_cnt, thr, tie, eps, n_problems};
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_palette_dbscan(Bt, labels, n_clusters, max_rows, max_slots, lw, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_palette_dbscan");
}

size_t rhccq_palette_minibatch_workspace_bytes(int max_rows, int n_problems) {
    // status copy + centres of every palette + one working slice per CTA (at most one per SM and per palette)
    if (max_rows < 1 || n_problems < 1) return 0;
    const size_t kmax = (size_t)max_rows / 10 + 2;
    int slices = n_problems < rhccq_sm_count() ? n_problems : rhccq_sm_count();
    return rhccq_carve_bytes((size_t)n_problems + 4 + 256, 4) + rhccq_carve_bytes((size_t)n_problems * 3 * kmax, 8)
           + (size_t)slices * rhccq_palette_minibatch_ws_bytes(max_rows);
}

int rhccq_palette_minibatch(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                            const double* quality, int n_problems, int32_t* labels, int32_t* n_clusters, int max_rows,
                            void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_REQUIRE(pal_keys && pal_off && pal_cnt && quality && labels && n_clusters && ws, "rhccq_palette_minibatch");
    RHCCQ_REQUIRE(n_problems >= 0 && max_rows >= 1, "rhccq_palette_minibatch");
    rhccq_palette_batch Bt = {pal_keys, pal_off, pal_cnt, nullptr, nullptr, nullptr, n_problems};
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_palette_minibatch(Bt, quality, labels, n_clusters, max_rows, lw, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_palette_minibatch");
}

int rhccq_palette_split(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt, int n_problems,
                        const int32_t* labels, const int32_t* n_clusters, const int32_t* max_cpc, const double* rng,
                        int rng_len, int32_t* leaf, int32_t* n_leaves, int max_rows, void* ws, size_t ws_bytes,
                        void* stream) {
    RHCCQ_REQUIRE(pal_keys && pal_off && pal_cnt && labels && max_cpc && rng && leaf && n_leaves, "rhccq_palette_split");
    RHCCQ_REQUIRE(n_problems >= 0 && max_rows >= 1 && rng_len >= 1, "rhccq_palette_split");
    rhccq_palette_batch Bt = {pal_keys, pal_off, pal_cnt, nullptr, nullptr, nullptr, n_problems};
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_palette_split(Bt, labels, n_clusters, max_cpc, rng, rng_len, leaf, n_leaves, max_rows, lw,
                                   stream) != 0)
        return -1;
    return rhccq_after_launch("rhccq_palette_split");
}

int rhccq_palette_finish(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt, int n_problems,
                         const int32_t* leaf, const int32_t* n_leaves, uint32_t* new_keys, int max_rows, void* ws,
                         size_t ws_bytes, void* stream) {
    RHCCQ_REQUIRE(pal_keys && pal_off && pal_cnt && leaf && n_leaves && new_keys, "rhccq_palette_finish");
    RHCCQ_REQUIRE(n_problems >= 0 && max_rows >= 1, "rhccq_palette_finish");
    rhccq_palette_batch Bt = {pal_keys, pal_off, pal_cnt, nullptr, nullptr, nullptr, n_problems};
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_palette_finish(Bt, leaf, n_leaves, new_keys, max_rows, lw, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_palette_finish");
}

int rhccq_remap_first(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                      const int32_t* pal_off, const int32_t* leaf, const int32_t* n_leaves,
                      const uint32_t* new_keys, const int32_t* ent_off, void* index_plane, int idx_bytes,
                      uint32_t* ent_color, uint32_t* ent_fpos, int max_leaves, void* stream) {
    RHCCQ_REQUIRE(crops && pal_off && leaf && n_leaves && new_keys && ent_off && index_plane && ent_color && ent_fpos,
                  "rhccq_remap_first");
    RHCCQ_REQUIRE(B > 0 && H > 0 && W > 0 && n_crops >= 0 && max_leaves >= 1, "rhccq_remap_first");
    RHCCQ_REQUIRE((long long)H * W < 0xFFFFFFFFll, "rhccq_remap_first");
    if (rhccq_launch_remap_first(seg, B, H, W, crops, n_crops, pal_off, leaf, n_leaves, new_keys, ent_off, index_plane,
                                 idx_bytes, ent_color, ent_fpos, max_leaves, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_remap_first");
}

int rhccq_merge_level(const uint32_t* color_in, const uint32_t* fpos_in, const int32_t* comp_start,
                      const int32_t* comp_cnt, const int32_t* grp_comp_off, int n_groups, uint32_t* color_out,
                      uint32_t* fpos_out, int32_t* out_off, int32_t* out_cnt, int32_t* out_present, int32_t* map,
                      int max_entries, int max_comps, void* ws, size_t ws_bytes, void* stream) {
    RHCCQ_REQUIRE(color_in && fpos_in && comp_start && comp_cnt && grp_comp_off && color_out && fpos_out && out_off
                  && out_cnt && out_present && map, "rhccq_merge_level");
    RHCCQ_REQUIRE(n_groups >= 0 && max_entries >= 1 && max_comps >= 1, "rhccq_merge_level");
    rhccq_merge_args M = {color_in, fpos_in, comp_start, comp_cnt, grp_comp_off, n_groups,
                           color_out, fpos_out, out_off, out_cnt, out_present, map};
    rhccq_launch_ws lw = {(unsigned char*)ws, ws_bytes};
    if (rhccq_launch_merge_level(M, max_entries, max_comps, lw, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_merge_level");
}

int rhccq_first_min(const int32_t* off, const int32_t* cnt, const int32_t* n_leaves, int n_groups,
                    const int32_t* leaf, const uint32_t* fpos_in, uint32_t* fpos_out, void* stream) {
    RHCCQ_REQUIRE(off && cnt && n_leaves && leaf && fpos_in && fpos_out && n_groups >= 0, "rhccq_first_min");
    if (rhccq_launch_first_min(off, cnt, n_leaves, n_groups, leaf, fpos_in, fpos_out, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_first_min");
}

int rhccq_compose_final(int n_segments, const int32_t* n_leaves1, const int32_t* ent_off0,
                        const int32_t* seg_region, const int32_t* region_group, const int32_t* group_image,
                        const int32_t* offA, const int32_t* mapA, const int32_t* offB, const int32_t* mapB,
                        const int32_t* leaf2, const uint32_t* color2, const int32_t* offC, const int32_t* mapC,
                        const int32_t* leaf3, const int32_t* presentC, int32_t* ent_final, void* stream) {
    RHCCQ_REQUIRE(n_leaves1 && ent_off0 && seg_region && region_group && group_image && offA && mapA && offB && mapB
                  && leaf2 && color2 && offC && mapC && leaf3 && presentC && ent_final && n_segments >= 0,
                  "rhccq_compose_final");
    rhccq_compose C = {n_segments, n_leaves1, ent_off0, seg_region, region_group, group_image, offA, mapA,
                       offB, mapB, leaf2, color2, offC, mapC, leaf3, presentC, ent_final};
    if (rhccq_launch_compose_final(C, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_compose_final");
}

int rhccq_paint(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops, const int32_t* ent_off,
                const int32_t* ent_final, int cls, const void* index_plane, int idx_bytes,
                uint16_t* out_plane, void* stream) {
    RHCCQ_REQUIRE(crops && ent_off && ent_final && index_plane && out_plane, "rhccq_paint");
    RHCCQ_REQUIRE(B > 0 && H > 0 && W > 0 && n_crops >= 0, "rhccq_paint");
    if (rhccq_launch_paint(seg, B, H, W, crops, n_crops, ent_off, ent_final, cls, index_plane, idx_bytes,
                           out_plane, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_paint");
}

int rhccq_comp_pass(const int32_t* comps, int n_comps, const int32_t* indices, int Hc, int Wc, int mode,
                    const int32_t* map, uint32_t* fpos, int32_t* prio, int32_t* canvas, void* stream) {
    RHCCQ_REQUIRE(comps && indices && n_comps >= 0 && Hc > 0 && Wc > 0 && mode >= 0 && mode <= 2, "rhccq_comp_pass");
    RHCCQ_REQUIRE(mode == 0 ? fpos != nullptr : (map != nullptr && prio != nullptr), "rhccq_comp_pass");
    RHCCQ_REQUIRE(mode != 2 || canvas != nullptr, "rhccq_comp_pass");
    if (rhccq_launch_comp_pass(comps, n_comps, indices, Hc, Wc, mode, map, fpos, prio, canvas, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_comp_pass");
}

int rhccq_decode_gather(const void* indices, int idx_bytes, long long n, const uint8_t* palette, int n_palette,
                        uint8_t* out_rgb, int32_t* bad, void* stream) {
    RHCCQ_REQUIRE(indices && palette && out_rgb && bad && n >= 0 && n_palette >= 0, "rhccq_decode_gather");
    if (rhccq_launch_decode_gather(indices, idx_bytes, n, palette, n_palette, out_rgb, bad, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_decode_gather");
}

int rhccq_sq_abs_err(const uint8_t* a, const uint8_t* b, long long n, long long* acc2, void* stream) {
    RHCCQ_REQUIRE(a && b && acc2 && n >= 0, "rhccq_sq_abs_err");
    if (rhccq_launch_sq_abs_err(a, b, n, acc2, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_sq_abs_err");
}

int rhccq_excl_scan(const int32_t* in, int n, int32_t* out, void* stream) {
    RHCCQ_REQUIRE(in && out && n >= 0, "rhccq_excl_scan");
    if (rhccq_launch_excl_scan(in, n, out, stream) != 0) return -1;
    return rhccq_after_launch("rhccq_excl_scan");
}

}  // extern "C"

// ---------------------------------------------------------------- a2 on the device
// max_cpc = ceil((-(q / 100) * n + n) / q), 0 -> 1: the Python expression of
// /root/reference/encoder/compression/clustering.py:129-133, evaluated in IEEE
// double with the same operation order.
__global__ void rhccq_k_cluster_params(const int* __restrict__ n_colors, const double* __restrict__ quality, int n,
                                       int* __restrict__ max_cpc) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double q = quality[i];
        const int nc = n_colors[i];
        int m = 1;
        if (nc > 0 && q > 0.0) {
            const double t = __ddiv_rn(q, 100.0);
            const double u = __dadd_rn(__dmul_rn(-t, (double)nc), (double)nc);
            const double v = ceil(__ddiv_rn(u, q));
            m = v < 1.0 ? 1 : (v > 2147483647.0 ? 2147483647 : (int)v);
        }
        max_cpc[i] = m;
    }
}

extern "C" int rhccq_cluster_params(const int32_t* n_colors, const double* quality, int n, int32_t* max_cpc,
                                    void* stream) {
    RHCCQ_REQUIRE(n_colors && quality && max_cpc && n >= 0, "rhccq_cluster_params");
    if (n == 0) return 0;
    const int grid = (n + 255) / 256;
    RHCCQ_LAUNCH(rhccq_k_cluster_params, grid, 256, 0, (cudaStream_t)stream, n_colors, quality, n, max_cpc);
    return rhccq_after_launch("rhccq_cluster_params");
}
