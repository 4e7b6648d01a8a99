// Internal declarations shared by the kernel translation units and the C-ABI
// glue (rhccq_api.cu).  Not installed; the public interface is include/rhccq.h.
#pragma once
#include <stdint.h>
#include <stddef.h>

#define RHCCQ_PALETTE_THREADS 512
#define RHCCQ_PIXEL_THREADS 256
#define RHCCQ_SMEM_BUDGET (200 * 1024)

// A batch of palettes: problem p owns rows pal_off[p] .. pal_off[p] + pal_cnt[p]
// of every per-row array.  thr/tie/eps describe the radius predicate on the
// 0..255 scale: accept d2 <= thr, except that with tie != 0 a pair at
// d2 == thr is evaluated in float64 as scikit-learn does (eps is the radius).
struct rhccq_palette_batch {
    const uint32_t* pal_keys;
    const int* pal_off;
    const int* pal_cnt;
    const int* thr;
    const int* tie;
    const double* eps;
    int n_problems;
};

// One level of the merge hierarchy (segments -> region, regions -> class
// canvas, classes -> image).  Component c owns entries comp_start[c] ..
// comp_start[c] + comp_cnt[c] of (color_in, fpos_in); group g owns components
// grp_comp_off[g] .. grp_comp_off[g + 1].  comp_start has one sentinel element.
struct rhccq_merge_args {
    const uint32_t* color_in;
    const uint32_t* fpos_in;
    const int* comp_start;
    const int* comp_cnt;
    const int* grp_comp_off;
    int n_groups;
    uint32_t* color_out;
    uint32_t* fpos_out;
    int* out_off;        // [n_groups + 1]
    int* out_cnt;        // [n_groups]; < 0 reports an error for the group
    int* out_present;    // [n_groups] components that took part (1 = passed through)
    int* map;            // per input entry: row of the group's merged palette
};

struct rhccq_compose {
    int n_segments;
    const int* n_leaves1; const int* ent_off0;
    const int* seg_region; const int* region_group; const int* group_image;
    const int* offA; const int* mapA;
    const int* offB; const int* mapB; const int* leaf2; const uint32_t* color2;
    const int* offC; const int* mapC; const int* leaf3; const int* presentC;
    int* ent_final;
};

struct rhccq_launch_ws {
    unsigned char* ws;
    size_t ws_bytes;
};

size_t rhccq_palette_dbscan_ws_bytes(int max_rows, int max_slots);
extern "C" int rhccq_palette_dbscan_slots(int thr, int n_rows);
size_t rhccq_palette_split_ws_bytes(int max_rows);
size_t rhccq_palette_finish_ws_bytes(int max_rows);
size_t rhccq_unique_ws_bytes(int max_valid);
size_t rhccq_merge_level_ws_bytes(int max_entries, int max_comps);

int rhccq_launch_palette_dbscan(const rhccq_palette_batch& B, int* labels, int* n_clusters, int max_rows,
                                int max_slots, rhccq_launch_ws ws, void* stream);
size_t rhccq_palette_minibatch_ws_bytes(int max_rows);
int rhccq_launch_palette_minibatch(const rhccq_palette_batch& B, const double* quality, int* labels, int* n_clusters,
                                   int max_rows, rhccq_launch_ws ws, void* stream);
int rhccq_launch_palette_split(const rhccq_palette_batch& B, const int* labels, const int* status_in, const int* max_cpc,
                               const double* rng, int rng_len, int* leaf, int* n_leaves, int max_rows,
                               rhccq_launch_ws ws, void* stream);
int rhccq_launch_palette_finish(const rhccq_palette_batch& B, const int* leaf, const int* n_leaves,
                                uint32_t* new_keys, int max_rows, rhccq_launch_ws ws, void* stream);
int rhccq_launch_unique(const uint8_t* img, const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                        const int* pal_off, uint32_t* pal_keys, int* pal_cnt, void* index_plane, int idx_bytes,
                        int repaint_black, int max_valid, rhccq_launch_ws ws, void* stream);
int rhccq_launch_remap_first(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                             const int* pal_off, const int* leaf, const int* n_leaves, const uint32_t* new_keys,
                             const int* ent_off, void* index_plane, int idx_bytes, uint32_t* ent_color,
                             uint32_t* ent_fpos, int max_leaves, void* stream);
int rhccq_launch_paint(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops, const int* ent_off,
                       const int* ent_final, int cls, const void* index_plane, int idx_bytes,
                       uint16_t* out_plane, void* stream);
int rhccq_launch_merge_level(const rhccq_merge_args& M, int max_entries, int max_comps, rhccq_launch_ws ws,
                             void* stream);
int rhccq_launch_first_min(const int* off, const int* cnt, const int* n_leaves, int n_groups, const int* leaf,
                           const uint32_t* fpos_in, uint32_t* fpos_out, void* stream);
int rhccq_launch_compose_final(const rhccq_compose& C, void* stream);
int rhccq_launch_comp_pass(const int32_t* comps, int n_comps, const int32_t* indices, int Hc, int Wc, int mode,
                           const int* map, uint32_t* fpos, int* prio, int32_t* canvas, void* stream);
int rhccq_launch_decode_gather(const void* idx, int idx_bytes, long long n, const uint8_t* pal, int n_pal, uint8_t* out,
                               int* bad, void* stream);
int rhccq_launch_sq_abs_err(const uint8_t* a, const uint8_t* b, long long n, long long* acc, void* stream);
int rhccq_launch_excl_scan(const int* in, int n, int* out, void* stream);

// Shared-memory or global-workspace placement of a per-CTA working set of
// `need` bytes; returns the grid size or -1 (error text set).
int rhccq_pick_grid(const void* kernel, size_t need, int n_problems, rhccq_launch_ws ws, size_t* smem,
                    unsigned char** gws, const char* what);
void rhccq_set_error(const char* fmt, ...);
int rhccq_smem_optin(const void* kernel, size_t bytes);
int rhccq_sm_count(void);
