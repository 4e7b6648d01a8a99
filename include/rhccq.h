/* rhccq.h — C ABI of the B200 hot path of the RHCCQ encoder.
 *
 * librhccq.so replaces the data-parallel core of the reference's hierarchical
 * palette quantiser (Riccardoalfieri2003/ROIBasedImageCompression).  The
 * reference is pure Python and has no FFI of its own; the functions below are
 * what a ctypes binding inside the reference's encoder/compression modules
 * would call in place of its Python/NumPy/scikit-learn loops.  Each entry
 * names the reference lines it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with host_;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - functions return 0 on success and a negative value on failure;
 *     rhccq_last_error() returns the text for the calling thread;
 *   - nothing is allocated: outputs and workspaces are caller-owned, the
 *     *_workspace_bytes functions say how much a call may use.  A working set
 *     that fits in shared memory ignores the workspace (it may be NULL);
 *   - per-problem result counters double as status: a negative count marks a
 *     problem the kernel refused (capacity exceeded, unsupported branch) — it
 *     never truncates or guesses;
 *   - colours are 24-bit keys R<<16|G<<8|B in uint32 (ascending key order ==
 *     lexicographic RGB order == np.unique(axis=0) order).
 *
 * Batches: problem p owns rows off[p] .. off[p] + cnt[p] of every per-row
 * array of a call.
 */
#ifndef RHCCQ_H
#define RHCCQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RHCCQ_ABI_VERSION 1

/* status codes stored in per-problem counters */
#define RHCCQ_ERR_CAPACITY (-1)      /* more rows/entries than the caller's max_* bound */
#define RHCCQ_ERR_UPSTREAM (-2)      /* an input counter was already negative, or the random table is too short */
#define RHCCQ_ERR_INDEX_WIDTH (-3)   /* more palette rows than the index type can address */
#define RHCCQ_ERR_MINIBATCH (-4)     /* >= 10000 non-black colours: the reference switches to MiniBatchKMeans
                                        (encoder/compression/clustering.py:207-218); rhccq_palette_dbscan marks
                                        the palette with this status and rhccq_palette_minibatch clusters it */

int rhccq_abi_version(void);
const char* rhccq_last_error(void);
/* 0 when a CUDA device of compute capability 10.x is usable by this process. */
int rhccq_device_check(void);

/* ------------------------------------------------------------------ host helpers (no GPU involved) */

/* First `count` outputs of numpy.random.RandomState(42).random_sample — the stream every
 * KMeans(random_state=42) of the reference consumes (clustering.py:751).  host_out: double[count]. */
int rhccq_kmeans_rng_fill(double* host_out, int count);
/* Doubles rhccq_palette_split may read for clusters of up to max_rows colours. */
int rhccq_kmeans_rng_need(int max_rows);
/* Slots of the cell table rhccq_palette_dbscan builds for a threshold and a row count. */
int rhccq_palette_dbscan_slots(int thr, int n_rows);

/* Per-problem working set of each op (bytes), and the size of the global workspace to pass for a
 * batch of n_problems with that working set: 0 when it fits in shared memory. */
size_t rhccq_workspace_total_bytes(size_t need, int n_problems);
size_t rhccq_unique_index_workspace_bytes(int max_valid);
size_t rhccq_palette_dbscan_workspace_bytes(int max_rows, int max_slots);
size_t rhccq_palette_split_workspace_bytes(int max_rows);
size_t rhccq_palette_finish_workspace_bytes(int max_rows);
size_t rhccq_merge_level_workspace_bytes(int max_entries, int max_comps);

/* ------------------------------------------------------------------ a1: unique colours of a segment
 * Replaces get_all_unique_colors (encoder/compression/clustering.py:4-103) together with the crop /
 * mask / black-repaint preparation of encoder/compression/subregions.py:315-421.
 *
 * img   uint8 [B,H,W,3]; seg int32 [K,B,H,W] label maps (one set per class: the ROI and non-ROI
 * calls have their own, encoder/compression/test.py:105-106) or NULL; crops int32 [n_crops,8] =
 * (image, row0, col0, height, width, segment id (0 = whole rectangle), class k, unused); the index
 * plane is [K,B,H,W] like seg.  For crop p the sorted
 * unique colours go to pal_keys[pal_off[p] ..] (capacity height*width), their number to pal_cnt[p],
 * and the palette row of every pixel OF THE SEGMENT to index_plane at the pixel's image position
 * (idx_bytes 2: uint16, 4: uint32; other pixels are left untouched — in the reference's crop they are
 * black, row 0).  repaint_black != 0 applies subregions.py:395-421.  max_valid bounds the number of
 * segment pixels of any crop. */
int rhccq_unique_index(const uint8_t* img, const int32_t* seg, int B, int H, int W,
                       const int32_t* crops, int n_crops, const int32_t* pal_off,
                       uint32_t* pal_keys, int32_t* pal_cnt, void* index_plane, int idx_bytes,
                       int repaint_black, int max_valid, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ a2: clustering parameters
 * compute_clustering_params (clustering.py:108-135) for device-resident colour counts:
 * max_cpc[p] = ceil(((1 - q/100) n + ... ) / q) exactly as the Python expression evaluates, 0 -> 1.
 * quality: double [n] on the device; n_colors: int32 [n] (negative counts give max_cpc 1). */
int rhccq_cluster_params(const int32_t* n_colors, const double* quality, int n, int32_t* max_cpc, void* stream);

/* ------------------------------------------------------------------ a3': DBSCAN(min_samples=1) on palettes
 * Replaces DBSCAN(eps/255, min_samples=1).fit_predict(palette/255.0) at clustering.py:204-205,233-235
 * (scikit-learn: sklearn/cluster/_dbscan.py:397-470).  Rows equal to black take no part (label -2,
 * clustering.py:185-192); the others get the number of their connected component of the eps-graph,
 * components numbered by their lowest row.  Predicate: d2 <= thr[p]; with tie[p] != 0 a pair at
 * d2 == thr[p] is evaluated in float64 on colours/255 against (eps[p]/255)^2 like the KD-tree does.
 * n_clusters[p] receives the number of components.  */
int rhccq_palette_dbscan(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                         const int32_t* thr, const int32_t* tie, const double* eps, int n_problems,
                         int32_t* labels, int32_t* n_clusters, int max_rows, int max_slots,
                         void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ a3, >= 10 000 colours: MiniBatchKMeans
 * Replaces MiniBatchKMeans(n_clusters=ceil(n*q/100/10), batch_size=1000, random_state=42, n_init='auto')
 * .fit_predict(non_black.astype(float)) at clustering.py:207-218 for every palette whose n_clusters[p] is
 * RHCCQ_ERR_MINIBATCH (as rhccq_palette_dbscan leaves it), in the exact arithmetic of
 * oracle/minibatch_restated.py.  labels[row] = cluster of the row (-2 for black rows), n_clusters[p] =
 * number of centres; other palettes are not touched.  quality: double [n_problems] on the device. */
size_t rhccq_palette_minibatch_workspace_bytes(int max_rows, int n_problems);
int rhccq_palette_minibatch(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                            const double* quality, int n_problems, int32_t* labels, int32_t* n_clusters, int max_rows,
                            void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ a3/a4: cluster -> new palette rows
 * Replaces clustering.py:253-355 and split_large_cluster (:720-775): black rows first, clusters of at
 * most max_cpc[p] colours one row each in ascending label order, larger clusters split recursively by
 * KMeans(k, random_state=42, n_init='auto') — scikit-learn's float64 arithmetic, decision for decision,
 * as restated in oracle/kmeans_sklearn.c — leaves in depth-first label order.  leaf[row] = new palette
 * row; n_leaves[p] = new palette size.  max_cpc[p] = -k runs that KMeans(k) once over the whole palette
 * (the operator of clustering.py:751-752 by itself): leaf[row] = rank of the row's label among the
 * non-empty labels.
 * rng: the rhccq_kmeans_rng_fill stream on the device.  n_clusters (may be NULL): the counters of
 * rhccq_palette_dbscan; a negative one is copied to n_leaves[p] and the problem is skipped. */
int rhccq_palette_split(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                        int n_problems, const int32_t* labels, const int32_t* n_clusters,
                        const int32_t* max_cpc, const double* rng, int rng_len, int32_t* leaf,
                        int32_t* n_leaves, int max_rows, void* ws, size_t ws_bytes, void* stream);

/* New colour of every leaf = per-channel truncated mean of its members (clustering.py:305,347).
 * new_keys[pal_off[p] + j], j < n_leaves[p]. */
int rhccq_palette_finish(const uint32_t* pal_keys, const int32_t* pal_off, const int32_t* pal_cnt,
                         int n_problems, const int32_t* leaf, const int32_t* n_leaves,
                         uint32_t* new_keys, int max_rows, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ a3 tail + a5 head: remap and first appearance
 * indices <- LUT[indices] (clustering.py:373-377) for the pixels of every crop, and per new palette
 * row the smallest raster position row*W+col at which it appears (0xFFFFFFFF: nowhere) — the order
 * merge_region_components_simple (encoder/compression/merging.py:52-82) hands out palette slots in.
 * Entries of crop p: ent_color/ent_fpos[ent_off[p] + j], j < n_leaves[p]. */
int rhccq_remap_first(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                      const int32_t* pal_off, const int32_t* leaf, const int32_t* n_leaves,
                      const uint32_t* new_keys, const int32_t* ent_off, void* index_plane, int idx_bytes,
                      uint32_t* ent_color, uint32_t* ent_fpos, int max_leaves, void* stream);

/* ------------------------------------------------------------------ a5: merge components of a canvas
 * merge_region_components_simple (merging.py:8-120) on entries.  Component c owns entries
 * comp_start[c] .. +comp_cnt[c] (comp_start has n_comps + 1 elements); group g owns components
 * grp_comp_off[g] .. grp_comp_off[g+1].  Output palette of group g at out_off[g] = comp_start[first
 * component] + g: black, then the colours in first-appearance order of the reversed paint sequence;
 * a single non-empty component passes through unchanged (:16-21).  map[entry] = its output row;
 * out_present[g] = components that took part. */
int rhccq_merge_level(const uint32_t* color_in, const uint32_t* fpos_in, const int32_t* comp_start,
                      const int32_t* comp_cnt, const int32_t* grp_comp_off, int n_groups,
                      uint32_t* color_out, uint32_t* fpos_out, int32_t* out_off, int32_t* out_cnt,
                      int32_t* out_present, int32_t* map, int max_entries, int max_comps,
                      void* ws, size_t ws_bytes, void* stream);

/* First appearance of clustered rows: fpos_out[off[g] + leaf[off[g]+j]] = min fpos_in[off[g]+j]. */
int rhccq_first_min(const int32_t* off, const int32_t* cnt, const int32_t* n_leaves, int n_groups,
                    const int32_t* leaf, const uint32_t* fpos_in, uint32_t* fpos_out, void* stream);

/* Compose segment entry -> final palette row through the three stages
 * (encoder/compression/test.py:105-142); -1 where the entry does not paint (black, merging.py:76). */
int rhccq_compose_final(int n_segments, const int32_t* n_leaves1, const int32_t* ent_off0,
                        const int32_t* seg_region, const int32_t* region_group, const int32_t* group_image,
                        const int32_t* offA, const int32_t* mapA, const int32_t* offB, const int32_t* mapB,
                        const int32_t* leaf2, const uint32_t* color2, const int32_t* offC, const int32_t* mapC,
                        const int32_t* leaf3, const int32_t* presentC, int32_t* ent_final, void* stream);

/* out_plane[pixel] = ent_final[ent_off[p] + index_plane[pixel]] for the pixels of the crops of class
 * `cls` (cls < 0: all crops) whose entry paints.  uint16 output plane [B,H,W].  Call once per class,
 * last listed class first, so that the first listed one wins overlaps (merging.py:52). */
int rhccq_paint(const int32_t* seg, int B, int H, int W, const int32_t* crops, int n_crops,
                const int32_t* ent_off, const int32_t* ent_final, int cls,
                const void* index_plane, int idx_bytes, uint16_t* out_plane, void* stream);

/* Operator-level a5 on component dicts that may overlap or leave the canvas (merging.py:52-82).
 * comps int32 [n,8] = (pixel offset into indices, h, w, row0, col0 relative to the canvas, palette
 * offset, palette rows, list position).  mode 0: fpos[palette row] = first canvas raster position;
 * mode 1: prio[pixel] = lowest list position among painting components; mode 2: canvas[pixel] =
 * map[palette row] for the component that owns the pixel.  Run 0, rhccq_merge_level, 1, 2. */
int rhccq_comp_pass(const int32_t* comps, int n_comps, const int32_t* indices, int Hc, int Wc, int mode,
                    const int32_t* map, uint32_t* fpos, int32_t* prio, int32_t* canvas, void* stream);

/* ------------------------------------------------------------------ a3' at scale: DBSCAN of a point cloud
 * Replaces sklearn.cluster.DBSCAN(eps, min_samples).fit_predict(X) — the third-party operator the
 * reference calls at encoder/compression/clustering.py:233-235 (semantics: sklearn/cluster/_dbscan.py:
 * 397-470, _dbscan_inner.pyx) — for float32 X [n, dims] with dims <= 6, e.g. (x, y, R, G, B) pixel
 * features, n up to 2^31 / 32.  Labels follow scikit-learn: clusters numbered by their lowest core index,
 * a non-core point within eps of core points takes the lowest-numbered of their clusters, the rest is -1.
 * The phases are separate calls (each is one or a few kernels) so that they can be timed one by one:
 *   plan_make (host) -> [bounds] -> bin -> count -> union -> border -> relabel. */
typedef struct rhccq_dbscan_plan {
    int n, dims, grid_dims, min_pts;
    double eps, side;                 /* cell side: eps * (1 + 2^-20) */
    double origin[3];
    int ncell[3];
    long long n_cells;
    int cells_per_tile;
    long long n_tiles;
} rhccq_dbscan_plan;

/* Host only.  lo / hi: bounds of the first grid_dims coordinates (rhccq_dbscan_bounds computes them). */
int rhccq_dbscan_plan_make(int n, int dims, int grid_dims, double eps, int min_pts, const double* host_lo,
                           const double* host_hi, rhccq_dbscan_plan* host_plan);
size_t rhccq_dbscan_workspace_bytes(const rhccq_dbscan_plan* host_plan);
/* out6 (device doubles): min of coordinates 0..2, max of coordinates 0..2; ws >= 24 KiB. */
int rhccq_dbscan_bounds(const float* pts, int n, int dims, int grid_dims, double* out6, void* ws, size_t ws_bytes,
                        void* stream);
/* cell of every point, cell histogram, exclusive scan, counting sort into 32-byte records */
int rhccq_dbscan_bin(const rhccq_dbscan_plan* host_plan, const float* pts, void* ws, size_t ws_bytes, void* stream);
/* neighbours within eps (self included); core[i] = count >= min_pts (uint8 [n], original order) */
int rhccq_dbscan_count(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, uint8_t* core, void* stream);
/* union-find over core pairs within eps; a set's root is its lowest original index */
int rhccq_dbscan_union(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, uint8_t* core, void* stream);
/* flatten + border attachment (= rhccq_dbscan_flatten followed by rhccq_dbscan_attach) */
int rhccq_dbscan_border(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, uint8_t* core, void* stream);
/* root (lowest index of its set) of every core point into the workspace's root array, -1 for the others */
int rhccq_dbscan_flatten(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, uint8_t* core, void* stream);
/* non-core points take the lowest root among the core points within eps (reads the root array) */
int rhccq_dbscan_attach(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, uint8_t* core, void* stream);
/* byte offset inside the workspace of: 0 the root array (int32 [n]), 1 scan scratch (int32 [n]),
 * 2 the union-find parents (int32 [n]), 3 the sorted records (8 x float32 [n]) */
size_t rhccq_dbscan_ws_offset(const rhccq_dbscan_plan* host_plan, int which);

/* ------------------------------------------------------------------ the same operator on an image lattice
 * X[i] = (x, y, R, G, B) of pixel i of an H x W image in raster order (the pixel features of BASELINE.json
 * configs 3-5): neighbours come from a (2 floor(eps) + 1)^2 stencil over packed 8-bit colours in a
 * shared-memory tile.  src_kind 0: float32 points [H*W, 5] (checked: *status = 1 when they are not such a
 * lattice — then use the generic path), 1: uint8 image [H, W, 3].  count: int32 [H*W], core: uint8 [H*W].
 * Labels are identical to the generic path's and to scikit-learn's. */
size_t rhccq_dbscan_lattice_workspace_bytes(int H, int W);
int rhccq_dbscan_lattice_count(const void* src, int src_kind, int H, int W, double eps, int min_pts, int32_t* count,
                               uint8_t* core, int32_t* status, void* ws, size_t ws_bytes, void* stream);
int rhccq_dbscan_lattice_union(int H, int W, double eps, int min_pts, void* ws, size_t ws_bytes, void* stream);
int rhccq_dbscan_lattice_border(int H, int W, double eps, int min_pts, const uint8_t* core, void* ws, size_t ws_bytes,
                                void* stream);
int rhccq_dbscan_lattice_relabel(int H, int W, void* ws, size_t ws_bytes, int32_t* labels, void* stream);
/* border = flatten (roots of core pixels into the root array) + attach (non-core pixels take the lowest adjacent root) */
int rhccq_dbscan_lattice_flatten(int H, int W, const uint8_t* core, void* ws, size_t ws_bytes, void* stream);
int rhccq_dbscan_lattice_attach(int H, int W, double eps, int min_pts, void* ws, size_t ws_bytes, void* stream);
/* byte offset inside the workspace of: 0 the root array, 1 root flags, 2 union-find parents, 3 packed colours (all
 * 4 bytes x H*W), 4 scan scratch */
size_t rhccq_dbscan_lattice_ws_offset(int H, int W, int which);

/* ------------------------------------------------------------------ strips of one point set across GPUs
 * (SURVEY.md 8e).  Every rank runs plan..flatten on its strip plus a halo of 2 eps; local indices are
 * global index - g0.  emit: for core points i in [lo, hi) of the boundary zone that are not their own root,
 * append (g0 + i, g0 + root) to edges (int32 pairs; *counter may exceed capacity: then enlarge and repeat).
 * merge: union-find over the gathered edges of all ranks in an open-addressing table (capacity a power of
 * two >= 4 n_edges), roots = lowest global index.  lookup: local roots -> global roots, in place (table NULL:
 * only the shift by g0).  own_roots: global indices of the roots among the rank's own points [own_lo,
 * own_hi), ascending.  rank_labels: label = position of the point's global root in the sorted list of all
 * ranks' roots (-1 for noise). */
int rhccq_uf_emit_edges(const int32_t* rootlab, int lo, int hi, int g0, int32_t* edges, int32_t* counter, int capacity,
                        void* stream);
int rhccq_uf_merge_edges(const int32_t* edges, int n_edges, int32_t* table_keys, int32_t* table_parent, int table_cap,
                         void* stream);
int rhccq_uf_lookup_roots(int32_t* rootlab, int n, int g0, const int32_t* table_keys, const int32_t* table_parent,
                          int table_cap, void* stream);
int rhccq_dbscan_own_roots(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, int own_lo, int own_hi, int g0,
                           int32_t* out_ids, int32_t* out_count, void* stream);
/* own_roots on a bare root array (global roots after rhccq_uf_lookup_roots); scratch: int32 [own_roots_scratch_ints(n)] */
size_t rhccq_uf_own_roots_scratch_ints(int n);
/* out_capacity > 0: ids beyond that many are dropped (out_count still holds the true number) */
int rhccq_uf_own_roots(const int32_t* rootlab, int n, int own_lo, int own_hi, int g0, int32_t* scratch, int32_t* out_ids,
                       int32_t* out_count, int out_capacity, void* stream);
int rhccq_uf_rank_labels(const int32_t* sorted_roots, int n_roots, const int32_t* rootlab, int lo, int hi, int32_t* labels,
                         void* stream);
/* The exchange step without the host in between: every rank contributes one block of stride_ints
 * int32 — [0] = number of payload rows, payload from int 2 (edge rows of two ints, or root ids) — and `gathered`
 * is the all-gather of the world's blocks in rank order.  merge: union-find over all edges (table_cap: a power of
 * two >= 4 * world * cap_rows); rank_labels: label of every own point = rank of its global root among all roots
 * (-2 everywhere when a block reports more rows than cap_rows: the caller must provide more room). */
int rhccq_uf_merge_edges_gathered(const int32_t* gathered, int world, int stride_ints, int cap_rows, int32_t* table_keys,
                                  int32_t* table_parent, int table_cap, void* stream);
int rhccq_uf_rank_labels_gathered(const int32_t* gathered, int world, int stride_ints, int cap_rows, const int32_t* rootlab,
                                  int lo, int hi, int32_t* labels, void* stream);
/* labels int32 [n], original order */
int rhccq_dbscan_relabel(const rhccq_dbscan_plan* host_plan, void* ws, size_t ws_bytes, int32_t* labels, void* stream);

/* ------------------------------------------------------------------ decoder side
 * out_rgb[i] = palette[indices[i]] (decoder/uncompression/uncompression.py:209); indices of 1, 2 or 4 bytes,
 * palette uint8 [n_palette,3]; *bad = 1 when an index is outside the palette (zero it first).
 * acc2 (int64 [2], zeroed by the caller) += (sum of squared differences, sum of absolute differences) of two
 * uint8 arrays: MSE, PSNR and MAE of decoder/uncompression/comparison.py:43-44,64-79. */
int rhccq_decode_gather(const void* indices, int idx_bytes, long long n, const uint8_t* palette, int n_palette,
                        uint8_t* out_rgb, int32_t* bad, void* stream);
int rhccq_sq_abs_err(const uint8_t* a, const uint8_t* b, long long n, long long* acc2, void* stream);

/* ------------------------------------------------------------------ container side
 * The 'i' entry of the .rhccq package is zlib(index bytes) (/root/reference/encoder/compression/compression.py:204-220,
 * read back by zlib.decompress, decoder/uncompression/uncompression.py:58-92).  rhccq_deflate_chunks turns every
 * rhccq_deflate_chunk_bytes() bytes of n_frames index planes (frame f at src + f * frame_stride, frame_bytes each;
 * elem_bytes per index, row_bytes per image row) into one fixed-Huffman deflate block followed by an empty stored
 * block, in the chunk's slot of rhccq_deflate_slot_bytes() bytes; slot_len[c] = bytes used, sums[2c], sums[2c+1] =
 * (sum of the chunk's bytes, sum of position-in-chunk * byte) for the Adler-32.  rhccq_deflate_pack copies slot c to
 * out + offsets[c].  The stream of a frame is 78 01 | its chunks | 03 00 | Adler-32 (big-endian). */
int rhccq_deflate_chunk_bytes(void);
int rhccq_deflate_slot_bytes(void);
int rhccq_deflate_chunks(const uint8_t* src, long long frame_stride, int frame_bytes, int n_frames, int elem_bytes,
                         int row_bytes, uint8_t* slots, int32_t* slot_len, unsigned long long* sums, void* stream);
int rhccq_deflate_pack(const uint8_t* slots, const int32_t* slot_len, const long long* offsets, long long n_slots,
                       uint8_t* out, void* stream);

/* out[i] = sum_{j<i} max(in[j],0), out[n] = total. */
int rhccq_excl_scan(const int32_t* in, int n, int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RHCCQ_H */
