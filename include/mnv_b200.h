/*
 * mnv_b200.h — C-ABI of the B200-native Mega-NeRF octree render path.
 *
 * This header is the drop-in boundary.  Every entry point
 * replaces one of the reference's host launchers / LibTorch call sites; the
 * reference interface each one stands in for is cited as
 * `<path under the reference tree>:<line>`.
 *
 * Conventions
 *   - plain C, no torch / glm / STL types in any signature;
 *   - every function returns an `int` status (MNV_OK == 0), never calls
 *     exit()/cudaDeviceReset() (the reference's cuda_assert does,
 *     src/cuda/common.cu:8-20);
 *   - pointers named *_dev are device pointers on the tree's device, *_host
 *     are host pointers (pinned memory makes the async copies truly async);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - there is NO CPU fallback: without a CUDA device every compute call
 *     returns MNV_ERR_NO_DEVICE.
 */
#ifndef MNV_B200_H
#define MNV_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNV_GLOBAL_BASIS_MAX 25 /* include/render_options.hpp:4 VIEWER_GLOBAL_BASIS_MAX */

enum {
    MNV_OK = 0,
    MNV_ERR_INVALID = 1,   /* bad argument / shape */
    MNV_ERR_NO_DEVICE = 2, /* no CUDA device: there is no CPU fallback */
    MNV_ERR_CUDA = 3,      /* CUDA runtime error, see mnv_last_error() */
    MNV_ERR_OOM = 4,
    MNV_ERR_IO = 5,      /* file missing / unreadable */
    MNV_ERR_FORMAT = 6,  /* .npz / model container schema violation */
    MNV_ERR_FULL = 7     /* tree capacity exhausted ("Full", cuda_renderer.cpp:228-231) */
};

/* include/data_format.hpp:7-22 */
enum { MNV_FORMAT_RGBA = 0, MNV_FORMAT_SH = 1 };

/* include/render_options.hpp:9-56 — same fields, same order, same defaults
 * (see mnv_render_options_default).  Layout-compatible with
 * viewer::RenderOptions on this ABI (bool = 1 byte). */
typedef struct mnv_render_options {
    float step_size;
    float sigma_thresh;
    float stop_thresh;
    float background_brightness;
    float render_bbox[6];
    int basis_minmax[2];
    float rot_dirs[3];
    bool show_grid;
    int grid_max_depth;
    bool render_depth;
    bool use_splitting;
    bool use_guided_sampling;
    int max_depth;
    int samples_per_corner;
    int split_batch_size;
    int nerf_batch_size;
    int max_sample_count;
    bool need_viewdir;
    int appearance_embedding;
    int max_guided_samples;
} mnv_render_options;

/* include/data_spec.hpp:9-23 CameraSpec + the 12 floats Camera::_update
 * uploads (src/camera.cpp:113-123).  c2w is glm::mat4x3 column-major:
 * right(3), up(3), back(3), center(3). */
typedef struct mnv_camera {
    int width, height;
    float fx, fy, cx, cy;
    float c2w[12];
} mnv_camera;

/* Host-side view of a loaded tree in the reference's AoS schema
 * (src/n3tree/n3tree.cpp:28-205). */
typedef struct mnv_tree_desc {
    int N;                 /* branching factor, only 2 is supported */
    int data_dim;          /* 3*basis_dim+1 (SH) or 4 (RGBA) */
    int format;            /* MNV_FORMAT_* */
    int basis_dim;         /* -1 for RGBA */
    int64_t capacity;      /* nodes in use */
    const uint16_t *data;  /* fp16 bits [capacity][8][data_dim] */
    const int32_t *child;  /* [capacity][8] relative offset, 0 = leaf */
    const int32_t *parent; /* [capacity] packed node*8+child, may be NULL */
    const int16_t *sample_counts; /* [capacity][8], NULL -> 8 everywhere (n3tree.cpp:191-193) */
    float scale[3];        /* invradius3 */
    float offset[3];
} mnv_tree_desc;

/* VQ-compressed leaf colours (svox `quant_colors` / `quant_map` / `data_retained` / `sigma`,
 * src/n3tree/n3tree.cpp:109-175): SH basis functions [n_retain, n_retain + n_quant) come from one 65536-entry
 * codebook each, the first n_retain are stored plainly, sigma is separate.  The reference decodes on the CPU with a
 * scalar triple loop (and writes `channel * n_basis` without `+ basis`, :145,:161); mnv_tree_create_vq uploads the
 * compressed arrays (2 B per basis function and leaf instead of 6) and decodes them on the device into the payload
 * plane, destination index channel * n_basis + basis. */
typedef struct mnv_vq_desc {
    int n_quant;                   /* quantised basis functions */
    int n_retain;                  /* plainly stored ones (0: data_retained may be NULL) */
    const uint16_t *quant_colors;  /* fp16 bits [n_quant][65536][3] */
    const uint16_t *quant_map;     /* u16 [n_quant][capacity][8] */
    const uint16_t *data_retained; /* fp16 bits [n_retain][capacity][8][3] */
    const uint16_t *sigma;         /* fp16 bits [capacity][8] */
} mnv_vq_desc;

typedef struct mnv_tree mnv_tree;   /* opaque device tree (SoA planes) */
typedef struct mnv_model mnv_model; /* opaque Mega-NeRF MLP container */

/* One Mega-NeRF sub-MLP, weights in PyTorch nn.Linear layout ([out][in], fp32, host).
 * Stands in for one `sub_module_i` of the TorchScript container the reference
 * loads in VolumeRenderer::load_model (src/renderer/cuda_renderer.cpp:518-543).
 * The architecture is not part of the reference repository; shapes follow
 * BASELINE.json / SURVEY.md §8 A9:
 *   h0 = PE(xyz, pe_xyz_freqs)                       (3 + 6*freqs values, include-input)
 *   h  = relu(trunk[l](l == skip_layer ? cat(h0, h) : h)),  l = 0 .. n_trunk_layers-1
 *   sigma = act(sigma_w . h + sigma_b)
 *   f  = final(h)                                    (width -> width, no activation)
 *   g  = relu(head1(cat(f, [PE(dir, pe_dir_freqs)], [embedding[appearance index]])))
 *   out = cat(head2(g), sigma)                       (out_rgb_dim + 1 columns)          */
typedef struct mnv_mlp_desc {
    int n_trunk_layers;   /* 8 */
    int width;            /* 256 (the only supported width) */
    int skip_layer;       /* 4 */
    int pe_xyz_freqs;     /* 12 */
    int pe_dir_freqs;     /* 4 */
    int need_viewdir;     /* container attr need_viewdir, cuda_renderer.cpp:536 */
    int appearance_dim;   /* 48; 0 = no appearance embedding */
    int n_appearance;     /* rows of the embedding table */
    int head_width;       /* 128 */
    int out_rgb_dim;      /* 3 * basis_dim (27 for SH9), 3 for RGB */
    int sigma_activation; /* 0 = ReLU, 1 = softplus */
    const float *trunk_w[12];
    const float *trunk_b[12];
    const float *sigma_w, *sigma_b;
    const float *final_w, *final_b;
    const float *embedding; /* [n_appearance][appearance_dim] */
    const float *head1_w, *head1_b;
    const float *head2_w, *head2_b;
} mnv_mlp_desc;

/* Per-frame statistics the traversal kernel can produce (bench / roofline). */
typedef struct mnv_frame_stats {
    uint64_t rays;
    uint64_t visits;        /* leaf visits = iterations of rt_core.cuh:220-324 */
    uint64_t shaded_visits; /* visits with sigma > sigma_thresh */
    uint64_t rays_hit;      /* rays that entered the bbox */
} mnv_frame_stats;

/* ---- misc ---------------------------------------------------------------- */
const char *mnv_version(void);
const char *mnv_last_error(void); /* thread-local message of the last failure */
int mnv_device_count(int *count);
void mnv_render_options_default(mnv_render_options *opt); /* render_options.hpp defaults */

/* ---- device memory helpers for host code that does not include CUDA headers ---------- */
int mnv_malloc(void **ptr_dev, size_t bytes, int device);
int mnv_free(void *ptr_dev);
int mnv_memset(void *ptr_dev, int value, size_t bytes, void *stream);
int mnv_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int mnv_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream); /* synchronises */
int mnv_fill_f32(float *ptr_dev, float value, int64_t n, void *stream);
int mnv_fill_i32(int32_t *ptr_dev, int32_t value, int64_t n, void *stream);
/* U[0,1) numbers (counter-based hash generator; stands in for torch::rand, cuda_renderer.cpp:250) */
int mnv_fill_uniform(float *ptr_dev, int64_t n, uint64_t seed, void *stream);
int mnv_malloc_host(void **ptr_host, size_t bytes); /* pinned (cudaHostAlloc) */
int mnv_free_host(void *ptr_host);
int mnv_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int mnv_stream_create(void **stream, int device); /* non-blocking cudaStream_t */
int mnv_stream_destroy(void *stream);
int mnv_stream_synchronize(void *stream);
/* cudaArray_t surfaces of the kind the viewer registers from its GL renderbuffers (RGBA8 colour, R32F
 * "fake depth", src/renderer/cuda_renderer.cpp:419-441) for hosts without GL: kind 0 = RGBA8, 1 = R32F. */
int mnv_array_create(void **array, int width, int height, int kind, int device);
int mnv_array_destroy(void *array);
int mnv_array_upload(void *array, const void *src_host, size_t row_bytes, int height);
int mnv_array_download(void *dst_host, void *array, size_t row_bytes, int height);

/* ---- tree: N3Tree::move_to_device, src/n3tree/n3tree.cpp:207-246 --------- */
int mnv_tree_create(mnv_tree **out, const mnv_tree_desc *desc, int64_t max_capacity, int device);
/* As mnv_tree_create with desc->data == NULL: the leaf payloads come from the VQ arrays, decoded on the device
 * (SH trees only; desc->data_dim == 3 * (n_quant + n_retain) + 1). */
int mnv_tree_create_vq(mnv_tree **out, const mnv_tree_desc *desc, const mnv_vq_desc *vq, int64_t max_capacity,
                       int device);
int mnv_tree_destroy(mnv_tree *tree);
int mnv_tree_capacity(const mnv_tree *tree, int64_t *capacity, int64_t *max_capacity);
int mnv_tree_device_bytes(const mnv_tree *tree, uint64_t *bytes);
/* Read the tree back in the reference's AoS schema (any pointer may be NULL). */
int mnv_tree_download(const mnv_tree *tree, int64_t first, int64_t count, uint16_t *data,
                      int32_t *child, int32_t *parent, int16_t *sample_counts);

/* The library keeps one cudaSurfaceObject_t per cudaArray_t it has been handed (the reference creates one per
 * launch and never destroys it, renderer_kernel.cu:377-385).  Call this before such an array is destroyed or
 * unregistered (the viewer: in resize, before cudaGraphicsUnregisterResource, cuda_renderer.cpp:417-421): it
 * synchronises the device and destroys the cached objects, so a recycled array handle is never mistaken for
 * the old one. */
int mnv_tree_release_surfaces(mnv_tree *tree);

/* Launch order of the march kernel's 16x8-pixel CTA tiles for frames with exactly n tiles
 * (ceil(W/16) * ceil(H/8)): CTA i renders tile order_dev[i] (a permutation of 0..n-1, caller-owned device
 * memory; NULL restores row-major).  Results do not depend on the order; the frame time does — the kernel cannot
 * end before its longest rays do, so tiles that hold them should start first. */
int mnv_tree_set_tile_order(mnv_tree *tree, const int32_t *order_dev, int n);

/* ---- point query: query_single_from_root, include/cuda/rt_core.cuh:117-159
 * xyz_dev: [n][3] tree-space coordinates; out_dev: [n][3] = chunk, child, depth. */
int mnv_query_points(const mnv_tree *tree, const float *xyz_dev, int64_t n, int32_t *out_dev,
                     void *stream);

/* ---- octree render: viewer::render_voxels, src/cuda/renderer_kernel.cu:396-437
 * (kernel :243-292, march include/cuda/rt_core.cuh:162-332).
 *   image_arr / depth_arr : cudaArray_t (RGBA8 / R32F surfaces, GL interop) or NULL
 *   image_linear_dev      : RGBA8 [height][width] linear device buffer or NULL
 *   to_split_dev/to_sample_dev : f32 [P][3] = (priority, chunk, child) or NULL
 *   visited_dev           : i32 [max_capacity] or NULL (then track_visit must be false)
 * Exactly one of image_arr / image_linear_dev must be given.  */
int mnv_render_voxels(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                      void *image_arr, void *depth_arr, uint8_t *image_linear_dev,
                      float *to_split_dev, float *to_sample_dev, int32_t *visited_dev,
                      bool track_visit, bool offscreen, void *stream);

/* Image-tile partition for multi-GPU rendering: render only
 * the 8x4-pixel-aligned tiles t with (t % tile_mod) == tile_rem of a
 * tile_w x tile_h tiling; pixels outside are left untouched. */
int mnv_render_voxels_tiles(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                            uint8_t *image_linear_dev, float *to_split_dev, float *to_sample_dev,
                            int tile_w, int tile_h, int tile_mod, int tile_rem, void *stream);

/* Debug/parity variant: additionally writes per ray an FNV-1a hash of the
 * visited (chunk*8+child) sequence, the visit count, and the first `log_cap`
 * packed leaf indices. All outputs are device pointers; any may be NULL. */
int mnv_render_voxels_logged(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                             uint8_t *image_linear_dev, uint64_t *visit_hash_dev,
                             int32_t *visit_count_dev, int32_t *shaded_count_dev,
                             int32_t *visit_log_dev, int log_cap, void *stream);

/* The call a host application makes per frame with HOST buffers: uploads the
 * camera (48 B, src/camera.cpp:113-123), renders offscreen, reads the RGBA8
 * frame back into rgba_host ([height][width][4]); synchronous on return.
 * When opt->use_splitting is set the split / re-sample candidates are produced
 * too (into tree-owned device buffers, see mnv_tree_trackers). */
int mnv_render_frame_host(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                          uint8_t *rgba_host, mnv_frame_stats *stats /* may be NULL */);

/* ---- guided sampling: viewer::get_samples_from_voxels, src/cuda/renderer_kernel.cu:439-485
 * (kernel :329-363, emitter include/cuda/rt_core.cuh:418-576) + the cumsum / mask
 * compaction of cuda_renderer.cpp:116-120, in CSR form.  Outputs (device):
 *   offsets_dev  i64 [P]  inclusive scan of per-ray sample counts (== torch::cumsum)
 *   z_vals_dev   f32 [V]  row column 0 of the reference's guided_samples
 *   rows_dev     f32 [V][row_stride]  MLP input rows: x,y,z,(view dir),(appearance index)
 *   cluster_dev  i16 [V]  sub-module id, grid rule of rt_core.cuh:541-549
 * V = *total_rows_host (written after the internal count pass; the call
 * synchronises the stream once, like the reference). MNV_ERR_FULL if V exceeds
 * capacity_rows. grid_dim / min_position / range are host pointers. */
int mnv_guided_samples(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                       void *depth_arr, bool offscreen, const int32_t grid_dim[2],
                       const float min_position[3], const float range[3], int64_t *offsets_dev,
                       float *z_vals_dev, float *rows_dev, int row_stride, int16_t *cluster_dev,
                       int64_t capacity_rows, int64_t *total_rows_host, float *to_split_dev,
                       float *to_sample_dev, int32_t *visited_dev, bool track_visit, void *stream);

/* ---- viewer::render_nerf_results, src/cuda/renderer_kernel.cu:365-394 (kernel :294-327,
 * compositor rt_core.cuh:334-416). sample_values_dev f32 [V][value_stride] (MLP outputs),
 * sigma_col < 0 keeps the reference's behaviour of reading sigma from column 3
 * (rt_core.cuh:365) whatever the data format. */
int mnv_render_nerf_results(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                            void *image_arr, uint8_t *image_linear_dev,
                            const float *sample_values_dev, int value_stride, int sigma_col,
                            const float *z_vals_dev, const int64_t *offsets_dev, bool offscreen,
                            void *stream);

/* ---- dynamic refinement --------------------------------------------------------------
 * viewer::add_children_and_generate_samples, src/cuda/renderer_kernel.cu:487-511 (kernel
 * :170-198): link 8 children under each of the n chosen leaves (parent_nodes_dev i32 [n][2] =
 * chunk, child) at node indices capacity .. capacity+n-1, and turn the U[0,1) numbers in
 * samples_dev (f32 [n*8][samples_per_corner][rand_dim], rand_dim = 3 + 3*need_viewdir +
 * (appearance_embedding != -1)) into world-space sample rows; cluster_dev i16
 * [n*8][samples_per_corner].  The children become visible to the march at once (sigma 0)
 * and count towards capacity after mnv_tree_commit_children.  MNV_ERR_FULL when the tree
 * cannot take n more nodes ("Full", cuda_renderer.cpp:228-231). */
int mnv_add_children_and_generate_samples(mnv_tree *tree, const mnv_render_options *opt,
                                          const int32_t *parent_nodes_dev, int n, float *samples_dev,
                                          int16_t *cluster_dev, int32_t *visited_dev,
                                          const int32_t grid_dim[2], const float min_position[3],
                                          const float range[3], void *stream);
/* Impl::expand_voxels tail, cuda_renderer.cpp:266-275: leaf payload = mean over the
 * samples_per_corner MLP outputs (results_dev f32 [n*8][samples_per_corner][result_stride],
 * first data_dim columns used) rounded to fp16, sample_counts = samples_per_corner,
 * capacity += n. */
int mnv_tree_commit_children(mnv_tree *tree, const mnv_render_options *opt, int n,
                             const float *results_dev, int result_stride, void *stream);
/* Multi-GPU form of the same tail (MLP rows sharded across the replicas, 1.8 MB of fp16
 * payloads exchanged instead of 30 MB of fp32 results).  mnv_tree_reduce_children turns the MLP outputs of
 * n_children consecutive new leaves (results_dev f32 [n_children][samples_per_corner][result_stride]) into their
 * payload records — records_dev, mnv_tree_record_bytes() bytes each, the device layout of one leaf — without
 * touching the tree; after the all-gather every replica calls mnv_tree_commit_children_records with the
 * records of all n * 8 children in child order.  Equal, bit for bit, to mnv_tree_commit_children. */
int mnv_tree_record_bytes(const mnv_tree *tree, int *bytes);
int mnv_tree_reduce_children(mnv_tree *tree, const mnv_render_options *opt, int n_children,
                             const float *results_dev, int result_stride, void *records_dev, void *stream);
int mnv_tree_commit_children_records(mnv_tree *tree, const mnv_render_options *opt, int n,
                                     const void *records_dev, void *stream);
/* viewer::generate_samples, renderer_kernel.cu:513-535 (kernel :200-213): sample rows for m
 * existing leaves (nodes_dev i32 [m][2]). */
int mnv_generate_samples(mnv_tree *tree, const mnv_render_options *opt, const int32_t *nodes_dev, int m,
                         float *samples_dev, int16_t *cluster_dev, const int32_t grid_dim[2],
                         const float min_position[3], const float range[3], void *stream);
/* Impl::get_more_samples tail, cuda_renderer.cpp:318-339: running-mean update of the m leaves
 * with samples_per_corner new MLP outputs each; sample_counts += samples_per_corner. */
int mnv_tree_update_samples(mnv_tree *tree, const mnv_render_options *opt, const int32_t *nodes_dev,
                            int m, const float *results_dev, int result_stride, void *stream);
/* Impl::prune_tree, cuda_renderer.cpp:343-381 + viewer::adjust_parents_and_children,
 * renderer_kernel.cu:537-550 (kernel :63-86): to_delete_dev u8 [capacity] (1 = drop node),
 * index_shifts_dev i32 [capacity] = inclusive cumsum of to_delete, num_deleted = its last
 * element.  Fixes child / parent links, compacts every plane in place, capacity -= num_deleted. */
int mnv_tree_prune(mnv_tree *tree, const uint8_t *to_delete_dev, const int32_t *index_shifts_dev,
                   int first_shift_index, int64_t num_deleted, void *stream);

/* The whole of Impl::prune_tree (cuda_renderer.cpp:343-381): drop every node whose
 * visited_dev[node] == 0 (node 0 is always kept), then zero visited_dev[1..max_capacity).
 * *num_deleted_host receives the number of reclaimed nodes (0: "Nothing can be pruned"). */
int mnv_tree_prune_unvisited(mnv_tree *tree, int32_t *visited_dev, int64_t *num_deleted_host, void *stream);

/* ---- Mega-NeRF MLP: torch::jit::load + Module::forward, cuda_renderer.cpp:165-203,518-543
 * Container attributes grid_dim / min_position / max_position (cluster rule,
 * rt_core.cuh:541-549) travel with the model. x rows are
 * [x, y, z, (dir x3 if need_viewdir), (appearance index as float if appearance_dim > 0)];
 * out rows are [rgb / SH (out_rgb_dim), sigma] at stride out_stride floats.
 * bf16 operands, fp32 accumulation in TMEM (tcgen05), fp32 bias / activations. */
int mnv_model_create(mnv_model **out, int n_submodules, const mnv_mlp_desc *descs,
                     const int32_t grid_dim[2], const float min_position[3],
                     const float max_position[3], int device);
int mnv_model_destroy(mnv_model *model);
int mnv_model_info(const mnv_model *model, int *n_submodules, int *in_dim, int *out_dim,
                   double *flops_per_row);
int mnv_mlp_forward(mnv_model *model, int submodule, const float *x_dev, int64_t rows, int in_dim,
                    float *out_dev, int out_stride, void *stream);

/* Impl::query_submodules, cuda_renderer.cpp:165-203: evaluate rows_dev f32 [V][in_dim] with
 * the sub-module named by cluster_dev i16 [V]; out_dev f32 [V][out_stride] receives
 * [rgb / SH, sigma] in the first out_dim columns of each row (scatter_ semantics). */
int mnv_query_submodules(mnv_model *model, const int16_t *cluster_dev, const float *rows_dev, int in_dim,
                         int64_t rows, float *out_dev, int out_stride, void *stream);

/* Candidate selection of Impl::expand_voxels (cuda_renderer.cpp:205-226): unique tracker rows
 * with their vote counts, keep rows voted by >= 2 rays, order by (-count, depth, chunk, child),
 * take the first max_n -> nodes_dev i32 [max_n][2] = (chunk, child).  *n_selected rows are
 * valid; *n_candidates is the reference's "Split candidates" count. */
int mnv_select_split_candidates(const float *to_split_dev, int64_t n_rays, int max_n, int32_t *nodes_dev,
                                int *n_selected, int *n_candidates, void *stream);
/* Impl::get_more_samples (cuda_renderer.cpp:281-293): unique rows ordered by
 * (sample count, chunk, child), first max_n. */
int mnv_select_sample_candidates(const float *to_sample_dev, int64_t n_rays, int max_n,
                                 int32_t *nodes_dev, int *n_selected, int *n_candidates, void *stream);

/* Multi-GPU refinement: instead of exchanging the raw [P][3] float tracker rows, each GPU
 * reduces its rows to vote records — u32 x 3 = (leaf id = chunk*8+child, priority, votes), unordered — which
 * are all-gathered and merged.  mnv_vote_reduce writes at most cap_records records into records_dev and
 * returns their number in *n_records (host); MNV_ERR_FULL if they do not fit (nothing is lost: *n_records
 * is still the true count).  mnv_select_candidates_from_votes is the selection of
 * mnv_select_split_candidates (kind 0) / mnv_select_sample_candidates (kind 1) over the union of optional
 * local tracker rows and n_records gathered records (records with votes == 0 are padding). */
int mnv_vote_reduce(const float *tracker_dev, int64_t n_rays, uint32_t *records_dev, int64_t cap_records,
                    int64_t *n_records, void *stream);
int mnv_select_candidates_from_votes(int kind, const float *tracker_dev, int64_t n_rays,
                                     const uint32_t *records_dev, int64_t n_records, int max_n,
                                     int32_t *nodes_dev, int *n_selected, int *n_candidates, void *stream);

/* The chunk column of the trackers: the reference's float VALUE while that is exact (chunk < 2^24, the
 * reference's own rows bit for bit), the id's raw integer BITS above (the reference's float rows lose odd
 * node ids beyond 16.7 M nodes, rt_core.cuh:238-240 with src/opts.cpp:24).  Host helpers for consumers. */
float mnv_tracker_encode_chunk(int32_t chunk);
int32_t mnv_tracker_decode_chunk(float column_value);

/* Device pointers of the tree-owned candidate buffers filled by the host frame
 * calls when opt->use_splitting is set: f32 [P][3] each (valid until the next
 * frame call with a different size). */
int mnv_tree_trackers(mnv_tree *tree, float **to_split_dev, float **to_sample_dev);

/* Multi-GPU form of mnv_render_frame_host (image tiles with the
 * tree replicated): the frame is cut into bands of band_rows full-width pixel
 * rows (band_rows a multiple of 8); this call renders the bands b with
 * (b % band_mod) == band_rem and copies exactly those bands to their place in
 * the full-size host frame rgba_host (one strided D2H copy). No collective. */
int mnv_render_frame_host_bands(mnv_tree *tree, const mnv_camera *cam,
                                const mnv_render_options *opt, uint8_t *rgba_host, int band_rows,
                                int band_mod, int band_rem, mnv_frame_stats *stats);

/* ---- sub-module split across GPUs (second mode; no reference counterpart —
 * the reference is single-GPU).  GPU g owns one spatial cell of the Mega-NeRF (y, z) grid and
 * marches every ray of the frame through that cell only; the pixel range [o*block, (o+1)*block)
 * of the frame is composited by owner o.
 *
 * cell_box (tree space, lo xyz / hi xyz; NULL: no cell) clips the march on top of opt->render_bbox,
 *   which stays the caller's.  Where a ray enters the cell through a face inside the render box the
 *   segment starts step_size behind that face — where the unsharded march would land coming from
 *   the previous leaf (rt_core.cuh:228-230) — instead of marching a sliver the full frame skips.
 *
 * mnv_render_voxels_partial: like mnv_render_voxels(offscreen) but each ray's premultiplied
 *   (r, g, b, alpha) goes, as one 16-byte store, to partial_dst[o] + (slot*block + p % block),
 *   o = p / block.  partial_dst are device pointers valid on this GPU: the local buffer or
 *   peers' buffers mapped with mnv_ipc_open (NVLink stores — the exchange is fused in the march).
 * mnv_signal_peers: after the march on the same stream, raises flag_dst[o][slot] = value in
 *   every owner (system-scope fence first).
 * mnv_composite_partials: owner side. partials_dev f32 [n][block][4] in slot order; boxes
 *   f32 [n][6] host, the cells in slot order; waits on flags_dev[0..n) >= wait_value on the
 *   device when flags_dev != NULL; orders the segments front to back per ray, composes, blends
 *   opt->background_brightness and writes RGBA8 for pixels [first_pixel, first_pixel+n_pixels)
 *   into rgba_dev[0..n_pixels). */
int mnv_render_voxels_partial(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                              const float cell_box[6], int n_owners, float *const *partial_dst,
                              int block_pixels, int slot, void *stream);
int mnv_signal_peers(uint32_t *const *flag_dst, int n, int slot, uint32_t value, void *stream);
int mnv_composite_partials(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                           const float *partials_dev, int n, int block_pixels, const float *boxes_host,
                           int64_t first_pixel, int n_pixels, uint8_t *rgba_dev, const uint32_t *flags_dev,
                           uint32_t wait_value, void *stream);
/* Guided sampling with the sub-modules sharded across GPUs (BASELINE.json configs[4]): rank g holds
 * cell g's subtree and sub-MLP only and handles the segment every ray has inside cell g.  The
 * emission loop of the unsharded frame (rt_core.cuh:321-357) carries two values along the ray — the
 * octree transmittance T (emission stops below stop_thresh) and the sample count (capped at
 * max_guided_samples) — and the compositor (rt_core.cuh:361-372) gives every sample the interval
 * to the NEXT sample.  One exchange of 16 bytes per ray per rank reproduces all three:
 * mnv_guided_segment_probe: march the cell with T = 1, count = 0; probe_dev f32 [P][4] =
 *   (T at the cell's exit, samples emitted, z of the first sample or 3e38, 0).
 *   The ranks all-gather these -> probe_all_dev f32 [n_cells][P][4].
 * mnv_guided_samples_segment: mnv_guided_samples for the segment, with T and the count at the
 *   cell's entry derived from the records of the segments in front of it (ordered by first z).
 * mnv_render_nerf_results_partial: render_nerf_results_kernel (renderer_kernel.cu:298-319) per
 *   segment: the segment's last sample gets the interval to the first sample of the next emitting
 *   segment, and only the ray's overall last sample takes the remaining transmittance.  Stores
 *   premultiplied (r, g, b, 1 - T_segment) into the pixel owner's buffer like
 *   mnv_render_voxels_partial.  render_depth is not available here (MNV_ERR_INVALID).
 * mnv_composite_partials_guided: mnv_composite_partials without the early-termination rule and
 *   with the frame opaque (out[3] = 1, renderer_kernel.cu:315-316). */
int mnv_guided_segment_probe(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                             const float cell_box[6], float *probe_dev, void *stream);
int mnv_guided_samples_segment(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                               const float cell_box[6], const int32_t grid_dim[2], const float min_position[3],
                               const float range[3], const float *probe_all_dev, int n_cells, int slot,
                               int64_t *offsets_dev, float *z_vals_dev, float *rows_dev, int row_stride,
                               int16_t *cluster_dev, int64_t capacity_rows, int64_t *total_rows_host,
                               void *stream);
int mnv_render_nerf_results_partial(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                                    const float *sample_values_dev, int value_stride, int sigma_col,
                                    const float *z_vals_dev, const int64_t *offsets_dev,
                                    const float *probe_all_dev, int n_cells, int slot, int n_owners,
                                    float *const *partial_dst, int block_pixels, void *stream);
int mnv_composite_partials_guided(mnv_tree *tree, const mnv_camera *cam, const mnv_render_options *opt,
                                  const float *partials_dev, int n, int block_pixels,
                                  const float *boxes_host, int64_t first_pixel, int n_pixels,
                                  uint8_t *rgba_dev, const uint32_t *flags_dev, uint32_t wait_value,
                                  void *stream);
/* CUDA IPC plumbing for the peer mappings (one process per GPU): a 64-byte handle of a buffer
 * from mnv_malloc, opened in another process with lazy peer access. */
int mnv_ipc_export(void *ptr_dev, uint8_t handle[64]);
int mnv_ipc_open(const uint8_t handle[64], void **ptr_dev, int device);
int mnv_ipc_close(void *ptr_dev);

/* ---- replica group: image tiles over the GPUs of one box, driven by ONE process (the viewer's shape) -----------
 * SURVEY.md §8(e), first mode; the reference is single-GPU (renderer_kernel.cu:17).  The tree is replicated on
 * every device; replica i marches the interleaved band_rows-row bands b % n == i; finished bands reach devices[0]
 * as one strided NVLink peer copy per replica, ordered by events on the device (no host synchronisation, no
 * collective).  The gathered frame — bit-identical to the one-GPU frame — lands in a linear RGBA8 buffer on
 * devices[0] and / or in the cudaArray behind the GL renderbuffer (image_arr_dev0; cuda_renderer.cpp:432-457). */
typedef struct mnv_group mnv_group;
int mnv_group_create(mnv_group **out, const mnv_tree_desc *desc, int64_t max_capacity, const int *devices,
                     int n_devices);
int mnv_group_destroy(mnv_group *group);
int mnv_group_size(const mnv_group *group, int *n);
int mnv_group_tree(mnv_group *group, int i, mnv_tree **tree); /* replica i, owned by the group */
int mnv_group_render_frame(mnv_group *group, const mnv_camera *cam, const mnv_render_options *opt,
                           uint8_t *image_linear_dev0, void *image_arr_dev0, int band_rows);
int mnv_group_render_frame_host(mnv_group *group, const mnv_camera *cam, const mnv_render_options *opt,
                                uint8_t *rgba_host, int band_rows);
int mnv_group_synchronize(mnv_group *group);
/* One frame with dynamic refinement ON across the group (Impl::render + expand_voxels, cuda_renderer.cpp:68-163,
 * :205-278): votes of each replica's bands reduced to records and exchanged with peer copies, identical selection
 * and linking everywhere, MLP rows sharded by child (models[i] lives on replica i's device), fp16 payload records
 * exchanged and committed everywhere.  All replicas hold the same tree afterwards.  The frame (marched before the
 * split, like the reference's) goes to any of the three targets that is not NULL. */
int mnv_group_refine_frame(mnv_group *group, mnv_model *const *models, const mnv_camera *cam,
                           const mnv_render_options *opt, const int32_t grid_dim[2], const float min_position[3],
                           const float range[3], uint64_t seed, uint8_t *image_linear_dev0, void *image_arr_dev0,
                           uint8_t *rgba_host, int band_rows, int *nodes_added);

#ifdef __cplusplus
}
#endif
#endif /* MNV_B200_H */
