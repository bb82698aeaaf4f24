// Internal declarations shared by the CUDA translation units of libmnv_b200.so.
// Nothing here is part of the public C-ABI (include/mnv_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/mnv_b200.h"

namespace mnv {

// ---------------------------------------------------------------- errors ----
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MNV_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::mnv::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------- device tree (SoA) --
//
// The reference keeps the tree as four AoS tensors (include/data_spec.hpp:25-50):
//   data[max][8][D] f16, child[max][8] i32 (relative), parent[max] i32,
//   sample_counts[max][8] i16.
// Here a leaf visit needs exactly ONE 4-byte load to learn everything the
// march needs to decide what to do with a cell:
//
//   cell[node*8 + slot]  (u32 plane)
//     bit31 = 0 : internal -> bits 0..30 = ABSOLUTE index of the child node
//     bit31 = 1 : leaf     -> bits 16..30 = sample count (i16 >= 0),
//                             bits 0..15  = sigma, fp16 bits
//
//   payload[(node*8 + slot) * rec_u4]  (uint4 plane, 16-byte aligned records)
//     the leaf's data_dim fp16 values (colour / SH coefficients, sigma last),
//     zero-padded to a multiple of 8 halfs: SH9 -> 28 halfs -> one 64-byte
//     record = two 32-byte sectors, fetched with 4 LDG.128 instead of the
//     reference's 27 scalar LDG.U16 of a 56-byte record straddling 3 sectors.
//
//   parent[node] (i32 plane, packed parent slot = parent_node*8 + child) is
//     only touched by refinement / pruning.
struct DeviceTree {
    uint32_t *cell = nullptr;
    uint4 *payload = nullptr;
    int32_t *parent = nullptr;
    int16_t *sample_counts = nullptr;  // [max*8] authoritative counts (refinement only)
    int N = 2;
    int data_dim = 0;
    int format = MNV_FORMAT_RGBA;
    int basis_dim = -1;
    int rec_u4 = 0;  // uint4 per payload record
    int64_t capacity = 0;
    int64_t max_capacity = 0;
    float scale[3] = {1, 1, 1};
    float offset[3] = {0, 0, 0};
    int device = 0;
    int pending_children = 0;  // nodes linked by add_children, not yet committed
    int max_leaf_depth = 1;  // deepest leaf (reference counting: root's children are depth 1); an upper bound
                             // while a refinement step's exact depth is still in flight (depth_pending)
    int *max_depth_dev = nullptr;   // exact deepest leaf, atomicMax'ed by add_children_kernel
    int *max_depth_host = nullptr;  // pinned copy, valid once depth_event has completed
    cudaEvent_t depth_event = nullptr;
    bool depth_pending = false;
    // Anchor grid (mnv_tree.cu): a dense 2^A x 2^A x 2^A table over the unit cube.  Entry (ix, iy, iz) tells the
    // march where the point's descent stands at level A: either the leaf itself when the tree ends at depth <= A
    // there (x = its cell word, y = depth-1 << 28 | slot) or the level-A node to continue from (x = node,
    // y = A << 28).  One 8-byte load replaces the A top levels of query_single_from_root (rt_core.cuh:117-159).
    // Rebuilt lazily (anchor_dirty) after any kernel that rewrites the cell plane.
    uint2 *anchor = nullptr;
    int anchor_level = 0;  // A (0: disabled)
    bool anchor_dirty = true;
    // scratch for mnv_render_frame_host
    int32_t *count_dev = nullptr;  // [P] per-ray sample counts (guided sampling)
    int64_t count_cap = 0;
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    uint8_t *frame_dev = nullptr;
    float *split_dev = nullptr, *sample_dev = nullptr;
    size_t frame_bytes = 0;
    unsigned long long *stats_dev = nullptr;
    const int32_t *tile_order_dev = nullptr;  // dev hook (mnv_tree_set_tile_order): caller-owned launch order
    int tile_order_n = 0;
    void **partial_table_dev = nullptr;  // 8 pointers: where the owners' partial buffers are mapped on this GPU
    void *partial_table_host[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;
};

constexpr uint32_t kLeafBit = 0x80000000u;

__host__ __device__ inline uint32_t make_leaf_cell(uint16_t sigma_bits, int sample_count) {
    const uint32_t sc = (uint32_t) (sample_count < 0 ? 0 : (sample_count > 32767 ? 32767 : sample_count));
    return kLeafBit | (sc << 16) | sigma_bits;
}

// ---- candidate trackers -------------------------------------------------------------------
// The reference's trackers are float rows (priority, chunk, child) (rt_core.cuh:238-252), and fp32 holds
// integers exactly only below 2^24 — while its own default capacity is 2*10^7 nodes (src/opts.cpp:24;
// SURVEY.md quirk 2): above 16.7 M nodes odd node ids silently round to a neighbour.  The column keeps the
// reference's encoding wherever that is exact (chunk < 2^24: the float VALUE, bit-identical to the
// reference's rows) and carries larger ids as their raw integer BITS.  The two ranges cannot collide:
// ids in [2^24, 2^28) are the bit patterns 0x01000000..0x0FFFFFFF — positive floats below 1e-29 — whereas
// every float-encoded id >= 1 has bits >= 0x3F800000, 0 is 0 and "none" is -1.f (sign bit set).
__host__ __device__ inline float tracker_encode_chunk(int32_t chunk) {
    if (chunk < (1 << 24)) return (float) chunk;
#ifdef __CUDA_ARCH__
    return __int_as_float(chunk);
#else
    union { int32_t i; float f; } u;
    u.i = chunk;
    return u.f;
#endif
}
__host__ __device__ inline int32_t tracker_decode_chunk(float v) {
#ifdef __CUDA_ARCH__
    const int32_t bits = __float_as_int(v);
#else
    union { float f; int32_t i; } u;
    u.f = v;
    const int32_t bits = u.i;
#endif
    return (bits >= (1 << 24) && bits < (1 << 28)) ? bits : (int32_t) v;
}

// Kernel-side view, passed by value.
struct TreeView {
    const uint32_t *__restrict__ cell;
    const uint4 *__restrict__ payload;
    int rec_u4;
    int data_dim;
    int format;
    int basis_dim;
    float scale[3];
    float offset[3];
};

inline TreeView make_view(const DeviceTree &t) {
    TreeView v;
    v.cell = t.cell;
    v.payload = t.payload;
    v.rec_u4 = t.rec_u4;
    v.data_dim = t.data_dim;
    v.format = t.format;
    v.basis_dim = t.basis_dim;
    for (int i = 0; i < 3; ++i) {
        v.scale[i] = t.scale[i];
        v.offset[i] = t.offset[i];
    }
    return v;
}

// Outputs / modes of one render launch.
struct RenderTargets {
    uint8_t *image_linear = nullptr;      // RGBA8 [H][W]
    cudaSurfaceObject_t image_surf = 0;   // or an RGBA8 surface (GL interop)
    cudaSurfaceObject_t depth_surf = 0;   // R32F t_max surface, read when !offscreen
    float *to_split = nullptr;            // [P][3]
    float *to_sample = nullptr;           // [P][3]
    int32_t *visited = nullptr;           // [max_capacity]
    bool track_visit = false;
    bool offscreen = true;
    // parity / statistics
    unsigned long long *visit_hash = nullptr;
    int32_t *visit_count = nullptr;
    int32_t *shaded_count = nullptr;
    int32_t *visit_log = nullptr;
    int log_cap = 0;
    unsigned long long *frame_stats = nullptr;  // [4] rays, visits, shaded, rays_hit
    // launch order of the 16x8-pixel CTA tiles: CTA i renders tile tile_order[i] (null: row-major)
    const int32_t *tile_order = nullptr;
    // image-tile partition (multi-GPU): render tiles with (tile % mod) == rem
    int tile_w = 0, tile_h = 0, tile_mod = 1, tile_rem = 0;
    // sub-module split (multi-GPU): instead of an image, write the ray's premultiplied partial
    // (r, g, b, alpha) into the memory of the GPU that owns the pixel — pixel p belongs to owner
    // p / partial_block and lands in its buffer [partial_slot][p % partial_block] (local or peer pointers)
    float4 *const *partial_dst = nullptr;  // device table of partial_n pointers
    int partial_n = 0, partial_block = 0, partial_slot = 0;
    bool has_cell = false;  // clip the march to cell_box (tree space) with the interior-entry rule of clip_to_cell
    float cell_box[6] = {0, 0, 0, 1, 1, 1};
};

int launch_render_voxels(DeviceTree &tree, const mnv_camera &cam,
                         const mnv_render_options &opt, const RenderTargets &tg,
                         cudaStream_t stream);
// (re)builds tree.anchor on `stream` when the cell plane changed since the last build
int ensure_anchor(DeviceTree &tree, cudaStream_t stream);
// picks up the exact deepest-leaf level of the last refinement step once its copy has landed
void refresh_max_leaf_depth(DeviceTree &tree);
// ancestors of visited nodes become visited (the reference marks every node of the root path per query)
int launch_propagate_visited(const DeviceTree &tree, int32_t *visited, cudaStream_t stream);

int launch_query_points(const DeviceTree &tree, const float *xyz, int64_t n, int32_t *out,
                        cudaStream_t stream);

const char *last_error_cstr();
int build_device_tree(DeviceTree &t, const mnv_tree_desc &desc, const mnv_vq_desc *vq = nullptr);
int download_device_tree(const DeviceTree &t, int64_t first, int64_t count, uint16_t *data,
                         int32_t *child, int32_t *parent, int16_t *sample_counts);

// ---- guided sampling (mnv_guided.cu) ------------------------------------------
struct GuidedIO {
    cudaSurfaceObject_t depth_surf = 0;
    bool offscreen = true;
    int64_t *offsets = nullptr;   // [P] inclusive scan of per-ray sample counts (out)
    float *z_vals = nullptr;      // [capacity_rows]
    float *rows = nullptr;        // [capacity_rows][row_stride]
    int16_t *cluster = nullptr;   // [capacity_rows]
    int row_stride = 0;
    int64_t capacity_rows = 0;
    int64_t *total_rows = nullptr;  // host, out
    float *to_split = nullptr, *to_sample = nullptr;
    int32_t *visited = nullptr;
    bool track_visit = false;
    int32_t grid_dim[2] = {1, 1};
    float min_position[3] = {0, 0, 0}, range[3] = {1, 1, 1};
    // sub-modules sharded across GPUs (mnv_guided_segment_probe / mnv_guided_samples_segment)
    float4 *seg_probe = nullptr;        // [P]: probe pass only, nothing else is written
    const float4 *seg_table = nullptr;  // [seg_n][P] all ranks' probe records
    int seg_n = 0, seg_slot = 0;
    bool has_cell = false;
    float cell_box[6] = {0, 0, 0, 1, 1, 1};
};
int launch_guided_samples(DeviceTree &tree, const mnv_camera &cam, const mnv_render_options &opt,
                          const GuidedIO &io, cudaStream_t stream);
// one rank's share of a guided-sampling frame whose sub-modules are sharded across GPUs
struct NerfSegment {
    const float4 *seg_table;     // [n_seg][P] all ranks' probe records
    int n_seg, slot;
    float4 *const *partial_dst;  // device table of the owners' buffers
    int partial_block;
};
int launch_composite_nerf(const DeviceTree &tree, const mnv_camera &cam,
                          const mnv_render_options &opt, uint8_t *image_linear,
                          cudaSurfaceObject_t image_surf, const float *values, int value_stride,
                          int sigma_col, const float *z_vals, const int64_t *offsets, bool offscreen,
                          cudaStream_t stream, const NerfSegment *seg = nullptr);

// ---- refinement (mnv_refine.cu) -------------------------------------------------
int refine_add_children(DeviceTree &t, const mnv_render_options &opt, const int32_t *parent_nodes_dev,
                        int n, float *samples_dev, int16_t *cluster_dev, int32_t *visited_dev,
                        const int32_t *grid_dim, const float *min_position, const float *range,
                        cudaStream_t stream);
int refine_commit_children(DeviceTree &t, const mnv_render_options &opt, int n, const float *results_dev,
                           int result_stride, cudaStream_t stream);
int refine_reduce_children(const DeviceTree &t, const mnv_render_options &opt, int n_children,
                           const float *results_dev, int result_stride, uint4 *records_dev, cudaStream_t stream);
int refine_commit_records(DeviceTree &t, const mnv_render_options &opt, int n, const uint4 *records_dev,
                          cudaStream_t stream);
int refine_generate_samples(DeviceTree &t, const mnv_render_options &opt, const int32_t *nodes_dev, int m,
                            float *samples_dev, int16_t *cluster_dev, const int32_t *grid_dim,
                            const float *min_position, const float *range, cudaStream_t stream);
int refine_update_samples(DeviceTree &t, const mnv_render_options &opt, const int32_t *nodes_dev, int m,
                          const float *results_dev, int result_stride, cudaStream_t stream);
int refine_prune(DeviceTree &t, const uint8_t *to_delete_dev, const int32_t *index_shifts_dev,
                 int first_shift_index, int64_t num_deleted, cudaStream_t stream);

// ---- fused MLP (mnv_mlp.cu) -------------------------------------------------
struct MlpModel;
MlpModel *mlp_create(const mnv_mlp_desc &d, int device, int *rc_out);
void mlp_destroy(MlpModel *m);
int mlp_forward(const MlpModel *m, const float *x_dev, int64_t rows, int in_dim, float *out_dev,
                int out_stride, cudaStream_t stream);
int mlp_forward_indexed(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, int64_t rows,
                        int in_dim, float *out_dev, int out_stride, cudaStream_t stream);
// rows / offset read on the device: dyn_dev[0] = first slot of row_index_dev, dyn_dev[1] = row count (<= max_rows)
int mlp_forward_bucket(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, const int32_t *dyn_dev,
                       int64_t max_rows, int in_dim, float *out_dev, int out_stride, cudaStream_t stream);
double mlp_flops_per_row(const MlpModel *m);
int mlp_in_dim(const MlpModel *m);
int mlp_out_dim(const MlpModel *m);

int fill_uniform(float *ptr, int64_t n, uint64_t seed, cudaStream_t stream);
int fill_f32(float *ptr, float v, int64_t n, cudaStream_t stream);
int fill_i32(int32_t *ptr, int32_t v, int64_t n, cudaStream_t stream);
int prune_unvisited(DeviceTree &t, int32_t *visited_dev, int64_t *num_deleted, cudaStream_t stream);

// ---- sub-module split across GPUs (mnv_multigpu.cu) ------------------------------------
int launch_signal_peers(uint32_t *const *flag_dst_host, int n, int slot, uint32_t value, cudaStream_t stream);
int launch_composite_partials(const DeviceTree &tree, const mnv_camera &cam, const mnv_render_options &opt,
                              const float *partials_dev, int n, int block, const float *boxes_host,
                              int64_t first_pixel, int n_pixels, uint8_t *rgba_dev, const uint32_t *flags_dev,
                              uint32_t wait_value, cudaStream_t stream, bool guided = false);

// ---- candidate selection (mnv_vote.cu) / sub-module dispatch (mnv_select.cu) ------------------
int select_split_candidates(const float *to_split_dev, int64_t P, int max_n, int32_t *nodes_dev,
                            int *n_selected, int *n_candidates, cudaStream_t stream);
int select_sample_candidates(const float *to_sample_dev, int64_t P, int max_n, int32_t *nodes_dev,
                             int *n_selected, int *n_candidates, cudaStream_t stream);
// rows: tracker [P][3] (may be null); pairs: vote records (id, priority, count) u32 x 3 (may be null)
int select_candidates(int kind, const float *rows_dev, int64_t P, const uint32_t *pairs_dev, int64_t n_pairs, int max_n,
                      int32_t *nodes_dev, int *n_selected, int *n_candidates, cudaStream_t stream);
int vote_reduce(const float *rows_dev, int64_t P, uint32_t *pairs_out_dev, int64_t cap, int64_t *n_out,
                cudaStream_t stream);
// launch / finish halves (one operation in flight per device; see mnv_vote.cu)
int select_candidates_launch(int kind, const float *rows_dev, int64_t P, const uint32_t *pairs_dev, int64_t n_pairs, int max_n,
                             int32_t *nodes_dev, cudaStream_t stream);
int select_candidates_finish(int *n_selected, int *n_candidates, cudaStream_t stream);
int vote_reduce_launch(const float *rows_dev, int64_t P, uint32_t *pairs_out_dev, int64_t cap, cudaStream_t stream);
int vote_reduce_finish(int64_t cap, int64_t *n_out, cudaStream_t stream);
int query_submodules(MlpModel *const *subs, int n_subs, const int16_t *cluster_dev, const float *rows_dev,
                     int in_dim, int64_t V, float *out_dev, int out_stride, cudaStream_t stream);

}  // namespace mnv
