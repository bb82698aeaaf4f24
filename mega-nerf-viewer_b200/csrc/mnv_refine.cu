// Dynamic octree refinement on the SoA device tree.
//
//   A10 add_children_and_generate_samples_kernel  src/cuda/renderer_kernel.cu:170-198
//       + the payload write-back of Impl::expand_voxels  cuda_renderer.cpp:266-275
//   A11 generate_samples_kernel                    renderer_kernel.cu:200-213 (+ :88-168)
//       + the running-mean update of Impl::get_more_samples  cuda_renderer.cpp:318-339
//   A12 adjust_parents_and_children_kernel         renderer_kernel.cu:63-86
//       + the gather compaction of Impl::prune_tree  cuda_renderer.cpp:360-377
//
// Differences from the reference that are deliberate:
//   * the new node's ancestry is taken from the parent list, not re-read from
//     tree.parent while sibling lanes are still writing it (latent race,
//     renderer_kernel.cu:189-197 vs :119);
//   * leaf payload, sigma/count cell word and sample-count plane are written together
//     so the march never sees a half-built leaf;
//   * compaction runs on the device without host round trips per 100k-row chunk.
#include <cuda_fp16.h>

#include <algorithm>

#include "mnv_internal.cuh"
#include "mnv_math.cuh"

namespace mnv {
namespace {

struct ClusterRule {
    int grid0, grid1;
    float min1, min2, range1, range2;
};

__device__ __forceinline__ int cluster_of(const ClusterRule &c, float y, float z) {
    const float g0 = (float) c.grid0, g1 = (float) c.grid1;
    const int a = (int) fmaxf(fminf(__fmul_rn(__fdiv_rn(__fadd_rn(y, -c.min1), c.range1), g0), __fadd_rn(g0, -1.f)), 0.f);
    const int b = (int) fmaxf(fminf(__fmul_rn(__fdiv_rn(__fadd_rn(z, -c.min2), c.range2), g1), __fadd_rn(g1, -1.f)), 0.f);
    return a * c.grid1 + b;
}

// generate_samples_inner, renderer_kernel.cu:88-168.  `packed` = node*8 + child of the voxel;
// for a voxel of a node that is being created, `first_parent` is that node's packed parent
// slot (else -1: read the parent plane).
__device__ __forceinline__ int generate_samples_inner(
        const int32_t *__restrict__ parent, const float *scale, const float *offset,
        const mnv_render_options &opt, float *__restrict__ samples /* [c][rand_dim] */,
        int16_t *__restrict__ cluster /* [c] */, int rand_dim, const ClusterRule &cr, int64_t packed,
        int64_t first_parent) {
    float corners[3] = {0.f, 0.f, 0.f};
    int depth = 0;
    int64_t cur = packed;
    bool first = true;
    for (;;) {
        const int k = (int) (cur & 1), j = (int) ((cur >> 1) & 1), i = (int) ((cur >> 2) & 1);
        const int64_t node = cur >> 3;
        corners[0] = __fdiv_rn(__fadd_rn(corners[0], (float) i), 2.f);
        corners[1] = __fdiv_rn(__fadd_rn(corners[1], (float) j), 2.f);
        corners[2] = __fdiv_rn(__fadd_rn(corners[2], (float) k), 2.f);
        if (node == 0) break;
        cur = (first && first_parent >= 0) ? first_parent : (int64_t) parent[node];
        first = false;
        ++depth;
    }
    const float length_local = __uint_as_float((uint32_t) (127 - depth - 1) << 23);  // pow(N, -depth-1)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float corner = __fdiv_rn(__fadd_rn(corners[a], -offset[a]), scale[a]);
        const float k = __fdiv_rn(length_local, scale[a]);
        for (int s = 0; s < opt.samples_per_corner; ++s)
            samples[s * rand_dim + a] = __fmaf_rn(samples[s * rand_dim + a], k, corner);
    }
    if (opt.need_viewdir) {
        for (int s = 0; s < opt.samples_per_corner; ++s) {
            samples[s * rand_dim + 3] = 1.f;
            samples[s * rand_dim + 4] = 0.f;
            samples[s * rand_dim + 5] = 0.f;
            if (opt.appearance_embedding != -1) samples[s * rand_dim + 6] = (float) opt.appearance_embedding;
        }
    } else if (opt.appearance_embedding != -1) {
        for (int s = 0; s < opt.samples_per_corner; ++s)
            samples[s * rand_dim + 3] = (float) opt.appearance_embedding;
    }
    for (int s = 0; s < opt.samples_per_corner; ++s)
        cluster[s] = (int16_t) cluster_of(cr, samples[s * rand_dim + 1], samples[s * rand_dim + 2]);
    return depth + 1;  // depth of the voxel in the reference's counting (root's children: 1)
}

struct RefineParams {
    uint32_t *cell;
    uint4 *payload;
    int32_t *parent;
    int16_t *counts;
    int rec_u4, data_dim;
    float scale[3], offset[3];
    int64_t capacity;
    mnv_render_options opt;
    ClusterRule cr;
    int *max_depth;  // deepest leaf of the tree (device), raised by add_children
};

__global__ void add_children_kernel(RefineParams p, const int32_t *__restrict__ parent_nodes, int n,
                                    float *__restrict__ samples, int16_t *__restrict__ cluster,
                                    int rand_dim, int32_t *visited) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n * 8) return;
    const int rel = tid >> 3, child = tid & 7;
    const int64_t abs_node = p.capacity + rel;
    const int32_t pn = parent_nodes[2 * rel], pc = parent_nodes[2 * rel + 1];
    const int64_t pslot = (int64_t) pn * 8 + pc;
    if (child == 0) {
        p.cell[pslot] = (uint32_t) abs_node;  // leaf -> internal, absolute child index
        p.parent[abs_node] = (int32_t) pslot;
        if (visited) visited[abs_node] = visited[pn];
    }
    // a new leaf: sigma 0 / count 0 until mnv_tree_commit_children fills it
    p.cell[abs_node * 8 + child] = make_leaf_cell(0, 0);
    const int c = p.opt.samples_per_corner;
    const int depth = generate_samples_inner(p.parent, p.scale, p.offset, p.opt, samples + (size_t) tid * c * rand_dim,
                                             cluster + (size_t) tid * c, rand_dim, p.cr, abs_node * 8 + child, pslot);
    if (child == 0 && p.max_depth) atomicMax(p.max_depth, depth);
}

__global__ void generate_samples_kernel(RefineParams p, const int32_t *__restrict__ nodes, int m,
                                        float *__restrict__ samples, int16_t *__restrict__ cluster,
                                        int rand_dim) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= m) return;
    const int c = p.opt.samples_per_corner;
    generate_samples_inner(p.parent, p.scale, p.offset, p.opt, samples + (size_t) tid * c * rand_dim,
                           cluster + (size_t) tid * c, rand_dim, p.cr,
                           (int64_t) nodes[2 * tid] * 8 + nodes[2 * tid + 1], -1);
}

// new leaf payload = mean over the c MLP outputs (torch::mean_out into the fp16 tensor,
// cuda_renderer.cpp:270), sample_counts = c (:272).  One thread per new leaf slot.
// `records` null: the record goes straight into the tree (slot capacity*8 + tid) together with its cell word and
// count; else only the payload record of child first_child + tid is produced, into records[tid] (multi-GPU: every
// rank reduces the children whose MLP rows it evaluated, the records are all-gathered and committed everywhere).
__global__ void commit_children_kernel(RefineParams p, int n_children, const float *__restrict__ results,
                                       int result_stride, int c, uint4 *__restrict__ records) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_children) return;
    const int64_t slot = p.capacity * 8 + tid;
    const float *r = results + (size_t) tid * c * result_stride;
    __half *rec = reinterpret_cast<__half *>(records ? records + (size_t) tid * p.rec_u4 : p.payload + slot * p.rec_u4);
    __half sig = __float2half(0.f);
    for (int k = 0; k < p.rec_u4 * 8; ++k) {
        float acc = 0.f;
        if (k < p.data_dim)
            for (int s = 0; s < c; ++s) acc += r[s * result_stride + k];
        const __half h = __float2half_rn(acc / (float) c);  // torch::mean: sum / c
        rec[k] = k < p.data_dim ? h : __float2half(0.f);
        if (k == p.data_dim - 1) sig = h;
    }
    if (records) return;
    p.counts[slot] = (int16_t) c;
    p.cell[slot] = make_leaf_cell(__half_as_ushort(sig), c);
}

// payload records produced elsewhere (commit_children_kernel with `records`) -> the tree's new leaves
__global__ void commit_records_kernel(RefineParams p, int n_children, const uint4 *__restrict__ records, int c) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;  // one uint4 of one record
    if (i >= (int64_t) n_children * p.rec_u4) return;
    const int64_t child = i / p.rec_u4;
    const int part = (int) (i - child * p.rec_u4);
    const int64_t slot = p.capacity * 8 + child;
    const uint4 v = records[i];
    p.payload[slot * p.rec_u4 + part] = v;
    const int sig_part = (p.data_dim - 1) / 8, sig_half = (p.data_dim - 1) % 8;
    if (part == sig_part) {
        const uint32_t w = sig_half < 2 ? v.x : sig_half < 4 ? v.y : sig_half < 6 ? v.z : v.w;
        const uint16_t sig = (uint16_t) ((sig_half & 1) ? (w >> 16) : (w & 0xffffu));
        p.counts[slot] = (int16_t) c;
        p.cell[slot] = make_leaf_cell(sig, c);
    }
}

// running mean: data += (sum(new) - c*data) / (count + c); count += c (cuda_renderer.cpp:318-339)
__global__ void update_samples_kernel(RefineParams p, const int32_t *__restrict__ nodes, int m,
                                      const float *__restrict__ results, int result_stride, int c) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= m) return;
    const int64_t slot = (int64_t) nodes[2 * tid] * 8 + nodes[2 * tid + 1];
    const float *r = results + (size_t) tid * c * result_stride;
    __half *rec = reinterpret_cast<__half *>(p.payload + slot * p.rec_u4);
    const int new_count = (int) p.counts[slot] + c;
    __half sig = rec[p.data_dim - 1];
    for (int k = 0; k < p.data_dim; ++k) {
        float acc = 0.f;
        for (int s = 0; s < c; ++s) acc += r[s * result_stride + k];
        const float old = __half2float(rec[k]);
        // the reference forms c*old in fp16 (int * Half tensor) before the fp32 subtraction
        const float c_old = __half2float(__float2half_rn((float) c * old));
        const float upd = (acc - c_old) / (float) new_count;
        const __half h = __float2half_rn(old + __half2float(__float2half_rn(upd)));
        rec[k] = h;
        if (k == p.data_dim - 1) sig = h;
    }
    p.counts[slot] = (int16_t) new_count;
    p.cell[slot] = make_leaf_cell(__half_as_ushort(sig), new_count);
}

// adjust_parents_and_children_kernel, renderer_kernel.cu:63-86, on absolute child links.
__global__ void adjust_links_kernel(RefineParams p, int first_shift_index,
                                    const uint8_t *__restrict__ to_delete,
                                    const int32_t *__restrict__ index_shifts) {
    const int64_t chunk = (int64_t) blockIdx.x * blockDim.x + threadIdx.x + first_shift_index;
    if (chunk >= p.capacity || chunk == 0) return;
    const int32_t pslot = p.parent[chunk];
    const int32_t pn = pslot >> 3;
    if (to_delete[chunk]) {
        // the parent's slot is a leaf again: its payload record and count were kept
        const __half *rec = reinterpret_cast<const __half *>(p.payload + (int64_t) pslot * p.rec_u4);
        p.cell[pslot] = make_leaf_cell(__half_as_ushort(rec[p.data_dim - 1]), p.counts[pslot]);
    } else {
        p.cell[pslot] = (uint32_t) (chunk - index_shifts[chunk]);
        p.parent[chunk] = (int32_t) (((int64_t) pn - index_shifts[pn]) * 8 + (pslot & 7));
    }
}

// Stable compaction of one plane (row_bytes per node) for nodes [begin, end): kept rows are
// gathered into `tmp`, then written at their new index (new = old - shift <= old, so ascending
// chunk order never overwrites unread rows).
__global__ void gather_rows_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ tmp,
                                   int row_u4, int64_t begin, int64_t end,
                                   const uint8_t *__restrict__ to_delete,
                                   const int32_t *__restrict__ shifts) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t node = begin + i / row_u4;
    if (node >= end || to_delete[node]) return;
    const int64_t before = begin > 0 ? shifts[begin - 1] : 0;
    const int64_t dst = (node - shifts[node]) - (begin - before);
    const int j = (int) (i % row_u4);
    reinterpret_cast<uint4 *>(tmp)[dst * row_u4 + j] = reinterpret_cast<const uint4 *>(src)[node * row_u4 + j];
}
__global__ void scatter_rows_kernel(uint8_t *__restrict__ dst_plane, const uint8_t *__restrict__ tmp,
                                    int row_u4, int64_t begin, int64_t end,
                                    const int32_t *__restrict__ shifts) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t before = begin > 0 ? shifts[begin - 1] : 0;
    const int64_t kept = (end - begin) - (shifts[end - 1] - before);
    const int64_t r = i / row_u4;
    if (r >= kept) return;
    const int64_t dst0 = begin - before;
    reinterpret_cast<uint4 *>(dst_plane)[(dst0 + r) * row_u4 + i % row_u4] =
            reinterpret_cast<const uint4 *>(tmp)[i];
}
// 4-byte rows (parent plane)
__global__ void gather_words_kernel(const int32_t *__restrict__ src, int32_t *__restrict__ tmp,
                                    int64_t begin, int64_t end, const uint8_t *__restrict__ to_delete,
                                    const int32_t *__restrict__ shifts) {
    const int64_t node = begin + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= end || to_delete[node]) return;
    const int64_t before = begin > 0 ? shifts[begin - 1] : 0;
    tmp[(node - shifts[node]) - (begin - before)] = src[node];
}
__global__ void scatter_words_kernel(int32_t *__restrict__ dst, const int32_t *__restrict__ tmp,
                                     int64_t begin, int64_t end, const int32_t *__restrict__ shifts) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t before = begin > 0 ? shifts[begin - 1] : 0;
    const int64_t kept = (end - begin) - (shifts[end - 1] - before);
    if (i >= kept) return;
    dst[begin - before + i] = tmp[i];
}

RefineParams make_params(const DeviceTree &t, const mnv_render_options &opt, const int32_t *grid_dim,
                         const float *min_position, const float *range) {
    RefineParams p;
    p.cell = t.cell;
    p.payload = t.payload;
    p.parent = t.parent;
    p.counts = t.sample_counts;
    p.rec_u4 = t.rec_u4;
    p.data_dim = t.data_dim;
    for (int i = 0; i < 3; ++i) {
        p.scale[i] = t.scale[i];
        p.offset[i] = t.offset[i];
    }
    p.capacity = t.capacity;
    p.opt = opt;
    p.max_depth = t.max_depth_dev;
    p.cr.grid0 = grid_dim ? grid_dim[0] : 1;
    p.cr.grid1 = grid_dim ? grid_dim[1] : 1;
    p.cr.min1 = min_position ? min_position[1] : 0.f;
    p.cr.min2 = min_position ? min_position[2] : 0.f;
    p.cr.range1 = range ? range[1] : 1.f;
    p.cr.range2 = range ? range[2] : 1.f;
    return p;
}

int rand_dim_of(const mnv_render_options &opt) {
    return 3 + (opt.need_viewdir ? 3 : 0) + (opt.appearance_embedding != -1 ? 1 : 0);
}

}  // namespace

int refine_add_children(DeviceTree &t, const mnv_render_options &opt, const int32_t *parent_nodes_dev,
                        int n, float *samples_dev, int16_t *cluster_dev, int32_t *visited_dev,
                        const int32_t *grid_dim, const float *min_position, const float *range,
                        cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    refresh_max_leaf_depth(t);
    if (!t.max_depth_dev) {  // first refinement step of this tree
        MNV_CUDA(cudaMalloc(&t.max_depth_dev, sizeof(int)));
        MNV_CUDA(cudaMallocHost(&t.max_depth_host, sizeof(int)));
        MNV_CUDA(cudaEventCreateWithFlags(&t.depth_event, cudaEventDisableTiming));
        *t.max_depth_host = t.max_leaf_depth;
        MNV_CUDA(cudaMemcpyAsync(t.max_depth_dev, t.max_depth_host, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    if (t.capacity + n > t.max_capacity) {
        set_error("Full: capacity %lld + %d > max %lld", (long long) t.capacity, n, (long long) t.max_capacity);
        return MNV_ERR_FULL;  // cuda_renderer.cpp:228-231
    }
    const RefineParams p = make_params(t, opt, grid_dim, min_position, range);
    const int th = 256;
    add_children_kernel<<<(n * 8 + th - 1) / th, th, 0, stream>>>(p, parent_nodes_dev, n, samples_dev,
                                                                  cluster_dev, rand_dim_of(opt), visited_dev);
    MNV_CUDA(cudaGetLastError());
    t.pending_children = n;
    t.anchor_dirty = true;
    // each split deepens the tree by at most one level: the bound holds until the exact value (the kernel's
    // atomicMax over the new leaves' depths) has reached the host — no synchronisation here
    t.max_leaf_depth = std::min(23, t.max_leaf_depth + 1);
    if (t.max_depth_host && t.depth_event) {
        MNV_CUDA(cudaMemcpyAsync(t.max_depth_host, t.max_depth_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MNV_CUDA(cudaEventRecord(t.depth_event, stream));
        t.depth_pending = true;
    }
    return MNV_OK;
}

int refine_commit_children(DeviceTree &t, const mnv_render_options &opt, int n, const float *results_dev,
                           int result_stride, cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    if (n != t.pending_children) {
        set_error("commit of %d children, %d pending", n, t.pending_children);
        return MNV_ERR_INVALID;
    }
    const RefineParams p = make_params(t, opt, nullptr, nullptr, nullptr);
    const int th = 128;
    commit_children_kernel<<<(n * 8 + th - 1) / th, th, 0, stream>>>(p, n * 8, results_dev, result_stride,
                                                                     opt.samples_per_corner, nullptr);
    MNV_CUDA(cudaGetLastError());
    t.capacity += n;  // cuda_renderer.cpp:275
    t.pending_children = 0;
    t.anchor_dirty = true;
    return MNV_OK;
}

// Multi-GPU refinement: the mean -> fp16 payload records of n_children new leaves whose MLP
// rows this rank evaluated (results_dev holds exactly those children's rows); the tree is not touched.
int refine_reduce_children(const DeviceTree &t, const mnv_render_options &opt, int n_children,
                           const float *results_dev, int result_stride, uint4 *records_dev, cudaStream_t stream) {
    if (n_children <= 0) return MNV_OK;
    const RefineParams p = make_params(t, opt, nullptr, nullptr, nullptr);
    const int th = 128;
    commit_children_kernel<<<(n_children + th - 1) / th, th, 0, stream>>>(p, n_children, results_dev, result_stride,
                                                                          opt.samples_per_corner, records_dev);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

// ... and the gathered records of all n * 8 children committed into this replica.
int refine_commit_records(DeviceTree &t, const mnv_render_options &opt, int n, const uint4 *records_dev,
                          cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    if (n != t.pending_children) {
        set_error("commit of %d children, %d pending", n, t.pending_children);
        return MNV_ERR_INVALID;
    }
    const RefineParams p = make_params(t, opt, nullptr, nullptr, nullptr);
    const int64_t work = (int64_t) n * 8 * t.rec_u4;
    commit_records_kernel<<<(unsigned) ((work + 255) / 256), 256, 0, stream>>>(p, n * 8, records_dev,
                                                                               opt.samples_per_corner);
    MNV_CUDA(cudaGetLastError());
    t.capacity += n;
    t.pending_children = 0;
    t.anchor_dirty = true;
    return MNV_OK;
}

int refine_generate_samples(DeviceTree &t, const mnv_render_options &opt, const int32_t *nodes_dev, int m,
                            float *samples_dev, int16_t *cluster_dev, const int32_t *grid_dim,
                            const float *min_position, const float *range, cudaStream_t stream) {
    if (m <= 0) return MNV_OK;
    const RefineParams p = make_params(t, opt, grid_dim, min_position, range);
    const int th = 256;
    generate_samples_kernel<<<(m + th - 1) / th, th, 0, stream>>>(p, nodes_dev, m, samples_dev, cluster_dev,
                                                                  rand_dim_of(opt));
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

int refine_update_samples(DeviceTree &t, const mnv_render_options &opt, const int32_t *nodes_dev, int m,
                          const float *results_dev, int result_stride, cudaStream_t stream) {
    if (m <= 0) return MNV_OK;
    const RefineParams p = make_params(t, opt, nullptr, nullptr, nullptr);
    const int th = 128;
    update_samples_kernel<<<(m + th - 1) / th, th, 0, stream>>>(p, nodes_dev, m, results_dev, result_stride,
                                                                opt.samples_per_corner);
    MNV_CUDA(cudaGetLastError());
    t.anchor_dirty = true;
    return MNV_OK;
}

int refine_prune(DeviceTree &t, const uint8_t *to_delete_dev, const int32_t *index_shifts_dev,
                 int first_shift_index, int64_t num_deleted, cudaStream_t stream) {
    if (num_deleted <= 0) return MNV_OK;
    t.anchor_dirty = true;
    mnv_render_options dummy{};
    const RefineParams p = make_params(t, dummy, nullptr, nullptr, nullptr);
    const int th = 256;
    const int64_t cap = t.capacity;
    const int64_t span = cap - first_shift_index;
    adjust_links_kernel<<<(unsigned) ((span + th - 1) / th), th, 0, stream>>>(p, first_shift_index,
                                                                              to_delete_dev, index_shifts_dev);
    MNV_CUDA(cudaGetLastError());
    // chunked stable compaction of the four planes
    const int64_t chunk = 1 << 16;
    const int pay_u4 = 8 * t.rec_u4;  // uint4 per node in the payload plane
    uint8_t *tmp = nullptr;
    MNV_CUDA(cudaMalloc(&tmp, (size_t) chunk * pay_u4 * 16));
    int rc = MNV_OK;
    for (int64_t b = first_shift_index; b < cap; b += chunk) {
        const int64_t e = std::min(cap, b + chunk);
        const int64_t nn = e - b;
        struct {
            uint8_t *plane;
            int row_u4;
        } planes[3] = {{reinterpret_cast<uint8_t *>(t.cell), 2},
                       {reinterpret_cast<uint8_t *>(t.payload), pay_u4},
                       {reinterpret_cast<uint8_t *>(t.sample_counts), 1}};
        for (auto &pl : planes) {
            const int64_t work = nn * pl.row_u4;
            gather_rows_kernel<<<(unsigned) ((work + th - 1) / th), th, 0, stream>>>(
                    pl.plane, tmp, pl.row_u4, b, e, to_delete_dev, index_shifts_dev);
            scatter_rows_kernel<<<(unsigned) ((work + th - 1) / th), th, 0, stream>>>(
                    pl.plane, tmp, pl.row_u4, b, e, index_shifts_dev);
        }
        gather_words_kernel<<<(unsigned) ((nn + th - 1) / th), th, 0, stream>>>(
                t.parent, reinterpret_cast<int32_t *>(tmp), b, e, to_delete_dev, index_shifts_dev);
        scatter_words_kernel<<<(unsigned) ((nn + th - 1) / th), th, 0, stream>>>(
                t.parent, reinterpret_cast<const int32_t *>(tmp), b, e, index_shifts_dev);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) {
            rc = cuda_fail(err, "prune compaction", __FILE__, __LINE__);
            break;
        }
    }
    cudaError_t err = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    if (rc == MNV_OK && err != cudaSuccess) rc = cuda_fail(err, "prune", __FILE__, __LINE__);
    if (rc == MNV_OK) t.capacity -= num_deleted;  // cuda_renderer.cpp:377
    return rc;
}

}  // namespace mnv
