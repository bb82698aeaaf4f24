// Device tree construction: reference AoS host arrays -> SoA planes in HBM.
// Replaces N3Tree::move_to_device (src/n3tree/n3tree.cpp:207-246); the layout
// is described in mnv_internal.cuh.
#include <algorithm>
#include <cstring>
#include <vector>

#include "mnv_internal.cuh"

namespace mnv {

namespace {

thread_local char g_err[512] = "";

// AoS (child i32[n][8], data f16[n][8][D], counts i16[n][8] or null) -> cell words.
__global__ void build_cells_kernel(const int32_t *__restrict__ child_aos,
                                   const uint16_t *__restrict__ data_aos,
                                   const int16_t *__restrict__ counts_aos, int64_t first_node,
                                   int64_t n_slots, int data_dim, uint32_t *__restrict__ cell,
                                   int16_t *__restrict__ counts_plane) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int64_t node = first_node + (i >> 3);
    const int32_t rel = child_aos[i];
    const int sc = counts_aos ? (int) counts_aos[i] : 8;  // n3tree.cpp:191-193
    const int64_t g = first_node * 8 + i;
    counts_plane[g] = (int16_t) sc;
    if (rel == 0) {
        cell[g] = make_leaf_cell(data_aos[i * data_dim + data_dim - 1], sc);
    } else {
        cell[g] = (uint32_t) (node + rel);
    }
}

// VQ decode (src/n3tree/n3tree.cpp:109-175 restated for the device): one thread per (slot, basis function) writes the
// three channels of that basis function into the staged AoS record; the sigma column by the basis-0 thread.
// map / retained / sigma are the chunk's slices: map[b][n_slots], retained[b][n_slots][3], sigma[n_slots].
__global__ void vq_decode_kernel(const uint16_t *__restrict__ book, const uint16_t *__restrict__ map,
                                 const uint16_t *__restrict__ retained, const uint16_t *__restrict__ sigma,
                                 int64_t n_slots, int n_quant, int n_retain, int data_dim,
                                 uint16_t *__restrict__ data_aos) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int n_basis = n_quant + n_retain;
    if (i >= n_slots * n_basis) return;
    const int64_t slot = i / n_basis;
    const int b = (int) (i - slot * n_basis);
    const uint16_t *c;
    if (b < n_retain) c = retained + ((size_t) b * n_slots + slot) * 3;
    else c = book + ((size_t) (b - n_retain) * 65536 + map[(size_t) (b - n_retain) * n_slots + slot]) * 3;
    uint16_t *dst = data_aos + slot * data_dim + b;
    dst[0] = c[0];
    dst[n_basis] = c[1];
    dst[2 * n_basis] = c[2];
    if (b == 0) data_aos[slot * data_dim + data_dim - 1] = sigma[slot];
}

// One thread per 16-byte chunk of a payload record.
__global__ void build_payload_kernel(const uint16_t *__restrict__ data_aos, int64_t first_slot,
                                     int64_t n_slots, int data_dim, int rec_u4,
                                     uint4 *__restrict__ payload) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots * rec_u4) return;
    const int64_t slot = i / rec_u4;
    const int j = (int) (i % rec_u4);
    const uint16_t *src = data_aos + slot * data_dim + j * 8;
    const int remain = data_dim - j * 8;
    uint16_t h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = k < remain ? src[k] : (uint16_t) 0;
    uint4 v;
    v.x = h[0] | ((uint32_t) h[1] << 16);
    v.y = h[2] | ((uint32_t) h[3] << 16);
    v.z = h[4] | ((uint32_t) h[5] << 16);
    v.w = h[6] | ((uint32_t) h[7] << 16);
    payload[(first_slot + slot) * rec_u4 + j] = v;
}

// SoA -> AoS for download.
__global__ void unpack_kernel(const uint32_t *__restrict__ cell, const uint4 *__restrict__ payload,
                              const int16_t *__restrict__ counts_plane, int64_t first_node,
                              int64_t n_slots, int data_dim, int rec_u4,
                              int32_t *__restrict__ child_aos, uint16_t *__restrict__ data_aos,
                              int16_t *__restrict__ counts_aos) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int64_t g = first_node * 8 + i;
    const int64_t node = first_node + (i >> 3);
    const uint32_t c = cell[g];
    if (child_aos) child_aos[i] = (c & kLeafBit) ? 0 : (int32_t) ((int64_t) c - node);
    if (counts_aos) counts_aos[i] = counts_plane[g];
    if (data_aos) {
        const uint16_t *rec = reinterpret_cast<const uint16_t *>(payload + g * rec_u4);
        for (int k = 0; k < data_dim; ++k) data_aos[i * data_dim + k] = rec[k];
    }
}

}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char *last_error_cstr() { return g_err; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int) e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? MNV_ERR_OOM
           : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? MNV_ERR_NO_DEVICE
                                                                          : MNV_ERR_CUDA;
}

int build_device_tree(DeviceTree &t, const mnv_tree_desc &d, const mnv_vq_desc *vq) {
    const int64_t cap = d.capacity;
    const int D = d.data_dim;
    t.rec_u4 = (D + 7) / 8;
    const int64_t max_slots = t.max_capacity * 8;
    MNV_CUDA(cudaMalloc(&t.cell, max_slots * sizeof(uint32_t)));
    MNV_CUDA(cudaMalloc(&t.payload, max_slots * t.rec_u4 * sizeof(uint4)));
    MNV_CUDA(cudaMalloc(&t.parent, t.max_capacity * sizeof(int32_t)));
    MNV_CUDA(cudaMalloc(&t.sample_counts, max_slots * sizeof(int16_t)));

    MNV_CUDA(cudaMemsetAsync(t.parent, 0, t.max_capacity * sizeof(int32_t), t.stream));
    if (d.parent)
        MNV_CUDA(cudaMemcpyAsync(t.parent, d.parent, cap * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 t.stream));

    // staged, chunked upload of the AoS arrays (bounded scratch: <= ~256 MiB)
    const int64_t chunk_nodes =
            std::max<int64_t>(1, std::min<int64_t>(cap, (256ll << 20) / (8ll * D * 2 + 32 + 16)));
    int32_t *s_child = nullptr;
    uint16_t *s_data = nullptr;
    int16_t *s_counts = nullptr;
    uint16_t *s_book = nullptr, *s_map = nullptr, *s_ret = nullptr, *s_sig = nullptr;  // VQ staging
    MNV_CUDA(cudaMalloc(&s_child, chunk_nodes * 8 * sizeof(int32_t)));
    MNV_CUDA(cudaMalloc(&s_data, chunk_nodes * 8 * D * sizeof(uint16_t)));
    if (d.sample_counts) MNV_CUDA(cudaMalloc(&s_counts, chunk_nodes * 8 * sizeof(int16_t)));
    if (vq) {
        MNV_CUDA(cudaMalloc(&s_book, (size_t) vq->n_quant * 65536 * 3 * sizeof(uint16_t)));
        MNV_CUDA(cudaMalloc(&s_map, (size_t) std::max(vq->n_quant, 1) * chunk_nodes * 8 * sizeof(uint16_t)));
        MNV_CUDA(cudaMalloc(&s_ret, (size_t) std::max(vq->n_retain, 1) * chunk_nodes * 8 * 3 * sizeof(uint16_t)));
        MNV_CUDA(cudaMalloc(&s_sig, (size_t) chunk_nodes * 8 * sizeof(uint16_t)));
        MNV_CUDA(cudaMemcpyAsync(s_book, vq->quant_colors, (size_t) vq->n_quant * 65536 * 3 * sizeof(uint16_t),
                                 cudaMemcpyHostToDevice, t.stream));
    }
    int rc = MNV_OK;
    for (int64_t first = 0; first < cap && rc == MNV_OK; first += chunk_nodes) {
        const int64_t n = std::min(chunk_nodes, cap - first);
        const int64_t slots = n * 8;
        cudaError_t e = cudaMemcpyAsync(s_child, d.child + first * 8, slots * sizeof(int32_t),
                                        cudaMemcpyHostToDevice, t.stream);
        if (e == cudaSuccess && !vq)
            e = cudaMemcpyAsync(s_data, d.data + first * 8 * D, slots * D * sizeof(uint16_t),
                                cudaMemcpyHostToDevice, t.stream);
        if (vq) {
            // the chunk's slices of the compressed arrays, then the decode fills the staged AoS records
            for (int b = 0; b < vq->n_quant && e == cudaSuccess; ++b)
                e = cudaMemcpyAsync(s_map + (size_t) b * slots, vq->quant_map + ((size_t) b * cap + first) * 8,
                                    slots * sizeof(uint16_t), cudaMemcpyHostToDevice, t.stream);
            for (int b = 0; b < vq->n_retain && e == cudaSuccess; ++b)
                e = cudaMemcpyAsync(s_ret + (size_t) b * slots * 3, vq->data_retained + ((size_t) b * cap + first) * 8 * 3,
                                    slots * 3 * sizeof(uint16_t), cudaMemcpyHostToDevice, t.stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(s_sig, vq->sigma + first * 8, slots * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                    t.stream);
            if (e == cudaSuccess) {
                const int64_t work = slots * (vq->n_quant + vq->n_retain);
                vq_decode_kernel<<<(unsigned) ((work + 255) / 256), 256, 0, t.stream>>>(
                        s_book, s_map, s_ret, s_sig, slots, vq->n_quant, vq->n_retain, D, s_data);
            }
        }
        if (e == cudaSuccess && s_counts)
            e = cudaMemcpyAsync(s_counts, d.sample_counts + first * 8, slots * sizeof(int16_t),
                                cudaMemcpyHostToDevice, t.stream);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "tree upload", __FILE__, __LINE__);
            break;
        }
        const int th = 256;
        build_cells_kernel<<<(unsigned) ((slots + th - 1) / th), th, 0, t.stream>>>(
                s_child, s_data, s_counts, first, slots, D, t.cell, t.sample_counts);
        const int64_t n_u4 = slots * t.rec_u4;
        build_payload_kernel<<<(unsigned) ((n_u4 + th - 1) / th), th, 0, t.stream>>>(
                s_data, first * 8, slots, D, t.rec_u4, t.payload);
        e = cudaStreamSynchronize(t.stream);  // staging buffers are reused
        if (e != cudaSuccess) rc = cuda_fail(e, "tree build", __FILE__, __LINE__);
    }
    cudaFree(s_child);
    cudaFree(s_data);
    if (s_counts) cudaFree(s_counts);
    cudaFree(s_book);
    cudaFree(s_map);
    cudaFree(s_ret);
    cudaFree(s_sig);
    return rc;
}

// ---- anchor grid ------------------------------------------------------------------------------------------
namespace {
__global__ void anchor_build_kernel(const uint32_t *__restrict__ cell, int A, uint2 *__restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1u << (3 * A))) return;
    const uint32_t m = (1u << A) - 1;
    const uint32_t iz = idx & m, iy = (idx >> A) & m, ix = idx >> (2 * A);
    uint32_t node = 0;
    for (int l = 0;; ++l) {
        const int bit = A - 1 - l;
        const uint32_t c = (((ix >> bit) & 1u) << 2) | (((iy >> bit) & 1u) << 1) | ((iz >> bit) & 1u);
        const uint32_t slot = node * 8u + c;
        const uint32_t cw = cell[slot];
        if (cw & kLeafBit) {
            // the leaf itself, unless its slot does not fit 28 bits (trees beyond 2^25 nodes): then the node holding it
            out[idx] = slot < (1u << 28) ? make_uint2(cw, ((uint32_t) l << 28) | slot) : make_uint2(node, (uint32_t) l << 28);
            return;
        }
        node = cw;
        if (l == A - 1) {
            out[idx] = make_uint2(node, (uint32_t) A << 28);
            return;
        }
    }
}

__global__ void propagate_visited_kernel(const int32_t *__restrict__ parent, int64_t capacity, int32_t *visited) {
    const int64_t node = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (node <= 0 || node >= capacity || visited[node] == 0) return;
    int64_t n = parent[node] >> 3;
    for (int guard = 0; guard < 32; ++guard) {
        if (visited[n] != 0) break;  // whoever marked it walks (or walked) the rest of the chain
        visited[n] = 1;
        if (n == 0) break;
        n = parent[n] >> 3;
    }
}
}  // namespace

int ensure_anchor(DeviceTree &t, cudaStream_t stream) {
    if (t.anchor_level <= 0) return MNV_OK;
    const size_t n = (size_t) 1 << (3 * t.anchor_level);
    if (!t.anchor) {
        MNV_CUDA(cudaMalloc(&t.anchor, n * sizeof(uint2)));
        t.anchor_dirty = true;
    }
    if (!t.anchor_dirty) return MNV_OK;
    anchor_build_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(t.cell, t.anchor_level, t.anchor);
    MNV_CUDA(cudaGetLastError());
    t.anchor_dirty = false;
    return MNV_OK;
}

void refresh_max_leaf_depth(DeviceTree &t) {
    if (!t.depth_pending || !t.depth_event) return;
    if (cudaEventQuery(t.depth_event) != cudaSuccess) {
        cudaGetLastError();  // not ready: keep the upper bound
        return;
    }
    t.max_leaf_depth = std::min(23, std::max(1, *t.max_depth_host));
    t.depth_pending = false;
}

int launch_propagate_visited(const DeviceTree &t, int32_t *visited, cudaStream_t stream) {
    if (t.capacity <= 1) return MNV_OK;
    propagate_visited_kernel<<<(unsigned) ((t.capacity + 255) / 256), 256, 0, stream>>>(t.parent, t.capacity, visited);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

int download_device_tree(const DeviceTree &t, int64_t first, int64_t count, uint16_t *data,
                         int32_t *child, int32_t *parent, int16_t *sample_counts) {
    if (first < 0 || count < 0 || first + count > t.capacity) {
        set_error("download range [%lld, %lld) outside capacity %lld", (long long) first,
                  (long long) (first + count), (long long) t.capacity);
        return MNV_ERR_INVALID;
    }
    if (count == 0) return MNV_OK;
    const int D = t.data_dim;
    if (parent)
        MNV_CUDA(cudaMemcpy(parent, t.parent + first, count * sizeof(int32_t),
                            cudaMemcpyDeviceToHost));
    const int64_t chunk_nodes =
            std::max<int64_t>(1, std::min<int64_t>(count, (256ll << 20) / (8ll * D * 2 + 32 + 16)));
    int32_t *s_child = nullptr;
    uint16_t *s_data = nullptr;
    int16_t *s_counts = nullptr;
    if (child) MNV_CUDA(cudaMalloc(&s_child, chunk_nodes * 8 * sizeof(int32_t)));
    if (data) MNV_CUDA(cudaMalloc(&s_data, chunk_nodes * 8 * D * sizeof(uint16_t)));
    if (sample_counts) MNV_CUDA(cudaMalloc(&s_counts, chunk_nodes * 8 * sizeof(int16_t)));
    int rc = MNV_OK;
    for (int64_t off = 0; off < count && rc == MNV_OK; off += chunk_nodes) {
        const int64_t n = std::min(chunk_nodes, count - off);
        const int64_t slots = n * 8;
        const int th = 256;
        unpack_kernel<<<(unsigned) ((slots + th - 1) / th), th, 0, t.stream>>>(
                t.cell, t.payload, t.sample_counts, first + off, slots, D, t.rec_u4, s_child, s_data,
                s_counts);
        cudaError_t e = cudaStreamSynchronize(t.stream);
        if (e == cudaSuccess && child)
            e = cudaMemcpy(child + off * 8, s_child, slots * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && data)
            e = cudaMemcpy(data + off * 8 * D, s_data, slots * D * sizeof(uint16_t),
                           cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sample_counts)
            e = cudaMemcpy(sample_counts + off * 8, s_counts, slots * sizeof(int16_t),
                           cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "tree download", __FILE__, __LINE__);
    }
    if (s_child) cudaFree(s_child);
    if (s_data) cudaFree(s_data);
    if (s_counts) cudaFree(s_counts);
    return rc;
}

}  // namespace mnv
