// Device-side replacements for the LibTorch glue around the refinement kernels
// (SURVEY.md §2d): the per-sub-module dispatch of Impl::query_submodules
// (cuda_renderer.cpp:165-203: sort + unique_consecutive + .item() loops + index gather +
// scatter_), the random sample source, and Impl::prune_tree's mask / cumsum (:343-381).
// Candidate selection (unique_dim / sort of expand_voxels / get_more_samples) lives in mnv_vote.cu.
//
// Everything stays on the device; the MLP runs with a row-index indirection so the
// gather / scatter copies of the reference disappear.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/functional.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/transform.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "mnv_internal.cuh"

namespace mnv {
namespace {

template <typename T>
__global__ void fill_kernel(T *p, T v, int64_t n) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// splitmix64 finaliser as a counter-based generator: 24 random bits -> [0, 1)
__global__ void uniform_kernel(float *p, int64_t n, unsigned long long seed) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long) (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = (float) (z >> 40) * (1.0f / 16777216.0f);
}
struct NotVisited {
    const int32_t *visited;
    __device__ uint8_t operator()(long long i) const { return (i > 0 && visited[i] == 0) ? 1 : 0; }
};

}  // namespace

int fill_uniform(float *ptr, int64_t n, uint64_t seed, cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    uniform_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(ptr, n, (unsigned long long) seed);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}
int fill_f32(float *ptr, float v, int64_t n, cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    fill_kernel<float><<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(ptr, v, n);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}
int fill_i32(int32_t *ptr, int32_t v, int64_t n, cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    fill_kernel<int32_t><<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(ptr, v, n);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

// Impl::prune_tree, cuda_renderer.cpp:343-381.
int prune_unvisited(DeviceTree &t, int32_t *visited_dev, int64_t *num_deleted, cudaStream_t stream) {
    auto pol = thrust::cuda::par.on(stream);
    const int64_t cap = t.capacity;
    uint8_t *del = nullptr;
    int32_t *shifts = nullptr;
    MNV_CUDA(cudaMalloc(&del, cap));
    MNV_CUDA(cudaMalloc(&shifts, cap * sizeof(int32_t)));
    int rc = MNV_OK;
    int64_t num = 0;
    try {
        thrust::device_ptr<uint8_t> d(del);
        thrust::device_ptr<int32_t> sh(shifts);
        thrust::transform(pol, thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(cap), d,
                          NotVisited{visited_dev});
        thrust::inclusive_scan(pol, d, d + cap, sh, thrust::plus<int32_t>());
        int32_t last = 0;
        MNV_CUDA(cudaMemcpyAsync(&last, shifts + cap - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        MNV_CUDA(cudaStreamSynchronize(stream));
        num = last;
        if (num > 0) rc = refine_prune(t, del, shifts, 0, num, stream);
        if (rc == MNV_OK && t.max_capacity > 1)  // visit_tracker.slice(0, 1, max).zero_()
            MNV_CUDA(cudaMemsetAsync(visited_dev + 1, 0, (size_t) (t.max_capacity - 1) * sizeof(int32_t), stream));
    } catch (const std::exception &e) {
        set_error("prune_unvisited: %s", e.what());
        cudaGetLastError();
        rc = MNV_ERR_CUDA;
    }
    cudaFree(del);
    cudaFree(shifts);
    if (num_deleted) *num_deleted = num;
    return rc;
}

// ---- scratch arena ---------------------------------------------------------------------
// cudaMalloc / cudaFree per call cost more than the passes they serve (and synchronise the
// device): the temporaries of query_submodules are bump-allocated from one per-device arena that
// grows to the high-water mark after the first frames.
namespace {
struct Arena {
    std::mutex mu;  // one user at a time per device (two host threads may drive one GPU)
    char *base = nullptr;
    size_t cap = 0, off = 0, want = 0;
    std::vector<void *> overflow;
    void *take(size_t n) {
        n = (n + 255) & ~size_t(255);
        want += n;
        if (off + n <= cap) {
            void *p = base + off;
            off += n;
            return p;
        }
        void *p = nullptr;
        if (cudaMalloc(&p, n) != cudaSuccess) throw std::bad_alloc();
        overflow.push_back(p);
        return p;
    }
    // call with the stream idle (every user below synchronises before returning)
    void reset() {
        for (void *p : overflow) cudaFree(p);
        if (!overflow.empty() || want > cap) {
            cudaFree(base);
            cap = want + want / 4 + (1 << 20);
            if (cudaMalloc(&base, cap) != cudaSuccess) {
                base = nullptr;
                cap = 0;
            }
        }
        overflow.clear();
        off = want = 0;
    }
};
Arena &arena_for_current_device() {
    static Arena arenas[16];
    int dev = 0;
    cudaGetDevice(&dev);
    return arenas[dev & 15];
}
struct ArenaScope {  // owns the arena for one call; resets it when the call is over (stream synchronised by then)
    Arena &a;
    cudaStream_t stream;
    ArenaScope(Arena &arena, cudaStream_t s) : a(arena), stream(s) { a.mu.lock(); }
    ~ArenaScope() {
        cudaStreamSynchronize(stream);
        a.reset();
        a.mu.unlock();
    }
};
}  // namespace

// ---- Impl::query_submodules ---------------------------------------------------------------
// Rows are bucketed by sub-module id with a counting pass and a warp-aggregated scatter of the row
// indices (no sort: rows are independent, any order inside a bucket gives the same per-row result);
// the bucket sizes never travel to the host — every sub-module's fused MLP is launched with the
// worst-case grid and reads its (offset, count) pair from device memory — and the MLP reads x /
// writes out through the index, so the reference's gather / scatter_ copies disappear too.
namespace {
constexpr int kMaxSubs = 64;

__global__ void bucket_count_kernel(const int16_t *__restrict__ cluster, int64_t V, int n_subs, int32_t *counts,
                                    int32_t *bad) {
    __shared__ int32_t h[kMaxSubs];
    for (int i = threadIdx.x; i < kMaxSubs; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t) gridDim.x * blockDim.x) {
        const int c = cluster[i];
        if (c < 0 || c >= n_subs) atomicExch(bad, 1);
        else atomicAdd(&h[c], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_subs; i += blockDim.x)
        if (h[i]) atomicAdd(&counts[i], h[i]);
}
// dyn[2*s] = first index slot of bucket s, dyn[2*s+1] = its row count; cursor[s] = dyn[2*s]
__global__ void bucket_offsets_kernel(const int32_t *counts, int n_subs, int32_t *dyn, int32_t *cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int32_t off = 0;
        for (int s = 0; s < n_subs; ++s) {
            dyn[2 * s] = off;
            dyn[2 * s + 1] = counts[s];
            cursor[s] = off;
            off += counts[s];
        }
    }
}
__global__ void bucket_scatter_kernel(const int16_t *__restrict__ cluster, int64_t V, int n_subs, int32_t *cursor,
                                      int32_t *__restrict__ idx) {
    __shared__ int32_t h[kMaxSubs], base[kMaxSubs];
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = threadIdx.x; k < kMaxSubs; k += blockDim.x) h[k] = 0;
    __syncthreads();
    int c = -1, rank = 0;
    if (i < V) {
        c = cluster[i];
        if (c < 0 || c >= n_subs) c = -1;
    }
    if (c >= 0) rank = atomicAdd(&h[c], 1);  // position inside this block's share of the bucket
    __syncthreads();
    for (int k = threadIdx.x; k < n_subs; k += blockDim.x)
        if (h[k]) base[k] = atomicAdd(&cursor[k], h[k]);
    __syncthreads();
    if (c >= 0) idx[base[c] + rank] = (int32_t) i;
}
}  // namespace

int query_submodules(MlpModel *const *subs, int n_subs, const int16_t *cluster_dev, const float *rows_dev,
                     int in_dim, int64_t V, float *out_dev, int out_stride, cudaStream_t stream) {
    if (V <= 0) return MNV_OK;
    if (n_subs > kMaxSubs || V >= (1ll << 31)) {
        set_error("query_submodules: %d sub-modules / %lld rows not supported", n_subs, (long long) V);
        return MNV_ERR_INVALID;
    }
    if (n_subs == 1) {  // nothing to bucket (ids are still validated by the model's clamp-free contract: all 0)
        return mlp_forward(subs[0], rows_dev, V, in_dim, out_dev, out_stride, stream);
    }
    Arena &A = arena_for_current_device();
    ArenaScope scope{A, stream};
    int rc = MNV_OK;
    try {
        auto *idx = static_cast<int32_t *>(A.take(V * sizeof(int32_t)));
        auto *small = static_cast<int32_t *>(A.take((4 * kMaxSubs + 1) * sizeof(int32_t)));
        int32_t *counts = small, *dyn = small + kMaxSubs, *cursor = small + 3 * kMaxSubs, *bad = small + 4 * kMaxSubs;
        MNV_CUDA(cudaMemsetAsync(small, 0, (4 * kMaxSubs + 1) * sizeof(int32_t), stream));
        const int th = 256;
        const unsigned blocks = (unsigned) ((V + th - 1) / th);
        bucket_count_kernel<<<std::min(blocks, 148u * 8u), th, 0, stream>>>(cluster_dev, V, n_subs, counts, bad);
        bucket_offsets_kernel<<<1, 32, 0, stream>>>(counts, n_subs, dyn, cursor);
        bucket_scatter_kernel<<<blocks, th, 0, stream>>>(cluster_dev, V, n_subs, cursor, idx);
        MNV_CUDA(cudaGetLastError());
        for (int s = 0; s < n_subs && rc == MNV_OK; ++s)
            rc = mlp_forward_bucket(subs[s], rows_dev, idx, dyn + 2 * s, V, in_dim, out_dev, out_stride, stream);
        int32_t bad_host = 0;
        MNV_CUDA(cudaMemcpyAsync(&bad_host, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        MNV_CUDA(cudaStreamSynchronize(stream));  // the index lives in the arena: done before it is recycled
        if (rc == MNV_OK && bad_host) {
            set_error("cluster id outside [0, %d)", n_subs);
            rc = MNV_ERR_INVALID;
        }
    } catch (const std::exception &e) {
        set_error("query_submodules: %s", e.what());
        cudaGetLastError();
        rc = MNV_ERR_CUDA;
    }
    return rc;
}

}  // namespace mnv
