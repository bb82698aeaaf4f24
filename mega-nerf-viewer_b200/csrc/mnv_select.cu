// Device-side replacements for the LibTorch glue around the refinement kernels
// (SURVEY.md §2d): candidate selection (unique_dim / cat / sort of
// Impl::expand_voxels, cuda_renderer.cpp:205-226, and Impl::get_more_samples,
// :281-293) and the per-sub-module dispatch of Impl::query_submodules (:165-203:
// sort + unique_consecutive + .item() loops + index gather + scatter_).
//
// Everything stays on the device; each call returns one or two small counts to the
// host (the reference syncs once per cluster per batch).  Sorting uses Thrust/CUB
// (library code for a non-hot op); the MLP runs with a row-index indirection so
// the gather / scatter copies of the reference disappear.
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/reduce.h>
#include <thrust/remove.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/binary_search.h>
#include <thrust/unique.h>
#include <thrust/transform.h>
#include <thrust/scan.h>
#include <thrust/functional.h>

#include <algorithm>
#include <vector>

#include "mnv_internal.cuh"

namespace mnv {
namespace {

// tracker rows are (priority, chunk, child) floats (rt_core.cuh:238-252); chunk < 0 = none.
// id = chunk*8 + child identifies the leaf (and therefore its depth / sample count).
struct RowToKey {
    const float *rows;
    __device__ unsigned long long operator()(long long i) const {
        const float chunk = rows[3 * i + 1];
        if (!(chunk >= 0.f)) return ~0ull;
        const unsigned long long id = (unsigned long long) ((long long) chunk * 8 + (long long) rows[3 * i + 2]);
        const unsigned long long prio = (unsigned long long) (long long) rows[3 * i + 0];
        return (prio << 32) | id;  // (priority, id): priority is a function of id
    }
};
struct IsNone {
    __device__ bool operator()(unsigned long long k) const { return k == ~0ull; }
};
struct SplitRank {  // order of unique_dim on rows (-count, depth, chunk, child)
    __device__ unsigned long long operator()(const thrust::tuple<unsigned long long, int> &t) const {
        const unsigned long long key = thrust::get<0>(t);
        const unsigned long long count = (unsigned long long) thrust::get<1>(t);
        const unsigned long long depth = key >> 32, id = key & 0xffffffffull;
        return ((0x3ffffffull - count) << 37) | ((depth & 0x3full) << 31) | (id & 0x7fffffffull);
    }
};
struct CountBelow2 {
    __device__ bool operator()(const thrust::tuple<unsigned long long, int> &t) const {
        return thrust::get<1>(t) < 2;  // "< -1" on the negated counts, cuda_renderer.cpp:214
    }
};
__global__ void write_nodes_kernel(const unsigned long long *ranked, int n, int shift_is_rank,
                                   int32_t *nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long id = shift_is_rank ? (ranked[i] & 0x7fffffffull) : (ranked[i] & 0xffffffffull);
    nodes[2 * i] = (int32_t) (id >> 3);
    nodes[2 * i + 1] = (int32_t) (id & 7);
}

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

int select_split_candidates(const float *to_split_dev, int64_t P, int max_n, int32_t *nodes_dev,
                            int *n_selected, int *n_candidates, cudaStream_t stream) {
    auto pol = thrust::cuda::par.on(stream);
    unsigned long long *keys = nullptr, *ukeys = nullptr;
    int *counts = nullptr;
    MNV_CUDA(cudaMalloc(&keys, P * sizeof(unsigned long long)));
    int rc = MNV_OK;
    try {
        thrust::device_ptr<unsigned long long> k(keys);
        thrust::transform(pol, thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(P), k,
                          RowToKey{to_split_dev});
        const long long valid = thrust::remove_if(pol, k, k + P, IsNone()) - k;
        thrust::sort(pol, k, k + valid);
        MNV_CUDA(cudaMalloc(&ukeys, std::max<long long>(valid, 1) * sizeof(unsigned long long)));
        MNV_CUDA(cudaMalloc(&counts, std::max<long long>(valid, 1) * sizeof(int)));
        thrust::device_ptr<unsigned long long> uk(ukeys);
        thrust::device_ptr<int> cnt(counts);
        const long long uniq = thrust::reduce_by_key(pol, k, k + valid, thrust::constant_iterator<int>(1), uk, cnt).first - uk;
        auto zb = thrust::make_zip_iterator(thrust::make_tuple(uk, cnt));
        const long long kept = thrust::remove_if(pol, zb, zb + uniq, CountBelow2()) - zb;
        // rank = (-count, depth, chunk, child); reuse `keys` for the ranks
        thrust::transform(pol, zb, zb + kept, k, SplitRank());
        thrust::sort(pol, k, k + kept);
        const int n = (int) std::min<long long>(kept, max_n);
        if (n > 0) write_nodes_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, 1, nodes_dev);
        MNV_CUDA(cudaStreamSynchronize(stream));
        if (n_selected) *n_selected = n;
        if (n_candidates) *n_candidates = (int) kept;
    } catch (const std::exception &e) {
        set_error("select_split_candidates: %s", e.what());
        cudaGetLastError();
        rc = MNV_ERR_CUDA;
    }
    cudaFree(keys);
    cudaFree(ukeys);
    cudaFree(counts);
    return rc;
}

int select_sample_candidates(const float *to_sample_dev, int64_t P, int max_n, int32_t *nodes_dev,
                             int *n_selected, int *n_candidates, cudaStream_t stream) {
    auto pol = thrust::cuda::par.on(stream);
    unsigned long long *keys = nullptr;
    MNV_CUDA(cudaMalloc(&keys, P * sizeof(unsigned long long)));
    int rc = MNV_OK;
    try {
        thrust::device_ptr<unsigned long long> k(keys);
        thrust::transform(pol, thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(P), k,
                          RowToKey{to_sample_dev});
        const long long valid = thrust::remove_if(pol, k, k + P, IsNone()) - k;
        thrust::sort(pol, k, k + valid);  // (sample count, chunk, child) == unique_dim order
        const long long uniq = thrust::unique(pol, k, k + valid) - k;
        const int n = (int) std::min<long long>(uniq, max_n);
        if (n > 0) write_nodes_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, 0, nodes_dev);
        MNV_CUDA(cudaStreamSynchronize(stream));
        if (n_selected) *n_selected = n;
        if (n_candidates) *n_candidates = (int) uniq;
    } catch (const std::exception &e) {
        set_error("select_sample_candidates: %s", e.what());
        cudaGetLastError();
        rc = MNV_ERR_CUDA;
    }
    cudaFree(keys);
    return rc;
}

// Impl::query_submodules: rows are grouped by sub-module id with one stable sort of the row
// indices; each sub-module then runs the fused MLP over its index range, reading x and
// writing out through the index (no gather / scatter copies).
int query_submodules(MlpModel *const *subs, int n_subs, const int16_t *cluster_dev, const float *rows_dev,
                     int in_dim, int64_t V, float *out_dev, int out_stride, cudaStream_t stream) {
    if (V <= 0) return MNV_OK;
    auto pol = thrust::cuda::par.on(stream);
    int16_t *keys = nullptr;
    int32_t *idx = nullptr;
    MNV_CUDA(cudaMalloc(&keys, V * sizeof(int16_t)));
    MNV_CUDA(cudaMalloc(&idx, V * sizeof(int32_t)));
    int rc = MNV_OK;
    try {
        thrust::device_ptr<int16_t> k(keys);
        thrust::device_ptr<int32_t> ix(idx);
        MNV_CUDA(cudaMemcpyAsync(keys, cluster_dev, V * sizeof(int16_t), cudaMemcpyDeviceToDevice, stream));
        thrust::sequence(pol, ix, ix + V);
        thrust::stable_sort_by_key(pol, k, k + V, ix);
        std::vector<int64_t> bounds(n_subs + 1);
        for (int s = 0; s <= n_subs; ++s)
            bounds[s] = thrust::lower_bound(pol, k, k + V, (int16_t) s) - k;
        if (bounds[0] != 0 || bounds[n_subs] != V) {
            set_error("cluster id outside [0, %d)", n_subs);
            rc = MNV_ERR_INVALID;
        }
        for (int s = 0; s < n_subs && rc == MNV_OK; ++s) {
            const int64_t cnt = bounds[s + 1] - bounds[s];
            if (cnt > 0)
                rc = mlp_forward_indexed(subs[s], rows_dev, idx + bounds[s], cnt, in_dim, out_dev, out_stride,
                                         stream);
        }
        if (rc == MNV_OK) MNV_CUDA(cudaStreamSynchronize(stream));  // idx / keys are freed below
    } catch (const std::exception &e) {
        set_error("query_submodules: %s", e.what());
        cudaGetLastError();
        rc = MNV_ERR_CUDA;
    }
    cudaFree(keys);
    cudaFree(idx);
    return rc;
}

}  // namespace mnv
