// Candidate selection for dynamic refinement, entirely on the device.
//
// Replaces the LibTorch glue of Impl::expand_voxels (src/renderer/cuda_renderer.cpp:205-226:
// mask -> unique_dim(counts) -> cat -> mask(count >= 2) -> unique_dim == lexicographic sort by
// (-count, depth, chunk, child) -> first split_batch_size rows) and of Impl::get_more_samples
// (:281-293: mask -> unique_dim == sort by (sample count, chunk, child) -> first rows).
//
// Round 1 did this with a Thrust chain (copy_if -> sort -> reduce_by_key -> copy_if -> sort): two radix sorts of
// up to P 64-bit keys and four implicit host synchronisations, 0.62 ms of a 3 ms frame.  Here:
//
//   1. vote_insert   every tracker row with a candidate votes into an open-addressing hash table keyed by
//                    (priority, leaf id); a warp first merges its equal keys with __match_any_sync (neighbouring
//                    rays mostly vote for the same leaf), so one atomic pair serves the whole group;
//   2. vote_collect  a coalesced sweep over the table, each block over its own share: count, reserve the output
//                    range with ONE global atomic, then write — live slots with enough votes become 64-bit rank
//                    keys — and the table is left clean for the next frame;
//   3. radix select  six 11-bit MSB-first passes (one launch each: histogram, the last block scans) find the max_n-th smallest rank —
//                    keys are unique, so exactly min(max_n, candidates) keys are <= that threshold;
//   4. gather + one-block bitonic sort of those <= max_n keys -> (chunk, child) rows in the reference's order.
//
// No step needs a count on the host: every kernel reads its sizes from a small control block in device
// memory.  The only synchronisation is the last one, which returns n_selected / n_candidates to the caller
// (it sizes the next launches, as `n` does in the reference).
//
// The same machinery serves the multi-GPU path: vote_reduce() turns a rank's tracker rows
// into (id, priority, count) records — typically a quarter of the rays — which are all-gathered instead of
// the raw [P, 3] float rows and merged by select_from_votes() on every replica.
#include <algorithm>
#include <mutex>

#include "mnv_internal.cuh"

namespace mnv {
namespace {

constexpr unsigned long long kEmpty = ~0ull;
constexpr int kBins = 2048, kDigitBits = 11, kPasses = 6;
constexpr int kMaxSortN = 16384;  // keys the one-block bitonic sort holds in shared memory (128 KB)

struct Ctl {
    unsigned long long prefix, prefix_mask, threshold;
    uint32_t n_uniq, n_cand, n_sel, n_pairs, k_rem, take_all, overflow, pad;  // pad: ticket counter of select_pass_kernel
    uint32_t hist[kBins];
};

struct Scratch {
    unsigned long long *keys = nullptr;  // [T] (priority << 32) | leaf id, kEmpty when free
    uint32_t *counts = nullptr;          // [T]
    uint32_t T = 0;
    unsigned long long *cand = nullptr;  // [cand_cap] rank keys
    size_t cand_cap = 0;
    unsigned long long *sel = nullptr;   // [kMaxSortN]
    Ctl *ctl = nullptr;
    uint32_t *host_out = nullptr;  // pinned: n_sel, n_cand, n_uniq, n_pairs, overflow
    std::mutex mu;
};

Scratch &scratch_for_current_device() {
    static Scratch s[16];
    int dev = 0;
    cudaGetDevice(&dev);
    return s[dev & 15];
}

__global__ void fill_u64_kernel(unsigned long long *p, unsigned long long v, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

int ensure(Scratch &s, int64_t voters, cudaStream_t stream) {
    uint32_t T = 1u << 16;
    while ((int64_t) T < 2 * voters && T < (1u << 30)) T <<= 1;
    if (T > s.T) {
        cudaFree(s.keys);
        cudaFree(s.counts);
        s.keys = nullptr;
        s.counts = nullptr;
        s.T = 0;
        MNV_CUDA(cudaMalloc(&s.keys, (size_t) T * sizeof(unsigned long long)));
        MNV_CUDA(cudaMalloc(&s.counts, (size_t) T * sizeof(uint32_t)));
        fill_u64_kernel<<<148 * 8, 256, 0, stream>>>(s.keys, kEmpty, T);
        MNV_CUDA(cudaMemsetAsync(s.counts, 0, (size_t) T * sizeof(uint32_t), stream));
        s.T = T;
    }
    if ((size_t) voters > s.cand_cap) {
        cudaFree(s.cand);
        s.cand = nullptr;
        s.cand_cap = 0;
        MNV_CUDA(cudaMalloc(&s.cand, (size_t) voters * sizeof(unsigned long long)));
        s.cand_cap = (size_t) voters;
    }
    if (!s.sel) MNV_CUDA(cudaMalloc(&s.sel, (size_t) kMaxSortN * sizeof(unsigned long long)));
    if (!s.ctl) MNV_CUDA(cudaMalloc(&s.ctl, sizeof(Ctl)));
    if (!s.host_out) MNV_CUDA(cudaMallocHost(&s.host_out, 8 * sizeof(uint32_t)));
    return MNV_OK;
}

// ---- 1. votes into the hash table -------------------------------------------------------------------------------
__device__ __forceinline__ void table_add(unsigned long long *keys, uint32_t *counts, uint32_t mask, int shift,
                                          unsigned long long key, uint32_t w) {
    uint32_t h = ((uint32_t) key * 0x9E3779B1u) >> shift;  // multiplicative hash of the leaf id
    for (;;) {
        const unsigned long long prev = atomicCAS(keys + h, kEmpty, key);
        if (prev == kEmpty || prev == key) {
            atomicAdd(counts + h, w);
            return;
        }
        h = (h + 1) & mask;
    }
}

// tracker rows are (priority, chunk, child) floats (rt_core.cuh:238-252); chunk < 0 = none
__device__ __forceinline__ unsigned long long row_key(const float *__restrict__ rows, int64_t i) {
    const float chunk_f = rows[3 * i + 1];
    if (!(chunk_f >= 0.f)) return kEmpty;
    const unsigned long long id =
            (unsigned long long) ((long long) tracker_decode_chunk(chunk_f) * 8 + (long long) rows[3 * i + 2]);
    const unsigned long long prio = (unsigned long long) (long long) rows[3 * i + 0];
    return (prio << 32) | id;  // priority (depth / sample count) is a function of the leaf
}

__global__ void __launch_bounds__(256) vote_insert_rows_kernel(const float *__restrict__ rows, int64_t P,
                                                                unsigned long long *keys, uint32_t *counts,
                                                                uint32_t mask, int shift) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long key = i < P ? row_key(rows, i) : kEmpty;
    // equal keys of a warp vote once, with their multiplicity
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key != kEmpty && (int) (threadIdx.x & 31) == __ffs(peers) - 1)
        table_add(keys, counts, mask, shift, key, (uint32_t) __popc(peers));
}

// pre-reduced votes of other ranks: records of three u32 (leaf id, priority, count)
__global__ void __launch_bounds__(256) vote_insert_pairs_kernel(const uint32_t *__restrict__ pairs, int64_t n,
                                                                 unsigned long long *keys, uint32_t *counts,
                                                                 uint32_t mask, int shift) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = pairs[3 * i], prio = pairs[3 * i + 1], c = pairs[3 * i + 2];
    if (c == 0) return;  // padding
    table_add(keys, counts, mask, shift, ((unsigned long long) prio << 32) | id, c);
}

// ---- 2. sweep: live slots -> rank keys (or vote records), table left clean ----------------------------------------
struct U32x3 {
    uint32_t a, b, c;
};

// block-wide exclusive prefix of one value per thread (256 threads); *total = the block's sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    *total = tot;
    return base + incl - v;
}

__device__ __forceinline__ bool vote_qualifies(int kind, uint32_t c) {
    return kind == 0 ? c >= 2   // "< -1" on the negated counts, cuda_renderer.cpp:214
                     : c >= 1;
}

// KIND 0: split ranks (-count, depth, id), votes >= 2; KIND 1: re-sample ranks (sample count, id); KIND 2: records.
// Each block owns one contiguous share of the table and visits it twice — count, then ONE global atomic reserves
// the block's output range, then write + clear — because a same-address atomic per 256 slots (65 536 of them for a
// 4K frame) serialises at the L2: 0.89 ms, against 0.06 ms for the two passes (the second reads L2-resident lines).
template <int KIND>
__global__ void __launch_bounds__(256) vote_collect_kernel(unsigned long long *keys, uint32_t *counts, uint32_t T,
                                                           unsigned long long *cand, U32x3 *pairs, uint32_t pair_cap,
                                                           Ctl *ctl) {
    __shared__ uint32_t s_base;
    const uint32_t share = ((T / gridDim.x + 255) / 256) * 256;  // T and the shares are multiples of 256
    const uint32_t begin = min(blockIdx.x * share, T), end = min(begin + share, T);
    uint32_t mine = 0, live = 0;
    for (uint32_t h = begin + threadIdx.x; h < end; h += 256) {
        if (keys[h] != kEmpty) {
            ++live;
            // every live slot holds at least one vote: only the split selection (votes >= 2) has to look at the count
            if (KIND != 0 || counts[h] >= 2) ++mine;
        }
    }
    uint32_t total = 0, live_total = 0;
    uint32_t at = block_exclusive_scan(mine, &total);
    __syncthreads();
    block_exclusive_scan(live, &live_total);
    if (threadIdx.x == 0) {
        s_base = total ? atomicAdd(KIND == 2 ? &ctl->n_pairs : &ctl->n_cand, total) : 0;
        if (live_total) atomicAdd(&ctl->n_uniq, live_total);
    }
    __syncthreads();
    at += s_base;
    // second visit, four slots per trip: the four key loads, then the live slots' count loads, are independent — a
    // one-slot loop is a chain of dependent global loads (measured 0.31-0.81 ms for the 4K frame's 16.7 M slots)
    for (uint32_t h0 = begin + threadIdx.x; h0 < end; h0 += 4 * 256) {
        unsigned long long k[4];
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t h = h0 + j * 256;
            k[j] = h < end ? keys[h] : kEmpty;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = k[j] != kEmpty ? counts[h0 + j * 256] : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k[j] == kEmpty) continue;
            const uint32_t h = h0 + j * 256;
            keys[h] = kEmpty;
            counts[h] = 0;
            if (!vote_qualifies(KIND, c[j])) continue;
            if (KIND == 2) {
                // a full output drops records (reported through ctl->overflow, the caller falls back to the raw rows)
                if (at < pair_cap) pairs[at] = U32x3{(uint32_t) k[j], (uint32_t) (k[j] >> 32), c[j]};
                else ctl->overflow = 1;
            } else if (KIND == 0) {
                const unsigned long long depth = k[j] >> 32, id = k[j] & 0xffffffffull;
                cand[at] = ((0x3ffffffull - (unsigned long long) min(c[j], 0x3ffffffu)) << 37) | ((depth & 0x3full) << 31) |
                           (id & 0x7fffffffull);
            } else {
                cand[at] = k[j];
            }
            ++at;
        }
    }
}

// ---- 3. radix select of the k-th smallest rank ---------------------------------------------------------------------
__global__ void select_begin_kernel(Ctl *ctl, uint32_t k) {
    ctl->prefix = 0;
    ctl->prefix_mask = 0;
    ctl->k_rem = k;
    ctl->take_all = ctl->n_cand <= k ? 1u : 0u;
    ctl->threshold = kEmpty;
    ctl->n_sel = 0;
}

// One radix-select pass in one launch: every block histograms its share of the live keys' current digit; the block
// that finishes LAST (ticket counter) scans the 2048 bins, fixes the digit whose bucket holds the k_rem-th key and
// clears the histogram for the next pass — no separate one-block "pick" launch between passes.
__global__ void __launch_bounds__(256) select_pass_kernel(const unsigned long long *__restrict__ cand, Ctl *ctl, int shift,
                                                          int last) {
    __shared__ uint32_t h[kBins];
    __shared__ uint32_t s_warp[8], s_last;
    if (ctl->take_all) return;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint32_t n = ctl->n_cand;
    const unsigned long long prefix = ctl->prefix, pmask = ctl->prefix_mask;
    // whole warps iterate together; equal digits of a warp are counted once (the leading digits of the split ranks are
    // identical for almost every key: one shared-memory bin would otherwise take every atomic)
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const unsigned long long k = i < n ? cand[i] : 0;
        const bool in = i < n && (k & pmask) == prefix;
        const uint32_t bin = in ? (uint32_t) (k >> shift) & (kBins - 1) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (in && (int) (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[bin], (uint32_t) __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x)
        if (h[i]) atomicAdd(&ctl->hist[i], h[i]);
    // ---- last block: pick the digit ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->pad, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int t = threadIdx.x;  // 8 consecutive bins per thread
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        c[j] = __ldcg(&ctl->hist[8 * t + j]);
        ctl->hist[8 * t + j] = 0;
        sum += c[j];
    }
    // exclusive prefix of the per-thread sums over the block
    const int lane = t & 31, warp = t >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = incl - sum;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const uint32_t k = ctl->k_rem;
    __syncthreads();
    if (before < k && k <= before + sum) {  // exactly one thread: the k-th key is in one of its bins
        uint32_t below = before;
        int bin = 8 * t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k <= below + c[j]) {
                bin = 8 * t + j;
                break;
            }
            below += c[j];
        }
        const unsigned long long digit_mask = (unsigned long long) (kBins - 1) << shift;
        ctl->prefix |= ((unsigned long long) bin << shift) & digit_mask;
        ctl->prefix_mask |= digit_mask;
        ctl->k_rem = k - below;
        if (last) ctl->threshold = ctl->prefix;
    }
    if (t == 0) ctl->pad = 0;  // ticket counter ready for the next pass
}

__global__ void __launch_bounds__(256) select_gather_kernel(const unsigned long long *__restrict__ cand, Ctl *ctl,
                                                            unsigned long long *sel, uint32_t sel_cap) {
    const uint32_t n = ctl->n_cand;
    const unsigned long long thr = ctl->threshold;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = cand[i];
        if (k <= thr) {
            const uint32_t at = atomicAdd(&ctl->n_sel, 1u);
            if (at < sel_cap) sel[at] = k;
        }
    }
}

// ---- 4. sort the <= max_n survivors, write (chunk, child) rows ------------------------------------------------------
__global__ void __launch_bounds__(1024) select_sort_kernel(const unsigned long long *__restrict__ sel, Ctl *ctl,
                                                           uint32_t sel_cap, int id_bits, int32_t *nodes,
                                                           uint32_t *host_out) {
    extern __shared__ unsigned long long s_keys[];
    const uint32_t n = min(ctl->n_sel, sel_cap);
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) s_keys[i] = i < n ? sel[i] : kEmpty;
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1));  // index with bit `stride` clear
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == up) {
                    s_keys[lo] = b;
                    s_keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    const unsigned long long idm = (1ull << id_bits) - 1;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long id = s_keys[i] & idm;
        nodes[2 * i] = (int32_t) (id >> 3);
        nodes[2 * i + 1] = (int32_t) (id & 7);
    }
    if (threadIdx.x == 0) {
        host_out[0] = n;
        host_out[1] = ctl->n_cand;
        host_out[2] = ctl->n_uniq;
    }
}

__global__ void report_kernel(const Ctl *ctl, uint32_t *host_out) {
    host_out[2] = ctl->n_uniq;
    host_out[3] = ctl->n_pairs;
    host_out[4] = ctl->overflow;
}

int run_select(Scratch &s, int kind, int max_n, int32_t *nodes_dev, cudaStream_t stream) {
    const uint32_t grid = 148 * 8;
    if (kind == 0) vote_collect_kernel<0><<<std::min(grid, s.T / 256), 256, 0, stream>>>(s.keys, s.counts, s.T, s.cand, nullptr, 0, s.ctl);
    else vote_collect_kernel<1><<<std::min(grid, s.T / 256), 256, 0, stream>>>(s.keys, s.counts, s.T, s.cand, nullptr, 0, s.ctl);
    select_begin_kernel<<<1, 1, 0, stream>>>(s.ctl, (uint32_t) max_n);
    for (int p = 0; p < kPasses; ++p) {
        const int shift = std::max(64 - kDigitBits * (p + 1), 0);  // 53, 42, 31, 20, 9, 0 (the last digit overlaps: harmless)
        select_pass_kernel<<<148 * 2, 256, 0, stream>>>(s.cand, s.ctl, shift, p == kPasses - 1);
    }
    select_gather_kernel<<<148 * 2, 256, 0, stream>>>(s.cand, s.ctl, s.sel, (uint32_t) max_n);
    uint32_t m = 1;
    while ((int) m < max_n) m <<= 1;
    const size_t smem = (size_t) m * sizeof(unsigned long long);
    static bool attr_set[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 15]) {
        MNV_CUDA(cudaFuncSetAttribute(select_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kMaxSortN * (int) sizeof(unsigned long long)));
        attr_set[dev & 15] = true;
    }
    select_sort_kernel<<<1, 1024, smem, stream>>>(s.sel, s.ctl, (uint32_t) max_n, kind == 0 ? 31 : 32, nodes_dev, s.host_out);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

int finish_select(Scratch &s, int *n_selected, int *n_candidates, cudaStream_t stream) {
    MNV_CUDA(cudaStreamSynchronize(stream));  // the one synchronisation: n sizes the caller's next launches
    if (n_selected) *n_selected = (int) s.host_out[0];
    if (n_candidates) *n_candidates = (int) s.host_out[1];
    return MNV_OK;
}

int begin(Scratch &s, int64_t voters, cudaStream_t stream) {
    int rc = ensure(s, voters, stream);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaMemsetAsync(s.ctl, 0, sizeof(Ctl), stream));
    return MNV_OK;
}

}  // namespace

// kind 0: split candidates (cuda_renderer.cpp:205-226); kind 1: re-sample candidates (:281-293).
// rows (tracker, may be null) and pairs (vote records of other ranks, may be null) are merged.
// The *_launch / *_finish pairs split every call into "enqueue on this device's stream" and "synchronise + read the counts":
// a host thread that drives several GPUs (mnv_group.cu) launches on all of them before it waits for any.  One operation
// may be in flight per device (its scratch and result words); launch locks the device's scratch, finish unlocks it.
int select_candidates_launch(int kind, const float *rows_dev, int64_t P, const uint32_t *pairs_dev, int64_t n_pairs, int max_n,
                             int32_t *nodes_dev, cudaStream_t stream) {
    if (max_n > kMaxSortN) {
        set_error("select_candidates: max_n %d exceeds %d", max_n, kMaxSortN);
        return MNV_ERR_INVALID;
    }
    Scratch &s = scratch_for_current_device();
    s.mu.lock();
    const int64_t voters = std::max<int64_t>((rows_dev ? P : 0) + (pairs_dev ? n_pairs : 0), 1);
    int rc = begin(s, voters, stream);
    if (rc == MNV_OK) {
        const int shift = 32 - __builtin_ctz(s.T);
        if (rows_dev && P > 0)
            vote_insert_rows_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, stream>>>(rows_dev, P, s.keys, s.counts, s.T - 1, shift);
        if (pairs_dev && n_pairs > 0)
            vote_insert_pairs_kernel<<<(unsigned) ((n_pairs + 255) / 256), 256, 0, stream>>>(pairs_dev, n_pairs, s.keys, s.counts, s.T - 1, shift);
        rc = run_select(s, kind, max_n, nodes_dev, stream);
    }
    if (rc != MNV_OK) s.mu.unlock();
    return rc;
}

int select_candidates_finish(int *n_selected, int *n_candidates, cudaStream_t stream) {
    Scratch &s = scratch_for_current_device();
    const int rc = finish_select(s, n_selected, n_candidates, stream);
    s.mu.unlock();
    return rc;
}

int select_candidates(int kind, const float *rows_dev, int64_t P, const uint32_t *pairs_dev, int64_t n_pairs, int max_n,
                      int32_t *nodes_dev, int *n_selected, int *n_candidates, cudaStream_t stream) {
    const int rc = select_candidates_launch(kind, rows_dev, P, pairs_dev, n_pairs, max_n, nodes_dev, stream);
    return rc != MNV_OK ? rc : select_candidates_finish(n_selected, n_candidates, stream);
}

int select_split_candidates(const float *to_split_dev, int64_t P, int max_n, int32_t *nodes_dev,
                            int *n_selected, int *n_candidates, cudaStream_t stream) {
    return select_candidates(0, to_split_dev, P, nullptr, 0, max_n, nodes_dev, n_selected, n_candidates, stream);
}

int select_sample_candidates(const float *to_sample_dev, int64_t P, int max_n, int32_t *nodes_dev,
                             int *n_selected, int *n_candidates, cudaStream_t stream) {
    return select_candidates(1, to_sample_dev, P, nullptr, 0, max_n, nodes_dev, n_selected, n_candidates, stream);
}

// One rank's tracker rows -> vote records (id, priority, count), unordered.  *n_out = records the rows reduce to;
// MNV_ERR_FULL when that exceeds `cap` (the caller falls back to exchanging raw rows).
int vote_reduce_launch(const float *rows_dev, int64_t P, uint32_t *pairs_out_dev, int64_t cap, cudaStream_t stream) {
    Scratch &s = scratch_for_current_device();
    s.mu.lock();
    int rc = begin(s, std::max<int64_t>(P, 1), stream);
    if (rc == MNV_OK) {
        const int shift = 32 - __builtin_ctz(s.T);
        if (P > 0)
            vote_insert_rows_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, stream>>>(rows_dev, P, s.keys, s.counts, s.T - 1, shift);
        vote_collect_kernel<2><<<std::min(148u * 8u, s.T / 256), 256, 0, stream>>>(
                s.keys, s.counts, s.T, nullptr, reinterpret_cast<U32x3 *>(pairs_out_dev),
                (uint32_t) std::min<int64_t>(cap, 0xffffffffll), s.ctl);
        report_kernel<<<1, 1, 0, stream>>>(s.ctl, s.host_out);
        if (cudaGetLastError() != cudaSuccess) rc = MNV_ERR_CUDA;
    }
    if (rc != MNV_OK) s.mu.unlock();
    return rc;
}

int vote_reduce_finish(int64_t cap, int64_t *n_out, cudaStream_t stream) {
    Scratch &s = scratch_for_current_device();
    int rc = MNV_OK;
    if (cudaStreamSynchronize(stream) != cudaSuccess) rc = MNV_ERR_CUDA;
    if (rc == MNV_OK) {
        if (n_out) *n_out = (int64_t) s.host_out[3];
        if (s.host_out[4]) {
            set_error("vote_reduce: %u records do not fit %lld", s.host_out[3], (long long) cap);
            rc = MNV_ERR_FULL;
        }
    }
    s.mu.unlock();
    return rc;
}

int vote_reduce(const float *rows_dev, int64_t P, uint32_t *pairs_out_dev, int64_t cap, int64_t *n_out, cudaStream_t stream) {
    const int rc = vote_reduce_launch(rows_dev, P, pairs_out_dev, cap, stream);
    return rc != MNV_OK ? rc : vote_reduce_finish(cap, n_out, stream);
}

}  // namespace mnv
