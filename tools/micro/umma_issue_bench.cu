// Dev microbenchmark: what does the issuing lane pay per tcgen05.mma (M=128 or 256 with cta_group::2, N=256, K=16, bf16,
// SWIZZLE_NONE K-major) when G MMAs share one tcgen05.commit (and optionally one mbarrier poll)?  One CTA (or CTA pair) per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_issue_bench umma_issue_bench.cu && ./umma_issue_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24);
}
template <bool PAIR>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    if constexpr (PAIR)
        asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void commit(uint32_t bar) {
    if constexpr (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t) 3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(parity) : "memory");
}
// G MMAs + one commit per trip; POLL: a poll of an already-complete barrier in front of every trip
template <bool PAIR, int G, bool POLL>
__global__ void __launch_bounds__(128, 1) k(int n, int trips, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_done, bar_ring[8], bar_ready;
    __shared__ uint32_t tm;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ready)));
        for (int j = 0; j < 8; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ring[j])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");  // phase 0 complete
    }
    if (threadIdx.x < 32) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t *) smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tm;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
        const uint32_t id = idesc(PAIR ? 256 : 128, n);
        uint64_t ad = desc(a0, 128, 4096), bd = desc(b0, 128, 256);
        const uint32_t ring0 = smem_u32(&bar_ring[0]), ready = smem_u32(&bar_ready);
        const long long t0 = clock64();
        uint32_t r = 0;
        for (int i = 0; i < trips; ++i) {
            if (POLL) {
                wait_bar(ready, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < G; ++j) mma<PAIR>(tmem + (i & 1) * 256, ad + 16 * j, bd + (PAIR ? 256 : 512) * j, id);
            commit<PAIR>(ring0 + 8 * r);
            r = (r + 1) & 7;
        }
        const long long t1 = clock64();
        commit<PAIR>(smem_u32(&bar_done));
        wait_bar(smem_u32(&bar_done), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    if (threadIdx.x < 32) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}
template <bool PAIR, int G, bool POLL>
void run(int n, long long *d) {
    const int trips = 4096 / G;
    cudaFuncSetAttribute(k<PAIR, G, POLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = PAIR ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k<PAIR, G, POLL>, n, trips, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s N %3d: %d MMAs per commit%s: issue %.1f clk/MMA, complete %.1f clk/MMA (%s)\n", PAIR ? "cta_group::2 M256" : "cta_group::1 M128", n, G,
           POLL ? " + poll" : "", (double) h[0] / (trips * G), (double) h[1] / (trips * G), cudaGetErrorString(e));
}
int main() {
    long long *d;
    cudaMalloc(&d, 16);
    for (int n : {256, 128}) {
        run<false, 1, false>(n, d); run<false, 2, false>(n, d); run<false, 4, false>(n, d); run<false, 8, false>(n, d);
        run<false, 1, true>(n, d); run<false, 2, true>(n, d); run<false, 4, true>(n, d);
        run<true, 1, false>(n, d); run<true, 2, false>(n, d); run<true, 4, false>(n, d); run<true, 8, false>(n, d);
        run<true, 1, true>(n, d); run<true, 2, true>(n, d); run<true, 4, true>(n, d);
    }
    return 0;
}
