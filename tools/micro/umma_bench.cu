// Dev microbenchmark: time of back-to-back tcgen05.mma (M=128, N, K=16, bf16) from shared memory
// for different smem layouts (SWIZZLE_NONE K-major core matrices vs 128B swizzle), 1 CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t) layout << 61);
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24);
}
__global__ void __launch_bounds__(128, 1) k(int mode, int n, int iters, int a_kcols, long long *out, int commit_every) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[8];
    __shared__ uint32_t tm;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int j = 0; j < 8; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[j])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t *) smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tm;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024 * 2;  // A: 128 KB region, B after it
        const uint32_t id = idesc(128, n);
        uint64_t ads[4], bds[4];
        for (int j = 0; j < 4; ++j) {
            if (mode == 0) {
                ads[j] = desc(a0 + j * 256, 128, (a_kcols / 8) * 128, 0);
                bds[j] = desc(b0 + j * 8192, 128, 256, 0);
            } else {
                ads[j] = desc(a0 + j * 32, 16, 1024, 2);
                bds[j] = desc(b0 + j * 32, 16, 1024, 2);
            }
        }
        if (commit_every == 1 || commit_every == 11) {
            volatile uint32_t *sch = (volatile uint32_t *) (smem + 190 * 1024);
            // a completed barrier to poll: arrive once on bar2[7] -> phase 0 complete
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2[7])) : "memory");
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t w0 = sch[3 * (i & 63)], w1 = sch[3 * (i & 63) + 1], w2 = sch[3 * (i & 63) + 2];
                const uint64_t ad = ads[0] + (w0 & 1) + (w1 & 1), bd = bds[0] + (w2 & 1);
                if (commit_every == 11) {
                    asm volatile("{.reg .pred p; W2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D2; bra W2; D2: }" ::"r"(smem_u32(&bar2[7])) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::
                             "r"(tmem + (i & 1) * 256), "l"(ad), "l"(bd), "r"(id), "r"(1u) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[i % 6])) : "memory");
            }
            long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{.reg .pred p; W3: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D3; bra W3; D3: }" ::"r"(smem_u32(&bar)) : "memory");
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        } else {
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::
                             "r"(tmem + (j & 1) * 256), "l"(ads[j]), "l"(bds[j]), "r"(id), "r"(1u) : "memory");
            if (commit_every == 4 || (commit_every == 2))
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(i >> 2) & 7])) : "memory");
            if (commit_every == 2)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[((i >> 2) + 1) & 7])) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {148})
        for (int mode : {0})
            for (int n : {256, 128})
                for (int akc : {0, 1, 11}) {
                    const int iters = 2000;
                    k<<<grid, 128, 200 * 1024>>>(mode, n, iters, 256, d, akc);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("grid %3d mode %s N %3d commits/4mma(0,1,2) key %3d: issue %.1f clk/mma, complete %.1f clk/mma (%s)\n", grid,
                           mode ? "SW128" : "NONE ", n, akc, (double) h[0] / iters, (double) h[1] / iters, cudaGetErrorString(e));
                }
    return 0;
}
