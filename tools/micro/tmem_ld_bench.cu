// Dev microbenchmark: tcgen05.ld throughput (32x32b.x32 = 4 KiB per warp instruction) with W warps reading, with and without
// a stream of M128 N256 K16 MMAs running on the same SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_bench tmem_ld_bench.cu && ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
}
// warps 0..W-1 read (warp w: lane quarter w % 4, columns 256 + 128 * (w / 4) ...); warp 8 issues MMAs into columns 0..255 if mma != 0
__global__ void __launch_bounds__(288, 1) k(int W, int mma, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 288) ((uint32_t *) smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tm;
    if (warp == 8) {
        if (mma && (threadIdx.x & 31) == 0) {
            const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
            const uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (256 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
            const uint64_t ad = desc(a0, 128, 4096), bd = desc(b0, 128, 256);
            const long long t0 = clock64();
            for (int i = 0; i < mma; ++i)
                asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(id) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
            asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar_done)) : "memory");
            if (blockIdx.x == 0) out[1] = clock64() - t0;
        }
    } else if (warp < W) {
        const uint32_t taddr = tmem + ((uint32_t) ((warp & 3) * 32) << 16) + 256 + 128 * (warp >> 2);
        uint32_t r[32], acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            tmem_ld32(taddr + 32 * (i & 3), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[2] = acc; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    long long *d, h[3];
    cudaMalloc(&d, 24);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mma : {0, 1500})
        for (int W : {0, 1, 4, 8}) {
            if (!mma && !W) continue;
            const int iters = 2000;
            cudaMemset(d, 0, 24);
            k<<<148, 288, 200 * 1024>>>(W, mma, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("%d reading warps, %4d MMAs alongside: %.1f clk per tcgen05.ld (4 KiB) per warp", W, mma, (double) h[0] / iters);
            if (W) printf(" = %.0f B/clk per SM", W * 4096.0 * iters / (double) h[0]);
            if (mma) printf("; %.1f clk per MMA", (double) h[1] / mma);
            printf(" (%s)\n", cudaGetErrorString(e));
        }
    return 0;
}
