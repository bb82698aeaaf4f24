// Dev microbenchmark: latency of one cp.async.bulk (global -> shared, L2-resident source, mbarrier complete_tx) as the issuing
// thread sees it, for 4 / 8 / 16 KiB copies, with one CTA or with one CTA on every SM doing the same, and the sustained rate with
// D copies in flight per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_copy_latency bulk_copy_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(parity) : "memory");
}
// non-blocking poll (mbarrier.test_wait): spins, never suspends
__device__ __forceinline__ void spin_bar(uint32_t bar, uint32_t parity) {
    asm volatile("{.reg .pred p; W: mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(parity) : "memory");
}
template <bool SPIN>
__global__ void __launch_bounds__(32, 1) k(const uint8_t *src, size_t src_bytes, int bytes, int depth, int iters, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t off = ((uint32_t) blockIdx.x * 65536u) & (uint32_t) (src_bytes / 2 - 1);  // power-of-two wrap: no 64-bit modulo in the loop
        uint32_t ph[16] = {0};
        // prime: depth copies in flight
        const long long t0 = clock64();
        for (int i = 0; i < iters + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) {
                if (SPIN) spin_bar(smem_u32(&bar[s]), ph[s]);
                else wait_bar(smem_u32(&bar[s]), ph[s]);
                ph[s] ^= 1;
            }
            if (i < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t) s * bytes)),
                             "l"(src + off), "r"(bytes), "r"(smem_u32(&bar[s])) : "memory");
                off = (off + (uint32_t) bytes) & (uint32_t) (src_bytes / 2 - 1);
            }
        }
        if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
}
int main() {
    const size_t src_bytes = 2 << 20;  // L2-resident like the 1.2 MB of weights
    uint8_t *src;
    long long *d, h;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int spin : {0, 1})
        for (int grid : {1, 148})
            for (int bytes : {4096, 16384})
                for (int depth : {1, 2, 4, 7, 12}) {
                    if ((size_t) depth * bytes > 196 * 1024) continue;
                    const int iters = 2000;
                    if (spin) {
                        k<true><<<grid, 32, 200 * 1024>>>(src, src_bytes, bytes, depth, 16, d);  // warm L2
                        k<true><<<grid, 32, 200 * 1024>>>(src, src_bytes, bytes, depth, iters, d);
                    } else {
                        k<false><<<grid, 32, 200 * 1024>>>(src, src_bytes, bytes, depth, 16, d);
                        k<false><<<grid, 32, 200 * 1024>>>(src, src_bytes, bytes, depth, iters, d);
                    }
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    const double clk = (double) h / iters;
                    printf("%s %3d CTA(s), %5d B per copy, %2d in flight: %.0f clk per copy (%s%.1f B/clk per SM) (%s)\n", spin ? "test_wait spin" : "try_wait      ",
                           grid, bytes, depth, clk, depth == 1 ? "= latency; " : "", bytes / clk, cudaGetErrorString(e));
                }
    return 0;
}
