// Fused Mega-NeRF sub-MLP forward on tcgen05 / TMEM (sm_100a).
//
// Replaces the TorchScript call at the reference's query_submodules
// (src/renderer/cuda_renderer.cpp:188-193: nerfs[i].forward({input, false}) under fp16
// autocast).  The architecture is NOT in /root/reference (un-pinned TorchScript
// artefact of cmusatyalab/mega-nerf); the shapes are the ones BASELINE.json /
// SURVEY.md §8 A9 name: PE(xyz, 12 freqs) = 75 -> 8 x [Linear 256 + ReLU] with the
// 75-vector re-concatenated in front of layer 4 (K = 331) -> sigma head 256 -> 1;
// feature Linear 256 -> 256; head Linear(256 [+27 dir PE] [+48 appearance]) -> 128 +
// ReLU; Linear 128 -> 3*basis.  Output row = [rgb / SH coefficients, sigma].
//
// One CTA owns TWO 128-row tiles and ping-pongs them through the 11 GEMMs so that the
// tensor pipe works on one tile while the epilogue warps drain the other:
//     tensor pipe : T0.L0  T1.L0  T0.L1  T1.L1  T0.L2 ...
//     epilogue    :        T0.L0  T1.L0  T0.L1  T1.L1 ...
// A (activations, bf16) lives in shared memory in the UMMA K-major no-swizzle core-matrix
// layout and is rewritten in place by the epilogue; B (weights, bf16, pre-packed on the
// host into the same layout, 16 K-columns = 8 KiB per chunk) is streamed from L2 by 1-D
// bulk TMA copies (cp.async.bulk -> UBLKCP) through an mbarrier ring whose stages hold PER
// chunks; D accumulates in TMEM (2 x 256 fp32 columns = all 512) via tcgen05.mma issued by
// one thread (N <= 256, K = 16: 128 clk each, measured); the epilogue reads D back with
// tcgen05.ld, applies ReLU while packing to bf16 (cvt.rn.relu.bf16x2.f32) and writes the
// next layer's A operand.  By default two CTAs of a cluster (one TPC) form a PAIR: the MMAs
// are cta_group::2 (M = 256 over both SMs, issued by the leader CTA), each CTA holds its own
// rows and HALF of every weight chunk (template parameters of mlp_forward_kernel).
//   warps 0-7 : epilogue of whichever tile completed — warp w owns TMEM lanes
//               32(w%4)..+31 (the hardware's lane-quarter rule) and column half w/4;
//               also the positional encodings of the next 256 rows
//   warp  8   : TMA producer (one elected lane)
//   warp  9   : TMEM allocation; MMA issue (one elected lane) in the leader CTA, "my half
//               has landed" relay in the peer CTA of a pair
// Biases are added by the epilogue from an fp32 row staged in shared memory (dev switch
// MNV_MLP_BIAS_EPILOGUE: let them ride in the GEMM as bf16 hi + lo against two columns of
// ones of the PE operand instead).  The appearance embedding enters head 1 as a per-index
// fp32 bias row (W_app . emb[i] + b, tabulated at model load with bf16-rounded operands)
// instead of 48 more K columns.
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mnv_internal.cuh"

// -DMNV_MLP_TIMING (EXTRA=-DMNV_MLP_TIMING make) compiles cycle counters into the issuing lane; with
// MNV_MLP_DEBUG=1 in the environment every launch then prints where that lane waited.
#ifdef MNV_MLP_TIMING
#define MLP_T(...) __VA_ARGS__
#else
#define MLP_T(...)
#endif
// Dev-only diagnostic builds (results are WRONG, timing only): bit 0 = the epilogue computes but does not store the next
// A operand, bit 1 = the producer signals a stage without copying weights into it, bit 2 = the epilogue does not read the
// accumulators out of TMEM (its registers hold a stand-in value).  They separate shared-memory
// bandwidth contention from the issuing lane's own instruction latency (profiles/r2_mlp_smem_diag.md).
#ifndef MNV_MLP_DIAG
#define MNV_MLP_DIAG 0
#endif

namespace mnv {
namespace {

constexpr int kMlpThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kTileM = 128;
constexpr int kTiles = 2;  // row tiles per CTA
constexpr int kMaxStages = 13;  // ring stages: whatever shared memory is left holds (mlp_stages)
constexpr int kChunkK = 16;
constexpr int kMaxN = 256;
constexpr int kStageBytes = kMaxN * kChunkK * 2;  // 8 KiB
constexpr int kActK = 256, kPeK = 80, kDirK = 32;
constexpr int kActBytes = kTileM * kActK * 2;  // 64 KiB
constexpr int kPeBytes = kTileM * kPeK * 2;    // 20 KiB
constexpr int kDirBytes = kTileM * kDirK * 2;  // 8 KiB
constexpr int kActSBO = (kActK / 8) * 128, kPeSBO = (kPeK / 8) * 128, kDirSBO = (kDirK / 8) * 128;
constexpr int kTmemCols = 512;
constexpr int kMaxLayers = 16, kMaxChunks = 256;
constexpr int kLayerWords = 12;  // smem copy of one layer's issue schedule

enum { kEpiReluAct = 0, kEpiReluActSigma = 1, kEpiLinearAct = 2, kEpiOut = 3, kEpiReluActApp = 4 };
enum { kSrcAct = 0, kSrcPE = 1, kSrcDir = 2 };

struct ChunkDesc {
    uint32_t gmem_off;  // byte offset of the packed chunk in the weight blob
    uint16_t n;         // padded N of the layer
    uint16_t a_src;     // which A buffer
    uint16_t a_k0;      // first K column inside that buffer
    uint16_t pad;
};

struct SegDesc {  // a run of consecutive 16-column chunks of one A buffer
    uint16_t a_src, a_k0, n_chunks, pad;
};
struct LayerDesc {
    uint16_t n;        // padded N (multiple of 32)
    uint16_t n_real;   // real output width
    uint16_t chunk_begin, chunk_end;
    uint16_t n_segs, pad;
    SegDesc segs[3];
    uint32_t epilogue;
    uint32_t bias_off;  // float offset of an fp32 bias row the epilogue adds; kNoBias: the bias rides in the GEMM
};
constexpr uint32_t kNoBias = 0xffffffffu;

struct MlpSchedule {
    int n_layers, n_chunks;
    LayerDesc layers[kMaxLayers];
    ChunkDesc chunks[kMaxChunks];
};

struct MlpParams {
    const uint8_t *__restrict__ weights;   // packed bf16 chunks, schedule order
    const float *__restrict__ sigma_w;     // [256] sigma head weights, then its bias
    const float *__restrict__ biases;      // fp32 bias rows added by the epilogue (layers with bias_off != kNoBias)
    const float *__restrict__ app_bias;    // [n_appearance][head_n] head-1 bias rows incl. appearance term (may be null)
    const MlpSchedule *__restrict__ sched;
    const float *__restrict__ x;           // [rows][in_dim]
    float *__restrict__ out;               // [rows][out_stride]
    const int32_t *__restrict__ row_index; // optional: logical row i reads x / writes out at row_index[i]
    const int32_t *__restrict__ dyn;       // optional: {first slot of row_index, row count} decided on the device
    int64_t rows;
    int in_dim, out_stride;
    int n_groups;  // groups of kTiles * kTileM rows
    int pe_xyz_freqs, pe_dir_freqs, ones_col;
    int need_viewdir, n_appearance, app_col, head_n;  // app_col: column of x holding the index (-1: none)
    int sigma_activation;          // 0 = ReLU, 1 = softplus
    int n_stages;                  // depth of the weight ring
    int out_real;                  // 3 * basis
    long long *dbg;                // per-CTA cycle counters, only in -DMNV_MLP_TIMING builds (MNV_MLP_DEBUG=1), else null
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t"
            "}" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
}
// one poll; may suspend the thread for a hardware-defined time if the phase has not completed yet.  The same form serves
// barriers that take arrivals from the peer CTA: cluster-scope acquire / release qualifiers on the poll and on the remote
// arrive cost a full fence each (measured: 570 clk per ring stage on the issuing lane) and buy nothing here — what is
// handed over sits in shared memory, written through the async proxy or fenced with fence.proxy.async by its writer.
template <bool CLUSTER>
__device__ __forceinline__ bool mbar_try(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of a shared::cta address as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(smem_u32(dst)),
            "l"(src), "r"(bytes), "r"(smem_u32(bar))
            : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// PAIR: the arrival is delivered to the barrier at this offset in BOTH CTAs of the pair
template <bool PAIR>
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_addr) {
    if constexpr (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
                     "h"((uint16_t) 3)
                     : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr)
                     : "memory");
}
// UMMA shared-memory descriptors (K-major, SWIZZLE_NONE, cute::UMMA::SmemDescriptor layout) are assembled from
// two 32-bit halves in the issue loop: low = start>>4 [0,14) | LBO>>4 [16,30); high = SBO>>4 [0,14) | version 1 [14,16).
// Instruction descriptor kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) |
           ((uint32_t) (m >> 4) << 24);
}
template <bool PAIR>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR)
        asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}" ::"r"(tmem_d),
                "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                : "memory");
    else
        asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}" ::"r"(tmem_d),
                "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                : "memory");
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// asynchronous: the registers are valid after tmem_wait()
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
#if MNV_MLP_DIAG & 4
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = taddr + j;
    return;
#endif
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
              "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {  // a -> low half
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {  // max(x, 0) fused into the conversion
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

// Byte offset of element (row, k) in an A buffer with K-major core-matrix layout.
__device__ __forceinline__ uint32_t a_off(int row, int k, int sbo) {
    return (uint32_t) ((row >> 3) * sbo + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

// sin / cos of a = 2^o * v: two-constant Cody-Waite reduction to [-pi, pi] (a is exact, the
// reduction error is ~1e-11 * |a|), then the SFU approximations (abs error 2^-21.4 there) —
// two orders below the bf16 rounding the value goes through next.
__device__ __forceinline__ void sincos_reduced(float a, float &s, float &c) {
    const float k = rintf(a * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, a);
    r = fmaf(k, 1.7484555e-07f, r);
    s = __sinf(r);
    c = __cosf(r);
}

__device__ __forceinline__ void put_bf16(uint8_t *buf, int row, int k, int sbo, float val) {
    *reinterpret_cast<__nv_bfloat16 *>(buf + a_off(row, k, sbo)) = __float2bfloat16_rn(val);
}

// Octaves [o0, o1) of the positional encoding of v[3] (NeRF "Embedding", include-input:
// [v, sin(2^0 v), cos(2^0 v), sin(2^1 v), cos(2^1 v), ...]) into columns 3 + 6*o ... of `buf`.
__device__ __forceinline__ void write_pe_octaves(uint8_t *buf, int row, int sbo, const float *v, int o0, int o1) {
    float f = exp2f((float) o0);
    for (int o = o0; o < o1; ++o) {
        const int k = 3 + 6 * o;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincos_reduced(f * v[c], sn, cs);
            put_bf16(buf, row, k + c, sbo, sn);
            put_bf16(buf, row, k + 3 + c, sbo, cs);
        }
        f *= 2.f;
    }
}

// Epilogue of one (layer, tile) for the layers that produce the next A operand, straight-line for a compile-time kind
// and column count: accumulators out of TMEM 32 columns at a time (the next load is in flight while this one is used),
// + fp32 bias row (shared memory, or the per-appearance-index row in global memory), [sigma-head partial dot product],
// ReLU fused into the bf16 conversion, four 16-byte stores into the K-major core-matrix layout.  The generic loop this
// replaces spent 3.5 instructions per useful one (register copies of every accumulator, per-block branches on the layer
// kind: profiles/r2_mlp_smem_diag.md) and its 2500 clk per unit, not the tensor pipe's 2048, paced the kernel.
//   taddr: TMEM address of this thread's first column; dst: &A[row][first column]; bias: &row[first column]
template <int KIND, int N_MINE>
__device__ __forceinline__ void epilogue_unit(const uint32_t taddr, uint8_t *__restrict__ dst, const float *__restrict__ bias,
                                              const float *__restrict__ sigma_w, float &sigma) {
    constexpr int kBlocks = N_MINE / 32;
    uint32_t r[2][32];
    tmem_ld32(taddr, r[0]);
#pragma unroll
    for (int blk = 0; blk < kBlocks; ++blk) {
        float4 b[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float4 *b4 = reinterpret_cast<const float4 *>(bias + 32 * blk) + g;
            b[g] = KIND == kEpiReluActApp ? __ldg(b4) : *b4;
        }
        float4 sw[8];
        if constexpr (KIND == kEpiReluActSigma) {
#pragma unroll
            for (int g = 0; g < 8; ++g) sw[g] = __ldg(reinterpret_cast<const float4 *>(sigma_w + 32 * blk) + g);
        }
        tmem_wait();
        if (blk + 1 < kBlocks) tmem_ld32(taddr + 32 * (blk + 1), r[(blk + 1) & 1]);  // overlaps the math below
        const uint32_t(&rr)[32] = r[blk & 1];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = __uint_as_float(rr[8 * g + j]) + (&b[2 * g].x)[j];
                v[4 + j] = __uint_as_float(rr[8 * g + 4 + j]) + (&b[2 * g + 1].x)[j];
            }
            if constexpr (KIND == kEpiReluActSigma) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    sigma = fmaf(fmaxf(v[j], 0.f), (&sw[2 * g].x)[j], sigma);
                    sigma = fmaf(fmaxf(v[4 + j], 0.f), (&sw[2 * g + 1].x)[j], sigma);
                }
            }
            uint4 o;
            if constexpr (KIND == kEpiLinearAct) {
                o.x = pack_bf16(v[0], v[1]);
                o.y = pack_bf16(v[2], v[3]);
                o.z = pack_bf16(v[4], v[5]);
                o.w = pack_bf16(v[6], v[7]);
            } else {
                o.x = pack_bf16_relu(v[0], v[1]);
                o.y = pack_bf16_relu(v[2], v[3]);
                o.z = pack_bf16_relu(v[4], v[5]);
                o.w = pack_bf16_relu(v[6], v[7]);
            }
#if MNV_MLP_DIAG & 1
            if (sigma_w == nullptr)  // never true: keeps the arithmetic, drops the store traffic
#endif
            *reinterpret_cast<uint4 *>(dst + (4 * blk + g) * 128) = o;  // 8 columns = one 16-byte row of a core matrix
        }
    }
}

// PAIR: two CTAs of a cluster (one TPC) run tcgen05.mma.cta_group::2 — one lane of the leader CTA issues M = 256 MMAs
// that span both SMs; each CTA keeps its own 2 x 128 rows (A operand, accumulators) and HALF of every weight chunk
// (its N/2 rows of B), so the same ring holds twice as many MMAs of look-ahead and each SM reads half the weight bytes.
// PER: weight chunks (MMAs) per ring stage — one "weights landed" poll and one stage-release commit per PER MMAs.
template <bool PAIR, int PER>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_forward_kernel(MlpParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t kSlot = PAIR ? kStageBytes / 2 : kStageBytes;  // one chunk (or this CTA's half of it)
    constexpr uint32_t kStageB = PER * kSlot;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    // the pair walks the 256-row groups together: CTA rank r of the pair at even block b takes group b + r, b + r + grid, ...
    const int g_first = PAIR ? (int) (blockIdx.x & ~1u) : (int) blockIdx.x;
    if (p.dyn) {  // bucket of a sub-module dispatch: size and position were computed by the previous kernels
        p.row_index += p.dyn[0];
        p.rows = p.dyn[1];
        p.n_groups = (int) ((p.rows + kTiles * kTileM - 1) / (kTiles * kTileM));
        if (g_first >= p.n_groups) return;  // uniform per CTA pair, before any barrier / TMEM allocation
    }
    const int kStages = p.n_stages;
    uint8_t *s_act = smem;                           // [kTiles][kActBytes]
    uint8_t *s_pe = s_act + kTiles * kActBytes;      // [kTiles][kPeBytes]
    uint8_t *s_dir = s_pe + kTiles * kPeBytes;       // [kTiles][kDirBytes], only with view directions
    uint8_t *s_stage = s_dir + (p.need_viewdir ? kTiles * kDirBytes : 0);  // [kStages][kStageB]
    float *s_bias = reinterpret_cast<float *>(s_stage + kStages * kStageB);  // [2][kMaxN] this / next layer's bias row
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_bias + 2 * kMaxN);
    uint64_t *bar_full = bars;                       // [kMaxStages] weights landed (PAIR: in both CTAs, leader's barrier)
    uint64_t *bar_empty = bars + kMaxStages;         // [kMaxStages] weights consumed
    uint64_t *bar_acc = bars + 2 * kMaxStages;       // [kTiles] accumulator of the current layer complete
    uint64_t *bar_act = bars + 2 * kMaxStages + kTiles; // [kTiles] A operand of the next layer written
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 2 * kTiles);
    // issuer-side schedule, per layer 2 + 3 * 3 words: n_segs | n_chunks << 8 | weight bytes per chunk << 16;
    // instruction descriptor; then per segment: low word of tile 0's A descriptor, high word,
    // n_chunks | (tile stride >> 4) << 16
    uint32_t *s_layer = s_tmem + 4;  // [kMaxLayers][kLayerWords]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const MlpSchedule &S = *p.sched;
    const int n_layers = S.n_layers;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            // PAIR: the leader's barrier also takes one arrival from the peer CTA (its half has landed)
            mbar_init(bar_full + s, (PAIR && leader) ? 2 : 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int t = 0; t < kTiles; ++t) {
            mbar_init(bar_acc + t, 1);
            // PAIR: one arrival per epilogue warp of both CTAs (the warp's lanes meet in __syncwarp first)
            mbar_init(bar_act + t, PAIR ? 2 * (kEpiThreads / 32) : kEpiThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    if (threadIdx.x < n_layers) {
        const LayerDesc ld = S.layers[threadIdx.x];
        uint32_t *w = s_layer + threadIdx.x * kLayerWords;
        w[0] = (uint32_t) ld.n_segs | ((uint32_t) (ld.chunk_end - ld.chunk_begin) << 8) | ((uint32_t) ld.n * kChunkK * 2u) << 16;
        w[1] = umma_idesc(PAIR ? 2 * kTileM : kTileM, ld.n);
        for (int i = 0; i < ld.n_segs; ++i) {
            const SegDesc sg = ld.segs[i];
            const uint8_t *abuf = sg.a_src == kSrcAct ? s_act : (sg.a_src == kSrcPE ? s_pe : s_dir);
            const uint32_t sbo = sg.a_src == kSrcAct ? kActSBO : (sg.a_src == kSrcPE ? kPeSBO : kDirSBO);
            const uint32_t stride = sg.a_src == kSrcAct ? kActBytes : (sg.a_src == kSrcPE ? kPeBytes : kDirBytes);
            w[2 + 3 * i] = (((smem_u32(abuf) + (sg.a_k0 >> 3) * 128) >> 4) & 0x3fffu) | ((128u >> 4) << 16);
            w[3 + 3 * i] = (sbo >> 4) | (1u << 14);
            w[4 + 3 * i] = (uint32_t) sg.n_chunks | ((stride >> 4) << 16);
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // barriers initialised and TMEM allocated in both CTAs
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 8) {
        // ===================== TMA producer =====================
        // every layer's chunks are streamed twice in a row (tile 0, then tile 1): the second pass hits L2.
        // PAIR: this CTA copies its half of every chunk (rows n of B in [rank * N/2, (rank + 1) * N/2)).
        if (lane == 0) {
            uint32_t s = 0, ph = 1;
            for (int gb = g_first; gb < p.n_groups; gb += gridDim.x) {
                const uint8_t *layer_src = p.weights;  // chunks are contiguous in schedule order
                for (int l = 0; l < n_layers; ++l) {
                    const uint32_t *w = s_layer + l * kLayerWords;
                    const uint32_t w0 = w[0];
                    const uint32_t n_segs = w0 & 0xffu, n_chunks = (w0 >> 8) & 0xffu, bytes = w0 >> 16;
                    const uint32_t part = PAIR ? bytes >> 1 : bytes;
                    for (int t = 0; t < kTiles; ++t) {
                        const uint8_t *src = layer_src + cta_rank * part;
                        for (uint32_t i = 0; i < n_segs; ++i) {
                            for (uint32_t left = w[4 + 3 * i] & 0xffffu; left > 0;) {
                                const uint32_t k = left < (uint32_t) PER ? left : (uint32_t) PER;
                                mbar_wait(bar_empty + s, ph);
#if MNV_MLP_DIAG & 2
                                mbar_arrive(bar_full + s);
#else
                                mbar_expect_tx(bar_full + s, k * part);
                                for (uint32_t j = 0; j < k; ++j)
                                    tma_bulk_g2s(s_stage + s * kStageB + j * kSlot, src + j * bytes, part, bar_full + s);
#endif
                                src += k * bytes;
                                left -= k;
                                if (++s == (uint32_t) kStages) {
                                    s = 0;
                                    ph ^= 1;
                                }
                            }
                        }
                    }
                    layer_src += n_chunks * bytes;
                }
            }
        }
    } else if (warp == 9 && PAIR && !leader) {
        // ===================== peer CTA: forward "my half has landed" to the leader's barrier =====================
        if (lane == 0) {
            const uint32_t full_leader = mapa_u32(smem_u32(bar_full), 0);
            uint32_t s = 0, ph = 0;
            for (int gb = g_first; gb < p.n_groups; gb += gridDim.x)
                for (int l = 0; l < n_layers; ++l) {
                    const uint32_t *w = s_layer + l * kLayerWords;
                    const uint32_t n_segs = w[0] & 0xffu;
                    for (int t = 0; t < kTiles; ++t)
                        for (uint32_t i = 0; i < n_segs; ++i)
                            for (uint32_t left = w[4 + 3 * i] & 0xffffu; left > 0; left -= (left < (uint32_t) PER ? left : (uint32_t) PER)) {
                                mbar_wait(bar_full + s, ph);
                                mbar_arrive_cluster(full_leader + 8 * s);
                                if (++s == (uint32_t) kStages) {
                                    s = 0;
                                    ph ^= 1;
                                }
                            }
                }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        // The whole warp runs the (warp-uniform) loop and one elected lane issues.  What the lane pays between two MMAs
        // is what bounds the kernel (a poll of the weights barrier ~70 clk and a commit cost more than the MMA itself:
        // tools/micro/umma_bench.cu, profiles/r2_mlp_smem_diag.md), so the ring position is carried as addresses,
        // one poll / one stage-release commit serve PER MMAs, and the poll of the NEXT stage is issued right behind the
        // commit so that its latency overlaps the loop control.
        const uint32_t b_lo0 = ((smem_u32(s_stage) >> 4) & 0x3fffu) | ((128u >> 4) << 16);
        const uint32_t b_hi = ((kChunkK / 8 * 128u) >> 4) | (1u << 14);
        const bool issue = elect_one();
        const uint32_t full0 = smem_u32(bar_full), empty0 = smem_u32(bar_empty);
        uint32_t s = 0, ph = 0, act_ph = 0;
        uint32_t full_a = full0, empty_a = empty0, b_lo = b_lo0;
        MLP_T(long long w_act = 0, w_full = 0, w_act_l[kMaxLayers] = {0}; const long long t_begin = p.dbg ? clock64() : 0;)
        bool ready = mbar_try<PAIR>(full_a, ph);
        for (int gb = g_first; gb < p.n_groups; gb += gridDim.x) {
            for (int l = 0; l < n_layers; ++l) {
                const uint32_t *w = s_layer + l * kLayerWords;
                const uint32_t n_segs = w[0] & 0xffu, idesc = w[1];
                for (int t = 0; t < kTiles; ++t) {
                    MLP_T(const long long t0 = p.dbg ? clock64() : 0;)
                    while (!mbar_try<PAIR>(smem_u32(bar_act + t), act_ph)) {}  // A operand of tile t ready, accumulator drained
                    MLP_T(if (p.dbg) { const long long d = clock64() - t0; w_act += d; w_act_l[l] += d; })
                    uint32_t acc = 0;
                    for (uint32_t i = 0; i < n_segs; ++i) {
                        const uint32_t w2 = w[4 + 3 * i];
                        uint32_t a_lo = w[2 + 3 * i] + t * (w2 >> 16);
                        const uint32_t a_hi = w[3 + 3 * i];
                        for (uint32_t left = w2 & 0xffffu; left > 0;) {
                            const uint32_t k = left < (uint32_t) PER ? left : (uint32_t) PER;
                            MLP_T(const long long t1 = p.dbg ? clock64() : 0;)
                            while (!ready) ready = mbar_try<PAIR>(full_a, ph);
                            MLP_T(if (p.dbg) w_full += clock64() - t1;)
                            tc_fence_after();
                            if (issue) {
#pragma unroll
                                for (uint32_t j = 0; j < (uint32_t) PER; ++j)
                                    if (j < k) {
                                        umma_bf16<PAIR>(tmem_base + t * kMaxN, ((uint64_t) a_hi << 32) | (a_lo + j * ((kChunkK / 8 * 128u) >> 4)),
                                                        ((uint64_t) b_hi << 32) | (b_lo + j * (kSlot >> 4)), idesc, j == 0 ? acc : 1u);
                                    }
                                tc_commit_addr<PAIR>(empty_a);  // frees the weight stage (in both CTAs)
                            }
                            acc = 1;
                            a_lo += k * ((kChunkK / 8 * 128u) >> 4);  // next K columns of the K-major A operand
                            left -= k;
                            full_a += 8;
                            empty_a += 8;
                            b_lo += kStageB >> 4;
                            if (++s == (uint32_t) kStages) {
                                s = 0;
                                ph ^= 1;
                                full_a = full0;
                                empty_a = empty0;
                                b_lo = b_lo0;
                            }
                            ready = mbar_try<PAIR>(full_a, ph);
                        }
                    }
                    if (issue) tc_commit_addr<PAIR>(smem_u32(bar_acc + t));  // accumulator complete -> epilogue of tile t
                    __syncwarp();
                }
                act_ph ^= 1;
            }
        }
#ifdef MNV_MLP_TIMING
        if (p.dbg && issue) {
            long long *o = p.dbg + (size_t) blockIdx.x * 32;
            o[0] = clock64() - t_begin;
            o[1] = w_act;
            o[2] = w_full;
            for (int l = 0; l < n_layers && l < 16; ++l) o[8 + l] = w_act_l[l];
        }
#endif
    } else {
        // ===================== PE + epilogue warps =====================
        // thread = (row of the tile, column half); both tiles in turn
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;  // 0..127, TMEM lane
        const uint32_t tlane = tmem_base + ((uint32_t) (q * 32) << 16);
        uint32_t acc_ph = 0;
        MLP_T(long long e_wait = 0, e_busy = 0, e_busy_l2 = 0, e_bar = 0; const long long e_begin = p.dbg ? clock64() : 0;)
        const uint32_t act_leader = PAIR ? mapa_u32(smem_u32(bar_act), 0) : 0u;
        // "A operand of tile t written / accumulator drained": every thread has fenced its own writes; PAIR: the warp's
        // lanes meet and one of them arrives on the LEADER's barrier (a remote arrival for the peer CTA)
        auto act_arrive = [&](int t) {
            if constexpr (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(act_leader + 8 * t);
            } else {
                mbar_arrive(bar_act + t);
            }
        };
        for (int gb = g_first; gb < p.n_groups; gb += gridDim.x) {
            const int grp = gb + (int) cta_rank;  // may lie past the last group in the peer CTA: its rows are invalid
            int64_t grow[kTiles];
            bool valid[kTiles];
            int ai[kTiles];
            // ---- inputs: positional encodings -> A buffers (octaves split between the two halves) ----
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
                const int64_t lrow = ((int64_t) grp * kTiles + t) * kTileM + row;
                valid[t] = lrow < p.rows;
                grow[t] = (valid[t] && p.row_index) ? (int64_t) p.row_index[lrow] : lrow;
                float xin[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) xin[c] = (valid[t] && c < p.in_dim) ? p.x[grow[t] * p.in_dim + c] : 0.f;
                uint8_t *pe = s_pe + t * kPeBytes;
                const int fh = p.pe_xyz_freqs >> 1;
                if (h == 0) {
                    for (int c = 0; c < 3; ++c) put_bf16(pe, row, c, kPeSBO, xin[c]);
                    write_pe_octaves(pe, row, kPeSBO, xin, 0, fh);
                } else {
                    write_pe_octaves(pe, row, kPeSBO, xin, fh, p.pe_xyz_freqs);
                    for (int k = 3 + 6 * p.pe_xyz_freqs; k < kPeK; ++k)  // padding, with the two bias columns = 1
                        put_bf16(pe, row, k, kPeSBO, (k == p.ones_col || k == p.ones_col + 1) ? 1.f : 0.f);
                    if (p.need_viewdir) {
                        uint8_t *dir = s_dir + t * kDirBytes;
                        for (int c = 0; c < 3; ++c) put_bf16(dir, row, c, kDirSBO, xin[3 + c]);
                        write_pe_octaves(dir, row, kDirSBO, xin + 3, 0, p.pe_dir_freqs);
                        for (int k = 3 + 6 * p.pe_dir_freqs; k < kDirK; ++k) put_bf16(dir, row, k, kDirSBO, 0.f);
                    }
                }
                ai[t] = 0;
                if (p.app_bias) {
                    const int a = p.app_col >= 0 ? (int) xin[p.app_col] : 0;
                    ai[t] = min(max(a, 0), p.n_appearance - 1);
                }
                fence_async_smem();
                tc_fence_before();
                act_arrive(t);
            }

            float sigma[kTiles] = {0.f, 0.f};
            // bias rows travel through shared memory one layer ahead (thread i fetches column i), so the
            // epilogue's adds never wait on global memory: s_bias[l & 1] is layer l's row
            {
                asm volatile("bar.sync 1, 256;" ::: "memory");  // the previous group's last epilogue is done with the rows
                const uint32_t bo = S.layers[0].bias_off;
                s_bias[threadIdx.x] = bo != kNoBias ? __ldg(p.biases + bo + threadIdx.x) : 0.f;
            }
            for (int l = 0; l < n_layers; ++l) {
                const LayerDesc ld = S.layers[l];
                const bool last_layer = ld.epilogue == kEpiOut;
                // all 256 threads are done with layer l-1 (and its bias row); layer l's row is complete
                MLP_T(const long long b0 = p.dbg ? clock64() : 0;)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                MLP_T(if (p.dbg) e_bar += clock64() - b0;)
                if (l + 1 < n_layers) {
                    const LayerDesc nx = S.layers[l + 1];
                    s_bias[((l + 1) & 1) * kMaxN + threadIdx.x] =
                            (nx.bias_off != kNoBias && threadIdx.x < nx.n) ? __ldg(p.biases + nx.bias_off + threadIdx.x) : 0.f;
                }
                const float *lbias = s_bias + (l & 1) * kMaxN;
                // columns of this thread: half of the layer (the 32-wide output layer: all, by half 0)
                const int n_mine = ld.n >= 64 ? (ld.n >> 1) : (h == 0 ? ld.n : 0);
                const int col0 = ld.n >= 64 ? h * n_mine : 0;
#pragma unroll
                for (int t = 0; t < kTiles; ++t) {
                    uint8_t *act = s_act + t * kActBytes;
                    MLP_T(const long long e0 = p.dbg ? clock64() : 0;)
                    mbar_wait(bar_acc + t, acc_ph);
                    MLP_T(const long long e1 = p.dbg ? clock64() : 0; e_wait += e1 - e0;)
                    tc_fence_after();
                    const uint32_t taddr = tlane + t * kMaxN + col0;
                    // the common shapes (256- and 128-wide layers that write the next A operand) run straight-line code
                    bool lean = !last_layer && (ld.n == 256 || ld.n == 128);
                    if (lean) {
                        uint8_t *dst = act + a_off(row, col0, kActSBO);
                        const float *sw = p.sigma_w + col0;
                        switch (ld.epilogue | (ld.n == 256 ? 0x100u : 0u)) {
                            case kEpiReluAct | 0x100u: epilogue_unit<kEpiReluAct, 128>(taddr, dst, lbias + col0, sw, sigma[t]); break;
                            case kEpiReluActSigma | 0x100u: epilogue_unit<kEpiReluActSigma, 128>(taddr, dst, lbias + col0, sw, sigma[t]); break;
                            case kEpiLinearAct | 0x100u: epilogue_unit<kEpiLinearAct, 128>(taddr, dst, lbias + col0, sw, sigma[t]); break;
                            case kEpiReluAct: epilogue_unit<kEpiReluAct, 64>(taddr, dst, lbias + col0, sw, sigma[t]); break;
                            case kEpiReluActApp:
                                epilogue_unit<kEpiReluActApp, 64>(taddr, dst, p.app_bias + (size_t) ai[t] * p.head_n + col0, sw, sigma[t]);
                                break;
                            default: lean = false; break;
                        }
                    }
                    uint32_t r[2][32];
                    if (!lean && n_mine > 0) tmem_ld32(taddr, r[0]);
#pragma unroll 1
                    for (int c0 = 0; c0 < (lean ? 0 : n_mine); c0 += 64) {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int cc = c0 + half * 32;
                            if (cc >= n_mine) break;
                            tmem_wait();
                            if (cc + 32 < n_mine) tmem_ld32(taddr + cc + 32, r[half ^ 1]);  // overlaps the math below
                            const uint32_t(&rr)[32] = r[half];
                            const int col = col0 + cc;
                            if (last_layer) {
                                if (valid[t]) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (col + j < p.out_real)
                                            p.out[grow[t] * p.out_stride + col + j] = __uint_as_float(rr[j]) + lbias[col + j];
                                }
                                continue;
                            }
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
                            if (ld.epilogue == kEpiReluActApp) {
                                const float4 *b4 = reinterpret_cast<const float4 *>(p.app_bias + (size_t) ai[t] * p.head_n + col);
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    const float4 b = __ldg(b4 + g);
                                    v[4 * g + 0] += b.x;
                                    v[4 * g + 1] += b.y;
                                    v[4 * g + 2] += b.z;
                                    v[4 * g + 3] += b.w;
                                }
                            } else if (ld.bias_off != kNoBias) {
                                const float4 *b4 = reinterpret_cast<const float4 *>(lbias + col);
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    const float4 b = b4[g];
                                    v[4 * g + 0] += b.x;
                                    v[4 * g + 1] += b.y;
                                    v[4 * g + 2] += b.z;
                                    v[4 * g + 3] += b.w;
                                }
                            }
                            if (ld.epilogue == kEpiReluActSigma) {
                                const float4 *w4 = reinterpret_cast<const float4 *>(p.sigma_w + col);
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    const float4 w = __ldg(w4 + g);
                                    sigma[t] = fmaf(fmaxf(v[4 * g + 0], 0.f), w.x, sigma[t]);
                                    sigma[t] = fmaf(fmaxf(v[4 * g + 1], 0.f), w.y, sigma[t]);
                                    sigma[t] = fmaf(fmaxf(v[4 * g + 2], 0.f), w.z, sigma[t]);
                                    sigma[t] = fmaf(fmaxf(v[4 * g + 3], 0.f), w.w, sigma[t]);
                                }
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 o;
                                if (ld.epilogue == kEpiLinearAct) {
                                    o.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]);
                                    o.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
                                    o.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]);
                                    o.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
                                } else {
                                    o.x = pack_bf16_relu(v[8 * g + 0], v[8 * g + 1]);
                                    o.y = pack_bf16_relu(v[8 * g + 2], v[8 * g + 3]);
                                    o.z = pack_bf16_relu(v[8 * g + 4], v[8 * g + 5]);
                                    o.w = pack_bf16_relu(v[8 * g + 6], v[8 * g + 7]);
                                }
#if MNV_MLP_DIAG & 1
                                if (p.rows < 0)  // never true: keeps the arithmetic, drops the store traffic
#endif
                                *reinterpret_cast<uint4 *>(act + a_off(row, col + 8 * g, kActSBO)) = o;
                            }
                        }
                    }
                    if (ld.epilogue == kEpiReluActSigma) {
                        // the sigma head needs the whole row: the two column halves meet through a
                        // shuffle-free exchange in the (now dead) TMEM-drained registers is not possible
                        // across warps, so half 1 parks its partial sum in global memory (its own row's
                        // sigma slot) and half 0 adds it at output time; the act barrier below orders them.
                        if (h == 1 && valid[t]) p.out[grow[t] * p.out_stride + p.out_real] = sigma[t];
                    }
                    if (last_layer) {
                        if (h == 0 && valid[t]) {
                            float *po = p.out + grow[t] * p.out_stride + p.out_real;
                            float sg = sigma[t] + __ldcg(po) + __ldg(p.sigma_w + 256);
                            sg = p.sigma_activation == 1 ? (sg > 20.f ? sg : log1pf(expf(sg))) : fmaxf(sg, 0.f);
                            *po = sg;
                        }
                        tc_fence_before();  // TMEM reads ordered before the next group's first MMA
                    } else {
                        if (ld.epilogue == kEpiReluActSigma) __threadfence_block();
                        fence_async_smem();  // generic-proxy smem writes -> visible to the UMMA proxy
                        tc_fence_before();
                        act_arrive(t);
                    }
                    MLP_T(if (p.dbg) { e_busy += clock64() - e1; if (l == 2) e_busy_l2 += clock64() - e1; })
                }
                acc_ph ^= 1;
            }
        }
#ifdef MNV_MLP_TIMING
        if (p.dbg && threadIdx.x == 0) {  // epilogue warp 0: waiting for accumulators / working / waiting for the other warps
            long long *o = p.dbg + (size_t) blockIdx.x * 32;
            o[3] = clock64() - e_begin;
            o[4] = e_wait;
            o[5] = e_busy;
            o[6] = e_bar;
            o[7] = e_busy_l2;
        }
#endif
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // PAIR: neither CTA leaves while the other may still signal it
    if (warp == 9) {
        tc_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

constexpr size_t kSmemLimit = 232448;  // opt-in dynamic shared memory of one sm_100 CTA
constexpr size_t mlp_smem_fixed(bool viewdir) {
    return kTiles * (kActBytes + kPeBytes + (viewdir ? kDirBytes : 0)) + 2 * kMaxN * 4 + (2 * kMaxStages + 2 * kTiles) * 8 + 16 +
           kMaxLayers * kLayerWords * 4;
}
// ring depth: everything that is left (7 x 8 KiB without view directions, 5 with; 13 / 10 stages of 4 KiB)
constexpr int mlp_stages(bool viewdir, size_t stage_bytes) {
    return (int) std::min<size_t>(kMaxStages, (kSmemLimit - mlp_smem_fixed(viewdir)) / stage_bytes);
}
constexpr size_t mlp_smem_bytes(bool viewdir, size_t stage_bytes) {
    return mlp_smem_fixed(viewdir) + mlp_stages(viewdir, stage_bytes) * stage_bytes;
}
static_assert(mlp_stages(false, kStageBytes) == 7 && mlp_stages(true, kStageBytes) == 5 && mlp_stages(true, 2 * kStageBytes) >= 2,
              "shared memory budget of one sm_100 CTA");

// launch shape (dev overrides): MNV_MLP_PAIR=0/1 (CTA pairs with cta_group::2 MMAs), MNV_MLP_PER=1/2/4 (MMAs per ring stage;
// 0 = by model: 4 — three 16 KiB stages per CTA — without view directions, 2 — five 8 KiB stages — with them).
// Measured at 262 144 rows (profiles/r2_mlp_modes.log): pair/4 0.346 ms, pair/1 0.376, pair/2 0.404, single/1 0.382, single/2 0.408.
struct MlpMode {
    bool pair;
    int per;
};
const MlpMode &mlp_mode() {
    static const MlpMode m = [] {
        MlpMode r{true, 0};
        if (const char *e = std::getenv("MNV_MLP_PAIR")) r.pair = std::atoi(e) != 0;
        if (const char *e = std::getenv("MNV_MLP_PER")) r.per = std::atoi(e);
        if (r.per != 1 && r.per != 2 && r.per != 4) r.per = 0;
        return r;
    }();
    return m;
}
int mlp_per(const MlpMode &mode, bool viewdir) { return mode.per ? mode.per : (viewdir ? 2 : 4); }
using MlpKernel = void (*)(MlpParams);
MlpKernel mlp_kernel(bool pair, int per) {
    if (pair) return per == 1 ? mlp_forward_kernel<true, 1> : (per == 2 ? mlp_forward_kernel<true, 2> : mlp_forward_kernel<true, 4>);
    return per == 1 ? mlp_forward_kernel<false, 1> : (per == 2 ? mlp_forward_kernel<false, 2> : mlp_forward_kernel<false, 4>);
}

// ------------------------------------------------------------------ host packing
uint16_t f2bf(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t) ((u >> 16) | 0x40u);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
    return (uint16_t) (u >> 16);
}
float bf2f(uint16_t h) {
    const uint32_t u = (uint32_t) h << 16;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

struct KSeg {  // a run of K columns of one layer's A operand
    int a_src, a_k0, k_len;       // in the A buffer (k_len multiple of 16)
    int w_col0, w_cols;           // real weight columns covered: [w_col0, w_col0 + w_cols)
};

}  // namespace

struct MlpModel {
    int device = 0;
    uint8_t *weights = nullptr;
    float *sigma_w = nullptr;   // [256] + bias
    float *biases = nullptr;    // fp32 bias rows for layers whose bias is not folded into the GEMM
    float *app_bias = nullptr;  // [n_appearance][head_n]
    MlpSchedule *sched_dev = nullptr;
    MlpSchedule sched;
    mnv_mlp_desc cfg;
    int head_n = 0, ones_col = 0;
    int num_sms = 0;
    int in_dim = 3;
    int out_dim = 0;
    double flops_per_row = 0;
};

MlpModel *mlp_create(const mnv_mlp_desc &d, int device, int *rc_out) {
    auto fail = [&](int rc, const char *msg) -> MlpModel * {
        set_error("%s", msg);
        *rc_out = rc;
        return nullptr;
    };
    if (d.width != 256 || d.n_trunk_layers < 2 || d.n_trunk_layers > 12 || d.head_width > 256 ||
        d.head_width % 32 || d.out_rgb_dim < 1 || d.out_rgb_dim > 32)
        return fail(MNV_ERR_INVALID, "unsupported MLP shape (width must be 256, head a multiple of 32 <= 256, "
                                     "at most 32 colour outputs)");
    const int pe = 3 + 6 * d.pe_xyz_freqs;
    const int pe_dir = d.need_viewdir ? 3 + 6 * d.pe_dir_freqs : 0;
    // two columns of ones behind the encoding carry the biases through the GEMM (same 16-column chunk)
    const int ones_col = (pe % 16 == 15) ? pe + 1 : pe;
    if (d.pe_xyz_freqs < 2 || ones_col + 2 > kPeK || pe_dir > kDirK)
        return fail(MNV_ERR_INVALID, "positional encoding too wide");
    if (d.appearance_dim > 0 && (d.n_appearance < 1 || !d.embedding))
        return fail(MNV_ERR_INVALID, "appearance_dim > 0 needs an embedding table");
    if (d.skip_layer >= d.n_trunk_layers || d.skip_layer < 1)
        return fail(MNV_ERR_INVALID, "skip_layer outside the trunk");
    const int pe_len = (ones_col + 2 + 15) / 16 * 16;  // K columns of the PE segment incl. the ones
    if (2 * (pe_len / kChunkK) + (d.n_trunk_layers + 1) * (256 / kChunkK + 1) + kDirK / kChunkK + 1 +
                d.head_width / kChunkK + 1 > kMaxChunks)
        return fail(MNV_ERR_INVALID, "MLP too deep for the on-chip schedule");

    auto *m = new MlpModel();
    m->device = device;
    m->cfg = d;
    m->ones_col = ones_col;
    m->in_dim = 3 + (d.need_viewdir ? 3 : 0) + (d.appearance_dim > 0 ? 1 : 0);
    m->out_dim = d.out_rgb_dim + 1;
    MlpSchedule &S = m->sched;
    std::memset(&S, 0, sizeof(S));
    std::vector<uint8_t> blob;
    std::vector<float> bias_rows(32, 0.f);
    double macs = 0;
    // bit l set -> layer l adds its fp32 bias in the epilogue; clear -> the bias rides in the GEMM as
    // bf16 hi + lo against two columns of ones.  Measured on B200: the GEMM route is 10 % faster but
    // moves 2.9 % of the rows beyond 1e-3 of the bf16-emulated reference (0.4 % with epilogue biases),
    // so the epilogue route is the default; MNV_MLP_BIAS_EPILOGUE=<mask> is a dev override.
    const unsigned epi_bias_mask = std::getenv("MNV_MLP_BIAS_EPILOGUE") ? (unsigned) std::strtoul(std::getenv("MNV_MLP_BIAS_EPILOGUE"), nullptr, 0) : ~0u;

    // W [n_real][k_real_total] row-major (nn.Linear); b may be null (bias applied by the epilogue)
    auto add_layer = [&](const float *W, const float *b, int n_real, int k_real_total, std::vector<KSeg> segs,
                         int epilogue) {
        LayerDesc &L = S.layers[S.n_layers];
        L.n = (uint16_t) ((n_real + 31) / 32 * 32);  // the epilogue reads 32 TMEM columns at a time
        L.n_real = (uint16_t) n_real;
        L.epilogue = (uint32_t) epilogue;
        L.chunk_begin = (uint16_t) S.n_chunks;
        L.bias_off = kNoBias;
        const int n_mma = L.n;
        if (b && ((epi_bias_mask >> S.n_layers) & 1u)) {
            L.bias_off = (uint32_t) bias_rows.size();
            for (int c = 0; c < n_mma; ++c) bias_rows.push_back(c < n_real ? b[c] : 0.f);
            b = nullptr;
        }
        const int ones_chunk0 = ones_col / 16 * 16;
        bool has_ones = false;
        for (const KSeg &sg : segs) has_ones |= sg.a_src == kSrcPE && sg.a_k0 <= ones_chunk0 && sg.a_k0 + sg.k_len >= ones_chunk0 + 16;
        if (b && !has_ones) segs.push_back({kSrcPE, ones_chunk0, 16, 0, 0});  // bias-only chunk
        L.n_segs = (uint16_t) segs.size();
        for (size_t si = 0; si < segs.size(); ++si)
            L.segs[si] = SegDesc{(uint16_t) segs[si].a_src, (uint16_t) segs[si].a_k0, (uint16_t) (segs[si].k_len / kChunkK), 0};
        for (const KSeg &sg : segs) {
            for (int k0 = 0; k0 < sg.k_len; k0 += kChunkK) {
                ChunkDesc &C = S.chunks[S.n_chunks++];
                C.gmem_off = (uint32_t) blob.size();
                C.n = (uint16_t) n_mma;
                C.a_src = (uint16_t) sg.a_src;
                C.a_k0 = (uint16_t) (sg.a_k0 + k0);
                const size_t bytes = (size_t) n_mma * kChunkK * 2;
                blob.resize(blob.size() + bytes, 0);
                uint8_t *dst = blob.data() + C.gmem_off;
                const int sbo = (kChunkK / 8) * 128;
                auto put = [&](int n, int k, uint16_t hbits) {
                    std::memcpy(dst + (n / 8) * sbo + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2, &hbits, 2);
                };
                for (int n = 0; n < n_real; ++n) {
                    for (int k = 0; k < kChunkK; ++k) {
                        const int kk = k0 + k;  // column inside the segment
                        if (kk < sg.w_cols) put(n, k, f2bf(W[(size_t) n * k_real_total + sg.w_col0 + kk]));
                    }
                    if (b && sg.a_src == kSrcPE && C.a_k0 == ones_chunk0) {  // bias = hi + lo against the ones
                        const uint16_t hi = f2bf(b[n]);
                        put(n, ones_col - ones_chunk0, hi);
                        put(n, ones_col + 1 - ones_chunk0, f2bf(b[n] - bf2f(hi)));
                    }
                }
            }
        }
        L.chunk_end = (uint16_t) S.n_chunks;
        macs += (double) n_real * k_real_total;
        ++S.n_layers;
    };

    // trunk
    for (int l = 0; l < d.n_trunk_layers; ++l) {
        std::vector<KSeg> segs;
        int k_total;
        if (l == 0) {
            segs.push_back({kSrcPE, 0, pe_len, 0, pe});
            k_total = pe;
        } else if (l == d.skip_layer) {  // cat([pe, h]) like the NeRF skip connection
            segs.push_back({kSrcAct, 0, 256, pe, 256});
            segs.push_back({kSrcPE, 0, pe_len, 0, pe});
            k_total = pe + 256;
        } else {
            segs.push_back({kSrcAct, 0, 256, 0, 256});
            k_total = 256;
        }
        add_layer(d.trunk_w[l], d.trunk_b[l], 256, k_total, segs,
                  l == d.n_trunk_layers - 1 ? kEpiReluActSigma : kEpiReluAct);
    }
    // feature layer (no activation)
    add_layer(d.final_w, d.final_b, 256, 256, {{kSrcAct, 0, 256, 0, 256}}, kEpiLinearAct);
    // head 1: [feature(256) | dir PE | appearance].  The appearance columns do not enter the GEMM:
    // their product with the (per-row constant) embedding row is a bias, tabulated per index below.
    std::vector<float> app_bias;
    int head_n = 0;
    {
        std::vector<KSeg> segs;
        segs.push_back({kSrcAct, 0, 256, 0, 256});
        int col = 256;
        if (pe_dir > 0) {
            segs.push_back({kSrcDir, 0, kDirK, col, pe_dir});
            col += pe_dir;
        }
        const int k_total = col + d.appearance_dim;
        const bool app = d.appearance_dim > 0;
        add_layer(d.head1_w, app ? nullptr : d.head1_b, d.head_width, k_total, segs, app ? kEpiReluActApp : kEpiReluAct);
        head_n = S.layers[S.n_layers - 1].n;
        if (app) {
            app_bias.assign((size_t) d.n_appearance * head_n, 0.f);
            for (int a = 0; a < d.n_appearance; ++a)
                for (int n = 0; n < d.head_width; ++n) {
                    double acc = 0;  // operand rounding of the tensor-core path, wide accumulation
                    for (int k = 0; k < d.appearance_dim; ++k)
                        acc += (double) bf2f(f2bf(d.head1_w[(size_t) n * k_total + col + k])) *
                               (double) bf2f(f2bf(d.embedding[(size_t) a * d.appearance_dim + k]));
                    app_bias[(size_t) a * head_n + n] = (float) (acc + (double) d.head1_b[n]);
                }
        }
    }
    // head 2
    add_layer(d.head2_w, d.head2_b, d.out_rgb_dim, d.head_width,
              {{kSrcAct, 0, d.head_width, 0, d.head_width}}, kEpiOut);
    macs += 256;  // sigma head
    m->flops_per_row = 2.0 * macs;
    std::vector<float> sig(d.sigma_w, d.sigma_w + 256);
    sig.push_back(d.sigma_b[0]);

    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->weights, blob.size());
    if (e == cudaSuccess) e = cudaMemcpy(m->weights, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->biases, bias_rows.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->biases, bias_rows.data(), bias_rows.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->sigma_w, sig.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->sigma_w, sig.data(), sig.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !app_bias.empty()) {
        e = cudaMalloc(&m->app_bias, app_bias.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(m->app_bias, app_bias.data(), app_bias.size() * 4, cudaMemcpyHostToDevice);
    }
    m->head_n = head_n;
    if (e == cudaSuccess) e = cudaMalloc(&m->sched_dev, sizeof(MlpSchedule));
    if (e == cudaSuccess) e = cudaMemcpy(m->sched_dev, &S, sizeof(S), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mlp_kernel(mlp_mode().pair, mlp_per(mlp_mode(), d.need_viewdir != 0)),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemLimit);
    if (e != cudaSuccess) {
        *rc_out = cuda_fail(e, "mlp_create", __FILE__, __LINE__);
        mlp_destroy(m);
        return nullptr;
    }
    *rc_out = MNV_OK;
    return m;
}

void mlp_destroy(MlpModel *m) {
    if (!m) return;
    cudaFree(m->weights);
    cudaFree(m->sigma_w);
    cudaFree(m->biases);
    cudaFree(m->app_bias);
    cudaFree(m->sched_dev);
    delete m;
}

static int mlp_launch(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, const int32_t *dyn_dev,
                      int64_t rows, int in_dim, float *out_dev, int out_stride, cudaStream_t stream);

int mlp_forward(const MlpModel *m, const float *x_dev, int64_t rows, int in_dim, float *out_dev,
                int out_stride, cudaStream_t stream) {
    return mlp_forward_indexed(m, x_dev, nullptr, rows, in_dim, out_dev, out_stride, stream);
}

int mlp_forward_indexed(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, int64_t rows,
                        int in_dim, float *out_dev, int out_stride, cudaStream_t stream) {
    return mlp_launch(m, x_dev, row_index_dev, nullptr, rows, in_dim, out_dev, out_stride, stream);
}

int mlp_forward_bucket(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, const int32_t *dyn_dev,
                       int64_t max_rows, int in_dim, float *out_dev, int out_stride, cudaStream_t stream) {
    return mlp_launch(m, x_dev, row_index_dev, dyn_dev, max_rows, in_dim, out_dev, out_stride, stream);
}

static int mlp_launch(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, const int32_t *dyn_dev,
                      int64_t rows, int in_dim, float *out_dev, int out_stride, cudaStream_t stream) {
    if (rows <= 0) return MNV_OK;
    if (in_dim != m->in_dim) {
        set_error("mlp_forward: in_dim %d, model expects %d", in_dim, m->in_dim);
        return MNV_ERR_INVALID;
    }
    if (out_stride < m->out_dim) {
        set_error("mlp_forward: out_stride %d < output width %d", out_stride, m->out_dim);
        return MNV_ERR_INVALID;
    }
    MlpParams p;
    p.weights = m->weights;
    p.sigma_w = m->sigma_w;
    p.biases = m->biases;
    p.app_bias = m->app_bias;
    p.sched = m->sched_dev;
    p.x = x_dev;
    p.out = out_dev;
    p.row_index = row_index_dev;
    p.dyn = dyn_dev;
    p.rows = rows;
    p.in_dim = in_dim;
    p.out_stride = out_stride;
    p.n_groups = (int) ((rows + kTiles * kTileM - 1) / (kTiles * kTileM));
    p.pe_xyz_freqs = m->cfg.pe_xyz_freqs;
    p.pe_dir_freqs = m->cfg.pe_dir_freqs;
    p.need_viewdir = m->cfg.need_viewdir;
    p.n_appearance = m->cfg.n_appearance;
    p.head_n = m->head_n;
    p.app_col = m->cfg.appearance_dim > 0 ? in_dim - 1 : -1;
    p.ones_col = m->ones_col;
    p.sigma_activation = m->cfg.sigma_activation;
    p.out_real = m->cfg.out_rgb_dim;
    const MlpMode &mode = mlp_mode();
    // pairs: an even number of CTAs, one pair per TPC
    const int grid = mode.pair ? std::min((p.n_groups + 1) & ~1, m->num_sms & ~1) : std::min(p.n_groups, m->num_sms);
    p.dbg = nullptr;
#ifdef MNV_MLP_TIMING
    static const bool debug = std::getenv("MNV_MLP_DEBUG") != nullptr;  // dev: where does the issuer wait?
#else
    const bool debug = false;
#endif
    if (debug) {
        MNV_CUDA(cudaMalloc(&p.dbg, (size_t) grid * 32 * sizeof(long long)));
        MNV_CUDA(cudaMemsetAsync(p.dbg, 0, (size_t) grid * 32 * sizeof(long long), stream));
    }
    const int per = mlp_per(mode, p.need_viewdir != 0);
    const size_t stage_bytes = (size_t) per * (mode.pair ? kStageBytes / 2 : kStageBytes);
    p.n_stages = mlp_stages(p.need_viewdir != 0, stage_bytes);
    if (p.n_stages < 2) {
        set_error("mlp_forward: MNV_MLP_PER=%d leaves %d ring stages", per, p.n_stages);
        return MNV_ERR_INVALID;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3(kMlpThreads);
    cfg.dynamicSmemBytes = mlp_smem_bytes(p.need_viewdir != 0, stage_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = mode.pair ? 1 : 0;
    MNV_CUDA(cudaLaunchKernelEx(&cfg, mlp_kernel(mode.pair, per), p));
    MNV_CUDA(cudaGetLastError());
    if (debug) {
        std::vector<long long> h((size_t) grid * 32);
        MNV_CUDA(cudaMemcpyAsync(h.data(), p.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        MNV_CUDA(cudaStreamSynchronize(stream));
        cudaFree(p.dbg);
        std::fprintf(stderr, "[mlp dbg] CTA0 issuer: total %lld clk = wait_act %lld + wait_full %lld + issue %lld; wait_act per layer:",
                     h[0], h[1], h[2], h[0] - h[1] - h[2]);
        for (int l = 0; l < m->sched.n_layers; ++l) std::fprintf(stderr, " %lld", h[8 + l]);
        std::fprintf(stderr, "\n[mlp dbg] CTA0 epilogue warp 0: total %lld clk = wait_acc %lld + busy %lld (layer 2: %lld) + layer barrier %lld + rest\n",
                     h[3], h[4], h[5], h[7], h[6]);
    }
    return MNV_OK;
}

double mlp_flops_per_row(const MlpModel *m) { return m->flops_per_row; }
int mlp_in_dim(const MlpModel *m) { return m->in_dim; }
int mlp_out_dim(const MlpModel *m) { return m->out_dim; }

}  // namespace mnv
