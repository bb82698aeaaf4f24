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
// One CTA owns a 128-row tile and keeps its activations on chip across all 11
// GEMMs:  A (activations, bf16) lives in shared memory in the UMMA K-major
// no-swizzle core-matrix layout, B (weights, bf16, pre-packed on the host into the
// same layout) is streamed from L2 by 1-D bulk TMA copies (cp.async.bulk ->
// UBLKCP) through a 3-stage mbarrier ring, D accumulates in TMEM (128 lanes x
// N<=256 fp32 columns) via tcgen05.mma issued by one thread, and the epilogue
// warps read D back with tcgen05.ld, add bias, apply ReLU, round to bf16 and write
// the next layer's A operand in place.
//   warps 0-3 : positional encoding, epilogues (warp w owns TMEM lanes 32w..32w+31)
//   warp  4   : TMA producer (one elected lane)
//   warp  5   : TMEM allocation + MMA issue (one elected lane)
#include <cuda_bf16.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "mnv_internal.cuh"

namespace mnv {
namespace {

constexpr int kMlpThreads = 192;
constexpr int kTileM = 128;
constexpr int kStages = 3;
constexpr int kChunkK = 64;
constexpr int kMaxN = 256;
constexpr int kStageBytes = kMaxN * kChunkK * 2;  // 32 KiB
constexpr int kActK = 256, kSideK = 80;
constexpr int kActBytes = kTileM * kActK * 2;    // 64 KiB
constexpr int kSideBytes = kTileM * kSideK * 2;  // 20 KiB
constexpr int kActSBO = (kActK / 8) * 128, kSideSBO = (kSideK / 8) * 128;
constexpr int kTmemCols = 256;
constexpr int kMaxLayers = 16, kMaxChunks = 64;

enum { kEpiReluAct = 0, kEpiReluActSigma = 1, kEpiLinearAct = 2, kEpiOut = 3 };
enum { kSrcAct = 0, kSrcPE = 1, kSrcAux = 2 };

struct ChunkDesc {
    uint32_t gmem_off;  // byte offset of the packed chunk in the weight blob
    uint32_t bytes;
    uint16_t n;     // padded N of the layer
    uint16_t kc;    // K of this chunk (multiple of 16, <= kChunkK)
    uint16_t a_src; // which A buffer
    uint16_t a_k0;  // first K column inside that buffer
    uint8_t first, last, layer, pad;
};

struct LayerDesc {
    uint16_t n;       // padded N
    uint16_t n_real;  // real output width
    uint32_t bias_off;  // float offset into the bias blob
    uint32_t epilogue;
};

struct MlpSchedule {
    int n_layers, n_chunks;
    LayerDesc layers[kMaxLayers];
    ChunkDesc chunks[kMaxChunks];
};

struct MlpParams {
    const uint8_t *__restrict__ weights;   // packed bf16 chunks
    const float *__restrict__ biases;      // fp32 biases, then sigma head weights + bias
    const float *__restrict__ embedding;   // [n_appearance][app_dim] fp32 (may be null)
    const MlpSchedule *__restrict__ sched;
    const float *__restrict__ x;           // [rows][in_dim]
    float *__restrict__ out;               // [rows][out_stride]
    const int32_t *__restrict__ row_index; // optional: logical row i reads x / writes out at row_index[i]
    int64_t rows;
    int in_dim, out_stride;
    int n_tiles;
    int pe_xyz_freqs, pe_dir_freqs;
    int need_viewdir, app_dim, n_appearance, app_col;  // app_col: column of x holding the index (-1: none)
    int sigma_w_off, sigma_b_off;  // float offsets into biases
    int sigma_activation;          // 0 = ReLU, 1 = softplus
    int out_real;                  // 3 * basis
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
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(bar))
                 : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) |
           ((uint32_t) (m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

// Byte offset of element (row, k) in an A buffer with K-major core-matrix layout.
__device__ __forceinline__ uint32_t a_off(int row, int k, int sbo) {
    return (uint32_t) ((row >> 3) * sbo + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

// Positional encoding of v[3] with `freqs` octaves into `feat` (3 + 6*freqs values):
// [v, sin(2^0 v), cos(2^0 v), sin(2^1 v), cos(2^1 v), ...] (NeRF "Embedding", include-input).
__device__ __forceinline__ void write_pe(uint8_t *buf, int row, int k0, int sbo, const float *v,
                                         int freqs, int k_pad_end) {
    auto put = [&](int k, float val) {
        *reinterpret_cast<__nv_bfloat16 *>(buf + a_off(row, k, sbo)) = __float2bfloat16_rn(val);
    };
    int k = k0;
    for (int c = 0; c < 3; ++c) put(k++, v[c]);
    float f = 1.f;
    for (int o = 0; o < freqs; ++o) {
        float s[3], cs[3];
        for (int c = 0; c < 3; ++c) sincosf(f * v[c], &s[c], &cs[c]);
        for (int c = 0; c < 3; ++c) put(k++, s[c]);
        for (int c = 0; c < 3; ++c) put(k++, cs[c]);
        f *= 2.f;
    }
    for (; k < k_pad_end; ++k) put(k, 0.f);
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_forward_kernel(const MlpParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_act = smem;
    uint8_t *s_pe = s_act + kActBytes;
    uint8_t *s_aux = s_pe + kSideBytes;
    uint8_t *s_stage = s_aux + kSideBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_stage + kStages * kStageBytes);
    uint64_t *bar_full = bars;                 // [kStages] weights landed
    uint64_t *bar_empty = bars + kStages;      // [kStages] weights consumed
    uint64_t *bar_acc = bars + 2 * kStages;    // accumulator of the current layer complete
    uint64_t *bar_act = bars + 2 * kStages + 1;  // A operand of the next layer written (128 arrivals)
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const MlpSchedule &S = *p.sched;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        mbar_init(bar_acc, 1);
        mbar_init(bar_act, kTileM);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(s_tmem)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int c = 0; c < S.n_chunks; ++c, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_empty + s, ((it / kStages) & 1) ^ 1);
                    const ChunkDesc cd = S.chunks[c];
                    mbar_expect_tx(bar_full + s, cd.bytes);
                    tma_bulk_g2s(s_stage + s * kStageBytes, p.weights + cd.gmem_off, cd.bytes,
                                 bar_full + s);
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, act_uses = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int c = 0; c < S.n_chunks; ++c, ++it) {
                    const ChunkDesc cd = S.chunks[c];
                    if (cd.first) {  // A operand of this layer ready, TMEM drained
                        mbar_wait(bar_act, act_uses & 1);
                        ++act_uses;
                        tc_fence_after();
                    }
                    const int s = it % kStages;
                    mbar_wait(bar_full + s, (it / kStages) & 1);
                    tc_fence_after();
                    const uint8_t *abuf = cd.a_src == kSrcAct ? s_act : (cd.a_src == kSrcPE ? s_pe : s_aux);
                    const uint32_t a_sbo = cd.a_src == kSrcAct ? kActSBO : kSideSBO;
                    const uint32_t a_base = smem_u32(abuf) + (cd.a_k0 >> 3) * 128;
                    const uint32_t b_base = smem_u32(s_stage + s * kStageBytes);
                    const uint32_t b_sbo = (cd.kc >> 3) * 128;
                    const uint32_t idesc = umma_idesc(kTileM, cd.n);
                    for (int k = 0; k < cd.kc; k += 16) {
                        const uint64_t ad = umma_desc(a_base + (k >> 3) * 128, 128, a_sbo);
                        const uint64_t bd = umma_desc(b_base + (k >> 3) * 128, 128, b_sbo);
                        umma_bf16(tmem_base, ad, bd, idesc, (cd.first && k == 0) ? 0u : 1u);
                    }
                    tc_commit(bar_empty + s);          // frees the weight stage
                    if (cd.last) tc_commit(bar_acc);   // accumulator complete -> epilogue
                }
            }
        }
    } else {
        // ===================== PE + epilogue warps (thread = one row of the tile) ==========
        const int row = threadIdx.x;  // 0..127, TMEM lane
        uint32_t acc_uses = 0;
        const float *bias = p.biases;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int64_t lrow = (int64_t) tile * kTileM + row;
            const bool valid = lrow < p.rows;
            const int64_t grow = (valid && p.row_index) ? (int64_t) p.row_index[lrow] : lrow;
            // ---- inputs: positional encodings + appearance embedding -> A buffers ----
            float xin[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) xin[c] = (valid && c < p.in_dim) ? p.x[grow * p.in_dim + c] : 0.f;
            write_pe(s_pe, row, 0, kSideSBO, xin, p.pe_xyz_freqs, kSideK);
            if (p.need_viewdir) {
                write_pe(s_aux, row, 0, kSideSBO, xin + 3, p.pe_dir_freqs, 32);
            } else {
                for (int k = 0; k < 32; k += 8)
                    *reinterpret_cast<uint4 *>(s_aux + a_off(row, k, kSideSBO)) = make_uint4(0, 0, 0, 0);
            }
            if (p.app_dim > 0) {
                int ai = p.app_col >= 0 ? (int) xin[p.app_col] : 0;
                ai = min(max(ai, 0), p.n_appearance - 1);
                const float *e = p.embedding + (size_t) ai * p.app_dim;
                for (int k = 0; k < kSideK - 32; ++k)
                    *reinterpret_cast<__nv_bfloat16 *>(s_aux + a_off(row, 32 + k, kSideSBO)) =
                            __float2bfloat16_rn(k < p.app_dim ? __ldg(e + k) : 0.f);
            } else {
                for (int k = 32; k < kSideK; k += 8)
                    *reinterpret_cast<uint4 *>(s_aux + a_off(row, k, kSideSBO)) = make_uint4(0, 0, 0, 0);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_act);

            float sigma = 0.f;
            for (int l = 0; l < S.n_layers; ++l) {
                const LayerDesc ld = S.layers[l];
                mbar_wait(bar_acc, acc_uses & 1);
                ++acc_uses;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t) (warp * 32) << 16);
                for (int c0 = 0; c0 < ld.n; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    if (ld.epilogue == kEpiOut) {
                        if (valid) {
                            for (int j = 0; j < 32; ++j) {
                                const int c = c0 + j;
                                if (c < p.out_real)
                                    p.out[grow * p.out_stride + c] =
                                            __uint_as_float(r[j]) + __ldg(bias + ld.bias_off + c);
                            }
                        }
                    } else {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            v[j] = __uint_as_float(r[j]) + __ldg(bias + ld.bias_off + c0 + j);
                            if (ld.epilogue != kEpiLinearAct) v[j] = fmaxf(v[j], 0.f);
                        }
                        if (ld.epilogue == kEpiReluActSigma) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                sigma = fmaf(v[j], __ldg(bias + p.sigma_w_off + c0 + j), sigma);
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 q;
                            q.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]);
                            q.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
                            q.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]);
                            q.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
                            *reinterpret_cast<uint4 *>(s_act + a_off(row, c0 + 8 * g, kActSBO)) = q;
                        }
                    }
                }
                if (ld.epilogue == kEpiOut) {
                    if (valid) {
                        float sg = sigma + __ldg(bias + p.sigma_b_off);
                        sg = p.sigma_activation == 1 ? (sg > 20.f ? sg : log1pf(expf(sg)))
                                                     : fmaxf(sg, 0.f);
                        p.out[grow * p.out_stride + p.out_real] = sg;
                    }
                    tc_fence_before();  // TMEM reads ordered before the next tile's first MMA
                } else {
                    fence_async_smem();  // generic-proxy smem writes -> visible to the UMMA proxy
                    tc_fence_before();
                    mbar_arrive(bar_act);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(kTmemCols));
    }
}

constexpr size_t kMlpSmemBytes =
        kActBytes + 2 * kSideBytes + kStages * kStageBytes + (2 * kStages + 2) * 8 + 16;

// ------------------------------------------------------------------ host packing
uint16_t f2bf(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t) ((u >> 16) | 0x40u);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
    return (uint16_t) (u >> 16);
}

struct KSeg {  // a run of K columns of one layer's A operand
    int a_src, a_k0, k_len;       // in the A buffer (k_len multiple of 16)
    int w_col0, w_cols;           // real weight columns covered: [w_col0, w_col0 + w_cols)
};

}  // namespace

struct MlpModel {
    int device = 0;
    uint8_t *weights = nullptr;
    float *biases = nullptr;
    float *embedding = nullptr;
    MlpSchedule *sched_dev = nullptr;
    MlpSchedule sched;
    mnv_mlp_desc cfg;
    int sigma_w_off = 0, sigma_b_off = 0;
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
        d.head_width % 16 || d.out_rgb_dim < 1 || d.out_rgb_dim > 255)
        return fail(MNV_ERR_INVALID, "unsupported MLP shape (width must be 256, head <= 256)");
    const int pe = 3 + 6 * d.pe_xyz_freqs;
    const int pe_dir = d.need_viewdir ? 3 + 6 * d.pe_dir_freqs : 0;
    if (pe > kSideK || pe_dir > 32 || d.appearance_dim > kSideK - 32)
        return fail(MNV_ERR_INVALID, "positional encoding / appearance embedding too wide");
    if (d.skip_layer >= d.n_trunk_layers)
        return fail(MNV_ERR_INVALID, "skip_layer outside the trunk");

    auto *m = new MlpModel();
    m->device = device;
    m->cfg = d;
    m->in_dim = 3 + (d.need_viewdir ? 3 : 0) + (d.appearance_dim > 0 ? 1 : 0);
    m->out_dim = d.out_rgb_dim + 1;
    MlpSchedule &S = m->sched;
    std::memset(&S, 0, sizeof(S));
    std::vector<uint8_t> blob;
    std::vector<float> bias;
    double macs = 0;

    auto add_layer = [&](const float *W, const float *b, int n_real, int k_real_total,
                         const std::vector<KSeg> &segs, int epilogue) {
        LayerDesc &L = S.layers[S.n_layers];
        L.n = (uint16_t) ((n_real + 31) / 32 * 32);  // the epilogue reads 32 TMEM columns at a time
        L.n_real = (uint16_t) n_real;
        L.bias_off = (uint32_t) bias.size();
        L.epilogue = (uint32_t) epilogue;
        const int n_mma = L.n;  // epilogue reads 32 columns at a time
        for (int c = 0; c < n_mma; ++c) bias.push_back(c < n_real ? b[c] : 0.f);
        bool first = true;
        for (size_t si = 0; si < segs.size(); ++si) {
            const KSeg &sg = segs[si];
            for (int k0 = 0; k0 < sg.k_len; k0 += kChunkK) {
                const int kc = std::min(kChunkK, sg.k_len - k0);
                ChunkDesc &C = S.chunks[S.n_chunks++];
                C.gmem_off = (uint32_t) blob.size();
                C.bytes = (uint32_t) (n_mma * kc * 2);
                C.n = (uint16_t) n_mma;
                C.kc = (uint16_t) kc;
                C.a_src = (uint16_t) sg.a_src;
                C.a_k0 = (uint16_t) (sg.a_k0 + k0);
                C.first = first;
                C.last = 0;
                C.layer = (uint8_t) S.n_layers;
                first = false;
                blob.resize(blob.size() + C.bytes, 0);
                uint8_t *dst = blob.data() + C.gmem_off;
                const int sbo = (kc / 8) * 128;
                for (int n = 0; n < n_real; ++n)
                    for (int k = 0; k < kc; ++k) {
                        const int kk = k0 + k;  // column inside the segment
                        if (kk >= sg.w_cols) continue;
                        const float w = W[(size_t) n * k_real_total + sg.w_col0 + kk];
                        const uint16_t h = f2bf(w);
                        std::memcpy(dst + (n / 8) * sbo + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2, &h, 2);
                    }
            }
        }
        S.chunks[S.n_chunks - 1].last = 1;
        macs += (double) n_real * k_real_total;
        ++S.n_layers;
    };

    // trunk
    for (int l = 0; l < d.n_trunk_layers; ++l) {
        std::vector<KSeg> segs;
        int k_total;
        if (l == 0) {
            segs.push_back({kSrcPE, 0, (pe + 15) / 16 * 16, 0, pe});
            k_total = pe;
        } else if (l == d.skip_layer) {  // cat([pe, h]) like the NeRF skip connection
            segs.push_back({kSrcAct, 0, 256, pe, 256});
            segs.push_back({kSrcPE, 0, (pe + 15) / 16 * 16, 0, pe});
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
    // head 1: [feature(256) | dir PE | appearance]
    {
        std::vector<KSeg> segs;
        segs.push_back({kSrcAct, 0, 256, 0, 256});
        int col = 256;
        if (pe_dir > 0) {
            segs.push_back({kSrcAux, 0, 32, col, pe_dir});
            col += pe_dir;
        }
        if (d.appearance_dim > 0) {
            segs.push_back({kSrcAux, 32, (d.appearance_dim + 15) / 16 * 16, col, d.appearance_dim});
            col += d.appearance_dim;
        }
        add_layer(d.head1_w, d.head1_b, d.head_width, col, segs, kEpiReluAct);
    }
    // head 2
    add_layer(d.head2_w, d.head2_b, d.out_rgb_dim, d.head_width,
              {{kSrcAct, 0, d.head_width, 0, d.head_width}}, kEpiOut);
    macs += 256;  // sigma head
    m->flops_per_row = 2.0 * macs;
    m->sigma_w_off = (int) bias.size();
    for (int c = 0; c < 256; ++c) bias.push_back(d.sigma_w[c]);
    m->sigma_b_off = (int) bias.size();
    bias.push_back(d.sigma_b[0]);

    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->weights, blob.size());
    if (e == cudaSuccess) e = cudaMemcpy(m->weights, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->biases, bias.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->biases, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d.appearance_dim > 0) {
        const size_t nb = (size_t) d.n_appearance * d.appearance_dim * 4;
        e = cudaMalloc(&m->embedding, nb);
        if (e == cudaSuccess) e = cudaMemcpy(m->embedding, d.embedding, nb, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&m->sched_dev, sizeof(MlpSchedule));
    if (e == cudaSuccess) e = cudaMemcpy(m->sched_dev, &S, sizeof(S), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMlpSmemBytes);
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
    cudaFree(m->biases);
    cudaFree(m->embedding);
    cudaFree(m->sched_dev);
    delete m;
}

int mlp_forward(const MlpModel *m, const float *x_dev, int64_t rows, int in_dim, float *out_dev,
                int out_stride, cudaStream_t stream) {
    return mlp_forward_indexed(m, x_dev, nullptr, rows, in_dim, out_dev, out_stride, stream);
}

int mlp_forward_indexed(const MlpModel *m, const float *x_dev, const int32_t *row_index_dev, int64_t rows,
                        int in_dim, float *out_dev, int out_stride, cudaStream_t stream) {
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
    p.biases = m->biases;
    p.embedding = m->embedding;
    p.sched = m->sched_dev;
    p.x = x_dev;
    p.out = out_dev;
    p.row_index = row_index_dev;
    p.rows = rows;
    p.in_dim = in_dim;
    p.out_stride = out_stride;
    p.n_tiles = (int) ((rows + kTileM - 1) / kTileM);
    p.pe_xyz_freqs = m->cfg.pe_xyz_freqs;
    p.pe_dir_freqs = m->cfg.pe_dir_freqs;
    p.need_viewdir = m->cfg.need_viewdir;
    p.app_dim = m->cfg.appearance_dim;
    p.n_appearance = m->cfg.n_appearance;
    p.app_col = m->cfg.appearance_dim > 0 ? in_dim - 1 : -1;
    p.sigma_w_off = m->sigma_w_off;
    p.sigma_b_off = m->sigma_b_off;
    p.sigma_activation = m->cfg.sigma_activation;
    p.out_real = m->cfg.out_rgb_dim;
    const int grid = std::min(p.n_tiles, m->num_sms);
    mlp_forward_kernel<<<grid, kMlpThreads, kMlpSmemBytes, stream>>>(p);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

double mlp_flops_per_row(const MlpModel *m) { return m->flops_per_row; }
int mlp_in_dim(const MlpModel *m) { return m->in_dim; }
int mlp_out_dim(const MlpModel *m) { return m->out_dim; }

}  // namespace mnv
