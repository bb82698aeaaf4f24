// Octree traversal + volume-rendering integration for sm_100a.
//
// Replaces the reference's render_voxels_kernel
// (src/cuda/renderer_kernel.cu:243-292) and device::render_voxels_trace_ray /
// query_single_from_root (include/cuda/rt_core.cuh:117-332).
//
// What is different from the reference (all of it layout / schedule, none of
// it arithmetic — see mnv_math.cuh for the numeric contract):
//   * query: the reference restarts at the root for every step of every ray
//     (depth dependent loads).  The descent x*=2; f=floor(x); x-=f is exact in
//     fp32, so the leaf containing pos is exactly the integer cell
//     floor(pos*2^d).  One 8-byte load of the ANCHOR GRID — a dense 2^A-cube
//     (A = 8) over the unit box, index from three FFMA.RM against the magic
//     constant 2^23 — resolves the top A levels: it is the leaf itself when
//     the tree ends at depth <= A there, else the level-A node, from which at
//     most depth - A cell words remain (three funnel shifts per level build
//     node * 8 + child).  Trees without an anchor grid (MNV_ANCHOR_LEVEL=0)
//     take the round-1 path: each ray keeps the node path of its previous
//     leaf in shared memory and re-descends below the deepest common ancestor.
//   * one 4-byte "cell" word per slot holds child link OR (leaf, sigma, sample
//     count): an empty leaf visit costs exactly one load.
//   * shaded leaves fetch one aligned 64-byte record with 4 x LDG.128.
//   * warps own 8x4-pixel tiles (rays of a warp stay in neighbouring leaves)
//     instead of 32x1 row segments.
//   * split / re-sample candidates live in registers and are written once per
//     ray; camera and options travel as kernel parameters (no 48-byte upload).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "mnv_internal.cuh"
#include "mnv_math.cuh"

namespace mnv {
namespace {

#ifndef MNV_TILE_H
#define MNV_TILE_H 8   // CTA tile = 16 x MNV_TILE_H pixels (8 -> 4 warps, 16 -> 8 warps)
#endif
#ifndef MNV_UNROLL2
#define MNV_UNROLL2 1  // two steps per loop trip: the previous-cell registers rotate instead of being copied
#endif
#ifndef MNV_UNROLL_TRACK
#define MNV_UNROLL_TRACK 0  // also unroll the candidate-tracking variants
#endif
#ifndef MNV_UNROLL_ANCHOR
#define MNV_UNROLL_ANCHOR 1  // loop trips unrolled in the anchor-grid variants
#endif
#ifndef MNV_LAZY_EMPTY
#define MNV_LAZY_EMPTY 1  // candidates from empty leaves kept in registers and committed once per ray
#endif
#ifndef MNV_TRACK_REGS
#define MNV_TRACK_REGS 1  // candidate trackers in registers instead of shared memory
#endif
#ifndef MNV_SMEM_STATE
#define MNV_SMEM_STATE 1  // park SH basis + shaded-only ray state in shared memory
#endif
#ifndef MNV_CTA_WARPS
// warps per CTA.  Every warp owns an 8x4-pixel tile and never talks to the others (no barrier in the kernel), so a CTA
// is only a scheduling unit: its warp slots, registers and shared memory stay allocated until its LAST warp's longest
// ray ends.  Fewer warps per CTA free those slots earlier and let the tail of the frame spread over more SMs.
#define MNV_CTA_WARPS (16 * MNV_TILE_H / 32)
#endif
#ifndef MNV_PACK_PRIO
// candidate priorities are not carried through the march: the split priority (a depth) rides in bits 8..15 of `flags`,
// the re-sample priority (the leaf's sample count) is re-read from the cell word once per ray
#define MNV_PACK_PRIO 1
#endif
#ifndef MNV_DDA_SIGN
#define MNV_DDA_SIGN 1  // exit distance of the unit cube by per-ray sign selection instead of three max()
#endif
constexpr int kWarps = MNV_CTA_WARPS;
constexpr int kThreads = 32 * kWarps;  // each warp owns an 8x4-pixel tile
#ifndef MNV_MIN_BLOCKS
#define MNV_MIN_BLOCKS (1024 / (32 * MNV_CTA_WARPS))  // resident CTAs per SM the register allocation targets
#endif
#ifndef MNV_MIN_BLOCKS_PLAIN
// without candidate tracking the anchored march fits 56 registers: 9 CTAs per SM (measured 0.940 -> 0.906 ms;
// the tracking variant spills at that budget and is slower, 1.004 -> 1.036 ms)
#define MNV_MIN_BLOCKS_PLAIN (MNV_CTA_WARPS == 4 ? 9 : (MNV_CTA_WARPS == 2 ? 18 : MNV_MIN_BLOCKS))
#endif
constexpr int kTileW = 16, kTileH = MNV_TILE_H;  // launch-order tile; the multi-GPU partition is a multiple of 16x8
constexpr int kWarpsPerTile = kTileW * kTileH / 32;
constexpr int kMaxLevel = 22;    // q carries 23 bits per axis: leaf depth <= 23

// words of per-ray state parked in shared memory (render_pixel)
enum { kRsDeltaScale = 0, kRsT, kRsOut0, kRsOut1, kRsOut2, kRsWordsBase,
       kRsMaxW = kRsWordsBase, kRsMaxSW, kRsSplitId, kRsSplitPrio, kRsSampId, kRsSampPrio,
       kRsWordsTrack };

struct RenderParams {
    TreeView tree;
    mnv_camera cam;
    mnv_render_options opt;
    RenderTargets tg;
    int mtiles_x;   // partition tiles per row (multi-GPU partition)
    int tiles_x;    // 16x8-pixel tiles per row
    int n_tiles;    // tiles of the frame
    int max_level;  // deepest level a descent may reach (tree max leaf depth - 1, <= 22)
    int path_levels;  // rows of the shared-memory node path (= max_level + 1; 0 with the anchor grid)
    uint32_t samp_limit;  // max_sample_count << 16, saturated: (cell word & 0x7fffffff) < samp_limit <=> count < max_sample_count
    int split_limit;  // min(opt.max_depth, 23): leaves at depth 23 cannot be split (23-bit cell coordinates)
    // anchor grid (mnv_internal.cuh): entry index = ((ax * dim + ay) * dim + az) - bias with a? the raw bits of
    // fma_rd(p?, 2^A, 2^23) — the 0x4B000000 exponents of the three terms fold into one constant
    const uint2 *anchor;
    int anchor_level;
    float anchor_scale;    // 2^A
    uint32_t anchor_dim;   // 2^A
    uint32_t anchor_bias;
};

template <int R>
__device__ __forceinline__ float rec_half(const uint32_t (&w)[R], int h) {
    const uint32_t v = w[h >> 1];
    return __half2float(__ushort_as_half((unsigned short) ((h & 1) ? (v >> 16) : (v & 0xffffu))));
}

// One colour channel: tmp = B0*C0 (+ per-degree groups, each an FMUL + FFMA
// chain then one FADD) — include/cuda/rt_core.cuh:257-286 as compiled.
template <int TERMS, int R>
__device__ __forceinline__ float sh_channel(const float (&B)[TERMS], const uint32_t (&w)[R],
                                            int off) {
    float tmp = __fmul_rn(B[0], rec_half(w, off));
    if constexpr (TERMS >= 25) {
        float s = __fmul_rn(B[17], rec_half(w, off + 17));
        s = __fmaf_rn(B[16], rec_half(w, off + 16), s);
#pragma unroll
        for (int k = 18; k <= 24; ++k) s = __fmaf_rn(B[k < TERMS ? k : 0], rec_half(w, off + k), s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 16) {
        float s = __fmul_rn(B[10 < TERMS ? 10 : 0], rec_half(w, off + 10));
        s = __fmaf_rn(B[9 < TERMS ? 9 : 0], rec_half(w, off + 9), s);
#pragma unroll
        for (int k = 11; k <= 15; ++k) s = __fmaf_rn(B[k < TERMS ? k : 0], rec_half(w, off + k), s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 9) {
        float s = __fmul_rn(B[5 < TERMS ? 5 : 0], rec_half(w, off + 5));
        s = __fmaf_rn(B[4 < TERMS ? 4 : 0], rec_half(w, off + 4), s);
#pragma unroll
        for (int k = 6; k <= 8; ++k) s = __fmaf_rn(B[k < TERMS ? k : 0], rec_half(w, off + k), s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 4) {
        float s = __fmul_rn(B[2 < TERMS ? 2 : 0], rec_half(w, off + 2));
        s = __fmaf_rn(B[1 < TERMS ? 1 : 0], rec_half(w, off + 1), s);
        s = __fmaf_rn(B[3 < TERMS ? 3 : 0], rec_half(w, off + 3), s);
        tmp = __fadd_rn(tmp, s);
    }
    return tmp;
}

// TERMS: 0 = RGBA, else SH basis dimension (1,4,9,16,25).
// TRACK: produce split / re-sample candidates.  LOGV: visit hash/count/log/stats.
// VISIT: mark visited nodes (track_visit).
template <int TERMS, bool TRACK, bool LOGV, bool VISIT, bool ANCHOR>
__device__ __forceinline__ void render_pixel(const RenderParams &p, const int x, const int y,
                                             int32_t *__restrict__ s_path,
                                             float *__restrict__ s_basis,
                                             float *__restrict__ s_ray) {
    const int W = p.cam.width;
    const int idx = y * W + x;
    const mnv_render_options &opt = p.opt;

    uint32_t rgbx_init = 0;
    if (!p.tg.offscreen) rgbx_init = surf2Dread<uint32_t>(p.tg.image_surf, x * 4, y, cudaBoundaryModeZero);

    // Per-ray state that only shaded leaves touch lives in shared memory (word k of
    // this thread at s_ray[k * kThreads]) so that the march itself needs few registers.
#if MNV_SMEM_STATE
#define RS(k) s_ray[(k) * kThreads]
#define RSI(k) reinterpret_cast<int32_t *>(s_ray)[(k) * kThreads]
#else
    float rs_regs[kRsWordsTrack];
#define RS(k) rs_regs[k]
#define RSI(k) reinterpret_cast<int32_t *>(rs_regs)[k]
#endif
    RS(kRsOut0) = 0.f;
    RS(kRsOut1) = 0.f;
    RS(kRsOut2) = 0.f;
    float out3 = 0.f;
    // candidate trackers: shared memory like the rest of the shaded-only state, or (MNV_TRACK_REGS) plain registers —
    // the compiler keeps register copies of them anyway and then pays the stores on top
#if MNV_TRACK_REGS
    float trk_f[2];
    int32_t trk_i[4];
#define TS(k) trk_f[(k) - kRsMaxW]
#define TSI(k) trk_i[(k) - kRsSplitId]
#else
#define TS(k) RS(k)
#define TSI(k) RSI(k)
#endif

    // ---- ray generation: screen2worlddir, renderer_kernel.cu:30-38 ----------
    const float *m = p.cam.c2w;
    const float vx = __fdiv_rn(__fadd_rn(__fadd_rn((float) x, 0.5f), -p.cam.cx), p.cam.fx);
    const float vy = __fdiv_rn(-__fadd_rn(__fadd_rn((float) y, 0.5f), -p.cam.cy), p.cam.fy);
    float d0 = __fadd_rn(__fmaf_rn(vx, m[0], __fmul_rn(vy, m[3])), -m[6]);
    float d1 = __fadd_rn(__fmaf_rn(vx, m[1], __fmul_rn(vy, m[4])), -m[7]);
    float d2 = __fadd_rn(__fmaf_rn(vx, m[2], __fmul_rn(vy, m[5])), -m[8]);
    {
        const float inv = __frcp_rn(ref_norm3(d0, d1, d2));
        d0 = __fmul_rn(d0, inv);
        d1 = __fmul_rn(d1, inv);
        d2 = __fmul_rn(d2, inv);
    }
    // cen = offset + scale * cen, renderer_kernel.cu:272-275 (FFMA)
    const float c0 = __fmaf_rn(p.tree.scale[0], m[9], p.tree.offset[0]);
    const float c1 = __fmaf_rn(p.tree.scale[1], m[10], p.tree.offset[1]);
    const float c2 = __fmaf_rn(p.tree.scale[2], m[11], p.tree.offset[2]);

    float tmax_bg = 1e9f;
    if (!p.tg.offscreen) tmax_bg = surf2Dread<float>(p.tg.depth_surf, x * 4, y, cudaBoundaryModeZero);

    float v0 = d0, v1 = d1, v2 = d2;  // view direction for the SH basis
    ref_rodrigues(opt.rot_dirs, v0, v1, v2);

    // ---- render_voxels_trace_ray, rt_core.cuh:162-332 ----------------------
    if (TRACK) {
        TSI(kRsSplitPrio) = opt.max_depth + 1;  // priorities stay integers during the march, converted once per ray
        TSI(kRsSampPrio) = opt.max_sample_count + 1;
        TSI(kRsSplitId) = -1;  // packed leaf id node*8 + child
        TSI(kRsSampId) = -1;
        TS(kRsMaxW) = -1.f;
        TS(kRsMaxSW) = -1.f;
    }
    unsigned long long vhash = 0xcbf29ce484222325ULL;
    int nvis = 0, nshaded = 0;
    bool hit = false;

    // _get_delta_scale, rt_core.cuh:102-115
    d0 = __fmul_rn(d0, p.tree.scale[0]);
    d1 = __fmul_rn(d1, p.tree.scale[1]);
    d2 = __fmul_rn(d2, p.tree.scale[2]);
    const float delta_scale = __frcp_rn(ref_norm3(d0, d1, d2));
    d0 = __fmul_rn(d0, delta_scale);
    d1 = __fmul_rn(d1, delta_scale);
    d2 = __fmul_rn(d2, delta_scale);
    tmax_bg = __fdiv_rn(tmax_bg, delta_scale);
    RS(kRsDeltaScale) = delta_scale;
    RS(kRsT) = 1.f;

    // invdir = 1.f / (dir + 1e-9): double add + double reciprocal, rt_core.cuh:188-190
    const float i0 = d2f(__drcp_rn(__dadd_rn((double) d0, 1e-9)));
    const float i1 = d2f(__drcp_rn(__dadd_rn((double) d1, 1e-9)));
    const float i2 = d2f(__drcp_rn(__dadd_rn((double) d2, 1e-9)));

    // _dda_world, rt_core.cuh:70-86 (double, rounded to float per term)
    float tmin = 0.f, tmax = 1e4f;
    {
        const float cc[3] = {c0, c1, c2};
        const float ii[3] = {i0, i1, i2};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double ci = (double) cc[i], inv = (double) ii[i];
            const float t1 = d2f(__dmul_rn(
                    __dadd_rn(__dadd_rn((double) opt.render_bbox[i], 1e-6), -ci), inv));
            const float t2 = d2f(__dmul_rn(
                    __dadd_rn(__dadd_rn((double) opt.render_bbox[i + 3], -1e-6), -ci), inv));
            tmin = fmaxf(tmin, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    if (p.tg.partial_n > 0 && p.tg.has_cell)
        clip_to_cell(p.tg.cell_box, c0, c1, c2, i0, i1, i2, opt.step_size, tmin, tmax);
    tmax = fminf(tmax, tmax_bg);

    if (tmax < 0.f || tmin > tmax) {
        if (opt.render_depth) out3 = 1.f;
    } else {
        hit = true;
        float B[TERMS > 0 ? TERMS : 1];
        if (TERMS > 0) {
            // SH basis of the view direction: needed only by shaded leaves; with
            // MNV_SMEM_STATE it is parked in shared memory instead of 9..25 registers
            ref_sh_basis<(TERMS > 0 ? TERMS : 1)>(v0, v1, v2, B);
#pragma unroll
            for (int k = 0; k < TERMS; ++k) {
                if (k < opt.basis_minmax[0] || k > opt.basis_minmax[1]) B[k] = 0.f;
#if MNV_SMEM_STATE
                s_basis[k * kThreads] = B[k];
#endif
            }
        }
        constexpr int REC_W = TERMS > 0 ? ((3 * TERMS + 1 + 7) / 8) * 4 : 4;  // u32 words / record

        float t = tmin;
        // bit 0 / 1: a shaded leaf already set the split / re-sample candidate
        // (max_weight / max_sample_weight != -1 in rt_core.cuh:308-321); bit 2: stopped early
        uint32_t flags = 0;
#if MNV_LAZY_EMPTY
        // last empty leaf that qualifies as split / re-sample candidate, whatever the flags say: read only if no
        // shaded leaf ever set the candidate (rt_core.cuh:308-321 updates them while max_weight == -1)
        uint32_t e_split = 0xffffffffu, e_samp = 0xffffffffu;
        int e_depth = 0;
#endif
        // raw bits of (floor(pos * 2^23) + 2^23) as float: 0x4B000000 | q, q = 23-bit cell coords
        uint32_t pqx = 0x4B000000u, pqy = 0x4B000000u, pqz = 0x4B000000u;
        int pdepth = 1;  // previous leaf depth: path valid for levels < pdepth
        const float clamp_hi = f_from_bits(0x3F7FFFEFu);  // 1.f - 1e-6f
        const uint32_t *__restrict__ cells = p.tree.cell;
#if MNV_DDA_SIGN
        const float ip0 = fmaxf(i0, 0.f), ip1 = fmaxf(i1, 0.f), ip2 = fmaxf(i2, 0.f);
#endif

        // two steps per loop trip (the previous-cell registers rotate instead of being copied); the tracking
        // variants are register-bound and keep one
        constexpr int kUnroll = ANCHOR ? MNV_UNROLL_ANCHOR : (MNV_UNROLL2 && (MNV_UNROLL_TRACK || !TRACK) ? 2 : 1);
#pragma unroll kUnroll
        while (t < tmax) {
            // pos = cen + t*dir (FFMA), clamp to [0, 1-1e-6] (rt_core.cuh:221-223,125-127);
            // .SAT gives the clamp to [0,1] for free, min() finishes it.
            const float px = fminf(__saturatef(__fmaf_rn(t, d0, c0)), clamp_hi);
            const float py = fminf(__saturatef(__fmaf_rn(t, d1, c1)), clamp_hi);
            const float pz = fminf(__saturatef(__fmaf_rn(t, d2, c2)), clamp_hi);
            uint32_t cw, slot;
            int lvl;
            if constexpr (ANCHOR) {
                // One 8-byte load of the anchor grid resolves the top A levels: either the leaf itself (tree ends at
                // depth <= A here) or the level-A node, from which at most (depth - A) cell words remain.  No per-ray
                // node path, no previous-cell registers: with 32 rays per warp the old path cache paid the deepest
                // lane's re-descent (warp-max 3.4-4 rounds per step) on almost every step.
                const uint32_t ax = __float_as_uint(__fmaf_rd(px, p.anchor_scale, 8388608.f));
                const uint32_t ay = __float_as_uint(__fmaf_rd(py, p.anchor_scale, 8388608.f));
                const uint32_t az = __float_as_uint(__fmaf_rd(pz, p.anchor_scale, 8388608.f));
                const uint2 e = __ldg(p.anchor + ((ax * p.anchor_dim + ay) * p.anchor_dim + az - p.anchor_bias));
                lvl = (int) (e.y >> 28);
                cw = e.x;
                slot = e.y & 0x0fffffffu;
                if ((int32_t) e.x >= 0) {
                    // exact integer cell coordinates at level 23: floor(p * 2^23) sits in the mantissa of
                    // fma_rd(p, 2^23, 2^23); the level-lvl child bit of each axis moves to bit 31
                    uint32_t node = e.x;
                    uint32_t sx = __float_as_uint(__fmaf_rd(px, 8388608.f, 8388608.f)) << (9 + lvl);
                    uint32_t sy = __float_as_uint(__fmaf_rd(py, 8388608.f, 8388608.f)) << (9 + lvl);
                    uint32_t sz = __float_as_uint(__fmaf_rd(pz, 8388608.f, 8388608.f)) << (9 + lvl);
                    for (;;) {
                        slot = __funnelshift_l(sz, __funnelshift_l(sy, __funnelshift_l(sx, node, 1), 1), 1);
                        cw = __ldg(cells + slot);
                        if ((int32_t) cw < 0 || lvl >= p.max_level) break;
                        node = cw;
                        ++lvl;
                        sx <<= 1;
                        sy <<= 1;
                        sz <<= 1;
                    }
                }
                if (VISIT) {
                    // the node holding the leaf; its ancestors are marked by launch_propagate_visited afterwards
                    // (the reference marks every node of the root path at every query, rt_core.cuh:132-135)
                    if (p.tg.visited[slot >> 3] == 0) p.tg.visited[slot >> 3] = 1;
                }
            } else {
                // exact integer cell coordinates at level 23: floor(p * 2^23) sits in the
                // mantissa of fma_rd(p, 2^23, 2^23)
                const uint32_t qx = __float_as_uint(__fmaf_rd(px, 8388608.f, 8388608.f));
                const uint32_t qy = __float_as_uint(__fmaf_rd(py, 8388608.f, 8388608.f));
                const uint32_t qz = __float_as_uint(__fmaf_rd(pz, 8388608.f, 8388608.f));
                const uint32_t diff = (qx ^ pqx) | (qy ^ pqy) | (qz ^ pqz);  // exponent bits cancel
                pqx = qx;
                pqy = qy;
                pqz = qz;
                // number of leading (from bit 22) bits shared with the previous cell
                lvl = min(__clz((int) diff) - 9, pdepth - 1);
                uint32_t node = lvl > 0 ? (uint32_t) s_path[lvl * kThreads] : 0u;
                // level-lvl child bit of each axis moved to bit 31; three funnel shifts append the x, y, z bits to
                // node: slot = node * 8 + child in 3 ALU instructions per level instead of 6
                uint32_t sx = qx << (9 + lvl), sy = qy << (9 + lvl), sz = qz << (9 + lvl);
                for (;;) {
                    if (VISIT) {
                        if (p.tg.visited[node] == 0) p.tg.visited[node] = 1;
                    }
                    slot = __funnelshift_l(sz, __funnelshift_l(sy, __funnelshift_l(sx, node, 1), 1), 1);
                    cw = __ldg(cells + slot);
                    if ((int32_t) cw < 0 || lvl >= p.max_level) break;
                    node = cw;
                    ++lvl;
                    sx <<= 1;
                    sy <<= 1;
                    sz <<= 1;
                    s_path[lvl * kThreads] = (int32_t) node;
                }
            }
            const int depth = lvl + 1;
            pdepth = depth;
            const float sigma = __half2float(__ushort_as_half((unsigned short) (cw & 0xffffu)));
            const bool shaded = (int32_t) cw < 0 && sigma > opt.sigma_thresh;
            const uint4 *rec = p.tree.payload + (size_t) slot * (REC_W / 4);
            if (LOGV) {
                const long long packed = (long long) slot;
                vhash = (vhash ^ (unsigned long long) packed) * 0x100000001b3ULL;
                if (p.tg.visit_log && nvis < p.tg.log_cap)
                    p.tg.visit_log[(size_t) idx * p.tg.log_cap + nvis] = (int32_t) packed;
                ++nvis;
            }

            // position inside the leaf, in leaf units: frac(pos * 2^depth) (exact)
            const float cube = __uint_as_float((uint32_t) (127 + depth) << 23);
            const float icube = __uint_as_float((uint32_t) (127 - depth) << 23);
            // floor(p*2^depth) via the 2^23 magic constant (exact), then an exact FFMA
            const float flx = __fadd_rn(__fmaf_rd(px, cube, 8388608.f), -8388608.f);
            const float fly = __fadd_rn(__fmaf_rd(py, cube, 8388608.f), -8388608.f);
            const float flz = __fadd_rn(__fmaf_rd(pz, cube, 8388608.f), -8388608.f);
            const float fx = __fmaf_rn(px, cube, -flx);
            const float fy = __fmaf_rn(py, cube, -fly);
            const float fz = __fmaf_rn(pz, cube, -flz);
            // _dda_unit, rt_core.cuh:88-100: FMUL then FADD (not fused in the reference build)
            float tm;
            {
#if MNV_DDA_SIGN
                // max(a1, a1 + i) is a1 + i for i > 0 and a1 for i < 0 (rounding is monotonic; i is finite and
                // non-zero by construction: 1 / (dir + 1e-9) in double); a1 >= +0 when i < 0, so adding +0 is exact
                const float a2 = __fadd_rn(__fmul_rn(-fx, i0), ip0);
                const float b2 = __fadd_rn(__fmul_rn(-fy, i1), ip1);
                const float e2 = __fadd_rn(__fmul_rn(-fz, i2), ip2);
                tm = fminf(fminf(fminf(a2, 1e4f), b2), e2);
#else
                const float a1 = __fmul_rn(-fx, i0), a2 = __fadd_rn(a1, i0);
                const float b1 = __fmul_rn(-fy, i1), b2 = __fadd_rn(b1, i1);
                const float e1 = __fmul_rn(-fz, i2), e2 = __fadd_rn(e1, i2);
                tm = fminf(fminf(fminf(fmaxf(a1, a2), 1e4f), fmaxf(b1, b2)), fmaxf(e1, e2));
#endif
            }
            // / cube_size (exact power of two), + step_size
            const float delta_t = __fadd_rn(__fmul_rn(tm, icube), opt.step_size);
            const int scount = (int) ((cw >> 16) & 0x7fffu);
            // scount < max_sample_count on the raw word (count in bits 16..30 above the sigma bits): one compare
            const bool samp_ok = (cw & 0x7fffffffu) < p.samp_limit;

            if (shaded) {
                if (LOGV) ++nshaded;
                float T = RS(kRsT);
                const float att =
                        ref_expf(__fmul_rn(__fmul_rn(RS(kRsDeltaScale), -delta_t), sigma));
                const float weight = __fmul_rn(T, __fadd_rn(1.f, -att));
                if (TRACK) {
                    if (weight > TS(kRsMaxW) && depth < p.split_limit) {
                        TSI(kRsSplitId) = (int32_t) slot;
                        TS(kRsMaxW) = weight;
#if MNV_PACK_PRIO
                        flags = (flags & 0xffff00ffu) | ((uint32_t) depth << 8) | 1u;
#else
                        TSI(kRsSplitPrio) = depth;
                        flags |= 1u;
#endif
                    }
                    if (weight > TS(kRsMaxSW) && samp_ok) {
                        TSI(kRsSampId) = (int32_t) slot;
#if !MNV_PACK_PRIO
                        TSI(kRsSampPrio) = scount;
#endif
                        TS(kRsMaxSW) = weight;
                        flags |= 2u;
                    }
                }
                float out0 = RS(kRsOut0), out1 = RS(kRsOut1), out2 = RS(kRsOut2);
                if (opt.render_depth) {
                    out0 = __fmaf_rn(t, weight, out0);
                } else {
                    uint32_t w[REC_W];
#pragma unroll
                    for (int j = 0; j < REC_W / 4; ++j) {
                        // plain __ldg: neighbouring warps re-use payload lines through L1 (L1::no_allocate / evict_first
                        // measured 5 % / 3 % slower, profiles/r2_variants.jsonl tags pl1 / pl2)
                        const uint4 v = __ldg(rec + j);
                        w[4 * j] = v.x;
                        w[4 * j + 1] = v.y;
                        w[4 * j + 2] = v.z;
                        w[4 * j + 3] = v.w;
                    }
                    if (TERMS > 0) {
#if MNV_SMEM_STATE
#pragma unroll
                        for (int k = 0; k < TERMS; ++k) B[k] = s_basis[k * kThreads];
#endif
                        out0 = __fadd_rn(out0, ref_weighted_sigmoid(
                                weight, sh_channel<(TERMS > 0 ? TERMS : 1), REC_W>(B, w, 0)));
                        out1 = __fadd_rn(out1, ref_weighted_sigmoid(
                                weight, sh_channel<(TERMS > 0 ? TERMS : 1), REC_W>(B, w, TERMS)));
                        out2 = __fadd_rn(out2, ref_weighted_sigmoid(
                                weight, sh_channel<(TERMS > 0 ? TERMS : 1), REC_W>(B, w, 2 * TERMS)));
                    } else {
                        out0 = __fmaf_rn(weight, rec_half(w, 0), out0);
                        out1 = __fmaf_rn(weight, rec_half(w, 1), out1);
                        out2 = __fmaf_rn(weight, rec_half(w, 2), out2);
                    }
                }
                T = __fmul_rn(T, att);
                RS(kRsT) = T;
                if (T < opt.stop_thresh) {
                    if (opt.render_depth) out0 = out1 = out2 = fminf(__fmul_rn(out0, 0.3f), 1.0f);
                    const float scale = __frcp_rn(__fadd_rn(1.f, -T));
                    RS(kRsOut0) = __fmul_rn(out0, scale);
                    RS(kRsOut1) = __fmul_rn(out1, scale);
                    RS(kRsOut2) = __fmul_rn(out2, scale);
                    out3 = 1.f;
                    flags |= 4u;  // terminated early
                    break;
                }
                RS(kRsOut0) = out0;
                RS(kRsOut1) = out1;
                RS(kRsOut2) = out2;
            } else if (TRACK) {
#if MNV_LAZY_EMPTY
                if (depth < p.split_limit) {
                    e_split = slot;
                    e_depth = depth;
                }
                if (samp_ok) e_samp = slot;
#else
                if (!(flags & 1u) && depth < p.split_limit) {
                    TSI(kRsSplitId) = (int32_t) slot;
                    TSI(kRsSplitPrio) = depth;
                }
                if (!(flags & 2u) && samp_ok) {
                    TSI(kRsSampId) = (int32_t) slot;
                    TSI(kRsSampPrio) = scount;
                }
#endif
            }
            t = __fadd_rn(t, delta_t);
        }
#if MNV_PACK_PRIO
        if (TRACK) {
            if (flags & 1u) TSI(kRsSplitPrio) = (int) ((flags >> 8) & 0xffu);
            if (flags & 2u) TSI(kRsSampPrio) = (int) ((__ldg(cells + TSI(kRsSampId)) >> 16) & 0x7fffu);
        }
#endif
#if MNV_LAZY_EMPTY
        if (TRACK) {
            if (!(flags & 1u) && e_split != 0xffffffffu) {
                TSI(kRsSplitId) = (int32_t) e_split;
                TSI(kRsSplitPrio) = e_depth;
            }
            if (!(flags & 2u) && e_samp != 0xffffffffu) {
                TSI(kRsSampId) = (int32_t) e_samp;
                TSI(kRsSampPrio) = (int) ((__ldg(cells + e_samp) >> 16) & 0x7fffu);
            }
        }
#endif
        if (!(flags & 4u)) {
            if (opt.render_depth) {
                const float dv = fminf(__fmul_rn(RS(kRsOut0), 0.3f), 1.0f);
                RS(kRsOut0) = dv;
                RS(kRsOut1) = dv;
                RS(kRsOut2) = dv;
                out3 = 1.f;
            } else {
                out3 = __fadd_rn(1.f, -RS(kRsT));
            }
        }
    }

    if (p.tg.partial_n > 0) {
        // sub-module split: this GPU's segment of the ray, premultiplied, straight into the owner's memory
        const int owner = idx / p.tg.partial_block;
        p.tg.partial_dst[owner][(size_t) p.tg.partial_slot * p.tg.partial_block + (idx - owner * p.tg.partial_block)] =
                make_float4(RS(kRsOut0), RS(kRsOut1), RS(kRsOut2), out3);
        return;
    }
    // ---- composite_and_write, renderer_kernel.cu:215-241 --------------------
    const float nalpha = __fadd_rn(1.f, -out3);
    float out0 = RS(kRsOut0), out1 = RS(kRsOut1), out2 = RS(kRsOut2);
    if (p.tg.offscreen) {
        const float remain = __fmul_rn(nalpha, opt.background_brightness);
        out0 = __fadd_rn(out0, remain);
        out1 = __fadd_rn(out1, remain);
        out2 = __fadd_rn(out2, remain);
    } else {
        const float r0 = __fdiv_rn((float) (rgbx_init & 0xffu), 255.f);
        const float r1 = __fdiv_rn((float) ((rgbx_init >> 8) & 0xffu), 255.f);
        const float r2 = __fdiv_rn((float) ((rgbx_init >> 16) & 0xffu), 255.f);
        out0 = __fmaf_rn(r0, nalpha, out0);
        out1 = __fadd_rn(__fmul_rn(r1, nalpha), out1);
        out2 = __fadd_rn(__fmul_rn(r2, nalpha), out2);
    }
    const uint32_t rgba = ref_to_u8(out0) | (ref_to_u8(out1) << 8) | (ref_to_u8(out2) << 16) | 0xff000000u;
    if (p.tg.image_linear)
        reinterpret_cast<uint32_t *>(p.tg.image_linear)[idx] = rgba;
    else
        surf2Dwrite(rgba, p.tg.image_surf, x * 4, y, cudaBoundaryModeZero);

    if (TRACK) {
        // (priority, chunk, child) as floats, like the reference's trackers (ids >= 2^24: tracker_encode_chunk)
        const int32_t sid = TSI(kRsSplitId), pid = TSI(kRsSampId);
        float *ts = p.tg.to_split + (size_t) idx * 3;
        ts[0] = (float) TSI(kRsSplitPrio);
        ts[1] = sid < 0 ? -1.f : tracker_encode_chunk(sid >> 3);
        ts[2] = sid < 0 ? -1.f : (float) (sid & 7);
        float *tp = p.tg.to_sample + (size_t) idx * 3;
        tp[0] = (float) TSI(kRsSampPrio);
        tp[1] = pid < 0 ? -1.f : tracker_encode_chunk(pid >> 3);
        tp[2] = pid < 0 ? -1.f : (float) (pid & 7);
    }
#undef RS
#undef RSI
#undef TS
#undef TSI
    if (LOGV) {
        if (p.tg.visit_hash) p.tg.visit_hash[idx] = vhash;
        if (p.tg.visit_count) p.tg.visit_count[idx] = nvis;
        if (p.tg.shaded_count) p.tg.shaded_count[idx] = nshaded;
        if (p.tg.frame_stats) {
            atomicAdd(p.tg.frame_stats + 0, 1ull);
            atomicAdd(p.tg.frame_stats + 1, (unsigned long long) nvis);
            atomicAdd(p.tg.frame_stats + 2, (unsigned long long) nshaded);
            atomicAdd(p.tg.frame_stats + 3, hit ? 1ull : 0ull);
        }
    }
}

// ---- ray-coherent tiles ----------------------------------------------------------
// One CTA = 4 warps = a 16x8-pixel tile; each warp owns an 8x4-pixel sub-tile, so the
// 32 rays of a warp stay in neighbouring leaves (the reference maps a warp to 32
// pixels of one image row).  CTAs are dispatched in row-major tile order.  A
// persistent variant (per-SM Morton segments + work stealing) was measured slower on
// B200 ("rejected"): the march is bound by dependent-load latency, and the
// plain grid keeps more independent warps in flight.
//
// TERMS: 0 = RGBA, else SH basis dimension (1,4,9,16,25).
// TRACK: produce split / re-sample candidates.  LOGV: visit hash/count/log/stats.
// VISIT: mark visited nodes (track_visit).
template <int TERMS, bool TRACK, bool LOGV, bool VISIT, bool ANCHOR>
__global__ void __launch_bounds__(kThreads, (ANCHOR && !TRACK && !LOGV && !VISIT) ? MNV_MIN_BLOCKS_PLAIN : MNV_MIN_BLOCKS)
render_voxels_kernel(const RenderParams p) {
    // [path_levels][kThreads] node path | [TERMS][kThreads] SH basis | [words][kThreads] ray state
    extern __shared__ int32_t s_dyn[];
    // warp v of the frame = sub-tile (v % 4) of 16x8-pixel tile (v / 4): the same pixel -> warp map for every kWarps
    const int lane = threadIdx.x & 31;
    const int vwarp = (int) blockIdx.x * kWarps + (int) (threadIdx.x >> 5);
    const int sub = vwarp % kWarpsPerTile;
    const int bt_raw = vwarp / kWarpsPerTile;
    if (bt_raw >= p.n_tiles) return;
    const int bt = p.tg.tile_order ? p.tg.tile_order[bt_raw] : bt_raw;
    const int bty = bt / p.tiles_x, btx = bt - bty * p.tiles_x;
    if (p.tg.tile_mod > 1) {
        const int mt = ((bty * kTileH) / p.tg.tile_h) * p.mtiles_x + (btx * kTileW) / p.tg.tile_w;
        if (mt % p.tg.tile_mod != p.tg.tile_rem) return;
    }
    const int x = btx * kTileW + (sub & 1) * 8 + (lane & 7);
    const int y = bty * kTileH + (sub >> 1) * 4 + (lane >> 3);
    if (x >= p.cam.width || y >= p.cam.height) return;
    render_pixel<TERMS, TRACK, LOGV, VISIT, ANCHOR>(
            p, x, y, s_dyn + threadIdx.x,
            reinterpret_cast<float *>(s_dyn + p.path_levels * kThreads) + threadIdx.x,
            reinterpret_cast<float *>(s_dyn + (p.path_levels + TERMS) * kThreads) + threadIdx.x);
}

// query_single_from_root for arbitrary points (include/cuda/rt_core.cuh:117-159).
__global__ void query_points_kernel(TreeView tree, const float *__restrict__ xyz, int64_t n,
                                    int32_t *__restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float hi = f_from_bits(0x3F7FFFEFu);
    const float px = fmaxf(fminf(xyz[3 * i], hi), 0.f);
    const float py = fmaxf(fminf(xyz[3 * i + 1], hi), 0.f);
    const float pz = fmaxf(fminf(xyz[3 * i + 2], hi), 0.f);
    const uint32_t qx = __float_as_uint(__fmaf_rd(px, 8388608.f, 8388608.f));
    const uint32_t qy = __float_as_uint(__fmaf_rd(py, 8388608.f, 8388608.f));
    const uint32_t qz = __float_as_uint(__fmaf_rd(pz, 8388608.f, 8388608.f));
    int32_t node = 0;
    int lvl = 0, cidx;
    for (;;) {
        const int bit = 22 - lvl;
        cidx = (((qx >> bit) & 1u) << 2) | (((qy >> bit) & 1u) << 1) | ((qz >> bit) & 1u);
        const uint32_t cw = __ldg(tree.cell + ((int64_t) node * 8 + cidx));
        if ((cw & kLeafBit) || lvl >= kMaxLevel) break;
        node = (int32_t) cw;
        ++lvl;
    }
    out[3 * i] = node;
    out[3 * i + 1] = cidx;
    out[3 * i + 2] = lvl + 1;
}

template <int TERMS, bool ANCHOR>
int dispatch(const RenderParams &p, bool track, bool logv, bool visit, dim3 grid, size_t smem,
             cudaStream_t stream) {
#define MNV_LAUNCH(T, L, V) \
    render_voxels_kernel<TERMS, T, L, V, ANCHOR><<<grid, kThreads, smem, stream>>>(p)
    if (logv) {
        if (track) MNV_LAUNCH(true, true, false);
        else MNV_LAUNCH(false, true, false);
    } else if (visit) {
        if (track) MNV_LAUNCH(true, false, true);
        else MNV_LAUNCH(false, false, true);
    } else {
        if (track) MNV_LAUNCH(true, false, false);
        else MNV_LAUNCH(false, false, false);
    }
#undef MNV_LAUNCH
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

}  // namespace

int launch_render_voxels(DeviceTree &tree, const mnv_camera &cam,
                         const mnv_render_options &opt, const RenderTargets &tg,
                         cudaStream_t stream) {
    if (cam.width <= 0 || cam.height <= 0) {
        set_error("camera size %dx%d", cam.width, cam.height);
        return MNV_ERR_INVALID;
    }
    if (tg.partial_n > 0) {
        if (tg.partial_n > 8 || tg.partial_block <= 0 || tg.partial_slot < 0 || tg.to_split || tg.visit_hash ||
            (int64_t) tg.partial_n * tg.partial_block < (int64_t) cam.width * cam.height || !tg.offscreen) {
            set_error("bad partial-output arguments");
            return MNV_ERR_INVALID;
        }
    } else if ((tg.image_linear == nullptr) == (tg.image_surf == 0)) {
        set_error("exactly one of image_linear / image surface must be given");
        return MNV_ERR_INVALID;
    }
    if ((tg.to_split == nullptr) != (tg.to_sample == nullptr)) {
        set_error("to_split and to_sample must be given together");
        return MNV_ERR_INVALID;
    }
    if (tg.track_visit && !tg.visited) {
        set_error("track_visit needs a visited buffer");
        return MNV_ERR_INVALID;
    }
    refresh_max_leaf_depth(tree);
    {
        const int rc = ensure_anchor(tree, stream);
        if (rc != MNV_OK) return rc;
    }
    RenderParams p;
    p.tree = make_view(tree);
    p.cam = cam;
    p.opt = opt;
    p.tg = tg;
    p.tiles_x = (cam.width + kTileW - 1) / kTileW;
    const int tiles_y = (cam.height + kTileH - 1) / kTileH;
    // a leaf at depth d is found in a node of level d-1; refinement may deepen the tree
    p.max_level = std::min(kMaxLevel, std::max(tree.max_leaf_depth, 1) - 1);
    const bool anchored = tree.anchor_level > 0 && tree.anchor != nullptr;
    p.path_levels = anchored ? 0 : p.max_level + 1;
    p.split_limit = std::min(opt.max_depth, 23);
    p.samp_limit = opt.max_sample_count <= 0 ? 0u : (opt.max_sample_count >= 32768 ? 0x80000000u : (uint32_t) opt.max_sample_count << 16);
    p.anchor = tree.anchor;
    p.anchor_level = tree.anchor_level;
    p.anchor_dim = 1u << tree.anchor_level;
    p.anchor_scale = (float) p.anchor_dim;
    p.anchor_bias = (0x4B000000u * p.anchor_dim + 0x4B000000u) * p.anchor_dim + 0x4B000000u;
    if (p.tg.tile_mod > 1) {
        if (p.tg.tile_w < kTileW || p.tg.tile_h < kTileH || p.tg.tile_w % kTileW ||
            p.tg.tile_h % kTileH) {
            set_error("tile size must be a multiple of %dx%d", kTileW, kTileH);
            return MNV_ERR_INVALID;
        }
        p.mtiles_x = (cam.width + p.tg.tile_w - 1) / p.tg.tile_w;
    } else {
        p.tg.tile_mod = 1;
        p.mtiles_x = 1;
    }
    const bool track = tg.to_split != nullptr;
    const bool logv = tg.visit_hash || tg.visit_count || tg.shaded_count || tg.visit_log ||
                      tg.frame_stats;
    const bool visit = tg.track_visit;
    if (logv && visit) {
        set_error("visit logging and track_visit cannot be combined");
        return MNV_ERR_INVALID;
    }
    if (!p.tg.tile_order && tree.tile_order_dev && tree.tile_order_n == p.tiles_x * tiles_y)
        p.tg.tile_order = tree.tile_order_dev;
    const int terms = tree.format == MNV_FORMAT_SH ? tree.basis_dim : 0;
    p.n_tiles = p.tiles_x * tiles_y;
    const dim3 grid((unsigned) ((p.n_tiles * kWarpsPerTile + kWarps - 1) / kWarps));
    const size_t smem = (size_t) (p.path_levels +
                                  (MNV_SMEM_STATE ? terms + (track ? kRsWordsTrack : kRsWordsBase) : 0)) *
                        kThreads * sizeof(int32_t);
    int rc;
#define MNV_TERMS(N) rc = anchored ? dispatch<N, true>(p, track, logv, visit, grid, smem, stream) \
                                   : dispatch<N, false>(p, track, logv, visit, grid, smem, stream)
    switch (terms) {
        case 0: MNV_TERMS(0); break;
        case 1: MNV_TERMS(1); break;
        case 4: MNV_TERMS(4); break;
        case 9: MNV_TERMS(9); break;
        case 16: MNV_TERMS(16); break;
        case 25: MNV_TERMS(25); break;
        default:
            set_error("unsupported SH basis_dim %d", terms);
            return MNV_ERR_INVALID;
    }
#undef MNV_TERMS
    // the reference marks every node of the root path at each query; the anchored march marks the node holding
    // the leaf and the ancestors follow here (same visited set)
    if (rc == MNV_OK && visit && anchored) rc = launch_propagate_visited(tree, tg.visited, stream);
    return rc;
}

int launch_query_points(const DeviceTree &tree, const float *xyz, int64_t n, int32_t *out,
                        cudaStream_t stream) {
    if (n <= 0) return MNV_OK;
    const int th = 256;
    query_points_kernel<<<(unsigned) ((n + th - 1) / th), th, 0, stream>>>(make_view(tree), xyz, n,
                                                                           out);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

}  // namespace mnv
