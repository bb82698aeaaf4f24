// Guided ray sampling.
//
//   A7  get_samples_from_voxels_kernel / device::get_samples_trace_ray
//       (src/cuda/renderer_kernel.cu:329-363, include/cuda/rt_core.cuh:418-576):
//       the same march as the octree render, but every shaded leaf emits one MLP
//       input row (sample position at the leaf ENTRY point, optional view direction
//       and appearance index), its z value and its sub-module (cluster) id.
//   A8  render_nerf_results_kernel / device::composite_nerf_results
//       (renderer_kernel.cu:294-327, rt_core.cuh:334-416): per-ray alpha compositing
//       of the MLP outputs over a CSR range.
//
// The reference writes samples into dense [P][128][sd] f32 and [P*128][D+1] f32
// buffers (30.8 GB at 1080p for the result buffer alone, cuda_renderer.cpp:478-493)
// and compacts them afterwards with boolean masks.  Here the march runs twice — a
// count pass, an inclusive scan of the per-ray counts (== torch::cumsum,
// cuda_renderer.cpp:116), and an emit pass that writes rows straight into compact
// CSR storage — so memory is exactly V rows.  Row order and values are those of the
// reference's compacted `valid_samples`.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "mnv_internal.cuh"
#include "mnv_march.cuh"

namespace mnv {
namespace {

constexpr int kGThreads = 128;  // 4 warps, 16x8-pixel CTA tile, 8x4 pixels per warp

struct GuidedParams {
    TreeView tree;
    mnv_camera cam;
    mnv_render_options opt;
    cudaSurfaceObject_t depth_surf;
    bool offscreen;
    int tiles_x, max_level, path_levels;
    // outputs
    int32_t *num_samples;        // [P] (count pass)
    const int64_t *offsets;      // [P] inclusive scan (emit pass)
    float *z_vals;               // [V]
    float *rows;                 // [V][row_stride]  = x,y,z,(dir),(appearance)
    int16_t *cluster;            // [V]
    int row_stride;
    float *to_split, *to_sample; // [P][3] or null
    int32_t *visited;
    // cluster rule (rt_core.cuh:541-549)
    int grid0, grid1;
    float min1, min2, range1, range2;
    // sub-modules sharded across GPUs: this rank marches one cell's segment of every ray
    float4 *seg_probe;        // [P] out of the probe pass: (T_segment, samples, z of the first sample, -)
    const float4 *seg_table;  // [seg_n][P] every rank's probe record (null: unsharded)
    int seg_n, seg_slot;
    bool has_cell;
    float cell_box[6];
};

constexpr float kNoSegment = 3.0e38f;

// What the march of the unsharded frame would carry into this rank's cell, and where the segment after
// this one starts — from every rank's probe record of the ray.  Cells are disjoint convex boxes, so the
// segments of a ray do not interleave: ordering them by the z of their first sample is the march order.
// A segment emits only while the transmittance is above stop_thresh and the ray has fewer than
// max_guided_samples samples (the two rules of the emission loop below, rt_core.cuh:335-352).
struct SegmentContext {
    float T_in;    // transmittance at the cell's entry
    int count_in;  // samples the ray already has
    float z_next;  // first sample of the next emitting segment, kNoSegment if this one is the last
};
__device__ __forceinline__ SegmentContext segment_context(const float4 *__restrict__ table, int n, int slot,
                                                          size_t P, size_t idx, float stop_thresh,
                                                          int max_samples) {
    SegmentContext c;
    c.T_in = 1.f;
    c.count_in = 0;
    c.z_next = kNoSegment;
    const float4 mine = table[(size_t) slot * P + idx];
    const float z_mine = mine.y > 0.f ? mine.z : kNoSegment;
    float T_after = 1.f;   // running transmittance / count over the segments behind this one, in z order
    int count_after = 0;
    float z_prev = -1.f;
    int prev_slot = -1;
    // walk the (at most 8) segments in z order without sorting: repeatedly take the smallest key
    // greater than the previous one
    for (int k = 0; k < n; ++k) {
        float zk = kNoSegment;
        int ck = -1;
        float4 rk = make_float4(1.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < n; ++s) {
            const float4 r = table[(size_t) s * P + idx];
            if (!(r.y > 0.f)) continue;
            const bool after_prev = r.z > z_prev || (r.z == z_prev && s > prev_slot);
            const bool before_best = r.z < zk || (r.z == zk && s < ck);
            if (after_prev && (ck < 0 || before_best)) {
                zk = r.z;
                ck = s;
                rk = r;
            }
        }
        if (ck < 0) break;
        z_prev = zk;
        prev_slot = ck;
        const bool before_me = zk < z_mine || (zk == z_mine && ck < slot);
        if (before_me) {
            c.T_in = __fmul_rn(c.T_in, rk.x);
            c.count_in += (int) rk.y;
            T_after = c.T_in;
            count_after = c.count_in;
        } else if (ck == slot) {
            T_after = __fmul_rn(c.T_in, rk.x);
            count_after = c.count_in + (int) rk.y;
        } else {
            // a later segment: does it emit anything?
            if (!(T_after < stop_thresh) && count_after < max_samples) {
                c.z_next = zk;
                break;
            }
            T_after = __fmul_rn(T_after, rk.x);
            count_after += (int) rk.y;
        }
    }
    return c;
}

template <bool EMIT, bool TRACK, bool VISIT>
__global__ void __launch_bounds__(kGThreads, 8) guided_samples_kernel(const GuidedParams p) {
    extern __shared__ int32_t s_dyn[];  // [path_levels][kGThreads]
    const int bt = blockIdx.x;
    const int bty = bt / p.tiles_x, btx = bt - bty * p.tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = btx * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = bty * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= p.cam.width || y >= p.cam.height) return;
    const int idx = y * p.cam.width + x;
    const mnv_render_options &opt = p.opt;
    int32_t *path = s_dyn + threadIdx.x;

    float tmax_bg = 1e9f;
    if (!p.offscreen) tmax_bg = surf2Dread<float>(p.depth_surf, x * 4, y, cudaBoundaryModeZero);
    Ray r;
    setup_ray(p.tree, p.cam, opt, x, y, tmax_bg, r, p.has_cell ? p.cell_box : nullptr);

    float split_prio = (float) (opt.max_depth + 1), samp_prio = (float) (opt.max_sample_count + 1);
    int32_t split_id = -1, samp_id = -1;
    float max_weight = -1.f, max_sample_weight = -1.f;
    int count = 0;
    const int64_t base = EMIT ? (idx == 0 ? 0 : p.offsets[idx - 1]) : 0;
    const float *cm = p.cam.c2w;
    float T_init = 1.f, z_first = kNoSegment;
    int count_init = 0;
    if (p.seg_table) {
        const SegmentContext sc = segment_context(p.seg_table, p.seg_n, p.seg_slot,
                                                  (size_t) p.cam.width * p.cam.height, (size_t) idx,
                                                  opt.stop_thresh, opt.max_guided_samples);
        T_init = sc.T_in;
        count_init = sc.count_in;
    }
    float T = T_init;

    if (r.hit && !(T_init < opt.stop_thresh) && count_init < opt.max_guided_samples) {
        float t = r.tmin;
        MarchState ms;
        while (t < r.tmax) {
            const Leaf lf = march_step<VISIT, /*FUSED_POS=*/false>(p.tree.cell, p.max_level, r, t, opt.step_size, ms, path,
                                              kGThreads, p.visited);
            const float sigma = leaf_sigma(lf.cw);
            const int scount = leaf_sample_count(lf.cw);
            if ((int32_t) lf.cw < 0 && sigma > opt.sigma_thresh) {
                const float att = ref_expf(__fmul_rn(__fmul_rn(r.delta_scale, -lf.delta_t), sigma));
                const float weight = __fmul_rn(T, __fadd_rn(1.f, -att));
                if (TRACK) {
                    if (weight > max_weight && lf.depth < opt.max_depth) {
                        split_id = (int32_t) (lf.node * 8u + lf.cidx);
                        split_prio = (float) lf.depth;
                        max_weight = weight;
                    }
                    if (weight > max_sample_weight && scount < opt.max_sample_count) {
                        samp_id = (int32_t) (lf.node * 8u + lf.cidx);
                        samp_prio = (float) scount;
                        max_sample_weight = weight;
                    }
                }
                if (count_init + count < opt.max_guided_samples) {
                    if (!EMIT && p.seg_probe && count == 0) {
                        const float z0 = __fdiv_rn(__fmul_rn(t, r.d0), p.tree.scale[0]);
                        const float z1 = __fdiv_rn(__fmul_rn(t, r.d1), p.tree.scale[1]);
                        const float z2 = __fdiv_rn(__fmul_rn(t, r.d2), p.tree.scale[2]);
                        z_first = ref_norm3(z0, z1, z2);
                    }
                    if (EMIT) {
                        // true_z = t*dir/scale ; z = |true_z| ; sample = true_cen + true_dir*z
                        const float z0 = __fdiv_rn(__fmul_rn(t, r.d0), p.tree.scale[0]);
                        const float z1 = __fdiv_rn(__fmul_rn(t, r.d1), p.tree.scale[1]);
                        const float z2 = __fdiv_rn(__fmul_rn(t, r.d2), p.tree.scale[2]);
                        const float z = ref_norm3(z0, z1, z2);
                        const float sx = __fmaf_rn(r.w0, z, cm[9]);
                        const float sy = __fmaf_rn(r.w1, z, cm[10]);
                        const float sz = __fmaf_rn(r.w2, z, cm[11]);
                        const int64_t row = base + count;
                        p.z_vals[row] = z;
                        float *o = p.rows + row * p.row_stride;
                        o[0] = sx;
                        o[1] = sy;
                        o[2] = sz;
                        int c = 3;
                        if (opt.need_viewdir) {
                            o[3] = r.v0;
                            o[4] = r.v1;
                            o[5] = r.v2;
                            c = 6;
                        }
                        if (opt.appearance_embedding != -1) o[c] = (float) opt.appearance_embedding;
                        const float g0 = (float) p.grid0, g1 = (float) p.grid1;
                        const int a = (int) fmaxf(
                                fminf(__fmul_rn(__fdiv_rn(__fadd_rn(sy, -p.min1), p.range1), g0),
                                      __fadd_rn(g0, -1.f)),
                                0.f);
                        const int b = (int) fmaxf(
                                fminf(__fmul_rn(__fdiv_rn(__fadd_rn(sz, -p.min2), p.range2), g1),
                                      __fadd_rn(g1, -1.f)),
                                0.f);
                        p.cluster[row] = (int16_t) (a * p.grid1 + b);
                    }
                    ++count;
                }
                T = __fmul_rn(T, att);
                if (T < opt.stop_thresh) break;
            } else if (TRACK) {
                if (max_weight == -1.f && lf.depth < opt.max_depth) {
                    split_id = (int32_t) (lf.node * 8u + lf.cidx);
                    split_prio = (float) lf.depth;
                }
                if (max_sample_weight == -1.f && scount < opt.max_sample_count) {
                    samp_id = (int32_t) (lf.node * 8u + lf.cidx);
                    samp_prio = (float) scount;
                }
            }
            t = __fadd_rn(t, lf.delta_t);
        }
    }
    if (!EMIT) {
        if (p.seg_probe) p.seg_probe[idx] = make_float4(T, (float) count, z_first, 0.f);
        else p.num_samples[idx] = count;
    }
    if (TRACK) {
        float *ts = p.to_split + (size_t) idx * 3;
        ts[0] = split_prio;
        ts[1] = split_id < 0 ? -1.f : tracker_encode_chunk(split_id >> 3);
        ts[2] = split_id < 0 ? -1.f : (float) (split_id & 7);
        float *tp = p.to_sample + (size_t) idx * 3;
        tp[0] = samp_prio;
        tp[1] = samp_id < 0 ? -1.f : tracker_encode_chunk(samp_id >> 3);
        tp[2] = samp_id < 0 ? -1.f : (float) (samp_id & 7);
    }
}

// ---------------------------------------------------------------- compositor (A8)
struct CompositeParams {
    mnv_camera cam;
    mnv_render_options opt;
    int basis_dim;  // -1: RGBA
    uint8_t *image_linear;
    cudaSurfaceObject_t image_surf;
    const float *values;  // [V][value_stride]
    int value_stride, sigma_col;
    const float *z_vals;
    const int64_t *offsets;
    bool offscreen;
    // segment mode (sub-modules sharded across GPUs): this rank's samples are one segment of every ray
    const float4 *seg_table;    // [n_seg][P] every rank's probe record
    int n_seg, slot;
    float4 *const *partial_dst;  // owner o's [n_seg][block] float4 buffer
    int partial_block;
};

template <int TERMS>
__device__ __forceinline__ float sh_channel_f32(const float (&B)[TERMS > 0 ? TERMS : 1],
                                                const float *__restrict__ c) {
    float tmp = __fmul_rn(B[0], c[0]);
    if constexpr (TERMS >= 25) {
        float s = __fmul_rn(B[17], c[17]);
        s = __fmaf_rn(B[16], c[16], s);
#pragma unroll
        for (int k = 18; k <= 24; ++k) s = __fmaf_rn(B[k], c[k], s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 16) {
        float s = __fmul_rn(B[10], c[10]);
        s = __fmaf_rn(B[9], c[9], s);
#pragma unroll
        for (int k = 11; k <= 15; ++k) s = __fmaf_rn(B[k], c[k], s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 9) {
        float s = __fmul_rn(B[5], c[5]);
        s = __fmaf_rn(B[4], c[4], s);
#pragma unroll
        for (int k = 6; k <= 8; ++k) s = __fmaf_rn(B[k], c[k], s);
        tmp = __fadd_rn(tmp, s);
    }
    if constexpr (TERMS >= 4) {
        float s = __fmul_rn(B[2], c[2]);
        s = __fmaf_rn(B[1], c[1], s);
        s = __fmaf_rn(B[3], c[3], s);
        tmp = __fadd_rn(tmp, s);
    }
    return tmp;
}

template <int TERMS, bool SEG>
__global__ void __launch_bounds__(256) composite_nerf_kernel(const CompositeParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = p.cam.width;
    if (idx >= W * p.cam.height) return;
    const int x = idx % W, y = idx / W;
    const mnv_render_options &opt = p.opt;
    uint32_t rgbx_init = 0;
    if (!SEG && !p.offscreen) rgbx_init = surf2Dread<uint32_t>(p.image_surf, x * 4, y, cudaBoundaryModeZero);

    float out0 = 0.f, out1 = 0.f, out2 = 0.f;
    const int64_t start = idx == 0 ? 0 : p.offsets[idx - 1], end = p.offsets[idx];
    // SEG: where the next segment of this ray begins (cells are disjoint convex boxes: segments do not
    // interleave), so that this segment's last sample gets the same delta as in the unsharded frame
    float z_next = kNoSegment, t_end = 1.f;
    if (SEG && start != end)
        z_next = segment_context(p.seg_table, p.n_seg, p.slot, (size_t) W * p.cam.height, (size_t) idx,
                                 opt.stop_thresh, opt.max_guided_samples).z_next;
    if (start != end) {
        // view direction: screen2worlddir + rodrigues (renderer_kernel.cu:311-314)
        const float *m = p.cam.c2w;
        const float vx = __fdiv_rn(__fadd_rn(__fadd_rn((float) x, 0.5f), -p.cam.cx), p.cam.fx);
        const float vy = __fdiv_rn(-__fadd_rn(__fadd_rn((float) y, 0.5f), -p.cam.cy), p.cam.fy);
        float v0 = __fadd_rn(__fmaf_rn(vx, m[0], __fmul_rn(vy, m[3])), -m[6]);
        float v1 = __fadd_rn(__fmaf_rn(vx, m[1], __fmul_rn(vy, m[4])), -m[7]);
        float v2 = __fadd_rn(__fmaf_rn(vx, m[2], __fmul_rn(vy, m[5])), -m[8]);
        const float inv = __frcp_rn(ref_norm3(v0, v1, v2));
        v0 = __fmul_rn(v0, inv);
        v1 = __fmul_rn(v1, inv);
        v2 = __fmul_rn(v2, inv);
        ref_rodrigues(opt.rot_dirs, v0, v1, v2);
        float B[TERMS > 0 ? TERMS : 1];
        if (TERMS > 0) {
            ref_sh_basis<(TERMS > 0 ? TERMS : 1)>(v0, v1, v2, B);
#pragma unroll
            for (int k = 0; k < TERMS; ++k)
                if (k < opt.basis_minmax[0] || k > opt.basis_minmax[1]) B[k] = 0.f;
        }
        float ti = 1.f, wc = 0.f;  // the reference leaves weight_component uninitialised for 1-sample rays
        for (int64_t i = start; i < end; ++i) {
            const float *sv = p.values + i * p.value_stride;
            float weight;
            if (i < end - 1 || (SEG && z_next < 1.0e38f)) {
                const float zn = i < end - 1 ? p.z_vals[i + 1] : z_next;
                const float delta = __fadd_rn(zn, -p.z_vals[i]);
                wc = ref_expf(__fmul_rn(delta, -sv[p.sigma_col]));
                weight = __fmul_rn(ti, __fadd_rn(1.f, -wc));
            } else {
                weight = ti;  // the last sample of the ray takes what is left
                if (SEG) wc = 0.f;
            }
            if (opt.render_depth) {
                out0 = __fmaf_rn(ti, weight, out0);
            } else if (TERMS > 0) {
                out0 = __fadd_rn(out0, ref_weighted_sigmoid(weight, sh_channel_f32<TERMS>(B, sv)));
                out1 = __fadd_rn(out1, ref_weighted_sigmoid(weight, sh_channel_f32<TERMS>(B, sv + TERMS)));
                out2 = __fadd_rn(out2, ref_weighted_sigmoid(weight, sh_channel_f32<TERMS>(B, sv + 2 * TERMS)));
            } else {
                out0 = __fmaf_rn(weight, sv[0], out0);
                out1 = __fmaf_rn(weight, sv[1], out1);
                out2 = __fmaf_rn(weight, sv[2], out2);
            }
            ti = __fmul_rn(ti, wc);
        }
        t_end = ti;
        if (!SEG && opt.render_depth) out0 = out1 = out2 = fminf(__fmul_rn(out0, 0.3f), 1.0f);
    }
    if constexpr (SEG) {
        // premultiplied colour + alpha of the segment, straight into the pixel owner's memory
        const int owner = idx / p.partial_block;
        p.partial_dst[owner][(size_t) p.slot * p.partial_block + (idx - owner * p.partial_block)] =
            make_float4(out0, out1, out2, __fadd_rn(1.f, -t_end));
        return;
    }
    // out[3] = 1 (renderer_kernel.cu:315-316): composite_and_write adds nothing
    const float nalpha = 0.f;
    if (p.offscreen) {
        const float remain = __fmul_rn(nalpha, opt.background_brightness);
        out0 = __fadd_rn(out0, remain);
        out1 = __fadd_rn(out1, remain);
        out2 = __fadd_rn(out2, remain);
    } else {
        out0 = __fmaf_rn(__fdiv_rn((float) (rgbx_init & 0xffu), 255.f), nalpha, out0);
        out1 = __fadd_rn(__fmul_rn(__fdiv_rn((float) ((rgbx_init >> 8) & 0xffu), 255.f), nalpha), out1);
        out2 = __fadd_rn(__fmul_rn(__fdiv_rn((float) ((rgbx_init >> 16) & 0xffu), 255.f), nalpha), out2);
    }
    const uint32_t rgba = ref_to_u8(out0) | (ref_to_u8(out1) << 8) | (ref_to_u8(out2) << 16) | 0xff000000u;
    if (p.image_linear)
        reinterpret_cast<uint32_t *>(p.image_linear)[idx] = rgba;
    else
        surf2Dwrite(rgba, p.image_surf, x * 4, y, cudaBoundaryModeZero);
}

struct ToI64 {
    __host__ __device__ int64_t operator()(const int32_t &v) const { return (int64_t) v; }
};

}  // namespace

int launch_guided_samples(DeviceTree &tree, const mnv_camera &cam, const mnv_render_options &opt,
                          const GuidedIO &io, cudaStream_t stream) {
    const int W = cam.width, H = cam.height;
    if (W <= 0 || H <= 0) return MNV_ERR_INVALID;
    const int64_t P = (int64_t) W * H;
    if ((io.to_split == nullptr) != (io.to_sample == nullptr)) {
        set_error("to_split and to_sample must be given together");
        return MNV_ERR_INVALID;
    }
    if (io.seg_table && (io.seg_n < 1 || io.seg_n > 8 || io.seg_slot < 0 || io.seg_slot >= io.seg_n)) {
        set_error("guided samples: bad segment description (%d of %d)", io.seg_slot, io.seg_n);
        return MNV_ERR_INVALID;
    }
    if (io.seg_probe) {
        // probe pass of the sharded frame: the march only
        GuidedParams q{};
        q.tree = make_view(tree);
        q.cam = cam;
        q.opt = opt;
        q.depth_surf = io.depth_surf;
        q.offscreen = io.offscreen;
        q.tiles_x = (W + 15) / 16;
        q.max_level = std::min(22, std::max(tree.max_leaf_depth, 1) - 1);
        q.path_levels = q.max_level + 1;
        q.seg_probe = io.seg_probe;
        q.has_cell = io.has_cell;
        for (int a = 0; a < 6; ++a) q.cell_box[a] = io.cell_box[a];
        const dim3 g((unsigned) (q.tiles_x * ((H + 7) / 8)));
        guided_samples_kernel<false, false, false>
            <<<g, kGThreads, (size_t) q.path_levels * kGThreads * sizeof(int32_t), stream>>>(q);
        MNV_CUDA(cudaGetLastError());
        return MNV_OK;
    }
    if (io.track_visit && !io.visited) {
        set_error("track_visit needs a visited buffer");
        return MNV_ERR_INVALID;
    }
    const int in_dim = 3 + (opt.need_viewdir ? 3 : 0) + (opt.appearance_embedding != -1 ? 1 : 0);
    if (io.row_stride < in_dim) {
        set_error("row_stride %d < %d input columns", io.row_stride, in_dim);
        return MNV_ERR_INVALID;
    }
    // scratch: per-ray counts + CUB temp storage, owned by the tree
    if (tree.count_cap < P) {
        cudaFree(tree.count_dev);
        tree.count_dev = nullptr;
        tree.count_cap = 0;
        MNV_CUDA(cudaMalloc(&tree.count_dev, P * sizeof(int32_t)));
        tree.count_cap = P;
    }
    auto it = thrust::make_transform_iterator((const int32_t *) tree.count_dev, ToI64());
    size_t temp_bytes = 0;
    MNV_CUDA(cub::DeviceScan::InclusiveSum(nullptr, temp_bytes, it, io.offsets, (int) P, stream));
    if (tree.scan_tmp_bytes < temp_bytes) {
        cudaFree(tree.scan_tmp);
        tree.scan_tmp = nullptr;
        tree.scan_tmp_bytes = 0;
        MNV_CUDA(cudaMalloc(&tree.scan_tmp, temp_bytes));
        tree.scan_tmp_bytes = temp_bytes;
    }

    GuidedParams p;
    p.tree = make_view(tree);
    p.cam = cam;
    p.opt = opt;
    p.depth_surf = io.depth_surf;
    p.offscreen = io.offscreen;
    p.tiles_x = (W + 15) / 16;
    p.max_level = std::min(22, std::max(tree.max_leaf_depth, 1) - 1);
    p.path_levels = p.max_level + 1;
    p.num_samples = tree.count_dev;
    p.offsets = io.offsets;
    p.z_vals = io.z_vals;
    p.rows = io.rows;
    p.cluster = io.cluster;
    p.row_stride = io.row_stride;
    p.to_split = io.to_split;
    p.to_sample = io.to_sample;
    p.visited = io.visited;
    p.grid0 = io.grid_dim[0];
    p.grid1 = io.grid_dim[1];
    p.min1 = io.min_position[1];
    p.min2 = io.min_position[2];
    p.range1 = io.range[1];
    p.range2 = io.range[2];
    p.seg_probe = nullptr;
    p.seg_table = io.seg_table;
    p.seg_n = io.seg_n;
    p.seg_slot = io.seg_slot;
    p.has_cell = io.has_cell;
    for (int a = 0; a < 6; ++a) p.cell_box[a] = io.cell_box[a];
    const dim3 grid((unsigned) (p.tiles_x * ((H + 7) / 8)));
    const size_t smem = (size_t) p.path_levels * kGThreads * sizeof(int32_t);

    // pass 1: count (no candidates, no visit marks: the emit pass does both)
    guided_samples_kernel<false, false, false><<<grid, kGThreads, smem, stream>>>(p);
    MNV_CUDA(cudaGetLastError());
    MNV_CUDA(cub::DeviceScan::InclusiveSum(tree.scan_tmp, temp_bytes, it, io.offsets, (int) P, stream));
    int64_t total = 0;
    MNV_CUDA(cudaMemcpyAsync(&total, io.offsets + (P - 1), sizeof(int64_t), cudaMemcpyDeviceToHost,
                             stream));
    MNV_CUDA(cudaStreamSynchronize(stream));  // the reference syncs here too (valid_samples.size(0))
    if (io.total_rows) *io.total_rows = total;
    if (total > io.capacity_rows) {
        set_error("guided samples: %lld rows needed, capacity %lld", (long long) total,
                  (long long) io.capacity_rows);
        return MNV_ERR_FULL;
    }
    // pass 2: emit
    const bool track = io.to_split != nullptr;
    if (io.track_visit) {
        if (track) guided_samples_kernel<true, true, true><<<grid, kGThreads, smem, stream>>>(p);
        else guided_samples_kernel<true, false, true><<<grid, kGThreads, smem, stream>>>(p);
    } else {
        if (track) guided_samples_kernel<true, true, false><<<grid, kGThreads, smem, stream>>>(p);
        else guided_samples_kernel<true, false, false><<<grid, kGThreads, smem, stream>>>(p);
    }
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

int launch_composite_nerf(const DeviceTree &tree, const mnv_camera &cam,
                          const mnv_render_options &opt, uint8_t *image_linear,
                          cudaSurfaceObject_t image_surf, const float *values, int value_stride,
                          int sigma_col, const float *z_vals, const int64_t *offsets, bool offscreen,
                          cudaStream_t stream, const NerfSegment *seg) {
    if (!seg && (image_linear == nullptr) == (image_surf == 0)) {
        set_error("exactly one of image_linear / image surface must be given");
        return MNV_ERR_INVALID;
    }
    if (seg && opt.render_depth) {
        set_error("composite_nerf: render_depth is not available in segment mode");
        return MNV_ERR_INVALID;
    }
    if (seg && (seg->n_seg < 1 || seg->n_seg > 8 || seg->slot < 0 || seg->slot >= seg->n_seg ||
                !seg->seg_table || !seg->partial_dst || seg->partial_block <= 0)) {
        set_error("composite_nerf: bad segment description");
        return MNV_ERR_INVALID;
    }
    CompositeParams p;
    p.seg_table = seg ? seg->seg_table : nullptr;
    p.n_seg = seg ? seg->n_seg : 0;
    p.slot = seg ? seg->slot : 0;
    p.partial_dst = seg ? seg->partial_dst : nullptr;
    p.partial_block = seg ? seg->partial_block : 1;
    p.cam = cam;
    p.opt = opt;
    p.basis_dim = tree.basis_dim;
    p.image_linear = image_linear;
    p.image_surf = image_surf;
    p.values = values;
    p.value_stride = value_stride;
    // reference quirk (rt_core.cuh:365): sigma is read from column 3 whatever the format
    p.sigma_col = sigma_col < 0 ? 3 : sigma_col;
    p.z_vals = z_vals;
    p.offsets = offsets;
    p.offscreen = offscreen;
    const int P = cam.width * cam.height;
    const int th = 256, blocks = (P + th - 1) / th;
    const int terms = tree.format == MNV_FORMAT_SH ? tree.basis_dim : 0;
    switch (terms) {
        case 0:
            if (seg) composite_nerf_kernel<0, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<0, false><<<blocks, th, 0, stream>>>(p);
            break;
        case 1:
            if (seg) composite_nerf_kernel<1, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<1, false><<<blocks, th, 0, stream>>>(p);
            break;
        case 4:
            if (seg) composite_nerf_kernel<4, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<4, false><<<blocks, th, 0, stream>>>(p);
            break;
        case 9:
            if (seg) composite_nerf_kernel<9, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<9, false><<<blocks, th, 0, stream>>>(p);
            break;
        case 16:
            if (seg) composite_nerf_kernel<16, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<16, false><<<blocks, th, 0, stream>>>(p);
            break;
        case 25:
            if (seg) composite_nerf_kernel<25, true><<<blocks, th, 0, stream>>>(p);
            else composite_nerf_kernel<25, false><<<blocks, th, 0, stream>>>(p);
            break;
        default: set_error("unsupported basis_dim %d", terms); return MNV_ERR_INVALID;
    }
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

}  // namespace mnv
