// Ray set-up and one march step, shared by the guided-sampling kernels
// (mnv_guided.cu).  Same arithmetic and the same integer-cell descent as
// render_pixel in mnv_render.cu (see the comments there and mnv_math.cuh); kept as
// small force-inlined helpers so that other per-ray kernels reproduce the
// reference's leaf-visit sequence exactly.
#pragma once

#include <cuda_fp16.h>

#include "mnv_internal.cuh"
#include "mnv_math.cuh"

namespace mnv {

struct Ray {
    float c0, c1, c2;     // origin in tree space
    float d0, d1, d2;     // direction in tree space (scaled, unit length)
    float i0, i1, i2;     // 1 / (dir + 1e-9)
    float delta_scale;    // rt_core.cuh:102-115
    float tmin, tmax;
    float w0, w1, w2;     // unit direction in world space ("true_dir")
    float v0, v1, v2;     // view direction after rot_dirs
    bool hit;
};

// screen2worlddir + cen transform + rodrigues + _get_delta_scale + invdir + _dda_world
// (renderer_kernel.cu:30-61,272-283; rt_core.cuh:70-115,188-199).
__device__ __forceinline__ void setup_ray(const TreeView &tree, const mnv_camera &cam,
                                          const mnv_render_options &opt, int x, int y,
                                          float tmax_bg, Ray &r, const float *cell_box = nullptr) {
    const float *m = cam.c2w;
    const float vx = __fdiv_rn(__fadd_rn(__fadd_rn((float) x, 0.5f), -cam.cx), cam.fx);
    const float vy = __fdiv_rn(-__fadd_rn(__fadd_rn((float) y, 0.5f), -cam.cy), cam.fy);
    float d0 = __fadd_rn(__fmaf_rn(vx, m[0], __fmul_rn(vy, m[3])), -m[6]);
    float d1 = __fadd_rn(__fmaf_rn(vx, m[1], __fmul_rn(vy, m[4])), -m[7]);
    float d2 = __fadd_rn(__fmaf_rn(vx, m[2], __fmul_rn(vy, m[5])), -m[8]);
    {
        const float inv = __frcp_rn(ref_norm3(d0, d1, d2));
        d0 = __fmul_rn(d0, inv);
        d1 = __fmul_rn(d1, inv);
        d2 = __fmul_rn(d2, inv);
    }
    r.w0 = d0;
    r.w1 = d1;
    r.w2 = d2;
    r.c0 = __fmaf_rn(tree.scale[0], m[9], tree.offset[0]);
    r.c1 = __fmaf_rn(tree.scale[1], m[10], tree.offset[1]);
    r.c2 = __fmaf_rn(tree.scale[2], m[11], tree.offset[2]);
    r.v0 = d0;
    r.v1 = d1;
    r.v2 = d2;
    ref_rodrigues(opt.rot_dirs, r.v0, r.v1, r.v2);

    d0 = __fmul_rn(d0, tree.scale[0]);
    d1 = __fmul_rn(d1, tree.scale[1]);
    d2 = __fmul_rn(d2, tree.scale[2]);
    r.delta_scale = __frcp_rn(ref_norm3(d0, d1, d2));
    r.d0 = __fmul_rn(d0, r.delta_scale);
    r.d1 = __fmul_rn(d1, r.delta_scale);
    r.d2 = __fmul_rn(d2, r.delta_scale);
    tmax_bg = __fdiv_rn(tmax_bg, r.delta_scale);
    r.i0 = d2f(__drcp_rn(__dadd_rn((double) r.d0, 1e-9)));
    r.i1 = d2f(__drcp_rn(__dadd_rn((double) r.d1, 1e-9)));
    r.i2 = d2f(__drcp_rn(__dadd_rn((double) r.d2, 1e-9)));
    float tmin = 0.f, tmax = 1e4f;
    const float cc[3] = {r.c0, r.c1, r.c2};
    const float ii[3] = {r.i0, r.i1, r.i2};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double ci = (double) cc[i], inv = (double) ii[i];
        const float t1 = d2f(__dmul_rn(__dadd_rn(__dadd_rn((double) opt.render_bbox[i], 1e-6), -ci), inv));
        const float t2 =
                d2f(__dmul_rn(__dadd_rn(__dadd_rn((double) opt.render_bbox[i + 3], -1e-6), -ci), inv));
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    if (cell_box) clip_to_cell(cell_box, r.c0, r.c1, r.c2, r.i0, r.i1, r.i2, opt.step_size, tmin, tmax);
    tmax = fminf(tmax, tmax_bg);
    r.tmin = tmin;
    r.tmax = tmax;
    r.hit = !(tmax < 0.f || tmin > tmax);
}

struct MarchState {
    uint32_t pqx = 0x4B000000u, pqy = 0x4B000000u, pqz = 0x4B000000u;
    int pdepth = 1;
};

struct Leaf {
    uint32_t cw, node, cidx;
    int depth;
    float delta_t;
};

// One iteration of the reference's while (t < tmax) loop up to and including
// delta_t (rt_core.cuh:221-230): locate the leaf of pos = cen + t*dir and the
// distance to its exit.  s_path: this thread's column of the shared node path.
// FUSED_POS: pos = FFMA(t, dir, cen) as in the reference's render_voxels_kernel; false:
// pos = FADD(cen, FMUL(t, dir)), which is what its get_samples_from_voxels_kernel executes
// (t*dir is reused for the sample's z there, so nvcc does not contract the add).
template <bool VISIT, bool FUSED_POS>
__device__ __forceinline__ Leaf march_step(const uint32_t *__restrict__ cells, const int max_level,
                                           const Ray &r, const float t, const float step_size,
                                           MarchState &ms, int32_t *__restrict__ s_path,
                                           const int path_stride, int32_t *visited) {
    const float clamp_hi = f_from_bits(0x3F7FFFEFu);
    const float rx = FUSED_POS ? __fmaf_rn(t, r.d0, r.c0) : __fadd_rn(r.c0, __fmul_rn(t, r.d0));
    const float ry = FUSED_POS ? __fmaf_rn(t, r.d1, r.c1) : __fadd_rn(r.c1, __fmul_rn(t, r.d1));
    const float rz = FUSED_POS ? __fmaf_rn(t, r.d2, r.c2) : __fadd_rn(r.c2, __fmul_rn(t, r.d2));
    const float px = fminf(__saturatef(rx), clamp_hi);
    const float py = fminf(__saturatef(ry), clamp_hi);
    const float pz = fminf(__saturatef(rz), clamp_hi);
    const uint32_t qx = __float_as_uint(__fmaf_rd(px, 8388608.f, 8388608.f));
    const uint32_t qy = __float_as_uint(__fmaf_rd(py, 8388608.f, 8388608.f));
    const uint32_t qz = __float_as_uint(__fmaf_rd(pz, 8388608.f, 8388608.f));
    const uint32_t diff = (qx ^ ms.pqx) | (qy ^ ms.pqy) | (qz ^ ms.pqz);
    ms.pqx = qx;
    ms.pqy = qy;
    ms.pqz = qz;
    int lvl = min(__clz((int) diff) - 9, ms.pdepth - 1);
    uint32_t node = lvl > 0 ? (uint32_t) s_path[lvl * path_stride] : 0u;
    // level-lvl child bit of each axis at bit 31; three funnel shifts append them to node: node * 8 + child
    // (same slot arithmetic as render_pixel, mnv_render.cu)
    uint32_t sx = qx << (9 + lvl), sy = qy << (9 + lvl), sz = qz << (9 + lvl);
    Leaf lf;
    uint32_t slot;
    for (;;) {
        if (VISIT) {
            if (visited[node] == 0) visited[node] = 1;
        }
        slot = __funnelshift_l(sz, __funnelshift_l(sy, __funnelshift_l(sx, node, 1), 1), 1);
        lf.cw = __ldg(cells + slot);
        if ((int32_t) lf.cw < 0 || lvl >= max_level) break;
        node = lf.cw;
        ++lvl;
        sx <<= 1;
        sy <<= 1;
        sz <<= 1;
        s_path[lvl * path_stride] = (int32_t) node;
    }
    lf.node = node;
    lf.cidx = slot & 7u;
    lf.depth = lvl + 1;
    ms.pdepth = lf.depth;
    const float cube = __uint_as_float((uint32_t) (127 + lf.depth) << 23);
    const float icube = __uint_as_float((uint32_t) (127 - lf.depth) << 23);
    const float flx = __fadd_rn(__fmaf_rd(px, cube, 8388608.f), -8388608.f);
    const float fly = __fadd_rn(__fmaf_rd(py, cube, 8388608.f), -8388608.f);
    const float flz = __fadd_rn(__fmaf_rd(pz, cube, 8388608.f), -8388608.f);
    const float fx = __fmaf_rn(px, cube, -flx);
    const float fy = __fmaf_rn(py, cube, -fly);
    const float fz = __fmaf_rn(pz, cube, -flz);
    const float a1 = __fmul_rn(-fx, r.i0), a2 = __fadd_rn(a1, r.i0);
    const float b1 = __fmul_rn(-fy, r.i1), b2 = __fadd_rn(b1, r.i1);
    const float e1 = __fmul_rn(-fz, r.i2), e2 = __fadd_rn(e1, r.i2);
    const float tm = fminf(fminf(fminf(fmaxf(a1, a2), 1e4f), fmaxf(b1, b2)), fmaxf(e1, e2));
    lf.delta_t = __fadd_rn(__fmul_rn(tm, icube), step_size);
    return lf;
}

__device__ __forceinline__ float leaf_sigma(uint32_t cw) {
    return __half2float(__ushort_as_half((unsigned short) (cw & 0xffffu)));
}
__device__ __forceinline__ int leaf_sample_count(uint32_t cw) { return (int) ((cw >> 16) & 0x7fffu); }

}  // namespace mnv
