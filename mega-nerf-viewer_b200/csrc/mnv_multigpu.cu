// Sub-module split across the GPUs of one box (second mode).
//
// The reference has no multi-GPU code; this is the B200-side design for octrees (and sub-MLPs)
// that are partitioned by Mega-NeRF spatial cell: GPU g owns the cell g of the (y, z) grid — its
// subtree — and marches EVERY ray of the frame through that cell only (render_bbox clip of
// rt_core.cuh:192-199).  The segment's premultiplied colour and alpha leave the march kernel as
// one 16-byte store per ray straight into the memory of the GPU that owns the pixel (NVLink peer
// store through a CUDA-IPC mapping: the all-to-all is fused into the producing kernel, there is no
// staging buffer and no NCCL call on the data path).  A 32-thread kernel then raises a per-source
// flag in every peer, and the owner's compositor waits on its flags on the device, orders the
// segments front to back along each ray (cells are disjoint boxes: entry distance), composes
//      C = sum_i C_i prod_{j<i} (1 - a_j),   T = prod_i (1 - a_i)
// with the reference's early-termination rule (rt_core.cuh:289-303) applied per segment, blends
// the background and truncates to RGBA8 exactly like composite_and_write (renderer_kernel.cu:215-241).
#include <cstring>

#include "mnv_internal.cuh"
#include "mnv_march.cuh"

namespace mnv {
namespace {

constexpr int kMaxCells = 8;

struct CompositeParams {
    TreeView tree;
    mnv_camera cam;
    mnv_render_options opt;
    const float4 *partials;  // [n][block]
    int n, block;
    long long first_pixel;
    int n_pixels;
    float boxes[kMaxCells][6];  // tree-space cell boxes, slot order
    uint32_t *out;              // RGBA8 [n_pixels]
    const uint32_t *flags;      // [n] raised by the producers (may be null)
    uint32_t wait_value;
    // segments of a guided-sampling frame (mnv_render_nerf_results_partial): no early-termination
    // rule, and the frame is opaque — out[3] = 1, renderer_kernel.cu:315-316
    bool guided;
};

__global__ void signal_peers_kernel(uint32_t *d0, uint32_t *d1, uint32_t *d2, uint32_t *d3, uint32_t *d4,
                                    uint32_t *d5, uint32_t *d6, uint32_t *d7, int n, int slot, uint32_t value) {
    uint32_t *dst[8] = {d0, d1, d2, d3, d4, d5, d6, d7};
    // the march kernel that precedes this launch on the stream has completed: its peer stores are
    // performed; the fence orders them before the flag for every observer in the system
    __threadfence_system();
    if ((int) threadIdx.x < n) {
        volatile uint32_t *f = dst[threadIdx.x] + slot;
        *f = value;
    }
}

__global__ void __launch_bounds__(256) composite_partials_kernel(const CompositeParams p) {
    if (p.flags) {
        if (threadIdx.x < (unsigned) p.n) {
            const volatile uint32_t *f = p.flags + threadIdx.x;
            // frame counters only grow: >= also covers a producer that is already a frame ahead
            while ((int32_t) (*f - p.wait_value) < 0) __nanosleep(64);
            __threadfence_system();
        }
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_pixels) return;
    const long long pix = p.first_pixel + i;
    const int x = (int) (pix % p.cam.width), y = (int) (pix / p.cam.width);
    Ray r;
    setup_ray(p.tree, p.cam, p.opt, x, y, 1e9f, r);
    // entry distance of every cell (the slab test of _dda_world, rt_core.cuh:70-86)
    float key[kMaxCells];
    int order[kMaxCells];
    const float cc[3] = {r.c0, r.c1, r.c2}, ii[3] = {r.i0, r.i1, r.i2};
#pragma unroll
    for (int c = 0; c < kMaxCells; ++c) {
        order[c] = c;
        key[c] = 3.0e38f;
        if (c < p.n) {
            float tmin = 0.f, tmax = 1e4f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ci = (double) cc[a], inv = (double) ii[a];
                const float t1 = d2f(__dmul_rn(__dadd_rn(__dadd_rn((double) p.boxes[c][a], 1e-6), -ci), inv));
                const float t2 = d2f(__dmul_rn(__dadd_rn(__dadd_rn((double) p.boxes[c][a + 3], -1e-6), -ci), inv));
                tmin = fmaxf(tmin, fminf(t1, t2));
                tmax = fminf(tmax, fmaxf(t1, t2));
            }
            if (!(tmax < 0.f || tmin > tmax)) key[c] = tmin;
        }
    }
#pragma unroll
    for (int a = 1; a < kMaxCells; ++a) {  // insertion sort, n <= 8
#pragma unroll
        for (int b = a; b > 0; --b) {
            if (key[b] < key[b - 1]) {
                const float tk = key[b];
                key[b] = key[b - 1];
                key[b - 1] = tk;
                const int to = order[b];
                order[b] = order[b - 1];
                order[b - 1] = to;
            }
        }
    }
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, T = 1.f, alpha = 0.f;
    bool done = false;
#pragma unroll
    for (int k = 0; k < kMaxCells; ++k) {
        if (k < p.n && key[k] < 1.0e38f && !done) {
            const float4 s = p.partials[(size_t) order[k] * p.block + i];
            c0 = fmaf(T, s.x, c0);
            c1 = fmaf(T, s.y, c1);
            c2 = fmaf(T, s.z, c2);
            T *= 1.f - s.w;
            if (!p.guided && T < p.opt.stop_thresh) {  // rt_core.cuh:289-303
                const float scale = 1.f / (1.f - T);
                c0 *= scale;
                c1 *= scale;
                c2 *= scale;
                alpha = 1.f;
                done = true;
            }
        }
    }
    if (!done) alpha = 1.f - T;
    if (p.guided) alpha = 1.f;
    const float remain = (1.f - alpha) * p.opt.background_brightness;
    c0 += remain;
    c1 += remain;
    c2 += remain;
    p.out[i] = ref_to_u8(c0) | (ref_to_u8(c1) << 8) | (ref_to_u8(c2) << 16) | 0xff000000u;
}

}  // namespace

int launch_signal_peers(uint32_t *const *dst, int n, int slot, uint32_t value, cudaStream_t stream) {
    if (n < 1 || n > kMaxCells) {
        set_error("signal_peers: n = %d", n);
        return MNV_ERR_INVALID;
    }
    uint32_t *d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n; ++i) d[i] = dst[i];
    signal_peers_kernel<<<1, 32, 0, stream>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], n, slot, value);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

int launch_composite_partials(const DeviceTree &tree, const mnv_camera &cam, const mnv_render_options &opt,
                              const float *partials_dev, int n, int block, const float *boxes_host,
                              int64_t first_pixel, int n_pixels, uint8_t *rgba_dev, const uint32_t *flags_dev,
                              uint32_t wait_value, cudaStream_t stream, bool guided) {
    if (n < 1 || n > kMaxCells || n_pixels < 0 || n_pixels > block || !partials_dev || !rgba_dev || !boxes_host) {
        set_error("composite_partials: bad arguments (n = %d, block = %d, pixels = %d)", n, block, n_pixels);
        return MNV_ERR_INVALID;
    }
    if (n_pixels == 0) return MNV_OK;
    CompositeParams p;
    p.tree = make_view(tree);
    p.cam = cam;
    p.opt = opt;
    p.partials = reinterpret_cast<const float4 *>(partials_dev);
    p.n = n;
    p.block = block;
    p.first_pixel = first_pixel;
    p.n_pixels = n_pixels;
    std::memcpy(p.boxes, boxes_host, (size_t) n * 6 * sizeof(float));
    for (int c = n; c < kMaxCells; ++c)
        for (int a = 0; a < 6; ++a) p.boxes[c][a] = 0.f;
    p.out = reinterpret_cast<uint32_t *>(rgba_dev);
    p.flags = flags_dev;
    p.wait_value = wait_value;
    p.guided = guided;
    composite_partials_kernel<<<(unsigned) ((n_pixels + 255) / 256), 256, 0, stream>>>(p);
    MNV_CUDA(cudaGetLastError());
    return MNV_OK;
}

}  // namespace mnv
