// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Headless C driver around the UNMODIFIED reference sources, compiled where
// they lie under /root/reference by oracle/build_ref.sh into
// oracle/_ref/libref_render.so (and libref_render_instr.so for the visit-log
// variant).  It exposes the reference's own host launchers
// (include/cuda/renderer_kernel.hpp:12-79) through a tiny extern "C" surface so
// that tests/ and bench.py (--impl reference) can run the reference's CUDA
// kernel rebuilt for sm_100 next to the B200-native path.
//
// Flow mirrors main.cpp / cuda_renderer.cpp: N3Tree::open -> move_to_device ->
// Camera::_update -> viewer::render_voxels(..., offscreen=true).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "include/cuda/renderer_kernel.hpp"  // resolved with -I$REF (the reference root)

#ifdef REF_VISIT_LOG
// defined by the generated patch in the instrumented copy of renderer_kernel.cu
extern "C" void ref_instr_set_buffers(unsigned long long *hash, int *count, int *log, int log_cap);
#endif

namespace viewer {
extern int cuda_n_threads;  // src/cuda/renderer_kernel.cu:12 (threads per block of every launcher)
}

namespace {
struct RefCtx {
    viewer::N3Tree tree;
    long max_cap = 0;
    viewer::Camera *cam = nullptr;
    cudaStream_t stream = nullptr;
    cudaArray_t img = nullptr;
    cudaArray_t depth = nullptr;
    int w = 0, h = 0;
    torch::Tensor split, sample, visited;
};

void ensure_target(RefCtx *c, int w, int h) {
    if (c->img && c->w == w && c->h == h) return;
    if (c->img) cudaFreeArray(c->img);
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    cudaMallocArray(&c->img, &desc, w, h, cudaArraySurfaceLoadStore);
    c->w = w;
    c->h = h;
    auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
    c->split = torch::ones({(long) w * h, 3}, f32) * -1;
    c->sample = torch::ones({(long) w * h, 3}, f32) * -1;
}

void set_camera(RefCtx *c, int w, int h, const float *intr, const float *c2w) {
    viewer::Camera &cam = *c->cam;
    cam.width = w;
    cam.height = h;
    cam.fx = intr[0];
    cam.fy = intr[1];
    cam.cx = intr[2];
    cam.cy = intr[3];
    for (int col = 0; col < 4; ++col)
        for (int r = 0; r < 3; ++r) cam.transform[col][r] = c2w[col * 3 + r];
    cam._update(/*transform_from_vecs=*/false, /*copy_cuda=*/true);
}
}  // namespace

extern "C" {

void *ref_open(const char *npz_path, long max_capacity) {
    auto *c = new RefCtx();
    c->tree.open(npz_path);
    if (c->tree.N == 0) {
        delete c;
        return nullptr;
    }
    if (max_capacity < c->tree.capacity) max_capacity = c->tree.capacity;
    c->max_cap = max_capacity;
    c->tree.move_to_device(max_capacity, true, true);
    // move_to_device leaves sample_counts uninitialised on the device
    // (src/n3tree/n3tree.cpp:235-241); the host-side value is 8 (:191-193).
    c->tree.sample_counts.fill_(8);
    c->cam = new viewer::Camera();
    cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault);
    c->visited = torch::zeros({max_capacity},
                              torch::TensorOptions().device(torch::kCUDA).dtype(torch::kInt32));
    return c;
}

// The reference's loader alone (N3Tree::open -> cnpy::npz_load -> load_npz, src/n3tree/n3tree.cpp:16-205), host
// tensors only: no CUDA call, so the loader cross-check runs on machines without a GPU.  ref_download /
// ref_sample_counts / ref_capacity / ref_data_dim / ref_format work on such a context; nothing else does.
void *ref_open_host(const char *npz_path) {
    auto *c = new RefCtx();
    c->tree.open(npz_path);
    if (c->tree.N == 0) {
        delete c;
        return nullptr;
    }
    return c;
}

// "SH9" / "RGBA" ... as the reference parsed it (data_format.cpp:5-24) -> buf
int ref_format(void *ctx, char *buf, int n) {
    const std::string s = static_cast<RefCtx *>(ctx)->tree.data_format.to_string();
    std::snprintf(buf, (size_t) n, "%s", s.c_str());
    return static_cast<RefCtx *>(ctx)->tree.data_format.basis_dim;
}

int ref_sample_counts(void *ctx, int16_t *out) {
    auto *c = static_cast<RefCtx *>(ctx);
    auto t = c->tree.sample_counts.slice(0, 0, c->tree.capacity).cpu().contiguous();
    std::memcpy(out, t.data_ptr(), t.numel() * 2);
    return 0;
}

void ref_close(void *ctx) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (!c) return;
    if (c->img) cudaFreeArray(c->img);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c->cam;
    delete c;
}

int ref_capacity(void *ctx) { return static_cast<RefCtx *>(ctx)->tree.capacity; }
int ref_data_dim(void *ctx) { return static_cast<RefCtx *>(ctx)->tree.data_dim; }

// Copies the loaded tree (first `capacity` rows) back to host arrays in the
// reference's own layout, so the caller can feed the identical tree to the
// B200-native path without going through a second loader.
int ref_download(void *ctx, uint16_t *data, int32_t *child, int32_t *parent, float *scale,
                 float *offset) {
    auto *c = static_cast<RefCtx *>(ctx);
    const long cap = c->tree.capacity;
    if (data) {
        auto t = c->tree.data.slice(0, 0, cap).cpu().contiguous();
        std::memcpy(data, t.data_ptr(), t.numel() * 2);
    }
    if (child) {
        auto t = c->tree.child.slice(0, 0, cap).cpu().contiguous();
        std::memcpy(child, t.data_ptr(), t.numel() * 4);
    }
    if (parent) {
        auto t = c->tree.parent.slice(0, 0, cap).cpu().contiguous();
        std::memcpy(parent, t.data_ptr(), t.numel() * 4);
    }
    if (scale) {
        auto t = c->tree.scale.cpu();
        std::memcpy(scale, t.data_ptr(), 12);
    }
    if (offset) {
        auto t = c->tree.offset.cpu();
        std::memcpy(offset, t.data_ptr(), 12);
    }
    return 0;
}

// viewer::render_voxels (src/cuda/renderer_kernel.cu:396-437) into an
// offscreen RGBA8 cudaArray; `iters` back-to-back launches, CUDA-event timed
// individually on the launching stream. Outputs are host pointers (nullable).
int ref_render_voxels(void *ctx, int w, int h, const float *intr, const float *c2w,
                      const void *opt_pod, int opt_size, unsigned char *rgba_out, float *split_out,
                      float *sample_out, int iters, float *ms_out) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) {
        fprintf(stderr, "ref_render_voxels: RenderOptions size mismatch %d vs %zu\n", opt_size,
                sizeof(viewer::RenderOptions));
        return 1;
    }
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < std::max(iters, 1); ++it) {
        c->split.fill_(-1);
        c->sample.fill_(-1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, c->stream);
        viewer::render_voxels(c->tree, *c->cam, opt, c->img, c->depth, c->stream, c->split,
                              c->sample, c->visited, /*track_visit=*/false, /*offscreen=*/true);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) ms_out[it] = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        fprintf(stderr, "ref_render_voxels: %s\n", cudaGetErrorString(err));
        return 2;
    }
    if (rgba_out)
        cudaMemcpy2DFromArray(rgba_out, (size_t) w * 4, c->img, 0, 0, (size_t) w * 4, h,
                              cudaMemcpyDeviceToHost);
    if (split_out) {
        auto t = c->split.cpu();
        std::memcpy(split_out, t.data_ptr(), t.numel() * 4);
    }
    if (sample_out) {
        auto t = c->sample.cpu();
        std::memcpy(sample_out, t.data_ptr(), t.numel() * 4);
    }
    return 0;
}

// The whole reference frame as a host application sees it with host buffers:
// camera upload + tracker fills + kernel + read-back of the RGBA8 frame.
// Wall-clock around the synchronous sequence; used for the e2e reference arm.
int ref_render_frame_host(void *ctx, int w, int h, const float *intr, const float *c2w,
                          const void *opt_pod, int opt_size, unsigned char *rgba_out) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    c->split.fill_(-1);   // cuda_renderer.cpp:97-98
    c->sample.fill_(-1);
    viewer::render_voxels(c->tree, *c->cam, opt, c->img, c->depth, c->stream, c->split, c->sample,
                          c->visited, false, true);
    cudaMemcpy2DFromArrayAsync(rgba_out, (size_t) w * 4, c->img, 0, 0, (size_t) w * 4, h,
                               cudaMemcpyDeviceToHost, c->stream);
    return cudaStreamSynchronize(c->stream) == cudaSuccess ? 0 : 2;
}

// viewer::get_samples_from_voxels (src/cuda/renderer_kernel.cu:439-485), offscreen, with the
// dense buffers of Impl::init_sample_tensor / init_split_tracker (cuda_renderer.cpp:460-496):
// guided_samples [P][S][sd] pre-filled with -1 in column 0 (:110), num_samples zeroed (:109).
int ref_get_samples(void *ctx, int w, int h, const float *intr, const float *c2w, const void *opt_pod,
                    int opt_size, const int *grid_dim, const float *min_position, const float *range,
                    int S, int sd, short *num_samples_out, float *samples_out, short *cluster_out,
                    float *split_out, float *sample_out) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    const long P = (long) w * h;
    auto cuda = torch::TensorOptions().device(torch::kCUDA);
    torch::Tensor num_samples = torch::zeros({P}, cuda.dtype(torch::kInt16));
    torch::Tensor samples = torch::ones({P, S, sd}, cuda.dtype(torch::kFloat32)) * -1;
    torch::Tensor cluster = torch::zeros({P, S}, cuda.dtype(torch::kInt16));
    torch::Tensor gd = torch::from_blob((void *) grid_dim, {2}, torch::kInt32).clone().to(torch::kCUDA);
    torch::Tensor mp = torch::from_blob((void *) min_position, {3}, torch::kFloat32).clone().to(torch::kCUDA);
    torch::Tensor rg = torch::from_blob((void *) range, {3}, torch::kFloat32).clone().to(torch::kCUDA);
    c->split.fill_(-1);
    c->sample.fill_(-1);
    viewer::get_samples_from_voxels(c->tree, *c->cam, opt, c->depth, c->stream, c->split, c->sample,
                                    c->visited, /*track_visit=*/false, /*offscreen=*/true,
                                    num_samples, samples, cluster, gd, mp, rg);
    if (cudaDeviceSynchronize() != cudaSuccess) return 2;
    auto ns = num_samples.cpu();
    std::memcpy(num_samples_out, ns.data_ptr(), P * 2);
    auto sm = samples.cpu();
    std::memcpy(samples_out, sm.data_ptr(), (size_t) P * S * sd * 4);
    auto cl = cluster.cpu();
    std::memcpy(cluster_out, cl.data_ptr(), (size_t) P * S * 2);
    if (split_out) {
        auto t = c->split.cpu();
        std::memcpy(split_out, t.data_ptr(), t.numel() * 4);
    }
    if (sample_out) {
        auto t = c->sample.cpu();
        std::memcpy(sample_out, t.data_ptr(), t.numel() * 4);
    }
    return 0;
}

// viewer::render_nerf_results (src/cuda/renderer_kernel.cu:365-394), offscreen.
int ref_render_nerf_results(void *ctx, int w, int h, const float *intr, const float *c2w,
                            const void *opt_pod, int opt_size, const float *sample_values, long V,
                            int vdim, const float *z_vals, const long *offsets,
                            unsigned char *rgba_out) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    const long P = (long) w * h;
    torch::Tensor sv = torch::from_blob((void *) sample_values, {V, vdim}, torch::kFloat32).clone().to(torch::kCUDA);
    torch::Tensor zv = torch::from_blob((void *) z_vals, {V}, torch::kFloat32).clone().to(torch::kCUDA);
    torch::Tensor of = torch::from_blob((void *) offsets, {P}, torch::kInt64).clone().to(torch::kCUDA);
    cudaGetLastError();
    // The reference never checks its launches.  Rebuilt for sm_100 this kernel needs 168
    // registers x 512 threads per block (auto_cuda_threads, renderer_kernel.cu:14-28) = 86016
    // > 65536 registers per SM: as shipped the launch fails with "too many resources requested".
    // The block size is the reference's own global knob (viewer::cuda_n_threads,
    // renderer_kernel.cu:12; auto_cuda_threads leaves any value other than -1 alone), so this
    // launch runs with 256 threads per block — same sources, same arithmetic, launch geometry
    // only — and the knob is restored afterwards.
    const int saved_threads = viewer::cuda_n_threads;
    viewer::cuda_n_threads = 256;
    viewer::render_nerf_results(c->tree, *c->cam, opt, c->img, c->stream, sv, zv, of, /*offscreen=*/true);
    viewer::cuda_n_threads = saved_threads;
    if (cudaGetLastError() != cudaSuccess) return 3;
    if (cudaDeviceSynchronize() != cudaSuccess) return 2;
    cudaMemcpy2DFromArray(rgba_out, (size_t) w * 4, c->img, 0, 0, (size_t) w * 4, h,
                          cudaMemcpyDeviceToHost);
    return 0;
}

// ---- refinement launchers (include/cuda/renderer_kernel.hpp:49-79) -------------------------
static void cluster_tensors(const int *grid_dim, const float *min_position, const float *range,
                            torch::Tensor &gd, torch::Tensor &mp, torch::Tensor &rg) {
    gd = torch::from_blob((void *) grid_dim, {2}, torch::kInt32).clone().to(torch::kCUDA);
    mp = torch::from_blob((void *) min_position, {3}, torch::kFloat32).clone().to(torch::kCUDA);
    rg = torch::from_blob((void *) range, {3}, torch::kFloat32).clone().to(torch::kCUDA);
}

// viewer::add_children_and_generate_samples; then capacity += n like expand_voxels
// (cuda_renderer.cpp:275).  samples: in = U[0,1) numbers, out = world-space rows.
int ref_add_children(void *ctx, const void *opt_pod, int opt_size, const int *parent_nodes, int n,
                     float *samples, int spc, int rand_dim, short *cluster_out, const int *grid_dim,
                     const float *min_position, const float *range) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    if (c->tree.capacity + n > c->max_cap) return 4;
    auto cuda = torch::TensorOptions().device(torch::kCUDA);
    torch::Tensor pn = torch::from_blob((void *) parent_nodes, {n, 2}, torch::kInt32).clone().to(torch::kCUDA);
    torch::Tensor sm = torch::from_blob(samples, {(long) n * 8, spc, rand_dim}, torch::kFloat32).clone().to(torch::kCUDA);
    torch::Tensor cl = torch::zeros({(long) n * 8, spc}, cuda.dtype(torch::kInt16));
    torch::Tensor gd, mp, rg;
    cluster_tensors(grid_dim, min_position, range, gd, mp, rg);
    viewer::add_children_and_generate_samples(c->tree, opt, pn, sm, cl, c->visited, gd, mp, rg);
    if (cudaDeviceSynchronize() != cudaSuccess) return 2;
    c->tree.capacity += n;
    auto smc = sm.cpu();
    std::memcpy(samples, smc.data_ptr(), smc.numel() * 4);
    auto clc = cl.cpu();
    std::memcpy(cluster_out, clc.data_ptr(), clc.numel() * 2);
    return 0;
}

// viewer::generate_samples for existing leaves.
int ref_generate_samples(void *ctx, const void *opt_pod, int opt_size, const int *nodes, int m,
                         float *samples, int spc, int rand_dim, short *cluster_out,
                         const int *grid_dim, const float *min_position, const float *range) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    auto cuda = torch::TensorOptions().device(torch::kCUDA);
    torch::Tensor nd = torch::from_blob((void *) nodes, {m, 2}, torch::kInt32).clone().to(torch::kCUDA);
    torch::Tensor sm = torch::from_blob(samples, {m, spc, rand_dim}, torch::kFloat32).clone().to(torch::kCUDA);
    torch::Tensor cl = torch::zeros({m, spc}, cuda.dtype(torch::kInt16));
    torch::Tensor gd, mp, rg;
    cluster_tensors(grid_dim, min_position, range, gd, mp, rg);
    viewer::generate_samples(c->tree, opt, nd, sm, cl, gd, mp, rg);
    if (cudaDeviceSynchronize() != cudaSuccess) return 2;
    auto smc = sm.cpu();
    std::memcpy(samples, smc.data_ptr(), smc.numel() * 4);
    auto clc = cl.cpu();
    std::memcpy(cluster_out, clc.data_ptr(), clc.numel() * 2);
    return 0;
}

// Impl::prune_tree (cuda_renderer.cpp:343-381) on host-provided marks: cumsum, argmin,
// viewer::adjust_parents_and_children, then the chunked gather of data/child/parent.
int ref_prune(void *ctx, const unsigned char *to_delete_host) {
    auto *c = static_cast<RefCtx *>(ctx);
    const long cap = c->tree.capacity;
    torch::Tensor to_delete = torch::from_blob((void *) to_delete_host, {cap}, torch::kUInt8).clone().to(torch::kCUDA).to(torch::kBool);
    int num_to_delete = to_delete.sum().item().toInt();
    if (num_to_delete == 0) return 0;
    torch::Tensor index_shifts = torch::cumsum(to_delete, 0, torch::kInt32);
    int first_shift_index = index_shifts.argmin().item().toInt();
    viewer::adjust_parents_and_children(c->tree, first_shift_index, to_delete, index_shifts);
    torch::Tensor keep = torch::arange(first_shift_index, cap, torch::TensorOptions().device(torch::kCUDA))
                                 .index({to_delete.slice(0, first_shift_index, cap) == false});
    const long kept = keep.size(0);
    c->tree.data.slice(0, first_shift_index, first_shift_index + kept) = c->tree.data.index({keep}).clone();
    c->tree.child.slice(0, first_shift_index, first_shift_index + kept) = c->tree.child.index({keep}).clone();
    c->tree.parent.slice(0, first_shift_index, first_shift_index + kept) = c->tree.parent.index({keep}).clone();
    c->tree.capacity -= num_to_delete;
    return cudaDeviceSynchronize() == cudaSuccess ? num_to_delete : -2;
}

#ifdef REF_VISIT_LOG
// Instrumented build only: one render with per-ray visit hash / count / log.
int ref_render_voxels_logged(void *ctx, int w, int h, const float *intr, const float *c2w,
                             const void *opt_pod, int opt_size, unsigned char *rgba_out,
                             unsigned long long *hash_out, int *count_out, int *log_out,
                             int log_cap) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    const size_t P = (size_t) w * h;
    unsigned long long *d_hash = nullptr;
    int *d_count = nullptr, *d_log = nullptr;
    cudaMalloc(&d_hash, P * 8);
    cudaMalloc(&d_count, P * 4);
    // FNV-1a offset basis, same as the native logged kernel
    std::vector<unsigned long long> init(P, 0xcbf29ce484222325ULL);
    cudaMemcpy(d_hash, init.data(), P * 8, cudaMemcpyHostToDevice);
    cudaMemset(d_count, 0, P * 4);
    if (log_out && log_cap > 0) {
        cudaMalloc(&d_log, P * (size_t) log_cap * 4);
        cudaMemset(d_log, 0xff, P * (size_t) log_cap * 4);
    }
    ref_instr_set_buffers(d_hash, d_count, d_log, d_log ? log_cap : 0);
    c->split.fill_(-1);
    c->sample.fill_(-1);
    viewer::render_voxels(c->tree, *c->cam, opt, c->img, c->depth, c->stream, c->split, c->sample,
                          c->visited, false, true);
    cudaError_t err = cudaDeviceSynchronize();
    ref_instr_set_buffers(nullptr, nullptr, nullptr, 0);
    if (err != cudaSuccess) {
        fprintf(stderr, "ref_render_voxels_logged: %s\n", cudaGetErrorString(err));
        return 2;
    }
    if (rgba_out)
        cudaMemcpy2DFromArray(rgba_out, (size_t) w * 4, c->img, 0, 0, (size_t) w * 4, h,
                              cudaMemcpyDeviceToHost);
    if (hash_out) cudaMemcpy(hash_out, d_hash, P * 8, cudaMemcpyDeviceToHost);
    if (count_out) cudaMemcpy(count_out, d_count, P * 4, cudaMemcpyDeviceToHost);
    if (log_out && d_log) cudaMemcpy(log_out, d_log, P * (size_t) log_cap * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_hash);
    cudaFree(d_count);
    if (d_log) cudaFree(d_log);
    return 0;
}
#endif

// Camera parity: runs the script of `mnv_headless --selftest-camera` on the reference's own
// viewer::Camera (src/camera.cpp) and dumps, per step, transform (12), K (16), w2c (16),
// fx, fy, cx, cy -> out[3][48].  Needs a CUDA device (the constructor cudaMallocs).
int ref_camera_trace(int w, int h, float fx, float *out) {
    viewer::Camera c(w, h, fx);
    auto dump = [&](int step) {
        float *o = out + step * 48;
        int k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) o[k++] = c.transform[i][j];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) o[k++] = c.K[i][j];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) o[k++] = c.w2c[i][j];
        o[k++] = c.fx;
        o[k++] = c.fy;
        o[k++] = c.cx;
        o[k++] = c.cy;
    };
    dump(0);
    c.begin_drag(100.f, 120.f, false, false);
    c.drag_update(260.f, 90.f);
    c.end_drag();
    c._update();
    dump(1);
    c.begin_drag(10.f, 10.f, true, false);
    c.drag_update(40.f, 70.f);
    c.end_drag();
    c.move(glm::vec3(0.1f, -0.2f, 0.3f));
    c._update();
    dump(2);
    return 0;
}

// viewer::render_voxels the way Impl::render calls it (cuda_renderer.cpp:141-142): offscreen = false, the
// RGBA8 surface already holds what GL drew (prior_rgba), the R32F surface the mesh depth t_max (depth), visit
// tracking optional.  Host pointers in / out; visited_out int32 [max_capacity].
int ref_render_voxels_interop(void *ctx, int w, int h, const float *intr, const float *c2w, const void *opt_pod,
                              int opt_size, const unsigned char *prior_rgba, const float *depth,
                              unsigned char *rgba_out, int track_visit, int *visited_out) {
    auto *c = static_cast<RefCtx *>(ctx);
    if (opt_size != (int) sizeof(viewer::RenderOptions)) return 1;
    viewer::RenderOptions opt;
    std::memcpy(&opt, opt_pod, sizeof(opt));
    ensure_target(c, w, h);
    set_camera(c, w, h, intr, c2w);
    cudaArray_t darr = nullptr;
    cudaChannelFormatDesc fdesc = cudaCreateChannelDesc<float>();
    cudaMallocArray(&darr, &fdesc, w, h, cudaArraySurfaceLoadStore);
    cudaMemcpy2DToArray(c->img, 0, 0, prior_rgba, (size_t) w * 4, (size_t) w * 4, h, cudaMemcpyHostToDevice);
    cudaMemcpy2DToArray(darr, 0, 0, depth, (size_t) w * 4, (size_t) w * 4, h, cudaMemcpyHostToDevice);
    c->split.fill_(-1);
    c->sample.fill_(-1);
    c->visited.zero_();
    c->visited[0] = 1;  // cuda_renderer.cpp:506
    cudaDeviceSynchronize();
    viewer::render_voxels(c->tree, *c->cam, opt, c->img, darr, c->stream, c->split, c->sample, c->visited,
                          track_visit != 0, /*offscreen=*/false);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        fprintf(stderr, "ref_render_voxels_interop: %s\n", cudaGetErrorString(err));
        cudaFreeArray(darr);
        return 2;
    }
    cudaMemcpy2DFromArray(rgba_out, (size_t) w * 4, c->img, 0, 0, (size_t) w * 4, h, cudaMemcpyDeviceToHost);
    if (visited_out) {
        auto t = c->visited.cpu();
        std::memcpy(visited_out, t.data_ptr(), t.numel() * 4);
    }
    cudaFreeArray(darr);
    return 0;
}

}  // extern "C"
