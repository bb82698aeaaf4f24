// Frame orchestration of the reference's VolumeRenderer::Impl (src/renderer/cuda_renderer.cpp)
// on top of the C-ABI: no LibTorch tensors, no per-cluster .item() syncs, no GL calls.
#include "renderer.hpp"

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdlib>
#include <map>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mnv_b200.h"
#include "model.hpp"

namespace viewer {

static_assert(sizeof(RenderOptions) == sizeof(mnv_render_options), "RenderOptions must match the C-ABI POD");
static_assert(offsetof(RenderOptions, max_guided_samples) == offsetof(mnv_render_options, max_guided_samples),
              "RenderOptions field order must match the C-ABI POD");

namespace {
void ck(int rc, const char *what) {
    if (rc != MNV_OK) throw std::runtime_error(std::string(what) + ": " + mnv_last_error());
}
struct DevBuf {  // grow-only device scratch
    void *p = nullptr;
    size_t bytes = 0;
    void reserve(size_t n) {
        if (n <= bytes) return;
        if (p) mnv_free(p);
        p = nullptr;
        bytes = 0;
        ck(mnv_malloc(&p, n, 0), "mnv_malloc");
        bytes = n;
    }
    void release() {
        if (p) mnv_free(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};
}  // namespace

struct VolumeRenderer::Impl {
    Impl(VolumeRenderer &owner) : self(owner), camera(owner.camera), options(owner.options) {
        timing = std::getenv("MNV_TIMING") != nullptr;
    }
    // dev: MNV_TIMING=1 synchronises after every stage and prints the per-stage totals at exit
    bool timing = false;
    std::map<std::string, std::pair<double, long>> stage_ms;
    std::chrono::steady_clock::time_point t_stage;
    void tick() {
        if (timing) {
            mnv_stream_synchronize(stream);
            t_stage = std::chrono::steady_clock::now();
        }
    }
    void tock(const char *label) {
        if (!timing) return;
        mnv_stream_synchronize(stream);
        const auto now = std::chrono::steady_clock::now();
        auto &e = stage_ms[label];
        e.first += std::chrono::duration<double, std::milli>(now - t_stage).count();
        e.second += 1;
        t_stage = now;
    }
    ~Impl() {
        if (timing)
            for (auto &kv : stage_ms)
                std::fprintf(stderr, "[mnv timing] %-28s %8.3f ms/call x %ld\n", kv.first.c_str(),
                             kv.second.first / kv.second.second, kv.second.second);
        for (DevBuf *b : {&frame, &split_tracker, &sample_tracker, &visit_tracker, &offsets, &z_vals, &rows,
                          &cluster, &values, &nodes, &rand, &rcluster, &results})
            b->release();
        if (frame_pinned) mnv_free_host(frame_pinned);
        if (stream) mnv_stream_destroy(stream);
        group_models.clear();
        if (group) mnv_group_destroy(group);
    }

    void start() {
        if (started) return;
        ck(mnv_stream_create(&stream, 0), "stream");  // cuda_renderer.cpp:45
        init_trackers();
        started = true;
    }

    int64_t pixels() const { return (int64_t) camera.width * camera.height; }

    // init_split_tracker, cuda_renderer.cpp:460-470
    void init_trackers() {
        const size_t P = (size_t) pixels();
        frame.reserve(P * 4);
        split_tracker.reserve(P * 3 * sizeof(float));
        sample_tracker.reserve(P * 3 * sizeof(float));
        if (frame_pinned_bytes < P * 4) {
            if (frame_pinned) mnv_free_host(frame_pinned);
            frame_pinned = nullptr;
            ck(mnv_malloc_host(&frame_pinned, P * 4), "pinned frame");
            frame_pinned_bytes = P * 4;
        }
        can_reuse_results = false;
    }

    const mnv_render_options *opt() const { return reinterpret_cast<const mnv_render_options *>(&options); }

    int in_dim() const {
        return 3 + (options.need_viewdir ? 3 : 0) + (options.appearance_embedding != -1 ? 1 : 0);
    }

    void render() {
        start();
        camera._update();
        self.last_frame = FrameInfo();
        if (tree == nullptr) return;
        if (group) {
            render_group();
            return;
        }
        if (tree->device_tree == nullptr) return;
        mnv_tree *dt = tree->device_tree;
        mnv_camera cam;
        camera.fill(cam);
        const int64_t P = pixels();
        init_trackers_if_resized();

        tick();
        ck(mnv_fill_f32(split_tracker.as<float>(), -1.f, P * 3, stream), "fill");
        ck(mnv_fill_f32(sample_tracker.as<float>(), -1.f, P * 3, stream), "fill");
        tock("fill trackers");
        const bool camera_has_changed = camera.has_changed();
        const bool track_visit =
                (camera_has_changed && tree->capacity > max_tree_capacity * 3 / 4) || prune_happened;
        if (camera_has_changed) can_reuse_results = false;

        void *image_arr = interop ? ca[buf_index * 2] : nullptr;
        void *depth_arr = interop ? ca[buf_index * 2 + 1] : nullptr;
        uint8_t *linear = interop ? nullptr : frame.as<uint8_t>();
        const bool offscreen = !interop;

        if (options.use_guided_sampling && !camera.is_dragging()) {
            if (!model.device_model) throw std::runtime_error("use_guided_sampling needs load_model()");
            if (!can_reuse_results) {
                guided_total = gather_guided_samples(dt, cam, depth_arr, offscreen, track_visit);
                tock("guided: emit samples");
                values.reserve((size_t) std::max<int64_t>(guided_total, 1) * (tree->data_dim + 1) * sizeof(float));
                ck(mnv_query_submodules(model.device_model, cluster.as<int16_t>(), rows.as<float>(), in_dim(),
                                        guided_total, values.as<float>(), tree->data_dim + 1, stream),
                   "query_submodules");
                tock("guided: query_submodules");
                self.last_frame.guided_rows = guided_total;
                can_reuse_results = true;
            }
            // The reference composites with sigma read from column 3 whatever the format
            // (rt_core.cuh:365), which is only right for RGBA trees; the column that holds sigma
            // (data_dim - 1, == 3 for RGBA) is used here.
            ck(mnv_render_nerf_results(dt, &cam, opt(), image_arr, linear, values.as<float>(), tree->data_dim + 1,
                                       tree->data_dim - 1, z_vals.as<float>(), offsets.as<int64_t>(), offscreen,
                                       stream),
               "render_nerf_results");
            tock("guided: composite");
        } else {
            ck(mnv_render_voxels(dt, &cam, opt(), image_arr, depth_arr, linear, split_tracker.as<float>(),
                                 sample_tracker.as<float>(), visit_tracker.as<int32_t>(), track_visit, offscreen,
                                 stream),
               "render_voxels");
            tock("render_voxels");
        }

        if (options.use_splitting && !camera.is_dragging()) expand_voxels(dt);

        if (max_tree_capacity - tree->capacity < options.split_batch_size) {
            prune_tree(dt);
            prune_happened = true;
        } else {
            prune_happened = false;
        }
        self.last_frame.capacity = tree->capacity;
        frame_host_valid = false;
        if (interop) {
            ck(mnv_stream_synchronize(stream), "sync");
            buf_index ^= 1;
        }
    }

    // Multi-GPU frame: interleaved bands on every replica, NVLink gather into devices[0], refinement across the group.
    void render_group() {
        mnv_camera cam;
        camera.fill(cam);
        init_trackers_if_resized();
        if (options.use_guided_sampling && !camera.is_dragging())
            throw std::runtime_error("guided sampling is a single-GPU feature: do not call set_devices()");
        camera.has_changed();
        void *image_arr = interop ? ca[buf_index * 2] : nullptr;
        uint8_t *linear = frame.as<uint8_t>();
        if (options.use_splitting && !camera.is_dragging()) {
            if (group_models.empty()) throw std::runtime_error("use_splitting needs load_model()");
            std::vector<mnv_model *> mh;
            for (auto &m : group_models) mh.push_back(m->device_model);
            int added = 0;
            ck(mnv_group_refine_frame(group, mh.data(), &cam, opt(), group_models[0]->grid_dim,
                                      group_models[0]->min_position, group_models[0]->range, self.rng_seed, linear,
                                      image_arr, nullptr, 8, &added),
               "group refine");
            frame_host_valid = false;
            self.last_frame.added = added;
            self.last_frame.split_candidates = added > 0 ? added : 0;
        } else {
            ck(mnv_group_render_frame(group, &cam, opt(), linear, image_arr, 8), "group frame");
            frame_host_valid = false;
        }
        // the gathered frame was produced on the replicas' streams: frame_host() / the GL blit read it from another
        // stream, so the frame is completed here (one host synchronisation per frame)
        ck(mnv_group_synchronize(group), "group sync");
        mnv_tree *t0 = nullptr;
        ck(mnv_group_tree(group, 0, &t0), "group tree");
        int64_t c = 0, m = 0;
        mnv_tree_capacity(t0, &c, &m);
        tree->capacity = (int) c;
        self.last_frame.capacity = c;
        if (interop) buf_index ^= 1;
    }

    void init_trackers_if_resized() {
        if (frame.bytes < (size_t) pixels() * 4) init_trackers();
    }

    // get_samples_from_voxels + cumsum + mask compaction (cuda_renderer.cpp:106-120), CSR form;
    // the row buffers grow on demand instead of the reference's dense [P][S] worst case.
    int64_t gather_guided_samples(mnv_tree *dt, const mnv_camera &cam, void *depth_arr, bool offscreen,
                                  bool track_visit) {
        const int64_t P = pixels();
        offsets.reserve((size_t) P * sizeof(int64_t));
        if (guided_capacity == 0) grow_guided(std::max<int64_t>(P * 8, 1 << 16));
        for (int attempt = 0; attempt < 2; ++attempt) {
            int64_t total = 0;
            const int rc = mnv_guided_samples(dt, &cam, opt(), depth_arr, offscreen, model.grid_dim,
                                              model.min_position, model.range, offsets.as<int64_t>(),
                                              z_vals.as<float>(), rows.as<float>(), in_dim(),
                                              cluster.as<int16_t>(), guided_capacity, &total,
                                              split_tracker.as<float>(), sample_tracker.as<float>(),
                                              visit_tracker.as<int32_t>(), track_visit, stream);
            if (rc == MNV_OK) return total;
            if (rc != MNV_ERR_FULL || attempt == 1) ck(rc, "guided_samples");
            grow_guided(total + total / 4);
        }
        return 0;
    }
    void grow_guided(int64_t n_rows) {
        guided_capacity = n_rows;
        z_vals.reserve((size_t) n_rows * sizeof(float));
        rows.reserve((size_t) n_rows * 7 * sizeof(float));  // widest row: xyz + dir + appearance
        cluster.reserve((size_t) n_rows * sizeof(int16_t));
    }

    int rand_dim() const { return in_dim(); }  // cuda_renderer.cpp:240-248

    // Impl::expand_voxels, cuda_renderer.cpp:205-278
    void expand_voxels(mnv_tree *dt) {
        if (!model.device_model) throw std::runtime_error("use_splitting needs load_model()");
        const int batch = std::max(options.split_batch_size, 1);
        nodes.reserve((size_t) batch * 2 * sizeof(int32_t));
        int n = 0, n_cand = 0;
        ck(mnv_select_split_candidates(split_tracker.as<float>(), pixels(), batch, nodes.as<int32_t>(), &n, &n_cand,
                                       stream),
           "select_split_candidates");
        tock("refine: select candidates");
        self.last_frame.split_candidates = n_cand;
        if (self.verbose) std::printf("Split candidates: %d\n", n_cand);
        if (n_cand == 0) {
            get_more_samples(dt);
            return;
        }
        if (tree->capacity + n > max_tree_capacity) {
            if (self.verbose) std::printf("Full\n");
            return;
        }
        const int c = options.samples_per_corner, rd = rand_dim(), D = tree->data_dim;
        const int64_t n_rows = (int64_t) n * 8 * c;
        rand.reserve((size_t) n_rows * rd * sizeof(float));
        rcluster.reserve((size_t) n_rows * sizeof(int16_t));
        results.reserve((size_t) n_rows * (D + 1) * sizeof(float));
        ck(mnv_fill_uniform(rand.as<float>(), n_rows * rd, self.rng_seed + (++rng_calls), stream), "rand");
        ck(mnv_add_children_and_generate_samples(dt, opt(), nodes.as<int32_t>(), n, rand.as<float>(),
                                                 rcluster.as<int16_t>(), visit_tracker.as<int32_t>(),
                                                 model.grid_dim, model.min_position, model.range, stream),
           "add_children_and_generate_samples");
        tock("refine: rand + add children");
        ck(mnv_query_submodules(model.device_model, rcluster.as<int16_t>(), rand.as<float>(), rd, n_rows,
                                results.as<float>(), D + 1, stream),
           "query_submodules");
        tock("refine: query_submodules");
        ck(mnv_tree_commit_children(dt, opt(), n, results.as<float>(), D + 1, stream), "commit_children");
        tock("refine: commit children");
        tree->sync_capacity();
        self.last_frame.added = n;
        if (self.verbose) std::printf("Added: %d, total size: %d\n", n, tree->capacity);
        can_reuse_results = false;
    }

    // Impl::get_more_samples, cuda_renderer.cpp:280-341.  rand_dim follows expand_voxels (the
    // reference allocates 3 columns here whatever the model needs, :301 — quirk 5 of SURVEY.md).
    void get_more_samples(mnv_tree *dt) {
        const int batch = std::max(options.split_batch_size, 1);
        int m = 0, n_cand = 0;
        ck(mnv_select_sample_candidates(sample_tracker.as<float>(), pixels(), batch, nodes.as<int32_t>(), &m,
                                        &n_cand, stream),
           "select_sample_candidates");
        if (n_cand == 0) return;
        if (self.verbose) std::printf("Sample candidates: %d\n", n_cand);
        const int c = options.samples_per_corner, rd = rand_dim(), D = tree->data_dim;
        const int64_t n_rows = (int64_t) m * c;
        rand.reserve((size_t) n_rows * rd * sizeof(float));
        rcluster.reserve((size_t) n_rows * sizeof(int16_t));
        results.reserve((size_t) n_rows * (D + 1) * sizeof(float));
        ck(mnv_fill_uniform(rand.as<float>(), n_rows * rd, self.rng_seed + (++rng_calls), stream), "rand");
        ck(mnv_generate_samples(dt, opt(), nodes.as<int32_t>(), m, rand.as<float>(), rcluster.as<int16_t>(),
                                model.grid_dim, model.min_position, model.range, stream),
           "generate_samples");
        ck(mnv_query_submodules(model.device_model, rcluster.as<int16_t>(), rand.as<float>(), rd, n_rows,
                                results.as<float>(), D + 1, stream),
           "query_submodules");
        ck(mnv_tree_update_samples(dt, opt(), nodes.as<int32_t>(), m, results.as<float>(), D + 1, stream),
           "update_samples");
        self.last_frame.resampled = m;
        can_reuse_results = false;
    }

    // Impl::prune_tree, cuda_renderer.cpp:343-381
    void prune_tree(mnv_tree *dt) {
        if (self.verbose) std::printf("Pruning\n");
        int64_t num = 0;
        ck(mnv_tree_prune_unvisited(dt, visit_tracker.as<int32_t>(), &num, stream), "prune");
        tree->sync_capacity();
        self.last_frame.pruned = num;
        if (self.verbose) {
            if (num == 0) std::printf("Nothing can be pruned\n");
            else std::printf("Pruning finished - reclaimed: %lld\n", (long long) num);
        }
    }

    // Impl::resize, cuda_renderer.cpp:383-458 (intrinsics rescale; GL storage is the caller's)
    void resize(int width, int height) {
        if (camera.width == width && camera.height == height) return;
        start();
        const float wr = (float) width / camera.width, hr = (float) height / camera.height;
        if (!initial_resize) {
            camera.fx *= wr;
            camera.default_fx *= wr;
            camera.fy *= hr;
            camera.default_fy *= hr;
            camera.cy *= hr;
            if (camera.default_cx != -1) camera.cx *= wr;
            if (camera.default_cy != -1) camera.cy *= hr;
        } else {
            initial_resize = false;
        }
        if (camera.default_cx == -1) camera.cx = (float) (width / 2);
        if (camera.default_cy == -1) camera.cy = (float) (height / 2);
        camera.width = width;
        camera.height = height;
        init_trackers();
        guided_capacity = 0;
        if (interop && tree && tree->device_tree) mnv_tree_release_surfaces(tree->device_tree);
        interop = false;  // surfaces of the old size are gone; the caller re-registers
    }

    // Impl::set, cuda_renderer.cpp:498-516
    void set(N3Tree &t, long max_capacity) {
        start();
        if (devices.size() > 1) {
            set_group(t, max_capacity);
            return;
        }
        t.move_to_device(max_capacity, true, true);
        tree = &t;
        int64_t c = 0, m = 0;
        mnv_tree_capacity(t.device_tree, &c, &m);
        max_tree_capacity = m;
        visit_tracker.reserve((size_t) m * sizeof(int32_t));
        ck(mnv_memset(visit_tracker.p, 0, (size_t) m * sizeof(int32_t), stream), "memset");
        ck(mnv_fill_i32(visit_tracker.as<int32_t>(), 1, 1, stream), "fill");
        options.basis_minmax[0] = 0;
        options.basis_minmax[1] = std::max(t.data_format.basis_dim - 1, 0);
        can_reuse_results = false;
        prune_happened = false;
    }

    // replicas on every device of the group, built from the host arrays (VQ files are decoded on the host here)
    void set_group(N3Tree &t, long max_capacity) {
        if (group) {
            mnv_group_destroy(group);
            group = nullptr;
        }
        t.decode_vq_host();
        const int64_t cap = t.child.size(0);
        mnv_tree_desc d;
        std::memset(&d, 0, sizeof(d));
        d.N = t.N;
        d.data_dim = t.data_dim;
        d.format = t.data_format.format == DataFormat::SH ? MNV_FORMAT_SH : MNV_FORMAT_RGBA;
        d.basis_dim = t.data_format.basis_dim;
        d.capacity = cap;
        d.data = t.data.data_ptr();
        d.child = t.child.data_ptr();
        d.parent = t.parent.numel() >= cap ? t.parent.data_ptr() : nullptr;
        d.sample_counts = t.sample_counts.numel() >= cap * 8 ? t.sample_counts.data_ptr() : nullptr;
        for (int i = 0; i < 3; ++i) {
            d.scale[i] = t.scale.v[i];
            d.offset[i] = t.offset.v[i];
        }
        ck(mnv_group_create(&group, &d, max_capacity, devices.data(), (int) devices.size()), "group create");
        tree = &t;
        mnv_tree *t0 = nullptr;
        ck(mnv_group_tree(group, 0, &t0), "group tree");
        int64_t c = 0, m = 0;
        mnv_tree_capacity(t0, &c, &m);
        max_tree_capacity = m;
        t.capacity = (int) c;
        options.basis_minmax[0] = 0;
        options.basis_minmax[1] = std::max(t.data_format.basis_dim - 1, 0);
        can_reuse_results = false;
        prune_happened = false;
    }

    void load_model(const std::filesystem::path &path) {
        if (self.verbose) std::printf("Loading model from: %s\n", path.c_str());
        if (devices.size() > 1) {  // one copy of the sub-modules per device of the group (1.2 MB of bf16 each)
            group_models.clear();
            for (int dev : devices) {
                group_models.push_back(std::make_unique<ModelContainer>());
                group_models.back()->load(path.string(), dev);
            }
            options.need_viewdir = group_models[0]->need_viewdir;
            if (options.appearance_embedding == -1 && group_models[0]->need_appearance_embedding)
                options.appearance_embedding = 0;
            return;
        }
        model.load(path.string(), 0);
        options.need_viewdir = model.need_viewdir;
        if (options.appearance_embedding == -1 && model.need_appearance_embedding) options.appearance_embedding = 0;
        can_reuse_results = false;
        if (self.verbose) std::printf("Model loaded\n");
    }

    VolumeRenderer &self;
    Camera &camera;
    RenderOptions &options;
    N3Tree *tree = nullptr;
    ModelContainer model;
    std::vector<int> devices;  // set_devices(): more than one -> replica group
    mnv_group *group = nullptr;
    std::vector<std::unique_ptr<ModelContainer>> group_models;
    void *stream = nullptr;
    bool started = false;
    bool initial_resize = true;
    bool can_reuse_results = false;
    bool prune_happened = false;
    int64_t max_tree_capacity = 0;
    int buf_index = 0;
    bool interop = false;
    void *ca[4] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf frame, split_tracker, sample_tracker, visit_tracker;
    DevBuf offsets, z_vals, rows, cluster, values;  // guided sampling (CSR)
    DevBuf nodes, rand, rcluster, results;          // refinement
    int64_t guided_capacity = 0, guided_total = 0;
    uint64_t rng_calls = 0;
    void *frame_pinned = nullptr;
    size_t frame_pinned_bytes = 0;
    bool frame_host_valid = false;
};

VolumeRenderer::VolumeRenderer() : impl_(std::make_unique<Impl>(*this)) {}
VolumeRenderer::~VolumeRenderer() {}

void VolumeRenderer::render() { impl_->render(); }
void VolumeRenderer::set(N3Tree &tree, long max_tree_capacity) { impl_->set(tree, max_tree_capacity); }
void VolumeRenderer::load_model(const std::filesystem::path &model_path) { impl_->load_model(model_path); }
void VolumeRenderer::clear() { impl_->tree = nullptr; }
void VolumeRenderer::resize(int width, int height) { impl_->resize(width, height); }
const char *VolumeRenderer::get_backend() { return "CUDA (sm_100a, mnv_b200)"; }

void VolumeRenderer::set_interop_surfaces(void *const cuda_arrays[4]) {
    // the previous arrays may be gone (resize re-registers the renderbuffers): forget their surface objects
    if (impl_->tree && impl_->tree->device_tree) mnv_tree_release_surfaces(impl_->tree->device_tree);
    impl_->interop = cuda_arrays != nullptr;
    for (int i = 0; i < 4; ++i) impl_->ca[i] = cuda_arrays ? cuda_arrays[i] : nullptr;
    impl_->buf_index = 0;
}

void VolumeRenderer::set_devices(const std::vector<int> &devices) { impl_->devices = devices; }

const uint8_t *VolumeRenderer::frame_device() const { return impl_->frame.as<uint8_t>(); }

const uint8_t *VolumeRenderer::frame_host() {
    Impl &I = *impl_;
    if (!I.started || I.interop) return nullptr;
    if (!I.frame_host_valid) {
        ck(mnv_memcpy_d2h_async(I.frame_pinned, I.frame.p, (size_t) I.pixels() * 4, I.stream), "frame read-back");
        ck(mnv_stream_synchronize(I.stream), "sync");
        I.frame_host_valid = true;
    }
    return static_cast<const uint8_t *>(I.frame_pinned);
}

}  // namespace viewer
