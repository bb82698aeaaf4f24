// viewer::VolumeRenderer — the public surface of the reference's include/renderer/renderer.hpp:9-40
// (render / set / load_model / clear / resize / get_backend + the public camera and options),
// driving the B200 kernels through the C-ABI of include/mnv_b200.h.  No LibTorch, no GL:
// the frame sequencing of Impl::render (src/renderer/cuda_renderer.cpp:68-163) is kept, the
// presentation target is either the caller's GL-interop surfaces (set_interop_surfaces, what
// cudaGraphicsSubResourceGetMappedArray returns, :447-455) or — the headless mode — a device
// RGBA8 frame that frame_host() reads back.
#pragma once

#include <cstdint>
#include <filesystem>
#include <memory>
#include <vector>

#include "camera.hpp"
#include "n3tree.hpp"
#include "render_options.hpp"

namespace viewer {

struct VolumeRenderer {
    explicit VolumeRenderer();
    ~VolumeRenderer();

    // Render the currently set tree (one frame; refinement / pruning as the options say)
    void render();

    // Set volumetric data to render.  Like the reference, keeps a pointer to `tree`
    // (the caller owns it and must outlive the renderer) and moves it to the device.
    void set(N3Tree &tree, long max_tree_capacity);

    // Load the Mega-NeRF sub-module container (see model.hpp for the file format)
    void load_model(const std::filesystem::path &model_path);

    // Clear the volumetric data
    void clear();

    // Resize the buffer
    void resize(int width, int height);

    // Get name identifying the renderer backend used e.g. CUDA
    const char *get_backend();

    // Camera instance
    Camera camera;

    // Rendering options
    RenderOptions options;

    // ---- B200-native additions (headless / interop plumbing) ----------------------------
    // GL interop: the four cudaArray_t of the reference's double-buffered colour (RGBA8) and
    // fake-depth (R32F) renderbuffers, in the order of Impl::ca (cuda_renderer.cpp:447-455).
    // With surfaces set, render() writes there with offscreen=false compositing.
    void set_interop_surfaces(void *const cuda_arrays[4]);
    // Multi-GPU (image tiles with the tree replicated): call before set().  The tree is replicated
    // on every listed device (mnv_group); render() marches interleaved 8-row bands on all of them and gathers the
    // bands over NVLink peer copies into devices[0] — the headless frame or the GL-interop colour surface — and, with
    // options.use_splitting, runs the refinement step across the group (votes and fp16 payloads exchanged by peer
    // copies, MLP rows sharded; load_model() places a copy of the sub-modules on every device).  Guided sampling
    // and pruning stay single-GPU features: with a group they are not available (render() throws / never prunes).
    void set_devices(const std::vector<int> &devices);
    // Headless: the frame render() produced, read back to host memory ([height][width][4]).
    const uint8_t *frame_host();
    // Device pointer of that frame (RGBA8 linear), valid until the next resize.
    const uint8_t *frame_device() const;
    // Statistics of the last render() call
    struct FrameInfo {
        int64_t guided_rows = 0;      // MLP rows evaluated for guided sampling (0: reused / off)
        int split_candidates = -1;    // "Split candidates: n" (-1: refinement did not run)
        int added = 0;                // nodes split this frame
        int resampled = 0;            // leaves that received more samples
        int64_t pruned = -1;          // nodes reclaimed (-1: no prune pass)
        int64_t capacity = 0;
    };
    FrameInfo last_frame;
    // Seed of the counter-based generator standing in for torch::rand (cuda_renderer.cpp:250)
    uint64_t rng_seed = 0x5eed;
    // Print the reference's progress lines ("Split candidates: ...") to stdout
    bool verbose = false;

   private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

}  // namespace viewer
