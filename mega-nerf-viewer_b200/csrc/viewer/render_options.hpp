// viewer::RenderOptions — field-for-field the reference's include/render_options.hpp:9-56
// (same names, order, types and defaults), so that code written against the reference
// compiles unchanged and the struct can be handed to the C-ABI as mnv_render_options.
#pragma once

#define VIEWER_GLOBAL_BASIS_MAX 25

namespace viewer {

struct RenderOptions {
    float step_size = 1e-4f;
    float sigma_thresh = 1e-2f;
    float stop_thresh = 1e-2f;
    float background_brightness = 1.f;
    float render_bbox[6] = {0.f, 0.f, 0.f, 1.f, 1.f, 1.f};
    int basis_minmax[2] = {0, VIEWER_GLOBAL_BASIS_MAX - 1};
    float rot_dirs[3] = {0.f, 0.f, 0.f};
    bool show_grid = false;
    int grid_max_depth = 4;
    bool render_depth = false;
    bool use_splitting = false;
    bool use_guided_sampling = false;
    int max_depth = 16;
    int samples_per_corner = 8;
    int split_batch_size = 4192;
    int nerf_batch_size = 1024;
    int max_sample_count = 256;
    bool need_viewdir = false;
    int appearance_embedding = -1;
    int max_guided_samples = 128;
};

}  // namespace viewer
