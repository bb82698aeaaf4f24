// Mega-NeRF sub-module container on disk — stands in for the TorchScript archive the reference
// loads with torch::jit::load (src/renderer/cuda_renderer.cpp:518-543).  Same attributes, flat
// tensors, stored as an .npz (written by mega_nerf_viewer_b200.save_model_container or
// tools/export_model.py from a Mega-NeRF checkpoint / TorchScript container):
//
//   grid_dim i32|i64 [2]      min_position f32 [3]     max_position f32 [3]
//   centroids f32 [M][3]      need_viewdir (any int/bool scalar)
//   need_appearance_embedding (any int/bool scalar)
//   sub_module_<i>/config  i32 [5] = n_trunk_layers, skip_layer, pe_xyz_freqs, pe_dir_freqs,
//                                    sigma_activation (0 ReLU, 1 softplus)
//   sub_module_<i>/trunk_w_<l> f32 [256][in]   trunk_b_<l> f32 [256]
//   sub_module_<i>/sigma_w f32 [1][256]  sigma_b f32 [1]
//   sub_module_<i>/final_w f32 [256][256] final_b f32 [256]
//   sub_module_<i>/embedding f32 [n_app][app_dim]          (optional)
//   sub_module_<i>/head1_w f32 [128][256 (+27) (+app_dim)]  head1_b f32 [128]
//   sub_module_<i>/head2_w f32 [out][128]  head2_b f32 [out]
#pragma once

#include <cstdint>
#include <string>
#include <vector>

struct mnv_model;

namespace viewer {

struct ModelContainer {
    int32_t grid_dim[2] = {1, 1};
    float min_position[3] = {0, 0, 0};
    float max_position[3] = {1, 1, 1};
    float range[3] = {1, 1, 1};  // max_position - min_position (cuda_renderer.cpp:527)
    std::vector<float> centroids;
    bool need_viewdir = false;
    bool need_appearance_embedding = false;
    int n_submodules = 0;
    int in_dim = 0, out_dim = 0;
    double flops_per_row = 0;
    mnv_model *device_model = nullptr;

    ModelContainer() = default;
    ModelContainer(const ModelContainer &) = delete;
    ModelContainer &operator=(const ModelContainer &) = delete;
    ~ModelContainer();

    // Throws std::runtime_error on a missing file, schema violation or device failure.
    void load(const std::string &path, int device = 0);
    void release();
};

}  // namespace viewer
