#include "model.hpp"

#include <cstring>
#include <stdexcept>

#include "../../../include/mnv_b200.h"
#include "npz.hpp"

namespace viewer {
namespace {

long scalar_int(const npz::Array &a) {
    if (a.bytes.empty()) throw std::runtime_error("model container: empty scalar");
    switch (a.kind) {
        case 'b':
        case '?': return a.bytes[0] != 0;
        case 'i':
        case 'u':
            if (a.word_size == 8) return (long) *a.checked<int64_t>(1, "scalar");
            if (a.word_size == 4) return *a.checked<int32_t>(1, "scalar");
            if (a.word_size == 2) return *a.checked<int16_t>(1, "scalar");
            return *a.checked<int8_t>(1, "scalar");
        case 'f': return a.word_size == 8 ? (long) *a.checked<double>(1, "scalar") : (long) *a.checked<float>(1, "scalar");
    }
    throw std::runtime_error("model container: unsupported scalar type");
}

const npz::Array &need(const npz::Archive &z, const std::string &key) {
    auto it = z.find(key);
    if (it == z.end()) throw std::runtime_error("model container: key missing: " + key);
    return it->second;
}

const float *f32(const npz::Archive &z, const std::string &key, size_t n_expected) {
    const npz::Array &a = need(z, key);
    if (a.kind != 'f' || a.word_size != 4) throw std::runtime_error("model container: " + key + " must be float32");
    if (n_expected && a.num_vals() != n_expected)
        throw std::runtime_error("model container: " + key + " has " + std::to_string(a.num_vals()) +
                                 " values, expected " + std::to_string(n_expected));
    return a.checked<float>(a.num_vals(), key.c_str());
}

}  // namespace

ModelContainer::~ModelContainer() { release(); }

void ModelContainer::release() {
    if (device_model) mnv_model_destroy(device_model);
    device_model = nullptr;
    n_submodules = 0;
}

void ModelContainer::load(const std::string &path, int device) {
    release();
    const npz::Archive z = npz::load(path);
    {
        const npz::Array &g = need(z, "grid_dim");
        if (g.num_vals() != 2) throw std::runtime_error("model container: grid_dim must have 2 entries");
        if (g.kind != 'i' && g.kind != 'u') throw std::runtime_error("model container: grid_dim must be an integer array");
        for (int i = 0; i < 2; ++i)
            grid_dim[i] = g.word_size == 8 ? (int32_t) g.checked<int64_t>(2, "grid_dim")[i] : g.checked<int32_t>(2, "grid_dim")[i];
    }
    std::memcpy(min_position, f32(z, "min_position", 3), 12);
    std::memcpy(max_position, f32(z, "max_position", 3), 12);
    for (int i = 0; i < 3; ++i) range[i] = max_position[i] - min_position[i];
    {
        const npz::Array &c = need(z, "centroids");
        if (c.shape.size() != 2 || c.shape[1] != 3) throw std::runtime_error("model container: centroids must be [M,3]");
        centroids.assign(f32(z, "centroids", 0), f32(z, "centroids", 0) + c.num_vals());
        n_submodules = (int) c.shape[0];
    }
    need_viewdir = scalar_int(need(z, "need_viewdir")) != 0;
    need_appearance_embedding = scalar_int(need(z, "need_appearance_embedding")) != 0;

    std::vector<mnv_mlp_desc> descs((size_t) n_submodules);
    for (int s = 0; s < n_submodules; ++s) {
        const std::string p = "sub_module_" + std::to_string(s) + "/";
        mnv_mlp_desc &d = descs[(size_t) s];
        std::memset(&d, 0, sizeof(d));
        const npz::Array &cfg = need(z, p + "config");
        if (cfg.num_vals() < 5 || cfg.word_size != 4) throw std::runtime_error("model container: bad " + p + "config");
        const int32_t *c = cfg.checked<int32_t>(5, "config");
        d.n_trunk_layers = c[0];
        d.skip_layer = c[1];
        d.pe_xyz_freqs = c[2];
        d.pe_dir_freqs = c[3];
        d.sigma_activation = c[4];
        d.need_viewdir = need_viewdir ? 1 : 0;
        if (d.n_trunk_layers < 1 || d.n_trunk_layers > 12) throw std::runtime_error("model container: bad trunk depth");
        const npz::Array &w0 = need(z, p + "trunk_w_0");
        if (w0.shape.size() != 2) throw std::runtime_error("model container: trunk_w_0 must be 2-D");
        d.width = (int) w0.shape[0];
        const int pe = 3 + 6 * d.pe_xyz_freqs;
        for (int l = 0; l < d.n_trunk_layers; ++l) {
            const int in = l == 0 ? pe : (l == d.skip_layer ? pe + d.width : d.width);
            d.trunk_w[l] = f32(z, p + "trunk_w_" + std::to_string(l), (size_t) d.width * in);
            d.trunk_b[l] = f32(z, p + "trunk_b_" + std::to_string(l), (size_t) d.width);
        }
        d.sigma_w = f32(z, p + "sigma_w", (size_t) d.width);
        d.sigma_b = f32(z, p + "sigma_b", 1);
        d.final_w = f32(z, p + "final_w", (size_t) d.width * d.width);
        d.final_b = f32(z, p + "final_b", (size_t) d.width);
        if (z.count(p + "embedding")) {
            const npz::Array &e = need(z, p + "embedding");
            if (e.shape.size() != 2) throw std::runtime_error("model container: embedding must be 2-D");
            d.n_appearance = (int) e.shape[0];
            d.appearance_dim = (int) e.shape[1];
            d.embedding = f32(z, p + "embedding", 0);
        }
        const npz::Array &h1 = need(z, p + "head1_w");
        if (h1.shape.size() != 2) throw std::runtime_error("model container: head1_w must be 2-D");
        d.head_width = (int) h1.shape[0];
        const size_t h1_in = (size_t) d.width + (need_viewdir ? 3 + 6 * d.pe_dir_freqs : 0) + d.appearance_dim;
        d.head1_w = f32(z, p + "head1_w", (size_t) d.head_width * h1_in);
        d.head1_b = f32(z, p + "head1_b", (size_t) d.head_width);
        const npz::Array &h2 = need(z, p + "head2_w");
        if (h2.shape.size() != 2 || (int) h2.shape[1] != d.head_width)
            throw std::runtime_error("model container: head2_w must be [out][head_width]");
        d.out_rgb_dim = (int) h2.shape[0];
        d.head2_w = f32(z, p + "head2_w", 0);
        d.head2_b = f32(z, p + "head2_b", (size_t) d.out_rgb_dim);
    }
    if (mnv_model_create(&device_model, n_submodules, descs.data(), grid_dim, min_position, max_position,
                         device) != MNV_OK) {
        device_model = nullptr;
        throw std::runtime_error(std::string("load_model: ") + mnv_last_error());
    }
    int n = 0;
    mnv_model_info(device_model, &n, &in_dim, &out_dim, &flops_per_row);
}

}  // namespace viewer
