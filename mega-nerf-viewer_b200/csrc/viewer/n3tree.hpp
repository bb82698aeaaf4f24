// viewer::N3Tree — the public surface of the reference's include/n3tree/n3tree.hpp:17-69:
// open(path) from the svox .npz schema, move_to_device(max_capacity, need_parent,
// need_sample_counts), gen_wireframe, pack_index / unpack_index, and the public members
// N, data_dim, data_format, scale, offset, data, child, parent, sample_counts, capacity.
// The reference holds torch::Tensors; here the members are plain host arrays (HostArray)
// with the operations main.cpp uses on them (slice(0,a,b), [i][j] = v, size(0)).
// The device copy is the SoA tree behind the C-ABI (mnv_tree).
#pragma once

#include <cstdint>
#include <string>
#include <tuple>
#include <vector>

#include "data_format.hpp"

struct mnv_tree;

namespace viewer {

template <typename T>
struct HostArray {
    std::vector<int64_t> shape;
    std::vector<T> v;

    int64_t size(int dim) const { return shape.at((size_t) dim); }
    int64_t numel() const { return (int64_t) v.size(); }
    int64_t row_elems() const { return shape.empty() || shape[0] == 0 ? 0 : numel() / shape[0]; }
    T *data_ptr() { return v.data(); }
    const T *data_ptr() const { return v.data(); }

    // rows [a, b) along dim 0 (the only slicing main.cpp performs, main.cpp:529-537)
    HostArray slice(int dim, int64_t a, int64_t b) const {
        HostArray out;
        if (dim != 0) return *this;
        out.shape = shape;
        out.shape[0] = b - a;
        out.v.assign(v.begin() + a * row_elems(), v.begin() + b * row_elems());
        return out;
    }
    struct Row {
        T *p;
        T &operator[](int64_t j) { return p[j]; }
    };
    Row operator[](int64_t i) { return Row{v.data() + i * row_elems()}; }
};

struct N3Tree {
    N3Tree();
    explicit N3Tree(const std::string &path);
    ~N3Tree();
    N3Tree(const N3Tree &) = delete;
    N3Tree &operator=(const N3Tree &) = delete;

    void open(const std::string &path);
    void move_to_device(long max_capacity, bool need_parent, bool need_sample_counts);
    std::vector<float> gen_wireframe(int max_depth = 100000) const;

    int N = 0;
    int data_dim = 0;
    DataFormat data_format;
    HostArray<float> scale, offset;        // [3]
    int64_t pack_index(int nd, int i, int j, int k);
    std::tuple<int, int, int, int> unpack_index(int64_t packed);
    HostArray<uint16_t> data;              // fp16 bits [capacity][8][data_dim]
    HostArray<int32_t> child;              // [capacity][8] relative offsets
    HostArray<int32_t> parent;             // [capacity] packed parent slot
    HostArray<int16_t> sample_counts;      // [capacity][8]
    int capacity = 0;

    // B200-native additions
    mnv_tree *device_tree = nullptr;       // set by move_to_device
    void sync_capacity();                  // refresh `capacity` after device-side refinement
    void download();                       // refresh the host arrays from the device tree
    // Write the tree in the svox schema open() reads (after download(): the refined tree).  The reference has no
    // writer; `parent_depth` column 1 is recomputed from the links.
    void save(const std::string &path) const;
    // VQ-compressed files (quant_colors / quant_map / data_retained / sigma, n3tree.cpp:109-175) stay compressed on
    // the host: move_to_device uploads them as they are and the GPU decodes (mnv_tree_create_vq).  `data` is empty
    // until decode_vq_host() (or download()) fills it for consumers that read leaf payloads on the host.
    bool is_vq_compressed() const;
    void decode_vq_host();

   private:
    int N2_ = 0, N3_ = 0;
    int vq_n_quant = 0, vq_n_retain = 0;
    std::vector<uint16_t> vq_book, vq_map, vq_retained, vq_sigma;
};

}  // namespace viewer
