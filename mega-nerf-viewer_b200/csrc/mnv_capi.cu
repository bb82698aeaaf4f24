// extern "C" surface of libmnv_b200.so (include/mnv_b200.h).
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "mnv_internal.cuh"

using namespace mnv;

struct mnv_model {
    std::vector<MlpModel *> subs;
    int32_t grid_dim[2] = {1, 1};
    float min_position[3] = {0, 0, 0}, max_position[3] = {1, 1, 1};
    int device = 0;
};

struct mnv_tree {
    DeviceTree t;
    std::map<void *, cudaSurfaceObject_t> surfaces;  // cudaArray_t -> surface, created once
};

namespace mnv {
DeviceTree &device_tree_of(mnv_tree *h) { return h->t; }  // for mnv_group.cu
}  // namespace mnv

namespace {

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MNV_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (have %d)", device, n);
        return MNV_ERR_INVALID;
    }
    return MNV_OK;
}

// The reference creates a cudaSurfaceObject_t per launch and never destroys it
// (src/cuda/renderer_kernel.cu:377-385,410-428); here one per array, cached.
int surface_for(mnv_tree *tree, void *arr, cudaSurfaceObject_t *out) {
    *out = 0;
    if (!arr) return MNV_OK;
    auto it = tree->surfaces.find(arr);
    if (it != tree->surfaces.end()) {
        *out = it->second;
        return MNV_OK;
    }
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = static_cast<cudaArray_t>(arr);
    cudaSurfaceObject_t s = 0;
    MNV_CUDA(cudaCreateSurfaceObject(&s, &rd));
    tree->surfaces[arr] = s;
    *out = s;
    return MNV_OK;
}

// Depth of the deepest leaf (root's children are depth 1), -1 on malformed links.
int host_max_leaf_depth(const int32_t *child, int64_t cap) {
    std::vector<int8_t> depth((size_t) cap, -1);  // level of each node, root = 0
    depth[0] = 0;
    int maxd = 1;
    bool pending = true;
    for (int pass = 0; pass < 64 && pending; ++pass) {  // one pass when children follow parents
        pending = false;
        for (int64_t n = 0; n < cap; ++n) {
            if (depth[n] < 0) {
                pending = true;
                continue;
            }
            for (int c = 0; c < 8; ++c) {
                const int32_t rel = child[n * 8 + c];
                if (rel == 0) continue;
                const int64_t m = n + rel;
                if (m <= 0 || m >= cap) return -1;
                if (depth[m] < 0) {
                    if (depth[n] >= 100) return -1;
                    depth[m] = (int8_t) (depth[n] + 1);
                    if (m < n) pending = true;
                    if (depth[m] + 1 > maxd) maxd = depth[m] + 1;
                }
            }
        }
    }
    return maxd;
}

}  // namespace

extern "C" {

const char *mnv_version(void) { return "mnv_b200 0.1 (sm_100a)"; }
const char *mnv_last_error(void) { return last_error_cstr(); }

int mnv_device_count(int *count) {
    if (!count) return MNV_ERR_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return MNV_OK;
}

int mnv_malloc(void **ptr_dev, size_t bytes, int device) {
    if (!ptr_dev) return MNV_ERR_INVALID;
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaSetDevice(device));
    MNV_CUDA(cudaMalloc(ptr_dev, bytes ? bytes : 1));
    return MNV_OK;
}
int mnv_free(void *ptr_dev) {
    if (ptr_dev) MNV_CUDA(cudaFree(ptr_dev));
    return MNV_OK;
}
int mnv_memset(void *ptr_dev, int value, size_t bytes, void *stream) {
    MNV_CUDA(cudaMemsetAsync(ptr_dev, value, bytes, static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}
int mnv_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream) {
    MNV_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}
int mnv_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    MNV_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    MNV_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}
int mnv_fill_f32(float *ptr_dev, float value, int64_t n, void *stream) {
    return fill_f32(ptr_dev, value, n, static_cast<cudaStream_t>(stream));
}
int mnv_fill_i32(int32_t *ptr_dev, int32_t value, int64_t n, void *stream) {
    return fill_i32(ptr_dev, value, n, static_cast<cudaStream_t>(stream));
}
int mnv_fill_uniform(float *ptr_dev, int64_t n, uint64_t seed, void *stream) {
    return fill_uniform(ptr_dev, n, seed, static_cast<cudaStream_t>(stream));
}

int mnv_malloc_host(void **ptr_host, size_t bytes) {
    if (!ptr_host) return MNV_ERR_INVALID;
    int rc = check_device(0);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaHostAlloc(ptr_host, bytes ? bytes : 1, cudaHostAllocDefault));
    return MNV_OK;
}
int mnv_free_host(void *ptr_host) {
    if (ptr_host) MNV_CUDA(cudaFreeHost(ptr_host));
    return MNV_OK;
}
int mnv_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    MNV_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}
int mnv_stream_create(void **stream, int device) {
    if (!stream) return MNV_ERR_INVALID;
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    MNV_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return MNV_OK;
}
int mnv_stream_destroy(void *stream) {
    if (stream) MNV_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}
int mnv_stream_synchronize(void *stream) {
    MNV_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return MNV_OK;
}

int mnv_array_create(void **array, int width, int height, int kind, int device) {
    if (!array || width <= 0 || height <= 0 || (kind != 0 && kind != 1)) return MNV_ERR_INVALID;
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaSetDevice(device));
    const cudaChannelFormatDesc desc = kind == 0 ? cudaCreateChannelDesc<uchar4>() : cudaCreateChannelDesc<float>();
    cudaArray_t a = nullptr;
    MNV_CUDA(cudaMallocArray(&a, &desc, (size_t) width, (size_t) height, cudaArraySurfaceLoadStore));
    *array = a;
    return MNV_OK;
}
int mnv_array_destroy(void *array) {
    if (array) MNV_CUDA(cudaFreeArray(static_cast<cudaArray_t>(array)));
    return MNV_OK;
}
int mnv_array_upload(void *array, const void *src_host, size_t row_bytes, int height) {
    if (!array || !src_host) return MNV_ERR_INVALID;
    MNV_CUDA(cudaMemcpy2DToArray(static_cast<cudaArray_t>(array), 0, 0, src_host, row_bytes, row_bytes, (size_t) height,
                                 cudaMemcpyHostToDevice));
    return MNV_OK;
}
int mnv_array_download(void *dst_host, void *array, size_t row_bytes, int height) {
    if (!array || !dst_host) return MNV_ERR_INVALID;
    MNV_CUDA(cudaMemcpy2DFromArray(dst_host, row_bytes, static_cast<cudaArray_t>(array), 0, 0, row_bytes,
                                   (size_t) height, cudaMemcpyDeviceToHost));
    return MNV_OK;
}

void mnv_render_options_default(mnv_render_options *o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->step_size = 1e-4f;
    o->sigma_thresh = 1e-2f;
    o->stop_thresh = 1e-2f;
    o->background_brightness = 1.f;
    o->render_bbox[3] = o->render_bbox[4] = o->render_bbox[5] = 1.f;
    o->basis_minmax[0] = 0;
    o->basis_minmax[1] = MNV_GLOBAL_BASIS_MAX - 1;
    o->grid_max_depth = 4;
    o->max_depth = 16;
    o->samples_per_corner = 8;
    o->split_batch_size = 4192;
    o->nerf_batch_size = 1024;
    o->max_sample_count = 256;
    o->appearance_embedding = -1;
    o->max_guided_samples = 128;
}

static int tree_create_impl(mnv_tree **out, const mnv_tree_desc *d, const mnv_vq_desc *vq, int64_t max_capacity,
                            int device);

int mnv_tree_create(mnv_tree **out, const mnv_tree_desc *d, int64_t max_capacity, int device) {
    return tree_create_impl(out, d, nullptr, max_capacity, device);
}

int mnv_tree_create_vq(mnv_tree **out, const mnv_tree_desc *d, const mnv_vq_desc *vq, int64_t max_capacity, int device) {
    if (!vq || !d) return MNV_ERR_INVALID;
    if (vq->n_quant < 0 || vq->n_retain < 0 || vq->n_quant + vq->n_retain < 1 || !vq->sigma ||
        (vq->n_quant > 0 && (!vq->quant_colors || !vq->quant_map)) || (vq->n_retain > 0 && !vq->data_retained)) {
        set_error("VQ tree: missing arrays");
        return MNV_ERR_INVALID;
    }
    if (d->format != MNV_FORMAT_SH || d->basis_dim != vq->n_quant + vq->n_retain) {
        set_error("VQ tree: SH format with basis_dim == n_quant + n_retain expected (got %d, %d + %d)", d->basis_dim,
                  vq->n_quant, vq->n_retain);
        return MNV_ERR_FORMAT;
    }
    return tree_create_impl(out, d, vq, max_capacity, device);
}

static int tree_create_impl(mnv_tree **out, const mnv_tree_desc *d, const mnv_vq_desc *vq, int64_t max_capacity,
                            int device) {
    if (!out || !d) return MNV_ERR_INVALID;
    *out = nullptr;
    if (d->N != 2) {
        set_error("only N == 2 octrees are supported (got N = %d)", d->N);
        return MNV_ERR_INVALID;
    }
    if (d->capacity <= 0 || (!d->data && !vq) || !d->child || d->data_dim <= 0) {
        set_error("empty tree / missing arrays");
        return MNV_ERR_INVALID;
    }
    if (d->format == MNV_FORMAT_SH) {
        const int b = d->basis_dim;
        if (!(b == 1 || b == 4 || b == 9 || b == 16 || b == 25) || d->data_dim != 3 * b + 1) {
            set_error("SH tree needs basis_dim in {1,4,9,16,25} and data_dim == 3*basis_dim+1 "
                      "(got %d, %d)", b, d->data_dim);
            return MNV_ERR_FORMAT;
        }
    } else if (d->format == MNV_FORMAT_RGBA) {
        if (d->data_dim != 4) {
            set_error("RGBA tree needs data_dim == 4 (got %d)", d->data_dim);
            return MNV_ERR_FORMAT;
        }
    } else {
        set_error("unknown data format %d", d->format);
        return MNV_ERR_FORMAT;
    }
    if (max_capacity < d->capacity) max_capacity = d->capacity;
    if (max_capacity >= (1ll << 28)) {
        set_error("max_capacity %lld too large (limit 2^28 nodes)", (long long) max_capacity);
        return MNV_ERR_INVALID;
    }
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaSetDevice(device));
    mnv_tree *h = new (std::nothrow) mnv_tree();
    if (!h) return MNV_ERR_OOM;
    DeviceTree &t = h->t;
    t.N = d->N;
    t.data_dim = d->data_dim;
    t.format = d->format;
    t.basis_dim = d->format == MNV_FORMAT_SH ? d->basis_dim : -1;
    t.capacity = d->capacity;
    t.max_capacity = max_capacity;
    t.device = device;
    for (int i = 0; i < 3; ++i) {
        t.scale[i] = d->scale[i];
        t.offset[i] = d->offset[i];
    }
    t.max_leaf_depth = host_max_leaf_depth(d->child, d->capacity);
    if (t.max_leaf_depth < 0) {
        delete h;
        set_error("child links do not form a tree rooted at node 0");
        return MNV_ERR_FORMAT;
    }
    if (t.max_leaf_depth > 23) {
        delete h;
        set_error("tree depth %d exceeds the supported maximum of 23", t.max_leaf_depth);
        return MNV_ERR_FORMAT;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    {
        // anchor grid level: 8 (2^24 entries, 128 MiB) unless the tree is shallower; MNV_ANCHOR_LEVEL=0..8 overrides
        // (measured on the 1080p bench frame: 7 -> 1.036 ms, 8 -> 1.004 ms; 4K Mill-19-scale: 3.46 -> 3.25 ms)
        int a = 8;
        if (const char *env = std::getenv("MNV_ANCHOR_LEVEL")) a = std::atoi(env);
        t.anchor_level = std::max(0, std::min(std::min(a, 8), std::max(t.max_leaf_depth, 1)));
    }
    rc = build_device_tree(t, *d, vq);
    if (rc != MNV_OK) {
        mnv_tree_destroy(h);
        return rc;
    }
    *out = h;
    return MNV_OK;
}

int mnv_tree_destroy(mnv_tree *h) {
    if (!h) return MNV_OK;
    cudaSetDevice(h->t.device);
    for (auto &kv : h->surfaces) cudaDestroySurfaceObject(kv.second);
    DeviceTree &t = h->t;
    cudaFree(t.cell);
    cudaFree(t.payload);
    cudaFree(t.parent);
    cudaFree(t.sample_counts);
    cudaFree(t.count_dev);
    cudaFree(t.scan_tmp);
    cudaFree(t.frame_dev);
    cudaFree(t.split_dev);
    cudaFree(t.sample_dev);
    cudaFree(t.stats_dev);
    cudaFree(t.partial_table_dev);
    cudaFree(t.anchor);
    cudaFree(t.max_depth_dev);
    if (t.max_depth_host) cudaFreeHost(t.max_depth_host);
    if (t.depth_event) cudaEventDestroy(t.depth_event);
    if (t.stream) cudaStreamDestroy(t.stream);
    delete h;
    return MNV_OK;
}

int mnv_tree_capacity(const mnv_tree *h, int64_t *capacity, int64_t *max_capacity) {
    if (!h) return MNV_ERR_INVALID;
    if (capacity) *capacity = h->t.capacity;
    if (max_capacity) *max_capacity = h->t.max_capacity;
    return MNV_OK;
}

int mnv_tree_device_bytes(const mnv_tree *h, uint64_t *bytes) {
    if (!h || !bytes) return MNV_ERR_INVALID;
    const DeviceTree &t = h->t;
    *bytes = (uint64_t) t.max_capacity * (8ull * (4 + 2 + 16ull * t.rec_u4) + 4);
    return MNV_OK;
}

int mnv_tree_download(const mnv_tree *h, int64_t first, int64_t count, uint16_t *data,
                      int32_t *child, int32_t *parent, int16_t *sample_counts) {
    if (!h) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return download_device_tree(h->t, first, count, data, child, parent, sample_counts);
}

int mnv_tree_release_surfaces(mnv_tree *h) {
    if (!h) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    MNV_CUDA(cudaDeviceSynchronize());  // no launch may still be using them
    for (auto &kv : h->surfaces) cudaDestroySurfaceObject(kv.second);
    h->surfaces.clear();
    return MNV_OK;
}

int mnv_tree_set_tile_order(mnv_tree *h, const int32_t *order_dev, int n) {
    if (!h || n < 0) return MNV_ERR_INVALID;
    h->t.tile_order_dev = order_dev;
    h->t.tile_order_n = order_dev ? n : 0;
    return MNV_OK;
}

int mnv_query_points(const mnv_tree *h, const float *xyz_dev, int64_t n, int32_t *out_dev,
                     void *stream) {
    if (!h || (n > 0 && (!xyz_dev || !out_dev))) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return launch_query_points(h->t, xyz_dev, n, out_dev, static_cast<cudaStream_t>(stream));
}

int mnv_render_voxels(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                      void *image_arr, void *depth_arr, uint8_t *image_linear_dev,
                      float *to_split_dev, float *to_sample_dev, int32_t *visited_dev,
                      bool track_visit, bool offscreen, void *stream) {
    if (!h || !cam || !opt) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    RenderTargets tg;
    tg.image_linear = image_linear_dev;
    int rc = surface_for(h, image_arr, &tg.image_surf);
    if (rc != MNV_OK) return rc;
    if (!offscreen) {
        if (!image_arr || !depth_arr) {
            set_error("offscreen == false needs image and depth surfaces");
            return MNV_ERR_INVALID;
        }
        rc = surface_for(h, depth_arr, &tg.depth_surf);
        if (rc != MNV_OK) return rc;
    }
    tg.to_split = to_split_dev;
    tg.to_sample = to_sample_dev;
    tg.visited = visited_dev;
    tg.track_visit = track_visit;
    tg.offscreen = offscreen;
    return launch_render_voxels(h->t, *cam, *opt, tg, static_cast<cudaStream_t>(stream));
}

int mnv_render_voxels_tiles(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                            uint8_t *image_linear_dev, float *to_split_dev, float *to_sample_dev,
                            int tile_w, int tile_h, int tile_mod, int tile_rem, void *stream) {
    if (!h || !cam || !opt || !image_linear_dev) return MNV_ERR_INVALID;
    if (tile_mod < 1 || tile_rem < 0 || tile_rem >= tile_mod) {
        set_error("bad tile partition %d / %d", tile_rem, tile_mod);
        return MNV_ERR_INVALID;
    }
    MNV_CUDA(cudaSetDevice(h->t.device));
    RenderTargets tg;
    tg.image_linear = image_linear_dev;
    tg.to_split = to_split_dev;
    tg.to_sample = to_sample_dev;
    tg.tile_w = tile_w;
    tg.tile_h = tile_h;
    tg.tile_mod = tile_mod;
    tg.tile_rem = tile_rem;
    return launch_render_voxels(h->t, *cam, *opt, tg, static_cast<cudaStream_t>(stream));
}

int mnv_render_voxels_logged(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                             uint8_t *image_linear_dev, uint64_t *visit_hash_dev,
                             int32_t *visit_count_dev, int32_t *shaded_count_dev,
                             int32_t *visit_log_dev, int log_cap, void *stream) {
    if (!h || !cam || !opt || !image_linear_dev) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    RenderTargets tg;
    tg.image_linear = image_linear_dev;
    tg.visit_hash = reinterpret_cast<unsigned long long *>(visit_hash_dev);
    tg.visit_count = visit_count_dev;
    tg.shaded_count = shaded_count_dev;
    tg.visit_log = visit_log_dev;
    tg.log_cap = visit_log_dev ? log_cap : 0;
    if (!tg.visit_hash && !tg.visit_count && !tg.shaded_count && !tg.visit_log) {
        set_error("no log output given");
        return MNV_ERR_INVALID;
    }
    return launch_render_voxels(h->t, *cam, *opt, tg, static_cast<cudaStream_t>(stream));
}

static int frame_host_impl(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                           uint8_t *rgba_host, int band_rows, int band_mod, int band_rem,
                           mnv_frame_stats *stats) {
    if (!h || !cam || !opt || !rgba_host) return MNV_ERR_INVALID;
    DeviceTree &t = h->t;
    MNV_CUDA(cudaSetDevice(t.device));
    const size_t bytes = (size_t) cam->width * cam->height * 4;
    if (bytes == 0) return MNV_ERR_INVALID;
    if (t.frame_bytes < bytes) {
        cudaFree(t.frame_dev);
        cudaFree(t.split_dev);
        cudaFree(t.sample_dev);
        t.frame_dev = nullptr;
        t.split_dev = t.sample_dev = nullptr;
        t.frame_bytes = 0;
        MNV_CUDA(cudaMalloc(&t.frame_dev, bytes));
        t.frame_bytes = bytes;
    }
    RenderTargets tg;
    tg.image_linear = t.frame_dev;
    if (opt->use_splitting) {
        if (!t.split_dev) {
            MNV_CUDA(cudaMalloc(&t.split_dev, bytes * 3));
            MNV_CUDA(cudaMalloc(&t.sample_dev, bytes * 3));
        }
        tg.to_split = t.split_dev;
        tg.to_sample = t.sample_dev;
    }
    if (band_mod > 1) {
        if (band_rows < 8 || band_rows % 8 || band_rem < 0 || band_rem >= band_mod) {
            set_error("bad band partition rows=%d %d/%d", band_rows, band_rem, band_mod);
            return MNV_ERR_INVALID;
        }
        tg.tile_w = ((cam->width + 15) / 16) * 16;
        tg.tile_h = band_rows;
        tg.tile_mod = band_mod;
        tg.tile_rem = band_rem;
    }
    if (stats) {
        if (!t.stats_dev) MNV_CUDA(cudaMalloc(&t.stats_dev, 4 * sizeof(unsigned long long)));
        MNV_CUDA(cudaMemsetAsync(t.stats_dev, 0, 4 * sizeof(unsigned long long), t.stream));
        tg.frame_stats = t.stats_dev;
    }
    int rc = launch_render_voxels(t, *cam, *opt, tg, t.stream);
    if (rc != MNV_OK) return rc;
    if (band_mod > 1) {
        // bands owned by this rank: contiguous band_bytes each, band_mod*band_bytes apart
        const size_t row_bytes = (size_t) cam->width * 4;
        const size_t band_bytes = row_bytes * band_rows;
        const int n_bands = (cam->height + band_rows - 1) / band_rows;
        const int full = cam->height / band_rows;  // complete bands
        int mine_full = 0;
        for (int b = band_rem; b < full; b += band_mod) ++mine_full;
        const size_t off = (size_t) band_rem * band_bytes;
        if (mine_full > 0)
            MNV_CUDA(cudaMemcpy2DAsync(rgba_host + off, band_bytes * band_mod, t.frame_dev + off,
                                       band_bytes * band_mod, band_bytes, mine_full,
                                       cudaMemcpyDeviceToHost, t.stream));
        if (n_bands > full && (n_bands - 1) % band_mod == band_rem) {  // ragged last band
            const size_t o2 = (size_t) full * band_bytes;
            MNV_CUDA(cudaMemcpyAsync(rgba_host + o2, t.frame_dev + o2, bytes - o2,
                                     cudaMemcpyDeviceToHost, t.stream));
        }
    } else {
        MNV_CUDA(cudaMemcpyAsync(rgba_host, t.frame_dev, bytes, cudaMemcpyDeviceToHost, t.stream));
    }
    if (stats) {
        unsigned long long hs[4];
        MNV_CUDA(cudaMemcpyAsync(hs, t.stats_dev, sizeof(hs), cudaMemcpyDeviceToHost, t.stream));
        MNV_CUDA(cudaStreamSynchronize(t.stream));
        stats->rays = hs[0];
        stats->visits = hs[1];
        stats->shaded_visits = hs[2];
        stats->rays_hit = hs[3];
    } else {
        MNV_CUDA(cudaStreamSynchronize(t.stream));
    }
    return MNV_OK;
}

int mnv_render_frame_host(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                          uint8_t *rgba_host, mnv_frame_stats *stats) {
    return frame_host_impl(h, cam, opt, rgba_host, 0, 1, 0, stats);
}

int mnv_render_frame_host_bands(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                                uint8_t *rgba_host, int band_rows, int band_mod, int band_rem,
                                mnv_frame_stats *stats) {
    if (band_mod < 1) return MNV_ERR_INVALID;
    return frame_host_impl(h, cam, opt, rgba_host, band_rows, band_mod, band_rem, stats);
}

int mnv_guided_samples(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                       void *depth_arr, bool offscreen, const int32_t grid_dim[2],
                       const float min_position[3], const float range[3], int64_t *offsets_dev,
                       float *z_vals_dev, float *rows_dev, int row_stride, int16_t *cluster_dev,
                       int64_t capacity_rows, int64_t *total_rows_host, float *to_split_dev,
                       float *to_sample_dev, int32_t *visited_dev, bool track_visit, void *stream) {
    if (!h || !cam || !opt || !grid_dim || !min_position || !range || !offsets_dev || !z_vals_dev ||
        !rows_dev || !cluster_dev) {
        set_error("mnv_guided_samples: missing argument");
        return MNV_ERR_INVALID;
    }
    MNV_CUDA(cudaSetDevice(h->t.device));
    GuidedIO io;
    io.offscreen = offscreen;
    if (!offscreen) {
        if (!depth_arr) {
            set_error("offscreen == false needs the depth surface");
            return MNV_ERR_INVALID;
        }
        int rc = surface_for(h, depth_arr, &io.depth_surf);
        if (rc != MNV_OK) return rc;
    }
    io.offsets = offsets_dev;
    io.z_vals = z_vals_dev;
    io.rows = rows_dev;
    io.cluster = cluster_dev;
    io.row_stride = row_stride;
    io.capacity_rows = capacity_rows;
    io.total_rows = total_rows_host;
    io.to_split = to_split_dev;
    io.to_sample = to_sample_dev;
    io.visited = visited_dev;
    io.track_visit = track_visit;
    for (int i = 0; i < 2; ++i) io.grid_dim[i] = grid_dim[i];
    for (int i = 0; i < 3; ++i) {
        io.min_position[i] = min_position[i];
        io.range[i] = range[i];
    }
    return launch_guided_samples(h->t, *cam, *opt, io, static_cast<cudaStream_t>(stream));
}

int mnv_render_nerf_results(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                            void *image_arr, uint8_t *image_linear_dev,
                            const float *sample_values_dev, int value_stride, int sigma_col,
                            const float *z_vals_dev, const int64_t *offsets_dev, bool offscreen,
                            void *stream) {
    if (!h || !cam || !opt || !offsets_dev) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    cudaSurfaceObject_t surf = 0;
    int rc = surface_for(h, image_arr, &surf);
    if (rc != MNV_OK) return rc;
    if (!offscreen && !image_arr) {
        set_error("offscreen == false needs the image surface");
        return MNV_ERR_INVALID;
    }
    return launch_composite_nerf(h->t, *cam, *opt, image_linear_dev, surf, sample_values_dev,
                                 value_stride, sigma_col, z_vals_dev, offsets_dev, offscreen,
                                 static_cast<cudaStream_t>(stream));
}

int mnv_add_children_and_generate_samples(mnv_tree *h, const mnv_render_options *opt,
                                          const int32_t *parent_nodes_dev, int n, float *samples_dev,
                                          int16_t *cluster_dev, int32_t *visited_dev,
                                          const int32_t grid_dim[2], const float min_position[3],
                                          const float range[3], void *stream) {
    if (!h || !opt || (n > 0 && (!parent_nodes_dev || !samples_dev || !cluster_dev))) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_add_children(h->t, *opt, parent_nodes_dev, n, samples_dev, cluster_dev, visited_dev,
                               grid_dim, min_position, range, static_cast<cudaStream_t>(stream));
}

int mnv_tree_commit_children(mnv_tree *h, const mnv_render_options *opt, int n, const float *results_dev,
                             int result_stride, void *stream) {
    if (!h || !opt || (n > 0 && !results_dev) || result_stride < h->t.data_dim) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_commit_children(h->t, *opt, n, results_dev, result_stride, static_cast<cudaStream_t>(stream));
}

int mnv_tree_record_bytes(const mnv_tree *h, int *bytes) {
    if (!h || !bytes) return MNV_ERR_INVALID;
    *bytes = h->t.rec_u4 * 16;
    return MNV_OK;
}

int mnv_tree_reduce_children(mnv_tree *h, const mnv_render_options *opt, int n_children, const float *results_dev,
                             int result_stride, void *records_dev, void *stream) {
    if (!h || !opt || n_children < 0 || (n_children > 0 && (!results_dev || !records_dev)) ||
        result_stride < h->t.data_dim)
        return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_reduce_children(h->t, *opt, n_children, results_dev, result_stride,
                                  static_cast<uint4 *>(records_dev), static_cast<cudaStream_t>(stream));
}

int mnv_tree_commit_children_records(mnv_tree *h, const mnv_render_options *opt, int n, const void *records_dev,
                                     void *stream) {
    if (!h || !opt || (n > 0 && !records_dev)) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_commit_records(h->t, *opt, n, static_cast<const uint4 *>(records_dev),
                                 static_cast<cudaStream_t>(stream));
}

int mnv_generate_samples(mnv_tree *h, const mnv_render_options *opt, const int32_t *nodes_dev, int m,
                         float *samples_dev, int16_t *cluster_dev, const int32_t grid_dim[2],
                         const float min_position[3], const float range[3], void *stream) {
    if (!h || !opt || (m > 0 && (!nodes_dev || !samples_dev || !cluster_dev))) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_generate_samples(h->t, *opt, nodes_dev, m, samples_dev, cluster_dev, grid_dim,
                                   min_position, range, static_cast<cudaStream_t>(stream));
}

int mnv_tree_update_samples(mnv_tree *h, const mnv_render_options *opt, const int32_t *nodes_dev, int m,
                            const float *results_dev, int result_stride, void *stream) {
    if (!h || !opt || (m > 0 && (!nodes_dev || !results_dev)) || result_stride < h->t.data_dim)
        return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_update_samples(h->t, *opt, nodes_dev, m, results_dev, result_stride,
                                 static_cast<cudaStream_t>(stream));
}

int mnv_tree_prune(mnv_tree *h, const uint8_t *to_delete_dev, const int32_t *index_shifts_dev,
                   int first_shift_index, int64_t num_deleted, void *stream) {
    if (!h || !to_delete_dev || !index_shifts_dev || first_shift_index < 0 || num_deleted < 0 ||
        num_deleted >= h->t.capacity)
        return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return refine_prune(h->t, to_delete_dev, index_shifts_dev, first_shift_index, num_deleted,
                        static_cast<cudaStream_t>(stream));
}

int mnv_tree_prune_unvisited(mnv_tree *h, int32_t *visited_dev, int64_t *num_deleted_host, void *stream) {
    if (!h || !visited_dev) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return prune_unvisited(h->t, visited_dev, num_deleted_host, static_cast<cudaStream_t>(stream));
}

// The owners' receive buffers (local or IPC-mapped peers) as a device-side pointer table, refreshed only
// when the caller's pointers change (they alternate between two frame parities).
static int upload_owner_table(DeviceTree &t, int n_owners, float *const *partial_dst, cudaStream_t stream) {
    if (!t.partial_table_dev) MNV_CUDA(cudaMalloc(&t.partial_table_dev, 8 * sizeof(void *)));
    bool changed = false;
    for (int i = 0; i < n_owners; ++i) {
        if (!partial_dst[i]) return MNV_ERR_INVALID;
        changed |= t.partial_table_host[i] != partial_dst[i];
        t.partial_table_host[i] = partial_dst[i];
    }
    if (changed)  // pageable source: the copy is staged before the call returns
        MNV_CUDA(cudaMemcpyAsync(t.partial_table_dev, t.partial_table_host, 8 * sizeof(void *), cudaMemcpyHostToDevice,
                                 stream));
    return MNV_OK;
}

int mnv_render_voxels_partial(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                              const float cell_box[6], int n_owners, float *const *partial_dst, int block_pixels,
                              int slot, void *stream) {
    if (!h || !cam || !opt || !partial_dst || n_owners < 1 || n_owners > 8) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    DeviceTree &t = h->t;
    if (int rc = upload_owner_table(t, n_owners, partial_dst, static_cast<cudaStream_t>(stream)); rc != MNV_OK)
        return rc;
    RenderTargets tg;
    tg.partial_n = n_owners;
    tg.partial_block = block_pixels;
    tg.partial_slot = slot;
    tg.partial_dst = reinterpret_cast<float4 *const *>(t.partial_table_dev);
    tg.has_cell = cell_box != nullptr;
    if (cell_box) std::memcpy(tg.cell_box, cell_box, sizeof(tg.cell_box));
    return launch_render_voxels(h->t, *cam, *opt, tg, static_cast<cudaStream_t>(stream));
}

int mnv_signal_peers(uint32_t *const *flag_dst, int n, int slot, uint32_t value, void *stream) {
    if (!flag_dst) return MNV_ERR_INVALID;
    return launch_signal_peers(flag_dst, n, slot, value, static_cast<cudaStream_t>(stream));
}

int mnv_composite_partials(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                           const float *partials_dev, int n, int block_pixels, const float *boxes_host,
                           int64_t first_pixel, int n_pixels, uint8_t *rgba_dev, const uint32_t *flags_dev,
                           uint32_t wait_value, void *stream) {
    if (!h || !cam || !opt) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return launch_composite_partials(h->t, *cam, *opt, partials_dev, n, block_pixels, boxes_host, first_pixel,
                                     n_pixels, rgba_dev, flags_dev, wait_value, static_cast<cudaStream_t>(stream));
}

int mnv_composite_partials_guided(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                                  const float *partials_dev, int n, int block_pixels, const float *boxes_host,
                                  int64_t first_pixel, int n_pixels, uint8_t *rgba_dev,
                                  const uint32_t *flags_dev, uint32_t wait_value, void *stream) {
    if (!h || !cam || !opt) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    return launch_composite_partials(h->t, *cam, *opt, partials_dev, n, block_pixels, boxes_host, first_pixel,
                                     n_pixels, rgba_dev, flags_dev, wait_value, static_cast<cudaStream_t>(stream),
                                     true);
}

int mnv_guided_segment_probe(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                             const float cell_box[6], float *probe_dev, void *stream) {
    if (!h || !cam || !opt || !probe_dev) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    GuidedIO io;
    io.has_cell = cell_box != nullptr;
    if (cell_box) std::memcpy(io.cell_box, cell_box, sizeof(io.cell_box));
    io.seg_probe = reinterpret_cast<float4 *>(probe_dev);
    return launch_guided_samples(h->t, *cam, *opt, io, static_cast<cudaStream_t>(stream));
}

int mnv_guided_samples_segment(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                               const float cell_box[6], const int32_t grid_dim[2], const float min_position[3], const float range[3],
                               const float *probe_all_dev, int n_cells, int slot, int64_t *offsets_dev,
                               float *z_vals_dev, float *rows_dev, int row_stride, int16_t *cluster_dev,
                               int64_t capacity_rows, int64_t *total_rows_host, void *stream) {
    if (!h || !cam || !opt || !grid_dim || !min_position || !range || !probe_all_dev || !offsets_dev ||
        !z_vals_dev || !rows_dev || !cluster_dev) {
        set_error("mnv_guided_samples_segment: missing argument");
        return MNV_ERR_INVALID;
    }
    MNV_CUDA(cudaSetDevice(h->t.device));
    GuidedIO io;
    io.offsets = offsets_dev;
    io.z_vals = z_vals_dev;
    io.rows = rows_dev;
    io.cluster = cluster_dev;
    io.row_stride = row_stride;
    io.capacity_rows = capacity_rows;
    io.total_rows = total_rows_host;
    for (int i = 0; i < 2; ++i) io.grid_dim[i] = grid_dim[i];
    for (int i = 0; i < 3; ++i) {
        io.min_position[i] = min_position[i];
        io.range[i] = range[i];
    }
    io.has_cell = cell_box != nullptr;
    if (cell_box) std::memcpy(io.cell_box, cell_box, sizeof(io.cell_box));
    io.seg_table = reinterpret_cast<const float4 *>(probe_all_dev);
    io.seg_n = n_cells;
    io.seg_slot = slot;
    return launch_guided_samples(h->t, *cam, *opt, io, static_cast<cudaStream_t>(stream));
}

int mnv_render_nerf_results_partial(mnv_tree *h, const mnv_camera *cam, const mnv_render_options *opt,
                                    const float *sample_values_dev, int value_stride, int sigma_col,
                                    const float *z_vals_dev, const int64_t *offsets_dev,
                                    const float *probe_all_dev, int n_cells, int slot, int n_owners,
                                    float *const *partial_dst, int block_pixels, void *stream) {
    if (!h || !cam || !opt || !partial_dst || n_owners < 1 || n_owners > 8 || !offsets_dev) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(h->t.device));
    DeviceTree &t = h->t;
    if (int rc = upload_owner_table(t, n_owners, partial_dst, static_cast<cudaStream_t>(stream)); rc != MNV_OK)
        return rc;
    NerfSegment seg;
    seg.seg_table = reinterpret_cast<const float4 *>(probe_all_dev);
    seg.n_seg = n_cells;
    seg.slot = slot;
    seg.partial_dst = reinterpret_cast<float4 *const *>(t.partial_table_dev);
    seg.partial_block = block_pixels;
    return launch_composite_nerf(h->t, *cam, *opt, nullptr, 0, sample_values_dev, value_stride, sigma_col,
                                 z_vals_dev, offsets_dev, true, static_cast<cudaStream_t>(stream), &seg);
}

int mnv_ipc_export(void *ptr_dev, uint8_t handle[64]) {
    if (!ptr_dev || !handle) return MNV_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t hd;
    MNV_CUDA(cudaIpcGetMemHandle(&hd, ptr_dev));
    std::memcpy(handle, &hd, 64);
    return MNV_OK;
}
int mnv_ipc_open(const uint8_t handle[64], void **ptr_dev, int device) {
    if (!ptr_dev || !handle) return MNV_ERR_INVALID;
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    MNV_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle, 64);
    MNV_CUDA(cudaIpcOpenMemHandle(ptr_dev, hd, cudaIpcMemLazyEnablePeerAccess));
    return MNV_OK;
}
int mnv_ipc_close(void *ptr_dev) {
    if (ptr_dev) MNV_CUDA(cudaIpcCloseMemHandle(ptr_dev));
    return MNV_OK;
}

int mnv_model_create(mnv_model **out, int n_submodules, const mnv_mlp_desc *descs,
                     const int32_t grid_dim[2], const float min_position[3],
                     const float max_position[3], int device) {
    if (!out || !descs || n_submodules < 1) return MNV_ERR_INVALID;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != MNV_OK) return rc;
    mnv_model *m = new (std::nothrow) mnv_model();
    if (!m) return MNV_ERR_OOM;
    m->device = device;
    for (int i = 0; i < 2; ++i) m->grid_dim[i] = grid_dim ? grid_dim[i] : 1;
    for (int i = 0; i < 3; ++i) {
        m->min_position[i] = min_position ? min_position[i] : 0.f;
        m->max_position[i] = max_position ? max_position[i] : 1.f;
    }
    for (int i = 0; i < n_submodules; ++i) {
        MlpModel *sub = mlp_create(descs[i], device, &rc);
        if (!sub) {
            mnv_model_destroy(m);
            return rc;
        }
        m->subs.push_back(sub);
    }
    *out = m;
    return MNV_OK;
}

int mnv_model_destroy(mnv_model *m) {
    if (!m) return MNV_OK;
    cudaSetDevice(m->device);
    for (MlpModel *s : m->subs) mlp_destroy(s);
    delete m;
    return MNV_OK;
}

int mnv_model_info(const mnv_model *m, int *n_submodules, int *in_dim, int *out_dim,
                   double *flops_per_row) {
    if (!m || m->subs.empty()) return MNV_ERR_INVALID;
    if (n_submodules) *n_submodules = (int) m->subs.size();
    if (in_dim) *in_dim = mlp_in_dim(m->subs[0]);
    if (out_dim) *out_dim = mlp_out_dim(m->subs[0]);
    if (flops_per_row) *flops_per_row = mlp_flops_per_row(m->subs[0]);
    return MNV_OK;
}

int mnv_mlp_forward(mnv_model *m, int submodule, const float *x_dev, int64_t rows, int in_dim,
                    float *out_dev, int out_stride, void *stream) {
    if (!m || submodule < 0 || submodule >= (int) m->subs.size() ||
        (rows > 0 && (!x_dev || !out_dev))) {
        set_error("mnv_mlp_forward: bad arguments");
        return MNV_ERR_INVALID;
    }
    MNV_CUDA(cudaSetDevice(m->device));
    return mlp_forward(m->subs[submodule], x_dev, rows, in_dim, out_dev, out_stride,
                       static_cast<cudaStream_t>(stream));
}

int mnv_query_submodules(mnv_model *m, const int16_t *cluster_dev, const float *rows_dev, int in_dim,
                         int64_t rows, float *out_dev, int out_stride, void *stream) {
    if (!m || m->subs.empty() || (rows > 0 && (!cluster_dev || !rows_dev || !out_dev))) return MNV_ERR_INVALID;
    MNV_CUDA(cudaSetDevice(m->device));
    return query_submodules(m->subs.data(), (int) m->subs.size(), cluster_dev, rows_dev, in_dim, rows, out_dev,
                            out_stride, static_cast<cudaStream_t>(stream));
}

int mnv_select_split_candidates(const float *to_split_dev, int64_t n_rays, int max_n, int32_t *nodes_dev,
                                int *n_selected, int *n_candidates, void *stream) {
    if (!to_split_dev || !nodes_dev || n_rays <= 0 || max_n <= 0) return MNV_ERR_INVALID;
    return select_split_candidates(to_split_dev, n_rays, max_n, nodes_dev, n_selected, n_candidates,
                                   static_cast<cudaStream_t>(stream));
}

int mnv_select_sample_candidates(const float *to_sample_dev, int64_t n_rays, int max_n,
                                 int32_t *nodes_dev, int *n_selected, int *n_candidates, void *stream) {
    if (!to_sample_dev || !nodes_dev || n_rays <= 0 || max_n <= 0) return MNV_ERR_INVALID;
    return select_sample_candidates(to_sample_dev, n_rays, max_n, nodes_dev, n_selected, n_candidates,
                                    static_cast<cudaStream_t>(stream));
}

int mnv_vote_reduce(const float *tracker_dev, int64_t n_rays, uint32_t *records_dev, int64_t cap_records,
                    int64_t *n_records, void *stream) {
    if (!tracker_dev || !records_dev || n_rays <= 0 || cap_records <= 0) return MNV_ERR_INVALID;
    return vote_reduce(tracker_dev, n_rays, records_dev, cap_records, n_records, static_cast<cudaStream_t>(stream));
}

int mnv_select_candidates_from_votes(int kind, const float *tracker_dev, int64_t n_rays,
                                     const uint32_t *records_dev, int64_t n_records, int max_n,
                                     int32_t *nodes_dev, int *n_selected, int *n_candidates, void *stream) {
    if ((kind != 0 && kind != 1) || !nodes_dev || max_n <= 0 || n_rays < 0 || n_records < 0 ||
        (n_rays > 0 && !tracker_dev) || (n_records > 0 && !records_dev))
        return MNV_ERR_INVALID;
    return select_candidates(kind, n_rays > 0 ? tracker_dev : nullptr, n_rays, n_records > 0 ? records_dev : nullptr,
                             n_records, max_n, nodes_dev, n_selected, n_candidates, static_cast<cudaStream_t>(stream));
}

float mnv_tracker_encode_chunk(int32_t chunk) { return tracker_encode_chunk(chunk); }
int32_t mnv_tracker_decode_chunk(float v) { return tracker_decode_chunk(v); }

int mnv_tree_trackers(mnv_tree *h, float **to_split_dev, float **to_sample_dev) {
    if (!h) return MNV_ERR_INVALID;
    if (to_split_dev) *to_split_dev = h->t.split_dev;
    if (to_sample_dev) *to_sample_dev = h->t.sample_dev;
    return MNV_OK;
}

}  // extern "C"
