// Replica group: the image-tile multi-GPU mode (first row) driven from ONE host process — the
// shape the viewer has (one process, one GL context on the display GPU; the reference is single-GPU,
// src/cuda/renderer_kernel.cu:17 hard-codes device 0).  bench.py scales with one process per GPU over
// torch.distributed; this is the native C++ host path for the same partition:
//
//   * the tree is replicated on every device of the group (each replica is an ordinary mnv_tree),
//   * replica i marches the interleaved bands b % n == i of the frame (bit-identical to the one-GPU frame: rays are
//     independent) into its own linear frame buffer, all launches issued back to back from the calling thread,
//   * finished bands travel to the display GPU (devices[0]) as ONE strided peer copy per replica over NVLink
//     (cudaMemcpy2DAsync between peer-enabled devices: the copy engines move them, no staging through the host),
//     ordered by events: the display GPU's stream waits for every replica's copy, nothing blocks the host,
//   * the gathered frame is either the caller's linear RGBA8 buffer or the cudaArray behind the GL renderbuffer
//     (one device-local 2-D copy into the interop surface, cuda_renderer.cpp:432-457).
//
// Dynamic refinement on the group (mnv_group_refine): every replica reduces the votes of its own bands to
// (leaf, priority, votes) records, the records are exchanged with peer copies, every replica runs the identical
// selection and links the children, the n*8*c MLP rows are sharded by child, the fp16 payload records are
// exchanged and committed everywhere — the C++ form of multigpu.ReplicatedPipeline.refine_frame, with peer copies
// in place of NCCL.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mnv_internal.cuh"

struct mnv_tree;
struct mnv_model;

namespace mnv {
DeviceTree &device_tree_of(mnv_tree *h);  // mnv_capi.cu
}

struct mnv_group {
    struct Replica {
        int device = 0;
        mnv_tree *tree = nullptr;
        uint8_t *frame = nullptr;  // RGBA8 [H][W] on this device
        size_t frame_bytes = 0;
        float *split = nullptr, *sample = nullptr;  // trackers [P][3] (refinement)
        size_t tracker_rays = 0;
        cudaEvent_t done = nullptr;  // bands rendered and delivered
        // refinement scratch
        uint32_t *records = nullptr;  // [n_dev][rec_cap][3] gathered vote records
        int64_t rec_cap = 0;
        int32_t *nodes = nullptr;
        float *samples = nullptr, *results = nullptr;
        int16_t *cluster = nullptr;
        uint8_t *payload = nullptr;  // [n_dev * per][record_bytes]
        size_t samples_cap = 0, results_cap = 0, payload_cap = 0;
    };
    std::vector<Replica> r;
    cudaEvent_t start = nullptr;  // on devices[0]: the display stream's position when the frame began
    uint64_t step = 0;
};

using namespace mnv;

namespace {

cudaStream_t stream_of(mnv_group::Replica &rp) { return device_tree_of(rp.tree).stream; }

// bands b % mod == rem of a [H][row_bytes] image: one strided copy for the complete bands, one for a ragged tail
int copy_bands(uint8_t *dst, const uint8_t *src, int height, size_t row_bytes, int band_rows, int mod, int rem,
               cudaStream_t stream) {
    const size_t band_bytes = row_bytes * (size_t) band_rows;
    const int n_bands = (height + band_rows - 1) / band_rows;
    const int full = height / band_rows;
    int mine = 0;
    for (int b = rem; b < full; b += mod) ++mine;
    const size_t off = (size_t) rem * band_bytes;
    if (mine > 0)
        MNV_CUDA(cudaMemcpy2DAsync(dst + off, band_bytes * mod, src + off, band_bytes * mod, band_bytes, mine,
                                   cudaMemcpyDefault, stream));
    if (n_bands > full && (n_bands - 1) % mod == rem) {
        const size_t o2 = (size_t) full * band_bytes;
        MNV_CUDA(cudaMemcpyAsync(dst + o2, src + o2, (size_t) height * row_bytes - o2, cudaMemcpyDefault, stream));
    }
    return MNV_OK;
}

int ensure_frame(mnv_group::Replica &rp, size_t bytes) {
    if (rp.frame_bytes >= bytes) return MNV_OK;
    MNV_CUDA(cudaSetDevice(rp.device));
    cudaFree(rp.frame);
    rp.frame = nullptr;
    rp.frame_bytes = 0;
    MNV_CUDA(cudaMalloc(&rp.frame, bytes));
    rp.frame_bytes = bytes;
    return MNV_OK;
}

}  // namespace

extern "C" {

int mnv_group_create(mnv_group **out, const mnv_tree_desc *desc, int64_t max_capacity, const int *devices,
                     int n_devices) {
    if (!out || !desc || !devices || n_devices < 1 || n_devices > 16) return MNV_ERR_INVALID;
    *out = nullptr;
    mnv_group *g = new (std::nothrow) mnv_group();
    if (!g) return MNV_ERR_OOM;
    g->r.resize((size_t) n_devices);
    int rc = MNV_OK;
    for (int i = 0; i < n_devices && rc == MNV_OK; ++i) {
        g->r[(size_t) i].device = devices[i];
        rc = mnv_tree_create(&g->r[(size_t) i].tree, desc, max_capacity, devices[i]);
        if (rc == MNV_OK && cudaEventCreateWithFlags(&g->r[(size_t) i].done, cudaEventDisableTiming) != cudaSuccess)
            rc = MNV_ERR_CUDA;
    }
    if (rc == MNV_OK) {
        // peer access both ways between every pair (copies and, for refinement, peer reads of the records)
        for (int i = 0; i < n_devices; ++i)
            for (int j = 0; j < n_devices; ++j) {
                if (i == j) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                if (!can) continue;
                cudaSetDevice(devices[i]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess) cudaGetLastError();  // already enabled by someone else: fine
            }
        cudaSetDevice(devices[0]);
        if (cudaEventCreateWithFlags(&g->start, cudaEventDisableTiming) != cudaSuccess) rc = MNV_ERR_CUDA;
    }
    if (rc != MNV_OK) {
        mnv_group_destroy(g);
        return rc;
    }
    *out = g;
    return MNV_OK;
}

int mnv_group_destroy(mnv_group *g) {
    if (!g) return MNV_OK;
    for (auto &rp : g->r) {
        cudaSetDevice(rp.device);
        if (rp.tree) {
            cudaStreamSynchronize(device_tree_of(rp.tree).stream);
            mnv_tree_destroy(rp.tree);
        }
        cudaFree(rp.frame);
        cudaFree(rp.split);
        cudaFree(rp.sample);
        cudaFree(rp.records);
        cudaFree(rp.nodes);
        cudaFree(rp.samples);
        cudaFree(rp.results);
        cudaFree(rp.cluster);
        cudaFree(rp.payload);
        if (rp.done) cudaEventDestroy(rp.done);
    }
    if (g->start) {
        cudaSetDevice(g->r[0].device);
        cudaEventDestroy(g->start);
    }
    delete g;
    return MNV_OK;
}

int mnv_group_size(const mnv_group *g, int *n) {
    if (!g || !n) return MNV_ERR_INVALID;
    *n = (int) g->r.size();
    return MNV_OK;
}

int mnv_group_tree(mnv_group *g, int i, mnv_tree **tree) {
    if (!g || !tree || i < 0 || i >= (int) g->r.size()) return MNV_ERR_INVALID;
    *tree = g->r[(size_t) i].tree;
    return MNV_OK;
}

// Shared by render_frame and refine: replica i marches its bands (with trackers when `track`), bands land in
// replica 0's frame buffer, replica 0's stream ends up ordered after every delivery.
static int group_march(mnv_group *g, const mnv_camera *cam, const mnv_render_options *opt, int band_rows, bool track) {
    const int n = (int) g->r.size();
    const int W = cam->width, H = cam->height;
    if (W <= 0 || H <= 0 || band_rows < 8 || band_rows % 8) {
        set_error("group frame: bad size %dx%d or band_rows %d (multiple of 8)", W, H, band_rows);
        return MNV_ERR_INVALID;
    }
    const size_t row_bytes = (size_t) W * 4, bytes = row_bytes * (size_t) H;
    const size_t rays = (size_t) W * H;
    for (auto &rp : g->r) {
        int rc = ensure_frame(rp, bytes);
        if (rc != MNV_OK) return rc;
        if (track && rp.tracker_rays < rays) {
            MNV_CUDA(cudaSetDevice(rp.device));
            cudaFree(rp.split);
            cudaFree(rp.sample);
            rp.split = rp.sample = nullptr;
            rp.tracker_rays = 0;
            MNV_CUDA(cudaMalloc(&rp.split, rays * 3 * sizeof(float)));
            MNV_CUDA(cudaMalloc(&rp.sample, rays * 3 * sizeof(float)));
            rp.tracker_rays = rays;
        }
    }
    mnv_group::Replica &r0 = g->r[0];
    MNV_CUDA(cudaSetDevice(r0.device));
    MNV_CUDA(cudaEventRecord(g->start, stream_of(r0)));
    const int tile_w = ((W + 15) / 16) * 16;
    for (int i = 0; i < n; ++i) {
        mnv_group::Replica &rp = g->r[(size_t) i];
        MNV_CUDA(cudaSetDevice(rp.device));
        cudaStream_t s = stream_of(rp);
        if (i > 0) MNV_CUDA(cudaStreamWaitEvent(s, g->start, 0));  // replica 0's buffer is free to be written
        if (track) {
            // rows of other replicas stay "no candidate": the vote reduction reads the whole array
            int rc = fill_f32(rp.split, -1.f, (int64_t) rays * 3, s);
            if (rc == MNV_OK) rc = fill_f32(rp.sample, -1.f, (int64_t) rays * 3, s);
            if (rc != MNV_OK) return rc;
        }
        int rc = mnv_render_voxels_tiles(rp.tree, cam, opt, rp.frame, track ? rp.split : nullptr,
                                         track ? rp.sample : nullptr, tile_w, band_rows, n, i, s);
        if (rc != MNV_OK) return rc;
        if (i > 0) {
            rc = copy_bands(r0.frame, rp.frame, H, row_bytes, band_rows, n, i, s);  // NVLink peer copy
            if (rc != MNV_OK) return rc;
            MNV_CUDA(cudaEventRecord(rp.done, s));
        }
    }
    MNV_CUDA(cudaSetDevice(r0.device));
    for (int i = 1; i < n; ++i) MNV_CUDA(cudaStreamWaitEvent(stream_of(r0), g->r[(size_t) i].done, 0));
    return MNV_OK;
}

// deliver replica 0's gathered frame: linear device buffer, cudaArray (GL interop surface) or host memory
static int group_deliver(mnv_group *g, const mnv_camera *cam, uint8_t *linear_dev0, void *image_arr_dev0,
                         uint8_t *rgba_host) {
    mnv_group::Replica &r0 = g->r[0];
    const size_t row_bytes = (size_t) cam->width * 4, bytes = row_bytes * (size_t) cam->height;
    cudaStream_t s0 = stream_of(r0);
    if (linear_dev0) MNV_CUDA(cudaMemcpyAsync(linear_dev0, r0.frame, bytes, cudaMemcpyDeviceToDevice, s0));
    if (image_arr_dev0)
        MNV_CUDA(cudaMemcpy2DToArrayAsync(static_cast<cudaArray_t>(image_arr_dev0), 0, 0, r0.frame, row_bytes, row_bytes,
                                          (size_t) cam->height, cudaMemcpyDeviceToDevice, s0));
    if (rgba_host) {
        MNV_CUDA(cudaMemcpyAsync(rgba_host, r0.frame, bytes, cudaMemcpyDeviceToHost, s0));
        MNV_CUDA(cudaStreamSynchronize(s0));
    }
    return MNV_OK;
}

int mnv_group_render_frame(mnv_group *g, const mnv_camera *cam, const mnv_render_options *opt,
                           uint8_t *image_linear_dev0, void *image_arr_dev0, int band_rows) {
    if (!g || !cam || !opt || (!image_linear_dev0 && !image_arr_dev0)) return MNV_ERR_INVALID;
    int rc = group_march(g, cam, opt, band_rows, false);
    if (rc != MNV_OK) return rc;
    return group_deliver(g, cam, image_linear_dev0, image_arr_dev0, nullptr);
}

int mnv_group_render_frame_host(mnv_group *g, const mnv_camera *cam, const mnv_render_options *opt, uint8_t *rgba_host,
                                int band_rows) {
    if (!g || !cam || !opt || !rgba_host) return MNV_ERR_INVALID;
    int rc = group_march(g, cam, opt, band_rows, false);
    if (rc != MNV_OK) return rc;
    return group_deliver(g, cam, nullptr, nullptr, rgba_host);
}

int mnv_group_synchronize(mnv_group *g) {
    if (!g) return MNV_ERR_INVALID;
    for (auto &rp : g->r) {
        MNV_CUDA(cudaSetDevice(rp.device));
        MNV_CUDA(cudaStreamSynchronize(stream_of(rp)));
    }
    return MNV_OK;
}

// One frame with dynamic refinement ON across the group (Impl::render + expand_voxels, cuda_renderer.cpp:68-163,
// 205-278).  models[i] is the sub-MLP container on replica i's device.  *nodes_added = leaves split (0: nothing to
// split or the tree is full).  Every replica ends the call with the same tree, bit for bit.
int mnv_group_refine_frame(mnv_group *g, mnv_model *const *models, const mnv_camera *cam, const mnv_render_options *opt,
                           const int32_t grid_dim[2], const float min_position[3], const float range[3], uint64_t seed,
                           uint8_t *image_linear_dev0, void *image_arr_dev0, uint8_t *rgba_host, int band_rows,
                           int *nodes_added) {
    if (!g || !models || !cam || !opt || !grid_dim || !min_position || !range) return MNV_ERR_INVALID;
    if (nodes_added) *nodes_added = 0;
    const int n = (int) g->r.size();
    int rc = group_march(g, cam, opt, band_rows, true);
    if (rc != MNV_OK) return rc;
    const int64_t rays = (int64_t) cam->width * cam->height;
    const int64_t cap_local = rays / n + (int64_t) cam->width * band_rows + 1024;
    const int max_n = opt->split_batch_size;
    const int c = opt->samples_per_corner;
    const int rd = 3 + (opt->need_viewdir ? 3 : 0) + (opt->appearance_embedding != -1 ? 1 : 0);
    std::vector<int64_t> n_rec((size_t) n, 0);
    std::vector<int> k((size_t) n, 0);
    ++g->step;
    // One host thread per replica for the whole exchange: issued from one thread, the ~30 runtime calls a replica needs per
    // frame (reduction, 7 record pulls + paddings, selection, sampling, MLP share, 7 payload pulls, commit) queue up behind
    // those of the other replicas — 4.0 ms per 4K frame on 8 GPUs against 3.4 ms on 4 (profiles/r2_headless_group_n8b.log).
    // The threads meet at three points: record counts known, selection sizes known, payload shares written.
    // (Replicas that share a device — tests — serialise on that device's scratch mutex inside launch / finish pairs.)
    struct Meet {
        std::mutex mu;
        std::condition_variable cv;
        int waiting = 0, generation = 0;
        void arrive_and_wait(int parties) {
            std::unique_lock<std::mutex> lk(mu);
            const int gen = generation;
            if (++waiting == parties) {
                waiting = 0;
                ++generation;
                cv.notify_all();
            } else {
                cv.wait(lk, [&] { return generation != gen; });
            }
        }
    } meet;
    std::atomic<int> failed{MNV_OK};
    std::atomic<int> full{0};
    std::vector<std::string> msgs((size_t) n);
    auto worker = [&](int i) {
        mnv_group::Replica &rp = g->r[(size_t) i];
        auto fail = [&](int code) {
            int expect = MNV_OK;
            if (failed.compare_exchange_strong(expect, code)) msgs[(size_t) i] = mnv_last_error();  // the message is thread-local
        };
#define GRP_CUDA(x)                                                   \
    do {                                                              \
        if (failed.load() == MNV_OK) {                                \
            const cudaError_t e_ = (x);                               \
            if (e_ != cudaSuccess) fail(cuda_fail(e_, #x, __FILE__, __LINE__)); \
        }                                                             \
    } while (0)
        cudaSetDevice(rp.device);
        cudaStream_t s = stream_of(rp);
        // 1. votes of this replica's own bands -> records, into slot i of its own gather buffer (the finish call
        //    synchronises the stream: the record count sizes the exchange)
        if (rp.rec_cap < cap_local) {
            cudaFree(rp.records);
            rp.records = nullptr;
            rp.rec_cap = 0;
            GRP_CUDA(cudaMalloc(&rp.records, (size_t) n * cap_local * 3 * sizeof(uint32_t)));
            if (rp.records) rp.rec_cap = cap_local;
        }
        if (!rp.nodes) GRP_CUDA(cudaMalloc(&rp.nodes, (size_t) 16384 * 2 * sizeof(int32_t)));
        if (failed.load() == MNV_OK) {
            int r = vote_reduce_launch(rp.split, rays, rp.records + (size_t) i * rp.rec_cap * 3, rp.rec_cap, s);
            if (r == MNV_OK) r = vote_reduce_finish(rp.rec_cap, &n_rec[(size_t) i], s);
            if (r != MNV_OK) fail(r);
        }
        meet.arrive_and_wait(n);
        // 2. exchange: pull the peers' records (peer copies); layout [j][rec_cap][3] with zero-vote padding between the
        //    blocks so that one contiguous range can be handed to the selection
        for (int j = 0; j < n && failed.load() == MNV_OK; ++j) {
            if (j != i && n_rec[(size_t) j] > 0)
                GRP_CUDA(cudaMemcpyAsync(rp.records + (size_t) j * rp.rec_cap * 3,
                                         g->r[(size_t) j].records + (size_t) j * g->r[(size_t) j].rec_cap * 3,
                                         (size_t) n_rec[(size_t) j] * 3 * sizeof(uint32_t), cudaMemcpyDefault, s));
            const int64_t pad = rp.rec_cap - n_rec[(size_t) j];
            if (pad > 0)
                GRP_CUDA(cudaMemsetAsync(rp.records + ((size_t) j * rp.rec_cap + (size_t) n_rec[(size_t) j]) * 3, 0,
                                         (size_t) pad * 3 * sizeof(uint32_t), s));
        }
        // 3. identical selection on every replica
        if (failed.load() == MNV_OK) {
            int cand = 0;
            int r = select_candidates_launch(0, nullptr, 0, rp.records, (int64_t) n * rp.rec_cap, max_n, rp.nodes, s);
            if (r == MNV_OK) r = select_candidates_finish(&k[(size_t) i], &cand, s);
            if (r != MNV_OK) fail(r);
        }
        meet.arrive_and_wait(n);
        if (failed.load() == MNV_OK && k[(size_t) i] != k[0]) {
            set_error("group refine: replicas selected %d vs %d leaves", k[(size_t) i], k[0]);
            fail(MNV_ERR_INVALID);
        }
        const int kk = k[0];
        const DeviceTree &t = device_tree_of(rp.tree);
        const bool is_full = kk > 0 && t.capacity + kk > t.max_capacity;  // "Full", cuda_renderer.cpp:228-231
        if (is_full) full.store(1);
        const int children = kk * 8;
        const int per = (children + n - 1) / n;
        int rec_bytes = 0;
        mnv_tree_record_bytes(rp.tree, &rec_bytes);
        const int out_stride = t.data_dim + 1;
        const bool go = kk > 0 && !is_full;
        if (go && failed.load() == MNV_OK) {
            const size_t need_s = (size_t) children * c * rd, need_r = (size_t) per * c * out_stride,
                         need_p = (size_t) n * per * rec_bytes;
            if (rp.samples_cap < need_s) {
                cudaFree(rp.samples);
                cudaFree(rp.cluster);
                rp.samples = nullptr;
                rp.cluster = nullptr;
                rp.samples_cap = 0;
                GRP_CUDA(cudaMalloc(&rp.samples, need_s * sizeof(float)));
                GRP_CUDA(cudaMalloc(&rp.cluster, (size_t) children * c * sizeof(int16_t)));
                if (rp.samples && rp.cluster) rp.samples_cap = need_s;
            }
            if (rp.results_cap < need_r) {
                cudaFree(rp.results);
                rp.results = nullptr;
                rp.results_cap = 0;
                GRP_CUDA(cudaMalloc(&rp.results, need_r * sizeof(float)));
                if (rp.results) rp.results_cap = need_r;
            }
            if (rp.payload_cap < need_p) {
                cudaFree(rp.payload);
                rp.payload = nullptr;
                rp.payload_cap = 0;
                GRP_CUDA(cudaMalloc(&rp.payload, need_p));
                if (rp.payload) rp.payload_cap = need_p;
            }
            if (failed.load() == MNV_OK) {
                // the same counter-based random numbers on every replica (torch::rand in the reference, :250)
                int r = fill_uniform(rp.samples, (int64_t) need_s, seed + g->step, s);
                if (r == MNV_OK)
                    r = mnv_add_children_and_generate_samples(rp.tree, opt, rp.nodes, kk, rp.samples, rp.cluster, nullptr,
                                                              grid_dim, min_position, range, s);
                // 4. this replica's share of the MLP rows -> payload records, into slot i of its own gather buffer
                const int lo = std::min(i * per, children), hi = std::min((i + 1) * per, children);
                if (r == MNV_OK) GRP_CUDA(cudaMemsetAsync(rp.payload + (size_t) i * per * rec_bytes, 0, (size_t) per * rec_bytes, s));
                if (r == MNV_OK && hi > lo) {
                    r = mnv_query_submodules(models[i], rp.cluster + (size_t) lo * c, rp.samples + (size_t) lo * c * rd, rd,
                                             (int64_t) (hi - lo) * c, rp.results, out_stride, s);
                    if (r == MNV_OK)
                        r = mnv_tree_reduce_children(rp.tree, opt, hi - lo, rp.results, out_stride,
                                                     rp.payload + (size_t) i * per * rec_bytes, s);
                }
                if (r != MNV_OK) fail(r);
                GRP_CUDA(cudaEventRecord(rp.done, s));
            }
        }
        meet.arrive_and_wait(n);
        // 5. payload exchange (64 B per child) and commit
        if (go && failed.load() == MNV_OK) {
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                GRP_CUDA(cudaStreamWaitEvent(s, g->r[(size_t) j].done, 0));
                GRP_CUDA(cudaMemcpyAsync(rp.payload + (size_t) j * per * rec_bytes,
                                         g->r[(size_t) j].payload + (size_t) j * per * rec_bytes, (size_t) per * rec_bytes,
                                         cudaMemcpyDefault, s));
            }
            if (failed.load() == MNV_OK) {
                const int r = mnv_tree_commit_children_records(rp.tree, opt, kk, rp.payload, s);
                if (r != MNV_OK) fail(r);
            }
        }
#undef GRP_CUDA
    };
    {
        std::vector<std::thread> workers;
        for (int i = 1; i < n; ++i) workers.emplace_back(worker, i);
        worker(0);
        for (auto &w : workers) w.join();
    }
    if (failed.load() != MNV_OK) {
        for (int i = 0; i < n; ++i)
            if (!msgs[(size_t) i].empty()) set_error("%s", msgs[(size_t) i].c_str());
        return failed.load();
    }
    if (nodes_added && !full.load()) *nodes_added = k[0];
    rc = group_deliver(g, cam, image_linear_dev0, image_arr_dev0, rgba_host);
    // the exchange buffers are read by the peers: nobody may start the next frame's reduction before all copies landed
    if (rc == MNV_OK) rc = mnv_group_synchronize(g);
    return rc;
}

}  // extern "C"
