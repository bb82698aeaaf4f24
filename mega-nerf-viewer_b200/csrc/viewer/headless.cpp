// mnv_headless — offscreen driver of viewer::VolumeRenderer (the "headless offscreen bench mode"
// BASELINE.json asks for).  Takes the options of the reference's CLI that concern the render
// path (src/opts.cpp) and replaces the GLFW loop of main.cpp:595-611 with a fixed camera orbit.
//
//   mnv_headless tree.npz [--model model.npz] [--width W --height H] [--frames N] [--poses K]
//                [--max_tree_capacity C] [--use_splitting] [--use_guided_sampling]
//                [--bg B] [--out frame.ppm] [--raw frame.rgba] [--save refined.npz] [--seed S] [--verbose]
//                [--interop] [--selftest-load] [--selftest-camera] [--selftest-wireframe D]
//
// Prints one JSON line with the wall-clock frame statistics (camera upload + render + refinement
// + RGBA8 read-back per frame).
#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mnv_b200.h"
#include "n3tree.hpp"
#include "renderer.hpp"

using namespace viewer;

namespace {

// 64-bit position-weighted byte sum (mega_nerf_viewer_b200.bytes_checksum restates it in numpy)
uint64_t checksum(const void *p, size_t n) {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    uint64_t h = 0;
    for (size_t i = 0; i < n; ++i) h += (uint64_t) (b[i] + 1u) * ((uint64_t) i * 2654435761ull + 1ull);
    return h;
}

// pose k of an n-pose orbit about z of the viewer's default camera (main.cpp:491-504 defaults
// scaled by 0.5, the config-2 camera of SURVEY.md §8(d))
void set_pose(Camera &cam, int k, int n) {
    const double ang = 2.0 * M_PI * k / n, c = std::cos(ang), s = std::sin(ang);
    const double cen[3] = {-3.5 * 0.5, 0.0, 3.5 * 0.5}, back[3] = {-0.7071068, 0.0, 0.7071068};
    cam.center = vec3((float) (c * cen[0] - s * cen[1]), (float) (s * cen[0] + c * cen[1]), (float) cen[2]);
    cam.v_back = vec3((float) (c * back[0] - s * back[1]), (float) (s * back[0] + c * back[1]), (float) back[2]);
    cam.v_world_up = vec3(0.f, 0.f, 1.f);
}

void print_camera(const Camera &c) {
    std::printf("{\"transform\": [");
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) std::printf("%s%.9g", (i || j) ? ", " : "", c.transform[i][j]);
    std::printf("], \"K\": [");
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) std::printf("%s%.9g", (i || j) ? ", " : "", c.K[i][j]);
    std::printf("], \"w2c\": [");
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) std::printf("%s%.9g", (i || j) ? ", " : "", c.w2c[i][j]);
    std::printf("], \"fx\": %.9g, \"fy\": %.9g, \"cx\": %.9g, \"cy\": %.9g, \"width\": %d, \"height\": %d}\n", c.fx,
                c.fy, c.cx, c.cy, c.width, c.height);
}

}  // namespace

int main(int argc, char **argv) {
    std::string tree_path, model_path, out_ppm, out_raw, save_path, resave_path;
    int width = 1920, height = 1080, frames = 16, poses = 16, wire_depth = -1, gpus = 1, replicas = 0;
    long max_cap = 0;
    bool splitting = false, guided = false, verbose = false, st_load = false, st_camera = false, interop = false;
    float bg = 0.f;
    uint64_t seed = 0x5eed;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value for %s\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--model") model_path = val();
        else if (a == "--width") width = std::atoi(val());
        else if (a == "--height") height = std::atoi(val());
        else if (a == "--frames") frames = std::atoi(val());
        else if (a == "--poses") poses = std::atoi(val());
        else if (a == "--max_tree_capacity") max_cap = std::atol(val());
        else if (a == "--use_splitting") splitting = true;
        else if (a == "--use_guided_sampling") guided = true;
        else if (a == "--bg") bg = (float) std::atof(val());
        else if (a == "--out") out_ppm = val();
        else if (a == "--raw") out_raw = val();
        else if (a == "--save") save_path = val();              // write the (refined) tree after the last frame
        else if (a == "--selftest-resave") resave_path = val(); // host only: load, write back
        else if (a == "--seed") seed = std::strtoull(val(), nullptr, 0);
        else if (a == "--verbose") verbose = true;
        else if (a == "--gpus") gpus = std::atoi(val());  // image tiles over this many devices (replica group)
        else if (a == "--replicas") replicas = std::atoi(val());  // dev/tests: group members, devices wrap around
        else if (a == "--interop") interop = true;  // present through cudaArray surfaces, like the GL viewer does
        else if (a == "--selftest-load") st_load = true;
        else if (a == "--selftest-camera") st_camera = true;
        else if (a == "--selftest-wireframe") wire_depth = std::atoi(val());
        else if (a[0] != '-') tree_path = a;
        else {
            std::fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }
    try {
        if (st_camera) {  // host-only: Camera defaults, _update, drag, resize-free
            Camera c(width, height, 1111.f * (width / 800.f));
            print_camera(c);
            c.begin_drag(100.f, 120.f, false, false);
            c.drag_update(260.f, 90.f);
            c.end_drag();
            c._update();
            print_camera(c);
            c.begin_drag(10.f, 10.f, true, false);
            c.drag_update(40.f, 70.f);
            c.end_drag();
            c.move(vec3(0.1f, -0.2f, 0.3f));
            c._update();
            print_camera(c);
            return 0;
        }
        if (tree_path.empty()) {
            std::fprintf(stderr, "usage: mnv_headless tree.npz [options]\n");
            return 2;
        }
        N3Tree tree(tree_path);
        if (tree.capacity == 0) return 3;
        if (!resave_path.empty()) {
            tree.save(resave_path);
            return 0;
        }
        if (st_load) {  // host-only: checksums of what N3Tree::open produced
            tree.decode_vq_host();  // host consumers read `data` (the GPU decode is exercised by a render)
            std::printf("{\"N\": %d, \"data_dim\": %d, \"format\": \"%s\", \"basis_dim\": %d, \"capacity\": %d, "
                        "\"scale\": [%.9g, %.9g, %.9g], \"offset\": [%.9g, %.9g, %.9g], "
                        "\"child\": \"%016" PRIx64 "\", \"parent\": \"%016" PRIx64 "\", \"data\": \"%016" PRIx64
                        "\", \"sample_counts_all_8\": %s, \"pack\": %lld, \"unpack\": [%d, %d, %d, %d]}\n",
                        tree.N, tree.data_dim, tree.data_format.to_string().c_str(), tree.data_format.basis_dim,
                        tree.capacity, tree.scale.v[0], tree.scale.v[1], tree.scale.v[2], tree.offset.v[0],
                        tree.offset.v[1], tree.offset.v[2], checksum(tree.child.v.data(), tree.child.v.size() * 4),
                        checksum(tree.parent.v.data(), tree.parent.v.size() * 4),
                        checksum(tree.data.v.data(), tree.data.v.size() * 2),
                        std::all_of(tree.sample_counts.v.begin(), tree.sample_counts.v.end(),
                                    [](int16_t s) { return s == 8; })
                                ? "true"
                                : "false",
                        (long long) tree.pack_index(5, 1, 0, 1), std::get<0>(tree.unpack_index(45)),
                        std::get<1>(tree.unpack_index(45)), std::get<2>(tree.unpack_index(45)),
                        std::get<3>(tree.unpack_index(45)));
            return 0;
        }
        if (wire_depth >= 0) {
            const std::vector<float> w = tree.gen_wireframe(wire_depth);
            std::printf("{\"floats\": %zu, \"hash\": \"%016" PRIx64 "\"}\n", w.size(), checksum(w.data(), w.size() * 4));
            return 0;
        }

        VolumeRenderer rend;
        rend.verbose = verbose;
        rend.rng_seed = seed;
        rend.options.background_brightness = bg;
        rend.options.use_splitting = splitting;
        rend.options.use_guided_sampling = guided;
        if (max_cap <= 0) max_cap = (long) tree.capacity + (splitting ? 4 * 4192L * (frames + 4) : 8192);
        if (gpus > 1 || replicas > 1) {
            int have = 0;
            mnv_device_count(&have);
            if (gpus > have) throw std::runtime_error("--gpus exceeds the visible devices");
            std::vector<int> devs;
            for (int i = 0; i < std::max(gpus, replicas); ++i) devs.push_back(i % std::max(gpus, 1));
            rend.set_devices(devs);
        }
        rend.set(tree, max_cap);
        if (!model_path.empty()) rend.load_model(model_path);
        // the reference constructs a 256x256 camera and resizes it to the window (main.cpp:593);
        // that first resize keeps the focal length, which is then set for the frame size
        rend.resize(width, height);
        rend.camera.fx = rend.camera.fy = 1111.f * (width / 800.f);

        // --interop: the presentation path of the GL viewer without GL — double-buffered RGBA8 + R32F cudaArray
        // surfaces (what cudaGraphicsSubResourceGetMappedArray hands out, cuda_renderer.cpp:447-455), cleared per
        // frame like glClearNamedFramebufferfv does (:72-79), composited in place with offscreen = false
        void *ca[4] = {nullptr, nullptr, nullptr, nullptr};
        std::vector<uint8_t> clear_rgba, frame_px;
        std::vector<float> clear_depth;
        if (interop) {
            for (int i = 0; i < 4; ++i)
                if (mnv_array_create(&ca[i], width, height, i & 1, 0) != MNV_OK) throw std::runtime_error(mnv_last_error());
            rend.set_interop_surfaces(ca);
            const uint8_t c8 = (uint8_t) std::min(255.f, std::max(0.f, bg * 255.f));
            clear_rgba.assign((size_t) width * height * 4, c8);
            for (size_t i = 3; i < clear_rgba.size(); i += 4) clear_rgba[i] = 255;
            clear_depth.assign((size_t) width * height, 1e9f);
            frame_px.resize((size_t) width * height * 4);
        }
        int buf = 0;
        std::vector<double> ms;
        int64_t guided_rows = 0, added = 0, resampled = 0;
        for (int f = -2; f < frames; ++f) {  // two untimed warm-up frames
            set_pose(rend.camera, ((f % poses) + poses) % poses, poses);
            if (interop) {
                mnv_array_upload(ca[buf * 2], clear_rgba.data(), (size_t) width * 4, height);
                mnv_array_upload(ca[buf * 2 + 1], clear_depth.data(), (size_t) width * 4, height);
            }
            const auto t0 = std::chrono::steady_clock::now();
            rend.render();
            const uint8_t *px = interop ? nullptr : rend.frame_host();
            if (interop) {
                mnv_array_download(frame_px.data(), ca[buf * 2], (size_t) width * 4, height);
                buf ^= 1;
            }
            const auto t1 = std::chrono::steady_clock::now();
            (void) px;
            if (f >= 0) {
                ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
                guided_rows += rend.last_frame.guided_rows;
                added += rend.last_frame.added;
                resampled += rend.last_frame.resampled;
            }
        }
        if (!save_path.empty()) {
            if (gpus > 1 || replicas > 1) throw std::runtime_error("--save is not available with a replica group");
            tree.download();
            tree.save(save_path);
        }
        const uint8_t *px = interop ? frame_px.data() : rend.frame_host();
        const size_t nbytes = (size_t) width * height * 4;
        if (!out_ppm.empty()) {
            std::ofstream o(out_ppm, std::ios::binary);
            o << "P6\n" << width << " " << height << "\n255\n";
            for (size_t i = 0; i < (size_t) width * height; ++i) o.write(reinterpret_cast<const char *>(px + 4 * i), 3);
        }
        if (!out_raw.empty()) {
            std::ofstream o(out_raw, std::ios::binary);
            o.write(reinterpret_cast<const char *>(px), (std::streamsize) nbytes);
        }
        std::vector<double> sorted = ms;
        std::sort(sorted.begin(), sorted.end());
        double mean = 0;
        for (double v : ms) mean += v;
        mean /= ms.empty() ? 1 : (double) ms.size();
        const double med = sorted.empty() ? 0 : sorted[sorted.size() / 2];
        std::printf("{\"backend\": \"%s\", \"width\": %d, \"height\": %d, \"frames\": %d, \"ms_per_frame_mean\": %.4f, "
                    "\"ms_per_frame_median\": %.4f, \"fps_median\": %.2f, \"mrays_per_s_median\": %.2f, "
                    "\"capacity\": %d, \"max_capacity\": %ld, \"guided_rows\": %lld, \"nodes_added\": %lld, "
                    "\"leaves_resampled\": %lld, \"frame_hash\": \"%016" PRIx64 "\", \"fx\": %.9g, \"cx\": %.9g, "
                    "\"cy\": %.9g, \"c2w\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]}\n",
                    rend.get_backend(), width, height, frames, mean, med, med > 0 ? 1000.0 / med : 0.0,
                    med > 0 ? (double) width * height / med / 1e3 : 0.0, tree.capacity, max_cap,
                    (long long) guided_rows, (long long) added, (long long) resampled, checksum(px, nbytes),
                    rend.camera.fx, rend.camera.cx, rend.camera.cy, rend.camera.transform[0][0],
                    rend.camera.transform[0][1], rend.camera.transform[0][2], rend.camera.transform[1][0],
                    rend.camera.transform[1][1], rend.camera.transform[1][2], rend.camera.transform[2][0],
                    rend.camera.transform[2][1], rend.camera.transform[2][2], rend.camera.transform[3][0],
                    rend.camera.transform[3][1], rend.camera.transform[3][2]);
        if (interop) rend.set_interop_surfaces(nullptr);  // drops the surface objects before the arrays go
        for (void *a : ca) mnv_array_destroy(a);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "mnv_headless: %s\n", e.what());
        return 1;
    }
    return 0;
}
