// viewer::Camera — the public surface of the reference's include/camera.hpp:12-87
// (pose vectors, intrinsics, drag helpers, has_changed, _update).  The reference exposes glm
// types; when glm is on the include path (`-DMNV_USE_GLM`) the same types are used, otherwise
// the small vec/mat types below provide the members the viewer touches (operator[], x/y/z).
// There is no device upload: the 4x3 c2w travels to the kernels as a launch parameter
// (the reference cudaMemcpyAsync's 48 bytes per frame, src/camera.cpp:113-123).
#pragma once

#include <memory>

#ifdef MNV_USE_GLM
#include "glm/mat4x3.hpp"
#include "glm/mat4x4.hpp"
#include "glm/vec2.hpp"
#include "glm/vec3.hpp"
namespace viewer {
using vec2 = glm::vec2;
using vec3 = glm::vec3;
using mat4x3 = glm::mat4x3;
using mat4x4 = glm::mat4x4;
}  // namespace viewer
#else
namespace viewer {
struct vec2 {
    float x = 0, y = 0;
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
    float v[4] = {0, 0, 0, 0};
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};
struct mat4x3 {  // 4 columns of vec3 (column-major like glm::mat4x3)
    vec3 c[4];
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
struct mat4x4 {
    vec4 c[4];
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
}  // namespace viewer
#endif

struct mnv_camera;

namespace viewer {

struct Camera {
    Camera(int width = 256, int height = 256, float fx = 1111.f, float fy = -1.f, float cx = -1.f,
           float cy = -1.f);
    ~Camera();

    void begin_drag(float x, float y, bool is_pan, bool about_origin);
    void drag_update(float x, float y);
    void end_drag();
    bool is_dragging() const;
    void move(const vec3 &xyz);
    bool has_changed();

    vec3 v_back, v_world_up, center;
    vec3 origin;
    vec3 v_up, v_right;
    mat4x3 transform;  // C2W: right, up, back, center
    mat4x4 K;
    mat4x4 w2c;
    int width, height;
    float fx, fy;
    float cx, cy;
    float default_fx, default_fy;
    float default_cx, default_cy;
    float movement_speed = 1.f;
    struct {
        float *transform = nullptr;  // kept for source compatibility; never allocated here
    } device;

    void _update(bool transform_from_vecs = true, bool copy_cuda = true);

    // POD view handed to the C-ABI (mnv_b200.h)
    void fill(mnv_camera &out) const;

   private:
    struct DragState;
    std::unique_ptr<DragState> drag_state_;
    bool has_changed_ = true;
    bool transform_changed_ = false;
    float last_fx = 0, last_fy = 0;
    int last_width = 0, last_height = 0;
};

}  // namespace viewer
