#include "camera.hpp"

#include <algorithm>
#include <cmath>

#include "../../../include/mnv_b200.h"

namespace viewer {
namespace {
#ifdef MNV_USE_GLM
inline float X(const vec3 &v) { return v.x; }
#endif
vec3 add(const vec3 &a, const vec3 &b) { return vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
vec3 sub(const vec3 &a, const vec3 &b) { return vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
vec3 mul(const vec3 &a, float s) { return vec3(a[0] * s, a[1] * s, a[2] * s); }
float dot(const vec3 &a, const vec3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
vec3 cross(const vec3 &a, const vec3 &b) {
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
vec3 normalize(const vec3 &a) { return mul(a, 1.f / std::sqrt(dot(a, a))); }
// rotate v about unit axis k by angle (Rodrigues) — what glm::rotate(mat4(1), angle, axis) * v does
vec3 rotate(const vec3 &v, float angle, const vec3 &axis) {
    const vec3 k = normalize(axis);
    const float c = std::cos(angle), s = std::sin(angle);
    return add(add(mul(v, c), mul(cross(k, v), s)), mul(k, dot(k, v) * (1.f - c)));
}
bool differs(const vec3 &a, const vec3 &b) { return a[0] != b[0] || a[1] != b[1] || a[2] != b[2]; }
}  // namespace

struct Camera::DragState {
    bool is_dragging = false, is_panning = false, about_origin = false;
    float start_x = 0, start_y = 0;
    vec3 drag_start_back, drag_start_right, drag_start_up, drag_start_center, drag_start_origin;
};

// Defaults of src/camera.cpp:29-46.
Camera::Camera(int width, int height, float fx, float fy, float cx, float cy)
    : width(width),
      height(height),
      fx(fx),
      fy(fy < 0.f ? fx : fy),
      cx(cx < 0.f ? (float) (width / 2) : cx),
      cy(cy < 0.f ? (float) (height / 2) : cy),
      default_fx(fx),
      default_fy(fy < 0.f ? fx : fy),
      default_cx(cx),
      default_cy(cy),
      drag_state_(std::make_unique<DragState>()) {
    center = vec3(-3.55f, 0.0f, 3.55f);
    v_back = vec3(-0.7071068f, 0.0f, 0.7071068f);
    v_world_up = vec3(0.0f, 0.0f, 1.0f);
    origin = vec3(0.0f, 0.0f, 0.0f);
    _update();
}

Camera::~Camera() = default;

// src/camera.cpp:54-130 without the device copy.
void Camera::_update(bool transform_from_vecs, bool copy_cuda) {
    if (transform_from_vecs) {
        v_back = normalize(v_back);
        v_right = normalize(cross(v_world_up, v_back));
        v_up = cross(v_back, v_right);
        const vec3 cols[4] = {v_right, v_up, v_back, center};
        for (int i = 0; i < 4; ++i) {
            if (differs(transform[i], cols[i])) transform_changed_ = true;
            transform[i] = cols[i];
        }
    }
    if (last_fx != fx || last_fy != fy || last_width != width || last_height != height) {
        transform_changed_ = true;
        last_fx = fx;
        last_fy = fy;
        last_width = width;
        last_height = height;
    }
    const float CLIP_NEAR = 1e-3f;
    K = mat4x4();
    K[0][0] = fx / (0.5f * width);
    K[1][1] = -fy / (0.5f * height);
    K[2][2] = -1.f;
    K[2][3] = -1.f;
    K[3][2] = -2 * CLIP_NEAR;
    // w2c = affine inverse of [R | t]: R^T, -R^T t
    w2c = mat4x4();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) w2c[c][r] = transform[r][c];
    for (int r = 0; r < 3; ++r) w2c[3][r] = -dot(transform[r], transform[3]);
    w2c[3][3] = 1.f;
    if (copy_cuda && transform_changed_) {
        has_changed_ = true;
        transform_changed_ = false;
    }
}

void Camera::fill(mnv_camera &out) const {
    out.width = width;
    out.height = height;
    out.fx = fx;
    out.fy = fy;
    out.cx = cx;
    out.cy = cy;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 3; ++r) out.c2w[c * 3 + r] = transform[c][r];
}

void Camera::begin_drag(float x, float y, bool is_pan, bool about_origin) {
    DragState &d = *drag_state_;
    d.is_dragging = true;
    d.start_x = x;
    d.start_y = y;
    d.drag_start_back = v_back;
    d.drag_start_right = v_right;
    d.drag_start_up = v_up;
    d.drag_start_center = center;
    d.drag_start_origin = origin;
    d.is_panning = is_pan;
    d.about_origin = about_origin;
}

// src/camera.cpp:143-187.
void Camera::drag_update(float x, float y) {
    DragState &d = *drag_state_;
    if (!d.is_dragging) return;
    const float s = -2.f * movement_speed / (float) std::max(width, height);
    float dx = (x - d.start_x) * s, dy = (y - d.start_y) * s;
    if (d.is_panning) {
        const vec3 shift = sub(mul(d.drag_start_right, dx), mul(d.drag_start_up, dy));
        center = add(d.drag_start_center, shift);
        if (d.about_origin) origin = add(d.drag_start_origin, shift);
        return;
    }
    if (d.about_origin) {
        dx = -dx;
        dy = -dy;
    }
    const vec3 back_tmp = rotate(d.drag_start_back, -dy, d.drag_start_right);
    if (dot(cross(v_world_up, back_tmp), d.drag_start_right) < 0.f) return;  // no flip over the pole
    const float yaw = std::fmod(-dx, 2.f * 3.14159265358979323846f);
    auto apply = [&](const vec3 &v) { return rotate(rotate(v, -dy, d.drag_start_right), yaw, v_world_up); };
    v_back = normalize(apply(d.drag_start_back));
    if (d.about_origin) center = add(apply(sub(d.drag_start_center, origin)), origin);
    _update(true, false);
}

bool Camera::is_dragging() const { return drag_state_->is_dragging; }
void Camera::end_drag() { drag_state_->is_dragging = false; }

void Camera::move(const vec3 &xyz) {
    center = add(center, mul(xyz, movement_speed));
    if (drag_state_->is_dragging)
        drag_state_->drag_start_center = add(drag_state_->drag_start_center, mul(xyz, movement_speed));
}

bool Camera::has_changed() {
    const bool r = has_changed_;
    has_changed_ = false;
    return r;
}

}  // namespace viewer
