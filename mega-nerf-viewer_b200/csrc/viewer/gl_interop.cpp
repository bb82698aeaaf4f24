// See gl_interop.hpp.  Reference: src/renderer/cuda_renderer.cpp:29-66, :70-95, :156-162, :383-458.
#ifdef MNV_WITH_GL
#include "gl_interop.hpp"

#ifdef MNV_GL_STUB_HEADERS
#include "gl_stub.h"
#else
#include <GL/glew.h>
#include <cuda_gl_interop.h>
#endif
#include <cuda_runtime_api.h>

#include <stdexcept>
#include <string>

namespace viewer {
namespace {
void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

GlPresenter::GlPresenter(VolumeRenderer &renderer) : rend_(renderer) {
    glCreateRenderbuffers(2, color_rb_.data());
    // the depth attachment cannot be read from CUDA: rasterised geometry also writes its linear depth into an
    // R32F colour attachment, which the kernels read as t_max (renderer_kernel.cu:277-280)
    glCreateRenderbuffers(2, fake_depth_rb_.data());
    glCreateRenderbuffers(2, depth_rb_.data());
    glCreateFramebuffers(2, fb_.data());
    for (int i = 0; i < 2; ++i) {
        glNamedFramebufferRenderbuffer(fb_[i], GL_COLOR_ATTACHMENT0, GL_RENDERBUFFER, color_rb_[i]);
        glNamedFramebufferRenderbuffer(fb_[i], GL_COLOR_ATTACHMENT1, GL_RENDERBUFFER, fake_depth_rb_[i]);
        glNamedFramebufferRenderbuffer(fb_[i], GL_DEPTH_ATTACHMENT, GL_RENDERBUFFER, depth_rb_[i]);
        const GLenum bufs[] = {GL_COLOR_ATTACHMENT0, GL_COLOR_ATTACHMENT1};
        glNamedFramebufferDrawBuffers(fb_[i], 2, bufs);
    }
}

GlPresenter::~GlPresenter() {
    rend_.set_interop_surfaces(nullptr);
    unregister_all();
    glDeleteRenderbuffers(2, color_rb_.data());
    glDeleteRenderbuffers(2, fake_depth_rb_.data());
    glDeleteRenderbuffers(2, depth_rb_.data());
    glDeleteFramebuffers(2, fb_.data());
}

void GlPresenter::unregister_all() {
    for (auto &r : res_) {
        if (r) cudaGraphicsUnregisterResource(r);
        r = nullptr;
    }
}

void GlPresenter::resize(int width, int height) {
    if (width == width_ && height == height_) return;
    rend_.set_interop_surfaces(nullptr);  // the arrays of the old size are about to disappear
    unregister_all();
    rend_.resize(width, height);
    const unsigned flags = cudaGraphicsRegisterFlagsSurfaceLoadStore | cudaGraphicsRegisterFlagsWriteDiscard;
    for (int i = 0; i < 2; ++i) {
        glNamedRenderbufferStorage(color_rb_[i], GL_RGBA8, width, height);
        glNamedRenderbufferStorage(fake_depth_rb_[i], GL_R32F, width, height);
        glNamedRenderbufferStorage(depth_rb_[i], GL_DEPTH_COMPONENT32F, width, height);
        const GLenum bufs[] = {GL_COLOR_ATTACHMENT0, GL_COLOR_ATTACHMENT1};
        glNamedFramebufferDrawBuffers(fb_[i], 2, bufs);
        ck(cudaGraphicsGLRegisterImage(&res_[i * 2], color_rb_[i], GL_RENDERBUFFER, flags), "register colour");
        ck(cudaGraphicsGLRegisterImage(&res_[i * 2 + 1], fake_depth_rb_[i], GL_RENDERBUFFER, flags), "register depth");
    }
    // the arrays behind the renderbuffers do not move while they stay registered: fetch them once
    void *arrays[4];
    ck(cudaGraphicsMapResources(4, res_.data(), 0), "map");
    for (int i = 0; i < 4; ++i) {
        cudaArray_t a = nullptr;
        ck(cudaGraphicsSubResourceGetMappedArray(&a, res_[i], 0, 0), "mapped array");
        arrays[i] = a;
    }
    ck(cudaGraphicsUnmapResources(4, res_.data(), 0), "unmap");
    rend_.set_interop_surfaces(arrays);
    width_ = width;
    height_ = height;
    buf_ = 0;
}

void GlPresenter::clear() {
    const GLfloat bg = rend_.options.background_brightness;
    const GLfloat colour[] = {bg, bg, bg, 1.f};
    const GLfloat depth_inf = 1e9f;
    glClearDepth(1.0);
    glClearNamedFramebufferfv(fb_[buf_], GL_COLOR, 0, colour);
    glClearNamedFramebufferfv(fb_[buf_], GL_COLOR, 1, &depth_inf);
    glClearNamedFramebufferfv(fb_[buf_], GL_DEPTH, 0, &depth_inf);
}

void GlPresenter::bind() {
    glDepthMask(GL_TRUE);
    glBindFramebuffer(GL_FRAMEBUFFER, fb_[buf_]);
}

void GlPresenter::unbind() { glBindFramebuffer(GL_FRAMEBUFFER, 0); }

void GlPresenter::map() {
    // GL -> CUDA hand-off of this frame's colour + fake-depth images (cuda_renderer.cpp:95).  The render kernels run
    // on the renderer's own non-blocking stream(s) — several devices with a replica group — so the hand-off is
    // completed on the host rather than ordered on one stream (the reference's loop calls glFinish per frame anyway,
    // main.cpp:614).
    ck(cudaGraphicsMapResources(2, &res_[buf_ * 2], 0), "map frame");
    ck(cudaStreamSynchronize(0), "map frame");
    mapped_ = true;
}

void GlPresenter::begin_frame() {
    clear();
    map();
}

void GlPresenter::end_frame() {
    if (mapped_) ck(cudaGraphicsUnmapResources(2, &res_[buf_ * 2], 0), "unmap frame");
    mapped_ = false;
    glNamedFramebufferReadBuffer(fb_[buf_], GL_COLOR_ATTACHMENT0);
    glBlitNamedFramebuffer(fb_[buf_], 0, 0, 0, width_, height_, 0, height_, width_, 0, GL_COLOR_BUFFER_BIT, GL_NEAREST);
    buf_ ^= 1;
}

}  // namespace viewer
#endif  // MNV_WITH_GL
