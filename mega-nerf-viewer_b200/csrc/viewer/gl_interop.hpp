// viewer::GlPresenter — the OpenGL half of the reference's VolumeRenderer::Impl
// (src/renderer/cuda_renderer.cpp:43-66 start, :383-458 resize, :70-95 / :156-162 per frame), kept OUT of
// VolumeRenderer so that the render path has no GL dependency: a double-buffered off-screen framebuffer whose
// RGBA8 colour and R32F "fake depth" renderbuffers are registered with CUDA and handed to
// VolumeRenderer::set_interop_surfaces, cleared / mapped before each frame and blitted (Y flipped) to the window
// afterwards.  The viewer's frame loop becomes
//
//     presenter.begin_frame();      // clear, [caller draws meshes / the wireframe into presenter.framebuffer()], map
//     renderer.render();            // kernels composite into the mapped surfaces (one GPU or a replica group)
//     presenter.end_frame();        // unmap, blit to the default framebuffer, swap buffers
//
// Built only with -DMNV_WITH_GL (needs a GL 4.5 context, GLEW and cuda_gl_interop.h; see gl_stub.h for the
// declarations-only check build this image uses).
#pragma once
#ifdef MNV_WITH_GL

#include <array>

#include "renderer.hpp"

struct cudaGraphicsResource;

namespace viewer {

class GlPresenter {
   public:
    explicit GlPresenter(VolumeRenderer &renderer);
    ~GlPresenter();
    GlPresenter(const GlPresenter &) = delete;
    GlPresenter &operator=(const GlPresenter &) = delete;

    // (Re)allocate the renderbuffers for a width x height window, register them with CUDA and pass the mapped
    // arrays to the renderer (which also rescales the camera intrinsics, cuda_renderer.cpp:383-427).
    void resize(int width, int height);
    // Clear colour to the background, fake depth to 1e9 and the depth buffer; the framebuffer of this frame is
    // then open for the caller's rasterised geometry (wireframe, meshes), whose colour the kernels composite over
    // and whose R32F depth clips the rays.  Ends by mapping the frame's two resources for CUDA.
    void begin_frame();
    // Call between begin_frame's clears and its map if geometry has to be drawn: begin_frame(draw) does both.
    template <typename DrawFn>
    void begin_frame(DrawFn draw) {
        clear();
        bind();
        draw();
        unbind();
        map();
    }
    // Unmap, blit the colour attachment to the default framebuffer with the Y flip (row 0 of the surface is the top
    // image row, cuda_renderer.cpp:159-161) and switch to the other buffer.
    void end_frame();
    unsigned framebuffer() const { return fb_[buf_]; }

   private:
    void clear();
    void bind();
    void unbind();
    void map();
    void unregister_all();

    VolumeRenderer &rend_;
    std::array<unsigned, 2> fb_{}, color_rb_{}, fake_depth_rb_{}, depth_rb_{};
    std::array<cudaGraphicsResource *, 4> res_{};  // colour0, depth0, colour1, depth1 (Impl::cgr order)
    int buf_ = 0, width_ = 0, height_ = 0;
    bool mapped_ = false;
};

}  // namespace viewer
#endif  // MNV_WITH_GL
