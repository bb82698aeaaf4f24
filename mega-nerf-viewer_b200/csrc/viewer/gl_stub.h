// Declarations of the handful of OpenGL 4.5 (direct state access) and CUDA-GL interop entry points gl_interop.cpp
// uses, for machines WITHOUT GL development packages (this build image has no GL/glew.h, GL/gl.h or
// cuda_gl_interop.h dependencies installed): `make gl-check` compiles gl_interop.cpp against these so that the
// presentation shim stays type-checked.  A real viewer build defines MNV_WITH_GL without MNV_GL_STUB_HEADERS and
// gets the same names from <GL/glew.h> / <cuda_gl_interop.h>.  Values are the Khronos registry's.
#pragma once
#include <cuda_runtime_api.h>

typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef double GLdouble;
typedef unsigned int GLbitfield;
typedef unsigned char GLboolean;

#define GL_RENDERBUFFER 0x8D41
#define GL_FRAMEBUFFER 0x8D40
#define GL_COLOR_ATTACHMENT0 0x8CE0
#define GL_COLOR_ATTACHMENT1 0x8CE1
#define GL_DEPTH_ATTACHMENT 0x8D00
#define GL_RGBA8 0x8058
#define GL_R32F 0x822E
#define GL_DEPTH_COMPONENT32F 0x8CAC
#define GL_COLOR 0x1800
#define GL_DEPTH 0x1801
#define GL_COLOR_BUFFER_BIT 0x00004000
#define GL_NEAREST 0x2600
#define GL_TRUE 1

extern "C" {
void glCreateRenderbuffers(GLsizei n, GLuint *renderbuffers);
void glCreateFramebuffers(GLsizei n, GLuint *framebuffers);
void glDeleteRenderbuffers(GLsizei n, const GLuint *renderbuffers);
void glDeleteFramebuffers(GLsizei n, const GLuint *framebuffers);
void glNamedFramebufferRenderbuffer(GLuint framebuffer, GLenum attachment, GLenum renderbuffertarget, GLuint renderbuffer);
void glNamedFramebufferDrawBuffers(GLuint framebuffer, GLsizei n, const GLenum *bufs);
void glNamedRenderbufferStorage(GLuint renderbuffer, GLenum internalformat, GLsizei width, GLsizei height);
void glClearNamedFramebufferfv(GLuint framebuffer, GLenum buffer, GLint drawbuffer, const GLfloat *value);
void glClearDepth(GLdouble depth);
void glDepthMask(GLboolean flag);
void glBindFramebuffer(GLenum target, GLuint framebuffer);
void glNamedFramebufferReadBuffer(GLuint framebuffer, GLenum src);
void glBlitNamedFramebuffer(GLuint readFramebuffer, GLuint drawFramebuffer, GLint srcX0, GLint srcY0, GLint srcX1,
                            GLint srcY1, GLint dstX0, GLint dstY0, GLint dstX1, GLint dstY1, GLbitfield mask,
                            GLenum filter);
// cuda_gl_interop.h (exported by libcudart)
cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource **resource, GLuint image, GLenum target,
                                        unsigned int flags);
}
