/* Minimal CL/GL interop declarations (Khronos cl_gl.h subset) for compiling
 * the reference's cl.c headless.  The GL sharing path is never taken: the shim
 * hides cl_khr_gl_sharing from the extension string. */
#ifndef FOSPHOR_STUB_CL_GL_H
#define FOSPHOR_STUB_CL_GL_H
#include <CL/cl.h>
#include <GL/gl.h>

#define CL_GL_CONTEXT_KHR   0x2008
#define CL_GLX_DISPLAY_KHR  0x200A
#define CL_WGL_HDC_KHR      0x200B

cl_mem clCreateFromGLBuffer(cl_context, cl_mem_flags, GLuint, cl_int *);
cl_mem clCreateFromGLTexture(cl_context, cl_mem_flags, GLenum, GLint, GLuint, cl_int *);
cl_mem clCreateFromGLTexture2D(cl_context, cl_mem_flags, GLenum, GLint, GLuint, cl_int *);
cl_int clEnqueueAcquireGLObjects(cl_command_queue, cl_uint, const cl_mem *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueReleaseGLObjects(cl_command_queue, cl_uint, const cl_mem *, cl_uint, const cl_event *, cl_event *);
#endif
