/* stub GL types: the reference's cl.c only needs these to compile */
#ifndef FOSPHOR_STUB_GL_H
#define FOSPHOR_STUB_GL_H
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
#define GL_TEXTURE_2D 0x0DE1
#endif
