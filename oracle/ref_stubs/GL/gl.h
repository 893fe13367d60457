// Stub of OUR OWN (not reference code): the GL types and two enums cuda_gl_interop.h and the reference CUDAQuickSurf.cu mention; no GL here.
#pragma once
typedef unsigned int GLuint; typedef unsigned int GLenum; typedef int GLint; typedef int GLsizei; typedef long GLsizeiptr;
#define GL_ARRAY_BUFFER 0x8892
#define GL_DYNAMIC_DRAW 0x88E8
