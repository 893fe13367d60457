// Minimal GL/GLU declarations so that geometry_calls_gl/src/CallTriMeshDataGL.cpp compiles without an
// OpenGL SDK. OUR stub; the functions are only referenced by Material texture loading, which the
// particle->surface path never executes. Definitions live in oracle/ref_stubs/gl_stub.cpp.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
#define GL_TEXTURE_2D 0x0DE1
#define GL_UNPACK_ALIGNMENT 0x0CF5
#define GL_UNPACK_ROW_LENGTH 0x0CF2
#define GL_PACK_ALIGNMENT 0x0D05
#define GL_PACK_ROW_LENGTH 0x0D02
#define GL_RGB 0x1907
#define GL_RGBA 0x1908
#define GL_UNSIGNED_BYTE 0x1401
#define GL_FLOAT 0x1406
#define GL_LUMINANCE 0x1909
#define GL_LUMINANCE_ALPHA 0x190A
void glGenTextures(GLsizei n, GLuint* textures);
void glBindTexture(GLenum target, GLuint texture);
void glDeleteTextures(GLsizei n, const GLuint* textures);
void glPixelStorei(GLenum pname, GLint param);
GLint gluBuild2DMipmaps(GLenum target, GLint internalFormat, GLsizei width, GLsizei height, GLenum format,
    GLenum type, const void* data);
#ifdef __cplusplus
}
#endif
