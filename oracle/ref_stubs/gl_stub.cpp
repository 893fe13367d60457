// Definitions for oracle/ref_stubs/GL/glu.h (never executed on the particle->surface path).
#include <GL/glu.h>
extern "C" {
void glGenTextures(GLsizei n, GLuint* t) { for (GLsizei i = 0; i < n; ++i) t[i] = 0; }
void glBindTexture(GLenum, GLuint) {}
void glDeleteTextures(GLsizei, const GLuint*) {}
void glPixelStorei(GLenum, GLint) {}
GLint gluBuild2DMipmaps(GLenum, GLint, GLsizei, GLsizei, GLenum, GLenum, const void*) { return 0; }
}
