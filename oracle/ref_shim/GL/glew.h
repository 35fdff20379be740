/* Minimal stand-in for <GL/glew.h>: the reference's octree.c / model.c include it
 * only for the scalar typedefs (octree.c L8).  No GL is linked. */
#ifndef QB_REF_SHIM_GLEW_H
#define QB_REF_SHIM_GLEW_H
typedef int          GLint;
typedef unsigned int GLuint;
typedef float        GLfloat;
typedef unsigned int GLenum;
#endif
