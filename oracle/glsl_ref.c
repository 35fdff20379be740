/*
 * glsl_ref.c -- runs the reference's UNMODIFIED shaders (octree_vsh.c / octree_fsh.c)
 * headless on Mesa llvmpipe and dumps the frame.
 *
 * TEST INFRASTRUCTURE ONLY.  The shader sources are NOT in this repository: the
 * Makefile embeds them as binary blobs from /root/reference/src/qubatron/shaders/
 * into oracle/_ref/ (git-ignored).  This harness restates only the host side of
 *   /root/reference/src/qubatron/octree_glc.c L93-247 (program, 6 data textures
 *   8192 texels wide, render target), L263-305 (uniforms, quad), L361-498
 *   (texture addressing: linear index -> (i mod 8192, i / 8192))
 * because octree_glc.c itself needs SDL2/GLEW, which this image does not have.
 *
 * GL comes from the Mesa 18.1.9 libGL.so.1 (llvmpipe, GLX/xlib flavour) bundled
 * with Nsight Compute, dlopen'ed after oracle/_ref/libX11.so.6 / libXext.so.6
 * (x11_stub.c) so that no X server is needed.
 *
 * usage: glsl_ref <in.bin> <out.rgba> [mode] [repeat]
 *   in.bin : int64 hdr[8] = {nodes_s, nodes_d, points_s, points_d, W, H, maxlevel, shoot}
 *            float  u[14]  = {camfp3, angle3, light3, basecube4... see below}
 *            then oct_s, oct_d (int32 x12 per node), col_s, nrm_s, col_d, nrm_d (float x3)
 *   mode 0 : the shader exactly as shipped -> RGBA8 frame
 *   mode 1 : aux dump of the static model index  (output statement patched only)
 *   mode 2 : aux dump of the dynamic model index
 *   mode 3 : aux dump of the shadow visibility step(sqr, 15.0)
 *   The aux modes append statements that copy a value to the output; the traversal
 *   and shading code is untouched.
 *   mode 40 : mode 0 followed by the presentation pass (texquad shaders + crosshair) into a $QB_WINDOW = WxH window;
 *             the output is the window image
 *   mode 20 - 22 : skeleton_vsh.c through transform feedback (see skin_main below)
 *   mode 30 / 31 : particle_vsh.c / dust_vsh.c through transform feedback (see particle_main below)
 * prints one JSON line with renderer, version and per-frame seconds.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

extern const char _binary_octree_fsh_c_start[], _binary_octree_fsh_c_end[];
extern const char _binary_octree_vsh_c_start[], _binary_octree_vsh_c_end[];
extern const char _binary_texquad_vsh_c_start[], _binary_texquad_vsh_c_end[];
extern const char _binary_texquad_fsh_c_start[], _binary_texquad_fsh_c_end[];
extern const char _binary_skeleton_vsh_c_start[], _binary_skeleton_vsh_c_end[];
extern const char _binary_skeleton_fsh_c_start[], _binary_skeleton_fsh_c_end[];
extern const char _binary_particle_vsh_c_start[], _binary_particle_vsh_c_end[];
extern const char _binary_particle_fsh_c_start[], _binary_particle_fsh_c_end[];
extern const char _binary_dust_vsh_c_start[], _binary_dust_vsh_c_end[];
extern const char _binary_dust_fsh_c_start[], _binary_dust_fsh_c_end[];

#ifndef MESA_DIR
    #define MESA_DIR "/opt/nvidia/nsight-compute/2025.2.1/host/linux-desktop-glibc_2_11_3-x64/Mesa"
#endif

typedef unsigned int GLenum, GLuint, GLbitfield;
typedef int          GLint, GLsizei;
typedef float        GLfloat;
typedef char         GLchar;
typedef long         GLsizeiptr;

#define GL_FRAGMENT_SHADER 0x8B30
#define GL_VERTEX_SHADER 0x8B31
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_NEAREST 0x2600
#define GL_LINEAR 0x2601
#define GL_RGB32F 0x8815
#define GL_RGBA32I 0x8D82
#define GL_RGB 0x1907
#define GL_RGBA 0x1908
#define GL_RGBA_INTEGER 0x8D99
#define GL_FLOAT 0x1406
#define GL_INT 0x1404
#define GL_UNSIGNED_BYTE 0x1401
#define GL_TEXTURE0 0x84C0
#define GL_FRAMEBUFFER 0x8D40
#define GL_COLOR_ATTACHMENT0 0x8CE0
#define GL_FRAMEBUFFER_COMPLETE 0x8CD5
#define GL_COLOR_BUFFER_BIT 0x4000
#define GL_DEPTH_BUFFER_BIT 0x0100
#define GL_ARRAY_BUFFER 0x8892
#define GL_DYNAMIC_DRAW 0x88E8
#define GL_TRIANGLES 4
#define GL_BLEND 0x0BE2
#define GL_RENDERER 0x1F01
#define GL_VERSION 0x1F02
#define GL_PACK_ALIGNMENT 0x0D05
#define GL_POINTS 0
#define GL_SCISSOR_TEST 0x0C11
#define GL_STATIC_DRAW 0x88E4
#define GL_STREAM_READ 0x88E1
#define GL_RASTERIZER_DISCARD 0x8C89
#define GL_TRANSFORM_FEEDBACK_BUFFER 0x8C8E
#define GL_SEPARATE_ATTRIBS 0x8C8D
#define GL_UNPACK_ALIGNMENT 0x0CF5

#define GLX_RGBA 4
#define GLX_DOUBLEBUFFER 5
#define GLX_RED_SIZE 8
#define GLX_GREEN_SIZE 9
#define GLX_BLUE_SIZE 10
#define GLX_DEPTH_SIZE 12

static void* (*p_glXGetProcAddress)(const char*);
#define GLFN(ret, name, ...)                                                                                          \
    static ret (*name)(__VA_ARGS__);
GLFN(GLuint, glCreateShader, GLenum)
GLFN(void, glShaderSource, GLuint, GLsizei, const GLchar**, const GLint*)
GLFN(void, glCompileShader, GLuint)
GLFN(void, glGetShaderiv, GLuint, GLenum, GLint*)
GLFN(void, glGetShaderInfoLog, GLuint, GLsizei, GLsizei*, GLchar*)
GLFN(GLuint, glCreateProgram, void)
GLFN(void, glAttachShader, GLuint, GLuint)
GLFN(void, glBindAttribLocation, GLuint, GLuint, const GLchar*)
GLFN(void, glLinkProgram, GLuint)
GLFN(void, glGetProgramiv, GLuint, GLenum, GLint*)
GLFN(void, glGetProgramInfoLog, GLuint, GLsizei, GLsizei*, GLchar*)
GLFN(void, glUseProgram, GLuint)
GLFN(GLint, glGetUniformLocation, GLuint, const GLchar*)
GLFN(void, glUniform1i, GLint, GLint)
GLFN(void, glUniform2fv, GLint, GLsizei, const GLfloat*)
GLFN(void, glUniform3fv, GLint, GLsizei, const GLfloat*)
GLFN(void, glUniform4fv, GLint, GLsizei, const GLfloat*)
GLFN(void, glUniformMatrix4fv, GLint, GLsizei, unsigned char, const GLfloat*)
GLFN(void, glGenTextures, GLsizei, GLuint*)
GLFN(void, glBindTexture, GLenum, GLuint)
GLFN(void, glTexParameteri, GLenum, GLenum, GLint)
GLFN(void, glTexImage2D, GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void*)
GLFN(void, glActiveTexture, GLenum)
GLFN(void, glGenFramebuffers, GLsizei, GLuint*)
GLFN(void, glBindFramebuffer, GLenum, GLuint)
GLFN(void, glFramebufferTexture2D, GLenum, GLenum, GLenum, GLuint, GLint)
GLFN(GLenum, glCheckFramebufferStatus, GLenum)
GLFN(void, glViewport, GLint, GLint, GLsizei, GLsizei)
GLFN(void, glClearColor, GLfloat, GLfloat, GLfloat, GLfloat)
GLFN(void, glClear, GLbitfield)
GLFN(void, glGenBuffers, GLsizei, GLuint*)
GLFN(void, glBindBuffer, GLenum, GLuint)
GLFN(void, glBufferData, GLenum, GLsizeiptr, const void*, GLenum)
GLFN(void, glGenVertexArrays, GLsizei, GLuint*)
GLFN(void, glBindVertexArray, GLuint)
GLFN(void, glEnableVertexAttribArray, GLuint)
GLFN(void, glVertexAttribPointer, GLuint, GLint, GLenum, unsigned char, GLsizei, const void*)
GLFN(void, glDrawArrays, GLenum, GLint, GLsizei)
GLFN(void, glFinish, void)
GLFN(void, glReadPixels, GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, void*)
GLFN(const unsigned char*, glGetString, GLenum)
GLFN(GLenum, glGetError, void)
GLFN(void, glEnable, GLenum)
GLFN(void, glPixelStorei, GLenum, GLint)
GLFN(void, glScissor, GLint, GLint, GLsizei, GLsizei)
GLFN(void, glDisable, GLenum)
GLFN(void, glTransformFeedbackVaryings, GLuint, GLsizei, const GLchar* const*, GLenum)
GLFN(void, glBindBufferBase, GLenum, GLuint, GLuint)
GLFN(void, glBeginTransformFeedback, GLenum)
GLFN(void, glEndTransformFeedback, void)
GLFN(void, glGetBufferSubData, GLenum, long, GLsizeiptr, void*)

static void die(const char* m)
{
    fprintf(stderr, "glsl_ref: %s\n", m);
    exit(2);
}

static void* need(void* h, const char* n)
{
    void* p = h ? dlsym(h, n) : p_glXGetProcAddress(n);
    if (!p)
    {
        fprintf(stderr, "glsl_ref: missing symbol %s\n", n);
        exit(2);
    }
    return p;
}

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* replace the first occurrence of `needle`; the shader text is only ever EXTENDED around output statements */
static char* patch(char* src, const char* needle, const char* repl)
{
    char* at = strstr(src, needle);
    if (!at) die("aux patch point not found in the reference shader");
    size_t a = (size_t) (at - src), nl = strlen(needle), rl = strlen(repl), sl = strlen(src);
    char*  out = malloc(sl - nl + rl + 1);
    memcpy(out, src, a);
    memcpy(out + a, repl, rl);
    memcpy(out + a + rl, at + nl, sl - a - nl + 1);
    free(src);
    return out;
}

static GLuint compile(GLenum type, const char* src)
{
    GLuint s = glCreateShader(type);
    glShaderSource(s, 1, &src, NULL);
    glCompileShader(s);
    GLint ok = 0;
    glGetShaderiv(s, GL_COMPILE_STATUS, &ok);
    if (!ok)
    {
        char log[4096];
        glGetShaderInfoLog(s, sizeof(log), NULL, log);
        fprintf(stderr, "glsl_ref: shader compile failed:\n%s\n", log);
        exit(2);
    }
    return s;
}

/* 8192-texel-wide data texture, rows = ceil(texels / 8192) (octree_glc.c L412-428) */
static GLuint data_texture(int unit, const void* data, size_t texels, int is_int)
{
    size_t rows = (texels + 8191) / 8192;
    if (rows == 0) rows = 1;
    if (rows > 8192) die("scene exceeds the reference's 8192x8192 texture cap");
    size_t texel_bytes = is_int ? 16 : 12;
    char*  padded      = calloc(rows * 8192, texel_bytes);
    if (texels) memcpy(padded, data, texels * texel_bytes);
    GLuint t;
    glGenTextures(1, &t);
    glActiveTexture(GL_TEXTURE0 + unit);
    glBindTexture(GL_TEXTURE_2D, t);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
    if (is_int)
        glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA32I, 8192, (GLsizei) rows, 0, GL_RGBA_INTEGER, GL_INT, padded);
    else
        glTexImage2D(GL_TEXTURE_2D, 0, GL_RGB32F, 8192, (GLsizei) rows, 0, GL_RGB, GL_FLOAT, padded);
    free(padded);
    return t;
}

static void gl_context(void)
{
    /* ---- GL via the stubbed X11 + Mesa llvmpipe ---- */
    char  path[4096];
    char* self = realpath("/proc/self/exe", NULL);
    char* dir  = self ? self : strdup("./x");
    *strrchr(dir, '/') = 0;
    snprintf(path, sizeof(path), "%s/libX11.so.6", dir);
    void* x11 = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!x11) die(dlerror());
    snprintf(path, sizeof(path), "%s/libXext.so.6", dir);
    if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) die(dlerror());
    const char* mesa = getenv("QB_MESA_LIBGL") ? getenv("QB_MESA_LIBGL") : MESA_DIR "/libGL.so.1";
    void*       gl   = dlopen(mesa, RTLD_NOW | RTLD_GLOBAL);
    if (!gl) die(dlerror());

    void* (*open_display)(int, int) = need(x11, "qb_stub_open_display");
    int win_w = 64, win_h = 64; /* the default framebuffer; mode 40 presents into it */
    if (getenv("QB_WINDOW")) sscanf(getenv("QB_WINDOW"), "%dx%d", &win_w, &win_h);
    void* dpy = open_display(win_w, win_h);

    void* (*glXChooseVisual)(void*, int, int*)        = need(gl, "glXChooseVisual");
    void* (*glXCreateContext)(void*, void*, void*, int) = need(gl, "glXCreateContext");
    int (*glXMakeCurrent)(void*, unsigned long, void*) = need(gl, "glXMakeCurrent");
    p_glXGetProcAddress                                = need(gl, "glXGetProcAddressARB");

    int   attribs[] = {GLX_RGBA, GLX_RED_SIZE, 8, GLX_GREEN_SIZE, 8, GLX_BLUE_SIZE, 8, GLX_DEPTH_SIZE, 24,
                       GLX_DOUBLEBUFFER, 0};
    void* vi        = glXChooseVisual(dpy, 0, attribs);
    if (!vi) die("glXChooseVisual failed");
    void* ctx = glXCreateContext(dpy, vi, NULL, 1);
    if (!ctx) die("glXCreateContext failed");
    if (!glXMakeCurrent(dpy, 1, ctx)) die("glXMakeCurrent failed");

#define LOAD(n) n = need(NULL, #n);
    LOAD(glCreateShader) LOAD(glShaderSource) LOAD(glCompileShader) LOAD(glGetShaderiv) LOAD(glGetShaderInfoLog)
    LOAD(glCreateProgram) LOAD(glAttachShader) LOAD(glBindAttribLocation) LOAD(glLinkProgram) LOAD(glGetProgramiv)
    LOAD(glGetProgramInfoLog) LOAD(glUseProgram) LOAD(glGetUniformLocation) LOAD(glUniform1i) LOAD(glUniform2fv)
    LOAD(glUniform3fv) LOAD(glUniform4fv) LOAD(glUniformMatrix4fv) LOAD(glGenTextures) LOAD(glBindTexture)
    LOAD(glTexParameteri) LOAD(glTexImage2D) LOAD(glActiveTexture) LOAD(glGenFramebuffers) LOAD(glBindFramebuffer)
    LOAD(glFramebufferTexture2D) LOAD(glCheckFramebufferStatus) LOAD(glViewport) LOAD(glClearColor) LOAD(glClear)
    LOAD(glGenBuffers) LOAD(glBindBuffer) LOAD(glBufferData) LOAD(glGenVertexArrays) LOAD(glBindVertexArray)
    LOAD(glEnableVertexAttribArray) LOAD(glVertexAttribPointer) LOAD(glDrawArrays) LOAD(glFinish) LOAD(glReadPixels)
    LOAD(glGetString) LOAD(glGetError) LOAD(glEnable) LOAD(glPixelStorei) LOAD(glScissor) LOAD(glDisable) LOAD(glTransformFeedbackVaryings)
    LOAD(glBindBufferBase) LOAD(glBeginTransformFeedback) LOAD(glEndTransformFeedback) LOAD(glGetBufferSubData)
}

static char* embedded(const char* b, const char* e)
{
    size_t n = (size_t) (e - b);
    char*  t = malloc(n + 1);
    memcpy(t, b, n);
    t[n] = 0;
    return t;
}

/*
 * mode 20 / 21: the reference's skinning vertex program skeleton_vsh.c through transform feedback, restating the
 * host side of /root/reference/src/qubatron/skeleton_glc.c L77-137 (program, varyings, VAO), L222-251 (uniforms,
 * the GL_POINTS draw with rasterizer discard) and L257-300 (buffers).
 *   in.bin : int64 hdr[2] = {n, maxlevel}; float basesize, pad; float oldbones[80], newbones[80];
 *            float positions[3n], normals[3n]
 *   out    : int32 oct14[4n], oct54[4n], oct94[4n]; float normal_out[3n]
 *   mode 20: the shader exactly as shipped
 *   mode 21: main()'s `pnt` captured in place of normal_out
 *   mode 22: oldbone_rot_quat, bonesangle_rot_quat and (bone, has_axis) of bone pair $QB_SKIN_BONE in place of the
 *            three digit vectors -- the only driver-dependent values of the program (sin / cos / acos)
 */
static int skin_main(const char* in, const char* outp, int mode, int repeat)
{
    FILE* f = fopen(in, "rb");
    if (!f) die("cannot open input");
    int64_t hdr[2];
    float   fb[2], ob[80], nb[80];
    if (fread(hdr, 8, 2, f) != 2 || fread(fb, 4, 2, f) != 2 || fread(ob, 4, 80, f) != 80 || fread(nb, 4, 80, f) != 80)
        die("short skin input header");
    size_t n   = (size_t) hdr[0];
    float* pos = malloc((n ? n : 1) * 12);
    float* nrm = malloc((n ? n : 1) * 12);
    if (fread(pos, 12, n, f) != n || fread(nrm, 12, n, f) != n) die("short skin input body");
    fclose(f);

    gl_context();
    char* vsh = embedded(_binary_skeleton_vsh_c_start, _binary_skeleton_vsh_c_end);
    char* fsh = embedded(_binary_skeleton_fsh_c_start, _binary_skeleton_fsh_c_end);
    const GLchar* vary[4] = {"oct14", "oct54", "oct94", "normal_out"}; /* skeleton_glc.c L99-100 */
    if (mode == 21) /* `pnt` captured in place of normal_out (llvmpipe allows 4 separate varyings) */
    {
        vsh     = patch(vsh, "flat out vec3  normal_out;", "flat out vec3  normal_out;\nflat out vec3 qb_pnt;");
        vsh     = patch(vsh, "    vec4 cube = basecube;", "    qb_pnt = pnt;\n    vec4 cube = basecube;");
        vary[3] = "qb_pnt";
    }
    if (mode == 22) /* the two rotation quaternions of bone pair QB_SKIN_BONE, as this GL evaluates them */
    {
        char buf[512];
        int  bone = getenv("QB_SKIN_BONE") ? atoi(getenv("QB_SKIN_BONE")) : 0;
        vsh = patch(vsh, "flat out vec3  normal_out;",
                    "flat out vec3  normal_out;\nflat out vec4 qb_rq;\nflat out vec4 qb_aq;\nflat out ivec4 qb_id;");
        vsh = patch(vsh, "    vec3  corner_points[POINT_COUNT];",
                    "    qb_rq = vec4(0.0); qb_aq = vec4(0.0); qb_id = ivec4(-1, 0, 0, 0);\n"
                    "    vec3  corner_points[POINT_COUNT];");
        snprintf(buf, sizeof(buf),
                 "vec4 oldbone_rot_quat = quat_from_axis_angle(oldbone_norm, newbones[i].w);\n"
                 "if (i == %d) { qb_rq = oldbone_rot_quat; qb_id.x = i; }", bone);
        vsh = patch(vsh, "vec4 oldbone_rot_quat = quat_from_axis_angle(oldbone_norm, newbones[i].w);", buf);
        snprintf(buf, sizeof(buf),
                 "vec4 bonesangle_rot_quat = quat_from_axis_angle(normalize(bones_axis), bones_angle);\n"
                 "if (i == %d) { qb_aq = bonesangle_rot_quat; qb_id.y = 1; }", bone);
        vsh = patch(vsh, "vec4 bonesangle_rot_quat = quat_from_axis_angle(normalize(bones_axis), bones_angle);", buf);
        vary[0] = "qb_rq", vary[1] = "qb_aq", vary[2] = "qb_id";
    }
    GLuint prog = glCreateProgram();
    glAttachShader(prog, compile(GL_VERTEX_SHADER, vsh));
    glAttachShader(prog, compile(GL_FRAGMENT_SHADER, fsh));
    const int nvary = 4;
    glTransformFeedbackVaryings(prog, nvary, vary, GL_SEPARATE_ATTRIBS);
    glBindAttribLocation(prog, 0, "position");
    glBindAttribLocation(prog, 1, "normal");
    glLinkProgram(prog);
    GLint ok = 0;
    glGetProgramiv(prog, GL_LINK_STATUS, &ok);
    if (!ok)
    {
        char log[4096];
        glGetProgramInfoLog(prog, sizeof(log), NULL, log);
        fprintf(stderr, "glsl_ref: link failed:\n%s\n", log);
        return 2;
    }

    GLuint vin[2], vao, vout[4];
    glGenBuffers(2, vin);
    glGenVertexArrays(1, &vao);
    glBindVertexArray(vao);
    glBindBuffer(GL_ARRAY_BUFFER, vin[0]);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (n * 12), pos, GL_STATIC_DRAW);
    glEnableVertexAttribArray(0);
    glVertexAttribPointer(0, 3, GL_FLOAT, 0, sizeof(GLfloat) * 3, 0);
    glBindBuffer(GL_ARRAY_BUFFER, vin[1]);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (n * 12), nrm, GL_STATIC_DRAW);
    glEnableVertexAttribArray(1);
    glVertexAttribPointer(1, 3, GL_FLOAT, 0, sizeof(GLfloat) * 3, 0);
    const size_t osz[4] = {n * 16, n * 16, n * 16, n * 12};
    glGenBuffers(4, vout);
    for (int i = 0; i < nvary; i++)
    {
        glBindBuffer(GL_ARRAY_BUFFER, vout[i]);
        glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (osz[i] ? osz[i] : 16), NULL, GL_STREAM_READ);
    }
    glBindBuffer(GL_ARRAY_BUFFER, 0);

    glEnable(GL_RASTERIZER_DISCARD);
    glUseProgram(prog);
    GLfloat basecube[4] = {0.0f, fb[0], fb[0], fb[0]};
    glUniform4fv(glGetUniformLocation(prog, "oldbones"), 20, ob);
    glUniform4fv(glGetUniformLocation(prog, "newbones"), 20, nb);
    glUniform4fv(glGetUniformLocation(prog, "basecube"), 1, basecube);
    glUniform1i(glGetUniformLocation(prog, "maxlevel"), (GLint) hdr[1]);
    for (int i = 0; i < nvary; i++) glBindBufferBase(GL_TRANSFORM_FEEDBACK_BUFFER, (GLuint) i, vout[i]);

    printf("{\"renderer\": \"%s\", \"version\": \"%s\", \"mode\": %d, \"frame_s\": [", glGetString(GL_RENDERER),
           glGetString(GL_VERSION), mode);
    for (int r = 0; r < repeat; r++)
    {
        double t0 = now();
        glBeginTransformFeedback(GL_POINTS);
        glDrawArrays(GL_POINTS, 0, (GLsizei) n);
        glEndTransformFeedback();
        glFinish();
        printf("%s%.6f", r ? ", " : "", now() - t0);
    }
    printf("], \"gl_error\": %u}\n", glGetError());

    f = fopen(outp, "wb");
    if (!f) die("cannot open output");
    for (int i = 0; i < nvary; i++)
    {
        void* host = malloc(osz[i] ? osz[i] : 16);
        glBindBuffer(GL_TRANSFORM_FEEDBACK_BUFFER, vout[i]);
        glGetBufferSubData(GL_TRANSFORM_FEEDBACK_BUFFER, 0, (GLsizeiptr) osz[i], host);
        fwrite(host, 1, osz[i], f);
        free(host);
    }
    fclose(f);
    return 0;
}

/*
 * mode 30 / 31: the reference's particle (particle_vsh.c) and dust (dust_vsh.c) simulation steps through transform
 * feedback, restating the host side of /root/reference/src/qubatron/particle_glc.c L43-113 (program, the two
 * varyings, VAO, static octree sampler on unit 11 = octree_glc.c's texture), L118-156 (uniforms, GL_POINTS draw with
 * rasterizer discard, read-back) and dust_glc.c L103-135.
 *   in.bin : int64 hdr[3] = {n, maxlevel, nodes_s}; float f[4] = {basesize, campos.xyz};
 *            float pos[3n], spd[3n]; int32 oct_s[12 * nodes_s]
 *   out    : float pos_out[3n], spd_out[3n]
 */
static int particle_main(const char* in, const char* outp, int mode, int repeat)
{
    FILE* f = fopen(in, "rb");
    if (!f) die("cannot open input");
    int64_t hdr[3];
    float   fb[4];
    if (fread(hdr, 8, 3, f) != 3 || fread(fb, 4, 4, f) != 4) die("short particle input header");
    size_t   n = (size_t) hdr[0], nodes = (size_t) hdr[2];
    float*   pos = malloc((n ? n : 1) * 12);
    float*   spd = malloc((n ? n : 1) * 12);
    int32_t* oct = malloc((nodes ? nodes : 1) * 48);
    if (fread(pos, 12, n, f) != n || fread(spd, 12, n, f) != n || fread(oct, 48, nodes, f) != nodes)
        die("short particle input body");
    fclose(f);

    gl_context();
    char* vsh = mode == 30 ? embedded(_binary_particle_vsh_c_start, _binary_particle_vsh_c_end)
                           : embedded(_binary_dust_vsh_c_start, _binary_dust_vsh_c_end);
    char* fsh = mode == 30 ? embedded(_binary_particle_fsh_c_start, _binary_particle_fsh_c_end)
                           : embedded(_binary_dust_fsh_c_start, _binary_dust_fsh_c_end);
    GLuint prog = glCreateProgram();
    glAttachShader(prog, compile(GL_VERTEX_SHADER, vsh));
    glAttachShader(prog, compile(GL_FRAGMENT_SHADER, fsh));
    const GLchar* vary[2] = {"pos_out", "spd_out"};
    glTransformFeedbackVaryings(prog, 2, vary, GL_SEPARATE_ATTRIBS);
    glBindAttribLocation(prog, 0, "pos");
    glBindAttribLocation(prog, 1, "spd");
    glLinkProgram(prog);
    GLint ok = 0;
    glGetProgramiv(prog, GL_LINK_STATUS, &ok);
    if (!ok)
    {
        char log[4096];
        glGetProgramInfoLog(prog, sizeof(log), NULL, log);
        fprintf(stderr, "glsl_ref: link failed:\n%s\n", log);
        return 2;
    }
    glUseProgram(prog);
    glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
    if (mode == 30)
    {
        data_texture(11, oct, nodes * 3, 1);
        glUniform1i(glGetUniformLocation(prog, "octtexbuf_s"), 11); /* particle_glc.c L106-107 */
        glUniform1i(glGetUniformLocation(prog, "octtexbuf_d"), 12);
    }

    GLuint vin[2], vao, vout[2];
    glGenBuffers(2, vin);
    glGenVertexArrays(1, &vao);
    glBindVertexArray(vao);
    glBindBuffer(GL_ARRAY_BUFFER, vin[0]);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (n * 12), pos, GL_STATIC_DRAW);
    glEnableVertexAttribArray(0);
    glVertexAttribPointer(0, 3, GL_FLOAT, 0, sizeof(GLfloat) * 3, 0);
    glBindBuffer(GL_ARRAY_BUFFER, vin[1]);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (n * 12), spd, GL_STATIC_DRAW);
    glEnableVertexAttribArray(1);
    glVertexAttribPointer(1, 3, GL_FLOAT, 0, sizeof(GLfloat) * 3, 0);
    glGenBuffers(2, vout);
    for (int i = 0; i < 2; i++)
    {
        glBindBuffer(GL_ARRAY_BUFFER, vout[i]);
        glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr) (n ? n * 12 : 16), NULL, GL_STREAM_READ);
    }
    glBindBuffer(GL_ARRAY_BUFFER, 0);

    glEnable(GL_RASTERIZER_DISCARD);
    GLfloat basecube[4] = {0.0f, fb[0], fb[0], fb[0]};
    glUniform4fv(glGetUniformLocation(prog, "basecube"), 1, basecube);
    glUniform1i(glGetUniformLocation(prog, "maxlevel"), (GLint) hdr[1]);
    if (mode == 31) glUniform3fv(glGetUniformLocation(prog, "campos"), 1, fb + 1);
    for (int i = 0; i < 2; i++) glBindBufferBase(GL_TRANSFORM_FEEDBACK_BUFFER, (GLuint) i, vout[i]);

    printf("{\"renderer\": \"%s\", \"version\": \"%s\", \"mode\": %d, \"frame_s\": [", glGetString(GL_RENDERER),
           glGetString(GL_VERSION), mode);
    for (int r = 0; r < repeat; r++)
    {
        double t0 = now();
        glBeginTransformFeedback(GL_POINTS);
        glDrawArrays(GL_POINTS, 0, (GLsizei) n);
        glEndTransformFeedback();
        glFinish();
        printf("%s%.6f", r ? ", " : "", now() - t0);
    }
    printf("], \"gl_error\": %u}\n", glGetError());

    f = fopen(outp, "wb");
    if (!f) die("cannot open output");
    for (int i = 0; i < 2; i++)
    {
        void* host = malloc(n ? n * 12 : 16);
        glBindBuffer(GL_TRANSFORM_FEEDBACK_BUFFER, vout[i]);
        glGetBufferSubData(GL_TRANSFORM_FEEDBACK_BUFFER, 0, (GLsizeiptr) (n * 12), host);
        fwrite(host, 1, n * 12, f);
        free(host);
    }
    fclose(f);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 3) die("usage: glsl_ref in.bin out.rgba [mode] [repeat]");
    int mode   = argc > 3 ? atoi(argv[3]) : 0;
    int repeat = argc > 4 ? atoi(argv[4]) : 1;
    if (mode >= 20 && mode <= 22) return skin_main(argv[1], argv[2], mode, repeat);
    if (mode == 30 || mode == 31) return particle_main(argv[1], argv[2], mode, repeat);

    /* ---- input ---- */
    FILE* f = fopen(argv[1], "rb");
    if (!f) die("cannot open input");
    int64_t hdr[8];
    float   u[16];
    if (fread(hdr, 8, 8, f) != 8 || fread(u, 4, 16, f) != 16) die("short input header");
    int64_t nodes_s = hdr[0], nodes_d = hdr[1], pts_s = hdr[2], pts_d = hdr[3];
    int     W = (int) hdr[4], H = (int) hdr[5], maxlevel = (int) hdr[6], shoot = (int) hdr[7];
    /* u: camfp[0..2] angle[3..5] light[6..8] basecube[9..12] dims[13..14] */
    int32_t* oct_s = malloc((size_t) (nodes_s ? nodes_s : 1) * 48);
    int32_t* oct_d = malloc((size_t) (nodes_d ? nodes_d : 1) * 48);
    float*   col_s = malloc((size_t) (pts_s ? pts_s : 1) * 12);
    float*   nrm_s = malloc((size_t) (pts_s ? pts_s : 1) * 12);
    float*   col_d = malloc((size_t) (pts_d ? pts_d : 1) * 12);
    float*   nrm_d = malloc((size_t) (pts_d ? pts_d : 1) * 12);
    if (fread(oct_s, 48, nodes_s, f) != (size_t) nodes_s || fread(oct_d, 48, nodes_d, f) != (size_t) nodes_d ||
        fread(col_s, 12, pts_s, f) != (size_t) pts_s || fread(nrm_s, 12, pts_s, f) != (size_t) pts_s ||
        fread(col_d, 12, pts_d, f) != (size_t) pts_d || fread(nrm_d, 12, pts_d, f) != (size_t) pts_d)
        die("short input body");
    fclose(f);
    if (W > 2048 || H > 2048) die("the reference's render target is 2048x2048 (octree_glc.c L237)");

    gl_context();

    /* ---- shaders: the reference text, byte for byte (mode 0) ---- */
    size_t fl  = (size_t) (_binary_octree_fsh_c_end - _binary_octree_fsh_c_start);
    size_t vl  = (size_t) (_binary_octree_vsh_c_end - _binary_octree_vsh_c_start);
    char*  fsh = malloc(fl + 1);
    char*  vsh = malloc(vl + 1);
    memcpy(fsh, _binary_octree_fsh_c_start, fl);
    fsh[fl] = 0;
    memcpy(vsh, _binary_octree_vsh_c_start, vl);
    vsh[vl] = 0;
    const char* pack = "{ int qv = floatBitsToInt(qb_keep); fragColor = vec4(float(qv & 255) / 255.0, "
                       "float((qv >> 8) & 255) / 255.0, float((qv >> 16) & 255) / 255.0, "
                       "float((qv >> 24) & 255) / 255.0); }";
    if (mode != 0 && mode != 40)
    {
        /* qb_out is written at the leaf of whichever trace ran last; qb_keep latches the PRIMARY trace's value */
        fsh = patch(fsh, "out vec4 fragColor;",
                    "out vec4 fragColor;\nfloat qb_out = intBitsToFloat(-1);\nfloat qb_keep = intBitsToFloat(-1);");
        fsh = patch(fsh, "fragColor = col;", pack);
        if (mode == 1 || mode == 2)
            fsh = patch(fsh, "ctlres res = cube_trace_line(camfp, csv);",
                        "ctlres res = cube_trace_line(camfp, csv);\nqb_keep = qb_out;");
    }
    if (mode == 1)
        fsh = patch(fsh, "int docti = oct_from_octets_for_index(8, stck[level].docti, octtexbuf_d, level);",
                    "int docti = oct_from_octets_for_index(8, stck[level].docti, octtexbuf_d, level);\n"
                    "qb_out = intBitsToFloat(socti);");
    if (mode == 2)
        fsh = patch(fsh, "int docti = oct_from_octets_for_index(8, stck[level].docti, octtexbuf_d, level);",
                    "int docti = oct_from_octets_for_index(8, stck[level].docti, octtexbuf_d, level);\n"
                    "qb_out = intBitsToFloat(docti);");
    if (mode == 9) /* debug: dump any float expression of main() evaluated after the ray set-up (QB_DEBUG_EXPR) */
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "csv = quat_rotate(qx, csv);\nqb_keep = %s;",
                 getenv("QB_DEBUG_EXPR") ? getenv("QB_DEBUG_EXPR") : "csv.x");
        fsh = patch(fsh, "csv = quat_rotate(qx, csv);", buf);
    }
    if (mode == 7) /* debug: patch QB_PATCH_NEEDLE -> QB_PATCH_REPL anywhere (e.g. qb_out = ...), dump QB_DEBUG_EXPR at the end */
    {
        char buf[512];
        if (getenv("QB_PATCH_NEEDLE")) fsh = patch(fsh, getenv("QB_PATCH_NEEDLE"), getenv("QB_PATCH_REPL"));
        snprintf(buf, sizeof(buf), "col.z *= 0.7;\nqb_keep = %s;", getenv("QB_DEBUG_EXPR") ? getenv("QB_DEBUG_EXPR") : "qb_out");
        fsh = patch(fsh, "col.z *= 0.7;", buf);
    }
    if (mode == 8) /* debug: any float expression of main() evaluated after the shading (QB_DEBUG_EXPR) */
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "col.z *= 0.7;\nqb_keep = %s;", getenv("QB_DEBUG_EXPR") ? getenv("QB_DEBUG_EXPR") : "sqr");
        fsh = patch(fsh, "col.z *= 0.7;", buf);
    }
    if (mode == 3)
        fsh = patch(fsh, "col.z *= 0.7;", "col.z *= 0.7;\nqb_keep = intBitsToFloat(int(step(sqr, 15.0)));");

    GLuint prog = glCreateProgram();
    glAttachShader(prog, compile(GL_VERTEX_SHADER, vsh));
    glAttachShader(prog, compile(GL_FRAGMENT_SHADER, fsh));
    glBindAttribLocation(prog, 0, "position"); /* octree_glc.c L113 (before link here) */
    glLinkProgram(prog);
    GLint ok = 0;
    glGetProgramiv(prog, GL_LINK_STATUS, &ok);
    if (!ok)
    {
        char log[4096];
        glGetProgramInfoLog(prog, sizeof(log), NULL, log);
        fprintf(stderr, "glsl_ref: link failed:\n%s\n", log);
        return 2;
    }
    glUseProgram(prog);

    /* ---- data textures on units unif+1 as in octree_glc.c L371-408 ---- */
    glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
    glPixelStorei(GL_PACK_ALIGNMENT, 1);
    struct
    {
        const char* name;
        int         unit;
        const void* data;
        size_t      texels;
        int         is_int;
    } tex[6] = {{"coltexbuf_s", 7, col_s, (size_t) pts_s, 0},       {"coltexbuf_d", 8, col_d, (size_t) pts_d, 0},
                {"nrmtexbuf_s", 9, nrm_s, (size_t) pts_s, 0},       {"nrmtexbuf_d", 10, nrm_d, (size_t) pts_d, 0},
                {"octtexbuf_s", 11, oct_s, (size_t) nodes_s * 3, 1}, {"octtexbuf_d", 12, oct_d, (size_t) nodes_d * 3, 1}};
    for (int i = 0; i < 6; i++)
    {
        data_texture(tex[i].unit, tex[i].data, tex[i].texels, tex[i].is_int);
        glUniform1i(glGetUniformLocation(prog, tex[i].name), tex[i].unit);
    }

    /* ---- 2048x2048 RGBA8 render target (octree_glc.c L230-244) ---- */
    GLuint fbo, rt;
    glGenFramebuffers(1, &fbo);
    glBindFramebuffer(GL_FRAMEBUFFER, fbo);
    glGenTextures(1, &rt);
    glActiveTexture(GL_TEXTURE0);
    glBindTexture(GL_TEXTURE_2D, rt);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA, 2048, 2048, 0, GL_RGBA, GL_UNSIGNED_BYTE, 0);
    glBindTexture(GL_TEXTURE_2D, 0);
    glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, rt, 0);
    if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) die("framebuffer incomplete");

    /* ---- uniforms (octree_glc.c L263-284); ortho = m4_defaultortho(0,ow,0,oh,-10,10), mt_matrix_4d.c L179-210 ---- */
    float ow = u[13], oh = u[14];
    float proj[16] = {0};
    proj[0]        = 2.0f / (ow - 0.0f);
    proj[5]        = 2.0f / (oh - 0.0f);
    proj[10]       = -2.0f / (10.0f - -10.0f);
    proj[12]       = -(ow + 0.0f) / (ow - 0.0f);
    proj[13]       = -(oh + 0.0f) / (oh - 0.0f);
    proj[14]       = -(10.0f + -10.0f) / (10.0f - -10.0f);
    proj[15]       = 1.0f;
    glEnable(GL_BLEND); /* L251 (blend func stays ONE/ZERO) */
    glUniformMatrix4fv(glGetUniformLocation(prog, "projection"), 1, 0, proj);
    glUniform3fv(glGetUniformLocation(prog, "camfp"), 1, u + 0);
    glUniform3fv(glGetUniformLocation(prog, "angle_in"), 1, u + 3);
    glUniform3fv(glGetUniformLocation(prog, "light"), 1, u + 6);
    glUniform4fv(glGetUniformLocation(prog, "basecube"), 1, u + 9);
    glUniform2fv(glGetUniformLocation(prog, "dimensions"), 1, u + 13);
    glUniform1i(glGetUniformLocation(prog, "maxlevel"), maxlevel);
    glUniform1i(glGetUniformLocation(prog, "shoot"), shoot);

    GLuint vbo, vao;
    glGenBuffers(1, &vbo);
    glBindBuffer(GL_ARRAY_BUFFER, vbo);
    glGenVertexArrays(1, &vao);
    glBindVertexArray(vao);
    glEnableVertexAttribArray(0);
    glVertexAttribPointer(0, 3, GL_FLOAT, 0, sizeof(GLfloat) * 3, 0);
    GLfloat quad[] = {0.0f, 0.0f, 0.0f, ow, 0.0f, 0.0f, 0.0f, oh, 0.0f, 0.0f, oh, 0.0f, ow, 0.0f, 0.0f, ow, oh, 0.0f};

    printf("{\"renderer\": \"%s\", \"version\": \"%s\", \"mode\": %d, \"frame_s\": [", glGetString(GL_RENDERER),
           glGetString(GL_VERSION), mode);
    for (int r = 0; r < repeat; r++)
    {
        double t0 = now();
        glViewport(0, 0, (GLsizei) ow, (GLsizei) oh); /* L288 */
        glClearColor(0.0f, 0.0f, 0.0f, 0.0f);
        glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
        glBufferData(GL_ARRAY_BUFFER, sizeof(quad), quad, GL_DYNAMIC_DRAW);
        glDrawArrays(GL_TRIANGLES, 0, 6);
        glFinish();
        printf("%s%.6f", r ? ", " : "", now() - t0);
    }
    printf("], \"gl_error\": %u}\n", glGetError());

    if (mode == 40)
    {
        /* ---- presentation, octree_glc.c L308-351: the render target drawn as a textured quad (texquad_vsh.c /
         * texquad_fsh.c, LINEAR filter) into the window, then the 2 x 2 crosshair ---- */
        int ww = 64, wh = 64;
        if (getenv("QB_WINDOW")) sscanf(getenv("QB_WINDOW"), "%dx%d", &ww, &wh);
        GLuint tq = glCreateProgram();
        glAttachShader(tq, compile(GL_VERTEX_SHADER, embedded(_binary_texquad_vsh_c_start, _binary_texquad_vsh_c_end)));
        glAttachShader(tq, compile(GL_FRAGMENT_SHADER, embedded(_binary_texquad_fsh_c_start, _binary_texquad_fsh_c_end)));
        glBindAttribLocation(tq, 0, "position");
        glBindAttribLocation(tq, 1, "texcoord");
        glLinkProgram(tq);
        glGetProgramiv(tq, GL_LINK_STATUS, &ok);
        if (!ok) die("texquad link failed");
        glBindFramebuffer(GL_FRAMEBUFFER, 0);
        glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
        glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
        glUseProgram(tq);
        glViewport(0, 0, ww, wh);
        glUniformMatrix4fv(glGetUniformLocation(tq, "projection"), 1, 0, proj);
        glActiveTexture(GL_TEXTURE0);
        glUniform1i(glGetUniformLocation(tq, "texture_base"), 0);
        glBindTexture(GL_TEXTURE_2D, rt);
        GLfloat vq[] = {0.0f,    0.0f,    0.0f, 0.0f, 0.0f, 2048.0f, 0.0f,    0.0f, 1.0f, 0.0f,
                        0.0f,    2048.0f, 0.0f, 0.0f, 1.0f, 0.0f,    2048.0f, 0.0f, 0.0f, 1.0f,
                        2048.0f, 0.0f,    0.0f, 1.0f, 0.0f, 2048.0f, 2048.0f, 0.0f, 1.0f, 1.0f};
        GLuint qvbo, qvao;
        glGenBuffers(1, &qvbo);
        glBindBuffer(GL_ARRAY_BUFFER, qvbo);
        glGenVertexArrays(1, &qvao);
        glBindVertexArray(qvao);
        glEnableVertexAttribArray(0);
        glEnableVertexAttribArray(1);
        glVertexAttribPointer(0, 3, GL_FLOAT, 0, sizeof(GLfloat) * 5, 0);
        glVertexAttribPointer(1, 2, GL_FLOAT, 0, sizeof(GLfloat) * 5, (const void*) 12);
        glBufferData(GL_ARRAY_BUFFER, sizeof(vq), vq, GL_DYNAMIC_DRAW);
        glDrawArrays(GL_TRIANGLES, 0, 6);
        if (getenv("QB_NO_CROSSHAIR")) goto skip_crosshair;
        glEnable(GL_SCISSOR_TEST);
        glClearColor(1.0f, 1.0f, 1.0f, 1.0f);
        glScissor(ww / 2 - 1, wh / 2 - 1, 2, 2);
        glClear(GL_COLOR_BUFFER_BIT);
        glDisable(GL_SCISSOR_TEST);
    skip_crosshair:
        glFinish();
        W = ww, H = wh;
    }
    unsigned char* out = malloc((size_t) W * H * 4);
    glReadPixels(0, 0, W, H, GL_RGBA, GL_UNSIGNED_BYTE, out);
    f = fopen(argv[2], "wb");
    if (!f) die("cannot open output");
    fwrite(out, 4, (size_t) W * H, f);
    fclose(f);
    return 0;
}
