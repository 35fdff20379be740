/*
 * octree_cuc.h -- C ABI of the B200 octree render connector ("cuc" = CUDA connector).
 *
 * Drop-in replacement for the reference's GL connector
 *   /root/reference/src/qubatron/octree_glc.c   (header part L1-87)
 * The three reference entry points keep their names, argument order, argument
 * meaning and (absent) error convention, so qubatron.c / modelutil.c compile
 * against this header unchanged:
 *
 *   octree_glc_init                   replaces octree_glc.c L65,  L93-247
 *   octree_glc_update                 replaces octree_glc.c L66-76, L249-352
 *   octree_glc_upload_texbuffer_data  replaces octree_glc.c L77-85, L361-498
 *   octree_glc_buffer_t               replaces octree_glc.c L16-24
 *   octree_glc_t                      replaces octree_glc.c L26-63 (only `memsize`
 *                                     is read outside the connector, qubatron.c L558)
 *
 * Everything named octree_cuc_* is an extension the headless / multi-GPU use
 * needs and the GL connector never had (the reference has no readback at all).
 *
 * Plain C, no CUDA or torch types: device buffers cross the boundary as
 * void* / uint64_t addresses.
 */
#ifndef OCTREE_CUC_H
#define OCTREE_CUC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mt_math/mt_vector_3d.c L6-10: passed BY VALUE to octree_glc_update */
#ifndef mt_vector_3d_h
typedef struct _v3_t v3_t;
struct _v3_t
{
    float x, y, z;
};
#endif

/* values of the GL enums the reference passes as `type` (modelutil.c L227-255) */
#define OCTREE_CUC_GL_INT 0x1404
#define OCTREE_CUC_GL_FLOAT 0x1406

/* octree_glc.c L16-24 */
typedef enum _octree_glc_buffer_t
{
    OCTREE_GLC_BUFFER_STATIC_COLOR,
    OCTREE_GLC_BUFFER_STATIC_NORMAL,
    OCTREE_GLC_BUFFER_STATIC_OCTREE,
    OCTREE_GLC_BUFFER_DYNAMIC_COLOR,
    OCTREE_GLC_BUFFER_DYNAMIC_NORMAL,
    OCTREE_GLC_BUFFER_DYNAMIC_OCTREE
} octree_glc_buffer_t;

/* octree_glc.c L26-63.  A value struct, returned by value from init and passed
 * by pointer afterwards.  `memsize` keeps the reference's name and meaning
 * (bytes of device storage, saturating at UINT32_MAX like the GLuint it was);
 * `memsize_bytes` is the unsaturated figure. */
typedef struct octree_glc_t
{
    void*        impl; /* connector state, owned by the connector */
    uint64_t     memsize_bytes;
    unsigned int memsize;
} octree_glc_t;

/* ---- reference API ------------------------------------------------------ */

/* octree_glc.c L93: `path` was the shader directory; accepted and ignored
 * (kernels are compiled into the library).  Uses the current CUDA device
 * (device 0 unless octree_cuc_select_device was called first).  Aborts with a
 * message when no CUDA device is usable -- there is no CPU fallback. */
octree_glc_t octree_glc_init(char* path);

/* octree_glc.c L249: render one frame.  Flushes pending range uploads, sets the
 * uniforms exactly as L263-284 (render size = width/(6 - quality/2), light from
 * lightc and lighta, base cube (0,S,S,S)), clears to (0,0,0,0) and runs the
 * trace+shadow+shade kernel.  Asynchronous like a GL draw: the frame is complete
 * after octree_cuc_sync / octree_cuc_read_frame.  The blit to a window and the
 * crosshair (L310-351) are presentation and are not done. */
void octree_glc_update(octree_glc_t* rc, float width, float height, v3_t position, v3_t angle, float lighta,
                       uint8_t quality, int maxlevel, float basesize, int shoot);

/* octree_glc.c L361: `data` points at the START of the whole logical array;
 * bytes [floor(start/itemsize)*itemsize, floor(end/itemsize)*itemsize) are
 * copied to the same offsets of the device array.  `size` only drives
 * capacity; when the array has to grow the call uploads [0,size) like the
 * reference does (L412-428).  type/itemsize: GL_FLOAT/12 for colour and
 * normal arrays, GL_INT/16 for octree arrays (a node = 3 items = 12 ints).
 * `data` has been consumed when the call returns. */
void octree_glc_upload_texbuffer_data(octree_glc_t* rc, void* data, int type, size_t size, size_t itemsize,
                                      size_t start, size_t end, octree_glc_buffer_t buftype);

/* ---- extensions ---------------------------------------------------------- */

/* Multi-GPU from the engine's one thread (SURVEY 8b "choose GPU count"; the reference is one C thread making one
 * call per frame, qubatron.c L116, L508-548).  Call right after octree_glc_init, before the first upload:
 * the connector then drives n devices of the box from the calling thread.  devices = n device ordinals
 * (devices[0] = the device octree_glc_init chose) or NULL for the next n - 1 devices after it.  From then on
 *   - every upload, tree build, skinning pass and setting reaches all n devices (bulk ones on all devices at once),
 *     so octree and point arrays are replicated;
 *   - octree_glc_update / octree_cuc_update_views render image tiles `mod n`: every device's kernel stores its
 *     tiles straight into the first device's framebuffer over NVLink (peer stores) and the last CTA of each
 *     kernel publishes the frame number into the first device's fence words; no collective, no host round trip;
 *   - read_frame / read_aux / read_window return the whole frame, counters are sums, last_frame_ms the slowest
 *     device's kernel, last_step_ms the time until every device's tiles have arrived.
 * The same ordinal may be given more than once (several shards on one GPU: how the single-GPU tests cover this). */
void octree_cuc_set_gpus(octree_glc_t* rc, int n, const int* devices);
int  octree_cuc_gpu_count(octree_glc_t* rc);

/* Errors.  The reference prints and carries on (octree_glc.c L243); the connector prints to stderr and abort()s on
 * any CUDA error or malformed argument -- it never returns partial state.  An engine that wants to log, save or
 * unwind first installs a handler: it is called with the message before the abort, and may choose not to return
 * (longjmp / exit).  Process-wide; NULL removes it. */
void octree_cuc_set_error_handler(void (*handler)(const char* message, void* user), void* user);

/* choose the CUDA device used by the next octree_glc_init (default: current) */
void octree_cuc_select_device(int device);

/* release all device memory of the connector */
void octree_cuc_destroy(octree_glc_t* rc);

/* wait for all queued uploads and frames */
void octree_cuc_sync(octree_glc_t* rc);

/* size of the last rendered frame */
void octree_cuc_frame_size(octree_glc_t* rc, int* width, int* height);

/* copy the last frame (RGBA8, row 0 = bottom like glReadPixels) to host memory;
 * synchronises.  Returns bytes written. */
size_t octree_cuc_read_frame(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity);

/* pipelined readback: queue the copy of the last frame to (page-locked) host memory on a separate copy stream and
 * return at once; the next octree_glc_update renders into a second framebuffer, so the copy of frame i overlaps the
 * rendering of frame i+1.  octree_cuc_wait_reads waits for all queued copies.  Returns the bytes that will be
 * written (0 if the buffer is too small). */
size_t octree_cuc_read_frame_async(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity);
void   octree_cuc_wait_reads(octree_glc_t* rc);
/* The same overlap for a frame whose render target must stay where it is -- rank 0's framebuffer that the other
 * ranks store their tiles into over NVLink (octree_cuc_ipc_export_frame): the finished frame is first copied device to
 * device (8 MB: a few microseconds, on the render stream, i.e. after whatever fence the caller queued there) into one
 * of two staging buffers, and goes to the page-locked host buffer from there on the copy stream while the next frame
 * renders.  octree_cuc_wait_reads waits for these copies too. */
size_t octree_cuc_read_frame_staged(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity);

/* device address of the RGBA8 frame (valid until the next resize) */
uint64_t octree_cuc_frame_device(octree_glc_t* rc);

/* render straight into caller-owned device memory (e.g. a torch tensor or a
 * peer-mapped buffer of another GPU); 0 restores the internal framebuffer.
 * pitch_pixels = row stride in pixels (0 = frame width). */
void octree_cuc_set_frame_target(octree_glc_t* rc, uint64_t device_ptr, size_t pitch_pixels);

/* parity planes: when enabled every frame also writes flags (1 byte/pixel,
 * OCTREE_CUC_FLAG_*) and aux (6 int32/pixel, OCTREE_CUC_AUX_*) */
void   octree_cuc_enable_aux(octree_glc_t* rc, int enable);
size_t octree_cuc_read_aux(octree_glc_t* rc, uint8_t* flags_host, int32_t* aux_host);

enum
{
    OCTREE_CUC_FLAG_DISCARD   = 1,
    OCTREE_CUC_FLAG_LEAF      = 2,
    OCTREE_CUC_FLAG_SHADED    = 4,
    OCTREE_CUC_FLAG_LIT       = 8,
    OCTREE_CUC_FLAG_DISC_TEST = 16,
    OCTREE_CUC_FLAG_DISC_ON   = 32
};
enum
{
    OCTREE_CUC_AUX_MODEL_S   = 0,
    OCTREE_CUC_AUX_MODEL_D   = 1,
    OCTREE_CUC_AUX_NODE_S    = 2,
    OCTREE_CUC_AUX_NODE_D    = 3,
    OCTREE_CUC_AUX_SH_NODE_S = 4,
    OCTREE_CUC_AUX_SH_NODE_D = 5,
    OCTREE_CUC_AUX_STRIDE    = 6
};

/* work counters of the last frame (SURVEY.md 8d counting rule); counting is a
 * separate kernel instantiation, enabled per frame */
typedef struct octree_cuc_counters
{
    int64_t rays_primary;
    int64_t rays_shadow;
    int64_t rays_disc;
    int64_t expand_s;
    int64_t expand_d;
    int64_t leaf_s;
    int64_t leaf_d;
    int64_t hits;
    int64_t discards;
    int64_t descents;
} octree_cuc_counters;
void octree_cuc_enable_counters(octree_glc_t* rc, int enable);
void octree_cuc_read_counters(octree_glc_t* rc, octree_cuc_counters* out);

/* image-tile sharding (multi-GPU): this connector renders only tiles whose
 * index (row-major over tile_w x tile_h tiles) is congruent to `rank` modulo
 * `world`.  world = 1 renders everything (default).  Pixels of other ranks'
 * tiles are left untouched. */
void octree_cuc_set_shard(octree_glc_t* rc, int rank, int world, int tile_w, int tile_h);

/* light override: when set, octree_glc_update uses it instead of lightc/lighta */
void octree_cuc_set_light(octree_glc_t* rc, const float* light_xyz_or_null);

/* kernel selection: 0 = auto (fast kernel when the base cube is exactly
 * representable at every level, generic otherwise), 1 = generic, 2 = fast */
void octree_cuc_set_kernel(octree_glc_t* rc, int which);
int  octree_cuc_last_kernel(octree_glc_t* rc);

/* division semantics of the pixel program.  The shader's `/` has two reproducible
 * executions and the connector can match either bit for bit:
 *   OCTREE_CUC_DIV_GLSL (default): a / b evaluated as a * (1.0 / b) -- what Mesa's
 *     GLSL front end makes of it (DIV_TO_MUL_RCP) when the reference shader runs
 *     headless on llvmpipe;
 *   OCTREE_CUC_DIV_IEEE: one correctly rounded divide -- what the reference's
 *     compiled CPU twin octree_trace_line (octree.c L302-339) computes. */
#define OCTREE_CUC_DIV_GLSL 0
#define OCTREE_CUC_DIV_IEEE 1
void octree_cuc_set_division(octree_glc_t* rc, int mode);

/* device time of the last frame's kernels in milliseconds (CUDA events on the
 * connector's stream); synchronises on the frame */
float octree_cuc_last_frame_ms(octree_glc_t* rc);

/* device time from the start of the last frame until it was complete: on the connector that collects a frame split
 * over several (rank 0 of a fence, the primary of a group) this includes the wait for the other devices' tiles;
 * elsewhere it equals octree_cuc_last_frame_ms */
float octree_cuc_last_step_ms(octree_glc_t* rc);

/* number of kernels this connector has launched since init */
uint64_t octree_cuc_launch_count(octree_glc_t* rc);

/* a batch of views of the same scene (multi-view config): n frames of equal
 * size rendered back to back into one device buffer of n*W*H pixels.
 * positions/angles: 3 floats per view. */
void octree_cuc_update_views(octree_glc_t* rc, int n, float width, float height, const float* positions,
                             const float* angles, float lighta, uint8_t quality, int maxlevel, float basesize,
                             int shoot);

/* run all of the connector's work (uploads, kernels, copies) on a caller-owned
 * CUDA stream, e.g. torch's current stream, so that collectives and event
 * timing issued by the caller order naturally with the frames; 0 restores the
 * connector's own stream. */
void octree_cuc_set_stream(octree_glc_t* rc, uint64_t cuda_stream);

/* allocate the internal framebuffer for `views` frames of width x height now
 * (normally done lazily by the first update) */
void octree_cuc_reserve_frame(octree_glc_t* rc, int width, int height, int views);

/* fused tile gather over NVLink: rank 0 exports its framebuffer as a 64-byte
 * CUDA IPC handle, the other ranks open it and render their tiles straight
 * into it with peer stores (octree_cuc_set_frame_target). */
void     octree_cuc_ipc_export_frame(octree_glc_t* rc, uint8_t* handle64);
void     octree_cuc_ipc_export_ptr(octree_glc_t* rc, uint64_t device_ptr, uint8_t* handle64);
uint64_t octree_cuc_ipc_open(octree_glc_t* rc, const uint8_t* handle64);
void     octree_cuc_ipc_close(octree_glc_t* rc, uint64_t device_ptr);

/* Completion fence for one frame split over several connectors -- one per process under torchrun, where
 * octree_cuc_set_gpus cannot be used -- by device-side flags instead of a collective per frame.  Every connector owns
 * a few fence words (octree_cuc_fence_device; export with octree_cuc_ipc_export_ptr, open with octree_cuc_ipc_open);
 * octree_cuc_set_fence(rank, world, ptrs) takes the addresses of all `world` connectors' words as this device sees
 * them (ptrs[rank] = its own).  Ranks > 0 render into rank 0's framebuffer (octree_cuc_set_frame_target); the last
 * CTA of their kernel publishes the frame number to rank 0 (__threadfence_system + release store over NVLink), rank
 * 0's stream waits for all of them in a one-warp kernel after its own tiles, and the first CTA of rank 0's NEXT
 * frame tells the others that the previous frame has been consumed -- their pixel stores (not their traversal)
 * wait for that.  Every rank must call octree_glc_update once per frame, rank 0 first when ranks share a device.
 * A peer that never answers traps the waiting kernel after ~4 s instead of hanging the GPU.  world <= 1 or
 * ptrs == NULL removes the fence. */
uint64_t octree_cuc_fence_device(octree_glc_t* rc);
void     octree_cuc_set_fence(octree_glc_t* rc, int rank, int world, const uint64_t* fence_ptrs);

/* "Next" row: build an octree on the GPU from per-point octant paths, with the reference's
 * node numbering, straight into the traversal layout -- what the engine does on the CPU every frame with
 * octree_reset + octree_insert_path (qubatron.c L439-452, octree.c L149-180).  oct14 / oct54 / oct94 are the
 * three int32[4 * n] digit buffers of the skinning pass (skeleton_glc.c L238-245), on the host or
 * (paths_on_device != 0) already on the device; model index of point i = first_modind + i.  Replaces the
 * whole tree named by `buftype`; returns the node count. */
size_t octree_cuc_build_octree_from_paths(octree_glc_t* rc, const int32_t* oct14, const int32_t* oct54,
                                          const int32_t* oct94, size_t n, int first_modind, int paths_on_device,
                                          octree_glc_buffer_t buftype);

/* "Next" row (first half): the skinning / octant-path pass of the reference's skeleton connector
 * (shaders/skeleton_vsh.c, dispatched by skeleton_glc.c L222-251) on the GPU, outputs kept on the device.
 *   octree_cuc_skeleton_alloc_in   replaces skeleton_glc_alloc_in (skeleton_glc.c L262): rest-pose points and
 *                                  normals, float[3] each, `bytes` = size of ONE of the two arrays
 *   octree_cuc_skeleton_update     replaces the transform-feedback draw: oldbones / newbones = 20 x vec4
 *                                  (zombie.oribones / newbones); skins points [0, model_count), writes their
 *                                  normals into the dynamic model and, with build_tree != 0, rebuilds the dynamic
 *                                  octree from their digits (what qubatron.c L439-452 + L508-529 do through the
 *                                  host).  Returns the node count of the rebuilt tree (0 if not built).
 *   octree_cuc_skeleton_read_out   the buffers the reference reads back every frame (skeleton_glc.c L238-245):
 *                                  oct14 / oct54 / oct94 (int32[4 * n]), normals and skinned points (float[3 * n]),
 *                                  any pointer may be NULL; returns n */
/*   octree_cuc_skeleton_set_rotations   sin / cos / acos have implementation-defined precision in GLSL ES, so the
 *                                  ten bone pairs' two rotation quaternions (skeleton_vsh.c L137, L151) are the one
 *                                  part of the program whose bits depend on the GL driver.  By default the connector
 *                                  evaluates them once per update on the host with libm (the shader's own TODO,
 *                                  L72: "rotation quaternions should be precalculated on the CPU per bone").  A host
 *                                  that has its own values passes them here: float[10][9] = rot_quat xyzw,
 *                                  axis_quat xyzw, has_axis (0 / 1) per pair; NULL returns to libm.  With a GL
 *                                  driver's values the outputs equal that driver's bit for bit (tests/golden). */
void   octree_cuc_skeleton_set_rotations(octree_glc_t* rc, const float* rotations90);
void   octree_cuc_skeleton_alloc_in(octree_glc_t* rc, const float* pntdata, const float* nrmdata, size_t bytes);
size_t octree_cuc_skeleton_update(octree_glc_t* rc, const float* oldbones80, const float* newbones80, int model_count,
                                  int maxlevel, float basesize, int build_tree);
size_t octree_cuc_skeleton_read_out(octree_glc_t* rc, int32_t* oct14, int32_t* oct54, int32_t* oct94, float* nrm_out,
                                    float* pnt_out);

/* Tile scheduling by measured cost (fast kernel, single-view frames; on by default).  The kernel records how many
 * warp-cycles every 64 x 64 tile took; the next frame of the same view -- or, for a view not seen before, of the most
 * recent one -- launches its tiles heaviest first, so that the longest rays of a frame do not end up running alone
 * at the end of the launch.  Pixels are independent: the frame is bit-identical either way. */
void octree_cuc_set_tile_feedback(octree_glc_t* rc, int on);

/* L2 residency A/B switch (BASELINE north_star design point 1, "hot upper levels ... in L2-persisting windows"): sets
 * aside `persist_bytes` of L2 (clamped to the device maximum) and marks the head of the static tree's node array -- as
 * much of it as the device's access-policy window allows -- persisting on the stream the frames run on, with the hit
 * ratio that fits the set-aside.  0 removes the window.  Frames are identical either way; what it buys is measured
 * in DESIGN.md (nothing on one GPU, where the node loads hit L1 94-98 % of the time). */
void octree_cuc_set_persisting_window(octree_glc_t* rc, size_t persist_bytes);

/* Resident CTAs per SM of the fast kernel (1..6; 0 = as many as fit, the default 7).  A frame split over many GPUs
 * leaves each of them a shard whose time is set by its longest rays, not by throughput: with fewer warps per SM every
 * warp runs faster.  Tuning knob, frames are identical. */
void octree_cuc_set_occupancy(octree_glc_t* rc, int ctas_per_sm);

/* "Next" row (second half): the presentation pass of octree_glc_update (octree_glc.c L308-351).  With
 * present enabled every single-view, unsharded octree_glc_update also produces what the reference leaves in the
 * window's back buffer: the frame drawn LINEAR-filtered into (int)width x (int)height pixels (all four channels),
 * plus the 2 x 2 white crosshair.  The window image stays on the device (octree_cuc_window_device: RGBA8, row 0 =
 * bottom -- what an interactive build registers with its window texture) or is copied out (octree_cuc_read_window,
 * which also reports the size; returns bytes copied or 0).  Off by default: the bench measures the frame.
 * Filter precision is implementation-defined in GL; the connector reproduces Mesa llvmpipe's RGBA8 filter bit for
 * bit (8-bit weights, like NVIDIA hardware); at quality 10 the pass is an exact copy on any implementation. */
void     octree_cuc_enable_present(octree_glc_t* rc, int on);
size_t   octree_cuc_read_window(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity, int* width, int* height);
uint64_t octree_cuc_window_device(octree_glc_t* rc);

/* "Next" row: the particle and dust simulation steps of the reference's transform-feedback programs
 * (shaders/particle_vsh.c dispatched by particle_glc.c L118-156 over the octree texture that octree_glc.c bound,
 * L105-113; shaders/dust_vsh.c dispatched by dust_glc.c L103-135).  `kind` selects the program.
 *   octree_cuc_particles_alloc_in   replaces particle_glc_alloc_in / dust_glc_alloc_in: positions and speeds,
 *                                   float[3] each, `bytes` = size of ONE of the two arrays
 *   octree_cuc_particles_update     replaces particle_glc_update / dust_glc_update: `steps` simulation steps over the
 *                                   first `count` particles, state kept on the device (the reference round-trips it
 *                                   through the host every frame, modelutil.c L715-724); campos is used by dust only
 *   octree_cuc_particles_read_out   the programs' two output buffers (either may be NULL); for OCTREE_CUC_PARTICLES
 *                                   returns how many of the `count` particles the last step left parked
 *                                   (speed.x < -900, the host's end-of-simulation test, modelutil.c L730-744) */
enum
{
    OCTREE_CUC_PARTICLES = 0, /* particle_vsh.c */
    OCTREE_CUC_DUST      = 1  /* dust_vsh.c */
};
void   octree_cuc_particles_alloc_in(octree_glc_t* rc, int kind, const float* posdata, const float* spddata, size_t bytes);
void   octree_cuc_particles_update(octree_glc_t* rc, int kind, int count, int maxlevel, float basesize, v3_t campos,
                                   int steps);
size_t octree_cuc_particles_read_out(octree_glc_t* rc, int kind, int count, float* pos_out, float* spd_out);

/* "Next" row: the offline voxeliser qmc (qmc.c: grid index at 2 * 2^levels cells per axis,
 * drop points outside the cube, x-major sort, first point of every occupied cell, colour = uchar / 255.0) and
 * the bulk tree build (octree_insert_point order) on the GPU, straight into the renderer's arrays.  pos / nrm:
 * float[3 * n], col_u8: uchar[3 * n] (on the host, or on the device when inputs_on_device != 0).  Fills the
 * colour/normal arrays and the octree of the static (dynamic = 0) or dynamic model; returns the number of
 * surviving points.  order_host (int64[n], optional) receives their source indices, pos_host (float[3*n],
 * optional) their positions -- what qmc writes to the .pnt file. */
size_t octree_cuc_voxelise_and_build(octree_glc_t* rc, const float* pos, const uint8_t* col_u8, const float* nrm,
                                     size_t n, int size, int levels, int inputs_on_device, int dynamic,
                                     int64_t* order_host, float* pos_host);

/* "Next" row: a batch of the engine's CPU ray queries octree_trace_line (octree.c L341-537)
 * against the device copy of ONE tree (dynamic_tree = 0 static, 1 dynamic).  pos / dir: float[3 * n] on the
 * host.  out_index[i] = oct[8] of the leaf hit or 0; out_tlf (float[4 * n], optional) = the leaf cube, left
 * untouched for misses like *otlf.  Same arithmetic as the compiled reference (IEEE division, its parallel-ray
 * sentinel): results are bit-identical to it. */
void octree_cuc_trace_lines(octree_glc_t* rc, size_t n, const float* pos, const float* dir, int dynamic_tree,
                            int maxlevel, float basesize, int32_t* out_index, float* out_tlf);

/* copy the device colour / normal arrays back as float[3] per point (parity checks); returns the point count */
size_t octree_cuc_download_points(octree_glc_t* rc, int dynamic, float* col_host, float* nrm_host,
                                  size_t capacity_points);

/* copy a device octree back in the reference's 12-int node format (parity checks); returns the node count,
 * writes nothing if capacity_nodes is too small */
size_t octree_cuc_download_octree(octree_glc_t* rc, octree_glc_buffer_t buftype, int32_t* nodes12_host,
                                  size_t capacity_nodes);

/* Optional: page-lock a host array the caller uploads from every frame (the dynamic octree, the skinning
 * output) so that bulk uploads run at full PCIe rate instead of through the driver's pageable staging.
 * The caller guarantees the memory stays allocated until it is unpinned; nothing is pinned implicitly. */
void octree_cuc_pin_host_buffer(octree_glc_t* rc, void* data, size_t bytes);
void octree_cuc_unpin_host_buffer(octree_glc_t* rc, void* data);

/* Bulk ranges (> 1 MiB) from PAGEABLE host memory -- what the unmodified engine passes -- are copied into page-locked
 * staging by `threads` host threads (+ the caller) while the DMA engine moves the previous piece, instead of through
 * the driver's pageable path (~10 GB/s).  Default 4; 0 = the driver's path.  Single-device connectors only: the
 * members of an octree_cuc_set_gpus group keep the driver's path.  Page-locked arrays (octree_cuc_pin_host_buffer)
 * go to the device directly either way. */
void octree_cuc_set_upload_threads(octree_glc_t* rc, int threads);

/* wall-clock milliseconds the connector spent inside upload calls (host side, including the copies it waited
 * for) since the last call of this function */
double octree_cuc_take_upload_ms(octree_glc_t* rc);

/* multi-GPU range updates across processes: with the replication log enabled, EVERY range the connector receives
 * through octree_glc_upload_texbuffer_data (batched or bulk, growth re-uploads included) is also recorded;
 * octree_cuc_export_pending drains the log as one packed blob (header + descriptors + payload; call with NULL for
 * the size), the caller broadcasts it (NCCL) and every other rank applies it.  apply_blob validates every
 * descriptor before it touches anything. */
void   octree_cuc_enable_replication_log(octree_glc_t* rc, int on);
size_t octree_cuc_export_pending(octree_glc_t* rc, void* blob_host, size_t capacity);
/* the same blob written straight into DEVICE memory of this GPU (the send buffer of the broadcast; 8-byte aligned):
 * the log keeps its payload in page-locked memory, so it moves at PCIe rate; 0 / too small -> the size needed */
size_t octree_cuc_export_pending_device(octree_glc_t* rc, uint64_t blob_device, size_t capacity);
void   octree_cuc_apply_blob(octree_glc_t* rc, const void* blob_host, size_t bytes);
/* the same for a blob that already sits in this GPU's memory (the receive buffer of the broadcast, 8-byte aligned):
 * the payload never travels through the host, one scatter launch applies it; the caller keeps the buffer alive until
 * the connector's stream has passed the call (octree_cuc_sync, or the next frame's completion) */
void   octree_cuc_apply_blob_device(octree_glc_t* rc, uint64_t blob_device, size_t bytes);

/* device self-test: the fast kernel's hoisted-reciprocal division against the IEEE
 * `/` on `count` operand pairs drawn like the traversal's (see octree_trace_fast.cuh);
 * returns the number of results that differ (must be 0) */
uint64_t octree_cuc_selftest_div(octree_glc_t* rc, uint64_t seed, uint64_t count);

/* host-side copy of the fast kernel's candidate-ordering table (octree_trace_fast.cuh, g_order_lut): 4096
 * entries, index = the 12 compare bits of the common expansion case, value = ordered candidate list + one-hot
 * octants.  No device needed; tests rebuild it from the reference's formulation (octree_fsh.c L276-311). */
void octree_cuc_debug_order_lut(uint64_t* out4096);

/* library self-description */
const char* octree_cuc_version(void);

#ifdef __cplusplus
}
#endif
#endif
