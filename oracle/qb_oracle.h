/*
 * qb_oracle.h -- CPU oracle for Qubatron's per-pixel octree hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (qubatron_b200/, the
 * C-ABI connector) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / reported baseline.
 *
 * It restates, statement by statement, the reference fragment shader
 *   /root/reference/src/qubatron/shaders/octree_fsh.c   (GLSL ES 3.00)
 * and the uniform set-up of
 *   /root/reference/src/qubatron/octree_glc.c L249-305  (octree_glc_update)
 * in strict IEEE fp32 (build with -O2 -ffp-contract=off, never -ffast-math,
 * -mfma or -march=native).
 *
 * Parity pinning (see DESIGN.md "Oracle"): the restatement is checked against
 *   (1) the reference's own compiled CPU twin octree_trace_line
 *       (octree.c L341-537, built unmodified into oracle/_ref/), and
 *   (2) the reference's unmodified GLSL run headless on Mesa llvmpipe
 *       (oracle/_ref/glsl_ref) with golden frames committed in tests/golden/.
 */
#ifndef QB_ORACLE_H
#define QB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* scene arrays exactly as the reference uploads them (octree_glc.c L361-498):
 * octrees = int32[12] per node (octree.c L11-14), colours/normals = float[3]
 * per point (model.c L14-25) */
typedef struct qb_scene
{
    const int32_t* oct_s;
    int64_t        nodes_s;
    const int32_t* oct_d;
    int64_t        nodes_d;
    const float*   col_s;
    const float*   nrm_s;
    int64_t        points_s;
    const float*   col_d;
    const float*   nrm_d;
    int64_t        points_d;
} qb_scene;

/* uniforms of octree_fsh.c L9-15 */
typedef struct qb_uniforms
{
    float   camfp[3];
    float   angle_in[3];
    float   light[3];
    float   basecube[4];
    float   dimensions[2];
    int32_t maxlevel;
    int32_t shoot;
    /* viewport in whole pixels (glViewport(0,0,ow,oh) truncates, octree_glc.c L288) */
    int32_t vp_w;
    int32_t vp_h;
} qb_uniforms;

/* per-pixel flags plane */
enum
{
    QB_FLAG_DISCARD   = 1,  /* `discard` hit in any of the traces: pixel keeps clear colour */
    QB_FLAG_LEAF      = 2,  /* primary trace returned a leaf (alpha == 1) */
    QB_FLAG_SHADED    = 4,  /* res.isp.w > 0: shadow ray traced and pixel shaded */
    QB_FLAG_LIT       = 8,  /* step(sqr,15.0) == 1 (valid when SHADED) */
    QB_FLAG_DISC_TEST = 16, /* camangle < 0.02: third trace executed */
    QB_FLAG_DISC_ON   = 32  /* light disc drawn (opaque white) */
};

/* aux plane: 6 int32 per pixel */
enum
{
    QB_AUX_MODEL_S = 0, /* oct[8] of the static leaf node (0 if none) */
    QB_AUX_MODEL_D = 1, /* oct[8] of the dynamic leaf node (0 if none) */
    QB_AUX_NODE_S  = 2, /* static leaf node index */
    QB_AUX_NODE_D  = 3, /* dynamic leaf node index */
    QB_AUX_SH_NODE_S = 4, /* shadow ray: static leaf node index (-1 no leaf) */
    QB_AUX_SH_NODE_D = 5, /* shadow ray: dynamic leaf node index */
    QB_AUX_STRIDE  = 6
};

/* deterministic work counters, SURVEY.md section 8(d) counting rule */
typedef struct qb_counters
{
    int64_t rays_primary;
    int64_t rays_shadow;
    int64_t rays_disc;
    int64_t expand_s; /* E_s */
    int64_t expand_d; /* E_d */
    int64_t leaf_s;   /* L_s */
    int64_t leaf_d;   /* L_d */
    int64_t hits;     /* H: pixels with isp.w > 0 */
    int64_t discards;
    int64_t descents;
} qb_counters;

/* octree_glc_update() uniform set-up (octree_glc.c L263-284) */
void qb_oracle_uniforms(qb_uniforms* u, float width, float height, const float position[3], const float angle[3],
                        float lighta, uint8_t quality, int maxlevel, float basesize, int shoot);

/* render rows [row0,row1) of the frame; rgba/flags/aux are FULL-frame planes
 * (row 0 = bottom, GL convention), any of flags/aux/counters may be NULL.
 * threads <= 0: all cores (OpenMP). */
void qb_oracle_render(const qb_scene* sc, const qb_uniforms* u, int row0, int row1, uint8_t* rgba, uint8_t* flags,
                      int32_t* aux, qb_counters* counters, int threads);

/* test hook: per-pixel `coord` planes ([vp_h][width] floats) used instead of the exact pixel centres, NULL to clear
 * (see octree_fsh_oracle.c; for comparisons against a rasteriser whose interpolation is not exact) */
void qb_oracle_set_coord_override(const float* cx, const float* cy, int width);

/* test hook: the view quaternions qz, qx (8 floats) instead of the libm ones, NULL to clear (octree_fsh_oracle.c) */
void qb_oracle_set_quat_override(const float* qz_qx8);

/* one cube_trace_line (octree_fsh.c L138-379).  returns 0 = miss (isp = 0),
 * 1 = leaf returned, -1 = discard.  out_isp[4], out_tlf[4], out_nodes[2],
 * out_models[2] may be NULL. */
int qb_oracle_trace(const qb_scene* sc, const qb_uniforms* u, const float pos[3], const float dir[3], float* out_isp,
                    float* out_tlf, int32_t* out_nodes, int32_t* out_models, qb_counters* counters);

/* batch of n rays: result[i] as above, models[i*2..] */
void qb_oracle_trace_batch(const qb_scene* sc, const qb_uniforms* u, int64_t n, const float* pos, const float* dir,
                           int32_t* result, int32_t* nodes, int32_t* models, float* isp, int threads);

/* primary-ray direction of pixel (px,py): octree_fsh.c L402-413 */
void qb_oracle_pixel_ray(const qb_uniforms* u, int px, int py, float dir[3]);

#ifdef __cplusplus
}
#endif
#endif
