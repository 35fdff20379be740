/*
 * octree_fsh_oracle.c -- CPU restatement of the reference pixel program.
 *
 * TEST INFRASTRUCTURE ONLY (see qb_oracle.h).  Follows
 *   /root/reference/src/qubatron/shaders/octree_fsh.c
 *     L62-99   is_cube_{x,y,z}plane          -> plane_hit()
 *     L127-136 oct_from_octets_for_index     -> node_slot()
 *     L138-379 cube_trace_line               -> trace_line()
 *     L381-395 quaternion helpers            -> quat_axis_angle(), quat_rotate()
 *     L399-464 main                          -> shade_pixel()
 *   /root/reference/src/qubatron/octree_glc.c
 *     L263-284 uniform set-up                -> qb_oracle_uniforms()
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (no -ffast-math / -mfma / -march=native).
 * Every float expression below keeps the operation order of the GLSL source;
 * mul and add are rounded separately.
 */
#include "qb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
    #include <omp.h>
#endif

/* Division.  Default: IEEE `/`, which is what the reference's own CPU twin (octree.c L302-339, compiled C)
 * does and what the CUDA path reproduces.  -DQB_DIV_MUL_RCP builds the variant that mirrors Mesa's GLSL
 * lowering (lower_instructions DIV_TO_MUL_RCP: a / b -> a * (1.0 / b), two roundings), used only to explain
 * the 0-2 pixels per frame on which llvmpipe and the IEEE oracle differ (tests/test_golden.py). */
#ifdef QB_DIV_MUL_RCP
    #define QB_DIV(a, b) ((a) * (1.0f / (b)))
#else
    #define QB_DIV(a, b) ((a) / (b))
#endif

typedef struct f4
{
    float x, y, z, w;
} f4;

typedef struct f3
{
    float x, y, z;
} f3;

/* octree_fsh.c L25-31 */
static const float k_xsft[8] = {0.f, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f, 1.f};
static const float k_ysft[8] = {0.f, 0.f, 1.f, 1.f, 0.f, 0.f, 1.f, 1.f};
static const float k_zsft[8] = {0.f, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f, 1.f};
static const int   k_hor[8]  = {1, 0, 3, 2, 5, 4, 7, 6};
static const int   k_ver[8]  = {2, 3, 0, 1, 6, 7, 4, 5};
static const int   k_dep[8]  = {4, 5, 6, 7, 0, 1, 2, 3};

#define QB_STACK_LEVELS 18 /* octree_fsh.c L151 */

/* octree_fsh.c L50-58 */
typedef struct level_t
{
    f4  cube;
    f4  isps[4];
    int octs[4];
    int ispsi;
    int socti;
    int docti;
} level_t;

typedef struct trace_res
{
    f4  isp;
    f4  col;
    f4  nrm;
    f4  tlf;
    int status;  /* 0 miss, 1 leaf, -1 discard */
    int node_s;  /* leaf node ids (valid when status == 1) */
    int node_d;
    int model_s; /* oct[8] values */
    int model_d;
} trace_res;

/* octree_fsh.c L62-99.  axis: 0 = x-plane, 1 = y-plane, 2 = z-plane */
static inline f4 plane_hit(int axis, float c, f3 lp, f3 lv)
{
    f4 r = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
    if (axis == 0)
    {
        if (lv.x != 0.0f)
        {
            r.w = QB_DIV(c - lp.x, lv.x);
            r.y = lp.y + lv.y * r.w;
            r.z = lp.z + lv.z * r.w;
            r.x = c;
        }
    }
    else if (axis == 1)
    {
        if (lv.y != 0.0f)
        {
            r.w = QB_DIV(c - lp.y, lv.y);
            r.x = lp.x + lv.x * r.w;
            r.z = lp.z + lv.z * r.w;
            r.y = c;
        }
    }
    else
    {
        if (lv.z != 0.0f)
        {
            r.w = QB_DIV(c - lp.z, lv.z);
            r.x = lp.x + lv.x * r.w;
            r.y = lp.y + lv.y * r.w;
            r.z = c;
        }
    }
    return r;
}

/* octree_fsh.c L127-136; texel addressing collapses to oct[i*12 + slot].
 * Out-of-range fetches return 0 (robust texelFetch). */
static inline int node_slot(int slot, int i, const int32_t* tree, int64_t nodes, int level)
{
    if (i == 0 && level > 0) return 0;
    if (i < 0 || (int64_t) i >= nodes) return 0;
    return tree[(int64_t) i * 12 + slot];
}

/* RGB32F texelFetch: .w reads 1.0; out-of-range reads (0,0,0,1) */
static inline f4 point_fetch(const float* arr, int64_t points, int idx)
{
    f4 r = {0.f, 0.f, 0.f, 1.f};
    if (arr != NULL && idx >= 0 && (int64_t) idx < points)
    {
        r.x = arr[(int64_t) idx * 3 + 0];
        r.y = arr[(int64_t) idx * 3 + 1];
        r.z = arr[(int64_t) idx * 3 + 2];
    }
    return r;
}

/* half-open face ranges, octree_fsh.c L166/L176/L186 */
static inline int in_xy(f4 a, f4 tlf, f4 brb) { return tlf.x < a.x && a.x <= brb.x && tlf.y > a.y && a.y >= brb.y; }
static inline int in_yz(f4 a, f4 tlf, f4 brb) { return tlf.y > a.y && a.y >= brb.y && tlf.z > a.z && a.z >= brb.z; }
static inline int in_xz(f4 a, f4 tlf, f4 brb) { return tlf.x < a.x && a.x <= brb.x && tlf.z > a.z && a.z >= brb.z; }

/* octree_fsh.c L138-379 */
static trace_res trace_line(const qb_scene* sc, const qb_uniforms* u, f3 pos, f3 dir, qb_counters* cnt)
{
    trace_res res;
    memset(&res, 0, sizeof(res));
    res.node_s = res.node_d = -1;

    int     level    = 0;
    int     maxlevel = u->maxlevel;
    level_t stck[QB_STACK_LEVELS];

    f4 basecube = {u->basecube[0], u->basecube[1], u->basecube[2], u->basecube[3]};

    stck[0].cube  = basecube;
    stck[0].socti = 0;
    stck[0].docti = 0;
    stck[0].ispsi = 0;

    f4 act;
    f4 tlf = basecube;
    f4 brb = {tlf.x + tlf.w, tlf.y - tlf.w, tlf.z - tlf.w, 0.0f};

    int hitc = 0;
    f4  hitp[6]; /* GLSL declares 4 and only [0],[1] are read; 6 keeps the writes in bounds */

    /* L164-192: front, back, left, right, top, bottom */
    act = plane_hit(2, tlf.z, pos, dir);
    if (in_xy(act, tlf, brb)) hitp[hitc++] = act;
    act = plane_hit(2, brb.z, pos, dir);
    if (in_xy(act, tlf, brb)) hitp[hitc++] = act;
    act = plane_hit(0, tlf.x, pos, dir);
    if (in_yz(act, tlf, brb)) hitp[hitc++] = act;
    act = plane_hit(0, brb.x, pos, dir);
    if (in_yz(act, tlf, brb)) hitp[hitc++] = act;
    act = plane_hit(1, tlf.y, pos, dir);
    if (in_xz(act, tlf, brb)) hitp[hitc++] = act;
    act = plane_hit(1, brb.y, pos, dir);
    if (in_xz(act, tlf, brb)) hitp[hitc++] = act;

    /* L195, L198 */
    if (hitc < 2)
    {
        res.status = -1;
        return res;
    }
    if (hitp[0].w < 0.0f && hitp[1].w < 0.0f)
    {
        res.status = -1;
        return res;
    }

    /* L205, L208 */
    if (hitp[1].w < hitp[0].w) hitp[0] = hitp[1];
    if (hitp[0].w < 0.0f)
    {
        hitp[0].x = pos.x;
        hitp[0].y = pos.y;
        hitp[0].z = pos.z;
        hitp[0].w = 0.0f;
    }

    stck[level].isps[0] = hitp[0];

    for (;;)
    {
        tlf = stck[level].cube;

        /* L218-248: leaf */
        if (level == maxlevel)
        {
            res.isp = stck[level].isps[0];
            res.tlf = tlf;

            int socti = node_slot(8, stck[level].socti, sc->oct_s, sc->nodes_s, level);
            int docti = node_slot(8, stck[level].docti, sc->oct_d, sc->nodes_d, level);

            res.col = point_fetch(sc->col_s, sc->points_s, socti);
            res.nrm = point_fetch(sc->nrm_s, sc->points_s, socti);
            if (docti > 0)
            {
                res.col = point_fetch(sc->col_d, sc->points_d, docti);
                res.nrm = point_fetch(sc->nrm_d, sc->points_d, docti);
            }

            res.status  = 1;
            res.node_s  = stck[level].socti;
            res.node_d  = stck[level].docti;
            res.model_s = socti;
            res.model_d = docti;
            if (cnt)
            {
                if (level == 0 || stck[level].socti != 0) cnt->leaf_s++;
                if (level == 0 || stck[level].docti != 0) cnt->leaf_d++;
            }
            return res;
        }

        /* L251-330: expand */
        if (stck[level].ispsi == 0)
        {
            stck[level].ispsi = 128;
            if (cnt)
            {
                if (level == 0 || stck[level].socti != 0) cnt->expand_s++;
                if (level == 0 || stck[level].docti != 0) cnt->expand_d++;
            }

            f4 b = {tlf.x + tlf.w, tlf.y - tlf.w, tlf.z - tlf.w, 0.0f};
            f4 hlf;
            hlf.x = b.x + (tlf.x - b.x) * 0.5f;
            hlf.y = b.y + (tlf.y - b.y) * 0.5f;
            hlf.z = b.z + (tlf.z - b.z) * 0.5f;
            hlf.w = b.w + (tlf.w - b.w) * 0.5f;

            f4 hp[4];
            int hc = 1;
            hp[0]  = stck[level].isps[0];

            act = plane_hit(2, hlf.z, pos, dir);
            if (act.w > 0.0f && in_xy(act, tlf, b)) hp[hc++] = act;
            act = plane_hit(0, hlf.x, pos, dir);
            if (act.w > 0.0f && in_yz(act, tlf, b)) hp[hc++] = act;
            act = plane_hit(1, hlf.y, pos, dir);
            if (act.w > 0.0f && in_xz(act, tlf, b)) hp[hc++] = act;

            int oct = 0;
            int pre = -1;

            for (int i = 0; i < hc; ++i)
            {
                if (i < hc - 1)
                {
                    for (int j = i + 1; j < hc; ++j)
                    {
                        if (hp[j].w < hp[i].w)
                        {
                            act   = hp[i];
                            hp[i] = hp[j];
                            hp[j] = act;
                        }
                    }
                }

                act = hp[i];
                oct = 0;
                if (act.x > hlf.x) oct = 1;
                if (act.y < hlf.y) oct += 2;
                if (act.z < hlf.z) oct += 4;
#ifdef QB_ORACLE_DEBUG
                fprintf(stderr, "L%d tlf %a %a %a %a hlf %a %a %a  cand %d/%d: %a %a %a w %a oct %d pre %d\n", level, tlf.x,
                        tlf.y, tlf.z, tlf.w, hlf.x, hlf.y, hlf.z, i, hc, act.x, act.y, act.z, act.w, oct, pre);
#endif

                if (oct == pre)
                {
                    if (act.x == hlf.x)
                        oct = k_hor[oct];
                    else if (act.y == hlf.y)
                        oct = k_ver[oct];
                    else if (act.z == hlf.z)
                        oct = k_dep[oct];
                }
                pre = oct;

                int socti = node_slot(oct, stck[level].socti, sc->oct_s, sc->nodes_s, level);
                int docti = node_slot(oct, stck[level].docti, sc->oct_d, sc->nodes_d, level);

                if (socti > 0 || docti > 0)
                {
                    int ind               = stck[level].ispsi;
                    int len               = ind & 0x0F;
                    stck[level].octs[len] = oct;
                    stck[level].isps[len] = act;
                    len++;
                    stck[level].ispsi = (ind & 0xF0) | len;
                }
            }
        }

        /* L334-375: descend or backtrack */
        int cur_len = stck[level].ispsi & 0x0F;
        if (cur_len > 0)
        {
            int nxt_ind = (stck[level].ispsi >> 4) & 7;
            f4  nxt_isp = stck[level].isps[nxt_ind];
            int nxt_oct = stck[level].octs[nxt_ind];

            float halfs = tlf.w / 2.0f;
            tlf.x += k_xsft[nxt_oct] * halfs;
            tlf.y -= k_ysft[nxt_oct] * halfs;
            tlf.z -= k_zsft[nxt_oct] * halfs;
            tlf.w = halfs;

            nxt_ind++;
            cur_len--;
            stck[level].ispsi = 128 | (nxt_ind << 4) | cur_len;

            int socti = node_slot(nxt_oct, stck[level].socti, sc->oct_s, sc->nodes_s, level);
            int docti = node_slot(nxt_oct, stck[level].docti, sc->oct_d, sc->nodes_d, level);

            level += 1;
            if (cnt) cnt->descents++;

            stck[level].cube    = tlf;
            stck[level].ispsi   = 0;
            stck[level].socti   = socti;
            stck[level].docti   = docti;
            stck[level].isps[0] = nxt_isp;
        }
        else
        {
            stck[level--].ispsi = 0;
            if (level < 0) return res; /* status 0, isp = 0, col = 0 */
        }
    }
}

/* GLSL cross(), dot(), normalize() on fp32 */
static inline f3 v_cross(f3 a, f3 b)
{
    f3 r = {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
    return r;
}
static inline float v_dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3    v_normalize(f3 a)
{
    float l = sqrtf(v_dot(a, a));
    f3    r = {QB_DIV(a.x, l), QB_DIV(a.y, l), QB_DIV(a.z, l)};
    return r;
}

/* octree_fsh.c L381-390 */
static inline f4 quat_axis_angle(f3 axis, float angle)
{
    f4    q;
    float half_angle = angle * 0.5f;
    q.x              = axis.x * sinf(half_angle);
    q.y              = axis.y * sinf(half_angle);
    q.z              = axis.z * sinf(half_angle);
    q.w              = cosf(half_angle);
    return q;
}

/* octree_fsh.c L392-395: v + 2.0 * cross(q.xyz, cross(q.xyz, v) + q.w * v) */
static inline f3 quat_rotate(f4 q, f3 v)
{
    f3 qv = {q.x, q.y, q.z};
    f3 c1 = v_cross(qv, v);
    f3 t  = {c1.x + q.w * v.x, c1.y + q.w * v.y, c1.z + q.w * v.z};
    f3 c2 = v_cross(qv, t);
    f3 r  = {v.x + 2.0f * c2.x, v.y + 2.0f * c2.y, v.z + 2.0f * c2.z};
    return r;
}

typedef struct frame_consts
{
    f4    qz, qx;
    f3    cfp;
    f3    camfp, light;
    f3    camlight_n; /* normalize(light - camfp) */
    float sx, sy;     /* coord scale when the viewport truncates ow/oh */
} frame_consts;

/* Optional view quaternions qz, qx (xyzw each).  They are the only place where the pixel program calls sin / cos
 * (quat_from_axis_angle, L406-408), whose precision GLSL ES leaves to the implementation: llvmpipe's polynomial and
 * libm's sinf / cosf agree to the last bit for most angles and differ by one ulp for some, which moves csv by an ulp
 * and flips a pixel once in ~10^6.  The oracle (like the connector) evaluates them with libm; when
 * it is compared against llvmpipe over random views it is given the driver's own two quaternions (glsl_ref mode 9). */
static const float* g_quat_override = NULL; /* qz[4], qx[4] */
void                qb_oracle_set_quat_override(const float* qz_qx8) { g_quat_override = qz_qx8; }

static void frame_setup(const qb_uniforms* u, frame_consts* fc)
{
    /* L403: tan(PI/4.0) folds to 1.0f in fp32 */
    const float tan_pi_4 = 1.0f;
    fc->cfp.x            = u->dimensions[0] / 2.0f;
    fc->cfp.y            = u->dimensions[1] / 2.0f;
    fc->cfp.z            = (u->dimensions[0] / 2.0f) / tan_pi_4;

    /* L406-408 */
    f3 yaxis = {0.0f, 1.0f, 0.0f};
    f3 negx  = {-1.0f, 0.0f, 0.0f};
    fc->qz   = quat_axis_angle(yaxis, -u->angle_in[0]);
    f3 vx    = quat_rotate(fc->qz, negx);
    fc->qx   = quat_axis_angle(vx, -u->angle_in[1]);
    if (g_quat_override)
    {
        const float* q = g_quat_override;
        fc->qz         = (f4){q[0], q[1], q[2], q[3]};
        fc->qx         = (f4){q[4], q[5], q[6], q[7]};
    }

    fc->camfp.x = u->camfp[0];
    fc->camfp.y = u->camfp[1];
    fc->camfp.z = u->camfp[2];
    fc->light.x = u->light[0];
    fc->light.y = u->light[1];
    fc->light.z = u->light[2];

    /* L417 */
    f3 camlight    = {fc->light.x - fc->camfp.x, fc->light.y - fc->camfp.y, fc->light.z - fc->camfp.z};
    fc->camlight_n = v_normalize(camlight);

    /* octree_vsh.c L11 + ortho quad: coord = pixel centre; exact (scale 1.0f)
     * whenever ow/oh are whole numbers (all BASELINE configs) */
    fc->sx = u->dimensions[0] / (float) u->vp_w;
    fc->sy = u->dimensions[1] / (float) u->vp_h;
}

/* Optional per-pixel `coord` planes ([vp_h][vp_w] floats each).  GL does not require the rasteriser's interpolation of
 * a varying to be exact, and llvmpipe's is not on every pixel (a few 1e-6 off on some rows / columns of some
 * viewports): the oracle evaluates coord = pixel centre exactly, and when it is compared against llvmpipe on
 * tie-prone views (scripts/oracle_fuzz_llvmpipe.py) it is given the coord llvmpipe actually produced (glsl_ref mode
 * 9), so that the comparison is about the traversal and shading, not about one driver's rasteriser. */
static const float* g_coord_x = NULL;
static const float* g_coord_y = NULL;
static int          g_coord_w = 0;
void qb_oracle_set_coord_override(const float* cx, const float* cy, int width)
{
    g_coord_x = cx;
    g_coord_y = cy;
    g_coord_w = width;
}

static inline f3 pixel_dir(const frame_consts* fc, int px, int py)
{
    /* L402-404, L412-413 */
    f3 ctp = {((float) px + 0.5f) * fc->sx, ((float) py + 0.5f) * fc->sy, 0.0f};
    if (g_coord_x)
    {
        ctp.x = g_coord_x[(size_t) py * g_coord_w + px];
        ctp.y = g_coord_y[(size_t) py * g_coord_w + px];
    }
    f3 csv = {ctp.x - fc->cfp.x, ctp.y - fc->cfp.y, ctp.z - fc->cfp.z};
    csv    = quat_rotate(fc->qz, csv);
    csv    = quat_rotate(fc->qx, csv);
    return csv;
}

static inline float f_max0(float a) { return a > 0.0f ? a : 0.0f; } /* max(a, 0.0); NaN -> 0 */

static inline uint8_t to_unorm8(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t) lrintf(v * 255.0f);
}

/* octree_fsh.c L399-464 */
static void shade_pixel(const qb_scene* sc, const qb_uniforms* u, const frame_consts* fc, int px, int py, uint8_t* rgba,
                        uint8_t* flags_out, int32_t* aux, qb_counters* cnt)
{
    int     flags = 0;
    int32_t a[QB_AUX_STRIDE] = {-1, -1, -1, -1, -1, -1};
    f4      col              = {0.f, 0.f, 0.f, 0.f};
    int     discard          = 0;

    f3 csv = pixel_dir(fc, px, py);

    /* L417-418 */
    f3    csv_n    = v_normalize(csv);
    float camdot   = v_dot(fc->camlight_n, csv_n);
    float camangle = acosf(camdot);

    if (cnt) cnt->rays_primary++;
    trace_res res = trace_line(sc, u, fc->camfp, csv, cnt);
    if (res.status < 0) discard = 1;

    if (!discard)
    {
        col = res.col;
        if (res.status == 1)
        {
            flags |= QB_FLAG_LEAF;
            a[QB_AUX_MODEL_S] = res.model_s;
            a[QB_AUX_MODEL_D] = res.model_d;
            a[QB_AUX_NODE_S]  = res.node_s;
            a[QB_AUX_NODE_D]  = res.node_d;
        }

        /* L424-450 */
        if (res.isp.w > 0.0f)
        {
            flags |= QB_FLAG_SHADED;
            if (cnt)
            {
                cnt->hits++;
                cnt->rays_shadow++;
            }
            f3 lghtv = {res.isp.x - fc->light.x, res.isp.y - fc->light.y, res.isp.z - fc->light.z};

            trace_res lcres = trace_line(sc, u, fc->light, lghtv, cnt);
            if (lcres.status < 0)
                discard = 1;
            else
            {
                if (lcres.status == 1)
                {
                    a[QB_AUX_SH_NODE_S] = lcres.node_s;
                    a[QB_AUX_SH_NODE_D] = lcres.node_d;
                }
                f3    ispv = {lcres.isp.x - res.isp.x, lcres.isp.y - res.isp.y, lcres.isp.z - res.isp.z};
                float sqr  = ispv.x * ispv.x + ispv.y * ispv.y + ispv.z * ispv.z;

                f3 nl = {-lghtv.x, -lghtv.y, -lghtv.z};
                f3 nc = {-csv.x, -csv.y, -csv.z};
                f3 nn = {res.nrm.x, res.nrm.y, res.nrm.z};
                nn    = v_normalize(nn);

                float lght_nrm_ang = f_max0(v_dot(v_normalize(nl), nn));
                float camv_nrm_ang = f_max0(v_dot(v_normalize(nc), nn));

                /* step(sqr, 15.0): 15.0 < sqr ? 0 : 1 */
                float vis = (15.0f < sqr) ? 0.0f : 1.0f;
                if (vis != 0.0f) flags |= QB_FLAG_LIT;

                float f = 0.1f + 0.2f * camv_nrm_ang + lght_nrm_ang * vis * 0.7f;
                col.x   = col.x * f;
                col.y   = col.y * f;
                col.z   = col.z * f;

                col.z *= 0.7f;

                float g = (float) u->shoot * camv_nrm_ang * 0.1f;
                col.x += g;
                col.y += g;
                col.z += g;
            }
        }
    }

    /* L452-459 */
    if (!discard && camangle < 0.02f)
    {
        flags |= QB_FLAG_DISC_TEST;
        if (cnt) cnt->rays_disc++;
        f3        lghtv = {fc->light.x - fc->camfp.x, fc->light.y - fc->camfp.y, fc->light.z - fc->camfp.z};
        trace_res lcres = trace_line(sc, u, fc->camfp, lghtv, cnt);
        if (lcres.status < 0)
            discard = 1;
        else
        {
            float resvx = lcres.isp.x - fc->camfp.x;
            if (QB_DIV(resvx, lghtv.x) > 1.0f)
            {
                flags |= QB_FLAG_DISC_ON;
                col.x = col.y = col.z = col.w = 1.0f;
            }
        }
    }

    if (discard)
    {
        flags = QB_FLAG_DISCARD;
        for (int i = 0; i < QB_AUX_STRIDE; i++) a[i] = -1;
        col.x = col.y = col.z = col.w = 0.0f;
        if (cnt) cnt->discards++;
    }

    /* L463 + RGBA8 unorm store of the clear-colour (0,0,0,0) target */
    rgba[0] = to_unorm8(col.x);
    rgba[1] = to_unorm8(col.y);
    rgba[2] = to_unorm8(col.z);
    rgba[3] = to_unorm8(col.w);
    if (flags_out) *flags_out = (uint8_t) flags;
    if (aux) memcpy(aux, a, sizeof(a));
}

/* octree_glc.c L263-284: note the double-precision intermediate arithmetic of
 * the C source (6.0, 20.0, 200.0 are double literals) */
void qb_oracle_uniforms(qb_uniforms* u, float width, float height, const float position[3], const float angle[3],
                        float lighta, uint8_t quality, int maxlevel, float basesize, int shoot)
{
    const float lightc[3] = {420.0f, 200.0f, 680.0f}; /* octree_glc.c L91 */
    memset(u, 0, sizeof(*u));
    u->basecube[0] = 0.0f;
    u->basecube[1] = basesize;
    u->basecube[2] = basesize;
    u->basecube[3] = basesize;
    u->light[0]    = lightc[0];
    u->light[1]    = (float) ((double) lightc[1] - (double) sinf(lighta) * 20.0);
    u->light[2]    = (float) ((double) lightc[2] - (double) sinf(lighta) * 200.0);
    u->angle_in[0] = angle[0];
    u->angle_in[1] = angle[1];
    u->angle_in[2] = 0.0f;
    u->camfp[0]    = position[0];
    u->camfp[1]    = position[1];
    u->camfp[2]    = position[2];

    float ow         = (float) ((double) width / (6.0 - (double) (float) quality / 2.0));
    float oh         = (float) ((double) height / (6.0 - (double) (float) quality / 2.0));
    u->dimensions[0] = ow;
    u->dimensions[1] = oh;
    u->vp_w          = (int32_t) ow;
    u->vp_h          = (int32_t) oh;
    u->maxlevel      = maxlevel;
    u->shoot         = shoot;
}

static void counters_add(qb_counters* dst, const qb_counters* src)
{
    dst->rays_primary += src->rays_primary;
    dst->rays_shadow += src->rays_shadow;
    dst->rays_disc += src->rays_disc;
    dst->expand_s += src->expand_s;
    dst->expand_d += src->expand_d;
    dst->leaf_s += src->leaf_s;
    dst->leaf_d += src->leaf_d;
    dst->hits += src->hits;
    dst->discards += src->discards;
    dst->descents += src->descents;
}

/* optional per-pixel work plane (descents of all rays of the pixel), for load-balance studies */
static int32_t* g_work_plane = NULL;
void            qb_oracle_set_work_plane(int32_t* plane) { g_work_plane = plane; }

void qb_oracle_render(const qb_scene* sc, const qb_uniforms* u, int row0, int row1, uint8_t* rgba, uint8_t* flags,
                      int32_t* aux, qb_counters* counters, int threads)
{
    frame_consts fc;
    frame_setup(u, &fc);
    int W = u->vp_w;
    if (row0 < 0) row0 = 0;
    if (row1 > u->vp_h) row1 = u->vp_h;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    qb_counters total;
    memset(&total, 0, sizeof(total));

#pragma omp parallel num_threads(threads)
    {
        qb_counters local;
        memset(&local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 1)
        for (int py = row0; py < row1; py++)
        {
            for (int px = 0; px < W; px++)
            {
                int64_t p  = (int64_t) py * W + px;
                int64_t d0 = local.descents;
                shade_pixel(sc, u, &fc, px, py, rgba + p * 4, flags ? flags + p : NULL,
                            aux ? aux + p * QB_AUX_STRIDE : NULL, (counters || g_work_plane) ? &local : NULL);
                if (g_work_plane) g_work_plane[p] = (int32_t) (local.descents - d0);
            }
        }
#pragma omp critical
        counters_add(&total, &local);
    }
    if (counters) counters_add(counters, &total);
}

int qb_oracle_trace(const qb_scene* sc, const qb_uniforms* u, const float pos[3], const float dir[3], float* out_isp,
                    float* out_tlf, int32_t* out_nodes, int32_t* out_models, qb_counters* counters)
{
    f3        p = {pos[0], pos[1], pos[2]};
    f3        d = {dir[0], dir[1], dir[2]};
    trace_res r = trace_line(sc, u, p, d, counters);
    if (out_isp)
    {
        out_isp[0] = r.isp.x;
        out_isp[1] = r.isp.y;
        out_isp[2] = r.isp.z;
        out_isp[3] = r.isp.w;
    }
    if (out_tlf)
    {
        out_tlf[0] = r.tlf.x;
        out_tlf[1] = r.tlf.y;
        out_tlf[2] = r.tlf.z;
        out_tlf[3] = r.tlf.w;
    }
    if (out_nodes)
    {
        out_nodes[0] = r.node_s;
        out_nodes[1] = r.node_d;
    }
    if (out_models)
    {
        out_models[0] = r.model_s;
        out_models[1] = r.model_d;
    }
    return r.status;
}

void qb_oracle_trace_batch(const qb_scene* sc, const qb_uniforms* u, int64_t n, const float* pos, const float* dir,
                           int32_t* result, int32_t* nodes, int32_t* models, float* isp, int threads)
{
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < n; i++)
    {
        result[i] = qb_oracle_trace(sc, u, pos + i * 3, dir + i * 3, isp ? isp + i * 4 : NULL, NULL,
                                    nodes ? nodes + i * 2 : NULL, models ? models + i * 2 : NULL, NULL);
    }
}

void qb_oracle_pixel_ray(const qb_uniforms* u, int px, int py, float dir[3])
{
    frame_consts fc;
    frame_setup(u, &fc);
    f3 d   = pixel_dir(&fc, px, py);
    dir[0] = d.x;
    dir[1] = d.y;
    dir[2] = d.z;
}
