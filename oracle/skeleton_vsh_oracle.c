/*
 * skeleton_vsh_oracle.c -- CPU restatement of the reference's skinning / octant-path vertex program.
 *
 * TEST INFRASTRUCTURE ONLY (see qb_oracle.h).  Follows, statement by statement,
 *   /root/reference/src/qubatron/shaders/skeleton_vsh.c
 *     L35-40   project_point
 *     L49-58   quat_from_axis_angle
 *     L63-70   qrot
 *     L74-186  main: per bone pair, weight by distance, rotate, blend
 *     L188-226 main: 12 octant digits of the skinned point
 * with the uniforms of skeleton_glc.c L222-227 (oldbones / newbones: 20 x vec4, basecube, maxlevel).
 *
 * "next" row SURVEY 8f #1.  Parity status: PINNED against the shader itself.  oracle/glsl_ref.c (modes 20-22) runs
 * skeleton_vsh.c unmodified through transform feedback on Mesa llvmpipe; tests/golden/skin_*.npz hold its outputs
 * and tests/test_golden.py checks this restatement against them bit for bit (digits, normals, skinned positions).
 * sin / cos / acos have implementation-defined precision in GLSL ES and are used only for the two rotation
 * quaternions of each bone pair, which depend on the bones alone (the shader's own TODO, L72: "rotation
 * quaternions should be precalculated on the CPU per bone"): the fixtures record the driver's 10 x 2 quaternions,
 * qb_oracle_skin_rot takes them as an input, and without them they are evaluated with libm.  The per-point
 * arithmetic is + - * / sqrt only.
 *
 * Built into both oracle libraries (IEEE `/` and, with -DQB_DIV_MUL_RCP, Mesa's a * (1/b) lowering).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef QB_DIV_MUL_RCP
    #define QB_DIV(a, b) ((a) * (1.0f / (b)))
#else
    #define QB_DIV(a, b) ((a) / (b))
#endif

typedef struct s3
{
    float x, y, z;
} s3;
typedef struct s4
{
    float x, y, z, w;
} s4;

static inline s3    sub3(s3 a, s3 b) { return (s3){a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline s3    add3(s3 a, s3 b) { return (s3){a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline s3    mul3(s3 a, float f) { return (s3){a.x * f, a.y * f, a.z * f}; }
static inline s3    div3(s3 a, float f) { return (s3){QB_DIV(a.x, f), QB_DIV(a.y, f), QB_DIV(a.z, f)}; }
static inline float dot3(s3 a, s3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len3(s3 a) { return sqrtf(dot3(a, a)); }
static inline s3    cross3(s3 a, s3 b) { return (s3){a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
static inline s3    norm3(s3 a)
{
    float l = len3(a);
    return div3(a, l);
}
/* L63-70 */
static inline s3 qrot(s4 q, s3 v)
{
    s3 qv = {q.x, q.y, q.z};
    s3 c1 = cross3(qv, v);
    s3 t  = {c1.x + q.w * v.x, c1.y + q.w * v.y, c1.z + q.w * v.z};
    s3 c2 = cross3(qv, t);
    return (s3){v.x + 2.0f * c2.x, v.y + 2.0f * c2.y, v.z + 2.0f * c2.z};
}
/* L49-58 */
static inline s4 quat_axis_angle(s3 axis, float angle)
{
    float h = angle * 0.5f;
    return (s4){axis.x * sinf(h), axis.y * sinf(h), axis.z * sinf(h), cosf(h)};
}

/* everything main() derives from one bone pair that does not depend on the point */
typedef struct qb_bone_consts
{
    s3    a, b;          /* oldbones[i].xyz, oldbones[i+1].xyz */
    float effect;        /* oldbones[i].w */
    s3    oldbone;       /* b - a */
    s3    midp;          /* a + oldbone / 2 */
    float half_len;      /* length(oldbone) / 2 */
    float ab_dot;        /* dot(AB, AB) of project_point */
    s3    newa;          /* newbones[i].xyz */
    s4    rot_quat;      /* quat_from_axis_angle(normalize(oldbone), newbones[i].w) */
    int   has_axis;      /* length(cross(oldbone_norm, currbone_norm)) > 0.000001 */
    s4    axis_quat;     /* quat_from_axis_angle(normalize(bones_axis), acos(bones_dot)) */
} qb_bone_consts;

void qb_oracle_bone_consts(const float* oldbones80, const float* newbones80, qb_bone_consts* out10)
{
    for (int k = 0; k < 10; k++)
    {
        int             i = 2 * k;
        qb_bone_consts* c = out10 + k;
        memset(c, 0, sizeof(*c));
        c->a      = (s3){oldbones80[i * 4], oldbones80[i * 4 + 1], oldbones80[i * 4 + 2]};
        c->b      = (s3){oldbones80[i * 4 + 4], oldbones80[i * 4 + 5], oldbones80[i * 4 + 6]};
        c->effect = oldbones80[i * 4 + 3];
        c->oldbone  = sub3(c->b, c->a);                                   /* L92 */
        c->midp     = add3(c->a, div3(c->oldbone, 2.0f));                 /* L93 */
        c->half_len = QB_DIV(len3(c->oldbone), 2.0f);                     /* L103 */
        c->ab_dot   = dot3(c->oldbone, c->oldbone);                       /* L39 */
        s3 na       = {newbones80[i * 4], newbones80[i * 4 + 1], newbones80[i * 4 + 2]};
        s3 nb       = {newbones80[i * 4 + 4], newbones80[i * 4 + 5], newbones80[i * 4 + 6]};
        c->newa     = na;
        s3 currbone = sub3(nb, na);                                       /* L119 */
        s3 on       = norm3(c->oldbone);                                  /* L132 */
        s3 cn       = norm3(currbone);                                    /* L133 */
        c->rot_quat = quat_axis_angle(on, newbones80[i * 4 + 3]);         /* L137 */
        float bones_dot   = dot3(on, cn);                                 /* L145 */
        float bones_angle = acosf(bones_dot);                             /* L146 */
        s3    bones_axis  = cross3(on, cn);                               /* L147 */
        c->has_axis       = len3(bones_axis) > 0.000001f;                 /* L149 */
        if (c->has_axis) c->axis_quat = quat_axis_angle(norm3(bones_axis), bones_angle); /* L151 */
    }
}

/* one point: skeleton_vsh.c main().  out_digits[12], out_pnt (the skinned position, not an output of the shader,
 * exported for tests), out_nrm = normal_out */
static void skin_point(const qb_bone_consts* bc, s3 position, s3 normal, const float basecube[4], int maxlevel,
                       int32_t* out_digits, float* out_nrm, float* out_pnt)
{
    s3    corner_points[20];
    s3    corner_normals[20];
    float corner_tozerow[20];
    int   corner_count  = 0;
    s3    corner_center = position;
    float tozerow_sum   = 0.0f;
    /* A point out of range of every bone pair takes corner_normals[0] (L172), which the shader never wrote: undefined
     * in GLSL (garbage on llvmpipe, scripts/oracle_fuzz_tf_llvmpipe.py).  Defined here, and in the connector, as 0. */
    corner_normals[0] = (s3){0.0f, 0.0f, 0.0f};

    for (int k = 0; k < 10; k++)
    {
        const qb_bone_consts* c = bc + k;
        /* L94: project_point(A, B, C) = A + dot(AC, AB) / dot(AB, AB) * AB */
        s3    AC               = sub3(position, c->a);
        float t                = QB_DIV(dot3(AC, c->oldbone), c->ab_dot);
        s3    point_on_oldbone = add3(c->a, mul3(c->oldbone, t));
        s3    point_on_oldbone_v    = sub3(point_on_oldbone, c->a);        /* L95 */
        s3    point_from_oldbone_v  = sub3(position, point_on_oldbone);    /* L96 */
        s3    point_from_halfbone_v = sub3(point_on_oldbone, c->midp);     /* L97 */

        float dist;
        if (len3(point_from_halfbone_v) < c->half_len) /* L103 */
            dist = len3(point_from_oldbone_v);
        else
        {
            float d0 = len3(sub3(position, c->a)), d1 = len3(sub3(position, c->b));
            dist     = d0 < d1 ? d0 : d1; /* min() */
        }
        float diff = dist - c->effect; /* L113 */
        if (diff < 0.0f)
        {
            s3    point_on_currbone_v   = point_on_oldbone_v;
            float remdist               = c->effect - dist;                  /* L124 */
            s3    point_from_currbone_v = qrot(c->rot_quat, point_from_oldbone_v); /* L138 */
            s3    currnormal            = qrot(c->rot_quat, normal);         /* L139 */
            if (c->has_axis) /* L149-155 */
            {
                point_on_currbone_v   = qrot(c->axis_quat, point_on_currbone_v);
                point_from_currbone_v = qrot(c->axis_quat, point_from_currbone_v);
                currnormal            = qrot(c->axis_quat, currnormal);
            }
            s3 currpos = add3(add3(c->newa, point_on_currbone_v), point_from_currbone_v); /* L157 */
            if (corner_count == 0) corner_center = currpos;                                /* L159 */
            corner_center = add3(corner_center, div3(sub3(currpos, corner_center), 2.0f)); /* L160 */
            corner_points[corner_count]  = currpos;
            corner_normals[corner_count] = currnormal;
            corner_tozerow[corner_count] = remdist;
            corner_count++;
            tozerow_sum += remdist;
        }
    }

    s3 pnt = corner_center;     /* L171 */
    s3 nrm = corner_normals[0]; /* L172 */
    if (corner_count > 1)
    {
        for (int i = 0; i < corner_count; i++)
        {
            float rat = QB_DIV(corner_tozerow[i], tozerow_sum);    /* L178 */
            s3    dir = sub3(corner_points[i], corner_center);     /* L179 */
            pnt       = add3(pnt, mul3(dir, rat));                 /* L180 */
            nrm       = div3(add3(nrm, corner_normals[i]), 2.0f);  /* L181 */
        }
    }
    out_nrm[0] = nrm.x, out_nrm[1] = nrm.y, out_nrm[2] = nrm.z;
    if (out_pnt) out_pnt[0] = pnt.x, out_pnt[1] = pnt.y, out_pnt[2] = pnt.z;

    /* L188-212: cube.w halves per level; only cube.w is ever read */
    float w = basecube[3];
    for (int level = 0; level < 12; level++) out_digits[level] = 0;
    for (int level = 0; level < maxlevel && level < 12; level++)
    {
        float size  = QB_DIV(w, 2.0f);
        int   octet = ((int) floorf(QB_DIV(pnt.x, size))) % 2;
        int   yi    = ((int) floorf(QB_DIV(pnt.y, size))) % 2;
        int   zi    = ((int) floorf(QB_DIV(pnt.z, size))) % 2;
        if (yi == 0) octet += 2;
        if (zi == 0) octet += 4;
        w                 = size;
        out_digits[level] = octet;
    }
}

/* n points: positions / normals float[3n] -> digits int32[12n] (oct14 | oct54 | oct94 per point), normals
 * float[3n], skinned positions float[3n] (optional) */
/* The ten bone pairs' rotations as 9 floats each: rot_quat[4], axis_quat[4], has_axis (0 / 1).  sin, cos and acos
 * have implementation-defined precision in GLSL ES (spec 4.5.1), so these 90 numbers are the one part of the shader
 * whose bits depend on the GL driver; everything downstream of them is + - * / sqrt. */
void qb_oracle_bone_rotations(const float* oldbones80, const float* newbones80, float* out90)
{
    qb_bone_consts bc[10];
    qb_oracle_bone_consts(oldbones80, newbones80, bc);
    for (int k = 0; k < 10; k++)
    {
        float* o = out90 + k * 9;
        o[0] = bc[k].rot_quat.x, o[1] = bc[k].rot_quat.y, o[2] = bc[k].rot_quat.z, o[3] = bc[k].rot_quat.w;
        o[4] = bc[k].axis_quat.x, o[5] = bc[k].axis_quat.y, o[6] = bc[k].axis_quat.z, o[7] = bc[k].axis_quat.w;
        o[8] = (float) bc[k].has_axis;
    }
}

void qb_oracle_skin_rot(const float* oldbones80, const float* newbones80, const float* rotations90,
                        const float basecube[4], int maxlevel, int64_t n, const float* positions,
                        const float* normals, int32_t* digits12, float* normals_out, float* points_out);

void qb_oracle_skin(const float* oldbones80, const float* newbones80, const float basecube[4], int maxlevel,
                    int64_t n, const float* positions, const float* normals, int32_t* digits12, float* normals_out,
                    float* points_out)
{
    qb_oracle_skin_rot(oldbones80, newbones80, NULL, basecube, maxlevel, n, positions, normals, digits12, normals_out,
                       points_out);
}

/* rotations90 != NULL: use these per-bone rotations (e.g. the ones a GL driver produced, oracle/glsl_ref.c mode 22)
 * instead of the libm ones */
void qb_oracle_skin_rot(const float* oldbones80, const float* newbones80, const float* rotations90,
                        const float basecube[4], int maxlevel, int64_t n, const float* positions,
                        const float* normals, int32_t* digits12, float* normals_out, float* points_out)
{
    qb_bone_consts bc[10];
    qb_oracle_bone_consts(oldbones80, newbones80, bc);
    if (rotations90)
        for (int k = 0; k < 10; k++)
        {
            const float* o   = rotations90 + k * 9;
            bc[k].rot_quat   = (s4){o[0], o[1], o[2], o[3]};
            bc[k].axis_quat  = (s4){o[4], o[5], o[6], o[7]};
            bc[k].has_axis   = o[8] != 0.0f;
        }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        s3 p  = {positions[i * 3], positions[i * 3 + 1], positions[i * 3 + 2]};
        s3 nr = {normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]};
        skin_point(bc, p, nr, basecube, maxlevel, digits12 + i * 12, normals_out + i * 3,
                   points_out ? points_out + i * 3 : NULL);
    }
}
