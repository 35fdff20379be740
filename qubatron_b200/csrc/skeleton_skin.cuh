// skeleton_skin.cuh -- per-point skinning by up to 10 bone pairs and the 12 octant digits of the skinned point:
// the reference's transform-feedback vertex program skeleton_vsh.c ("next" row SURVEY 8f #1, first half).
//
// Behaviour to reproduce: /root/reference/src/qubatron/shaders/skeleton_vsh.c main() L74-226 with the uniforms
// of skeleton_glc.c L222-227.  Everything that needs sin / cos / acos depends on the bones only (the shader's own
// TODO at L72) and arrives precomputed in BoneConsts (host libm, octree_cuc.cu); the per-point work below is
// + - * / sqrt, compiled with -fmad=false, so it reproduces the CPU restatement bit for bit.
// The outputs stay on the device: digits feed the tree build (octree_build.cuh), normals go straight into the
// dynamic model's point records.
#pragma once
#include "octree_build.cuh"
#include "octree_render.cuh"

namespace qb
{

struct BoneConsts
{
    float a[3], b[3];   // oldbones[i].xyz, oldbones[i+1].xyz
    float effect;       // oldbones[i].w
    float oldbone[3];   // b - a
    float midp[3];      // a + oldbone / 2
    float half_len;     // length(oldbone) / 2
    float ab_dot;       // dot(oldbone, oldbone)
    float newa[3];      // newbones[i].xyz
    float rot_quat[4];  // rotation about the old bone by newbones[i].w
    int   has_axis;     // bones not parallel
    float axis_quat[4]; // rotation old bone -> current bone
    float cull_r2;      // (half_len + effect + 1)^2, rounded up: farther from midp than this cannot be in range
};

struct SkinParams
{
    BoneConsts bones[10];
    float      basesize;
    int        maxlevel;
};

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 mul3(float3 a, float f) { return make_float3(a.x * f, a.y * f, a.z * f); }
__device__ __forceinline__ float  len3(float3 a) { return sqrtf(dot3(a, a)); }
template <int DIV>
__device__ __forceinline__ float3 half3(float3 a)
{
    return make_float3(qdiv<DIV>(a.x, 2.0f), qdiv<DIV>(a.y, 2.0f), qdiv<DIV>(a.z, 2.0f));
}

template <int DIV>
__global__ void skin_kernel(const SkinParams S, size_t n, const float* __restrict__ positions,
                            const float* __restrict__ normals, int4* __restrict__ p14, int4* __restrict__ p54,
                            int4* __restrict__ p94, float* __restrict__ rec, float* __restrict__ pnt_out,
                            unsigned long long* __restrict__ build_keys)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 position = ld3(positions + i * 3);
    const float3 normal   = ld3(normals + i * 3);

    // L82-88.  The shader stores every in-range bone's result and blends afterwards; the blend needs the final
    // centre, so the per-bone results are kept (10 x 7 floats of thread-local memory at most).
    float3 corner_points[10];
    float3 corner_normals[10];
    float  corner_tozerow[10];
    int    corner_count  = 0;
    float3 corner_center = position;
    float  tozerow_sum   = 0.0f;
    corner_normals[0]    = make_float3(0.0f, 0.0f, 0.0f);

#pragma unroll 1
    for (int k = 0; k < 10; k++)
    {
        const BoneConsts& c = S.bones[k];
        // Exact cull.  In real arithmetic `dist` below is the distance to the bone SEGMENT (perpendicular distance
        // when the foot lies inside it, else the nearer end), hence >= |position - midp| - half_len.  A point
        // farther from midp than half_len + effect + 1 has dist > effect + 1; fp32 rounding of dist is orders of
        // magnitude below that margin, so the reference takes the `diff >= 0` branch and the bone contributes
        // nothing.  (NaN compares false and falls through to the full evaluation.)
        const float3 from_mid = sub3(position, ld3(c.midp));
        if (dot3(from_mid, from_mid) > c.cull_r2) continue;
        const float3      A  = ld3(c.a);
        const float3      AB = ld3(c.oldbone);
        const float3      AC = sub3(position, A);
        const float       t  = qdiv<DIV>(dot3(AC, AB), c.ab_dot);      // L39
        const float3 point_on_oldbone      = add3(A, mul3(AB, t));      // L94
        const float3 point_on_oldbone_v    = sub3(point_on_oldbone, A); // L95
        const float3 point_from_oldbone_v  = sub3(position, point_on_oldbone);
        const float3 point_from_halfbone_v = sub3(point_on_oldbone, ld3(c.midp));
        float        dist;
        if (len3(point_from_halfbone_v) < c.half_len) // L103
            dist = len3(point_from_oldbone_v);
        else
        {
            const float d0 = len3(sub3(position, A)), d1 = len3(sub3(position, ld3(c.b)));
            dist           = d0 < d1 ? d0 : d1;
        }
        if (dist - c.effect < 0.0f) // L113-114
        {
            float3      point_on_currbone_v   = point_on_oldbone_v;
            const float remdist               = c.effect - dist;
            float3      point_from_currbone_v = quat_rotate(c.rot_quat, point_from_oldbone_v); // L138
            float3      currnormal            = quat_rotate(c.rot_quat, normal);               // L139
            if (c.has_axis) // L149-155
            {
                point_on_currbone_v   = quat_rotate(c.axis_quat, point_on_currbone_v);
                point_from_currbone_v = quat_rotate(c.axis_quat, point_from_currbone_v);
                currnormal            = quat_rotate(c.axis_quat, currnormal);
            }
            const float3 currpos = add3(add3(ld3(c.newa), point_on_currbone_v), point_from_currbone_v); // L157
            if (corner_count == 0) corner_center = currpos;
            corner_center = add3(corner_center, half3<DIV>(sub3(currpos, corner_center))); // L160
            corner_points[corner_count]  = currpos;
            corner_normals[corner_count] = currnormal;
            corner_tozerow[corner_count] = remdist;
            corner_count++;
            tozerow_sum += remdist;
        }
    }

    float3 pnt = corner_center;     // L171
    float3 nrm = corner_normals[0]; // L172
    if (corner_count > 1)
    {
        for (int k = 0; k < corner_count; k++)
        {
            const float rat = qdiv<DIV>(corner_tozerow[k], tozerow_sum); // L178
            pnt             = add3(pnt, mul3(sub3(corner_points[k], corner_center), rat));
            nrm             = half3<DIV>(add3(nrm, corner_normals[k]));
        }
    }
    // normal_out -> the dynamic model's normal (the engine uploads skelglc.nrm_out as DYNAMIC_NORMAL, qubatron.c L521-529)
    rec[i * 8 + 4] = nrm.x;
    rec[i * 8 + 5] = nrm.y;
    rec[i * 8 + 6] = nrm.z;
    rec[i * 8 + 7] = 0.0f;
    if (pnt_out)
    {
        pnt_out[i * 3 + 0] = pnt.x;
        pnt_out[i * 3 + 1] = pnt.y;
        pnt_out[i * 3 + 2] = pnt.z;
    }

    // L188-226: twelve octant digits
    float w = S.basesize;
    int   d[12];
#pragma unroll
    for (int level = 0; level < 12; level++)
    {
        d[level] = 0;
        if (level < S.maxlevel)
        {
            const float size  = qdiv<DIV>(w, 2.0f);
            int         octet = ((int) floorf(qdiv<DIV>(pnt.x, size))) % 2;
            const int   yi    = ((int) floorf(qdiv<DIV>(pnt.y, size))) % 2;
            const int   zi    = ((int) floorf(qdiv<DIV>(pnt.z, size))) % 2;
            if (yi == 0) octet += 2;
            if (zi == 0) octet += 4;
            w        = size;
            d[level] = octet;
        }
    }
    p14[i] = make_int4(d[0], d[1], d[2], d[3]);
    p54[i] = make_int4(d[4], d[5], d[6], d[7]);
    p94[i] = make_int4(d[8], d[9], d[10], d[11]);
    if (build_keys) // the tree build's sort word (octree_build.cuh build_key_kernel), saving it a pass over the digits
    {
        unsigned long long k = 0;
#pragma unroll
        for (int level = 0; level < 12; level++)
            if (level < S.maxlevel) k = (k << 3) | (unsigned long long) (d[level] & 7);
        build_keys[i] = (k << BUILD_INDEX_BITS) | (unsigned long long) i;
    }
}

// normals of the first n point records -> float[3n] (the reference's nrm_out buffer)
__global__ void gather_normals_kernel(const float* __restrict__ rec, size_t n, float* __restrict__ out3)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    out3[i * 3 + 0] = rec[i * 8 + 4];
    out3[i * 3 + 1] = rec[i * 8 + 5];
    out3[i * 3 + 2] = rec[i * 8 + 6];
}

} // namespace qb
