// octree_trace_generic.cuh -- reference-order octree traversal, any base cube.
//
// One thread = one ray.  This is the variant that keeps the full per-level
// state of the reference pixel program (cube, candidate points, octants) in a
// thread-local stack, so it is valid for every `basesize` / `maxlevel` <= 17.
// The register-stack variant for exactly representable grids lives in
// octree_trace_fast.cuh; both must agree bit for bit with the oracle.
//
// Behaviour to reproduce: /root/reference/src/qubatron/shaders/octree_fsh.c
//   L62-99 (plane hits), L127-136 (child lookup), L138-379 (cube_trace_line).
// The file is compiled with -fmad=false: every mul/add below is rounded
// separately, divisions are IEEE (-prec-div=true), denormals kept (-ftz=false).
#pragma once
#include "octree_types.cuh"
#include <cfloat>

namespace qb
{

struct TraceResult
{
    float ix, iy, iz, iw; // res.isp
    int   status;         // 0 miss, 1 leaf, -1 discard
    int   node_s, node_d; // leaf nodes
    int   model_s, model_d;
    float tx, ty, tz, tw; // res.tlf: the leaf cube (generic tracer only)
};

struct RayCounters
{
    unsigned int v[CNT_COUNT];
};

// Division semantics.  The reference's `/` has two reproducible executions:
//   DIV_GLSL: Mesa lowers the shader's a / b to a * (1.0 / b) (lower_instructions
//             DIV_TO_MUL_RCP; llvmpipe evaluates 1.0 / b with an IEEE divide): two
//             roundings.  This is what the reference SHADER computes when run
//             headless, and the connector's default.
//   DIV_IEEE: one correctly rounded divide, what the reference's compiled CPU twin
//             octree_trace_line (octree.c L302-339) computes.
constexpr int DIV_GLSL = 0;
constexpr int DIV_IEEE = 1;

template <int DIV>
__device__ __forceinline__ float qdiv(float n, float d)
{
    if (DIV == DIV_GLSL) return n * (1.0f / d);
    return n / d;
}

// The traversal exists in three places of the reference with small, deliberate-or-not differences; TWIN selects whose
// behaviour is reproduced:
//   TRACE_FSH      the pixel program octree_fsh.c (the hot path)
//   TRACE_CPU      the engine's CPU function octree_trace_line (octree.c L341-537)
//   TRACE_PARTICLE the particle program particle_vsh.c L67-343 (static tree only)
constexpr int TRACE_FSH      = 0;
constexpr int TRACE_CPU      = 1;
constexpr int TRACE_PARTICLE = 2;

// parallel ray: the pixel program's sentinel is FLT_MAX in every component (octree_fsh.c L64); the CPU function
// uses (0,0,0,FLT_MAX) (octree.c L304, L317, L330) and the particle program vec4(0.0) (particle_vsh.c L69, L82,
// L95), whose xyz CAN pass a range test -- kept as they are
template <int TWIN>
__device__ __forceinline__ float4 parallel_sentinel()
{
    if (TWIN == TRACE_CPU) return make_float4(0.0f, 0.0f, 0.0f, FLT_MAX);
    if (TWIN == TRACE_PARTICLE) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    return make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
}

// ray / axis-plane intersection (octree_fsh.c L62-99): w = (c - o)/d, the other
// two coordinates o + d*w, the plane coordinate exactly c.  A ray parallel to
// the plane yields FLT_MAX in every component so all range tests fail.
template <int DIV, int TWIN = 0>
__device__ __forceinline__ float4 plane_hit_x(float c, float3 o, float3 d)
{
    float4 r = parallel_sentinel<TWIN>();
    if (d.x != 0.0f)
    {
        r.w = qdiv<DIV>(c - o.x, d.x);
        r.y = o.y + d.y * r.w;
        r.z = o.z + d.z * r.w;
        r.x = c;
    }
    return r;
}
template <int DIV, int TWIN = 0>
__device__ __forceinline__ float4 plane_hit_y(float c, float3 o, float3 d)
{
    float4 r = parallel_sentinel<TWIN>();
    if (d.y != 0.0f)
    {
        r.w = qdiv<DIV>(c - o.y, d.y);
        r.x = o.x + d.x * r.w;
        r.z = o.z + d.z * r.w;
        r.y = c;
    }
    return r;
}
template <int DIV, int TWIN = 0>
__device__ __forceinline__ float4 plane_hit_z(float c, float3 o, float3 d)
{
    float4 r = parallel_sentinel<TWIN>();
    if (d.z != 0.0f)
    {
        r.w = qdiv<DIV>(c - o.z, d.z);
        r.x = o.x + d.x * r.w;
        r.y = o.y + d.y * r.w;
        r.z = c;
    }
    return r;
}

// the same three with the quotient supplied by the caller (per-ray reciprocals hoisted)
template <class Q>
__device__ __forceinline__ float4 plane_hit_x_q(float c, float3 o, float3 d, Q q)
{
    float4 r = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
    if (d.x != 0.0f)
    {
        r.w = q(c - o.x, 0);
        r.y = o.y + d.y * r.w;
        r.z = o.z + d.z * r.w;
        r.x = c;
    }
    return r;
}
template <class Q>
__device__ __forceinline__ float4 plane_hit_y_q(float c, float3 o, float3 d, Q q)
{
    float4 r = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
    if (d.y != 0.0f)
    {
        r.w = q(c - o.y, 1);
        r.x = o.x + d.x * r.w;
        r.z = o.z + d.z * r.w;
        r.y = c;
    }
    return r;
}
template <class Q>
__device__ __forceinline__ float4 plane_hit_z_q(float c, float3 o, float3 d, Q q)
{
    float4 r = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
    if (d.z != 0.0f)
    {
        r.w = q(c - o.z, 2);
        r.x = o.x + d.x * r.w;
        r.y = o.y + d.y * r.w;
        r.z = c;
    }
    return r;
}

// half-open cube ranges: x in (x0, x1], y in [y0, y1), z in [z0, z1)
struct Cube
{
    float x0, x1; // tlf.x, brb.x
    float y1, y0; // tlf.y, brb.y
    float z1, z0; // tlf.z, brb.z
};
__device__ __forceinline__ bool in_x(const Cube& c, float x) { return c.x0 < x && x <= c.x1; }
__device__ __forceinline__ bool in_y(const Cube& c, float y) { return c.y1 > y && y >= c.y0; }
__device__ __forceinline__ bool in_z(const Cube& c, float z) { return c.z1 > z && z >= c.z0; }

// the 8 child indices (device indices, octree_types.cuh) of a node.  Device node 0 is the empty dummy that "this
// tree has nothing here" (octree_fsh.c L129) leads to; nodes past the uploaded range read 0.
struct Children
{
    int4 lo, hi;
};
__device__ __forceinline__ Children load_children(const TreeDev& t, int node, int /*level*/)
{
    Children c;
    node = tree_clamp(t, node);
    c.lo = __ldg(t.child + 2 * (size_t) node);
    c.hi = __ldg(t.child + 2 * (size_t) node + 1);
    c.lo.x &= (int) CHILD_INDEX_MASK; // words 0/1 carry the child-exists mask in their top nibble
    c.lo.y &= (int) CHILD_INDEX_MASK;
    return c;
}
__device__ __forceinline__ int child_of(const Children& c, int oct)
{
    int a = (oct & 1) ? c.lo.y : c.lo.x;
    int b = (oct & 1) ? c.lo.w : c.lo.z;
    int e = (oct & 1) ? c.hi.y : c.hi.x;
    int f = (oct & 1) ? c.hi.w : c.hi.z;
    int l = (oct & 2) ? b : a;
    int h = (oct & 2) ? f : e;
    return (oct & 4) ? h : l;
}
__device__ __forceinline__ int model_of(const TreeDev& t, int node, int /*level*/)
{
    return __ldg(t.model + tree_clamp(t, node));
}

// base-cube entry (octree_fsh.c L157-211).  Returns false on `discard`.
// q(n, axis) returns n / dir[axis] in the caller's division semantics.
template <class Q>
__device__ __forceinline__ bool base_cube_entry_q(const float* basecube, float3 pos, float3 dir, float4& entry, Q q)
{
    Cube c;
    c.x0 = basecube[0];
    c.x1 = basecube[0] + basecube[3];
    c.y1 = basecube[1];
    c.y0 = basecube[1] - basecube[3];
    c.z1 = basecube[2];
    c.z0 = basecube[2] - basecube[3];

    int    hitc = 0;
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, act;

#define QB_FACE(ACT, COND)                                                                                            \
    act = ACT;                                                                                                        \
    if (COND)                                                                                                         \
    {                                                                                                                 \
        if (hitc == 0) h0 = act;                                                                                      \
        if (hitc == 1) h1 = act;                                                                                      \
        hitc++;                                                                                                       \
    }
    QB_FACE(plane_hit_z_q(c.z1, pos, dir, q), in_x(c, act.x) && in_y(c, act.y)) // front
    QB_FACE(plane_hit_z_q(c.z0, pos, dir, q), in_x(c, act.x) && in_y(c, act.y)) // back
    QB_FACE(plane_hit_x_q(c.x0, pos, dir, q), in_y(c, act.y) && in_z(c, act.z)) // left
    QB_FACE(plane_hit_x_q(c.x1, pos, dir, q), in_y(c, act.y) && in_z(c, act.z)) // right
    QB_FACE(plane_hit_y_q(c.y1, pos, dir, q), in_x(c, act.x) && in_z(c, act.z)) // top
    QB_FACE(plane_hit_y_q(c.y0, pos, dir, q), in_x(c, act.x) && in_z(c, act.z)) // bottom
#undef QB_FACE

    if (hitc < 2) return false;                   // L195
    if (h0.w < 0.0f && h1.w < 0.0f) return false; // L198
    if (h1.w < h0.w) h0 = h1;                     // L205
    if (h0.w < 0.0f) h0 = make_float4(pos.x, pos.y, pos.z, 0.0f); // L208
    entry = h0;
    return true;
}

template <int DIV, int TWIN = 0>
__device__ __forceinline__ bool base_cube_entry(const float* basecube, float3 pos, float3 dir, float4& entry)
{
    Cube c;
    c.x0 = basecube[0];
    c.x1 = basecube[0] + basecube[3];
    c.y1 = basecube[1];
    c.y0 = basecube[1] - basecube[3];
    c.z1 = basecube[2];
    c.z0 = basecube[2] - basecube[3];

    int    hitc = 0;
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, act;

#define QB_FACE(ACT, COND)                                                                                            \
    act = ACT;                                                                                                        \
    if ((TWIN != TRACE_CPU || act.w < FLT_MAX) && (COND)) /* octree.c L360-386: w < FLT_MAX on the six faces */     \
    {                                                                                                                 \
        if (hitc == 0) h0 = act;                                                                                      \
        if (hitc == 1) h1 = act;                                                                                      \
        hitc++;                                                                                                       \
    }
    QB_FACE((plane_hit_z<DIV, TWIN>(c.z1, pos, dir)), in_x(c, act.x) && in_y(c, act.y)) // front
    QB_FACE((plane_hit_z<DIV, TWIN>(c.z0, pos, dir)), in_x(c, act.x) && in_y(c, act.y)) // back
    QB_FACE((plane_hit_x<DIV, TWIN>(c.x0, pos, dir)), in_y(c, act.y) && in_z(c, act.z)) // left
    QB_FACE((plane_hit_x<DIV, TWIN>(c.x1, pos, dir)), in_y(c, act.y) && in_z(c, act.z)) // right
    QB_FACE((plane_hit_y<DIV, TWIN>(c.y1, pos, dir)), in_x(c, act.x) && in_z(c, act.z)) // top
    QB_FACE((plane_hit_y<DIV, TWIN>(c.y0, pos, dir)), in_x(c, act.x) && in_z(c, act.z)) // bottom
#undef QB_FACE

    if (hitc < 2) return false;                   // L195
    if (h0.w < 0.0f && h1.w < 0.0f) return false; // L198
    if (h1.w < h0.w) h0 = h1;                     // L205
    if (h0.w < 0.0f) h0 = make_float4(pos.x, pos.y, pos.z, 0.0f); // L208
    entry = h0;
    return true;
}

// per-level record of the reference's stck_t (octree_fsh.c L50-58); octants of
// the four candidate slots are packed 3 bits each
struct GenericLevel
{
    float4 cube;
    float4 isps[4];
    int    octs;
    int    ispsi;
    int    socti;
    int    docti;
};

constexpr int GENERIC_STACK = 18; // octree_fsh.c L151

template <int DIV, bool COUNT, int TWIN = 0>
__device__ __noinline__ TraceResult trace_generic(const FrameParams& P, float3 pos, float3 dir, RayCounters& cnt)
{
    TraceResult res;
    res.ix = res.iy = res.iz = res.iw = 0.0f;
    res.status                       = 0;
    res.node_s = res.node_d = -1;
    res.model_s = res.model_d = 0;
    res.tx = res.ty = res.tz = res.tw = 0.0f;

    float4 entry;
    if (!base_cube_entry<DIV, TWIN>(P.basecube, pos, dir, entry))
    {
        res.status = -1;
        return res;
    }

    GenericLevel stck[GENERIC_STACK];
    int          level = 0;
    stck[0].cube       = make_float4(P.basecube[0], P.basecube[1], P.basecube[2], P.basecube[3]);
    stck[0].socti      = ROOT_NODE;
    stck[0].docti      = TWIN == TRACE_PARTICLE ? 0 : ROOT_NODE;
    stck[0].ispsi      = 0;
    stck[0].octs       = 0;
    stck[0].isps[0]    = entry;

    const int maxlevel = P.maxlevel;

    for (;;)
    {
        float4 tlf = stck[level].cube;

        if (level == maxlevel) // L218-248
        {
            float4 isp  = stck[level].isps[0];
            res.ix      = isp.x;
            res.iy      = isp.y;
            res.iz      = isp.z;
            res.iw      = isp.w;
            res.status  = 1;
            res.tx = tlf.x, res.ty = tlf.y, res.tz = tlf.z, res.tw = tlf.w;
            res.node_s  = ref_node(stck[level].socti);
            res.node_d  = ref_node(stck[level].docti);
            res.model_s = model_of(P.tree_s, stck[level].socti, level);
            res.model_d = model_of(P.tree_d, stck[level].docti, level);
            if (COUNT)
            {
                if (stck[level].socti != 0) cnt.v[CNT_LEAF_S]++;
                if (stck[level].docti != 0) cnt.v[CNT_LEAF_D]++;
            }
            return res;
        }

        const int sn = stck[level].socti;
        const int dn = stck[level].docti;
        // both trees' child blocks: needed by the expansion and by the descent,
        // not by a level that is only passed through while backtracking
        Children cs, cd;
        cs.lo = cs.hi = cd.lo = cd.hi = make_int4(0, 0, 0, 0);
        if (stck[level].ispsi == 0 || (stck[level].ispsi & 0x0F) != 0)
        {
            cs = load_children(P.tree_s, sn, level);
            if (TWIN != TRACE_PARTICLE) cd = load_children(P.tree_d, dn, level);
            // particle_vsh.c L108-119 addresses texel (3i mod 8192) + octi/4 of row 3i / 8192 WITHOUT wrapping into the
            // next row (octree_fsh.c L130-135 wraps): for a node whose first texel is the last of its row, children
            // 4..7 are fetched outside the texture and read 0
            if (TWIN == TRACE_PARTICLE && sn > 0 && ((3 * (sn - 1)) & 8191) == 8191) cs.hi = make_int4(0, 0, 0, 0);
        }

        if (stck[level].ispsi == 0) // L251-330
        {
            if (COUNT)
            {
                if (sn != 0) cnt.v[CNT_EXPAND_S]++;
                if (dn != 0) cnt.v[CNT_EXPAND_D]++;
            }
            Cube c;
            c.x0 = tlf.x;
            c.x1 = tlf.x + tlf.w;
            c.y1 = tlf.y;
            c.y0 = tlf.y - tlf.w;
            c.z1 = tlf.z;
            c.z0 = tlf.z - tlf.w;
            // hlf = brb + (tlf - brb) * 0.5
            const float hx = c.x1 + (c.x0 - c.x1) * 0.5f;
            const float hy = c.y0 + (c.y1 - c.y0) * 0.5f;
            const float hz = c.z0 + (c.z1 - c.z0) * 0.5f;

            float4 hp[4];
            int    hc = 1;
            hp[0]     = stck[level].isps[0];
            float4 act;
            act = plane_hit_z<DIV, TWIN>(hz, pos, dir);
            if (act.w > 0.0f && in_x(c, act.x) && in_y(c, act.y)) hp[hc++] = act;
            act = plane_hit_x<DIV, TWIN>(hx, pos, dir);
            if (act.w > 0.0f && in_y(c, act.y) && in_z(c, act.z)) hp[hc++] = act;
            act = plane_hit_y<DIV, TWIN>(hy, pos, dir);
            if (act.w > 0.0f && in_x(c, act.x) && in_z(c, act.z)) hp[hc++] = act;

            int pre  = -1;
            int len  = 0;
            int octs = 0;
            for (int i = 0; i < hc; ++i)
            {
                for (int j = i + 1; j < hc; ++j) // L279-290 exchange sort, strict <
                {
                    if (hp[j].w < hp[i].w)
                    {
                        float4 t = hp[i];
                        hp[i]    = hp[j];
                        hp[j]    = t;
                    }
                }
                act     = hp[i];
                int oct = 0;
                if (act.x > hx) oct = 1;
                if (act.y < hy) oct += 2;
                if (act.z < hz) oct += 4;
                if (oct == pre) // L301-309
                {
                    if (act.x == hx)
                        oct ^= 1;
                    else if (act.y == hy)
                        oct ^= 2;
                    else if (act.z == hz)
                        oct ^= 4;
                }
                pre = oct;
                if (child_of(cs, oct) > 0 || child_of(cd, oct) > 0) // L317
                {
                    stck[level].isps[len] = act;
                    octs |= oct << (3 * len);
                    len++;
                }
            }
            stck[level].octs  = octs;
            stck[level].ispsi = 128 | len;
        }

        int state   = stck[level].ispsi;
        int cur_len = state & 0x0F;
        if (cur_len > 0) // L336-367
        {
            int    nxt_ind = (state >> 4) & 7;
            float4 nxt_isp = stck[level].isps[nxt_ind];
            int    nxt_oct = (stck[level].octs >> (3 * nxt_ind)) & 7;

            float halfs = tlf.w / 2.0f;
            tlf.x += ((nxt_oct & 1) ? 1.0f : 0.0f) * halfs;
            tlf.y -= ((nxt_oct & 2) ? 1.0f : 0.0f) * halfs;
            tlf.z -= ((nxt_oct & 4) ? 1.0f : 0.0f) * halfs;
            tlf.w = halfs;

            stck[level].ispsi = 128 | ((nxt_ind + 1) << 4) | (cur_len - 1);

            int socti = child_of(cs, nxt_oct);
            int docti = child_of(cd, nxt_oct);

            level += 1;
            if (COUNT) cnt.v[CNT_DESCENTS]++;
            stck[level].cube    = tlf;
            stck[level].ispsi   = 0;
            stck[level].octs    = 0;
            stck[level].socti   = socti;
            stck[level].docti   = docti;
            stck[level].isps[0] = nxt_isp;
        }
        else // L368-375
        {
            stck[level].ispsi = 0;
            level--;
            if (level < 0) return res;
        }
    }
}

} // namespace qb
