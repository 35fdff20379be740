// octree_view_host.h -- the per-frame constants of octree_glc_update evaluated on the HOST in fp32: everything main()
// of octree_fsh.c derives from uniforms that is constant over the frame and the render size
// of octree_glc.c L263-284.  Plain C++ (no CUDA calls): included by octree_cuc.cu, and by the host build of the
// traversal that tests/host_emu/ uses for logic tests.  Compile with -ffp-contract=off (csrc/Makefile does).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "octree_types.cuh"

namespace qb
{
namespace viewhost
{

// smallest float d with acosf(d) < 0.02f: the host-libm form of the light-disc
// test `camangle < 0.02` (octree_fsh.c L418, L452).  acosf is monotone over the
// scanned interval (checked by tests/test_oracle.py::test_disc_threshold).
inline float disc_dot_min()
{
    static float cached = 0.0f;
    if (cached != 0.0f) return cached;
    uint32_t lo, hi;
    float    flo = 0.99f, fhi = 1.0f;
    memcpy(&lo, &flo, 4);
    memcpy(&hi, &fhi, 4);
    while (lo < hi) // acosf(lo) >= 0.02 (false), acosf(hi) = 0 < 0.02 (true)
    {
        uint32_t mid = lo + (hi - lo) / 2;
        float    fm;
        memcpy(&fm, &mid, 4);
        if (acosf(fm) < 0.02f)
            hi = mid;
        else
            lo = mid + 1;
    }
    memcpy(&cached, &hi, 4);
    return cached;
}

inline void host_cross(const float* a, const float* b, float* r)
{
    volatile float x = a[1] * b[2] - b[1] * a[2];
    volatile float y = a[2] * b[0] - b[2] * a[0];
    volatile float z = a[0] * b[1] - b[0] * a[1];
    r[0] = x, r[1] = y, r[2] = z;
}
// octree_fsh.c L392-395 evaluated on the host in fp32 (per-frame constants)
inline void host_quat_rotate(const float* q, const float* v, float* out)
{
    float c1[3], t[3], c2[3];
    host_cross(q, v, c1);
    for (int i = 0; i < 3; i++)
    {
        volatile float m = q[3] * v[i];
        volatile float s = c1[i] + m;
        t[i]             = s;
    }
    host_cross(q, t, c2);
    for (int i = 0; i < 3; i++)
    {
        volatile float m = 2.0f * c2[i];
        volatile float s = v[i] + m;
        out[i]           = s;
    }
}
inline void host_quat_axis_angle(const float* axis, float angle, float* q)
{
    volatile float half = angle * 0.5f;
    float          sn = sinf(half), cs = cosf(half);
    for (int i = 0; i < 3; i++)
    {
        volatile float m = axis[i] * sn;
        q[i]             = m;
    }
    q[3] = cs;
}

// light_override: null, or the light position to use instead of octree_glc.c L264's; div_glsl: normalize() as
// v * (1 / len) (the GLSL lowering) instead of v / len
inline void fill_view(ViewParams& V, const float* position, const float* angle, float lighta, int shoot,
                      const float* light_override, bool div_glsl)
{
    const float lightc[3] = {420.0f, 200.0f, 680.0f}; // octree_glc.c L91
    V.camfp[0]            = position[0];
    V.camfp[1]            = position[1];
    V.camfp[2]            = position[2];
    if (light_override)
    {
        V.light[0] = light_override[0];
        V.light[1] = light_override[1];
        V.light[2] = light_override[2];
    }
    else
    {
        // octree_glc.c L264: double arithmetic, rounded once on store
        V.light[0] = lightc[0];
        V.light[1] = (float) ((double) lightc[1] - (double) sinf(lighta) * 20.0);
        V.light[2] = (float) ((double) lightc[2] - (double) sinf(lighta) * 200.0);
    }
    const float yaxis[3] = {0.0f, 1.0f, 0.0f};
    const float negx[3]  = {-1.0f, 0.0f, 0.0f};
    float       vx[3];
    host_quat_axis_angle(yaxis, -angle[0], V.qz); // octree_fsh.c L406-408
    host_quat_rotate(V.qz, negx, vx);
    host_quat_axis_angle(vx, -angle[1], V.qx);

    float cl[3];
    for (int i = 0; i < 3; i++)
    {
        volatile float d = V.light[i] - V.camfp[i];
        cl[i]            = d;
    }
    volatile float xx = cl[0] * cl[0], yy = cl[1] * cl[1], zz = cl[2] * cl[2];
    volatile float s1 = xx + yy;
    volatile float s2 = s1 + zz;
    float          l  = sqrtf(s2);
    volatile float inv = 1.0f / l;
    for (int i = 0; i < 3; i++)
    {
        // normalize(): v * (1/len) under the GLSL lowering, v / len with IEEE division
        volatile float q = div_glsl ? cl[i] * inv : cl[i] / l;
        V.camlight_n[i]  = q;
    }
    V.disc_dot_min = disc_dot_min();
    V.shoot        = shoot;
}

// octree_glc.c L263-284: render size (double arithmetic, rounded once) and viewport
inline void render_size(float width, float height, uint8_t quality, float& ow, float& oh, int& W, int& H)
{
    ow = (float) ((double) width / (6.0 - (double) (float) quality / 2.0));
    oh = (float) ((double) height / (6.0 - (double) (float) quality / 2.0));
    W  = (int) ow;
    H  = (int) oh;
}

// octree_fsh.c L403 (tan(PI/4.0) folds to 1.0f)
inline void fill_cfp(ViewParams& V, float ow, float oh)
{
    V.cfp[0] = ow / 2.0f;
    V.cfp[1] = oh / 2.0f;
    V.cfp[2] = (ow / 2.0f) / 1.0f;
}

} // namespace viewhost
} // namespace qb
