// present.cuh -- the reference's presentation pass ("next" row SURVEY 8f #4, second half).
//
// Behaviour to reproduce: /root/reference/src/qubatron/octree_glc.c L308-351.  The render target (2048 x 2048 RGBA8,
// cleared to (0,0,0,0), the frame in its lower-left (int)ow x (int)oh texels) is drawn as a LINEAR-filtered textured
// quad (shaders/texquad_vsh.c / texquad_fsh.c) into the width x height window -- all four channels stored, blending
// is ONE/ZERO -- and a 2 x 2 white crosshair is cleared at the centre.  GL leaves the filter precision to the
// implementation; this follows Mesa llvmpipe's RGBA8 path exactly (oracle/present_oracle.c, pinned on every pixel of
// tests/golden/present_*.npz): 8.8 fixed-point texel coordinate k = round(X * 256) - 128, X = (p + 0.5) * ow / width,
// weight k & 255, lerp(a, b, w) = a + (((b - a) * w) >> 8) per channel, x first, then y; REPEAT wrap (the texel left
// of column 0 is column 2047, i.e. cleared).  At quality 10 all weights are 0: a copy plus the crosshair.
// One thread per window pixel; the four taps come from L2 (the frame was just written).
#pragma once
#include "octree_types.cuh"

namespace qb
{

struct PresentParams
{
    const uchar4* frame; // rendered frame, row 0 = bottom
    size_t        pitch; // pixels per frame row
    int           vp_w, vp_h;
    double        sx, sy; // ow / width, oh / height
    int           width, height;
    uchar4*       window; // width x height, row 0 = bottom
};

__device__ __forceinline__ int lerp8(int a, int b, int w) { return a + (((b - a) * w) >> 8); }

__device__ __forceinline__ uchar4 present_texel(const PresentParams& P, int x, int y)
{
    if (x < P.vp_w && y < P.vp_h) return P.frame[(size_t) y * P.pitch + x];
    return make_uchar4(0, 0, 0, 0); // the cleared rest of the 2048 x 2048 render target
}

__global__ void present_kernel(const PresentParams P)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= P.width || py >= P.height) return;
    uchar4 out;
    if (px >= P.width / 2 - 1 && px < P.width / 2 + 1 && py >= P.height / 2 - 1 && py < P.height / 2 + 1)
        out = make_uchar4(255, 255, 255, 255); // glScissor(width/2 - 1, height/2 - 1, 2, 2) + white clear
    else
    {
        const long long kx = __double2ll_rn(((double) px + 0.5) * P.sx * 256.0) - 128;
        const long long ky = __double2ll_rn(((double) py + 0.5) * P.sy * 256.0) - 128;
        const int x0 = (int) ((kx >> 8) & 2047), x1 = (x0 + 1) & 2047, wx = (int) (kx & 255);
        const int y0 = (int) ((ky >> 8) & 2047), y1 = (y0 + 1) & 2047, wy = (int) (ky & 255);
        const uchar4 a = present_texel(P, x0, y0), b = present_texel(P, x1, y0);
        const uchar4 c = present_texel(P, x0, y1), d = present_texel(P, x1, y1);
        out.x = (unsigned char) lerp8(lerp8(a.x, b.x, wx), lerp8(c.x, d.x, wx), wy);
        out.y = (unsigned char) lerp8(lerp8(a.y, b.y, wx), lerp8(c.y, d.y, wx), wy);
        out.z = (unsigned char) lerp8(lerp8(a.z, b.z, wx), lerp8(c.z, d.z, wx), wy);
        out.w = (unsigned char) lerp8(lerp8(a.w, b.w, wx), lerp8(c.w, d.w, wx), wy);
    }
    P.window[(size_t) py * P.width + px] = out;
}

} // namespace qb
