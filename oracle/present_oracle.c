/*
 * present_oracle.c -- CPU restatement of the reference's presentation pass.
 *
 * TEST INFRASTRUCTURE ONLY (see qb_oracle.h).  Follows /root/reference/src/qubatron/octree_glc.c L308-351: the
 * 2048 x 2048 RGBA8 render target (cleared to (0,0,0,0) every frame, the frame in its lower-left (int)ow x (int)oh
 * texels, L254-260 + L288) is drawn as a textured quad (shaders/texquad_vsh.c, texquad_fsh.c; ortho(0,ow,0,oh);
 * quad 0..2048 with texcoords 0..1; MIN/MAG filter LINEAR, L236-237; wrap left at the GL default REPEAT; blending
 * enabled with the default ONE/ZERO function = a plain store of all four channels) into a width x height window,
 * then a 2 x 2 white crosshair is cleared at (width/2 - 1, height/2 - 1) (L347-351).
 *
 * "next" row SURVEY 8f #4 (second half).  GL leaves the precision of LINEAR filtering to the implementation.  This
 * restatement is pinned to what Mesa llvmpipe does for RGBA8 (found by fitting, then exact on every pixel of the
 * golden windows, tests/golden/present_*.npz, glsl_ref mode 40): the texel coordinate is taken in 8.8 fixed point,
 * k = round(X * 256) - 128 with X = (px + 0.5) * ow / width; texels k >> 8 and (k >> 8) + 1 (mod 2048); weight
 * w = k & 255; lerp(a, b, w) = a + (((b - a) * w) >> 8) per 8-bit channel, along x for both rows, then along y.
 * NVIDIA hardware also filters RGBA8 with 8-bit weights; other drivers may differ by 1-2 / 255 at non-integer scales.
 * At quality 10 (1:1, the reference's default) every weight is 0 and the pass is a copy.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline int lerp8(int a, int b, int w) { return a + (((b - a) * w) >> 8); }

/* frame: RGBA8 [vp_h][vp_w] (row 0 = bottom); window: RGBA8 [height][width] (row 0 = bottom) */
void qb_oracle_present(const uint8_t* frame, int vp_w, int vp_h, float ow, float oh, int width, int height,
                       uint8_t* window)
{
    const double sx = (double) ow / (double) width, sy = (double) oh / (double) height;
#pragma omp parallel for schedule(static)
    for (int py = 0; py < height; py++)
    {
        const long ky = lrint(((double) py + 0.5) * sy * 256.0) - 128;
        const int  y0 = (int) (((ky >> 8) % 2048 + 2048) % 2048), y1 = (y0 + 1) % 2048, wy = (int) (ky & 255);
        for (int px = 0; px < width; px++)
        {
            const long kx = lrint(((double) px + 0.5) * sx * 256.0) - 128;
            const int  x0 = (int) (((kx >> 8) % 2048 + 2048) % 2048), x1 = (x0 + 1) % 2048, wx = (int) (kx & 255);
            for (int c = 0; c < 4; c++)
            {
#define QB_TEXEL(X, Y) (((X) < vp_w && (Y) < vp_h) ? (int) frame[((size_t) (Y) * vp_w + (X)) * 4 + c] : 0)
                const int top = lerp8(QB_TEXEL(x0, y0), QB_TEXEL(x1, y0), wx);
                const int bot = lerp8(QB_TEXEL(x0, y1), QB_TEXEL(x1, y1), wx);
#undef QB_TEXEL
                window[((size_t) py * width + px) * 4 + c] = (uint8_t) lerp8(top, bot, wy);
            }
        }
    }
    for (int y = height / 2 - 1; y < height / 2 + 1; y++) /* glScissor(width/2 - 1, height/2 - 1, 2, 2) */
        for (int x = width / 2 - 1; x < width / 2 + 1; x++)
            if (x >= 0 && y >= 0 && x < width && y < height)
                for (int c = 0; c < 4; c++) window[((size_t) y * width + x) * 4 + c] = 255;
}
