/*
 * particle_vsh_oracle.c -- CPU restatement of the reference's particle and dust simulation vertex programs.
 *
 * TEST INFRASTRUCTURE ONLY (see qb_oracle.h).  Follows, statement by statement,
 *   /root/reference/src/qubatron/shaders/particle_vsh.c
 *     L67-104   is_cube_{x,y,z}plane: a ray parallel to the plane yields vec4(0.0) (NOT the fragment shader's
 *               FLT_MAX sentinel), whose xyz can pass a range test
 *     L108-119  oct_from_octets_for_index: texel x = (3i mod 8192) + octi/4 WITHOUT wrapping to the next row (the
 *               fragment shader wraps, octree_fsh.c L130-135); a fetch past column 8191 is outside the texture
 *     L121-343  cube_trace_line over the STATIC tree only; a ray that misses the base cube returns a zero result
 *               instead of discarding
 *     L345-374  main: gravity, trace along the speed vector, stick to the hit leaf when it is closer than 10 units,
 *               else move; park below y = -10
 *   /root/reference/src/qubatron/shaders/dust_vsh.c L21-37 (main; random() L16-19 is never called)
 * with the uniforms of particle_glc.c L118-133 / dust_glc.c L103-118.
 *
 * "next" row SURVEY 8f #2.  Parity status: PINNED -- oracle/glsl_ref.c modes 30 / 31 run both programs unmodified
 * through transform feedback on Mesa llvmpipe; tests/golden/particles_*.npz / dust_*.npz hold their outputs and
 * tests/test_golden.py compares this restatement with them bit for bit.
 * Out-of-texture fetch: undefined in GLSL ES 3.00 (texelFetch, section 8.8); llvmpipe returns 0, which is also what
 * robust buffer access mandates; restated as 0 and covered by the golden fixtures (scenes with > 2730 nodes).
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

typedef struct p4
{
    float x, y, z, w;
} p4;

typedef struct ptree
{
    const int32_t* oct; /* 12 ints per node */
    int64_t        nodes;
} ptree;

/* L67-104 */
static p4 plane_x(float x, const float* lp, const float* lv)
{
    p4 r = {0.0f, 0.0f, 0.0f, 0.0f};
    if (lv[0] != 0.0f)
    {
        r.w = QB_DIV(x - lp[0], lv[0]);
        r.y = lp[1] + lv[1] * r.w;
        r.z = lp[2] + lv[2] * r.w;
        r.x = x;
    }
    return r;
}
static p4 plane_y(float y, const float* lp, const float* lv)
{
    p4 r = {0.0f, 0.0f, 0.0f, 0.0f};
    if (lv[1] != 0.0f)
    {
        r.w = QB_DIV(y - lp[1], lv[1]);
        r.x = lp[0] + lv[0] * r.w;
        r.z = lp[2] + lv[2] * r.w;
        r.y = y;
    }
    return r;
}
static p4 plane_z(float z, const float* lp, const float* lv)
{
    p4 r = {0.0f, 0.0f, 0.0f, 0.0f};
    if (lv[2] != 0.0f)
    {
        r.w = QB_DIV(z - lp[2], lv[2]);
        r.x = lp[0] + lv[0] * r.w;
        r.y = lp[1] + lv[1] * r.w;
        r.z = z;
    }
    return r;
}

/* L108-119 */
static int oct_lookup(const ptree* t, int octi, int i, int level)
{
    if (i == 0 && level > 0) return 0;
    int64_t e  = (int64_t) i * 3;
    int64_t cy = e / 8192;
    int64_t cx = e - cy * 8192 + octi / 4;
    if (cx >= 8192) return 0;                      /* outside the texture: no row wrap in this program */
    int64_t texel = cy * 8192 + cx;
    if (texel >= t->nodes * 3 || texel < 0) return 0; /* zero padding behind the uploaded nodes */
    return t->oct[texel * 4 + (octi - (octi / 4) * 4)];
}

typedef struct plevel
{
    p4  cube;
    p4  isps[4];
    int octs[4];
    int ispsi, socti;
} plevel;

typedef struct pres
{
    p4 isp, tlf;
} pres;

static const float xsft[8] = {0, 1, 0, 1, 0, 1, 0, 1}, ysft[8] = {0, 0, 1, 1, 0, 0, 1, 1}, zsft[8] = {0, 0, 0, 0, 1, 1, 1, 1};
static const int   horpairs[8] = {1, 0, 3, 2, 5, 4, 7, 6}, verpairs[8] = {2, 3, 0, 1, 6, 7, 4, 5},
                 deppairs[8] = {4, 5, 6, 7, 0, 1, 2, 3};

/* L121-343 */
static pres particle_trace(const ptree* t, const float basecube[4], int maxlevel, const float* pos, const float* dir)
{
    pres res;
    memset(&res, 0, sizeof(res));
    plevel stck[18];
    int    level = 0;
    stck[0].cube  = (p4){basecube[0], basecube[1], basecube[2], basecube[3]};
    stck[0].socti = 0;
    stck[0].ispsi = 0;

    p4 act;
    p4 tlf = stck[0].cube;
    p4 brb = {tlf.x + tlf.w, tlf.y - tlf.w, tlf.z - tlf.w, 0.0f};
    int hitc = 0;
    p4  hitp[8];

    act = plane_z(tlf.z, pos, dir); /* front */
    if (tlf.x < act.x && act.x <= brb.x && tlf.y > act.y && act.y >= brb.y) hitp[hitc++] = act;
    act = plane_z(brb.z, pos, dir); /* back */
    if (tlf.x < act.x && act.x <= brb.x && tlf.y > act.y && act.y >= brb.y) hitp[hitc++] = act;
    act = plane_x(tlf.x, pos, dir); /* left */
    if (tlf.y > act.y && act.y >= brb.y && tlf.z > act.z && act.z >= brb.z) hitp[hitc++] = act;
    act = plane_x(brb.x, pos, dir); /* right */
    if (tlf.y > act.y && act.y >= brb.y && tlf.z > act.z && act.z >= brb.z) hitp[hitc++] = act;
    act = plane_y(tlf.y, pos, dir); /* top */
    if (tlf.x < act.x && act.x <= brb.x && tlf.z > act.z && act.z >= brb.z) hitp[hitc++] = act;
    act = plane_y(brb.y, pos, dir); /* bottom */
    if (tlf.x < act.x && act.x <= brb.x && tlf.z > act.z && act.z >= brb.z) hitp[hitc++] = act;

    if (hitc < 2) return res;                               /* L180 */
    if (hitp[0].w < 0.0f && hitp[1].w < 0.0f) return res;   /* L183 */
    if (hitp[1].w < hitp[0].w) hitp[0] = hitp[1];           /* L190 */
    if (hitp[0].w < 0.0f) hitp[0] = (p4){pos[0], pos[1], pos[2], 0.0f}; /* L193 */
    stck[level].isps[0] = hitp[0];

    for (;;)
    {
        tlf = stck[level].cube;
        if (level == maxlevel) /* L203-219; colour and normal are fetched but never used by main() */
        {
            res.isp = stck[level].isps[0];
            res.tlf = tlf;
            return res;
        }
        if (stck[level].ispsi == 0) /* L222-300 */
        {
            stck[level].ispsi = 128;
            p4 b   = {tlf.x + tlf.w, tlf.y - tlf.w, tlf.z - tlf.w, 0.0f};
            p4 hlf = {b.x + (tlf.x - b.x) * 0.5f, b.y + (tlf.y - b.y) * 0.5f, b.z + (tlf.z - b.z) * 0.5f, 0.0f};
            hitc    = 1;
            hitp[0] = stck[level].isps[0];
            act = plane_z(hlf.z, pos, dir);
            if (act.w > 0.0f && tlf.x < act.x && act.x <= b.x && tlf.y > act.y && act.y >= b.y) hitp[hitc++] = act;
            act = plane_x(hlf.x, pos, dir);
            if (act.w > 0.0f && tlf.y > act.y && act.y >= b.y && tlf.z > act.z && act.z >= b.z) hitp[hitc++] = act;
            act = plane_y(hlf.y, pos, dir);
            if (act.w > 0.0f && tlf.x < act.x && act.x <= b.x && tlf.z > act.z && act.z >= b.z) hitp[hitc++] = act;

            int oct = 0, pre = -1;
            for (int i = 0; i < hitc; ++i)
            {
                if (i < hitc - 1)
                    for (int j = i + 1; j < hitc; ++j)
                        if (hitp[j].w < hitp[i].w)
                        {
                            act     = hitp[i];
                            hitp[i] = hitp[j];
                            hitp[j] = act;
                        }
                act = hitp[i];
                oct = 0;
                if (act.x > hlf.x) oct = 1;
                if (act.y < hlf.y) oct += 2;
                if (act.z < hlf.z) oct += 4;
                if (oct == pre)
                {
                    if (act.x == hlf.x) oct = horpairs[oct];
                    else if (act.y == hlf.y) oct = verpairs[oct];
                    else if (act.z == hlf.z) oct = deppairs[oct];
                }
                pre = oct;
                int socti = oct_lookup(t, oct, stck[level].socti, level);
                if (socti > 0)
                {
                    int ind = stck[level].ispsi, len = ind & 0x0F;
                    stck[level].octs[len] = oct;
                    stck[level].isps[len] = act;
                    len++;
                    stck[level].ispsi = (ind & 0xF0) | len;
                }
            }
        }
        int cur_len = stck[level].ispsi & 0x0F;
        if (cur_len > 0) /* L306-333 */
        {
            int   nxt_ind = (stck[level].ispsi >> 4) & 7;
            p4    nxt_isp = stck[level].isps[nxt_ind];
            int   nxt_oct = stck[level].octs[nxt_ind];
            float halfs   = QB_DIV(tlf.w, 2.0f);
            tlf.x += xsft[nxt_oct] * halfs;
            tlf.y -= ysft[nxt_oct] * halfs;
            tlf.z -= zsft[nxt_oct] * halfs;
            tlf.w = halfs;
            nxt_ind++;
            cur_len--;
            stck[level].ispsi = 128 | (nxt_ind << 4) | cur_len;
            int socti = oct_lookup(t, nxt_oct, stck[level].socti, level);
            level += 1;
            stck[level].cube    = tlf;
            stck[level].ispsi   = 0;
            stck[level].socti   = socti;
            stck[level].isps[0] = nxt_isp;
        }
        else /* L334-341 */
        {
            stck[level--].ispsi = 0;
            if (level < 0) return res;
        }
    }
}

/* particle_vsh.c main() L345-374 for n particles: pos / spd float[3n] -> pos_out / spd_out; hit_out (optional,
 * not an output of the shader): 1 where the particle stuck to a leaf in this step */
void qb_oracle_particles(const int32_t* oct_s, int64_t nodes_s, const float basecube[4], int maxlevel, int64_t n,
                         const float* pos, const float* spd, float* pos_out, float* spd_out, int32_t* hit_out)
{
    ptree t = {oct_s, nodes_s};
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++)
    {
        const float* p = pos + i * 3;
        float        s[3] = {spd[i * 3], spd[i * 3 + 1], spd[i * 3 + 2]};
        float*       po = pos_out + i * 3;
        float*       so = spd_out + i * 3;
        po[0] = p[0], po[1] = p[1], po[2] = p[2];
        int stuck = 0;
        if (s[0] > -90000.0f)
        {
            s[1] -= 0.4f;
            pres res = particle_trace(&t, basecube, maxlevel, p, s);
            if (res.isp.w > 0.0f)
            {
                float dx = res.tlf.x - p[0], dy = res.tlf.y - p[1], dz = res.tlf.z - p[2];
                if (sqrtf(dx * dx + dy * dy + dz * dz) < 10.0f)
                {
                    po[0] = res.tlf.x, po[1] = res.tlf.y, po[2] = res.tlf.z;
                    s[0]  = -100000.0f;
                    stuck = 1;
                }
            }
            if (!stuck)
            {
                po[0] = p[0] + s[0], po[1] = p[1] + s[1], po[2] = p[2] + s[2];
                if (po[1] < -10.0f) s[0] = -100000.0f;
            }
        }
        so[0] = s[0], so[1] = s[1], so[2] = s[2];
        if (hit_out) hit_out[i] = stuck;
    }
}

/* dust_vsh.c main() L21-37 */
void qb_oracle_dust(const float campos[3], int64_t n, const float* pos, const float* spd, float* pos_out,
                    float* spd_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        float np[3] = {pos[i * 3] + spd[i * 3], pos[i * 3 + 1] + spd[i * 3 + 1], pos[i * 3 + 2] + spd[i * 3 + 2]};
        float dx = campos[0] - np[0], dy = campos[1] - np[1], dz = campos[2] - np[2];
        if (sqrtf(dx * dx + dy * dy + dz * dz) < 100.0f)
        {
            np[0] += np[0] - campos[0];
            np[1] += np[1] - campos[1];
            np[2] += np[2] - campos[2];
        }
        if (np[0] < 400.0f) np[0] = 800.0f;
        if (np[1] < 0.0f) np[1] = 300.0f;
        if (np[2] < 0.0f) np[2] = 400.0f;
        if (np[0] > 800.0f) np[0] = 400.0f;
        if (np[1] > 300.0f) np[1] = 0.0f;
        if (np[2] > 400.0f) np[2] = 0.0f;
        pos_out[i * 3] = np[0], pos_out[i * 3 + 1] = np[1], pos_out[i * 3 + 2] = np[2];
        spd_out[i * 3] = spd[i * 3], spd_out[i * 3 + 1] = spd[i * 3 + 1], spd_out[i * 3 + 2] = spd[i * 3 + 2];
    }
}
