// octree_render.cuh -- the per-pixel program: camera ray, primary trace, shadow
// ray from the light, shading, light disc, RGBA8 store.  One launch does the
// whole frame (or this rank's tiles of it); traces are a template policy so the
// generic and the register-stack traversal share everything else.
//
// Behaviour to reproduce: /root/reference/src/qubatron/shaders/octree_fsh.c
// main() L399-464 and octree_vsh.c L11 (coord = pixel centre).
// Compiled with -fmad=false (see octree_trace_generic.cuh).
#pragma once
#include "octree_trace_generic.cuh"

namespace qb
{

__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
    return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <int DIV>
__device__ __forceinline__ float3 normalize3(float3 a)
{
    const float l = sqrtf(dot3(a, a));
    if (DIV == DIV_GLSL)
    {
        const float inv = 1.0f / l;
        return make_float3(a.x * inv, a.y * inv, a.z * inv);
    }
    return make_float3(a.x / l, a.y / l, a.z / l);
}
// octree_fsh.c L392-395
__device__ __forceinline__ float3 quat_rotate(const float* q, float3 v)
{
    float3 qv = make_float3(q[0], q[1], q[2]);
    float3 c1 = cross3(qv, v);
    float3 t  = make_float3(c1.x + q[3] * v.x, c1.y + q[3] * v.y, c1.z + q[3] * v.z);
    float3 c2 = cross3(qv, t);
    return make_float3(v.x + 2.0f * c2.x, v.y + 2.0f * c2.y, v.z + 2.0f * c2.z);
}
__device__ __forceinline__ float max0(float a) { return a > 0.0f ? a : 0.0f; }
__device__ __forceinline__ unsigned int unorm8(float v)
{
    if (!(v > 0.0f)) return 0u;
    if (v > 1.0f) v = 1.0f;
    return (unsigned int) __float2int_rn(v * 255.0f);
}

template <int DIV>
struct GenericTracer
{
    static constexpr int div_mode = DIV;
    template <bool COUNT>
    static __device__ __forceinline__ TraceResult trace(const FrameParams& P, float3 pos, float3 dir, RayCounters& c)
    {
        return trace_generic<DIV, COUNT>(P, pos, dir, c);
    }
};

// Batched octree_trace_line ("next" row SURVEY 8f #4): the engine's CPU queries (collision probes qubatron.c
// L214-240, shooting L268-269, foot IK zombie.c L296/L310, ragdoll L477-486) against ONE tree, with the CPU
// twin's own quirks (octree.c L302-339 parallel-ray sentinel, L360-386 face tests) and IEEE division: the
// results equal the reference's compiled function bit for bit.  out_index = oct[8] of the leaf or 0;
// out_tlf (optional) = the leaf cube, left untouched on a miss like *otlf.
__global__ void trace_lines_kernel(const FrameParams P, size_t n, const float* __restrict__ pos,
                                   const float* __restrict__ dir, int* __restrict__ out_index,
                                   float* __restrict__ out_tlf)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayCounters       cnt;
    const TraceResult r = trace_generic<DIV_IEEE, false, TRACE_CPU>(P, make_float3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]),
                                                               make_float3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]), cnt);
    out_index[i]        = r.status == 1 ? r.model_s : 0;
    if (out_tlf && r.status == 1)
    {
        out_tlf[i * 4 + 0] = r.tx;
        out_tlf[i * 4 + 1] = r.ty;
        out_tlf[i * 4 + 2] = r.tz;
        out_tlf[i * 4 + 3] = r.tw;
    }
}

// ---------------------------------------------------------------------------
// multi-GPU completion fence (FenceDev, octree_types.cuh): device-side flags instead of a collective per frame
// ---------------------------------------------------------------------------
using ptx::ld_acquire_sys;
using ptx::st_release_sys;
constexpr long long FENCE_TIMEOUT_CYCLES = 8000000000ll; // ~4 s: a peer that never answers traps instead of hanging
__device__ __noinline__ void fence_spin(const unsigned* p, unsigned want)
{
    const long long t0 = clock64();
    while ((int) (ld_acquire_sys(p) - want) < 0)
    {
        __nanosleep(64);
        if (clock64() - t0 > FENCE_TIMEOUT_CYCLES)
        {
            printf("octree_cuc: multi-GPU fence timed out waiting for frame %u (have %u)\n", want, ld_acquire_sys(p));
            __trap();
        }
    }
}
// rank 0, first CTA of the frame's kernel: everything queued on rank 0's stream before this kernel has finished, so
// the previous frame has been consumed and the peers may overwrite it
__device__ __forceinline__ void fence_prologue(const FenceDev& F)
{
    if (F.peers && blockIdx.x == 0 && threadIdx.x > 0 && threadIdx.x < (unsigned) F.n)
        st_release_sys(F.peers[threadIdx.x] + FENCE_CONSUMED, F.seq - 1u);
}
// ranks != 0, before a warp's pixels go into rank 0's framebuffer (the traversal itself never waits)
__device__ __forceinline__ void fence_gate(const FenceDev& F)
{
    if (F.gate)
    {
        if ((threadIdx.x & 31) == 0) fence_spin(F.gate, F.seq - 1u);
        __syncwarp();
    }
}
// ranks != 0, on their stream right after the render kernel: publish the frame number to rank 0.  The kernel boundary
// orders the frame's peer stores before this kernel; it fences at system scope and release-stores the number.  (A
// first version did this from the last CTA of the render kernel, which takes a __threadfence_system per CTA after
// its NVLink stores: short CTAs -- the open-sky pose -- then cost more than their traversal and the pose stopped
// scaling at all; one 1-warp launch per frame costs ~2 us.)  A rank without a tile of the frame waits for the
// "consumed" signal itself, like its pixel stores would have.
__global__ void fence_signal_kernel(FenceDev F, int wait_gate)
{
    if (F.peers && threadIdx.x > 0 && threadIdx.x < (unsigned) F.n)
        st_release_sys(F.peers[threadIdx.x] + FENCE_CONSUMED, F.seq - 1u);
    if (F.done && threadIdx.x == 0)
    {
        if (wait_gate) fence_spin(F.gate, F.seq - 1u);
        __threadfence_system();
        st_release_sys(F.done, F.seq);
    }
}
// rank 0, after its own kernel: wait until every rank has published this frame
__global__ void fence_wait_done_kernel(const unsigned* local, int n, unsigned seq)
{
    if (threadIdx.x > 0 && threadIdx.x < (unsigned) n) fence_spin(local + threadIdx.x, seq);
}

// order[r] = the tile with the r-th largest cost (ties in tile order); clears the cost array of the next frame.
// One warp per tile, its lanes stride over the other tiles (n = 510 at 1080p: 16 independent loads per lane; the
// one-thread-per-tile loop this replaces took 11 us of every step, 1.6 % of a 1080p frame).
__global__ void tile_rank_kernel(const unsigned* __restrict__ cost, int n, int* __restrict__ order,
                                 unsigned* __restrict__ next_cost)
{
    const int i    = (int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = (int) (threadIdx.x & 31u);
    if (i >= n) return; // whole warps: blockDim is a multiple of 32
    const unsigned c = __ldg(cost + i);
    int            r = 0;
    for (int j = lane; j < n; j += 32)
    {
        const unsigned cj = __ldg(cost + j);
        r += (cj > c || (cj == c && j < i)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0)
    {
        order[r]     = i;
        next_cost[i] = 0;
    }
}
__global__ void tile_identity_kernel(int* __restrict__ order, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[i] = i;
}

// pixel of thread `tid` inside the CTA's 16x8 block: each warp covers an 8x4
// patch (lanes row-major inside it) so the 32 rays of a warp stay coherent
__device__ __forceinline__ void block_pixel(int tid, int& lx, int& ly)
{
    int warp = tid >> 5, lane = tid & 31;
    lx       = ((warp & 1) << 3) | (lane & 7);
    ly       = ((warp >> 1) << 2) | (lane >> 3);
}

template <class TRACER, bool AUX, bool COUNT>
__global__ void __launch_bounds__(BLOCK_THREADS) render_kernel(const FrameParams P)
{
    constexpr int DIV = TRACER::div_mode;
    // CTA -> (view, shard tile, block inside the tile)
    const int blocks_per_tile = P.blocks_per_tile_x * P.blocks_per_tile_y;
    int       b               = blockIdx.x;
    const int sub             = b % blocks_per_tile;
    b /= blocks_per_tile;
    const int tile_local = b % P.tiles_mine;
    const int view       = b / P.tiles_mine;
    const int tile       = P.rank + tile_local * P.world;
    const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
    const int by = sub / P.blocks_per_tile_x, bx = sub - by * P.blocks_per_tile_x;

    int lx, ly;
    block_pixel(threadIdx.x, lx, ly);
    const int px = tx * P.tile_w + bx * BLOCK_W + lx;
    const int py = ty * P.tile_h + by * BLOCK_H + ly;

    RayCounters cnt;
    if (COUNT)
    {
#pragma unroll
        for (int i = 0; i < CNT_COUNT; i++) cnt.v[i] = 0;
    }

    fence_prologue(P.fence);
    const bool active = px < P.W && py < P.H;
    if (active)
    {
        const ViewParams& V = P.views[view];

        const float3 camfp = make_float3(V.camfp[0], V.camfp[1], V.camfp[2]);
        const float3 light = make_float3(V.light[0], V.light[1], V.light[2]);

        // L402-404, L412-413
        float3 csv = make_float3(((float) px + 0.5f) * P.sx - V.cfp[0], ((float) py + 0.5f) * P.sy - V.cfp[1],
                                 0.0f - V.cfp[2]);
        csv        = quat_rotate(V.qz, csv);
        csv        = quat_rotate(V.qx, csv);

        // L417-418; acos(d) < 0.02 is evaluated as d >= disc_dot_min (host libm threshold)
        const float3 csv_n  = normalize3<DIV>(csv);
        const float  camdot = dot3(make_float3(V.camlight_n[0], V.camlight_n[1], V.camlight_n[2]), csv_n);
        const bool   disc   = camdot >= V.disc_dot_min && camdot <= 1.0f;

        int   flags   = 0;
        bool  discard = false;
        float cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
        int   a0 = -1, a1 = -1, a2 = -1, a3 = -1, a4 = -1, a5 = -1;

        if (COUNT) cnt.v[CNT_RAYS_PRIMARY]++;
        TraceResult res = TRACER::template trace<COUNT>(P, camfp, csv, cnt);
        if (res.status < 0) discard = true;

        if (!discard && res.status == 1)
        {
            flags |= 2;
            a0 = res.model_s;
            a1 = res.model_d;
            a2 = res.node_s;
            a3 = res.node_d;

            // L226-241: static point `model_s` (point 0 when the leaf is dynamic
            // only), overridden by the dynamic point iff model_d > 0
            const PointsDev& pts = (res.model_d > 0) ? P.pts_d : P.pts_s;
            const int        idx = (res.model_d > 0) ? res.model_d : res.model_s;
            float4           col = make_float4(0.f, 0.f, 0.f, 1.f), nrm = make_float4(0.f, 0.f, 0.f, 1.f);
            if ((unsigned) idx < (unsigned) pts.points)
            {
                col = __ldg(pts.rec + 2 * (size_t) idx);
                nrm = __ldg(pts.rec + 2 * (size_t) idx + 1);
            }
            cr = col.x;
            cg = col.y;
            cb = col.z;
            ca = 1.0f;

            if (res.iw > 0.0f) // L424-450
            {
                flags |= 4;
                if (COUNT)
                {
                    cnt.v[CNT_HITS]++;
                    cnt.v[CNT_RAYS_SHADOW]++;
                }
                const float3 lghtv = make_float3(res.ix - light.x, res.iy - light.y, res.iz - light.z);
                TraceResult  lc    = TRACER::template trace<COUNT>(P, light, lghtv, cnt);
                if (lc.status < 0)
                    discard = true;
                else
                {
                    if (lc.status == 1)
                    {
                        a4 = lc.node_s;
                        a5 = lc.node_d;
                    }
                    const float dx = lc.ix - res.ix, dy = lc.iy - res.iy, dz = lc.iz - res.iz;
                    const float sqr = dx * dx + dy * dy + dz * dz;

                    const float3 nn  = normalize3<DIV>(make_float3(nrm.x, nrm.y, nrm.z));
                    const float3 nl  = normalize3<DIV>(make_float3(-lghtv.x, -lghtv.y, -lghtv.z));
                    const float3 nc  = normalize3<DIV>(make_float3(-csv.x, -csv.y, -csv.z));
                    const float  lna = max0(dot3(nl, nn));
                    const float  cna = max0(dot3(nc, nn));
                    const float  vis = (15.0f < sqr) ? 0.0f : 1.0f; // step(sqr, 15.0)
                    if (vis != 0.0f) flags |= 8;

                    const float f = 0.1f + 0.2f * cna + lna * vis * 0.7f;
                    cr            = cr * f;
                    cg            = cg * f;
                    cb            = cb * f;
                    cb *= 0.7f;
                    const float g = (float) V.shoot * cna * 0.1f;
                    cr += g;
                    cg += g;
                    cb += g;
                }
            }
        }

        if (!discard && disc) // L452-459
        {
            flags |= 16;
            if (COUNT) cnt.v[CNT_RAYS_DISC]++;
            const float3 lghtv = make_float3(light.x - camfp.x, light.y - camfp.y, light.z - camfp.z);
            TraceResult  lc    = TRACER::template trace<COUNT>(P, camfp, lghtv, cnt);
            if (lc.status < 0)
                discard = true;
            else
            {
                const float resvx = lc.ix - camfp.x;
                if (qdiv<DIV>(resvx, lghtv.x) > 1.0f)
                {
                    flags |= 32;
                    cr = cg = cb = ca = 1.0f;
                }
            }
        }

        if (discard)
        {
            flags = 1;
            a0 = a1 = a2 = a3 = a4 = a5 = -1;
            cr = cg = cb = ca = 0.0f;
            if (COUNT) cnt.v[CNT_DISCARDS]++;
        }

        if (P.fence.gate) fence_spin(P.fence.gate, P.fence.seq - 1u);
        const size_t p = (size_t) view * P.view_stride + (size_t) py * P.pitch + px;
        P.frame[p]     = make_uchar4((unsigned char) unorm8(cr), (unsigned char) unorm8(cg), (unsigned char) unorm8(cb),
                                     (unsigned char) unorm8(ca));
        if (AUX)
        {
            const size_t q = (size_t) view * P.W * P.H + (size_t) py * P.W + px;
            P.flags[q]     = (uint8_t) flags;
            int* a         = P.aux + q * 6;
            a[0]           = a0;
            a[1]           = a1;
            a[2]           = a2;
            a[3]           = a3;
            a[4]           = a4;
            a[5]           = a5;
        }
    }

    if (COUNT)
    {
#pragma unroll
        for (int i = 0; i < CNT_COUNT; i++)
        {
            unsigned int v = cnt.v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(P.counters + i, (unsigned long long) v);
        }
    }
}

} // namespace qb
