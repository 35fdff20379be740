// particle_sim.cuh -- one simulation step of the reference's debris particles and dust motes ("next" row SURVEY 8f #2).
//
// Behaviour to reproduce:
//   /root/reference/src/qubatron/shaders/particle_vsh.c main() L345-374: gravity on the speed, ONE static-tree trace
//     from the particle along its speed vector (cube_trace_line L121-343, the TRACE_PARTICLE flavour of
//     octree_trace_generic.cuh), stick to the hit leaf's corner when it is closer than 10 units, else move; park
//     (speed.x = -100000) when stuck or below y = -10.  Dispatched by particle_glc.c L118-156 as a transform-feedback
//     draw over the octree texture that octree_glc.c bound (L105-113) -- the coupling a CUDA connector breaks.
//   /root/reference/src/qubatron/shaders/dust_vsh.c main() L21-37: move, bounce away from the camera, wrap in a box.
// State (position, speed: float[3] each) ping-pongs between two device buffers; nothing returns to the host unless
// asked for.  Compiled like the renderer: no contraction, IEEE sqrt / divide, both division modes.
#pragma once
#include "octree_trace_fast.cuh"

namespace qb
{

// FAST: the base cube's grid is exact in fp32 -> the trace runs through the fast traversal (shared-memory stack,
// ordering table; trace_fast_single), except for particles with a zero speed component, whose mid-plane "hits" are
// the particle program's vec4(0.0) sentinel; launched with BLOCK_THREADS threads and the fast kernel's shared stack.
template <int DIV, bool FAST>
__global__ void __launch_bounds__(FAST ? BLOCK_THREADS : 256)
    particle_step_kernel(const FrameParams P, size_t n, const float* __restrict__ pos, const float* __restrict__ spd,
                         float* __restrict__ pos_out, float* __restrict__ spd_out, unsigned* __restrict__ finished)
{
    QB_DYN_SHARED(int, s_stack);
    __shared__ unsigned   s_compact_sel[16];
    FastSmem              fsm{0u, 0u};
    if (FAST) fsm = fast_single_smem(s_stack, s_compact_sel);
    size_t i    = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    bool   done = false;
    if (i < n)
    {
        const float3 p = make_float3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]);
        float3       s = make_float3(spd[i * 3], spd[i * 3 + 1], spd[i * 3 + 2]);
        float3       o = p;
        if (s.x > -90000.0f) // L352
        {
            s.y -= 0.4f;
            RayCounters cnt;
            TraceResult r;
            if (FAST && fast_single_ok(s))
                r = trace_fast_single<DIV, TRACE_PARTICLE>(P, p, s, fsm);
            else
                r = trace_generic<DIV, false, TRACE_PARTICLE>(P, p, s, cnt);
            bool stuck = false;
            if (r.status == 1 && r.iw > 0.0f) // L358
            {
                const float dx = r.tx - p.x, dy = r.ty - p.y, dz = r.tz - p.z;
                if (sqrtf(dx * dx + dy * dy + dz * dz) < 10.0f) // L362
                {
                    o     = make_float3(r.tx, r.ty, r.tz);
                    s.x   = -100000.0f;
                    stuck = true;
                }
            }
            if (!stuck)
            {
                o = make_float3(p.x + s.x, p.y + s.y, p.z + s.z); // L370
                if (o.y < -10.0f) s.x = -100000.0f;
            }
        }
        pos_out[i * 3] = o.x, pos_out[i * 3 + 1] = o.y, pos_out[i * 3 + 2] = o.z;
        spd_out[i * 3] = s.x, spd_out[i * 3 + 1] = s.y, spd_out[i * 3 + 2] = s.z;
        done = s.x < -900.0f; // the host's end-of-simulation test, modelutil.c L735
    }
    // particles parked after this step, one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, done);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(finished, (unsigned) __popc(m));
}

__global__ void dust_step_kernel(float3 campos, size_t n, const float* __restrict__ pos, const float* __restrict__ spd,
                                 float* __restrict__ pos_out, float* __restrict__ spd_out)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 s = make_float3(spd[i * 3], spd[i * 3 + 1], spd[i * 3 + 2]);
    float3       q = make_float3(pos[i * 3] + s.x, pos[i * 3 + 1] + s.y, pos[i * 3 + 2] + s.z);
    const float  dx = campos.x - q.x, dy = campos.y - q.y, dz = campos.z - q.z;
    if (sqrtf(dx * dx + dy * dy + dz * dz) < 100.0f) // L26
    {
        q.x += q.x - campos.x;
        q.y += q.y - campos.y;
        q.z += q.z - campos.z;
    }
    if (q.x < 400.0f) q.x = 800.0f; // L28-34
    if (q.y < 0.0f) q.y = 300.0f;
    if (q.z < 0.0f) q.z = 400.0f;
    if (q.x > 800.0f) q.x = 400.0f;
    if (q.y > 300.0f) q.y = 0.0f;
    if (q.z > 400.0f) q.z = 0.0f;
    pos_out[i * 3] = q.x, pos_out[i * 3 + 1] = q.y, pos_out[i * 3 + 2] = q.z;
    spd_out[i * 3] = s.x, spd_out[i * 3 + 1] = s.y, spd_out[i * 3 + 2] = s.z;
}

} // namespace qb
