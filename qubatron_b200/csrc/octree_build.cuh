// octree_build.cuh -- GPU build of an octree from per-point octant paths, with the
// reference's node numbering.  ("Next" row SURVEY 8f #1, second half: the engine
// rebuilds the dynamic tree on the CPU every frame with octree_reset +
// octree_insert_path, /root/reference/src/qubatron/qubatron.c L439-452,
// octree.c L149-180.)
//
// Sequential semantics to reproduce: points are inserted in index order; walking
// the 12 digits of a point, every missing child is created with the next free
// node index and oct[8] = the point's model index.  Hence a node exists per
// distinct path prefix, its creator is the smallest point index having that
// prefix, and node indices follow the order (creator, level).
//
// Parallel form:
//   1. level by level, every point proposes itself for the slot (its node at this
//      level, its digit) with atomicMin; occupied slots become the next level's
//      nodes under temporary ids (scan), and every point steps into its child;
//   2. temporary nodes are sorted by (creator, level) -- one radix sort -- which
//      yields the reference's indices;
//   3. children are renumbered and written straight into the traversal layout
//      (octree_types.cuh: index + child-exists mask nibbles, model array).
// No host round trip: the paths may already live on the device.
#pragma once
#include "octree_types.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace qb
{

constexpr int BUILD_LEVELS = 12; // digits per point: oct14 / oct54 / oct94 (skeleton_vsh.c L212-226)

__device__ __forceinline__ int path_digit(const int* __restrict__ p14, const int* __restrict__ p54,
                                          const int* __restrict__ p94, size_t i, int level)
{
    // octree.c L153-156: levels 0-3 from the first buffer, 4-7 from the second, 8-11 from the third
    const int* p = level < 4 ? p14 : (level < 8 ? p54 : p94);
    return p[i * 4 + (level & 3)] & 7;
}

__global__ void build_propose_kernel(const int* __restrict__ p14, const int* __restrict__ p54,
                                     const int* __restrict__ p94, size_t n, int level, const int* __restrict__ cur,
                                     int level_base, int* __restrict__ table)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    // slot this point proposes itself for; points beyond n propose nothing
    long long slot = -1;
    if (i < n) slot = (long long) (cur[i] - level_base) * 8 + path_digit(p14, p54, p94, i, level);
    // Neighbouring points usually share the slot (the model is spatially sorted): within a warp indices
    // ascend, so only the first lane of a run of equal slots can hold the minimum -- one atomic per run
    // instead of one per point (the top levels would otherwise serialise 10 M atomics on 8 addresses).
    const long long prev = __shfl_up_sync(0xffffffffu, slot, 1);
    const bool      head = (threadIdx.x & 31) == 0 || prev != slot;
    if (slot >= 0 && head) atomicMin(&table[slot], (int) i);
}

__global__ void build_flags_kernel(const int* __restrict__ table, size_t slots, int* __restrict__ flags)
{
    size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (j < slots) flags[j] = table[j] != 0x7fffffff;
}

__global__ void build_create_kernel(const int* __restrict__ table, const int* __restrict__ pos, size_t slots,
                                    int level, int level_base, int next, int* __restrict__ tmp_child,
                                    unsigned* __restrict__ tmp_key)
{
    size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (j >= slots) return;
    const int creator = table[j];
    if (creator == 0x7fffffff) return;
    const int id                                             = next + pos[j];
    tmp_child[(size_t) (level_base + (int) (j >> 3)) * 8 + (j & 7)] = id;
    tmp_key[id] = ((unsigned) creator << 4) | (unsigned) level; // order of creation: (creator, level)
}

__global__ void build_step_kernel(const int* __restrict__ p14, const int* __restrict__ p54,
                                  const int* __restrict__ p94, size_t n, int level, int* __restrict__ cur,
                                  const int* __restrict__ tmp_child)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    cur[i] = tmp_child[(size_t) cur[i] * 8 + path_digit(p14, p54, p94, i, level)];
}

__global__ void build_fill_kernel(int* __restrict__ v, size_t n, int value)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i < n) v[i] = value;
}

__global__ void build_iota_kernel(int* __restrict__ v, int n, int first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = first + i;
}

// sorted_ids[r] = temporary id of the node with final index r + 1
__global__ void build_rank_kernel(const int* __restrict__ sorted_ids, int m, int* __restrict__ final_of_tmp)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) final_of_tmp[sorted_ids[r]] = r + 1;
    if (r == 0) final_of_tmp[0] = 0; // the root keeps index 0
}

__global__ void build_emit_kernel(const int* __restrict__ tmp_child, const unsigned* __restrict__ tmp_key,
                                  const int* __restrict__ final_of_tmp, int total, int first_modind,
                                  int4* __restrict__ child, int* __restrict__ model)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int f = final_of_tmp[t];
    int       c[8];
    unsigned  m = 0;
#pragma unroll
    for (int s = 0; s < 8; s++)
    {
        const int k = tmp_child[(size_t) t * 8 + s];
        c[s]        = k ? final_of_tmp[k] : 0;
        m |= (k ? 1u : 0u) << s;
    }
    child[2 * (size_t) f] = make_int4(c[0] | (int) ((m & 15u) << CHILD_MASK_SHIFT), c[1] | (int) ((m >> 4) << CHILD_MASK_SHIFT),
                                      c[2], c[3]);
    child[2 * (size_t) f + 1] = make_int4(c[4], c[5], c[6], c[7]);
    model[f]                  = t == 0 ? 0 : first_modind + (int) (tmp_key[t] >> 4); // octree.c L163: oct[8] = modind
}

// device layout -> the reference's 12-int nodes (octree.c L11-14), for parity checks
__global__ void export_nodes_kernel(const int4* __restrict__ child, const int* __restrict__ model, size_t nodes,
                                    int* __restrict__ out12)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= nodes) return;
    const int4 lo = child[2 * i], hi = child[2 * i + 1];
    int*       o  = out12 + i * 12;
    o[0]          = lo.x & (int) CHILD_INDEX_MASK;
    o[1]          = lo.y & (int) CHILD_INDEX_MASK;
    o[2]          = lo.z;
    o[3]          = lo.w;
    o[4]          = hi.x;
    o[5]          = hi.y;
    o[6]          = hi.z;
    o[7]          = hi.w;
    o[8]          = model[i];
    o[9] = o[10] = o[11] = 0;
}

} // namespace qb

// ---------------------------------------------------------------------------
// GPU voxeliser ("next" row SURVEY 8f #3): what the reference's offline tool qmc does
// (/root/reference/src/qubatron/qmc.c L62-83 grid index and drop, L130-165 + L259
// x-major sort, L291-327 first point of every occupied cell) followed by the bulk
// tree build from the survivors' octant digits (octree_insert_point, octree.c L95-147).
// ---------------------------------------------------------------------------
namespace qb
{

// sort key = (xi, yi, zi) packed x-major, 16 bits per axis; points outside the cube get the all-ones key
__global__ void voxel_key_kernel(const float* __restrict__ pos, size_t n, float precision, int division,
                                 unsigned long long* __restrict__ keys, unsigned* __restrict__ idx)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    // qmc.c L62-68: floor(p / precision), evaluated in fp32
    const float xi = floorf(pos[i * 3 + 0] / precision);
    const float yi = floorf(pos[i * 3 + 1] / precision);
    const float zi = floorf(pos[i * 3 + 2] / precision);
    const float dv = (float) division;
    const bool  in = xi >= 0.0f && xi < dv && yi >= 0.0f && yi < dv && zi >= 0.0f && zi < dv;
    keys[i] = in ? ((unsigned long long) xi << 32) | ((unsigned long long) yi << 16) | (unsigned long long) zi
                 : ~0ull;
    idx[i]  = (unsigned) i;
}

// first point of every distinct, valid key (the input order inside a cell is kept by the stable sort)
__global__ void voxel_flag_kernel(const unsigned long long* __restrict__ keys, size_t n, int* __restrict__ flags)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    flags[i]                   = k != ~0ull && (i == 0 || keys[i - 1] != k);
}

// survivors: point record (colour = uchar / 255.0 in double, qmc.c L52-54; normal), source index, and the
// octant digits of octree_insert_point (size halves per level; (int)(p / size) % 2, +2 / +4 for the LOWER halves)
__global__ void voxel_emit_kernel(const float* __restrict__ pos, const unsigned char* __restrict__ col,
                                  const float* __restrict__ nrm, const unsigned* __restrict__ sorted_idx,
                                  const int* __restrict__ flags, const int* __restrict__ slot, size_t n,
                                  float basesize, int levels, float* __restrict__ rec, float* __restrict__ pos_out,
                                  long long* __restrict__ order, int* __restrict__ p14, int* __restrict__ p54,
                                  int* __restrict__ p94)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const size_t src = sorted_idx[i];
    const size_t j   = (size_t) slot[i];
    const float  px = pos[src * 3 + 0], py = pos[src * 3 + 1], pz = pos[src * 3 + 2];
    order[j]           = (long long) src;
    pos_out[j * 3 + 0] = px;
    pos_out[j * 3 + 1] = py;
    pos_out[j * 3 + 2] = pz;
    float* r           = rec + j * 8;
    r[0]               = (float) ((double) col[src * 3 + 0] / 255.0);
    r[1]               = (float) ((double) col[src * 3 + 1] / 255.0);
    r[2]               = (float) ((double) col[src * 3 + 2] / 255.0);
    r[3]               = 1.0f;
    r[4]               = nrm[src * 3 + 0];
    r[5]               = nrm[src * 3 + 1];
    r[6]               = nrm[src * 3 + 2];
    r[7]               = 0.0f;
    float size = basesize;
    int   d[12];
#pragma unroll
    for (int l = 0; l < 12; l++)
    {
        d[l] = 0;
        if (l < levels)
        {
            size    = (float) ((double) size / 2.0); // octree.c L102
            int o   = ((int) (px / size)) % 2;
            int yi  = ((int) (py / size)) % 2;
            int zi  = ((int) (pz / size)) % 2;
            if (yi == 0) o += 2;
            if (zi == 0) o += 4;
            d[l] = o;
        }
    }
    reinterpret_cast<int4*>(p14)[j] = make_int4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<int4*>(p54)[j] = make_int4(d[4], d[5], d[6], d[7]);
    reinterpret_cast<int4*>(p94)[j] = make_int4(d[8], d[9], d[10], d[11]);
}

} // namespace qb
