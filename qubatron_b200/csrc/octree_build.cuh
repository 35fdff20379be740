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
