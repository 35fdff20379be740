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
// Parallel form (v2, sort based -- the first version walked all points down the tree level by level with one
// atomicMin per point and level, which spent 2.1 of its 3.4 ms on contended atomics):
//   1. key = the point's digits packed 3 bits each; one stable radix sort of (key, index); unique-by-key leaves the
//      distinct leaves in path order, each with its smallest point index;
//   2. a leaf whose key first differs from its predecessor's at digit t heads new nodes at depths t+1 .. levels: one
//      scan numbers every node of the tree (temporary ids), a binary search finds the leaf heading the parent of the
//      topmost new node;
//   3. depth by depth from the leaves up, every node takes the minimum of its children's creators (at most eight
//      contenders per address) and registers with its parent;
//   4. temporary nodes are sorted by (creator, level) -- one radix sort -- which yields the reference's indices;
//   5. children are renumbered and written straight into the traversal layout (octree_types.cuh: index +
//      child-exists mask nibbles, model array).
// No host round trip for the data: the paths may already live on the device.
#pragma once
#include "octree_types.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

namespace qb
{

constexpr int BUILD_LEVELS = 12; // digits per point: oct14 / oct54 / oct94 (skeleton_vsh.c L212-226)

// digits of one point packed most significant first (octree.c L153-156: levels 0-3 from the first buffer, 4-7 from
// the second, 8-11 from the third) above the point's index: one 64-bit word per point, so the sort moves keys only
// and, being stable on the path bits, leaves equal paths in index order.  36 path bits + 28 index bits.
constexpr int BUILD_INDEX_BITS = 28;

__global__ void build_key_kernel(const int4* __restrict__ p14, const int4* __restrict__ p54,
                                 const int4* __restrict__ p94, size_t n, int levels,
                                 unsigned long long* __restrict__ keys)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 a = p14[i], b = p54[i], c = p94[i];
    const int  d[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    unsigned long long k = 0;
#pragma unroll
    for (int l = 0; l < 12; l++)
        if (l < levels) k = (k << 3) | (unsigned long long) (d[l] & 7);
    keys[i] = (k << BUILD_INDEX_BITS) | (unsigned long long) i;
}

// sorted (path, index) words -> 1 for the first word of every distinct path
__global__ void build_head_flags_kernel(const unsigned long long* __restrict__ sorted, size_t n,
                                        unsigned char* __restrict__ flags)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = i == 0 || (sorted[i] >> BUILD_INDEX_BITS) != (sorted[i - 1] >> BUILD_INDEX_BITS);
}

// K: the distinct leaves in ascending path order, each word = path << 28 | smallest point index.  first_diff[k] = t: K[k] shares exactly t leading digits with K[k-1]
// (0 for the first leaf), so leaf k heads new nodes at depths t+1 .. levels (depth = digits in the prefix).
__global__ void build_leafinfo_kernel(const unsigned long long* __restrict__ K, int U, int levels,
                                      unsigned char* __restrict__ first_diff, int* __restrict__ created)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= U) return;
    int t = 0;
    if (k > 0)
        t = levels - 1 - (63 - __clzll((long long) ((K[k] >> BUILD_INDEX_BITS) ^ (K[k - 1] >> BUILD_INDEX_BITS)))) / 3;
    first_diff[k] = (unsigned char) t;
    created[k]    = levels - t;
}

// The parent of leaf k's topmost new node (depth t+1) is the depth-t node holding k; it is headed by the first leaf
// carrying k's t-digit prefix: a lower bound in the sorted keys.
__global__ void build_parent_kernel(const unsigned long long* __restrict__ K, int U, int levels,
                                    const unsigned char* __restrict__ first_diff, int* __restrict__ parent_leaf)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= U) return;
    const int t = first_diff[k];
    if (t == 0)
    {
        parent_leaf[k] = -1; // child of the root
        return;
    }
    const int                sh     = 3 * (levels - t) + BUILD_INDEX_BITS;
    const unsigned long long prefix = (K[k] >> sh) << sh; // index bits zero: below every word with this prefix
    int                      lo = 0, hi = k; // K[k] >= prefix, the answer is in [0, k]
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (K[mid] < prefix) lo = mid + 1;
        else hi = mid;
    }
    parent_leaf[k] = lo;
}

// temporary id of the node at depth d headed by leaf k (exists iff first_diff[k] < d)
__device__ __forceinline__ int build_tmp_id(const int* __restrict__ base, const unsigned char* __restrict__ first_diff,
                                            int k, int d)
{
    return 1 + base[k] + (d - (int) first_diff[k] - 1);
}

// One depth, leaves upwards: creator[id] already holds the minimum over the node's other children (previous launch);
// fold in the own chain, publish the creation key, register with the parent.
__global__ void build_link_kernel(const unsigned long long* __restrict__ K, int U, int levels, int d, const unsigned char* __restrict__ first_diff,
                                  const int* __restrict__ base, const int* __restrict__ parent_leaf,
                                  int* __restrict__ creator, int* __restrict__ tmp_child,
                                  unsigned* __restrict__ tmp_key)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= U) return;
    const int t = first_diff[k];
    if (t >= d) return;
    const int id = 1 + base[k] + (d - t - 1);
    const unsigned long long w = K[k];
    int cr = d == levels ? (int) (w & ((1ull << BUILD_INDEX_BITS) - 1)) : creator[id + 1]; // own node one depth down
    cr           = min(cr, creator[id]);
    creator[id]  = cr;
    tmp_key[id]  = ((unsigned) cr << 4) | (unsigned) (d - 1); // order of creation: (creator, level)
    const int digit = (int) ((w >> (3 * (levels - d) + BUILD_INDEX_BITS)) & 7ull);
    int       parent;
    if (d - 1 > t) parent = id - 1; // same chain; its creator is folded in by the next launch
    else if (d == 1) parent = 0;
    else
    {
        const int p = parent_leaf[k];
        parent      = build_tmp_id(base, first_diff, p, d - 1);
        atomicMin(&creator[parent], cr);
    }
    tmp_child[(size_t) parent * 8 + digit] = id;
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
    // `child` / `model` point at reference node 0 (= device node 1); child words hold DEVICE indices (octree_types.cuh)
    const int f = final_of_tmp[t];
    int       c[8];
    unsigned  m = 0;
#pragma unroll
    for (int s = 0; s < 8; s++)
    {
        const int k = tmp_child[(size_t) t * 8 + s];
        c[s]        = k ? final_of_tmp[k] + 1 : 0;
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
    // `child` / `model` point at reference node 0 (= device node 1); child words are device indices
    const int4 lo = child[2 * i], hi = child[2 * i + 1];
    int*       o  = out12 + i * 12;
    o[0]          = ref_node(lo.x & (int) CHILD_INDEX_MASK);
    o[1]          = ref_node(lo.y & (int) CHILD_INDEX_MASK);
    o[2]          = ref_node(lo.z);
    o[3]          = ref_node(lo.w);
    o[4]          = ref_node(hi.x);
    o[5]          = ref_node(hi.y);
    o[6]          = ref_node(hi.z);
    o[7]          = ref_node(hi.w);
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
