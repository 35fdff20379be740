// octree_types.cuh -- device-side data layout and per-frame parameter block.
//
// HBM layout (see DESIGN.md "Data layout"):
//   child_{s,d} : int4[2*(nodes+2)] -- device node 0 = all-zero dummy, reference node n
//                                   at device node n + 1 (see TreeDev below);
//                                   the 8 child indices of a node, 32 B, one
//                                   aligned sector per node expansion
//                                   (reference node = int32[12], octree.c L11-14).
//                                   Bits 0-27 of a word = child DEVICE index
//                                   (0 = absent); the top 4 bits of words 0 and 1
//                                   hold the node's 8-bit "child exists" mask
//                                   (bit k of the mask = child k > 0), maintained by
//                                   the upload kernels, so an expansion needs 8 bytes
//                                   of the node and a descent one more word.
//   slot_{s,d}  : uint2[8*(nodes+2)] -- derived from child_*: record 8 n + k = { device index of child k of node n,
//                                   child-exists mask of THAT CHILD, in every byte of the word (SLOT_MASK_REP) }.  A descent of the fast traversal reads one
//                                   record and has the child's node index and everything the child's expansion
//                                   needs; the expansion itself loads nothing (the connector keeps the records
//                                   current through every upload, build and blob: octree_cuc.cu derive_slots)
//   model_{s,d} : int32[nodes]   -- oct[8] of the node (point index), split out
//                                   as the reference author wanted
//                                   (octree_fsh.c L125)
//   pts_{s,d}   : float4[2*pts]  -- {colour rgb, 1}, {normal xyz, 0}: one 32 B
//                                   record per point (model.c L14-25 keeps two
//                                   float[3] arrays)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "octree_ptx.cuh"

namespace qb
{

// Node n of the reference array lives at DEVICE index n + 1 and child words hold device indices, so that 0 -- the
// reference's "no child" -- addresses device node 0, an all-zero dummy (no children, model 0).  Descending into an
// absent subtree therefore needs no test: it reads the dummy and stays there (octree_fsh.c L129 returns 0 for
// "node 0 below the root level").  The root is device node 1.  Never-uploaded nodes are zero as well, i.e. empty.
// `nodes` = device index of a slot that is kept zero just past the uploaded extent: an index beyond the extent is
// clamped to it with one min, which keeps "reads beyond the uploaded range return 0" without a compare-and-branch.
struct TreeDev
{
    const uint2* slot; // 8 x { child device index, child-exists mask OF THAT CHILD } per device node (fast traversal)
    const int4* child; // 2 x int4 per device node
    const int*  model; // oct[8] per device node
    int         nodes; // reference nodes uploaded so far + 1 = the zero slot every larger index is clamped to
};
__device__ __forceinline__ int tree_clamp(const TreeDev& t, int node)
{
    return (int) min((unsigned) node, (unsigned) t.nodes);
}
// device index -> reference index for reporting (0 for "absent", like the reference's stack)
__device__ __forceinline__ int ref_node(int dev) { return dev > 0 ? dev - 1 : 0; }
constexpr int ROOT_NODE = 1;

// The mask half of a slot record (v19): the child's 8-bit child-exists mask REPLICATED into the four bytes of the word --
// the form the expansion tests it in (the one-hot octants of up to four candidates, one per byte, against the mask in
// one AND: octree_trace_fast.cuh, g_order_lut value.y).  The replication is paid once, where the record is derived,
// instead of one multiply per traversal step, and the two trees' words are merged by the same three-input logic
// instruction that applies them, at the END of the expansion: the record loads of a descent have a whole traversal
// step to arrive (v18 merged the two masks at the end of the descent: 62 % of the kernel's long-scoreboard stall
// samples sat on that one instruction, profiles/r2_ncu_v18_*).  It pays only together with the unconditional first ray
// set-up of render_fast_kernel (QB_FIRST_RAY_ALIVE_ONLY there): by itself ptxas parks a wait for the camera-position
// load, which shares the records' scoreboard, on the first instructions of the expansion
// (profiles/r2_variants_ab_v19.json, r2_ncu_v19_experiment_scoreboards.txt).  -DQB_MASK_EARLY restores v18's form.
#if !defined(QB_MASK_EARLY) && !defined(QB_MASK_LATE)
    #define QB_MASK_LATE 1
#endif
#ifdef QB_MASK_LATE
constexpr unsigned SLOT_MASK_REP = 0x01010101u;
#else
constexpr unsigned SLOT_MASK_REP = 1u;
#endif

struct PointsDev
{
    const float4* rec; // 2 x float4 per point
    int           points;
};

// work counters, same order as octree_cuc_counters
enum
{
    CNT_RAYS_PRIMARY = 0,
    CNT_RAYS_SHADOW,
    CNT_RAYS_DISC,
    CNT_EXPAND_S,
    CNT_EXPAND_D,
    CNT_LEAF_S,
    CNT_LEAF_D,
    CNT_HITS,
    CNT_DISCARDS,
    CNT_DESCENTS,
    CNT_COUNT
};

// one view: everything main() of octree_fsh.c derives from the uniforms that is
// constant over the frame, computed on the host in fp32
struct ViewParams
{
    float camfp[3];
    float light[3];
    float qz[4];         // octree_fsh.c L406
    float qx[4];         // L408
    float cfp[3];        // L403
    float camlight_n[3]; // normalize(light - camfp), L417-418
    float disc_dot_min;  // smallest dot with acosf(dot) < 0.02f (L452), host libm
    int   shoot;
};

// Multi-GPU completion fence (octree_cuc_set_fence / octree_cuc_set_gpus).  One frame is split by image tiles over
// `n` connectors ("ranks"), all storing into rank 0's framebuffer; each connector owns FENCE_WORDS 32-bit words:
//   word k (k < 64)       on rank 0: the last frame number rank k has finished storing (written by a 1-warp kernel
//                         that follows rank k's render kernel on its stream)
//   word FENCE_CONSUMED   on rank k: the last frame number rank 0 is done with (written by rank 0's next kernel)
// Frame numbers count from 1 and are compared as signed differences.  All pointers null = no fence.
constexpr int FENCE_CONSUMED = 64;
constexpr int FENCE_WORDS    = 128;
struct FenceDev
{
    unsigned*        done;      // ranks != 0: where this rank publishes `seq` (word `rank` of rank 0's fence words)
    const unsigned*  gate;      // ranks != 0: own FENCE_CONSUMED word; pixel stores of frame `seq` wait for >= seq - 1
    unsigned* const* peers;     // rank 0: device table of every rank's fence words (entry 0 unused)
    unsigned         seq;       // number of this frame
    int              n;         // ranks
};

struct FrameParams
{
    TreeDev   tree_s, tree_d;
    PointsDev pts_s, pts_d;

    float basecube[4]; // (0, S, S, S), octree_glc.c L263
    int   maxlevel;
    float leaf_size;   // S / 2^maxlevel (fast kernel, exact-grid mode)
    float inv_leaf_size;

    int   W, H;   // viewport in pixels
    float sx, sy; // coord scale = ow / W (1.0 for whole-number render sizes)

    // outputs
    uchar4*             frame;        // RGBA8, row 0 = bottom
    size_t              pitch;        // pixels per row
    size_t              view_stride;  // pixels between views of a batch
    uint8_t*            flags;        // optional parity planes
    int*                aux;
    unsigned long long* counters;

    // image-tile sharding
    int tile_w, tile_h, tiles_x, tiles_y;
    int rank, world;
    int blocks_per_tile_x, blocks_per_tile_y;
    unsigned tiles_x_magic, bptx_magic; // ceil(2^32 / tiles_x), ceil(2^32 / blocks_per_tile_x): n / d = umulhi(n, magic) for n < 65536; 0 when d = 1
    int tiles_mine; // number of tiles this rank renders per view

    // tile scheduling by measured cost (fast kernel, single view): launch position -> tile of this rank, heaviest
    // first, learned from an earlier frame of the same or the most recent view; the kernel accumulates every
    // tile's warp-cycles for the next frame.  Both null = tiles in image order.
    const int* tile_order;
    unsigned*  tile_cost;

    // views
    const ViewParams* views; // device array, n_views entries
    int               n_views;

    FenceDev fence;
};

constexpr unsigned CHILD_INDEX_MASK = 0x0FFFFFFFu; // 2^28 nodes per tree
constexpr int      CHILD_MASK_SHIFT = 28;

constexpr int BLOCK_W       = 16; // pixels per CTA in x
constexpr int BLOCK_H       = 8;  // pixels per CTA in y
constexpr int BLOCK_THREADS = BLOCK_W * BLOCK_H;

} // namespace qb
