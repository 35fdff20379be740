// emu.cpp -- TEST INFRASTRUCTURE: the fast traversal kernels of qubatron_b200/csrc compiled for the HOST by g++ and
// run one thread at a time, so that `-m "not gpu"` tests can check the LOGIC of the product's own kernel source
// (candidate ordering, backtrack, ray sequencing, shading, base-cube entry) against the oracle without a GPU.
// What it cannot show is the device's code generation (FMA contraction, MUFU seeds, packed fp32): that is what the
// `-m gpu` parity tests are for.  Nothing under qubatron_b200/ includes, links or loads this file.
//
// The scene arrays arrive in the reference's formats (12-int nodes octree.c L11-14, float[3] colours / normals
// model.c L14-25) and are laid out here exactly as the connector's upload kernels lay them out in HBM
// (octree_types.cuh: child blocks with mask nibbles, slot records, model array, 32-byte point records).
#include "cuda_host_shim.h"
extern "C" int qb_emu_dyn_shared[3 * (16 + 4) * 128 + 64];
int            qb_emu_dyn_shared[3 * (16 + 4) * 128 + 64];

#include "octree_trace_fast.cuh"
#include "octree_view_host.h"

#include <vector>

using namespace qb;

namespace
{
struct Tree
{
    std::vector<int4>  child; // 2 per device node
    std::vector<uint2> slot;  // 8 per device node
    std::vector<int>   model;
    long               ref_nodes = 0;
    TreeDev dev() const
    {
        TreeDev D;
        D.child = child.data();
        D.slot  = slot.data();
        D.model = model.data();
        D.nodes = (int) ref_nodes + 1;
        return D;
    }
};
struct Points
{
    std::vector<float4> rec;
    long                points = 0;
};
struct Scene
{
    Tree   tree[2];
    Points pts[2];
};

// relayout_octree_nodes_kernel + derive_slots_kernel / propagate_masks_kernel of octree_cuc.cu, sequentially
void build_tree(Tree& T, const int* oct, long nodes)
{
    T.ref_nodes = nodes;
    const long dev_nodes = nodes + 2; // dummy in front, one zero slot behind
    T.child.assign(2 * dev_nodes, make_int4(0, 0, 0, 0));
    T.model.assign(dev_nodes, 0);
    T.slot.assign(8 * dev_nodes, make_uint2(0u, 0u));
    const unsigned max_dev = (unsigned) (dev_nodes - 1);
    auto           ix      = [max_dev](int v) -> int {
        const unsigned idx = v > 0 ? (((unsigned) v + 1u) & CHILD_INDEX_MASK) : 0u;
        return (int) (idx <= max_dev ? idx : 0u);
    };
    for (long n = 0; n < nodes; n++)
    {
        const int* o = oct + 12 * n;
        unsigned   m = 0;
        for (int k = 0; k < 8; k++) m |= o[k] > 0 ? 1u << k : 0u;
        T.child[2 * (n + 1)]     = make_int4(ix(o[0]) | (int) ((m & 15u) << CHILD_MASK_SHIFT),
                                             ix(o[1]) | (int) ((m >> 4) << CHILD_MASK_SHIFT), ix(o[2]), ix(o[3]));
        T.child[2 * (n + 1) + 1] = make_int4(ix(o[4]), ix(o[5]), ix(o[6]), ix(o[7]));
        T.model[n + 1]           = o[8];
    }
    const int* c0 = (const int*) T.child.data();
    auto own_mask = [c0](unsigned node) {
        return ((unsigned) c0[(size_t) node * 8] >> CHILD_MASK_SHIFT) |
               (((unsigned) c0[(size_t) node * 8 + 1] >> CHILD_MASK_SHIFT) << 4);
    };
    for (long e = 0; e < 8 * dev_nodes; e++)
    {
        const unsigned idx = (unsigned) c0[e] & CHILD_INDEX_MASK;
        T.slot[e]          = make_uint2(idx, idx ? own_mask(idx) * SLOT_MASK_REP : 0u);
    }
}
void build_points(Points& Q, const float* col, const float* nrm, long n)
{
    Q.points = n;
    Q.rec.assign(2 * (n > 0 ? n : 1), make_float4(0.f, 0.f, 0.f, 0.f));
    for (long i = 0; i < n; i++)
    {
        Q.rec[2 * i]     = make_float4(col[3 * i], col[3 * i + 1], col[3 * i + 2], 1.0f);
        Q.rec[2 * i + 1] = make_float4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], 0.0f);
    }
}

void fill_common(FrameParams& P, const Scene& S, int maxlevel, float basesize)
{
    memset(&P, 0, sizeof(P));
    P.tree_s        = S.tree[0].dev();
    P.tree_d        = S.tree[1].dev();
    P.pts_s.rec     = S.pts[0].rec.data();
    P.pts_s.points  = (int) S.pts[0].points;
    P.pts_d.rec     = S.pts[1].rec.data();
    P.pts_d.points  = (int) S.pts[1].points;
    P.basecube[0]   = 0.0f;
    P.basecube[1] = P.basecube[2] = P.basecube[3] = basesize;
    P.maxlevel      = maxlevel;
    P.leaf_size     = ldexpf(basesize, -maxlevel);
    P.inv_leaf_size = 1.0f / P.leaf_size;
}

template <int DIV, bool DYN>
void run_frame(FrameParams P, int cta_row0, int cta_row1)
{
    blockDim.x = BLOCK_THREADS, blockDim.y = blockDim.z = 1;
    gridDim.x = (unsigned) (P.blocks_per_tile_x * P.blocks_per_tile_y), gridDim.y = (unsigned) P.tiles_mine, gridDim.z = 1;
    {
        // the CTA-wide tables a real CTA fills cooperatively before its barrier: one pass with an empty viewport
        FrameParams Q = P;
        Q.W = Q.H = 0;
        blockIdx.x = blockIdx.y = blockIdx.z = 0;
        for (unsigned t = 0; t < BLOCK_THREADS; t++)
        {
            threadIdx.x = t;
            render_fast_kernel<DIV, DYN, true, false>(Q);
        }
    }
    for (unsigned ty = 0; ty < gridDim.y; ty++)
        for (unsigned sub = 0; sub < gridDim.x; sub++)
        {
            // pixel rows of this CTA, to skip the ones outside the requested band
            const int tile = (int) ty, tyy = tile / P.tiles_x;
            const int by   = (int) sub / P.blocks_per_tile_x;
            const int py0  = tyy * P.tile_h + by * BLOCK_H;
            if (py0 + BLOCK_H <= cta_row0 || py0 >= cta_row1) continue;
            blockIdx.x = sub, blockIdx.y = ty, blockIdx.z = 0;
            for (unsigned t = 0; t < BLOCK_THREADS; t++)
            {
                threadIdx.x = t;
                render_fast_kernel<DIV, DYN, true, false>(P);
            }
        }
}
} // namespace

extern "C"
{
void* qb_emu_scene_create(const int* oct_s, long nodes_s, const int* oct_d, long nodes_d, const float* col_s,
                          const float* nrm_s, long pts_s, const float* col_d, const float* nrm_d, long pts_d)
{
    Scene* S = new Scene;
    build_tree(S->tree[0], oct_s, nodes_s);
    build_tree(S->tree[1], oct_d, nodes_d);
    build_points(S->pts[0], col_s, nrm_s, pts_s);
    build_points(S->pts[1], col_d, nrm_d, pts_d);
    return S;
}
void qb_emu_scene_destroy(void* s) { delete (Scene*) s; }

// viewport of octree_glc_update(width, height, quality)
void qb_emu_render_size(float width, float height, int quality, int* W, int* H)
{
    float ow, oh;
    viewhost::render_size(width, height, (uint8_t) quality, ow, oh, *W, *H);
}

// One frame of render_fast_kernel (parity planes on): rgba [H][W][4], flags [H][W], aux [H][W][6]; only pixel rows
// [row0, row1) are rendered (whole CTA rows).  div: 0 = GLSL a * (1 / b), 1 = IEEE.  Returns 0, or -1 when the
// base cube is not one the fast kernel accepts.
int qb_emu_render(void* scene, float width, float height, int quality, const float* position, const float* angle,
                  float lighta, int maxlevel, float basesize, int shoot, const float* light_override, int div, int row0,
                  int row1, uint8_t* rgba, uint8_t* flags, int* aux)
{
    const Scene& S = *(const Scene*) scene;
    float        ow, oh;
    int          W, H;
    viewhost::render_size(width, height, (uint8_t) quality, ow, oh, W, H);
    if (W <= 0 || H <= 0 || maxlevel < 1 || maxlevel > FAST_MAX_LEVELS) return -1;
    ViewParams V;
    memset(&V, 0, sizeof(V));
    viewhost::fill_view(V, position, angle, lighta, shoot, light_override, div == DIV_GLSL);
    viewhost::fill_cfp(V, ow, oh);

    FrameParams P;
    fill_common(P, S, maxlevel, basesize);
    P.W = W, P.H = H;
    P.sx = ow / (float) W, P.sy = oh / (float) H;
    // an odd pitch keeps the frame store on its one-pixel-per-thread path (the 16-byte path hands pixels across lanes)
    const size_t           pitch = (size_t) W | 1u;
    std::vector<uchar4>    frame(pitch * (size_t) H, make_uchar4(0, 0, 0, 0));
    P.frame       = frame.data();
    P.pitch       = pitch;
    P.view_stride = pitch * (size_t) H;
    P.flags       = flags;
    P.aux         = aux;
    P.tile_w = 64, P.tile_h = 64;
    P.tiles_x = (W + 63) / 64, P.tiles_y = (H + 63) / 64;
    P.rank = 0, P.world = 1;
    P.blocks_per_tile_x = P.tile_w / BLOCK_W, P.blocks_per_tile_y = P.tile_h / BLOCK_H;
    P.tiles_mine = P.tiles_x * P.tiles_y;
    auto magic = [](int d) { return d <= 1 ? 0u : (unsigned) (((1ull << 32) + (unsigned) d - 1) / (unsigned) d); }; // 0: d = 1
    P.tiles_x_magic = magic(P.tiles_x);
    P.bptx_magic    = magic(P.blocks_per_tile_x);
    P.views      = &V;
    P.n_views    = 1;

    const bool dyn = P.tree_d.nodes > 2;
    if (div == DIV_GLSL)
        dyn ? run_frame<DIV_GLSL, true>(P, row0, row1) : run_frame<DIV_GLSL, false>(P, row0, row1);
    else
        dyn ? run_frame<DIV_IEEE, true>(P, row0, row1) : run_frame<DIV_IEEE, false>(P, row0, row1);
    for (int y = 0; y < H; y++) memcpy(rgba + (size_t) y * W * 4, frame.data() + (size_t) y * pitch, (size_t) W * 4);
    return 0;
}

// base_cube_entry_compact (the fast kernel's statement of octree_fsh.c L157-211) against base_cube_entry_q (the
// face-by-face statement the generic kernel and the oracle-checked v14 use), same quotient rule: returns how many of
// the n rays differ in the discard decision or in any bit of the entry point; first_bad = index of the first
long qb_emu_entry_compare(long n, const float* pos, const float* dir, const float* basecube, int div, long* first_bad)
{
    long bad = 0;
    for (long i = 0; i < n; i++)
    {
        const float ox = pos[3 * i], oy = pos[3 * i + 1], oz = pos[3 * i + 2];
        const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        const float rx = 1.0f / dx, ry = 1.0f / dy, rz = 1.0f / dz;
        auto quot = [&](float nn, int axis) -> float {
            const float d = axis == 0 ? dx : (axis == 1 ? dy : dz);
            const float r = axis == 0 ? rx : (axis == 1 ? ry : rz);
            return div == DIV_GLSL ? nn * r : nn / d;
        };
        float4     a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        const bool ha = base_cube_entry_q(basecube, make_float3(ox, oy, oz), make_float3(dx, dy, dz), a, quot);
        const bool hb = base_cube_entry_compact(basecube, ox, oy, oz, dx, dy, dz, b, quot);
        const bool same = ha == hb && (!ha || memcmp(&a, &b, sizeof(a)) == 0);
        if (!same && bad++ == 0 && first_bad) *first_bad = i;
    }
    return bad;
}

// octree_cuc_trace_lines on an exact grid: trace_lines_fast_kernel, one thread at a time
int qb_emu_trace_lines(void* scene, long n, const float* pos, const float* dir, int dynamic_tree, int maxlevel,
                       float basesize, int* out_index, float* out_tlf)
{
    const Scene& S = *(const Scene*) scene;
    FrameParams  P;
    fill_common(P, S, maxlevel, basesize);
    P.tree_s       = S.tree[dynamic_tree ? 1 : 0].dev(); // as octree_cuc_trace_lines: the queried tree in the first
    P.tree_d       = P.tree_s;                           // slot, nothing in the second (every index clamps to the dummy)
    P.tree_d.nodes = 0;
    blockDim.x = BLOCK_THREADS, blockDim.y = blockDim.z = 1;
    gridDim.x = (unsigned) ((n + BLOCK_THREADS - 1) / BLOCK_THREADS), gridDim.y = gridDim.z = 1;
    blockIdx.x = blockIdx.y = blockIdx.z = 0;
    for (unsigned t = 0; t < BLOCK_THREADS; t++) // the CTA-wide selector table
    {
        threadIdx.x = t;
        trace_lines_fast_kernel(P, 0, pos, dir, out_index, out_tlf);
    }
    for (unsigned b = 0; b < gridDim.x; b++)
        for (unsigned t = 0; t < BLOCK_THREADS; t++)
        {
            blockIdx.x = b, threadIdx.x = t;
            trace_lines_fast_kernel(P, (size_t) n, pos, dir, out_index, out_tlf);
        }
    return 0;
}
}
