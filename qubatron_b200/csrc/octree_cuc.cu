// octree_cuc.cu -- the C-ABI connector (include/octree_cuc.h): device memory,
// range uploads with on-device relayout, per-frame uniform set-up and kernel
// launches.  Replaces /root/reference/src/qubatron/octree_glc.c.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
// (see qubatron_b200/csrc/Makefile).  No CPU fallback: every entry point that
// needs the GPU aborts with a message if CUDA is unusable.
#include "../../include/octree_cuc.h"

#include "octree_build.cuh"
#include "octree_render.cuh"
#include "octree_trace_fast.cuh"
#include "skeleton_skin.cuh"
#include "particle_sim.cuh"
#include "present.cuh"
#include "octree_view_host.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace
{

using namespace qb;

// Errors: the message goes to stderr and to the host's handler (octree_cuc_set_error_handler), then the process
// aborts -- unless the handler does not return (longjmp, exit).  The connector never returns partial state and never
// falls back to the CPU.
typedef void (*error_fn_t)(const char*, void*);
error_fn_t g_error_fn   = nullptr;
void*      g_error_user = nullptr;

[[noreturn]] void die(const char* msg)
{
    fprintf(stderr, "octree_cuc: %s\n", msg);
    if (g_error_fn) g_error_fn(msg, g_error_user);
    abort();
}

#define CUDA_OK(call)                                                                                                 \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e__ = (call);                                                                                     \
        if (e__ != cudaSuccess)                                                                                       \
        {                                                                                                             \
            char m__[512];                                                                                            \
            snprintf(m__, sizeof(m__), "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,       \
                     cudaGetErrorString(e__));                                                                        \
            die(m__);                                                                                                 \
        }                                                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// relayout kernels: the host side speaks the reference's formats (12-int nodes,
// float[3] points); the device arrays are the traversal layout of
// octree_types.cuh.  Both kernels move 4-byte words.
// ---------------------------------------------------------------------------

// one 32-bit word of the reference node array -> traversal layout.  `child` / `model` point at reference node 0,
// which is device node 1 (octree_types.cuh); child words keep the DEVICE index (reference index + 1, 0 = absent) in
// bits 0-27; the node's child-exists mask lives in the top nibble of words 0 and 1.  Atomics on disjoint bit fields
// make concurrent updates of one node by several threads safe.
// A child word that points past the device array (`max_dev` = its last device index) can only come from an
// inconsistent host tree; it is stored as "the dummy" -- the fast kernel follows child words without a bounds test.
// The child-exists bit keeps the reference's meaning (index > 0): such a candidate is kept and leads nowhere.
__device__ __forceinline__ unsigned device_child(int v, unsigned max_dev)
{
    const unsigned idx = v > 0 ? (((unsigned) v + 1u) & CHILD_INDEX_MASK) : 0u;
    return idx <= max_dev ? idx : 0u;
}
__device__ __forceinline__ void put_node_word(int* child, int* model, size_t node, int slot, int v, unsigned max_dev)
{
    if (slot < 8)
    {
        const unsigned idx  = device_child(v, max_dev);
        unsigned*      w    = (unsigned*) child + node * 8 + slot;
        unsigned*      mw   = (unsigned*) child + node * 8 + (slot >> 2);
        const unsigned bit  = 1u << (CHILD_MASK_SHIFT + (slot & 3));
        if (slot < 2)
        {
            atomicAnd(w, ~CHILD_INDEX_MASK);
            atomicOr(w, idx);
        }
        else
            *w = idx;
        if (v > 0)
            atomicOr(mw, bit);
        else
            atomicAnd(mw, ~bit);
    }
    else if (slot == 8)
        model[node] = v;
}

// words [first_word, first_word + nwords) of a 12-int node array -> child/model
__global__ void relayout_octree_kernel(const int* __restrict__ src, size_t first_word, size_t nwords,
                                       int* __restrict__ child, int* __restrict__ model, unsigned max_dev)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    size_t k    = first_word + i;
    size_t node = k / 12;
    put_node_word(child, model, node, (int) (k - node * 12), src[i], max_dev);
}

// whole nodes (the common bulk case): one thread converts one 48-byte node
__global__ void relayout_octree_nodes_kernel(const int4* __restrict__ src, size_t first_node, size_t nnodes,
                                             int4* __restrict__ child, int* __restrict__ model, unsigned max_dev)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    int4 a = src[3 * i], b = src[3 * i + 1], c = src[3 * i + 2];
    unsigned m = (a.x > 0 ? 1u : 0u) | (a.y > 0 ? 2u : 0u) | (a.z > 0 ? 4u : 0u) | (a.w > 0 ? 8u : 0u) |
                 (b.x > 0 ? 16u : 0u) | (b.y > 0 ? 32u : 0u) | (b.z > 0 ? 64u : 0u) | (b.w > 0 ? 128u : 0u);
    auto ix = [max_dev](int v) { return (int) device_child(v, max_dev); }; // device index
    int4 lo = make_int4(ix(a.x) | (int) ((m & 15u) << CHILD_MASK_SHIFT), ix(a.y) | (int) ((m >> 4) << CHILD_MASK_SHIFT),
                        ix(a.z), ix(a.w));
    int4 hi = make_int4(ix(b.x), ix(b.y), ix(b.z), ix(b.w));
    child[2 * (first_node + i)]     = lo;
    child[2 * (first_node + i) + 1] = hi;
    model[first_node + i]           = c.x;
}

// words of a float[3] array -> 32-byte point records; which = 0 colour, 1 normal
__global__ void relayout_points_kernel(const float* __restrict__ src, size_t first_word, size_t nwords,
                                       float* __restrict__ rec, int which)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    size_t k                       = first_word + i;
    size_t pt                      = k / 3;
    int    c                       = (int) (k - pt * 3);
    rec[pt * 8 + which * 4 + c]    = src[i];
    if (c == 0) rec[pt * 8 + which * 4 + 3] = which ? 0.0f : 1.0f;
}

// batched small ranges ("zero-and-append" node uploads, modelutil.c L429-501):
// one launch applies every pending range
struct RangeDesc
{
    unsigned long long dst_word; // word offset in the logical (reference-format) array
    unsigned int       src_word; // word offset in the packed payload
    unsigned int       nwords;
    int                buftype;
    int                pad;
};
struct ScatterTargets
{
    int*     child[2];
    int*     model[2];
    float*   rec[2];
    unsigned max_dev[2]; // last device node index of each tree's arrays
};
__global__ void scatter_ranges_kernel(const RangeDesc* __restrict__ descs, int ndesc, const int* __restrict__ payload,
                                      unsigned int total_words, ScatterTargets T)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_words) return;
    int lo = 0, hi = ndesc - 1; // last descriptor with src_word <= i
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (descs[mid].src_word <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    const RangeDesc d = descs[lo];
    size_t          k = d.dst_word + (i - d.src_word);
    int             v = payload[i];
    switch (d.buftype)
    {
        case OCTREE_GLC_BUFFER_STATIC_OCTREE:
        case OCTREE_GLC_BUFFER_DYNAMIC_OCTREE:
        {
            int    t    = d.buftype == OCTREE_GLC_BUFFER_DYNAMIC_OCTREE;
            size_t node = k / 12;
            put_node_word(T.child[t], T.model[t], node, (int) (k - node * 12), v, T.max_dev[t]);
            break;
        }
        default:
        {
            int t     = (d.buftype == OCTREE_GLC_BUFFER_DYNAMIC_COLOR || d.buftype == OCTREE_GLC_BUFFER_DYNAMIC_NORMAL);
            int which = (d.buftype == OCTREE_GLC_BUFFER_STATIC_NORMAL || d.buftype == OCTREE_GLC_BUFFER_DYNAMIC_NORMAL);
            size_t pt = k / 3;
            int    c  = (int) (k - pt * 3);
            T.rec[t][pt * 8 + which * 4 + c] = __int_as_float(v);
            if (c == 0) T.rec[t][pt * 8 + which * 4 + 3] = which ? 0.0f : 1.0f;
            break;
        }
    }
}

// ---------------------------------------------------------------------------
// Slot records of the fast traversal (octree_types.cuh, TreeDev::slot): slot[8 n + k] = { device index of child k of
// node n, child-exists mask OF THAT CHILD }, derived from the child array above (which stays the statement of the
// tree: the generic tracer, the builders and the download read it).  A descent then fetches everything the next
// expansion needs from the PARENT's record in one 8-byte load, and the expansion loads nothing.
// Keeping them current: whenever words of node n change, (A) its eight slot records are recomputed and every child
// learns where its record lives (`parent`), then (B) n's own mask -- it may have changed -- is written into n's
// record in ITS parent.  Two launches: B reads the parent links and masks A left complete; the check against the
// record's index skips links a re-pointed slot left stale.  Children have larger indices than their parents, so a
// bulk upload in chunks fixes a parent's records when the child's chunk arrives.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned own_mask_of(const int* child_dev0, unsigned dev_node)
{
    const int2 w = *(const int2*) (child_dev0 + (size_t) dev_node * 8);
    return ((unsigned) w.x >> CHILD_MASK_SHIFT) | (((unsigned) w.y >> CHILD_MASK_SHIFT) << 4);
}
// ... as a slot record carries it (octree_types.cuh SLOT_MASK_REP: the mask in every byte of the word)
__device__ __forceinline__ unsigned slot_mask_of(const int* child_dev0, unsigned dev_node)
{
    return own_mask_of(child_dev0, dev_node) * SLOT_MASK_REP;
}
// (A) for device nodes [first, first + count): all eight records + the children's parent links
__global__ void derive_slots_kernel(const int* __restrict__ child_dev0, uint2* __restrict__ slot,
                                    unsigned* __restrict__ parent, size_t first, size_t count)
{
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= count * 8) return;
    const size_t   e   = first * 8 + i; // record index = device node * 8 + octant
    const unsigned idx = (unsigned) child_dev0[e] & CHILD_INDEX_MASK;
    slot[e]            = make_uint2(idx, idx ? slot_mask_of(child_dev0, idx) : 0u);
    if (idx) parent[idx] = (unsigned) e;
}
// (B) for device nodes [first, first + count): the node's own mask into its record in its parent
__global__ void propagate_masks_kernel(const int* __restrict__ child_dev0, uint2* __restrict__ slot,
                                       const unsigned* __restrict__ parent, size_t first, size_t count)
{
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned n  = (unsigned) (first + i);
    const unsigned pl = parent[n];
    if (pl != 0u && slot[pl].x == n) slot[pl].y = slot_mask_of(child_dev0, n);
}
// the same two steps for the nodes touched by a batch of small ranges (one thread per payload word)
struct SlotTargets
{
    const int* child_dev0[2];
    uint2*     slot[2];
    unsigned*  parent[2];
};
template <int PHASE>
__global__ void derive_ranges_kernel(const RangeDesc* __restrict__ descs, int ndesc, unsigned int total_words,
                                     SlotTargets T)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_words) return;
    int lo = 0, hi = ndesc - 1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (descs[mid].src_word <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    const RangeDesc d = descs[lo];
    if (d.buftype != OCTREE_GLC_BUFFER_STATIC_OCTREE && d.buftype != OCTREE_GLC_BUFFER_DYNAMIC_OCTREE) return;
    const int      t    = d.buftype == OCTREE_GLC_BUFFER_DYNAMIC_OCTREE;
    const size_t   k    = d.dst_word + (i - d.src_word);
    const size_t   node = k / 12;
    const int      s    = (int) (k - node * 12);
    if (s >= 8) return;
    const unsigned n = (unsigned) node + 1u; // device node
    if (PHASE == 0)
    {
        const size_t   e   = (size_t) n * 8 + s;
        const unsigned idx = (unsigned) T.child_dev0[t][e] & CHILD_INDEX_MASK;
        T.slot[t][e]       = make_uint2(idx, idx ? slot_mask_of(T.child_dev0[t], idx) : 0u);
        if (idx) T.parent[t][idx] = (unsigned) e;
    }
    else
    {
        const unsigned pl = T.parent[t][n];
        if (pl != 0u && T.slot[t][pl].x == n) T.slot[t][pl].y = slot_mask_of(T.child_dev0[t], n);
    }
}

// ---------------------------------------------------------------------------
// connector state
// ---------------------------------------------------------------------------

struct DevArray
{
    void*  ptr   = nullptr;
    size_t bytes = 0;
};

struct Tree
{
    DevArray child; // 32 B per device node: [0] the all-zero dummy, [n + 1] reference node n, one spare zero slot
    DevArray model; // 4 B per device node
    DevArray slot;   // 64 B per device node: eight { child index, that child's mask } records (fast traversal)
    DevArray parent; // 4 B per device node: the record (8 * node + octant) that points to it, 0 = none
    size_t   cap_nodes = 0;        // reference nodes the arrays hold (device nodes: cap_nodes + 2)
    size_t   nodes     = 0;        // highest reference node uploaded + 1
    size_t   sealed    = (size_t) -1; // extent whose clamp slot (device node nodes + 1) has been zeroed
    // where the upload / build / export kernels see reference node 0
    int* up_child() const { return (int*) child.ptr + 8; }
    int* up_model() const { return (int*) model.ptr + 1; }
    // last device node index the arrays hold (device nodes: cap_nodes + 2)
    unsigned max_dev() const { return (unsigned) (cap_nodes + 1); }
};
struct Points
{
    DevArray rec; // 32 B per point
    size_t   cap_points = 0;
    size_t   points     = 0;
};

constexpr size_t STAGE_BYTES   = 48u * 1398101u; // ~64 MiB device staging chunk, a whole number of nodes and points
constexpr size_t BATCH_BYTES   = 8u << 20;  // pinned staging for small ranges
constexpr size_t BATCH_MAXDESC = 1u << 16;
constexpr size_t SMALL_RANGE   = 256u << 10; // ranges up to this size are batched
constexpr int    VIEW_RING     = 4;

// One persistent host thread per member device of an octree_cuc_set_gpus group, for the per-frame launches: queued
// from the engine's one thread one device after the other they cost ~14 us each (0.11 ms at eight devices, more than
// half of an 8-GPU 1080p frame); the workers queue them side by side.  A worker spins for ~1 ms after a frame (the
// next one of a running engine is there by then) and sleeps on a condition variable otherwise.
struct FrameWorker
{
    std::thread             th;
    std::mutex              m;
    std::condition_variable cv;
    std::function<void()>   job;
    std::atomic<uint64_t>   posted{0}, done{0};
    std::atomic<bool>       quit{false};

    FrameWorker() { th = std::thread([this]() { loop(); }); }
    ~FrameWorker()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            quit.store(true);
        }
        cv.notify_one();
        if (th.joinable()) th.join();
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;)
        {
            int spins = 0;
            while (posted.load(std::memory_order_acquire) == seen && !quit.load())
            {
                if (++spins < 20000)
                    __builtin_ia32_pause();
                else
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&]() { return posted.load() != seen || quit.load(); });
                }
            }
            if (quit.load()) return;
            seen = posted.load(std::memory_order_acquire);
            job();
            done.store(seen, std::memory_order_release);
        }
    }
    void post(std::function<void()> j)
    {
        job = std::move(j);
        {
            std::lock_guard<std::mutex> lk(m);
            posted.fetch_add(1, std::memory_order_release);
        }
        cv.notify_one();
    }
    void wait()
    {
        while (done.load(std::memory_order_acquire) != posted.load(std::memory_order_acquire)) __builtin_ia32_pause();
    }
};

// Bulk uploads from PAGEABLE host memory (what the unmodified engine passes to octree_glc_upload_texbuffer_data:
// the [minmodi, maxmodi) colour / normal ranges of a shot, modelutil.c L486-501; the level's arrays at start-up).
// A plain cudaMemcpy from pageable memory goes through the driver's bounce buffer at ~9-11 GB/s.  Here the range is
// cut into CHUNK-sized pieces; a few persistent host threads copy piece i + 1 into one of SLOTS page-locked buffers
// while the DMA engine moves piece i from another at PCIe rate.  Used by single-device connectors; the members of an
// in-process group keep the driver's path (each device's thread would copy the same source again).
struct HostStager
{
    static constexpr size_t CHUNK = 4u << 20;
    static constexpr int    SLOTS = 3;
    char*       pin[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[SLOTS]  = {nullptr, nullptr, nullptr};
    bool        used[SLOTS] = {false, false, false};
    int         next = 0;

    std::vector<std::thread> th;
    std::mutex               mu;
    std::condition_variable  cv_go, cv_done;
    const char*              src = nullptr;
    char*                    dst = nullptr;
    size_t                   bytes = 0;
    unsigned                 gen = 0;
    int                      pending = 0;
    bool                     stop = false;

    // slice k of parts: whole cache lines, the tail goes to the last one
    static void slice(size_t bytes, int parts, int k, size_t& a, size_t& b)
    {
        const size_t per = ((bytes / (size_t) parts) + 63) & ~(size_t) 63;
        a = per * (size_t) k < bytes ? per * (size_t) k : bytes;
        b = k == parts - 1 ? bytes : (per * (size_t) (k + 1) < bytes ? per * (size_t) (k + 1) : bytes);
    }
    void start(int threads)
    {
        for (int i = 0; i < SLOTS; i++)
        {
            CUDA_OK(cudaMallocHost(&pin[i], CHUNK));
            CUDA_OK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        const int parts = threads + 1; // the calling thread copies a slice too
        for (int k = 0; k < threads; k++)
            th.emplace_back([this, k, parts]() {
                unsigned seen = 0;
                for (;;)
                {
                    const char* s;
                    char*       d;
                    size_t      n;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv_go.wait(lk, [&]() { return stop || gen != seen; });
                        if (stop) return;
                        seen = gen, s = src, d = dst, n = bytes;
                    }
                    size_t a, b;
                    slice(n, parts, k + 1, a, b);
                    if (b > a) memcpy(d + a, s + a, b - a);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (--pending == 0) cv_done.notify_one();
                    }
                }
            });
    }
    // memcpy(d, s, n) on all threads
    void copy(char* d, const char* s, size_t n)
    {
        const int parts = (int) th.size() + 1;
        if (th.empty() || n < (256u << 10))
        {
            memcpy(d, s, n);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            src = s, dst = d, bytes = n, pending = (int) th.size();
            gen++;
        }
        cv_go.notify_all();
        size_t a, b;
        slice(n, parts, 0, a, b);
        if (b > a) memcpy(d + a, s + a, b - a);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&]() { return pending == 0; });
    }
    // dev[0, n) = host[0, n) queued on `st`; returns once the host range has been consumed
    void h2d(void* dev, const char* host, size_t n, cudaStream_t st)
    {
        for (size_t off = 0; off < n; off += CHUNK)
        {
            const size_t m = n - off < CHUNK ? n - off : CHUNK;
            const int    k = next;
            next           = (next + 1) % SLOTS;
            if (used[k]) CUDA_OK(cudaEventSynchronize(ev[k])); // the DMA that last read this buffer
            copy(pin[k], host + off, m);
            CUDA_OK(cudaMemcpyAsync((char*) dev + off, pin[k], m, cudaMemcpyHostToDevice, st));
            CUDA_OK(cudaEventRecord(ev[k], st));
            used[k] = true;
        }
    }
    void shutdown()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& t : th) t.join();
        th.clear();
        for (int i = 0; i < SLOTS; i++)
        {
            if (ev[i]) cudaEventSynchronize(ev[i]), cudaEventDestroy(ev[i]);
            if (pin[i]) cudaFreeHost(pin[i]);
            pin[i] = nullptr, ev[i] = nullptr, used[i] = false;
        }
    }
};

struct Impl
{
    int          device = 0;
    cudaStream_t stream = nullptr;     // where all work is queued
    cudaStream_t own_stream = nullptr; // created at init; `stream` unless the caller set one
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    bool         timed = false;

    Tree   tree[2];
    Points pts[2];

    void* stage_dev = nullptr; // STAGE_BYTES
    // bulk uploads from pageable memory: page-locked staging filled by `upload_threads` host threads (0 = the driver's
    // own pageable path); created at the first such upload
    int                         upload_threads = 4;
    std::unique_ptr<HostStager> stager;

    // pending small ranges
    char*                          batch_host = nullptr; // pinned, BATCH_BYTES
    RangeDesc*                     desc_host  = nullptr; // pinned
    void*                          batch_dev  = nullptr;
    RangeDesc*                     desc_dev   = nullptr;
    size_t                         batch_used = 0;
    std::vector<RangeDesc>         descs;
    std::map<unsigned long long, std::pair<unsigned long long, int>> intervals[6]; // start -> (end, desc index)

    // frame
    uchar4*             frame       = nullptr;
    size_t              frame_cap   = 0; // pixels
    // pipelined readback (octree_cuc_read_frame_async): a second framebuffer so that the copy of frame i to the
    // host overlaps the rendering of frame i+1
    uchar4*             frame_alt   = nullptr;
    size_t              frame_alt_cap = 0;
    bool                ring_on     = false;
    int                 ring_cur    = 0;      // which of {frame, frame_alt} holds the last rendered frame
    cudaStream_t        copy_stream = nullptr;
    cudaEvent_t         ev_render[2] = {};
    cudaEvent_t         ev_copy[2]   = {};
    bool                copy_pending[2] = {};
    // staged async readback (render target stays in place: frames other ranks store into over NVLink)
    uchar4*     stage_frame[2]   = {nullptr, nullptr};
    size_t      stage_cap        = 0; // pixels
    cudaEvent_t ev_stage_ready[2] = {};
    cudaEvent_t ev_stage_done[2]  = {};
    bool        stage_pending[2]  = {};
    bool        stage_on          = false;
    int         stage_k           = 0;
    uint64_t            ext_target  = 0;
    size_t              ext_pitch   = 0;
    uint8_t*            flags       = nullptr;
    int*                aux         = nullptr;
    size_t              aux_cap     = 0; // pixels
    bool                aux_on      = false;
    bool                count_on    = false;
    unsigned long long* counters    = nullptr;
    ViewParams*         views_host  = nullptr; // pinned
    ViewParams*         views_dev   = nullptr;
    int                 views_cap   = 0;
    cudaEvent_t         slot_ev[VIEW_RING] = {};
    bool                slot_used[VIEW_RING] = {};
    uint64_t            frame_seq   = 0;
    int                 W = 0, H = 0, n_views = 0;

    int   shard_rank = 0, shard_world = 1, tile_w = 64, tile_h = 64;
    bool  light_override = false;
    float light[3]       = {0, 0, 0};
    int   kernel_choice  = 0;
    int   last_kernel    = 0;
    int   div_mode       = DIV_GLSL; // division semantics, octree_trace_generic.cuh

    // skinning ("next" row 8f #1): inputs uploaded once, outputs kept on the device
    float* skin_pos = nullptr;
    float* skin_nrm = nullptr;
    size_t skin_n   = 0;
    int4 * skin_p14 = nullptr, *skin_p54 = nullptr, *skin_p94 = nullptr;
    float* skin_pnt_out = nullptr;
    size_t skin_count   = 0; // points of the last update
    // tile scheduling by measured cost (fast kernel): per remembered view an order of this rank's tiles
    static constexpr int ORDER_SLOTS = 8;
    struct ViewKey
    {
        float pos[3], ang[3], width, height, lighta;
        int   quality, maxlevel, tiles;
    };
    bool      feedback_on  = true;
    int*      order_dev    = nullptr; // [ORDER_SLOTS][order_cap]
    unsigned* cost_dev[2]  = {nullptr, nullptr};
    int       order_cap    = 0;
    int       cost_phase   = 0;
    ViewKey   order_key[ORDER_SLOTS];
    uint64_t  order_used[ORDER_SLOTS] = {0};
    bool      order_valid[ORDER_SLOTS] = {false};
    int       order_last   = -1;
    uint64_t  order_clock  = 0;
    // presentation ("next" row 8f #4b): the window image of the last frame
    bool    present_on  = false;
    uchar4* window      = nullptr;
    size_t  window_cap  = 0;
    int     window_w = 0, window_h = 0;
    // particle / dust simulation state ("next" row 8f #2): [kind][in, out] position and speed, ping-pong
    float*    part_pos[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    float*    part_spd[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t    part_n[2]      = {0, 0};
    int       part_cur[2]    = {0, 0}; // which buffer holds the current state
    unsigned* part_finished  = nullptr; // particles parked by the last step (device counter)
    bool   skin_rot_set = false; // per-bone rotations supplied by the host instead of libm
    float  skin_rot[90];

    uint64_t launches = 0;
    uint64_t memsize  = 0;
    double   upload_ms = 0.0; // host wall time spent in upload calls

    // multi-GPU completion fence (octree_cuc_set_fence, FenceDev in octree_types.cuh)
    unsigned*   fence_words     = nullptr; // FENCE_WORDS words on this device
    unsigned**  fence_peers_dev = nullptr; // rank 0: device table of every rank's fence words
    unsigned*   fence_peer0     = nullptr; // ranks != 0: rank 0's fence words as addressed from this device
    int         fence_rank = 0, fence_n = 1;
    unsigned    fence_seq  = 0;
    cudaEvent_t ev_done    = nullptr; // rank 0: every rank's tiles of the last frame are in the framebuffer
    bool                  smem_optin[2][2] = {{false, false}, {false, false}};
    int                   cta_cap          = 0; // octree_cuc_set_occupancy: resident CTAs per SM of the fast kernel (0 = all)
    size_t                l2_window_bytes  = 0; // octree_cuc_set_persisting_window
    cudaStream_t          l2_window_stream = nullptr;
    const void*           l2_window_base   = nullptr;
    bool                  l2_window_dirty  = false;
    bool                  defer_completion = false; // group primary: see render_views / group_render
    std::function<void()> deferred;
    // in-process multi-GPU group (octree_cuc_set_gpus): the connectors of the other devices, driven by the same calls
    std::vector<octree_glc_t> replicas;
    std::vector<std::unique_ptr<FrameWorker>> workers; // one per replica when every member has a device of its own
    bool                      is_replica = false;
    uint8_t*                  ext_flags  = nullptr; // replicas: the primary's parity planes
    int*                      ext_aux    = nullptr;
    // replication log (octree_cuc_enable_replication_log): every range this connector applied since the last export
    bool                   repl_on = false;
    std::vector<RangeDesc> repl_descs;
    // payload of the logged ranges in page-locked memory: it goes to the device at PCIe rate (export_pending_device)
    struct PinnedLog
    {
        char*  ptr = nullptr;
        size_t used = 0, cap = 0;
        size_t size() const { return used; }
        bool   empty() const { return used == 0; }
        void   clear() { used = 0; }
        const char* data() const { return ptr; }
        void append(const char* src, size_t n)
        {
            if (used + n > cap)
            {
                size_t ncap = (used + n) + (used + n) / 2 + (1u << 20);
                char*  np   = nullptr;
                if (cudaMallocHost(&np, ncap) != cudaSuccess) die("replication log: cannot allocate page-locked memory");
                if (used) memcpy(np, ptr, used);
                if (ptr) cudaFreeHost(ptr);
                ptr = np;
                cap = ncap;
            }
            memcpy(ptr + used, src, n);
            used += n;
        }
        void release()
        {
            if (ptr) cudaFreeHost(ptr);
            ptr  = nullptr;
            used = cap = 0;
        }
    } repl_payload;
};

// run `call` (which names the member connector `m`) on every connector of the group behind a primary
#define REPLAY(I, call)                                                                                               \
    do                                                                                                                \
    {                                                                                                                 \
        if (!(I)->replicas.empty())                                                                                   \
        {                                                                                                             \
            for (auto& r__ : (I)->replicas)                                                                           \
            {                                                                                                         \
                octree_glc_t* m = &r__;                                                                               \
                call;                                                                                                 \
            }                                                                                                         \
            CUDA_OK(cudaSetDevice((I)->device));                                                                      \
        }                                                                                                             \
    } while (0)

int g_selected_device = -1;

// A heavy call (bulk upload, tree build, skinning) on every connector of a group AT THE SAME TIME: each device's
// share runs on a thread of its own, because these calls wait for their device (count read-backs, staging reuse)
// and N of them one after the other would take N times as long.  fn(member, is_primary).
void run_on_group(octree_glc_t* rc, const std::function<void(octree_glc_t*, bool)>& fn)
{
    Impl* I = (Impl*) rc->impl;
    if (I->replicas.empty())
    {
        fn(rc, true);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(I->replicas.size());
    for (auto& r : I->replicas)
    {
        octree_glc_t* m = &r;
        th.emplace_back([&fn, m]() { fn(m, false); });
    }
    fn(rc, true);
    for (auto& t : th) t.join();
    CUDA_OK(cudaSetDevice(I->device));
}

Impl* impl_of(octree_glc_t* rc)
{
    if (!rc || !rc->impl) die("connector used before octree_glc_init");
    Impl* I = (Impl*) rc->impl;
    CUDA_OK(cudaSetDevice(I->device));
    return I;
}

void publish_memsize(octree_glc_t* rc, Impl* I)
{
    rc->memsize_bytes = I->memsize;
    rc->memsize       = I->memsize > 0xffffffffull ? 0xffffffffu : (unsigned int) I->memsize;
}

void dev_alloc(Impl* I, DevArray& a, size_t bytes)
{
    CUDA_OK(cudaMalloc(&a.ptr, bytes));
    CUDA_OK(cudaMemsetAsync(a.ptr, 0, bytes, I->stream));
    a.bytes = bytes;
    I->memsize += bytes;
}
void dev_free(Impl* I, DevArray& a)
{
    if (a.ptr)
    {
        CUDA_OK(cudaFree(a.ptr));
        I->memsize -= a.bytes;
    }
    a.ptr   = nullptr;
    a.bytes = 0;
}
// grow to `bytes`, keeping the old content, zero-filling the rest
void dev_grow(Impl* I, DevArray& a, size_t bytes)
{
    DevArray n;
    dev_alloc(I, n, bytes);
    if (a.ptr)
    {
        CUDA_OK(cudaMemcpyAsync(n.ptr, a.ptr, a.bytes, cudaMemcpyDeviceToDevice, I->stream));
        CUDA_OK(cudaStreamSynchronize(I->stream));
        dev_free(I, a);
    }
    a = n;
}

size_t grown(size_t want) { return want + want / 4 + 4096; } // capacity policy: +25 %

ScatterTargets scatter_targets(Impl* I)
{
    ScatterTargets T;
    for (int t = 0; t < 2; t++)
    {
        T.child[t] = I->tree[t].up_child();
        T.model[t] = I->tree[t].up_model();
        T.rec[t]   = (float*) I->pts[t].rec.ptr;
        T.max_dev[t] = I->tree[t].max_dev();
    }
    return T;
}

bool is_octree(int buftype);

// slot records of device nodes [first_ref + 1, first_ref + 1 + count) after their words changed (see derive_slots_kernel)
void derive_slots(Impl* I, int t, size_t first_ref, size_t count)
{
    if (count == 0) return;
    Tree&      T  = I->tree[t];
    const int* c0 = (const int*) T.child.ptr;
    derive_slots_kernel<<<(unsigned) ((count * 8 + 255) / 256), 256, 0, I->stream>>>(c0, (uint2*) T.slot.ptr,
                                                                                     (unsigned*) T.parent.ptr, first_ref + 1, count);
    propagate_masks_kernel<<<(unsigned) ((count + 255) / 256), 256, 0, I->stream>>>(c0, (uint2*) T.slot.ptr,
                                                                                   (const unsigned*) T.parent.ptr, first_ref + 1, count);
    CUDA_OK(cudaGetLastError());
    I->launches += 2;
}
SlotTargets slot_targets(Impl* I)
{
    SlotTargets S;
    for (int t = 0; t < 2; t++)
    {
        S.child_dev0[t] = (const int*) I->tree[t].child.ptr;
        S.slot[t]       = (uint2*) I->tree[t].slot.ptr;
        S.parent[t]     = (unsigned*) I->tree[t].parent.ptr;
    }
    return S;
}
void derive_ranges(Impl* I, const RangeDesc* descs_dev, int nd, unsigned words)
{
    derive_ranges_kernel<0><<<(words + 255) / 256, 256, 0, I->stream>>>(descs_dev, nd, words, slot_targets(I));
    derive_ranges_kernel<1><<<(words + 255) / 256, 256, 0, I->stream>>>(descs_dev, nd, words, slot_targets(I));
    CUDA_OK(cudaGetLastError());
    I->launches += 2;
}

void flush_pending(Impl* I)
{
    if (I->descs.empty()) return;
    const size_t nd = I->descs.size();
    memcpy(I->desc_host, I->descs.data(), nd * sizeof(RangeDesc));
    CUDA_OK(cudaMemcpyAsync(I->batch_dev, I->batch_host, I->batch_used, cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaMemcpyAsync(I->desc_dev, I->desc_host, nd * sizeof(RangeDesc), cudaMemcpyHostToDevice, I->stream));
    unsigned int words = (unsigned int) (I->batch_used / 4);
    scatter_ranges_kernel<<<(words + 255) / 256, 256, 0, I->stream>>>(I->desc_dev, (int) nd, (const int*) I->batch_dev,
                                                                       words, scatter_targets(I));
    CUDA_OK(cudaGetLastError());
    I->launches++;
    bool any_octree = false;
    for (const RangeDesc& d : I->descs) any_octree = any_octree || is_octree(d.buftype);
    if (any_octree) derive_ranges(I, I->desc_dev, (int) nd, words);
    // the pinned staging is reused by the next batch
    CUDA_OK(cudaStreamSynchronize(I->stream));
    I->descs.clear();
    I->batch_used = 0;
    for (auto& m : I->intervals) m.clear();
}

bool is_octree(int buftype)
{
    return buftype == OCTREE_GLC_BUFFER_STATIC_OCTREE || buftype == OCTREE_GLC_BUFFER_DYNAMIC_OCTREE;
}
int tree_index(int buftype)
{
    return (buftype == OCTREE_GLC_BUFFER_DYNAMIC_COLOR || buftype == OCTREE_GLC_BUFFER_DYNAMIC_NORMAL ||
            buftype == OCTREE_GLC_BUFFER_DYNAMIC_OCTREE)
               ? 1
               : 0;
}

// make sure the device arrays behind `buftype` hold `size` logical bytes;
// returns true when they had to grow
bool ensure_capacity(Impl* I, int buftype, size_t size)
{
    const int t = tree_index(buftype);
    if (is_octree(buftype))
    {
        Tree&  T     = I->tree[t];
        size_t nodes = (size + 47) / 48;
        if (nodes > (size_t) CHILD_INDEX_MASK - 1) die("octree arrays are limited to 2^28 - 2 nodes per tree");
        if (nodes <= T.cap_nodes && T.child.ptr) return false;
        flush_pending(I);
        size_t cap = grown(nodes);
        dev_grow(I, T.child, (cap + 2) * 32); // + the dummy in front and the clamp slot behind
        dev_grow(I, T.model, (cap + 2) * 4);
        dev_grow(I, T.slot, (cap + 2) * 64);
        dev_grow(I, T.parent, (cap + 2) * 4);
        T.cap_nodes = cap;
        T.sealed    = (size_t) -1;
        return true;
    }
    Points& Pn     = I->pts[t];
    size_t  points = (size + 11) / 12;
    if (points <= Pn.cap_points) return false;
    flush_pending(I);
    size_t cap = grown(points);
    dev_grow(I, Pn.rec, cap * 32);
    Pn.cap_points = cap;
    return true;
}

void note_extent(Impl* I, int buftype, size_t end_byte)
{
    const int t = tree_index(buftype);
    if (is_octree(buftype))
    {
        size_t n = (end_byte + 47) / 48;
        if (n > I->tree[t].nodes) I->tree[t].nodes = n;
    }
    else
    {
        size_t n = (end_byte + 11) / 12;
        if (n > I->pts[t].points) I->pts[t].points = n;
    }
}

// `src` = the first byte of the range, i.e. logical byte s of the host array
void upload_bulk(Impl* I, const char* src, int buftype, size_t s, size_t e)
{
    flush_pending(I);
    const int t = tree_index(buftype);
    // a large range from pageable memory goes through page-locked staging filled by several host threads (HostStager)
    bool staged = false;
    if (I->upload_threads > 0 && e - s >= (1u << 20) && I->replicas.empty() && !I->is_replica)
    {
        cudaPointerAttributes at;
        const cudaError_t     q = cudaPointerGetAttributes(&at, src);
        if (q != cudaSuccess) (void) cudaGetLastError();
        staged = q != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
        if (staged && !I->stager)
        {
            I->stager.reset(new HostStager);
            I->stager->start(I->upload_threads);
        }
    }
    for (size_t off = s; off < e; off += STAGE_BYTES)
    {
        size_t n = e - off < STAGE_BYTES ? e - off : STAGE_BYTES;
        // pageable source: the call returns once the range has been consumed
        if (staged)
            I->stager->h2d(I->stage_dev, src + (off - s), n, I->stream);
        else
            CUDA_OK(cudaMemcpyAsync(I->stage_dev, src + (off - s), n, cudaMemcpyHostToDevice, I->stream));
        size_t   words  = n / 4;
        unsigned blocks = (unsigned) ((words + 255) / 256);
        if (is_octree(buftype) && off % 48 == 0 && n % 48 == 0)
        {
            size_t nn = n / 48;
            relayout_octree_nodes_kernel<<<(unsigned) ((nn + 255) / 256), 256, 0, I->stream>>>(
                (const int4*) I->stage_dev, off / 48, nn, (int4*) I->tree[t].up_child(), I->tree[t].up_model(),
                I->tree[t].max_dev());
        }
        else if (is_octree(buftype))
            relayout_octree_kernel<<<blocks, 256, 0, I->stream>>>((const int*) I->stage_dev, off / 4, words,
                                                                  I->tree[t].up_child(), I->tree[t].up_model(),
                                                                  I->tree[t].max_dev());
        if (is_octree(buftype)) derive_slots(I, t, off / 48, (off + n + 47) / 48 - off / 48);
        else
        {
            int which = (buftype == OCTREE_GLC_BUFFER_STATIC_NORMAL || buftype == OCTREE_GLC_BUFFER_DYNAMIC_NORMAL);
            relayout_points_kernel<<<blocks, 256, 0, I->stream>>>((const float*) I->stage_dev, off / 4, words,
                                                                  (float*) I->pts[t].rec.ptr, which);
        }
        CUDA_OK(cudaGetLastError());
        I->launches++;
    }
    // stage_dev is reused by the next call and `data` may be pageable: finish here
    CUDA_OK(cudaStreamSynchronize(I->stream));
}

void upload_batched(Impl* I, const char* src, int buftype, size_t s, size_t e)
{
    const size_t bytes = e - s;
    auto&        iv    = I->intervals[buftype];

    // identical range uploaded again: newest payload wins, in place
    auto same = iv.find(s);
    if (same != iv.end() && same->second.first == e)
    {
        const RangeDesc& d = I->descs[same->second.second];
        memcpy(I->batch_host + (size_t) d.src_word * 4, src, bytes);
        return;
    }
    // a different, overlapping pending range: apply what is queued first so that
    // ranges take effect in call order
    bool overlap = false;
    auto it      = iv.lower_bound(s);
    if (it != iv.end() && it->first < e) overlap = true;
    if (it != iv.begin())
    {
        auto pr = std::prev(it);
        if (pr->second.first > s) overlap = true;
    }
    if (overlap || I->batch_used + bytes > BATCH_BYTES || I->descs.size() + 1 > BATCH_MAXDESC) flush_pending(I);

    RangeDesc d;
    d.dst_word = s / 4;
    d.src_word = (unsigned int) (I->batch_used / 4);
    d.nwords   = (unsigned int) (bytes / 4);
    d.buftype  = buftype;
    d.pad      = 0;
    memcpy(I->batch_host + I->batch_used, src, bytes);
    I->batch_used += bytes;
    I->intervals[buftype][s] = std::make_pair((unsigned long long) e, (int) I->descs.size());
    I->descs.push_back(d);
}

// one texel-granular range [s, e) of the logical array behind `buftype`; `src` = its first byte
void apply_range(Impl* I, const char* src, int buftype, size_t s, size_t e)
{
    note_extent(I, buftype, e);
    if (e - s <= SMALL_RANGE)
        upload_batched(I, src, buftype, s, e);
    else
        upload_bulk(I, src, buftype, s, e);
}

// The kernels clamp every node index to device node `nodes + 1`, which must read as an empty node.  Memory behind
// the extent is zero from allocation unless an earlier, larger extent left nodes there (octree_reset + rebuild):
// zero that one slot whenever the extent has changed since the last launch.
TreeDev tree_dev(Impl* I, int t)
{
    Tree& T = I->tree[t];
    if (T.sealed != T.nodes)
    {
        CUDA_OK(cudaMemsetAsync((char*) T.child.ptr + (T.nodes + 1) * 32, 0, 32, I->stream));
        CUDA_OK(cudaMemsetAsync((char*) T.model.ptr + (T.nodes + 1) * 4, 0, 4, I->stream));
        T.sealed = T.nodes;
    }
    TreeDev D;
    D.child = (const int4*) T.child.ptr;
    D.slot  = (const uint2*) T.slot.ptr;
    D.model = (const int*) T.model.ptr;
    D.nodes = (int) T.nodes + 1;
    return D;
}

// bits(m) + maxlevel <= 24 for basesize = m * 2^e, m odd: every cube corner and
// centre k * basesize / 2^maxlevel is then exact in fp32, level sizes included
bool grid_is_exact(float basesize, int maxlevel)
{
    if (!(basesize > 0.0f) || !std::isfinite(basesize) || maxlevel < 1 || maxlevel > 16) return false;
    int   e;
    float m  = frexpf(basesize, &e); // basesize = m * 2^e, 0.5 <= m < 1
    uint32_t mi = (uint32_t) ldexpf(m, 24); // 24-bit integer mantissa
    while (mi && !(mi & 1)) mi >>= 1;
    int bits = 0;
    while (mi >> bits) bits++;
    if (bits + maxlevel > 23) return false;
    return e - 24 - maxlevel > -120; // far from the subnormal range
}

void fill_view(Impl* I, ViewParams& V, float ow, const float* position, const float* angle, float lighta, int shoot)
{
    (void) ow;
    viewhost::fill_view(V, position, angle, lighta, shoot, I->light_override ? I->light : nullptr, I->div_mode == DIV_GLSL);
}

template <int DIV>
void launch_generic_div(Impl* I, const FrameParams& P, unsigned blocks)
{
    if (I->aux_on && I->count_on)
        render_kernel<GenericTracer<DIV>, true, true><<<blocks, BLOCK_THREADS, 0, I->stream>>>(P);
    else if (I->aux_on)
        render_kernel<GenericTracer<DIV>, true, false><<<blocks, BLOCK_THREADS, 0, I->stream>>>(P);
    else if (I->count_on)
        render_kernel<GenericTracer<DIV>, false, true><<<blocks, BLOCK_THREADS, 0, I->stream>>>(P);
    else
        render_kernel<GenericTracer<DIV>, false, false><<<blocks, BLOCK_THREADS, 0, I->stream>>>(P);
}

void launch_generic(Impl* I, const FrameParams& P, unsigned blocks)
{
    if (I->div_mode == DIV_GLSL)
        launch_generic_div<DIV_GLSL>(I, P, blocks);
    else
        launch_generic_div<DIV_IEEE>(I, P, blocks);
}

template <int DIV, bool DYN>
void launch_fast_dyn(Impl* I, const FrameParams& P, unsigned blocks)
{
    size_t smem = (size_t) 3 * (P.maxlevel + FAST_RESULT_ROWS) * BLOCK_THREADS * sizeof(int); // levels + result rows
    // resident CTAs per SM capped by padding the dynamic shared memory (octree_cuc_set_occupancy): fewer warps per
    // SM run each of them faster, which is what a latency-bound shard of a frame split over many GPUs wants
    if (I->cta_cap > 0 && I->cta_cap < QB_MINBLOCKS)
    {
        const size_t per_cta = (size_t) (227 * 1024) / (size_t) (I->cta_cap + 1) + 1024; // k fit, k + 1 do not
        if (per_cta > smem) smem = per_cta;
        if (smem > 47 * 1024 && !(I->aux_on || I->count_on))
        {
            if (!I->smem_optin[DIV][DYN]) // per connector = per device: the attribute belongs to the device's context
            {
                CUDA_OK(cudaFuncSetAttribute(render_fast_kernel<DIV, DYN, false, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                I->smem_optin[DIV][DYN] = true;
            }
        }
        else if (smem > 47 * 1024)
            smem = 47 * 1024; // parity / counting instantiations stay below the 48 KB that needs no opt-in (static
                              // shared memory counts too): the cap is honoured down to 4 CTAs per SM there
    }
#ifdef QB_GRID_LINEAR
    const unsigned grid = blocks;
#else
    // (blocks of a tile, this shard's tiles, views): the kernel reads its place from blockIdx instead of dividing
    if (P.tiles_mine > 65535 || P.n_views > 65535) die("update: more than 65535 tiles per shard or views in one launch");
    const dim3 grid((unsigned) (P.blocks_per_tile_x * P.blocks_per_tile_y), (unsigned) P.tiles_mine, (unsigned) P.n_views);
    if ((size_t) grid.x * grid.y * grid.z != blocks) die("update: launch grid does not cover the frame");
#endif
    if (I->aux_on && I->count_on)
        render_fast_kernel<DIV, DYN, true, true><<<grid, BLOCK_THREADS, smem, I->stream>>>(P);
    else if (I->aux_on)
        render_fast_kernel<DIV, DYN, true, false><<<grid, BLOCK_THREADS, smem, I->stream>>>(P);
    else if (I->count_on)
        render_fast_kernel<DIV, DYN, false, true><<<grid, BLOCK_THREADS, smem, I->stream>>>(P);
    else
        render_fast_kernel<DIV, DYN, false, false><<<grid, BLOCK_THREADS, smem, I->stream>>>(P);
}

void launch_fast(Impl* I, const FrameParams& P, unsigned blocks)
{
    // a dynamic tree that is only a root (octree_reset, octree.c L89-93) has no geometry
    const bool dyn = P.tree_d.nodes > 2; // .nodes counts the dummy in front
    if (I->div_mode == DIV_GLSL)
        dyn ? launch_fast_dyn<DIV_GLSL, true>(I, P, blocks) : launch_fast_dyn<DIV_GLSL, false>(I, P, blocks);
    else
        dyn ? launch_fast_dyn<DIV_IEEE, true>(I, P, blocks) : launch_fast_dyn<DIV_IEEE, false>(I, P, blocks);
}

// octree_glc.c L268-269, L288: render size from the window size and the quality setting
using viewhost::render_size;

// the connector's own framebuffer / parity planes for `total` pixels (all views of a batch)
void ensure_frame(Impl* I, size_t total)
{
    if (total <= I->frame_cap) return;
    if (I->frame)
    {
        CUDA_OK(cudaStreamSynchronize(I->stream));
        if (I->copy_stream) CUDA_OK(cudaStreamSynchronize(I->copy_stream));
        CUDA_OK(cudaFree(I->frame));
        I->memsize -= I->frame_cap * 4;
    }
    CUDA_OK(cudaMalloc(&I->frame, total * 4));
    CUDA_OK(cudaMemsetAsync(I->frame, 0, total * 4, I->stream));
    I->frame_cap = total;
    I->memsize += I->frame_cap * 4;
}
void ensure_aux(Impl* I, size_t total)
{
    if (total <= I->aux_cap) return;
    if (I->flags)
    {
        CUDA_OK(cudaStreamSynchronize(I->stream));
        CUDA_OK(cudaFree(I->flags));
        CUDA_OK(cudaFree(I->aux));
        I->memsize -= I->aux_cap * 25;
    }
    CUDA_OK(cudaMalloc(&I->flags, total));
    CUDA_OK(cudaMalloc(&I->aux, total * 6 * sizeof(int)));
    // tiles of other ranks stay untouched: they must not read as garbage
    CUDA_OK(cudaMemsetAsync(I->flags, 0, total, I->stream));
    CUDA_OK(cudaMemsetAsync(I->aux, 0, total * 6 * sizeof(int), I->stream));
    I->aux_cap = total;
    I->memsize += I->aux_cap * 25;
}

// octree_cuc_set_persisting_window: see include/octree_cuc.h
void apply_l2_window(Impl* I)
{
    const void* base = I->tree[0].child.ptr;
    if (!I->l2_window_dirty && I->l2_window_stream == I->stream && I->l2_window_base == base) return;
    if (!I->l2_window_bytes && !I->l2_window_dirty) return;
    int max_persist = 0, max_window = 0;
    CUDA_OK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, I->device));
    CUDA_OK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, I->device));
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    if (I->l2_window_bytes && base && max_persist > 0 && max_window > 0)
    {
        const size_t persist = I->l2_window_bytes < (size_t) max_persist ? I->l2_window_bytes : (size_t) max_persist;
        CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
        size_t win = (I->tree[0].nodes + 2) * 32;
        if (win > (size_t) max_window) win = (size_t) max_window;
        a.accessPolicyWindow.base_ptr  = const_cast<void*>(base);
        a.accessPolicyWindow.num_bytes = win;
        a.accessPolicyWindow.hitRatio  = win <= persist ? 1.0f : (float) ((double) persist / (double) win);
        a.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
        a.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
        fprintf(stderr, "octree_cuc: L2 window %.1f MB of the static node array, %.1f MB persisting (device max %.1f / %.1f MB)\n",
                win / 1e6, persist / 1e6, max_window / 1e6, max_persist / 1e6);
    }
    else
    {
        a.accessPolicyWindow.num_bytes = 0; // disables the window
        if (I->l2_window_bytes == 0) cudaCtxResetPersistingL2Cache();
    }
    CUDA_OK(cudaStreamSetAttribute(I->stream, cudaStreamAttributeAccessPolicyWindow, &a));
    I->l2_window_stream = I->stream;
    I->l2_window_base   = base;
    I->l2_window_dirty  = false;
}

void render_views(octree_glc_t* rc, int n, float width, float height, const float* positions, const float* angles,
                  float lighta, uint8_t quality, int maxlevel, float basesize, int shoot)
{
    Impl* I = impl_of(rc);
    if (n <= 0) return;
    if (maxlevel < 0 || maxlevel > GENERIC_STACK - 1) die("maxlevel out of range (reference stack is 18 levels)");
    flush_pending(I);

    float ow, oh;
    int   W, H;
    render_size(width, height, quality, ow, oh, W, H);
    if (W <= 0 || H <= 0) return;
    const size_t pixels = (size_t) W * H;
    const bool   fenced = I->fence_n > 1;

    if (!I->ext_target) ensure_frame(I, pixels * n);
    if (I->aux_on && !I->ext_aux) ensure_aux(I, pixels * n);
    if (n > I->views_cap)
    {
        if (I->views_host)
        {
            CUDA_OK(cudaStreamSynchronize(I->stream));
            CUDA_OK(cudaFreeHost(I->views_host));
            CUDA_OK(cudaFree(I->views_dev));
        }
        I->views_cap = n < 64 ? 64 : n;
        CUDA_OK(cudaMallocHost(&I->views_host, VIEW_RING * I->views_cap * sizeof(ViewParams)));
        CUDA_OK(cudaMalloc(&I->views_dev, VIEW_RING * I->views_cap * sizeof(ViewParams)));
    }
    // per-frame constants travel through a small ring of pinned slots so that
    // queuing a frame never waits for the previous one
    const int slot = (int) (I->frame_seq++ % VIEW_RING);
    if (I->slot_used[slot]) CUDA_OK(cudaEventSynchronize(I->slot_ev[slot]));
    ViewParams* const vh = I->views_host + (size_t) slot * I->views_cap;
    ViewParams* const vd = I->views_dev + (size_t) slot * I->views_cap;

    for (int v = 0; v < n; v++)
    {
        ViewParams& V = vh[v];
        fill_view(I, V, ow, positions + 3 * v, angles + 3 * v, lighta, shoot);
        viewhost::fill_cfp(V, ow, oh);
    }
    CUDA_OK(cudaMemcpyAsync(vd, vh, n * sizeof(ViewParams), cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaEventRecord(I->slot_ev[slot], I->stream));
    I->slot_used[slot] = true;

    FrameParams P;
    memset(&P, 0, sizeof(P));
    for (int t = 0; t < 2; t++)
    {
        (t ? P.tree_d : P.tree_s) = tree_dev(I, t);
        PointsDev& Q = t ? P.pts_d : P.pts_s;
        Q.rec        = (const float4*) I->pts[t].rec.ptr;
        Q.points     = (int) I->pts[t].points;
    }
    P.basecube[0] = 0.0f; // octree_glc.c L263
    P.basecube[1] = basesize;
    P.basecube[2] = basesize;
    P.basecube[3] = basesize;
    P.maxlevel    = maxlevel;
    P.leaf_size   = ldexpf(basesize, -maxlevel);
    P.inv_leaf_size = 1.0f / P.leaf_size;
    P.W           = W;
    P.H           = H;
    P.sx          = ow / (float) W;
    P.sy          = oh / (float) H;
    if (I->ext_target)
    {
        P.frame       = (uchar4*) (uintptr_t) I->ext_target;
        P.pitch       = I->ext_pitch ? I->ext_pitch : (size_t) W;
        P.view_stride = P.pitch * H;
    }
    else
    {
        if (I->ring_on && !fenced)
        {
            I->ring_cur ^= 1;
            if (I->ring_cur == 1 && pixels * n > I->frame_alt_cap)
            {
                if (I->frame_alt)
                {
                    CUDA_OK(cudaStreamSynchronize(I->copy_stream));
                    CUDA_OK(cudaFree(I->frame_alt));
                    I->memsize -= I->frame_alt_cap * 4;
                }
                CUDA_OK(cudaMalloc(&I->frame_alt, pixels * n * 4));
                CUDA_OK(cudaMemsetAsync(I->frame_alt, 0, pixels * n * 4, I->stream)); // tiles of other ranks
                I->frame_alt_cap = pixels * n;
                I->memsize += I->frame_alt_cap * 4;
            }
            // the copy that last read this buffer must be done before the kernel overwrites it
            if (I->copy_pending[I->ring_cur]) CUDA_OK(cudaStreamWaitEvent(I->stream, I->ev_copy[I->ring_cur], 0));
        }
        P.frame       = (I->ring_on && !fenced && I->ring_cur == 1) ? I->frame_alt : I->frame;
        P.pitch       = W;
        P.view_stride = pixels;
    }
    P.flags    = I->ext_aux ? I->ext_flags : I->flags;
    P.aux      = I->ext_aux ? I->ext_aux : I->aux;
    P.counters = I->counters;
    if (fenced)
    {
        // completion fence of a frame split over several connectors (FenceDev, octree_types.cuh)
        P.fence.seq = ++I->fence_seq;
        P.fence.n   = I->fence_n;
        if (I->fence_rank == 0)
            P.fence.peers = I->fence_peers_dev;
        else
        {
            if (!I->ext_target) die("update: a fenced connector of rank > 0 renders into rank 0's frame (set_frame_target)");
            P.fence.done      = I->fence_peer0 + I->fence_rank;
            P.fence.gate      = I->fence_words + FENCE_CONSUMED;
        }
    }

    P.tile_w            = I->tile_w;
    P.tile_h            = I->tile_h;
    P.tiles_x           = (W + P.tile_w - 1) / P.tile_w;
    P.tiles_y           = (H + P.tile_h - 1) / P.tile_h;
    P.rank              = I->shard_rank;
    P.world             = I->shard_world;
    P.blocks_per_tile_x = P.tile_w / BLOCK_W;
    P.blocks_per_tile_y = P.tile_h / BLOCK_H;
    const int tiles     = P.tiles_x * P.tiles_y;
    if (tiles >= 65536 || P.blocks_per_tile_x * P.blocks_per_tile_y >= 65536) die("update: 65536 tiles or more in a frame");
    auto magic = [](int d) { return d <= 1 ? 0u : (unsigned) (((1ull << 32) + (unsigned) d - 1) / (unsigned) d); }; // 0: d = 1
    P.tiles_x_magic = magic(P.tiles_x);
    P.bptx_magic    = magic(P.blocks_per_tile_x);
    P.tiles_mine        = tiles > P.rank ? (tiles - P.rank + P.world - 1) / P.world : 0;
    P.views             = vd;
    P.n_views           = n;

    I->W       = W;
    I->H       = H;
    I->n_views = n;

    if (I->l2_window_bytes || I->l2_window_dirty) apply_l2_window(I);
    if (I->count_on) CUDA_OK(cudaMemsetAsync(I->counters, 0, CNT_COUNT * sizeof(unsigned long long), I->stream));
    // glClear(0,0,0,0) (octree_glc.c L289-290) needs no pass of its own: the
    // kernel stores every pixel of the tiles it owns, discarded ones as (0,0,0,0)

    const unsigned blocks = (unsigned) ((size_t) P.tiles_mine * P.blocks_per_tile_x * P.blocks_per_tile_y * n);
    std::function<void()> rank_tiles; // the tile-ranking launch of this frame, when it is queued after the completion
    // ranks != 0 of a fence: the frame number goes to rank 0 from a 1-warp kernel right behind the render kernel
    auto publish_done = [&]() {
        if (!fenced || I->fence_rank == 0) return;
        FenceDev F = P.fence;
        F.peers    = nullptr;
        fence_signal_kernel<<<1, 32, 0, I->stream>>>(F, 0);
        I->launches++;
    };
    CUDA_OK(cudaEventRecord(I->ev0, I->stream));
    bool ev1_recorded = false;
    if (blocks)
    {
        bool fast = I->kernel_choice == 2 || (I->kernel_choice == 0 && grid_is_exact(basesize, maxlevel));
        if (fast && !grid_is_exact(basesize, maxlevel)) die("fast kernel requested for a base cube that is not exact");
        int order_slot = -1;
        if (fast && n == 1 && I->feedback_on && P.tiles_mine > 1)
        {
            // the order learned for this exact view if it has been rendered before, else a copy of the most recent one
            const int nt = P.tiles_mine;
            if (nt > I->order_cap)
            {
                CUDA_OK(cudaStreamSynchronize(I->stream));
                if (I->order_dev) CUDA_OK(cudaFree(I->order_dev));
                for (int k = 0; k < 2; k++)
                    if (I->cost_dev[k]) CUDA_OK(cudaFree(I->cost_dev[k]));
                I->order_cap = nt + nt / 4 + 64;
                CUDA_OK(cudaMalloc(&I->order_dev, sizeof(int) * Impl::ORDER_SLOTS * I->order_cap));
                for (int k = 0; k < 2; k++)
                {
                    CUDA_OK(cudaMalloc(&I->cost_dev[k], sizeof(unsigned) * I->order_cap));
                    CUDA_OK(cudaMemsetAsync(I->cost_dev[k], 0, sizeof(unsigned) * I->order_cap, I->stream));
                }
                for (int k = 0; k < Impl::ORDER_SLOTS; k++) I->order_valid[k] = false;
                I->order_last = -1;
            }
            Impl::ViewKey key;
            memset(&key, 0, sizeof(key));
            memcpy(key.pos, positions, 12);
            memcpy(key.ang, angles, 12);
            key.width = width, key.height = height, key.lighta = lighta;
            key.quality = quality, key.maxlevel = maxlevel, key.tiles = nt;
            int lru = -1; // a free slot if there is one, else the least recently used
            for (int k = 0; k < Impl::ORDER_SLOTS; k++)
            {
                if (I->order_valid[k] && memcmp(&I->order_key[k], &key, sizeof(key)) == 0) order_slot = k;
                if (lru < 0)
                    lru = k;
                else if (I->order_valid[lru] && (!I->order_valid[k] || I->order_used[k] < I->order_used[lru]))
                    lru = k;
            }
            if (I->order_last >= 0 && I->order_key[I->order_last].tiles != nt)
            {
                // another tiling: measurements of the old one mean nothing here
                for (int k = 0; k < 2; k++)
                    CUDA_OK(cudaMemsetAsync(I->cost_dev[k], 0, sizeof(unsigned) * I->order_cap, I->stream));
            }
            if (order_slot < 0)
            {
                order_slot   = lru;
                int* dst     = I->order_dev + (size_t) order_slot * I->order_cap;
                const int ls = I->order_last;
                if (ls >= 0 && I->order_valid[ls] && I->order_key[ls].tiles == nt && ls != order_slot)
                    CUDA_OK(cudaMemcpyAsync(dst, I->order_dev + (size_t) ls * I->order_cap, sizeof(int) * nt,
                                            cudaMemcpyDeviceToDevice, I->stream));
                else if (!(ls == order_slot && I->order_valid[ls] && I->order_key[ls].tiles == nt))
                    tile_identity_kernel<<<(nt + 255) / 256, 256, 0, I->stream>>>(dst, nt);
                I->order_key[order_slot]   = key;
                I->order_valid[order_slot] = true;
            }
            I->order_used[order_slot] = ++I->order_clock;
            I->order_last             = order_slot;
            P.tile_order              = I->order_dev + (size_t) order_slot * I->order_cap;
            P.tile_cost               = I->cost_dev[I->cost_phase];
        }
        if (fast)
        {
            launch_fast(I, P, blocks);
            publish_done();
            if (order_slot >= 0)
            {
                // after the frame: this view's next order from the costs just measured (not part of ev0..ev1)
                CUDA_OK(cudaEventRecord(I->ev1, I->stream));
                ev1_recorded = true;
                const unsigned* cost_now  = I->cost_dev[I->cost_phase];
                unsigned*       cost_next = I->cost_dev[I->cost_phase ^ 1];
                int*            order_out = I->order_dev + (size_t) order_slot * I->order_cap;
                const int       nt        = P.tiles_mine;
                I->cost_phase ^= 1;
                rank_tiles = [=]() {
                    tile_rank_kernel<<<(nt * 32 + 255) / 256, 256, 0, I->stream>>>(cost_now, nt, order_out, cost_next);
                    I->launches++;
                };
                // the collecting rank of a fence ranks its tiles AFTER the wait for the other ranks' tiles (below):
                // the frame is complete a launch earlier
                if (!(fenced && I->fence_rank == 0))
                {
                    rank_tiles();
                    rank_tiles = nullptr;
                }
            }
        }
        else
        {
            launch_generic(I, P, blocks);
            publish_done();
        }
        CUDA_OK(cudaGetLastError());
        I->launches++;
        I->last_kernel = fast ? 2 : 1;
    }
    else if (fenced)
    {
        // no tile of this frame here: the fence signals still have to go out
        fence_signal_kernel<<<1, 64, 0, I->stream>>>(P.fence, 1);
        CUDA_OK(cudaGetLastError());
        I->launches++;
    }
    if (!ev1_recorded) CUDA_OK(cudaEventRecord(I->ev1, I->stream));
    // What follows the frame's own kernel on this stream: the wait for the other ranks' tiles and the presentation
    // pass.  The primary of an in-process group queues it only after every member's kernel has been launched
    // (group_render), so that nothing the host does in between -- an allocation that synchronises the device, when
    // members share a GPU -- can sit between the wait and the kernels it waits for.
    const uchar4* const done_frame = P.frame;
    const size_t        done_pitch = P.pitch;
    const unsigned      done_seq   = P.fence.seq;
    auto completion = [=]() {
        const int  ww = (int) width, wh = (int) height;
        const bool present = I->present_on && n == 1 && (I->shard_world == 1 || (fenced && I->fence_rank == 0)) &&
                             ww > 0 && wh > 0;
        // allocations first: nothing that may synchronise the device goes between the wait below and its peers
        if (present && (size_t) ww * wh > I->window_cap)
        {
            if (I->window)
            {
                CUDA_OK(cudaStreamSynchronize(I->stream));
                CUDA_OK(cudaFree(I->window));
                I->memsize -= I->window_cap * 4;
            }
            CUDA_OK(cudaMalloc(&I->window, (size_t) ww * wh * 4));
            I->window_cap = (size_t) ww * wh;
            I->memsize += I->window_cap * 4;
        }
        if (fenced && I->fence_rank == 0)
        {
            // the frame is complete once every rank has published its number (peer stores + release, no collective)
            fence_wait_done_kernel<<<1, 64, 0, I->stream>>>(I->fence_words, I->fence_n, done_seq);
            CUDA_OK(cudaGetLastError());
            I->launches++;
            CUDA_OK(cudaEventRecord(I->ev_done, I->stream));
        }
        if (rank_tiles) rank_tiles();
        // octree_glc.c L308-351; after ev1: not part of the frame time
        if (present)
        {
            PresentParams Q;
            Q.frame  = done_frame;
            Q.pitch  = done_pitch;
            Q.vp_w   = W;
            Q.vp_h   = H;
            Q.sx     = (double) ow / (double) ww;
            Q.sy     = (double) oh / (double) wh;
            Q.width  = ww;
            Q.height = wh;
            Q.window = I->window;
            const dim3 blk(32, 8), grd((ww + 31) / 32, (wh + 7) / 8);
            present_kernel<<<grd, blk, 0, I->stream>>>(Q);
            CUDA_OK(cudaGetLastError());
            I->launches++;
            I->window_w = ww;
            I->window_h = wh;
        }
        if (I->ring_on && !fenced && !I->ext_target) CUDA_OK(cudaEventRecord(I->ev_render[I->ring_cur], I->stream));
    };
    if (I->defer_completion)
        I->deferred = completion;
    else
        completion();
    I->timed = true;
    publish_memsize(rc, I);
}

// One frame (or batch of views) on every device of an in-process group: the primary's framebuffer is the target of
// all of them, each connector renders its tiles (octree_cuc_set_gpus has set the shards) and the device-side fence
// completes the frame on the primary's stream.
void group_render(octree_glc_t* rc, int n, float width, float height, const float* positions, const float* angles,
                  float lighta, uint8_t quality, int maxlevel, float basesize, int shoot)
{
    Impl* I = impl_of(rc);
    if (I->replicas.empty())
    {
        render_views(rc, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot);
        return;
    }
    if (n <= 0) return;
    float ow, oh;
    int   W, H;
    render_size(width, height, quality, ow, oh, W, H);
    if (W <= 0 || H <= 0) return;
    if (I->ext_target) die("update: an external frame target cannot be combined with octree_cuc_set_gpus");
    const size_t total = (size_t) W * H * n;
    ensure_frame(I, total);
    if (I->aux_on) ensure_aux(I, total);
    for (auto& r : I->replicas)
    {
        Impl* R       = (Impl*) r.impl;
        R->ext_target = (uint64_t) (uintptr_t) I->frame;
        R->ext_pitch  = (size_t) W;
        R->ext_flags  = I->aux_on ? I->flags : nullptr;
        R->ext_aux    = I->aux_on ? I->aux : nullptr;
    }
    I->defer_completion = true;
    if (!I->workers.empty())
    {
        // every member on a device of its own: all launches side by side (the order between devices is free,
        // the gate in the kernels makes the others' stores wait for the primary's kernel)
        for (size_t k = 0; k < I->replicas.size(); k++)
        {
            octree_glc_t* m = &I->replicas[k];
            I->workers[k]->post([=]() {
                render_views(m, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot);
            });
        }
        render_views(rc, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot);
        for (auto& w : I->workers) w->wait();
        I->defer_completion = false;
        CUDA_OK(cudaSetDevice(I->device));
    }
    else
    {
        // members share a device: the primary first, its kernel carries the "previous frame consumed" signal the
        // others' stores wait for
        render_views(rc, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot);
        I->defer_completion = false;
        REPLAY(I, render_views(m, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot));
    }
    if (I->deferred)
    {
        I->deferred();
        I->deferred = nullptr;
    }
    // device memory of the whole group
    uint64_t mem = I->memsize;
    for (auto& r : I->replicas) mem += ((Impl*) r.impl)->memsize;
    rc->memsize_bytes = mem;
    rc->memsize       = mem > 0xffffffffull ? 0xffffffffu : (unsigned int) mem;
}

unsigned* fence_alloc(Impl* I)
{
    if (!I->fence_words)
    {
        CUDA_OK(cudaMalloc(&I->fence_words, FENCE_WORDS * sizeof(unsigned)));
        CUDA_OK(cudaMemset(I->fence_words, 0, FENCE_WORDS * sizeof(unsigned)));
        CUDA_OK(cudaMalloc(&I->fence_peers_dev, 64 * sizeof(unsigned*)));
        CUDA_OK(cudaMemset(I->fence_peers_dev, 0, 64 * sizeof(unsigned*)));
        CUDA_OK(cudaEventCreate(&I->ev_done));
    }
    return I->fence_words;
}

void fence_wire(Impl* I, int rank, int world, const uint64_t* ptrs)
{
    CUDA_OK(cudaSetDevice(I->device));
    flush_pending(I);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    if (world <= 1 || ptrs == nullptr)
    {
        I->fence_rank = 0, I->fence_n = 1, I->fence_seq = 0;
        return;
    }
    if (world > 64 || rank < 0 || rank >= world) die("set_fence: need 0 <= rank < world <= 64");
    fence_alloc(I);
    if ((uint64_t) (uintptr_t) I->fence_words != ptrs[rank]) die("set_fence: entry `rank` must be this connector's own fence words");
    CUDA_OK(cudaMemset(I->fence_words, 0, FENCE_WORDS * sizeof(unsigned)));
    unsigned* table[64] = {};
    for (int k = 0; k < world; k++) table[k] = (unsigned*) (uintptr_t) ptrs[k];
    CUDA_OK(cudaMemcpy(I->fence_peers_dev, table, sizeof(table), cudaMemcpyHostToDevice));
    I->fence_peer0 = table[0];
    I->fence_rank  = rank;
    I->fence_n     = world;
    I->fence_seq   = 0;
}

} // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------

extern "C" {

const char* octree_cuc_version(void) { return "octree_cuc 0.1 (sm_100a)"; }

void octree_cuc_select_device(int device) { g_selected_device = device; }

octree_glc_t octree_glc_init(char* path)
{
    (void) path; // shader directory of the GL connector: kernels are built in
    octree_glc_t rc;
    memset(&rc, 0, sizeof(rc));

    int         ndev = 0;
    cudaError_t e    = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        fprintf(stderr, "octree_cuc: no usable CUDA device (%s); this connector has no CPU path\n",
                cudaGetErrorString(e));
        abort();
    }
    Impl* I = new Impl();
    if (g_selected_device >= 0)
        I->device = g_selected_device;
    else
        CUDA_OK(cudaGetDevice(&I->device));
    CUDA_OK(cudaSetDevice(I->device));
    {
        // scratch of the tree build comes from the stream-ordered pool: keep freed blocks cached across
        // synchronisations (the default threshold of 0 returns them to the driver every time)
        cudaMemPool_t pool;
        CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, I->device));
        uint64_t keep = ~0ull;
        CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CUDA_OK(cudaStreamCreateWithFlags(&I->own_stream, cudaStreamNonBlocking));
    I->stream = I->own_stream;
    if (getenv("QB_TILE_FEEDBACK")) I->feedback_on = atoi(getenv("QB_TILE_FEEDBACK")) != 0; // for A/B runs
    if (getenv("QB_CTA_CAP")) I->cta_cap = atoi(getenv("QB_CTA_CAP"));
    if (getenv("QB_UPLOAD_THREADS")) I->upload_threads = atoi(getenv("QB_UPLOAD_THREADS")); // for A/B runs
    if (getenv("QB_L2_WINDOW_MB")) I->l2_window_bytes = (size_t) atoi(getenv("QB_L2_WINDOW_MB")) << 20, I->l2_window_dirty = true;
    CUDA_OK(cudaEventCreate(&I->ev0));
    CUDA_OK(cudaEventCreate(&I->ev1));
    for (int i = 0; i < VIEW_RING; i++) CUDA_OK(cudaEventCreateWithFlags(&I->slot_ev[i], cudaEventDisableTiming));
    CUDA_OK(cudaMalloc(&I->stage_dev, STAGE_BYTES));
    CUDA_OK(cudaMalloc(&I->batch_dev, BATCH_BYTES));
    CUDA_OK(cudaMalloc(&I->desc_dev, BATCH_MAXDESC * sizeof(RangeDesc)));
    CUDA_OK(cudaMallocHost(&I->batch_host, BATCH_BYTES));
    CUDA_OK(cudaMallocHost(&I->desc_host, BATCH_MAXDESC * sizeof(RangeDesc)));
    CUDA_OK(cudaMalloc(&I->counters, CNT_COUNT * sizeof(unsigned long long)));
    CUDA_OK(cudaMemset(I->counters, 0, CNT_COUNT * sizeof(unsigned long long)));
    I->memsize = STAGE_BYTES + BATCH_BYTES + BATCH_MAXDESC * sizeof(RangeDesc);
    // every tree has a root (octree.c L63): node 0 reads as "no children" until uploaded
    for (int t = 0; t < 2; t++)
    {
        ensure_capacity(I, t ? OCTREE_GLC_BUFFER_DYNAMIC_OCTREE : OCTREE_GLC_BUFFER_STATIC_OCTREE, 48);
        ensure_capacity(I, t ? OCTREE_GLC_BUFFER_DYNAMIC_COLOR : OCTREE_GLC_BUFFER_STATIC_COLOR, 12);
    }
    CUDA_OK(cudaStreamSynchronize(I->stream));
    rc.impl = I;
    publish_memsize(&rc, I);
    return rc;
}

void octree_cuc_destroy(octree_glc_t* rc)
{
    if (!rc || !rc->impl) return;
    Impl* I = impl_of(rc);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    I->workers.clear(); // joins the frame workers
    for (auto& r : I->replicas) octree_cuc_destroy(&r);
    I->replicas.clear();
    CUDA_OK(cudaSetDevice(I->device));
    I->repl_payload.release();
    if (I->fence_words)
    {
        cudaFree(I->fence_words);
        cudaFree(I->fence_peers_dev);
        cudaEventDestroy(I->ev_done);
    }
    for (int t = 0; t < 2; t++)
    {
        dev_free(I, I->tree[t].child);
        dev_free(I, I->tree[t].model);
        dev_free(I, I->tree[t].slot);
        dev_free(I, I->tree[t].parent);
        dev_free(I, I->pts[t].rec);
    }
    if (I->stager) I->stager->shutdown();
    I->stager.reset();
    cudaFree(I->stage_dev);
    cudaFree(I->batch_dev);
    cudaFree(I->desc_dev);
    cudaFreeHost(I->batch_host);
    cudaFreeHost(I->desc_host);
    cudaFree(I->counters);
    if (I->frame) cudaFree(I->frame);
    if (I->frame_alt) cudaFree(I->frame_alt);
    if (I->window) cudaFree(I->window);
    if (I->order_dev) cudaFree(I->order_dev);
    for (int k = 0; k < 2; k++)
        if (I->cost_dev[k]) cudaFree(I->cost_dev[k]);
    for (void* p : {(void*) I->skin_pos, (void*) I->skin_nrm, (void*) I->skin_p14, (void*) I->skin_p54,
                    (void*) I->skin_p94, (void*) I->skin_pnt_out, (void*) I->part_finished})
        if (p) cudaFree(p);
    for (int k = 0; k < 2; k++)
        for (int b = 0; b < 2; b++)
        {
            if (I->part_pos[k][b]) cudaFree(I->part_pos[k][b]);
            if (I->part_spd[k][b]) cudaFree(I->part_spd[k][b]);
        }
    if (I->copy_stream) cudaStreamSynchronize(I->copy_stream);
    if (I->ring_on)
        for (int i = 0; i < 2; i++)
        {
            cudaEventDestroy(I->ev_render[i]);
            cudaEventDestroy(I->ev_copy[i]);
        }
    if (I->stage_on)
        for (int i = 0; i < 2; i++)
        {
            cudaEventDestroy(I->ev_stage_ready[i]);
            cudaEventDestroy(I->ev_stage_done[i]);
            if (I->stage_frame[i]) cudaFree(I->stage_frame[i]);
        }
    if (I->copy_stream) cudaStreamDestroy(I->copy_stream);
    if (I->flags) cudaFree(I->flags);
    if (I->aux) cudaFree(I->aux);
    if (I->views_host) cudaFreeHost(I->views_host);
    if (I->views_dev) cudaFree(I->views_dev);
    for (int i = 0; i < VIEW_RING; i++) cudaEventDestroy(I->slot_ev[i]);
    cudaEventDestroy(I->ev0);
    cudaEventDestroy(I->ev1);
    cudaStreamDestroy(I->own_stream);
    delete I;
    rc->impl          = nullptr;
    rc->memsize       = 0;
    rc->memsize_bytes = 0;
}

} // extern "C"
namespace
{
void upload_one(octree_glc_t* rc, void* data, int type, size_t size, size_t itemsize, size_t start, size_t end,
                octree_glc_buffer_t buftype)
{
    Impl* I = impl_of(rc);
    struct Timer
    {
        Impl*                                 I;
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~Timer() { I->upload_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    } timer{I};
    if ((int) buftype < 0 || (int) buftype > 5) die("upload: unknown buffer type");
    if (itemsize == 0 || data == nullptr) return;
    const bool oct = is_octree(buftype);
    if (oct && (type != OCTREE_CUC_GL_INT || itemsize != 16)) die("upload: octree buffers take GL_INT items of 16 bytes");
    if (!oct && (type != OCTREE_CUC_GL_FLOAT || itemsize != 12))
        die("upload: colour/normal buffers take GL_FLOAT items of 12 bytes");

    // octree_glc.c L412-428: growth re-uploads the whole array
    if (ensure_capacity(I, buftype, size))
    {
        start = 0;
        end   = size;
    }
    // octree_glc.c L432-433: texel granularity, rounded down
    size_t s = (start / itemsize) * itemsize;
    size_t e = (end / itemsize) * itemsize;
    if (e > (size / itemsize) * itemsize) e = (size / itemsize) * itemsize;
    if (e > s)
    {
        apply_range(I, (const char*) data + s, buftype, s, e);
        if (I->repl_on) // for the connectors of other processes (octree_cuc_export_pending)
        {
            RangeDesc d;
            d.dst_word = s / 4;
            d.src_word = (unsigned int) (I->repl_payload.size() / 4);
            d.nwords   = (unsigned int) ((e - s) / 4);
            d.buftype  = buftype;
            d.pad      = 0;
            if (I->repl_payload.size() + (e - s) >= ((size_t) 1 << 34)) die("replication log exceeds 16 GiB: export it");
            I->repl_descs.push_back(d);
            I->repl_payload.append((const char*) data + s, e - s);
        }
    }
    publish_memsize(rc, I);
}
} // namespace
extern "C" {

void octree_glc_upload_texbuffer_data(octree_glc_t* rc, void* data, int type, size_t size, size_t itemsize,
                                      size_t start, size_t end, octree_glc_buffer_t buftype)
{
    Impl* I = impl_of(rc);
    if (I->replicas.empty())
    {
        upload_one(rc, data, type, size, itemsize, start, end, buftype);
        return;
    }
    // every device of the group takes the same call (same state, so the same growth decisions).  The thousands of
    // 48-byte range uploads of a shot are a memcpy into pinned staging each: one after the other; a bulk range or a
    // growing array waits for its device: all devices at once.
    bool heavy = end > start && end - start > SMALL_RANGE;
    if ((int) buftype >= 0 && (int) buftype <= 5)
    {
        const int t = tree_index(buftype);
        heavy = heavy || (is_octree(buftype) ? (size + 47) / 48 > I->tree[t].cap_nodes : (size + 11) / 12 > I->pts[t].cap_points);
    }
    if (heavy)
        run_on_group(rc, [=](octree_glc_t* m, bool) { upload_one(m, data, type, size, itemsize, start, end, buftype); });
    else
    {
        upload_one(rc, data, type, size, itemsize, start, end, buftype);
        REPLAY(I, upload_one(m, data, type, size, itemsize, start, end, buftype));
    }
}

void octree_glc_update(octree_glc_t* rc, float width, float height, v3_t position, v3_t angle, float lighta,
                       uint8_t quality, int maxlevel, float basesize, int shoot)
{
    const float pos[3] = {position.x, position.y, position.z};
    const float ang[3] = {angle.x, angle.y, angle.z};
    group_render(rc, 1, width, height, pos, ang, lighta, quality, maxlevel, basesize, shoot);
}

void octree_cuc_update_views(octree_glc_t* rc, int n, float width, float height, const float* positions,
                             const float* angles, float lighta, uint8_t quality, int maxlevel, float basesize,
                             int shoot)
{
    group_render(rc, n, width, height, positions, angles, lighta, quality, maxlevel, basesize, shoot);
}

void octree_cuc_sync(octree_glc_t* rc)
{
    Impl* I = impl_of(rc);
    flush_pending(I);
    REPLAY(I, octree_cuc_sync(m)); // the primary's stream waits for the others' frames: they finish first
    CUDA_OK(cudaStreamSynchronize(I->stream));
}

void octree_cuc_frame_size(octree_glc_t* rc, int* width, int* height)
{
    Impl* I = impl_of(rc);
    if (width) *width = I->W;
    if (height) *height = I->H;
}

size_t octree_cuc_read_frame(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity)
{
    Impl*  I     = impl_of(rc);
    size_t bytes = (size_t) I->W * I->H * 4 * (I->n_views ? I->n_views : 1);
    if (I->ext_target) die("read_frame: frame target is external, read it there");
    if (bytes == 0 || capacity < bytes) return 0;
    const uchar4* src = (I->ring_on && I->ring_cur == 1) ? I->frame_alt : I->frame;
    CUDA_OK(cudaMemcpyAsync(rgba_host, src, bytes, cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return bytes;
}

void octree_cuc_enable_present(octree_glc_t* rc, int on) { impl_of(rc)->present_on = on != 0; }

void octree_cuc_set_tile_feedback(octree_glc_t* rc, int on)
{
    Impl* I        = impl_of(rc);
    I->feedback_on = on != 0;
    REPLAY(I, octree_cuc_set_tile_feedback(m, on));
}

// L2 persisting access-policy window over the head of the static tree's node array (north_star design point 1 as an
// A/B switch): `persist_bytes` of L2 are set aside (clamped to the device maximum) and the window's lines are marked
// persisting with the hit ratio that fits them; 0 removes the window.  Applied to the stream the frames run on.
void octree_cuc_set_error_handler(void (*handler)(const char* message, void* user), void* user)
{
    g_error_fn   = handler;
    g_error_user = user;
}

void octree_cuc_set_occupancy(octree_glc_t* rc, int ctas_per_sm)
{
    Impl* I    = impl_of(rc);
    I->cta_cap = ctas_per_sm;
    REPLAY(I, octree_cuc_set_occupancy(m, ctas_per_sm));
}

void octree_cuc_set_persisting_window(octree_glc_t* rc, size_t persist_bytes)
{
    Impl* I             = impl_of(rc);
    I->l2_window_bytes  = persist_bytes;
    I->l2_window_stream = nullptr; // re-apply at the next frame
    I->l2_window_dirty  = true;
    REPLAY(I, octree_cuc_set_persisting_window(m, persist_bytes));
}

size_t octree_cuc_read_window(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity, int* width, int* height)
{
    Impl*        I     = impl_of(rc);
    const size_t bytes = (size_t) I->window_w * I->window_h * 4;
    if (width) *width = I->window_w;
    if (height) *height = I->window_h;
    if (!rgba_host || bytes == 0 || capacity < bytes) return 0;
    CUDA_OK(cudaMemcpyAsync(rgba_host, I->window, bytes, cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return bytes;
}

uint64_t octree_cuc_window_device(octree_glc_t* rc) { return (uint64_t) (uintptr_t) impl_of(rc)->window; }

uint64_t octree_cuc_frame_device(octree_glc_t* rc)
{
    Impl* I = impl_of(rc);
    if (I->ext_target) return I->ext_target;
    return (uint64_t) (uintptr_t) ((I->ring_on && I->ring_cur == 1) ? I->frame_alt : I->frame);
}

size_t octree_cuc_read_frame_async(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity)
{
    Impl*  I     = impl_of(rc);
    size_t bytes = (size_t) I->W * I->H * 4 * (I->n_views ? I->n_views : 1);
    if (I->ext_target) die("read_frame_async: frame target is external, read it there");
    if (bytes == 0 || capacity < bytes) return 0;
    // a frame other connectors store into stays where it is: snapshot, then copy (octree_cuc_read_frame_staged)
    if (I->fence_n > 1) return octree_cuc_read_frame_staged(rc, rgba_host, capacity);
    if (!I->ring_on)
    {
        // first use: from now on frames alternate between two buffers; the frame just rendered is in `frame`
        if (!I->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&I->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++)
        {
            CUDA_OK(cudaEventCreateWithFlags(&I->ev_render[i], cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&I->ev_copy[i], cudaEventDisableTiming));
        }
        I->ring_on  = true;
        I->ring_cur = 0;
        CUDA_OK(cudaEventRecord(I->ev_render[0], I->stream));
    }
    const int     k   = I->ring_cur;
    const uchar4* src = k == 1 ? I->frame_alt : I->frame;
    CUDA_OK(cudaStreamWaitEvent(I->copy_stream, I->ev_render[k], 0));
    CUDA_OK(cudaMemcpyAsync(rgba_host, src, bytes, cudaMemcpyDeviceToHost, I->copy_stream));
    CUDA_OK(cudaEventRecord(I->ev_copy[k], I->copy_stream));
    I->copy_pending[k] = true;
    return bytes;
}

size_t octree_cuc_read_frame_staged(octree_glc_t* rc, uint8_t* rgba_host, size_t capacity)
{
    Impl*        I      = impl_of(rc);
    const size_t pixels = (size_t) I->W * I->H * (I->n_views ? I->n_views : 1);
    const size_t bytes  = pixels * 4;
    if (I->ext_target && I->ext_pitch && I->ext_pitch != (size_t) I->W) die("read_frame_staged: pitched external target");
    if (bytes == 0 || capacity < bytes) return 0;
    if (!I->stage_on)
    {
        if (!I->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&I->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++)
        {
            CUDA_OK(cudaEventCreateWithFlags(&I->ev_stage_ready[i], cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&I->ev_stage_done[i], cudaEventDisableTiming));
        }
        I->stage_on = true;
    }
    if (pixels > I->stage_cap)
    {
        CUDA_OK(cudaStreamSynchronize(I->copy_stream));
        CUDA_OK(cudaStreamSynchronize(I->stream));
        for (int i = 0; i < 2; i++)
        {
            if (I->stage_frame[i]) CUDA_OK(cudaFree(I->stage_frame[i]));
            CUDA_OK(cudaMalloc(&I->stage_frame[i], bytes));
            I->stage_pending[i] = false;
        }
        I->memsize += 2 * (pixels - I->stage_cap) * 4;
        I->stage_cap = pixels;
        publish_memsize(rc, I);
    }
    const int   k   = I->stage_k;
    I->stage_k ^= 1;
    const void* src = (const void*) (uintptr_t) octree_cuc_frame_device(rc);
    // the host copy that last read this staging buffer must be done before it is overwritten
    if (I->stage_pending[k]) CUDA_OK(cudaStreamWaitEvent(I->stream, I->ev_stage_done[k], 0));
    CUDA_OK(cudaMemcpyAsync(I->stage_frame[k], src, bytes, cudaMemcpyDeviceToDevice, I->stream));
    CUDA_OK(cudaEventRecord(I->ev_stage_ready[k], I->stream));
    CUDA_OK(cudaStreamWaitEvent(I->copy_stream, I->ev_stage_ready[k], 0));
    CUDA_OK(cudaMemcpyAsync(rgba_host, I->stage_frame[k], bytes, cudaMemcpyDeviceToHost, I->copy_stream));
    CUDA_OK(cudaEventRecord(I->ev_stage_done[k], I->copy_stream));
    I->stage_pending[k] = true;
    return bytes;
}

void octree_cuc_wait_reads(octree_glc_t* rc)
{
    Impl* I = impl_of(rc);
    if (I->copy_stream) CUDA_OK(cudaStreamSynchronize(I->copy_stream));
}

void octree_cuc_set_frame_target(octree_glc_t* rc, uint64_t device_ptr, size_t pitch_pixels)
{
    Impl* I       = impl_of(rc);
    I->ext_target = device_ptr;
    I->ext_pitch  = pitch_pixels;
}

void octree_cuc_enable_aux(octree_glc_t* rc, int enable)
{
    Impl* I   = impl_of(rc);
    I->aux_on = enable != 0;
    REPLAY(I, octree_cuc_enable_aux(m, enable));
}

size_t octree_cuc_read_aux(octree_glc_t* rc, uint8_t* flags_host, int32_t* aux_host)
{
    Impl*  I      = impl_of(rc);
    size_t pixels = (size_t) I->W * I->H * (I->n_views ? I->n_views : 1);
    if (!I->aux_on || !I->flags || pixels == 0) return 0;
    if (flags_host) CUDA_OK(cudaMemcpyAsync(flags_host, I->flags, pixels, cudaMemcpyDeviceToHost, I->stream));
    if (aux_host)
        CUDA_OK(cudaMemcpyAsync(aux_host, I->aux, pixels * 6 * sizeof(int), cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return pixels;
}

void octree_cuc_enable_counters(octree_glc_t* rc, int enable)
{
    Impl* I     = impl_of(rc);
    I->count_on = enable != 0;
    REPLAY(I, octree_cuc_enable_counters(m, enable));
}

void octree_cuc_read_counters(octree_glc_t* rc, octree_cuc_counters* out)
{
    Impl*              I = impl_of(rc);
    unsigned long long h[CNT_COUNT];
    CUDA_OK(cudaMemcpyAsync(h, I->counters, sizeof(h), cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    out->rays_primary = (int64_t) h[CNT_RAYS_PRIMARY];
    out->rays_shadow  = (int64_t) h[CNT_RAYS_SHADOW];
    out->rays_disc    = (int64_t) h[CNT_RAYS_DISC];
    out->expand_s     = (int64_t) h[CNT_EXPAND_S];
    out->expand_d     = (int64_t) h[CNT_EXPAND_D];
    out->leaf_s       = (int64_t) h[CNT_LEAF_S];
    out->leaf_d       = (int64_t) h[CNT_LEAF_D];
    out->hits         = (int64_t) h[CNT_HITS];
    out->discards     = (int64_t) h[CNT_DISCARDS];
    out->descents     = (int64_t) h[CNT_DESCENTS];
    for (auto& r : I->replicas) // a group's counters are the sums over its devices
    {
        octree_cuc_counters c;
        octree_cuc_read_counters(&r, &c);
        int64_t*       a = &out->rays_primary;
        const int64_t* b = &c.rays_primary;
        for (int i = 0; i < CNT_COUNT; i++) a[i] += b[i];
    }
    if (!I->replicas.empty()) CUDA_OK(cudaSetDevice(I->device));
}

void octree_cuc_set_shard(octree_glc_t* rc, int rank, int world, int tile_w, int tile_h)
{
    Impl* I = impl_of(rc);
    if (!I->replicas.empty()) die("set_shard: the shards of an octree_cuc_set_gpus group are managed by the connector");
    if (world < 1 || rank < 0 || rank >= world) die("set_shard: need 0 <= rank < world");
    if (tile_w <= 0 || tile_h <= 0 || tile_w % BLOCK_W || tile_h % BLOCK_H)
        die("set_shard: tile size must be a multiple of 16 x 8 pixels");
    I->shard_rank  = rank;
    I->shard_world = world;
    I->tile_w      = tile_w;
    I->tile_h      = tile_h;
}

void octree_cuc_set_light(octree_glc_t* rc, const float* light)
{
    Impl* I           = impl_of(rc);
    I->light_override = light != nullptr;
    if (light) memcpy(I->light, light, sizeof(I->light));
    REPLAY(I, octree_cuc_set_light(m, light));
}

void octree_cuc_set_kernel(octree_glc_t* rc, int which)
{
    if (which < 0 || which > 2) die("set_kernel: 0 auto, 1 generic, 2 fast");
    Impl* I          = impl_of(rc);
    I->kernel_choice = which;
    REPLAY(I, octree_cuc_set_kernel(m, which));
}

int octree_cuc_last_kernel(octree_glc_t* rc) { return impl_of(rc)->last_kernel; }

void octree_cuc_set_division(octree_glc_t* rc, int mode)
{
    if (mode != DIV_GLSL && mode != DIV_IEEE) die("set_division: 0 = GLSL a*(1/b), 1 = IEEE a/b");
    Impl* I     = impl_of(rc);
    I->div_mode = mode;
    REPLAY(I, octree_cuc_set_division(m, mode));
}

float octree_cuc_last_frame_ms(octree_glc_t* rc)
{
    Impl* I = impl_of(rc);
    if (!I->timed) return 0.0f;
    CUDA_OK(cudaEventSynchronize(I->ev1));
    float ms = 0.0f;
    CUDA_OK(cudaEventElapsedTime(&ms, I->ev0, I->ev1));
    for (auto& r : I->replicas) // a group's kernel time is its slowest device's
    {
        const float v = octree_cuc_last_frame_ms(&r);
        if (v > ms) ms = v;
    }
    if (!I->replicas.empty()) CUDA_OK(cudaSetDevice(I->device));
    return ms;
}

float octree_cuc_last_step_ms(octree_glc_t* rc)
{
    Impl* I = impl_of(rc);
    if (!I->timed) return 0.0f;
    if (!(I->fence_n > 1 && I->fence_rank == 0)) return octree_cuc_last_frame_ms(rc);
    CUDA_OK(cudaEventSynchronize(I->ev_done));
    float ms = 0.0f;
    CUDA_OK(cudaEventElapsedTime(&ms, I->ev0, I->ev_done));
    return ms;
}

uint64_t octree_cuc_launch_count(octree_glc_t* rc)
{
    Impl*    I = impl_of(rc);
    uint64_t n = I->launches;
    for (auto& r : I->replicas) n += ((Impl*) r.impl)->launches;
    return n;
}

void octree_cuc_set_stream(octree_glc_t* rc, uint64_t cuda_stream)
{
    Impl* I = impl_of(rc);
    flush_pending(I);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    I->stream = cuda_stream ? (cudaStream_t) (uintptr_t) cuda_stream : I->own_stream;
}

void octree_cuc_reserve_frame(octree_glc_t* rc, int width, int height, int views)
{
    Impl*  I      = impl_of(rc);
    size_t pixels = (size_t) width * height * (views > 0 ? views : 1);
    ensure_frame(I, pixels);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    publish_memsize(rc, I);
}

void octree_cuc_ipc_export_frame(octree_glc_t* rc, uint8_t* handle64)
{
    Impl* I = impl_of(rc);
    if (!I->frame) die("ipc_export_frame: reserve or render a frame first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, I->frame));
    memcpy(handle64, &h, 64);
}

uint64_t octree_cuc_ipc_open(octree_glc_t* rc, const uint8_t* handle64)
{
    impl_of(rc);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return (uint64_t) (uintptr_t) p;
}

void octree_cuc_ipc_close(octree_glc_t* rc, uint64_t device_ptr)
{
    Impl* I = impl_of(rc);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    if (I->ext_target == device_ptr) I->ext_target = 0;
    CUDA_OK(cudaIpcCloseMemHandle((void*) (uintptr_t) device_ptr));
}

void octree_cuc_ipc_export_ptr(octree_glc_t* rc, uint64_t device_ptr, uint8_t* handle64)
{
    impl_of(rc);
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, (void*) (uintptr_t) device_ptr));
    memcpy(handle64, &h, 64);
}

uint64_t octree_cuc_fence_device(octree_glc_t* rc) { return (uint64_t) (uintptr_t) fence_alloc(impl_of(rc)); }

void octree_cuc_set_fence(octree_glc_t* rc, int rank, int world, const uint64_t* fence_ptrs)
{
    Impl* I = impl_of(rc);
    if (!I->replicas.empty() || I->is_replica) die("set_fence: the fence of an octree_cuc_set_gpus group is managed by the connector");
    fence_wire(I, rank, world, fence_ptrs);
}

int octree_cuc_gpu_count(octree_glc_t* rc) { return 1 + (int) impl_of(rc)->replicas.size(); }

void octree_cuc_set_gpus(octree_glc_t* rc, int n, const int* devices)
{
    Impl* I = impl_of(rc);
    if (I->is_replica) die("set_gpus: not on a member of a group");
    if (!I->replicas.empty() || I->fence_n > 1) die("set_gpus: the group is already set up");
    if (n < 1 || n > 64) die("set_gpus: 1 <= n <= 64");
    for (int t = 0; t < 2; t++)
        if (I->tree[t].nodes || I->pts[t].points || !I->descs.empty())
            die("set_gpus: call it right after octree_glc_init, before the first upload");
    if (n == 1) return;
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (devices && devices[0] != I->device) die("set_gpus: devices[0] must be the device octree_glc_init chose");
    if (!devices && n > ndev) die("set_gpus: more GPUs asked for than the box has");
    uint64_t words[64];
    words[0] = (uint64_t) (uintptr_t) fence_alloc(I);
    const int saved = g_selected_device;
    for (int k = 1; k < n; k++)
    {
        const int dev = devices ? devices[k] : (I->device + k) % ndev;
        if (dev < 0 || dev >= ndev) die("set_gpus: no such device");
        if (dev != I->device)
        {
            // the render kernels of device `dev` store into the primary's framebuffer and fence words, and the
            // primary's kernel into theirs
            int can = 0;
            CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, I->device));
            if (!can) die("set_gpus: the devices cannot access each other's memory (no NVLink / PCIe peer path)");
            for (int dir = 0; dir < 2; dir++)
            {
                CUDA_OK(cudaSetDevice(dir ? I->device : dev));
                cudaError_t e = cudaDeviceEnablePeerAccess(dir ? dev : I->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    (void) cudaGetLastError();
                else
                    CUDA_OK(e);
            }
        }
        g_selected_device = dev;
        octree_glc_t r    = octree_glc_init(nullptr);
        Impl*        R    = (Impl*) r.impl;
        R->is_replica     = true;
        R->div_mode       = I->div_mode;
        R->kernel_choice  = I->kernel_choice;
        R->feedback_on    = I->feedback_on;
        R->aux_on         = I->aux_on;
        R->count_on       = I->count_on;
        R->light_override = I->light_override;
        memcpy(R->light, I->light, sizeof(I->light));
        words[k] = (uint64_t) (uintptr_t) fence_alloc(R);
        I->replicas.push_back(r);
    }
    g_selected_device = saved;
    // image tiles `mod n`, all stored into the primary's framebuffer, completed by the device-side fence
    for (int k = 0; k < n; k++)
    {
        Impl* M        = k ? (Impl*) I->replicas[k - 1].impl : I;
        M->shard_rank  = k;
        M->shard_world = n;
        fence_wire(M, k, n, words);
    }
    // frame workers when no two members share a device (shards of one GPU -- the test mode -- launch in order)
    bool distinct = true;
    for (int a = 0; a < n && distinct; a++)
        for (int b = a + 1; b < n; b++)
        {
            const int da = a ? ((Impl*) I->replicas[a - 1].impl)->device : I->device;
            const int db = ((Impl*) I->replicas[b - 1].impl)->device;
            if (da == db) distinct = false;
        }
    if (distinct && !getenv("QB_NO_FRAME_WORKERS"))
        for (int k = 1; k < n; k++) I->workers.emplace_back(new FrameWorker());
    CUDA_OK(cudaSetDevice(I->device));
}

void octree_cuc_set_upload_threads(octree_glc_t* rc, int threads)
{
    Impl* I = impl_of(rc);
    if (threads < 0 || threads > 64) die("set_upload_threads: 0..64");
    if (threads == I->upload_threads) return;
    CUDA_OK(cudaStreamSynchronize(I->stream));
    if (I->stager) I->stager->shutdown();
    I->stager.reset();
    I->upload_threads = threads;
}

void octree_cuc_pin_host_buffer(octree_glc_t* rc, void* data, size_t bytes)
{
    impl_of(rc);
    CUDA_OK(cudaHostRegister(data, bytes, cudaHostRegisterPortable)); // page-locked for every device of a group
}

void octree_cuc_unpin_host_buffer(octree_glc_t* rc, void* data)
{
    Impl* I = impl_of(rc);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    CUDA_OK(cudaHostUnregister(data));
}

double octree_cuc_take_upload_ms(octree_glc_t* rc)
{
    Impl*  I  = impl_of(rc);
    double ms = I->upload_ms;
    I->upload_ms = 0.0;
    return ms;
}

// ---------------------------------------------------------------------------
// GPU tree build from octant paths (octree_build.cuh) and export in the reference format
// ---------------------------------------------------------------------------
} // extern "C"

namespace
{
template <class T>
T* scratch(Impl* I, size_t count)
{
    void* p = nullptr;
    CUDA_OK(cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), I->stream));
    return (T*) p;
}
void            scratch_free(Impl* I, void* p) { CUDA_OK(cudaFreeAsync(p, I->stream)); }
inline unsigned nblk(size_t n) { return (unsigned) ((n + 255) / 256); }

// paths on the device -> tree `t` in the traversal layout; returns the node count
// keys_ready: optional scratch buffer of n sort words already filled by the producer (skin kernel); taken over
size_t build_from_device_paths(Impl* I, int buftype, const int* p14, const int* p54, const int* p94, size_t n,
                               int first_modind, int levels, unsigned long long* keys_ready = nullptr)
{
    const int    t  = tree_index(buftype);
    cudaStream_t st = I->stream;
    size_t       launches = 0;
    int          total = 1; // nodes including the root
    int*         tmp_child = nullptr; // children by temporary id
    unsigned*    tmp_key   = nullptr; // creation key (creator << 4 | level) per temporary node
    if (n > 0 && levels > 0)
    {
        using u64 = unsigned long long;
        // 1. one word per point (path << 28 | index) sorted on the path bits; the first word of every distinct path
        //    carries the smallest index
        u64* keys_a = keys_ready ? keys_ready : scratch<u64>(I, n);
        u64* keys_b = scratch<u64>(I, n);
        if (!keys_ready)
            build_key_kernel<<<nblk(n), 256, 0, st>>>((const int4*) p14, (const int4*) p54, (const int4*) p94, n, levels,
                                                      keys_a);
        size_t tb = 0;
        CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys_a, keys_b, (int) n, BUILD_INDEX_BITS,
                                               BUILD_INDEX_BITS + 3 * levels, st));
        void* tmp = scratch<char>(I, tb);
        CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tb, keys_a, keys_b, (int) n, BUILD_INDEX_BITS,
                                               BUILD_INDEX_BITS + 3 * levels, st));
        scratch_free(I, tmp);
        unsigned char* heads = scratch<unsigned char>(I, n);
        build_head_flags_kernel<<<nblk(n), 256, 0, st>>>(keys_b, n, heads);
        int* counts = scratch<int>(I, 4);
        tb          = 0;
        CUDA_OK(cub::DeviceSelect::Flagged(nullptr, tb, keys_b, heads, keys_a, counts, (int) n, st));
        tmp = scratch<char>(I, tb);
        CUDA_OK(cub::DeviceSelect::Flagged(tmp, tb, keys_b, heads, keys_a, counts, (int) n, st));
        scratch_free(I, tmp);
        scratch_free(I, heads);
        int U = 0;
        CUDA_OK(cudaMemcpyAsync(&U, counts, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        const u64* K = keys_a;

        // 2. nodes every leaf heads, temporary ids by one scan, parent links
        unsigned char* first_diff  = scratch<unsigned char>(I, (size_t) U);
        int*           created     = scratch<int>(I, (size_t) U);
        int*           base        = scratch<int>(I, (size_t) U);
        int*           parent_leaf = scratch<int>(I, (size_t) U);
        build_leafinfo_kernel<<<nblk(U), 256, 0, st>>>(K, U, levels, first_diff, created);
        tb = 0;
        CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, (const unsigned*) created, (unsigned*) base, U, st));
        tmp = scratch<char>(I, tb);
        CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, (const unsigned*) created, (unsigned*) base, U, st));
        scratch_free(I, tmp);
        build_parent_kernel<<<nblk(U), 256, 0, st>>>(K, U, levels, first_diff, parent_leaf);
        unsigned last[2] = {0, 0};
        CUDA_OK(cudaMemcpyAsync(&last[0], base + (U - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(&last[1], created + (U - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        // the scan ran in unsigned arithmetic (at most 12 * 2^28 < 2^32); the node format holds 2^28 nodes
        const unsigned long long all_nodes = 1ull + last[0] + last[1];
        if (all_nodes >= (1ull << 28)) die("build_octree_from_paths: the tree would exceed 2^28 nodes");
        total = (int) all_nodes;

        // 3. creators and child tables, leaves upwards
        tmp_child    = scratch<int>(I, (size_t) total * 8);
        tmp_key      = scratch<unsigned>(I, (size_t) total);
        int* creator = scratch<int>(I, (size_t) total);
        CUDA_OK(cudaMemsetAsync(tmp_child, 0, (size_t) total * 8 * sizeof(int), st));
        build_fill_kernel<<<nblk(total), 256, 0, st>>>(creator, (size_t) total, 0x7fffffff);
        for (int d = levels; d >= 1; d--)
            build_link_kernel<<<nblk(U), 256, 0, st>>>(K, U, levels, d, first_diff, base, parent_leaf, creator,
                                                       tmp_child, tmp_key);
        CUDA_OK(cudaGetLastError());
        launches += 9 + (size_t) levels;
        scratch_free(I, creator);
        scratch_free(I, parent_leaf);
        scratch_free(I, base);
        scratch_free(I, created);
        scratch_free(I, first_diff);
        scratch_free(I, counts);
        scratch_free(I, keys_b);
        scratch_free(I, keys_a);
    }
    else
    {
        if (keys_ready) scratch_free(I, keys_ready);
        tmp_child = scratch<int>(I, 8);
        tmp_key   = scratch<unsigned>(I, 1);
        CUDA_OK(cudaMemsetAsync(tmp_child, 0, 8 * sizeof(int), st));
    }


    // reference numbering = rank in (creator, level) order
    int* final_of_tmp = scratch<int>(I, (size_t) total);
    if (total > 1)
    {
        const int m      = total - 1;
        int*      ids_in = scratch<int>(I, m);
        int*      ids_out = scratch<int>(I, m);
        unsigned* keys_out = scratch<unsigned>(I, m);
        build_iota_kernel<<<nblk(m), 256, 0, st>>>(ids_in, m, 1);
        size_t tb = 0;
        CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, tmp_key + 1, keys_out, ids_in, ids_out, m, 0, 32, st));
        void* tmp = scratch<char>(I, tb);
        CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tb, tmp_key + 1, keys_out, ids_in, ids_out, m, 0, 32, st));
        build_rank_kernel<<<nblk(m), 256, 0, st>>>(ids_out, m, final_of_tmp);
        launches += 3;
        scratch_free(I, tmp);
        scratch_free(I, ids_in);
        scratch_free(I, ids_out);
        scratch_free(I, keys_out);
    }
    else
        CUDA_OK(cudaMemsetAsync(final_of_tmp, 0, sizeof(int), st));

    // straight into the traversal layout of this tree
    CUDA_OK(cudaStreamSynchronize(st));
    ensure_capacity(I, buftype, (size_t) total * 48);
    build_emit_kernel<<<nblk(total), 256, 0, st>>>(tmp_child, tmp_key, final_of_tmp, total, first_modind,
                                                   (int4*) I->tree[t].up_child(), I->tree[t].up_model());
    CUDA_OK(cudaGetLastError());
    launches += 1;
    I->tree[t].nodes = (size_t) total; // like octree_reset + rebuild: the old extent is gone
    I->launches += launches;
    derive_slots(I, t, 0, (size_t) total);

    scratch_free(I, final_of_tmp);
    scratch_free(I, tmp_child);
    scratch_free(I, tmp_key);
    return (size_t) total;
}
} // namespace

extern "C" {

} // extern "C"
namespace
{
size_t build_octree_one(octree_glc_t* rc, const int32_t* oct14, const int32_t* oct54, const int32_t* oct94, size_t n,
                        int first_modind, int paths_on_device, octree_glc_buffer_t buftype)
{
    Impl* I = impl_of(rc);
    if (!is_octree(buftype)) die("build_octree_from_paths: buftype must be an octree buffer");
    if (n >= ((size_t) 1 << 28)) die("build_octree_from_paths: at most 2^28 points");
    flush_pending(I);
    cudaStream_t st = I->stream;

    const int *p14 = oct14, *p54 = oct54, *p94 = oct94;
    int*       staged = nullptr;
    if (!paths_on_device && n)
    {
        staged = scratch<int>(I, n * 12);
        CUDA_OK(cudaMemcpyAsync(staged, oct14, n * 16, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(staged + n * 4, oct54, n * 16, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(staged + n * 8, oct94, n * 16, cudaMemcpyHostToDevice, st));
        p14 = staged, p54 = staged + n * 4, p94 = staged + n * 8;
    }
    const size_t total = build_from_device_paths(I, buftype, p14, p54, p94, n, first_modind, BUILD_LEVELS);
    if (staged) scratch_free(I, staged);
    CUDA_OK(cudaStreamSynchronize(st)); // host path arrays have been consumed
    publish_memsize(rc, I);
    return total;
}

size_t voxelise_one(octree_glc_t* rc, const float* pos, const uint8_t* col_u8, const float* nrm, size_t n, int size,
                    int levels, int inputs_on_device, int dynamic, int64_t* order_host, float* pos_host)
{
    Impl* I = impl_of(rc);
    if (levels < 1 || levels > 12) die("voxelise_and_build: 1 <= levels <= 12");
    if (n >= ((size_t) 1 << 28)) die("voxelise_and_build: at most 2^28 points");
    flush_pending(I);
    cudaStream_t st = I->stream;
    const int    octbuf = dynamic ? OCTREE_GLC_BUFFER_DYNAMIC_OCTREE : OCTREE_GLC_BUFFER_STATIC_OCTREE;
    const int    colbuf = dynamic ? OCTREE_GLC_BUFFER_DYNAMIC_COLOR : OCTREE_GLC_BUFFER_STATIC_COLOR;
    const int    t      = dynamic ? 1 : 0;

    const float*   d_pos = pos;
    const uint8_t* d_col = col_u8;
    const float*   d_nrm = nrm;
    void *         s0 = nullptr, *s1 = nullptr, *s2 = nullptr;
    if (!inputs_on_device && n)
    {
        s0 = scratch<float>(I, n * 3);
        s1 = scratch<uint8_t>(I, n * 3);
        s2 = scratch<float>(I, n * 3);
        CUDA_OK(cudaMemcpyAsync(s0, pos, n * 12, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(s1, col_u8, n * 3, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(s2, nrm, n * 12, cudaMemcpyHostToDevice, st));
        d_pos = (const float*) s0, d_col = (const uint8_t*) s1, d_nrm = (const float*) s2;
    }

    int division = 2;
    for (int i = 0; i < levels; i++) division *= 2;                // qmc.c L217
    const float precision = (float) size / (float) division;        // qmc.c L218
    size_t      m         = 0;
    if (n)
    {
        unsigned long long* keys  = scratch<unsigned long long>(I, n);
        unsigned long long* keys2 = scratch<unsigned long long>(I, n);
        unsigned*           idx   = scratch<unsigned>(I, n);
        unsigned*           idx2  = scratch<unsigned>(I, n);
        voxel_key_kernel<<<nblk(n), 256, 0, st>>>(d_pos, n, precision, division, keys, idx);
        size_t tb = 0;
        CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, idx2, (long long) n, 0, 48, st));
        void* tmp = scratch<char>(I, tb);
        // LSD radix sort is stable: equal cells keep their source order, like the reference's sort + first-of-cell
        CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, idx2, (long long) n, 0, 48, st));
        int* flags = scratch<int>(I, n);
        int* slot  = scratch<int>(I, n);
        voxel_flag_kernel<<<nblk(n), 256, 0, st>>>(keys2, n, flags);
        size_t tb2 = 0;
        CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, flags, slot, (long long) n, st));
        void* tmp2 = scratch<char>(I, tb2);
        CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp2, tb2, flags, slot, (long long) n, st));
        int last_slot = 0, last_flag = 0;
        CUDA_OK(cudaMemcpyAsync(&last_slot, slot + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(&last_flag, flags + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        m = (size_t) last_slot + (size_t) last_flag;

        ensure_capacity(I, colbuf, m * 12);
        long long* order = scratch<long long>(I, m);
        float*     pout  = scratch<float>(I, m * 3);
        int*       p14   = scratch<int>(I, m * 4);
        int*       p54   = scratch<int>(I, m * 4);
        int*       p94   = scratch<int>(I, m * 4);
        voxel_emit_kernel<<<nblk(n), 256, 0, st>>>(d_pos, d_col, d_nrm, idx2, flags, slot, n, (float) size, levels,
                                                    (float*) I->pts[t].rec.ptr, pout, order, p14, p54, p94);
        CUDA_OK(cudaGetLastError());
        I->launches += 5;
        I->pts[t].points = m;
        build_from_device_paths(I, octbuf, p14, p54, p94, m, 0, levels);
        if (order_host) CUDA_OK(cudaMemcpyAsync(order_host, order, m * sizeof(long long), cudaMemcpyDeviceToHost, st));
        if (pos_host) CUDA_OK(cudaMemcpyAsync(pos_host, pout, m * 12, cudaMemcpyDeviceToHost, st));
        for (void* p : {(void*) keys, (void*) keys2, (void*) idx, (void*) idx2, tmp, (void*) flags, (void*) slot, tmp2,
                        (void*) order, (void*) pout, (void*) p14, (void*) p54, (void*) p94})
            scratch_free(I, p);
    }
    else
    {
        I->pts[t].points = 0;
        build_from_device_paths(I, octbuf, nullptr, nullptr, nullptr, 0, 0, levels);
    }
    if (s0) scratch_free(I, s0);
    if (s1) scratch_free(I, s1);
    if (s2) scratch_free(I, s2);
    CUDA_OK(cudaStreamSynchronize(st));
    publish_memsize(rc, I);
    return m;
}
} // namespace

extern "C" {

size_t octree_cuc_build_octree_from_paths(octree_glc_t* rc, const int32_t* oct14, const int32_t* oct54,
                                          const int32_t* oct94, size_t n, int first_modind, int paths_on_device,
                                          octree_glc_buffer_t buftype)
{
    Impl* I = impl_of(rc);
    if (paths_on_device && !I->replicas.empty())
        die("build_octree_from_paths: device-resident paths belong to one GPU; a multi-GPU group takes host paths");
    size_t total = 0;
    run_on_group(rc, [&](octree_glc_t* m, bool primary) {
        const size_t r = build_octree_one(m, oct14, oct54, oct94, n, first_modind, paths_on_device, buftype);
        if (primary) total = r;
    });
    return total;
}

size_t octree_cuc_voxelise_and_build(octree_glc_t* rc, const float* pos, const uint8_t* col_u8, const float* nrm,
                                     size_t n, int size, int levels, int inputs_on_device, int dynamic,
                                     int64_t* order_host, float* pos_host)
{
    Impl* I = impl_of(rc);
    if (inputs_on_device && !I->replicas.empty())
        die("voxelise_and_build: device-resident inputs belong to one GPU; a multi-GPU group takes host arrays");
    size_t m0 = 0;
    run_on_group(rc, [&](octree_glc_t* m, bool primary) {
        const size_t r = voxelise_one(m, pos, col_u8, nrm, n, size, levels, inputs_on_device, dynamic,
                                      primary ? order_host : nullptr, primary ? pos_host : nullptr);
        if (primary) m0 = r;
    });
    return m0;
}

} // extern "C"

namespace
{
struct h3
{
    float x, y, z;
};
// host fp32 helpers with separately rounded operations (the file is built with -ffp-contract=off)
inline h3    hsub(h3 a, h3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline h3    hadd(h3 a, h3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float hdot(h3 a, h3 b)
{
    volatile float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z;
    volatile float s = x + y;
    return s + z;
}
inline float hlen(h3 a) { return sqrtf(hdot(a, a)); }
inline float hdiv(float a, float b, int div)
{
    if (div == DIV_GLSL)
    {
        volatile float r = 1.0f / b;
        return a * r;
    }
    return a / b;
}
inline h3 hdiv3(h3 a, float f, int div) { return {hdiv(a.x, f, div), hdiv(a.y, f, div), hdiv(a.z, f, div)}; }
inline h3 hcross(h3 a, h3 b)
{
    volatile float x1 = a.y * b.z, x2 = b.y * a.z, y1 = a.z * b.x, y2 = b.z * a.x, z1 = a.x * b.y, z2 = b.x * a.y;
    return {x1 - x2, y1 - y2, z1 - z2};
}
inline void hquat(h3 axis, float angle, float* q)
{
    volatile float h = angle * 0.5f;
    const float    sn = sinf(h), cs = cosf(h);
    volatile float x = axis.x * sn, y = axis.y * sn, z = axis.z * sn;
    q[0] = x, q[1] = y, q[2] = z, q[3] = cs;
}

// skeleton_vsh.c L92-93, L103, L119, L132-151 for one bone pair: everything that does not depend on the point
void bone_consts(const float* ob, const float* nb, int div, BoneConsts* out10)
{
    for (int k = 0; k < 10; k++)
    {
        const int   i = 2 * k;
        BoneConsts& c = out10[k];
        memset(&c, 0, sizeof(c));
        h3 a = {ob[i * 4], ob[i * 4 + 1], ob[i * 4 + 2]}, b = {ob[i * 4 + 4], ob[i * 4 + 5], ob[i * 4 + 6]};
        h3 oldbone = hsub(b, a);
        h3 midp    = hadd(a, hdiv3(oldbone, 2.0f, div));
        h3 na = {nb[i * 4], nb[i * 4 + 1], nb[i * 4 + 2]}, nbb = {nb[i * 4 + 4], nb[i * 4 + 5], nb[i * 4 + 6]};
        h3 currbone = hsub(nbb, na);
        h3 on       = hdiv3(oldbone, hlen(oldbone), div);
        h3 cn       = hdiv3(currbone, hlen(currbone), div);
        c.a[0] = a.x, c.a[1] = a.y, c.a[2] = a.z;
        c.b[0] = b.x, c.b[1] = b.y, c.b[2] = b.z;
        c.effect     = ob[i * 4 + 3];
        c.oldbone[0] = oldbone.x, c.oldbone[1] = oldbone.y, c.oldbone[2] = oldbone.z;
        c.midp[0] = midp.x, c.midp[1] = midp.y, c.midp[2] = midp.z;
        c.half_len = hdiv(hlen(oldbone), 2.0f, div);
        {
            const double r = (double) c.half_len + fabs((double) ob[i * 4 + 3]) + 1.0;
            c.cull_r2      = (float) (r * r * 1.001);
        }
        c.ab_dot   = hdot(oldbone, oldbone);
        c.newa[0] = na.x, c.newa[1] = na.y, c.newa[2] = na.z;
        hquat(on, nb[i * 4 + 3], c.rot_quat);
        const float bones_dot   = hdot(on, cn);
        const float bones_angle = acosf(bones_dot);
        const h3    axis        = hcross(on, cn);
        c.has_axis              = hlen(axis) > 0.000001f;
        if (c.has_axis) hquat(hdiv3(axis, hlen(axis), div), bones_angle, c.axis_quat);
    }
}
} // namespace

extern "C" {

} // extern "C"
namespace
{
void skeleton_alloc_one(octree_glc_t* rc, const float* pntdata, const float* nrmdata, size_t bytes)
{
    Impl*        I = impl_of(rc);
    const size_t n = bytes / 12;
    CUDA_OK(cudaStreamSynchronize(I->stream));
    for (void* p : {(void*) I->skin_pos, (void*) I->skin_nrm, (void*) I->skin_p14, (void*) I->skin_p54,
                    (void*) I->skin_p94, (void*) I->skin_pnt_out})
        if (p) CUDA_OK(cudaFree(p));
    I->memsize -= I->skin_n * (12 + 12 + 48 + 12);
    I->skin_n = n;
    CUDA_OK(cudaMalloc(&I->skin_pos, (n ? n : 1) * 12));
    CUDA_OK(cudaMalloc(&I->skin_nrm, (n ? n : 1) * 12));
    CUDA_OK(cudaMalloc(&I->skin_p14, (n ? n : 1) * 16));
    CUDA_OK(cudaMalloc(&I->skin_p54, (n ? n : 1) * 16));
    CUDA_OK(cudaMalloc(&I->skin_p94, (n ? n : 1) * 16));
    CUDA_OK(cudaMalloc(&I->skin_pnt_out, (n ? n : 1) * 12));
    I->memsize += n * (12 + 12 + 48 + 12);
    CUDA_OK(cudaMemcpyAsync(I->skin_pos, pntdata, n * 12, cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaMemcpyAsync(I->skin_nrm, nrmdata, n * 12, cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    ensure_capacity(I, OCTREE_GLC_BUFFER_DYNAMIC_NORMAL, n * 12);
    publish_memsize(rc, I);
}

size_t skeleton_update_one(octree_glc_t* rc, const float* oldbones80, const float* newbones80, int model_count,
                           int maxlevel, float basesize, int build_tree)
{
    Impl* I = impl_of(rc);
    if (model_count < 0 || (size_t) model_count > I->skin_n) die("skeleton_update: more points than skeleton_alloc_in gave");
    if (maxlevel < 1 || maxlevel > 12) die("skeleton_update: 1 <= maxlevel <= 12");
    flush_pending(I);
    SkinParams S;
    bone_consts(oldbones80, newbones80, I->div_mode, S.bones);
    if (I->skin_rot_set)
        for (int k = 0; k < 10; k++)
        {
            const float* o = I->skin_rot + k * 9;
            memcpy(S.bones[k].rot_quat, o, 16);
            memcpy(S.bones[k].axis_quat, o + 4, 16);
            S.bones[k].has_axis = o[8] != 0.0f;
        }
    S.basesize = basesize;
    S.maxlevel = maxlevel;
    const size_t n = (size_t) model_count;
    size_t       nodes = 0;
    unsigned long long* keys = (build_tree && n) ? scratch<unsigned long long>(I, n) : nullptr;
    if (n)
    {
        float* rec = (float*) I->pts[1].rec.ptr;
        if (I->div_mode == DIV_GLSL)
            skin_kernel<DIV_GLSL><<<nblk(n), 256, 0, I->stream>>>(S, n, I->skin_pos, I->skin_nrm, I->skin_p14,
                                                                  I->skin_p54, I->skin_p94, rec, I->skin_pnt_out, keys);
        else
            skin_kernel<DIV_IEEE><<<nblk(n), 256, 0, I->stream>>>(S, n, I->skin_pos, I->skin_nrm, I->skin_p14,
                                                                  I->skin_p54, I->skin_p94, rec, I->skin_pnt_out, keys);
        CUDA_OK(cudaGetLastError());
        I->launches++;
        if (n > I->pts[1].points) I->pts[1].points = n;
    }
    I->skin_count = n;
    if (build_tree)
        nodes = build_from_device_paths(I, OCTREE_GLC_BUFFER_DYNAMIC_OCTREE, (const int*) I->skin_p14,
                                        (const int*) I->skin_p54, (const int*) I->skin_p94, n, 0, maxlevel, keys);
    publish_memsize(rc, I);
    return nodes;
}
} // namespace

extern "C" {

void octree_cuc_skeleton_alloc_in(octree_glc_t* rc, const float* pntdata, const float* nrmdata, size_t bytes)
{
    impl_of(rc);
    run_on_group(rc, [=](octree_glc_t* m, bool) { skeleton_alloc_one(m, pntdata, nrmdata, bytes); });
}

// Every device of a group skins and builds for itself: the inputs are 160 floats, the outputs (the dynamic tree and
// 10 M normals) would be hundreds of megabytes to send around, and the kernels are deterministic.
size_t octree_cuc_skeleton_update(octree_glc_t* rc, const float* oldbones80, const float* newbones80, int model_count,
                                  int maxlevel, float basesize, int build_tree)
{
    impl_of(rc);
    size_t nodes = 0;
    run_on_group(rc, [&](octree_glc_t* m, bool primary) {
        const size_t r = skeleton_update_one(m, oldbones80, newbones80, model_count, maxlevel, basesize, build_tree);
        if (primary) nodes = r;
    });
    return nodes;
}

void octree_cuc_skeleton_set_rotations(octree_glc_t* rc, const float* rotations90)
{
    Impl* I         = impl_of(rc);
    I->skin_rot_set = rotations90 != nullptr;
    if (rotations90) memcpy(I->skin_rot, rotations90, sizeof(I->skin_rot));
    REPLAY(I, octree_cuc_skeleton_set_rotations(m, rotations90));
}

size_t octree_cuc_skeleton_read_out(octree_glc_t* rc, int32_t* oct14, int32_t* oct54, int32_t* oct94, float* nrm_out,
                                    float* pnt_out)
{
    Impl*        I = impl_of(rc);
    const size_t n = I->skin_count;
    cudaStream_t st = I->stream;
    if (oct14) CUDA_OK(cudaMemcpyAsync(oct14, I->skin_p14, n * 16, cudaMemcpyDeviceToHost, st));
    if (oct54) CUDA_OK(cudaMemcpyAsync(oct54, I->skin_p54, n * 16, cudaMemcpyDeviceToHost, st));
    if (oct94) CUDA_OK(cudaMemcpyAsync(oct94, I->skin_p94, n * 16, cudaMemcpyDeviceToHost, st));
    if (pnt_out) CUDA_OK(cudaMemcpyAsync(pnt_out, I->skin_pnt_out, n * 12, cudaMemcpyDeviceToHost, st));
    float* packed = nullptr;
    if (nrm_out && n)
    {
        packed = scratch<float>(I, n * 3);
        gather_normals_kernel<<<nblk(n), 256, 0, st>>>((const float*) I->pts[1].rec.ptr, n, packed);
        CUDA_OK(cudaGetLastError());
        I->launches++;
        CUDA_OK(cudaMemcpyAsync(nrm_out, packed, n * 12, cudaMemcpyDeviceToHost, st));
        scratch_free(I, packed);
    }
    CUDA_OK(cudaStreamSynchronize(st));
    return n;
}

void octree_cuc_particles_alloc_in(octree_glc_t* rc, int kind, const float* posdata, const float* spddata, size_t bytes)
{
    Impl* I = impl_of(rc);
    if (kind != OCTREE_CUC_PARTICLES && kind != OCTREE_CUC_DUST) die("particles_alloc_in: unknown kind");
    const size_t n = bytes / 12;
    CUDA_OK(cudaStreamSynchronize(I->stream));
    if (n > I->part_n[kind] || !I->part_pos[kind][0])
    {
        for (int b = 0; b < 2; b++)
        {
            if (I->part_pos[kind][b]) CUDA_OK(cudaFree(I->part_pos[kind][b]));
            if (I->part_spd[kind][b]) CUDA_OK(cudaFree(I->part_spd[kind][b]));
            CUDA_OK(cudaMalloc(&I->part_pos[kind][b], (n ? n : 1) * 12));
            CUDA_OK(cudaMalloc(&I->part_spd[kind][b], (n ? n : 1) * 12));
        }
        I->memsize += (n - I->part_n[kind]) * 48;
        I->part_n[kind] = n;
    }
    if (!I->part_finished)
    {
        CUDA_OK(cudaMalloc(&I->part_finished, sizeof(unsigned)));
        CUDA_OK(cudaMemsetAsync(I->part_finished, 0, sizeof(unsigned), I->stream));
    }
    I->part_cur[kind] = 0;
    CUDA_OK(cudaMemcpyAsync(I->part_pos[kind][0], posdata, n * 12, cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaMemcpyAsync(I->part_spd[kind][0], spddata, n * 12, cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream)); // the host arrays may change right after the call
    publish_memsize(rc, I);
}

void octree_cuc_particles_update(octree_glc_t* rc, int kind, int count, int maxlevel, float basesize, v3_t campos,
                                 int steps)
{
    Impl* I = impl_of(rc);
    if (kind != OCTREE_CUC_PARTICLES && kind != OCTREE_CUC_DUST) die("particles_update: unknown kind");
    if (count < 0 || (size_t) count > I->part_n[kind]) die("particles_update: more particles than particles_alloc_in gave");
    if (maxlevel < 0 || maxlevel > GENERIC_STACK - 1) die("particles_update: maxlevel out of range");
    if (count == 0 || steps <= 0) return;
    flush_pending(I);
    cudaStream_t st = I->stream;
    FrameParams  P;
    memset(&P, 0, sizeof(P));
    P.tree_s       = tree_dev(I, 0); // particle_glc.c L106-107: the static octree only
    P.tree_d       = P.tree_s;
    P.tree_d.nodes = 0; // every index clamps to the dummy
    P.basecube[0]  = 0.0f;
    P.basecube[1] = P.basecube[2] = P.basecube[3] = basesize;
    P.maxlevel                                    = maxlevel;
    // exact grid: the fast traversal (kernel choice 1 forces the generic one, like for frames)
    const bool   fast = I->kernel_choice != 1 && maxlevel >= 1 && grid_is_exact(basesize, maxlevel);
    const size_t smem = (size_t) 3 * maxlevel * BLOCK_THREADS * sizeof(int);
    if (fast)
    {
        P.leaf_size     = ldexpf(basesize, -maxlevel);
        P.inv_leaf_size = 1.0f / P.leaf_size;
    }
    const size_t n = (size_t) count;
    for (int s = 0; s < steps; s++)
    {
        const int a = I->part_cur[kind], b = a ^ 1;
        if (kind == OCTREE_CUC_PARTICLES)
        {
            CUDA_OK(cudaMemsetAsync(I->part_finished, 0, sizeof(unsigned), st));
            auto*  pa = I->part_pos[kind][a];
            auto*  sa = I->part_spd[kind][a];
            auto*  pb = I->part_pos[kind][b];
            auto*  sb = I->part_spd[kind][b];
            if (fast)
            {
                const unsigned g = (unsigned) ((n + BLOCK_THREADS - 1) / BLOCK_THREADS);
                if (I->div_mode == DIV_GLSL)
                    particle_step_kernel<DIV_GLSL, true><<<g, BLOCK_THREADS, smem, st>>>(P, n, pa, sa, pb, sb, I->part_finished);
                else
                    particle_step_kernel<DIV_IEEE, true><<<g, BLOCK_THREADS, smem, st>>>(P, n, pa, sa, pb, sb, I->part_finished);
            }
            else if (I->div_mode == DIV_GLSL)
                particle_step_kernel<DIV_GLSL, false><<<nblk(n), 256, 0, st>>>(P, n, pa, sa, pb, sb, I->part_finished);
            else
                particle_step_kernel<DIV_IEEE, false><<<nblk(n), 256, 0, st>>>(P, n, pa, sa, pb, sb, I->part_finished);
        }
        else
            dust_step_kernel<<<nblk(n), 256, 0, st>>>(make_float3(campos.x, campos.y, campos.z), n, I->part_pos[kind][a],
                                                      I->part_spd[kind][a], I->part_pos[kind][b], I->part_spd[kind][b]);
        CUDA_OK(cudaGetLastError());
        I->launches++;
        I->part_cur[kind] = b;
    }
}

size_t octree_cuc_particles_read_out(octree_glc_t* rc, int kind, int count, float* pos_out, float* spd_out)
{
    Impl* I = impl_of(rc);
    if (kind != OCTREE_CUC_PARTICLES && kind != OCTREE_CUC_DUST) die("particles_read_out: unknown kind");
    if (count < 0 || (size_t) count > I->part_n[kind]) die("particles_read_out: count beyond particles_alloc_in");
    const int a = I->part_cur[kind];
    if (pos_out) CUDA_OK(cudaMemcpyAsync(pos_out, I->part_pos[kind][a], (size_t) count * 12, cudaMemcpyDeviceToHost, I->stream));
    if (spd_out) CUDA_OK(cudaMemcpyAsync(spd_out, I->part_spd[kind][a], (size_t) count * 12, cudaMemcpyDeviceToHost, I->stream));
    unsigned fin = 0;
    if (I->part_finished) CUDA_OK(cudaMemcpyAsync(&fin, I->part_finished, sizeof(unsigned), cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return kind == OCTREE_CUC_PARTICLES ? (size_t) fin : 0;
}

void octree_cuc_trace_lines(octree_glc_t* rc, size_t n, const float* pos, const float* dir, int dynamic_tree,
                            int maxlevel, float basesize, int32_t* out_index, float* out_tlf)
{
    Impl* I = impl_of(rc);
    if (n == 0) return;
    if (maxlevel < 0 || maxlevel > GENERIC_STACK - 1) die("trace_lines: maxlevel out of range");
    flush_pending(I);
    cudaStream_t st = I->stream;
    FrameParams  P;
    memset(&P, 0, sizeof(P));
    const int t     = dynamic_tree ? 1 : 0;
    P.tree_s        = tree_dev(I, t); // the queried tree in the first slot, nothing in the second
    P.tree_d        = P.tree_s;
    P.tree_d.nodes  = 0; // every index clamps to the dummy
    P.basecube[0]   = 0.0f;
    P.basecube[1] = P.basecube[2] = P.basecube[3] = basesize;
    P.maxlevel                                    = maxlevel;
    float* d_pos = scratch<float>(I, n * 3);
    float* d_dir = scratch<float>(I, n * 3);
    int*   d_idx = scratch<int>(I, n);
    float* d_tlf = out_tlf ? scratch<float>(I, n * 4) : nullptr;
    CUDA_OK(cudaMemcpyAsync(d_pos, pos, n * 12, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_dir, dir, n * 12, cudaMemcpyHostToDevice, st));
    if (d_tlf) CUDA_OK(cudaMemcpyAsync(d_tlf, out_tlf, n * 16, cudaMemcpyHostToDevice, st)); // misses keep the caller's values
    if (I->kernel_choice != 1 && maxlevel >= 1 && grid_is_exact(basesize, maxlevel))
    {
        P.leaf_size     = ldexpf(basesize, -maxlevel);
        P.inv_leaf_size = 1.0f / P.leaf_size;
        trace_lines_fast_kernel<<<(unsigned) ((n + BLOCK_THREADS - 1) / BLOCK_THREADS), BLOCK_THREADS,
                                  (size_t) 3 * maxlevel * BLOCK_THREADS * sizeof(int), st>>>(P, n, d_pos, d_dir, d_idx, d_tlf);
    }
    else
        trace_lines_kernel<<<nblk(n), 256, 0, st>>>(P, n, d_pos, d_dir, d_idx, d_tlf);
    CUDA_OK(cudaGetLastError());
    I->launches++;
    CUDA_OK(cudaMemcpyAsync(out_index, d_idx, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (d_tlf) CUDA_OK(cudaMemcpyAsync(out_tlf, d_tlf, n * 16, cudaMemcpyDeviceToHost, st));
    scratch_free(I, d_pos);
    scratch_free(I, d_dir);
    scratch_free(I, d_idx);
    if (d_tlf) scratch_free(I, d_tlf);
    CUDA_OK(cudaStreamSynchronize(st));
}

size_t octree_cuc_download_points(octree_glc_t* rc, int dynamic, float* col_host, float* nrm_host, size_t capacity_points)
{
    Impl*        I = impl_of(rc);
    const int    t = dynamic ? 1 : 0;
    const size_t m = I->pts[t].points;
    if (capacity_points < m || (!col_host && !nrm_host)) return m;
    flush_pending(I);
    std::vector<float> rec(m * 8);
    CUDA_OK(cudaMemcpyAsync(rec.data(), I->pts[t].rec.ptr, m * 32, cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    for (size_t i = 0; i < m; i++)
        for (int c = 0; c < 3; c++)
        {
            if (col_host) col_host[i * 3 + c] = rec[i * 8 + c];
            if (nrm_host) nrm_host[i * 3 + c] = rec[i * 8 + 4 + c];
        }
    return m;
}

size_t octree_cuc_download_octree(octree_glc_t* rc, octree_glc_buffer_t buftype, int32_t* nodes12_host,
                                  size_t capacity_nodes)
{
    Impl* I = impl_of(rc);
    if (!is_octree(buftype)) die("download_octree: buftype must be an octree buffer");
    flush_pending(I);
    const int    t = tree_index(buftype);
    const size_t n = I->tree[t].nodes;
    if (nodes12_host == nullptr || capacity_nodes < n) return n;
    int* out = scratch<int>(I, n * 12);
    export_nodes_kernel<<<nblk(n), 256, 0, I->stream>>>((const int4*) I->tree[t].up_child(), I->tree[t].up_model(), n,
                                                        out);
    CUDA_OK(cudaGetLastError());
    I->launches++;
    CUDA_OK(cudaMemcpyAsync(nodes12_host, out, n * 48, cudaMemcpyDeviceToHost, I->stream));
    scratch_free(I, out);
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return n;
}

uint64_t octree_cuc_selftest_div(octree_glc_t* rc, uint64_t seed, uint64_t count)
{
    Impl* I = impl_of(rc);
    CUDA_OK(cudaMemsetAsync(I->counters, 0, sizeof(unsigned long long), I->stream));
    selftest_div_kernel<<<148 * 8, 256, 0, I->stream>>>(seed, count, I->counters);
    CUDA_OK(cudaGetLastError());
    I->launches++;
    unsigned long long bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, I->counters, sizeof(bad), cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    return (uint64_t) bad;
}

void octree_cuc_debug_order_lut(uint64_t* out4096)
{
    for (int i = 0; i < ORDER_LUT_SIZE; i++) out4096[i] = (uint64_t) h_order_lut_copy.v[i];
}

// blob = { uint64 ndesc, uint64 payload_bytes, RangeDesc[ndesc], payload }
void octree_cuc_enable_replication_log(octree_glc_t* rc, int on)
{
    Impl* I    = impl_of(rc);
    I->repl_on = on != 0;
    I->repl_descs.clear();
    I->repl_payload.clear();
}

size_t octree_cuc_export_pending(octree_glc_t* rc, void* blob_host, size_t capacity)
{
    Impl* I = impl_of(rc);
    if (!I->repl_on) die("export_pending: call octree_cuc_enable_replication_log first (ranges already applied are gone)");
    const size_t nd   = I->repl_descs.size();
    const size_t need = 16 + nd * sizeof(RangeDesc) + I->repl_payload.size();
    if (blob_host == nullptr || capacity < need) return need;
    uint64_t hdr[2] = {(uint64_t) nd, (uint64_t) I->repl_payload.size()};
    char*    p      = (char*) blob_host;
    memcpy(p, hdr, 16);
    if (nd) memcpy(p + 16, I->repl_descs.data(), nd * sizeof(RangeDesc));
    if (!I->repl_payload.empty()) memcpy(p + 16 + nd * sizeof(RangeDesc), I->repl_payload.data(), I->repl_payload.size());
    I->repl_descs.clear();
    I->repl_payload.clear();
    return need;
}

// The same blob written straight into DEVICE memory (the send buffer of the broadcast), payload from the page-locked
// log at PCIe rate; the connector's stream is synchronised before the log is cleared.
size_t octree_cuc_export_pending_device(octree_glc_t* rc, uint64_t blob_device, size_t capacity)
{
    Impl* I = impl_of(rc);
    if (!I->repl_on) die("export_pending: call octree_cuc_enable_replication_log first (ranges already applied are gone)");
    const size_t nd   = I->repl_descs.size();
    const size_t need = 16 + nd * sizeof(RangeDesc) + I->repl_payload.size();
    if (blob_device == 0 || capacity < need) return need;
    uint64_t hdr[2] = {(uint64_t) nd, (uint64_t) I->repl_payload.size()};
    char*    p      = (char*) (uintptr_t) blob_device;
    CUDA_OK(cudaMemcpyAsync(p, hdr, 16, cudaMemcpyHostToDevice, I->stream));
    if (nd)
        CUDA_OK(cudaMemcpyAsync(p + 16, I->repl_descs.data(), nd * sizeof(RangeDesc), cudaMemcpyHostToDevice, I->stream));
    if (!I->repl_payload.empty())
        CUDA_OK(cudaMemcpyAsync(p + 16 + nd * sizeof(RangeDesc), I->repl_payload.data(), I->repl_payload.size(),
                                cudaMemcpyHostToDevice, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    I->repl_descs.clear();
    I->repl_payload.clear();
    return need;
}

void octree_cuc_apply_blob(octree_glc_t* rc, const void* blob_host, size_t bytes)
{
    Impl* I = impl_of(rc);
    if (bytes < 16) return;
    const char* p = (const char*) blob_host;
    uint64_t    hdr[2];
    memcpy(hdr, p, 16);
    if (hdr[0] > (bytes - 16) / sizeof(RangeDesc) || 16 + hdr[0] * sizeof(RangeDesc) + hdr[1] != bytes || hdr[1] % 4)
        die("apply_blob: malformed blob");
    const char* payload = p + 16 + hdr[0] * sizeof(RangeDesc);
    // every descriptor is checked before anything is applied
    for (uint64_t i = 0; i < hdr[0]; i++)
    {
        RangeDesc d;
        memcpy(&d, p + 16 + i * sizeof(RangeDesc), sizeof(d));
        if (d.buftype < 0 || d.buftype > 5) die("apply_blob: descriptor with an unknown buffer type");
        if (d.nwords == 0 || ((uint64_t) d.src_word + d.nwords) * 4 > hdr[1]) die("apply_blob: descriptor outside the payload");
        const uint64_t item = is_octree(d.buftype) ? 16 : 12;
        if ((d.dst_word * 4) % item || ((uint64_t) d.nwords * 4) % item) die("apply_blob: range is not texel-aligned");
        if (is_octree(d.buftype) && (d.dst_word + d.nwords + 11) / 12 > (uint64_t) CHILD_INDEX_MASK - 1)
            die("apply_blob: range beyond the node limit");
    }
    for (uint64_t i = 0; i < hdr[0]; i++)
    {
        RangeDesc d;
        memcpy(&d, p + 16 + i * sizeof(RangeDesc), sizeof(d));
        const size_t s = (size_t) d.dst_word * 4, e = s + (size_t) d.nwords * 4;
        ensure_capacity(I, d.buftype, e);
        apply_range(I, payload + (size_t) d.src_word * 4, d.buftype, s, e);
    }
    REPLAY(I, octree_cuc_apply_blob(m, blob_host, bytes));
    publish_memsize(rc, I);
}

// The same blob already in DEVICE memory of this connector's GPU (the receive buffer of the broadcast): descriptors
// are validated from a small host copy, the payload never leaves the device -- one scatter launch applies it.
void octree_cuc_apply_blob_device(octree_glc_t* rc, uint64_t blob_device, size_t bytes)
{
    Impl* I = impl_of(rc);
    if (!I->replicas.empty()) die("apply_blob_device: a group replicates its uploads itself");
    if (bytes < 16) return;
    const char* p = (const char*) (uintptr_t) blob_device;
    if (((uintptr_t) p & 7) != 0) die("apply_blob_device: the blob must be 8-byte aligned");
    uint64_t hdr[2];
    CUDA_OK(cudaMemcpyAsync(hdr, p, 16, cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    if (hdr[0] > (bytes - 16) / sizeof(RangeDesc) || 16 + hdr[0] * sizeof(RangeDesc) + hdr[1] != bytes || hdr[1] % 4 ||
        hdr[1] / 4 > 0xffffffffull)
        die("apply_blob: malformed blob");
    if (hdr[0] == 0) return;
    std::vector<RangeDesc> descs(hdr[0]);
    CUDA_OK(cudaMemcpyAsync(descs.data(), p + 16, hdr[0] * sizeof(RangeDesc), cudaMemcpyDeviceToHost, I->stream));
    CUDA_OK(cudaStreamSynchronize(I->stream));
    uint64_t next_src = 0;
    for (const RangeDesc& d : descs)
    {
        if (d.buftype < 0 || d.buftype > 5) die("apply_blob: descriptor with an unknown buffer type");
        if (d.nwords == 0 || ((uint64_t) d.src_word + d.nwords) * 4 > hdr[1]) die("apply_blob: descriptor outside the payload");
        if (d.src_word != next_src) die("apply_blob: descriptors must tile the payload in order");
        next_src = (uint64_t) d.src_word + d.nwords;
        const uint64_t item = is_octree(d.buftype) ? 16 : 12;
        if ((d.dst_word * 4) % item || ((uint64_t) d.nwords * 4) % item) die("apply_blob: range is not texel-aligned");
        if (is_octree(d.buftype) && (d.dst_word + d.nwords + 11) / 12 > (uint64_t) CHILD_INDEX_MASK - 1)
            die("apply_blob: range beyond the node limit");
    }
    if (next_src * 4 != hdr[1]) die("apply_blob: descriptors must tile the payload in order");
    flush_pending(I); // call order is preserved against ranges this connector received directly
    for (const RangeDesc& d : descs)
    {
        const size_t e = ((size_t) d.dst_word + d.nwords) * 4;
        ensure_capacity(I, d.buftype, e);
        note_extent(I, d.buftype, e);
    }
    const unsigned words = (unsigned) (hdr[1] / 4);
    scatter_ranges_kernel<<<(words + 255) / 256, 256, 0, I->stream>>>(
        (const RangeDesc*) (p + 16), (int) hdr[0], (const int*) (p + 16 + hdr[0] * sizeof(RangeDesc)), words,
        scatter_targets(I));
    CUDA_OK(cudaGetLastError());
    I->launches++;
    derive_ranges(I, (const RangeDesc*) (p + 16), (int) hdr[0], words);
    publish_memsize(rc, I);
}

} // extern "C"
