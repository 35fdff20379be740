/*
 * qb_host.c -- host-side data model that feeds the connector: the 12-int
 * octree array and the voxelised point model, in the formats the reference
 * engine hands to octree_glc_upload_texbuffer_data().
 *
 * Own implementation (growable flat arrays, radix-sorted voxeliser); the
 * semantics it must reproduce are those of
 *   /root/reference/src/qubatron/octree.c  L55-67  octree_create
 *                                          L95-147 octree_insert_point
 *                                          L149-180 octree_insert_path
 *                                          L182-218 octree_remove_point
 *   /root/reference/src/qubatron/qmc.c     L62-83, L130-165, L217-218, L259,
 *                                          L291-327 (grid index, x-major sort,
 *                                          first-point-per-cell compaction)
 * and tests/test_host_model.py checks node-for-node / point-for-point equality
 * with the reference's compiled code.
 *
 * Plain C ABI, used through ctypes by qubatron_b200/scene.py and from C by
 * examples/host_demo.c.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define QB_NODE_INTS 12 /* octree.c L11-14: 8 children, model index, 3 pad */

typedef struct qb_octree
{
    int32_t* nodes; /* QB_NODE_INTS per node, node 0 = root */
    int64_t  len;
    int64_t  cap;
    float    basesize;
    int      levels;
} qb_octree;

static void tree_reserve(qb_octree* t, int64_t want)
{
    if (want <= t->cap) return;
    int64_t cap = t->cap ? t->cap : 1024;
    while (cap < want) cap += cap / 2 + 1024;
    int32_t* n = realloc(t->nodes, (size_t) cap * QB_NODE_INTS * sizeof(int32_t));
    if (!n)
    {
        fprintf(stderr, "qb_host: out of memory growing octree to %lld nodes\n", (long long) cap);
        abort();
    }
    memset(n + t->cap * QB_NODE_INTS, 0, (size_t) (cap - t->cap) * QB_NODE_INTS * sizeof(int32_t));
    t->nodes = n;
    t->cap   = cap;
}

qb_octree* qb_octree_create(float basesize, int levels)
{
    qb_octree* t = calloc(1, sizeof(*t));
    t->basesize  = basesize;
    t->levels    = levels;
    tree_reserve(t, 1024);
    t->len = 1; /* root, all zero */
    return t;
}

void qb_octree_delete(qb_octree* t)
{
    if (!t) return;
    free(t->nodes);
    free(t);
}

/* octree.c L89-93 */
void qb_octree_reset(qb_octree* t)
{
    memset(t->nodes, 0, QB_NODE_INTS * sizeof(int32_t));
    t->len = 1;
}

/* take over an existing node array (a level loaded from its flat file, model.c L53-111) */
void qb_octree_adopt(qb_octree* t, const int32_t* nodes, int64_t n)
{
    tree_reserve(t, n + n / 8 + 1024);
    memcpy(t->nodes, nodes, (size_t) n * QB_NODE_INTS * sizeof(int32_t));
    t->len = n;
}

/* The engine's loop over the nodes a shot touched (modelutil.c L429-437, L486-501): one
 * octree_glc_upload_texbuffer_data call per 48-byte node, `data` = the whole node array.  `upload` is that entry
 * point of the connector in use (this library stays independent of it). */
typedef void (*qb_upload_fn)(void* rc, void* data, int type, size_t size, size_t itemsize, size_t start, size_t end,
                             int buftype);
void qb_upload_node_ranges(qb_upload_fn upload, void* rc, qb_octree* t, const int32_t* node_index, int64_t count,
                           int buftype)
{
    const size_t node = QB_NODE_INTS * sizeof(int32_t);
    for (int64_t i = 0; i < count; i++)
        upload(rc, t->nodes, 0x1404 /* GL_INT */, (size_t) t->len * node, 4 * sizeof(int32_t),
               (size_t) node_index[i] * node, (size_t) (node_index[i] + 1) * node, buftype);
}

int64_t  qb_octree_len(const qb_octree* t) { return t->len; }
int32_t* qb_octree_nodes(qb_octree* t) { return t->nodes; }

/* octant of a point at one level (octree.c L102-109): bit0 = upper x half,
 * +2 = LOWER y half, +4 = LOWER z half; size is halved before use */
static inline int octant_of(float px, float py, float pz, float size)
{
    int o  = ((int) (px / size)) % 2;
    int yi = ((int) (py / size)) % 2;
    int zi = ((int) (pz / size)) % 2;
    if (yi == 0) o += 2;
    if (zi == 0) o += 4;
    return o;
}

static inline int64_t step_or_create(qb_octree* t, int64_t index, int octant, int64_t modind, int* created)
{
    int32_t child = t->nodes[index * QB_NODE_INTS + octant];
    if (child == 0)
    {
        tree_reserve(t, t->len + 1);
        child                                  = (int32_t) t->len;
        t->nodes[index * QB_NODE_INTS + octant] = child;
        int32_t* n                             = t->nodes + (int64_t) child * QB_NODE_INTS;
        memset(n, 0, QB_NODE_INTS * sizeof(int32_t));
        n[8] = (int32_t) modind;
        t->len++;
        if (created) *created = 1;
    }
    return child;
}

/* insert one point; touched[13] (may be NULL) receives, like octindarr of
 * octree.c L113-119, the parent and the new node of every level that created
 * a node (0 elsewhere) */
void qb_octree_insert_point(qb_octree* t, const float* p, int64_t modind, int32_t* touched)
{
    float   size  = t->basesize;
    int64_t index = 0;
    for (int level = 0; level < t->levels; level++)
    {
        size        = (float) (size / 2.0);
        int o       = octant_of(p[0], p[1], p[2], size);
        int created = 0;
        int64_t nxt = step_or_create(t, index, o, modind, &created);
        if (created && touched)
        {
            touched[level]     = (int32_t) index;
            touched[level + 1] = (int32_t) nxt;
        }
        index = nxt;
    }
}

/* modelutil.c L203-217: points in array order, model index = first_modind + i */
void qb_octree_insert_points(qb_octree* t, const float* pts, int64_t n, int64_t first_modind)
{
    for (int64_t i = 0; i < n; i++) qb_octree_insert_point(t, pts + i * 3, first_modind + i, NULL);
}

/* qubatron.c L439-452 / octree.c L149-180: precomputed octant digits, 12 per point */
void qb_octree_insert_paths(qb_octree* t, const int32_t* paths, int64_t n, int64_t first_modind)
{
    for (int64_t i = 0; i < n; i++)
    {
        int64_t index = 0;
        for (int level = 0; level < t->levels; level++)
            index = step_or_create(t, index, paths[i * 12 + level], first_modind + i, NULL);
    }
}

/* octree.c L182-218: walk levels+1 steps; at the leaf report its model index
 * and its PARENT node, and zero the parent's slot.  Outputs stay -1 when the
 * walk falls off the tree. */
void qb_octree_remove_point(qb_octree* t, const float* p, int32_t* modind, int32_t* octind)
{
    float   size   = t->basesize;
    int64_t index  = 0;
    int64_t lindex = 0;
    int     loct   = 0;
    *modind        = -1;
    *octind        = -1;
    for (int level = 0; level < t->levels + 1; level++)
    {
        size  = (float) (size / 2.0);
        int o = octant_of(p[0], p[1], p[2], size);
        if (level == t->levels)
        {
            *modind                                = t->nodes[index * QB_NODE_INTS + 8];
            *octind                                = (int32_t) lindex;
            t->nodes[lindex * QB_NODE_INTS + loct] = 0;
        }
        lindex = index;
        loct   = o;
        index  = t->nodes[index * QB_NODE_INTS + o];
        if (index == 0) return;
    }
}

/* ------------------------------------------------------------------------ */
/* voxeliser (qmc): grid index at 2*2^levels cells per axis, drop outside,   */
/* stable x-major sort, keep the first point of each occupied cell.          */
/* Returns the number of surviving points; order[] receives their source     */
/* indices in output order (caller gathers pos/col/nrm with it).             */
/* ------------------------------------------------------------------------ */
int64_t qb_voxelise_order(const float* pts, int64_t n, int size, int levels, int64_t* order, int64_t* dropped_out)
{
    int division = 2;
    for (int i = 0; i < levels; i++) division *= 2;
    float precision = (float) size / (float) division; /* qmc.c L217-218 */

    uint16_t* key = malloc((size_t) n * 3 * sizeof(uint16_t));
    uint32_t* a   = malloc((size_t) n * sizeof(uint32_t));
    uint32_t* b   = malloc((size_t) n * sizeof(uint32_t));
    int64_t*  cnt = malloc(((size_t) division + 1) * sizeof(int64_t));
    if (!key || !a || !b || !cnt || division > 65536 || n > 0xffffffffLL)
    {
        fprintf(stderr, "qb_host: voxelise cannot handle n=%lld division=%d\n", (long long) n, division);
        abort();
    }

    int64_t m = 0, dropped = 0;
    for (int64_t i = 0; i < n; i++)
    {
        /* qmc.c L62-68: floor(p / precision) evaluated in fp32 then widened */
        float xi = floorf(pts[i * 3 + 0] / precision);
        float yi = floorf(pts[i * 3 + 1] / precision);
        float zi = floorf(pts[i * 3 + 2] / precision);
        if (xi >= 0.0f && xi < (float) division && yi >= 0.0f && yi < (float) division && zi >= 0.0f &&
            zi < (float) division)
        {
            key[i * 3 + 0] = (uint16_t) xi;
            key[i * 3 + 1] = (uint16_t) yi;
            key[i * 3 + 2] = (uint16_t) zi;
            a[m++]         = (uint32_t) i;
        }
        else
            dropped++;
    }

    /* LSD radix: z, then y, then x => x-major order, stable in source order */
    for (int pass = 2; pass >= 0; pass--)
    {
        memset(cnt, 0, ((size_t) division + 1) * sizeof(int64_t));
        for (int64_t i = 0; i < m; i++) cnt[key[(int64_t) a[i] * 3 + pass] + 1]++;
        for (int k = 0; k < division; k++) cnt[k + 1] += cnt[k];
        for (int64_t i = 0; i < m; i++) b[cnt[key[(int64_t) a[i] * 3 + pass]]++] = a[i];
        uint32_t* tmp = a;
        a             = b;
        b             = tmp;
    }

    int64_t out = 0;
    for (int64_t i = 0; i < m; i++)
    {
        const uint16_t* k = key + (int64_t) a[i] * 3;
        if (i > 0)
        {
            const uint16_t* p = key + (int64_t) a[i - 1] * 3;
            if (p[0] == k[0] && p[1] == k[1] && p[2] == k[2]) continue;
        }
        order[out++] = a[i];
    }

    free(key);
    free(a);
    free(b);
    free(cnt);
    if (dropped_out) *dropped_out = dropped;
    return out;
}

/* gather rows of a float[3] array: dst[i] = src[order[i]] */
void qb_gather_f3(const float* src, const int64_t* order, int64_t n, float* dst)
{
    for (int64_t i = 0; i < n; i++)
    {
        const float* s = src + order[i] * 3;
        dst[i * 3 + 0] = s[0];
        dst[i * 3 + 1] = s[1];
        dst[i * 3 + 2] = s[2];
    }
}
