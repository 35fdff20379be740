/*
 * ref_wrap.c -- batch entry points around the UNMODIFIED reference data model.
 *
 * TEST INFRASTRUCTURE ONLY.  This translation unit contains no reference code:
 * it #includes /root/reference/src/qubatron/octree.c in "header mode" (the
 * reference's __INCLUDE_LEVEL__ idiom, octree.c L1-39) to get the declarations
 * and is linked against octree.c compiled where it lies (oracle/Makefile,
 * outputs only under oracle/_ref/).  It lets tests/bench drive
 *   octree_create / octree_insert_point / octree_insert_path /
 *   octree_remove_point (octree.c L55-218) and octree_trace_line (L341-537)
 * over numpy arrays.
 */
#include "octree.c"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
    #include <omp.h>
#endif

octree_t* qbref_tree_create(float basesize, int levels)
{
    octree_t* t = malloc(sizeof(octree_t));
    *t          = octree_create((v4_t){0.0f, basesize, basesize, basesize}, levels);
    return t;
}

void qbref_tree_delete(octree_t* t)
{
    octree_delete(t);
    free(t);
}

void qbref_tree_reset(octree_t* t) { octree_reset(t, t->basecube); }

/* modelutil.c L203-217: insert points in file order, model index = point index */
void qbref_tree_insert_points(octree_t* t, const float* pts, int64_t n, int64_t first_model_index)
{
    for (int64_t i = 0; i < n; i++)
        octree_insert_point(t, 0, (size_t) (first_model_index + i), (v3_t){pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]},
                            NULL);
}

/* single insert returning the touched-node list (modelutil.c L486-501) */
void qbref_tree_insert_point(octree_t* t, const float* pnt, int64_t modind, int32_t* octindarr13)
{
    int arr[14] = {0};
    octree_insert_point(t, 0, (size_t) modind, (v3_t){pnt[0], pnt[1], pnt[2]}, arr);
    for (int i = 0; i < 13; i++) octindarr13[i] = arr[i];
}

/* qubatron.c L439-452: paths[i*12 + level] */
void qbref_tree_insert_paths(octree_t* t, const int32_t* paths, int64_t n, int64_t first_model_index)
{
    for (int64_t i = 0; i < n; i++)
    {
        int p[12];
        for (int k = 0; k < 12; k++) p[k] = paths[i * 12 + k];
        octree_insert_path(t, 0, (size_t) (first_model_index + i), p, p + 4, p + 8);
    }
}

void qbref_tree_remove_point(octree_t* t, const float* pnt, int32_t* modind, int32_t* octind)
{
    int m = -1, o = -1;
    octree_remove_point(t, (v3_t){pnt[0], pnt[1], pnt[2]}, &m, &o);
    *modind = m;
    *octind = o;
}

int64_t        qbref_tree_len(octree_t* t) { return (int64_t) t->len; }
const int32_t* qbref_tree_data(octree_t* t) { return (const int32_t*) t->octs; }

/* octree_trace_line over n rays; out_index[i] = returned oct[8], out_tlf optional */
void qbref_trace_batch(octree_t* t, int64_t n, const float* pos, const float* dir, int32_t* out_index, float* out_tlf,
                       int threads)
{
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < n; i++)
    {
        v4_t tlf     = {0};
        out_index[i] = octree_trace_line(t, (v3_t){pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]},
                                         (v3_t){dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]}, &tlf);
        if (out_tlf)
        {
            out_tlf[i * 4 + 0] = tlf.x;
            out_tlf[i * 4 + 1] = tlf.y;
            out_tlf[i * 4 + 2] = tlf.z;
            out_tlf[i * 4 + 3] = tlf.w;
        }
    }
}
