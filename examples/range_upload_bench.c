/*
 * range_upload_bench.c -- cost of the engine's "zero-and-append" upload pattern at the C level:
 * N single-node 48-byte octree_glc_upload_texbuffer_data calls (modelutil.c L429-437, L486-501) followed by
 * the frame that applies them (one H2D copy + one scatter kernel).  Prints microseconds per call and the flush.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../include/octree_cuc.h"

#define GL_INT 0x1404

static double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

int main(int argc, char** argv)
{
    const size_t nodes = 4u << 20; /* a 4 M-node static tree (192 MB host array) */
    const int    calls = argc > 1 ? atoi(argv[1]) : 5000;
    int32_t*     octs  = calloc(nodes * 12, sizeof(int32_t));
    for (size_t i = 0; i + 1 < nodes; i++) octs[i * 12 + (i & 7)] = (int32_t) (i + 1);

    octree_glc_t rc = octree_glc_init("");
    octree_glc_upload_texbuffer_data(&rc, octs, GL_INT, nodes * 48, 16, 0, nodes * 48, OCTREE_GLC_BUFFER_STATIC_OCTREE);
    octree_cuc_sync(&rc);

    v3_t pos = {700.0f, 150.0f, 350.0f}, ang = {0.4636f, 0.0f, 0.0f};
    octree_glc_update(&rc, 640.0f, 360.0f, pos, ang, 0.0f, 10, 12, 1800.0f, 0);
    octree_cuc_sync(&rc);

    for (int round = 0; round < 3; round++)
    {
        uint32_t seed = 12345u + round;
        double   t0   = now_ms();
        for (int k = 0; k < calls; k++)
        {
            seed            = seed * 1664525u + 1013904223u;
            size_t node     = seed % nodes;
            octs[node * 12] = 0; /* zero a child slot, then upload that node */
            octree_glc_upload_texbuffer_data(&rc, octs, GL_INT, nodes * 48, 16, node * 48, (node + 1) * 48,
                                             OCTREE_GLC_BUFFER_STATIC_OCTREE);
        }
        double t1 = now_ms();
        octree_glc_update(&rc, 640.0f, 360.0f, pos, ang, 0.0f, 10, 12, 1800.0f, 0); /* flushes the batch */
        octree_cuc_sync(&rc);
        double t2 = now_ms();
        printf("round %d: %d node uploads in %.3f ms (%.3f us per call); frame incl. batch flush %.3f ms\n", round,
               calls, t1 - t0, 1e3 * (t1 - t0) / calls, t2 - t1);
    }
    octree_cuc_destroy(&rc);
    free(octs);
    return 0;
}
