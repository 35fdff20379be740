/*
 * host_demo.c -- a C host that drives the connector with exactly the calls the
 * reference engine makes, to show the drop-in at the C level:
 *
 *   main_init()            qubatron.c L96-205   -> octree_glc_init
 *   modelutil_load_test()  modelutil.c L78-170  -> 5-point OCTTEST scene, 4 uploads
 *   main_loop()            qubatron.c L508-548  -> dynamic octree upload + octree_glc_update
 *
 * The scene is built with the host data model (qubatron_b200/host/qb_host.c,
 * the octree.c equivalent).  Links only against liboctree_cuc.so and libqb_host.so;
 * no CUDA headers are needed on the host side.  Prints the number of lit pixels
 * and writes frame.ppm.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/octree_cuc.h"

/* host data model (qb_host.c) */
typedef struct qb_octree qb_octree;
qb_octree* qb_octree_create(float basesize, int levels);
void       qb_octree_insert_points(qb_octree* t, const float* pts, int64_t n, int64_t first_modind);
int64_t    qb_octree_len(const qb_octree* t);
int32_t*   qb_octree_nodes(qb_octree* t);
void       qb_octree_delete(qb_octree* t);

#define GL_INT 0x1404
#define GL_FLOAT 0x1406

int main(void)
{
    /* modelutil.c L89-110 */
    float points[15]  = {10, 690, 10, 10, 340, 10, 10, 340, 690, 10, 10, 10, 690, 10, 690};
    float normals[15] = {0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1};
    float colors[15];
    for (int i = 0; i < 15; i++) colors[i] = 1.0f;

    qb_octree* statoctr = qb_octree_create(1800.0f, 12); /* qubatron.c L126-132 */
    qb_octree* dynaoctr = qb_octree_create(1800.0f, 12);
    qb_octree_insert_points(statoctr, points, 5, 0);

    octree_glc_t rc = octree_glc_init("shaders/"); /* qubatron.c L116 */

    /* modelutil.c L131-169 */
    octree_glc_upload_texbuffer_data(&rc, colors, GL_FLOAT, 5 * sizeof(float) * 3, sizeof(float) * 3, 0,
                                     5 * sizeof(float) * 3, OCTREE_GLC_BUFFER_STATIC_COLOR);
    octree_glc_upload_texbuffer_data(&rc, normals, GL_FLOAT, 5 * sizeof(float) * 3, sizeof(float) * 3, 0,
                                     5 * sizeof(float) * 3, OCTREE_GLC_BUFFER_STATIC_NORMAL);
    size_t sbytes = (size_t) qb_octree_len(statoctr) * sizeof(int32_t) * 12;
    octree_glc_upload_texbuffer_data(&rc, qb_octree_nodes(statoctr), GL_INT, sbytes, sizeof(int32_t) * 4, 0, sbytes,
                                     OCTREE_GLC_BUFFER_STATIC_OCTREE);
    size_t dbytes = (size_t) qb_octree_len(dynaoctr) * sizeof(int32_t) * 12;
    octree_glc_upload_texbuffer_data(&rc, qb_octree_nodes(dynaoctr), GL_INT, dbytes, sizeof(int32_t) * 4, 0, dbytes,
                                     OCTREE_GLC_BUFFER_DYNAMIC_OCTREE);

    /* qubatron.c L136, L538-548: OCTTEST camera, window 1200x800 at quality 10 */
    v3_t lookpos   = {900.0f, 900.0f, 3000.0f};
    v3_t lookangle = {0.0f, 0.0f, 0.0f};
    octree_glc_update(&rc, 1200.0f, 800.0f, lookpos, lookangle, 0.0f, 10, 12, 1800.0f, 0);

    int w = 0, h = 0;
    octree_cuc_frame_size(&rc, &w, &h);
    uint8_t* frame = malloc((size_t) w * h * 4);
    if (!octree_cuc_read_frame(&rc, frame, (size_t) w * h * 4)) return 1;

    long lit = 0;
    for (long i = 0; i < (long) w * h; i++) lit += frame[i * 4 + 3] == 255;
    printf("frame %dx%d, %ld leaf pixels, device memory %.1f MB, %.3f ms on the GPU\n", w, h, lit,
           rc.memsize_bytes / 1e6, octree_cuc_last_frame_ms(&rc));

    FILE* f = fopen("frame.ppm", "wb");
    if (f)
    {
        fprintf(f, "P6\n%d %d\n255\n", w, h);
        for (int y = h - 1; y >= 0; y--) /* row 0 = bottom */
            for (int x = 0; x < w; x++) fwrite(frame + ((size_t) y * w + x) * 4, 1, 3, f);
        fclose(f);
    }
    free(frame);
    octree_cuc_destroy(&rc);
    qb_octree_delete(statoctr);
    qb_octree_delete(dynaoctr);
    return lit > 0 ? 0 : 2;
}
