/*
 * host_demo.c -- a C host that drives the connector with exactly the calls the
 * reference engine makes, to show the drop-in at the C level:
 *
 *   main_init()            qubatron.c L96-205   -> octree_glc_init
 *   modelutil_load_test()  modelutil.c L78-170  -> 5-point OCTTEST scene, 4 uploads
 *   main_loop()            qubatron.c L508-548  -> dynamic octree upload + octree_glc_update
 * and then one frame of the dynamic pipeline through the device-side replacements of the reference's other GL
 * connectors (the "next" rows of DESIGN.md section 9):
 *   skeleton_glc_update()  skeleton_glc.c L222-251 + qubatron.c L439-452, L508-529 -> octree_cuc_skeleton_update
 *   particle_glc_update()  particle_glc.c L118-156                                 -> octree_cuc_particles_update
 *   presentation           octree_glc.c L308-351                                   -> octree_cuc_enable_present
 *
 * With --gpus N the same call sequence drives N GPUs (image tiles mod N, replicated octree, peer stores into the
 * first GPU's framebuffer, device-side completion fence): nothing else in the host changes.
 *
 * The scene is built with the host data model (qubatron_b200/host/qb_host.c,
 * the octree.c equivalent).  Links only against liboctree_cuc.so and libqb_host.so;
 * no CUDA headers are needed on the host side.  Prints the number of lit pixels
 * and writes frame.ppm.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

/* host_demo [--gpus N] [--same-device]
 *   --gpus N        drive N GPUs of the box from this one thread (octree_cuc_set_gpus): same calls, same frame
 *   --same-device   put all N shards on GPU 0 (how a one-GPU box exercises the multi-GPU path) */
int main(int argc, char** argv)
{
    int gpus = 1, same_device = 0;
    for (int i = 1; i < argc; i++)
    {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc)
            gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--same-device"))
            same_device = 1;
    }

    /* modelutil.c L89-110 */
    float points[15]  = {10, 690, 10, 10, 340, 10, 10, 340, 690, 10, 10, 10, 690, 10, 690};
    float normals[15] = {0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1};
    float colors[15];
    for (int i = 0; i < 15; i++) colors[i] = 1.0f;

    qb_octree* statoctr = qb_octree_create(1800.0f, 12); /* qubatron.c L126-132 */
    qb_octree* dynaoctr = qb_octree_create(1800.0f, 12);
    qb_octree_insert_points(statoctr, points, 5, 0);

    octree_glc_t rc = octree_glc_init("shaders/"); /* qubatron.c L116 */
    if (gpus > 1)
    {
        int devices[64] = {0}; /* all on the device init chose (0) when --same-device */
        if (gpus > 64) return 3;
        octree_cuc_set_gpus(&rc, gpus, same_device ? devices : NULL);
        printf("driving %d GPU shard(s) from one thread%s\n", octree_cuc_gpu_count(&rc),
               same_device ? " (all on GPU 0)" : "");
    }

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

    /* ---- one frame of the dynamic pipeline, everything after the uploads stays on the device ---- */
    enum { RING = 64, ROWS = 120 };
    const int fig_n = RING * ROWS; /* a 60-unit upright tube of points around (700, 40..100, 700) */
    float*    fig_p = malloc(sizeof(float) * 3 * fig_n);
    float*    fig_nr = malloc(sizeof(float) * 3 * fig_n);
    float*    fig_c = malloc(sizeof(float) * 3 * fig_n);
    for (int r = 0; r < ROWS; r++)
        for (int k = 0; k < RING; k++)
        {
            const float a = 6.2831853f * (float) k / RING;
            float*      q = fig_p + 3 * (r * RING + k);
            float*      m = fig_nr + 3 * (r * RING + k);
            q[0] = 700.0f + 8.0f * cosf(a), q[1] = 40.0f + 0.5f * r, q[2] = 700.0f + 8.0f * sinf(a);
            m[0] = cosf(a), m[1] = 0.0f, m[2] = sinf(a);
            fig_c[3 * (r * RING + k)] = 0.3f, fig_c[3 * (r * RING + k) + 1] = 0.9f, fig_c[3 * (r * RING + k) + 2] = 0.4f;
        }
    octree_glc_upload_texbuffer_data(&rc, fig_c, GL_FLOAT, sizeof(float) * 3 * fig_n, sizeof(float) * 3, 0,
                                     sizeof(float) * 3 * fig_n, OCTREE_GLC_BUFFER_DYNAMIC_COLOR);
    octree_cuc_skeleton_alloc_in(&rc, fig_p, fig_nr, sizeof(float) * 3 * fig_n); /* skeleton_glc_alloc_in */
    float oldbones[80] = {0}, newbones[80] = {0}; /* 20 joints; pair 0 = the tube's axis, the others far away */
    for (int j = 0; j < 20; j++)
    {
        oldbones[4 * j] = newbones[4 * j] = 100.0f + 5.0f * j; /* unused pairs: short bones far from the figure */
        oldbones[4 * j + 1] = newbones[4 * j + 1] = 1500.0f + (j & 1);
        oldbones[4 * j + 2] = newbones[4 * j + 2] = 100.0f;
        oldbones[4 * j + 3] = 1.0f;
    }
    oldbones[0] = 700.0f, oldbones[1] = 100.0f, oldbones[2] = 700.0f, oldbones[3] = 40.0f; /* top joint, effect 40 */
    oldbones[4] = 700.0f, oldbones[5] = 40.0f, oldbones[6] = 700.0f, oldbones[7] = 40.0f;  /* bottom joint */
    newbones[0] = 715.0f, newbones[1] = 98.0f, newbones[2] = 700.0f, newbones[3] = 0.3f;   /* leaning, twisted 0.3 rad */
    newbones[4] = 700.0f, newbones[5] = 40.0f, newbones[6] = 700.0f, newbones[7] = 0.0f;
    const size_t dyn_nodes = octree_cuc_skeleton_update(&rc, oldbones, newbones, fig_n, 12, 1800.0f, 1);

    enum { PARTS = 2000 };
    float* part_p = malloc(sizeof(float) * 3 * PARTS);
    float* part_s = malloc(sizeof(float) * 3 * PARTS);
    for (int i = 0; i < PARTS; i++) /* debris falling onto the OCTTEST points */
    {
        const float* target = points + 3 * (i % 5);
        part_p[3 * i] = target[0] + 0.01f * (i % 7), part_p[3 * i + 1] = target[1] + 30.0f + 0.02f * (i % 13),
                  part_p[3 * i + 2] = target[2] + 0.01f * (i % 11);
        part_s[3 * i] = 0.001f, part_s[3 * i + 1] = -1.0f, part_s[3 * i + 2] = 0.001f;
    }
    octree_cuc_particles_alloc_in(&rc, OCTREE_CUC_PARTICLES, part_p, part_s, sizeof(float) * 3 * PARTS);
    v3_t nowhere = {0.0f, 0.0f, 0.0f};
    octree_cuc_particles_update(&rc, OCTREE_CUC_PARTICLES, PARTS, 12, 1800.0f, nowhere, 20);
    const size_t parked = octree_cuc_particles_read_out(&rc, OCTREE_CUC_PARTICLES, PARTS, part_p, part_s);

    octree_cuc_enable_present(&rc, 1);
    v3_t figpos = {700.0f, 70.0f, 950.0f}; /* not on a mid plane of the root (z = 900): the reference loses such rays */
    octree_glc_update(&rc, 1200.0f, 800.0f, figpos, lookangle, 0.0f, 8, 12, 1800.0f, 0); /* half-size render, upscaled */
    int ww = 0, wh = 0;
    octree_cuc_read_window(&rc, NULL, 0, &ww, &wh);
    uint8_t* window = malloc((size_t) ww * wh * 4);
    if (!octree_cuc_read_window(&rc, window, (size_t) ww * wh * 4, NULL, NULL)) return 1;
    long figure_px = 0;
    for (long i = 0; i < (long) ww * wh; i++) figure_px += window[i * 4 + 1] > window[i * 4] && window[i * 4 + 3] == 255;
    const uint8_t* cross = window + ((size_t) (wh / 2) * ww + ww / 2) * 4;
    printf("dynamic pipeline: %zu dynamic nodes built on the device, %zu of %d particles parked, window %dx%d, "
           "%ld figure pixels, crosshair %d %d %d\n",
           dyn_nodes, parked, PARTS, ww, wh, figure_px, cross[0], cross[1], cross[2]);
    const int dynamic_ok = dyn_nodes > 1000 && parked > 0 && figure_px > 1000 && cross[0] == 255;
    free(window), free(part_p), free(part_s), free(fig_p), free(fig_nr), free(fig_c);
    octree_cuc_destroy(&rc);
    qb_octree_delete(statoctr);
    qb_octree_delete(dynaoctr);
    return lit > 0 && dynamic_ok ? 0 : 2;
}
