/* tests/host_harness/stage_pool_test.c -- CPU test driver for hydrium_b200/csrc/stage_pool.c (test tool, not
 * a product path): many staging jobs of different shapes, helper counts and callers, every byte checked. */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "stage_pool.h"

static void plain_copy(uint8_t *d, const uint8_t *s, size_t n) { memcpy(d, s, n); }

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

static uint32_t rng(uint32_t *s) {
    *s = *s * 1664525u + 1013904223u;
    return *s >> 8;
}

/* one random job; returns the number of wrong bytes */
static size_t one_job(uint32_t *seed, uint32_t workers) {
    const uint32_t planes = rng(seed) % 3 == 0 ? 3 : 1;
    const uint32_t rows = 1 + rng(seed) % 300, bytes = 1 + rng(seed) % 3000;
    const size_t pitch = bytes + rng(seed) % 5000;
    const int flip = rng(seed) & 1;
    const size_t plane_src = pitch * rows, plane_dst = (size_t)bytes * rows;
    uint8_t *src = malloc(plane_src * planes), *dst = calloc(plane_dst * planes + 1, 1);
    for (size_t i = 0; i < plane_src * planes; i++)
        src[i] = (uint8_t)rng(seed);
    HydStageJob job;
    memset(&job, 0, sizeof(job));
    job.planes = planes;
    job.rows = rows;
    job.bytes = bytes;
    job.src_pitch = flip ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch;
    job.dst_pitch = bytes;
    job.copy = plain_copy;
    for (uint32_t k = 0; k < planes; k++) {
        job.src[k] = src + k * plane_src + (flip ? pitch * (rows - 1) : 0);
        job.dst[k] = dst + k * plane_dst;
    }
    hyd_stage_run(&job, workers);
    size_t bad = dst[plane_dst * planes] != 0;
    for (uint32_t k = 0; k < planes; k++)
        for (uint32_t y = 0; y < rows; y++) {
            const uint8_t *s = src + k * plane_src + pitch * (flip ? rows - 1 - y : y);
            if (memcmp(s, dst + k * plane_dst + (size_t)y * bytes, bytes))
                bad++;
        }
    free(src);
    free(dst);
    return bad;
}

static void *caller(void *arg) {
    uint32_t seed = (uint32_t)(uintptr_t)arg;
    size_t bad = 0;
    for (int i = 0; i < 300; i++)
        bad += one_job(&seed, rng(&seed) % 8);
    return (void *)(uintptr_t)bad;
}

/* stage_pool_test bench HELPERS IMAGES: time IMAGES passes of 256 tile copies (a 4096x4096 RGB8 image) with a
 * pause between passes, print mean / median / worst -- run several at once to see oversubscription */
static int cmp_d(const void *a, const void *b) { return *(const double *)a < *(const double *)b ? -1 : 1; }
static int bench(uint32_t workers, int images) {
    const size_t w = 4096 * 3, n = w * 4096;
    uint8_t *img = malloc(n), *stage = malloc(n);
    double *ms = malloc(sizeof(double) * (size_t)images);
    memset(img, 1, n);
    memset(stage, 0, n);
    for (int k = 0; k < images; k++) {
        const double t0 = now_ms();
        for (uint32_t t = 0; t < 256; t++) {
            HydStageJob job = {1, 256, 768, {img + (size_t)(t / 16) * 256 * w + (size_t)(t % 16) * 768, NULL, NULL},
                               {stage + (size_t)t * 256 * 768, NULL, NULL}, (ptrdiff_t)w, 768, plain_copy, NULL};
            hyd_stage_run(&job, workers);
        }
        ms[k] = now_ms() - t0;
        struct timespec ts = {0, 3 * 1000 * 1000};   /* the GPU tail: helpers go to sleep */
        nanosleep(&ts, NULL);
    }
    double sum = 0;
    for (int k = 0; k < images; k++)
        sum += ms[k];
    qsort(ms, (size_t)images, sizeof(double), cmp_d);
    printf("helpers %u: mean %.2f ms, median %.2f, p90 %.2f, worst %.2f\n", workers, sum / images, ms[images / 2], ms[images * 9 / 10], ms[images - 1]);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "bench"))
        return bench((uint32_t)atoi(argv[2]), atoi(argv[3]));
    size_t bad = 0;
    uint32_t seed = 7;
    for (int i = 0; i < 400; i++)
        bad += one_job(&seed, (uint32_t)i % 9);
    /* helpers asleep between jobs */
    for (int i = 0; i < 5; i++) {
        struct timespec ts = {0, 20 * 1000 * 1000};
        nanosleep(&ts, NULL);
        bad += one_job(&seed, 3);
    }
    /* several callers at once: one gets the helpers, the others copy on their own */
    pthread_t th[4];
    for (int i = 0; i < 4; i++)
        pthread_create(&th[i], NULL, caller, (void *)(uintptr_t)(100 + i));
    for (int i = 0; i < 4; i++) {
        void *r = NULL;
        pthread_join(th[i], &r);
        bad += (size_t)(uintptr_t)r;
    }
    /* a rough rate: 256 tiles of 256 rows x 768 bytes out of a 12 KB pitch */
    const size_t w = 4096 * 3, n = w * 4096;
    uint8_t *img = malloc(n), *stage = malloc(n);
    memset(img, 1, n);
    memset(stage, 0, n);
    for (uint32_t workers = 0; workers <= 3; workers += 3) {
        const double t0 = now_ms();
        for (uint32_t t = 0; t < 256; t++) {
            HydStageJob job = {1, 256, 768, {img + (size_t)(t / 16) * 256 * w + (size_t)(t % 16) * 768, NULL, NULL},
                               {stage + (size_t)t * 256 * 768, NULL, NULL}, (ptrdiff_t)w, 768, plain_copy, NULL};
            hyd_stage_run(&job, workers);
        }
        fprintf(stderr, "stage_pool_test: 256 tiles with %u helper(s): %.2f ms\n", workers, now_ms() - t0);
    }
    bad += memcmp(img, stage, 16) != 0;
    free(img);
    free(stage);
    printf("%zu\n", bad);
    return bad != 0;
}
