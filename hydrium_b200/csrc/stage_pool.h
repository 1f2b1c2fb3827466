/*
 * hydrium_b200/csrc/stage_pool.h -- the staging copy of hyd_send_tile, spread over a few host threads.
 *
 * hyd_send_tile must have copied the caller's samples before it returns (the caller may reuse its buffer,
 * libhydrium.h:213-257), and on the nine-symbol path that copy -- 50 MB of 768-byte rows for a 4096x4096
 * image, one core's worth of memory bandwidth -- was what the GPU waited for.  The pool splits a tile's rows
 * over the calling thread and up to HYD_STAGE_MAX_WORKERS helpers.  Helpers spin for a short while after
 * their last job (the next hyd_send_tile follows within microseconds) and then sleep on a condition
 * variable.  Portable C with pthreads and the GCC/Clang __atomic builtins; no codec work happens here.
 */
#ifndef HYDRIUM_B200_STAGE_POOL_H
#define HYDRIUM_B200_STAGE_POOL_H

#include <stddef.h>
#include <stdint.h>

#define HYD_STAGE_MAX_WORKERS 7
#define HYD_STAGE_MAX_PLANES 3

typedef void (*HydStageCopyFn)(uint8_t *dst, const uint8_t *src, size_t n);

typedef struct HydStageJob {
    uint32_t planes;                         /* 1 (interleaved rows) or 3 (planar) */
    uint32_t rows;                           /* rows per plane */
    size_t bytes;                            /* bytes copied per row */
    const uint8_t *src[HYD_STAGE_MAX_PLANES];
    uint8_t *dst[HYD_STAGE_MAX_PLANES];
    ptrdiff_t src_pitch;                     /* bytes between rows of the caller's buffer (may be negative) */
    size_t dst_pitch;
    HydStageCopyFn copy;                     /* row copy (non-temporal on x86) */
    void (*fence)(void);                     /* orders the copies of a thread before its completion count */
} HydStageJob;

/* Copies the job's rows and returns when every byte is in place.  workers = helpers wanted besides the
 * caller (0: the caller alone; clipped to HYD_STAGE_MAX_WORKERS).  Any number of threads may call this at
 * once: one of them gets the helpers, the others copy on their own. */
void hyd_stage_run(const HydStageJob *job, uint32_t workers);

/* helpers the pool would use by default: HYDRIUM_B200_THREADS (total copying threads, 1 = caller only),
 * else min(6, half the CPUs this process may run on) - 1 */
uint32_t hyd_stage_default_workers(void);

#endif
