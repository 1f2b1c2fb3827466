/*
 * hydrium_b200/csrc/stage_pool.c -- see stage_pool.h.  A process-wide pool of helper threads for the
 * staging copy behind hyd_send_tile (the reference copies nothing: it converts the caller's samples in
 * place, format.c:142-194; here the samples have to reach page-locked memory before the call returns).
 *
 * One job at a time.  The job's rows are handed out in blocks through one 64-bit word,
 * (sequence << 32) | (rows of the job << 16) | next row, advanced by compare-and-swap.  A claim succeeds only
 * while the word carries the sequence number the helper read the descriptor under and a next row below the
 * job's row count (taken from the word itself, never from the descriptor), i.e. while the job is
 * incomplete; the caller rewrites the descriptor only after every row of the job has been claimed and
 * copied, so a successful claim always belongs to the descriptor it was made with.
 */
#define _GNU_SOURCE
#include "stage_pool.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define cpu_relax() _mm_pause()
#else
#define cpu_relax() ((void)0)
#endif

#define BLOCK_ROWS 16u
#define SPIN_PAUSES 2000u         /* a helper spins on pause for this many iterations (the next tile is microseconds away), ... */
#define HOT_MS 5.0                /* ... then on sched_yield until this long after its last job (the GPU tail of an image; the next
                                   * image of a busy caller), and only then sleeps.  A helper that sleeps is woken next to the
                                   * caller's core and the two then share it for a scheduler tick or more -- measured as images
                                   * that stage in 25 ms instead of 3 -- so helpers of a busy encoder should not sleep at all;
                                   * yielding keeps them out of the way of other runnable threads on a crowded host. */
#define SMALL_JOB_BYTES (96u * 1024u)
#define SLOW_WAIT 2000u           /* iterations (256 pauses, then yields) the owner may wait for the helpers' last blocks; a whole tile copies in 20 us */
#define CALM_JOBS 512u            /* jobs without a slow wait before another helper is invited again */

/* The descriptor is published and read word by word with relaxed atomics: a helper that lost the race for
 * the last rows of job k may still be reading while the caller writes job k + 1; its claim then fails (see
 * above) and what it read is discarded, but the accesses themselves must not be a data race. */
#define JOB_WORDS ((sizeof(HydStageJob) + sizeof(uintptr_t) - 1) / sizeof(uintptr_t))
typedef union JobWords {
    HydStageJob j;
    uintptr_t w[JOB_WORDS];
} JobWords;

static struct {
    pthread_mutex_t owner;       /* held by the thread whose job the helpers are working on */
    pthread_mutex_t mu;          /* sleeping / waking helpers, starting threads */
    pthread_cond_t cv;
    pthread_t th[HYD_STAGE_MAX_WORKERS];
    uint32_t nthreads;
    uint64_t ticket;             /* (sequence << 32) | (rows << 16) | next unclaimed row */
    JobWords job;
    uint32_t invited;
    uint32_t done_rows;
    uint32_t sleepers;
    int stop;
    /* Oversubscribed hosts (several encoder processes, fewer free cores than threads): a helper that holds a
     * block and loses its core makes the caller wait for a scheduler quantum.  The owner of a job measures how
     * long it waited after finishing its own share; long waits lower the number of helpers invited, calm
     * stretches raise it again. */
    uint32_t cap, calm;
} P = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, {{0}}, 0, 0, 0, 0, HYD_STAGE_MAX_WORKERS, 0};

static void copy_rows(const HydStageJob *j, uint32_t first, uint32_t n) {
    for (uint32_t r = first; r < first + n; r++) {
        const uint32_t k = r / j->rows, y = r - k * j->rows;
        j->copy(j->dst[k] + (size_t)y * j->dst_pitch, j->src[k] + (ptrdiff_t)y * j->src_pitch, j->bytes);
    }
    if (j->fence)
        j->fence();
}

/* claim and copy blocks of job `seq` until none is left (or the pool has moved on) */
static void work_on(uint32_t seq) {
    JobWords jw;
    int have = 0;
    for (;;) {
        uint64_t t = __atomic_load_n(&P.ticket, __ATOMIC_ACQUIRE);
        if ((uint32_t)(t >> 32) != seq)
            return;
        if (!have) {
            for (size_t i = 0; i < JOB_WORDS; i++)
                jw.w[i] = __atomic_load_n(&P.job.w[i], __ATOMIC_RELAXED);
            have = 1;
        }
        const uint32_t cur = (uint32_t)t & 0xFFFFu, total = (uint32_t)(t >> 16) & 0xFFFFu;
        if (cur >= total)
            return;
        const uint32_t n = total - cur < BLOCK_ROWS ? total - cur : BLOCK_ROWS;
        if (!__atomic_compare_exchange_n(&P.ticket, &t, t + n, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            continue;
        copy_rows(&jw.j, cur, n);
        __atomic_add_fetch(&P.done_rows, n, __ATOMIC_RELEASE);
    }
}

static double mono_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

static void *helper_main(void *arg) {
    const uint32_t me = (uint32_t)(uintptr_t)arg;
    uint32_t seen = 0, spin = 0;
    double idle_since = 0;
    for (;;) {
        const uint64_t t = __atomic_load_n(&P.ticket, __ATOMIC_SEQ_CST);
        if ((uint32_t)(t >> 32) == seen) {
            if (__atomic_load_n(&P.stop, __ATOMIC_RELAXED))
                return NULL;
            if (++spin < SPIN_PAUSES) {
                cpu_relax();
                continue;
            }
            if (spin == SPIN_PAUSES)
                idle_since = mono_ms();
            if (mono_ms() - idle_since < HOT_MS) {
                sched_yield();
                continue;
            }
            pthread_mutex_lock(&P.mu);
            __atomic_add_fetch(&P.sleepers, 1, __ATOMIC_SEQ_CST);
            while ((uint32_t)(__atomic_load_n(&P.ticket, __ATOMIC_SEQ_CST) >> 32) == seen && !__atomic_load_n(&P.stop, __ATOMIC_RELAXED))
                pthread_cond_wait(&P.cv, &P.mu);
            __atomic_sub_fetch(&P.sleepers, 1, __ATOMIC_SEQ_CST);
            pthread_mutex_unlock(&P.mu);
            spin = 0;
            continue;
        }
        seen = (uint32_t)(t >> 32);
        spin = 0;
        if (me < __atomic_load_n(&P.invited, __ATOMIC_RELAXED))
            work_on(seen);
    }
}

/* a forked child has none of the parent's threads: start again from an empty pool */
static void after_fork_in_child(void) {
    pthread_mutex_init(&P.owner, NULL);
    pthread_mutex_init(&P.mu, NULL);
    pthread_cond_init(&P.cv, NULL);
    P.nthreads = 0;
    P.sleepers = 0;
}

static uint32_t ensure_threads(uint32_t want) {
    if (__atomic_load_n(&P.nthreads, __ATOMIC_ACQUIRE) >= want)
        return want;
    pthread_mutex_lock(&P.mu);
    static int fork_hook;
    if (!fork_hook) {
        pthread_atfork(NULL, NULL, after_fork_in_child);
        fork_hook = 1;
    }
    while (P.nthreads < want) {
        pthread_attr_t at;
        pthread_attr_init(&at);
        pthread_attr_setstacksize(&at, 256 * 1024);
        const int rc = pthread_create(&P.th[P.nthreads], &at, helper_main, (void *)(uintptr_t)P.nthreads);
        pthread_attr_destroy(&at);
        if (rc)
            break;
        __atomic_store_n(&P.nthreads, P.nthreads + 1, __ATOMIC_RELEASE);
    }
    const uint32_t n = P.nthreads < want ? P.nthreads : want;
    pthread_mutex_unlock(&P.mu);
    return n;
}

void hyd_stage_run(const HydStageJob *job, uint32_t workers) {
    const uint32_t total = job->planes * job->rows;
    if (!total)
        return;
    if (workers > HYD_STAGE_MAX_WORKERS)
        workers = HYD_STAGE_MAX_WORKERS;
    if (!workers || (size_t)total * job->bytes < SMALL_JOB_BYTES || total < 2 * BLOCK_ROWS || total > 0xFFFFu ||
        pthread_mutex_trylock(&P.owner) != 0) {
        copy_rows(job, 0, total);
        return;
    }
    if (workers > P.cap)
        workers = P.cap;
    workers = ensure_threads(workers);
    JobWords jw;
    memset(&jw, 0, sizeof(jw));
    jw.j = *job;
    for (size_t i = 0; i < JOB_WORDS; i++)
        __atomic_store_n(&P.job.w[i], jw.w[i], __ATOMIC_RELAXED);
    __atomic_store_n(&P.invited, workers, __ATOMIC_RELAXED);
    __atomic_store_n(&P.done_rows, 0, __ATOMIC_RELAXED);
    const uint32_t seq = (uint32_t)(__atomic_load_n(&P.ticket, __ATOMIC_RELAXED) >> 32) + 1;
    __atomic_store_n(&P.ticket, (uint64_t)seq << 32 | (uint64_t)total << 16, __ATOMIC_SEQ_CST);
    if (workers && __atomic_load_n(&P.sleepers, __ATOMIC_SEQ_CST)) {
        pthread_mutex_lock(&P.mu);
        pthread_cond_broadcast(&P.cv);
        pthread_mutex_unlock(&P.mu);
    }
    work_on(seq);
    uint32_t waited = 0;
    while (__atomic_load_n(&P.done_rows, __ATOMIC_ACQUIRE) != total) {
        if (++waited < 256u)
            cpu_relax();
        else
            sched_yield();   /* whoever holds the last block may be waiting for this very core */
    }
    if (waited > SLOW_WAIT) {          /* a helper was not running: invite one fewer from now on */
        P.cap = workers ? workers - 1 : 0;
        P.calm = 0;
    } else if (++P.calm >= CALM_JOBS) {   /* try one more again */
        if (P.cap < HYD_STAGE_MAX_WORKERS)
            P.cap++;
        P.calm = 0;
    }
    pthread_mutex_unlock(&P.owner);
}

uint32_t hyd_stage_default_workers(void) {
    const char *v = getenv("HYDRIUM_B200_THREADS");
    if (v && *v) {
        const long n = strtol(v, NULL, 10);
        if (n >= 1)
            return n - 1 > HYD_STAGE_MAX_WORKERS ? HYD_STAGE_MAX_WORKERS : (uint32_t)(n - 1);
    }
    int cpus = 1;
#if defined(__linux__)
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
        cpus = CPU_COUNT(&set);
#endif
    /* up to six copying threads, at most half of the CPUs this process may run on */
    int n = cpus / 2;
    if (n > 6)
        n = 6;
    return n > 1 ? (uint32_t)n - 1 : 0;
}

/* the library may be unloaded (dlclose): its code must not disappear under running helpers */
__attribute__((destructor)) static void stage_pool_shutdown(void) {
    pthread_mutex_lock(&P.mu);
    __atomic_store_n(&P.stop, 1, __ATOMIC_RELAXED);
    pthread_cond_broadcast(&P.cv);
    const uint32_t n = P.nthreads;
    pthread_mutex_unlock(&P.mu);
    for (uint32_t i = 0; i < n; i++)
        pthread_join(P.th[i], NULL);
}
