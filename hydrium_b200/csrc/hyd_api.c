/*
 * hydrium_b200/csrc/hyd_api.c -- the nine libhydrium entry points, portable C99.
 *
 * Host side of the drop-in boundary: encoder object, argument validation with the reference's
 * status codes and error strings, the output-buffer hand-off protocol, tile staging, and the
 * calls into the C-ABI CUDA layer (engine.cu, hydb_* in include/hydrium_b200.h).  No pixel is
 * touched here except to copy it into page-locked staging memory; all codec work is on the GPU
 * and there is no CPU fallback.
 *
 * Reference behaviour mirrored (file:line of the reference in each function):
 *   - libhydrium.c:16-203   lifecycle, metadata validation, buffer protocol, hyd_send_tile
 *   - encoder.c:437-508     tile bounds, tile size at image edges, last-tile rule, header once
 * Differences, by design (DESIGN.md "Boundary"):
 *   - every byte (image header, frame header, TOC, payload) surfaces through hyd_flush, in send
 *     order; the reference writes the header fields straight into the lent buffer during
 *     hyd_send_tile.  The concatenation of all surfaced bytes is identical.
 *   - hyd_send_tile is asynchronous by default: it copies the tile into page-locked staging memory and
 *     returns; tiles are encoded a chunk (32 tiles, or one multi-group frame) at a time by engine jobs
 *     that run while the caller stages the next ones, and hyd_flush hands out whatever has finished, in
 *     send order.  Everything is delivered by the end of the flush loop that follows the last tile
 *     (libhydrium.c:147-166), or by a second consecutive hyd_flush.  hydb_encoder_set_batch(1) /
 *     HYDRIUM_B200_BATCH=1 restores the reference's timing: every tile's bytes are ready when its
 *     hyd_send_tile returns.  Float tiles are always encoded synchronously (error timing, format.c:123-126).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/hydrium_b200.h"
#include "stage_pool.h"

/* Staging copy.  The destination is page-locked memory that only the GPU's copy engine reads afterwards, so
 * on x86 the rows are written with non-temporal stores: no read-for-ownership of the destination lines and
 * no cache pollution (about 1.7x the rate of memcpy for 768-byte rows out of a 12 KB-strided image).
 * stage_fence() orders them before the DMA is queued.  Elsewhere: plain memcpy. */
#if defined(__SSE2__)
#include <emmintrin.h>
static void stage_copy(uint8_t *d, const uint8_t *s, size_t n) {
    if (n < 256) {
        memcpy(d, s, n);
        return;
    }
    const size_t head = (size_t)(-(uintptr_t)d & 15u);
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    size_t k = n >> 6;
    while (k--) {
        const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16)),
                      c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
        _mm_stream_si128((__m128i *)d, a);
        _mm_stream_si128((__m128i *)(d + 16), b);
        _mm_stream_si128((__m128i *)(d + 32), c);
        _mm_stream_si128((__m128i *)(d + 48), e);
        s += 64;
        d += 64;
    }
    memcpy(d, s, n & 63u);
}
static void stage_fence(void) { _mm_sfence(); }
#else
static void stage_copy(uint8_t *d, const uint8_t *s, size_t n) { memcpy(d, s, n); }
static void stage_fence(void) {}
#endif

#define TILE 256u
/* device bytes reserved per workspace slot when a job's frames have to be gathered again (worst case) */
#define TILE_OUT_BYTES (768u * 1024u)
#define MAX_CHUNKS 8

enum { CH_FREE = 0, CH_FILLING, CH_INFLIGHT, CH_DONE };

typedef struct Chunk {
    uint8_t *stage_host, *stage_dev, *out_host;
    HydbFrame *frames;       /* [units] descriptors of what is staged (a classic tile = a frame of one group) */
    uint32_t nframes;
    size_t used;             /* staged bytes */
    uint32_t slot0;          /* first engine workspace slot of this chunk */
    int state;
    uint32_t job, job_slots;
    double t_submit;
    /* one-frame mode over several LF groups: the chunk holds one LF group (frame part) */
    int is_lf_part, closes_image;
    uint32_t lfid, groups;
} Chunk;

typedef struct Seg {
    uint8_t *heap;           /* owned block, or NULL when p points into a chunk's output area */
    const uint8_t *p;
    size_t len, pos;
    int chunk;               /* chunk to recycle once consumed, -1 for heap blocks */
} Seg;

typedef struct Gpu {
    int device;
    uint32_t nchunks, units, slots_per_chunk;
    size_t chunk_cap, out_cap;
    HydbEngine *engine;
    uint8_t *stage_host, *stage_dev, *out_host;
    HydbFrame *frames;
} Gpu;

struct HYDEncoder {
    HYDImageMetadata metadata;
    int have_metadata;
    int one_frame;
    uint32_t tile_w, tile_h;        /* pixels one hyd_send_tile covers (2048 in one-frame mode) */
    uint32_t groups_per_tile;       /* 256x256 groups in a full tile: > 1 means multi-group frames */
    const char *error;

    /* lent output buffer (libhydrium.c:114-145) */
    uint8_t *out;
    size_t out_len, out_pos;

    /* bytes produced but not yet handed to the caller, in send order */
    Seg *segs;
    size_t seg_head, seg_tail, seg_cap;

    int wrote_header;
    int last_tile;
    int flush_idle;                 /* the previous call was a hyd_flush that had nothing more to give */
    uint64_t tiles_sent, tiles_total;
    HYDStatusCode async_rc;         /* first error of an asynchronous job: sticky */

    /* suggested ICC profile, already in the form the image header codes (libhydrium.c:242-305) */
    uint8_t *icc;
    size_t icc_size;

    /* GPU side */
    int device;
    uint32_t batch;                 /* 0 = automatic chunks, 1 = synchronous per tile, n = chunks of n tiles */
    uint32_t depth;                 /* chunks in flight (0 = default) */
    uint32_t stage_workers;         /* helper threads of the staging copy (stage_pool.h; HYDRIUM_B200_THREADS) */
    uint32_t outcap_kb;             /* HYDRIUM_B200_OUTCAP_KB: output area per chunk (tests: forces the re-gather path) */
    Gpu gpu;
    Chunk chunks[MAX_CHUNKS];
    int cur;                        /* chunk being filled, -1 = none */
    uint64_t ring_head, ring_next;  /* oldest chunk not yet retired / next chunk to open (counters, index = % nchunks) */
    char errbuf[256];
    double tr_stage, tr_submit, tr_flush, tr_first, tr_last_send;   /* HYDRIUM_B200_APITRACE: where the caller's thread spent its time */

    /* one-frame mode over several LF groups (encoder.c:752-1011): every hyd_send_tile encodes one
     * 2048x2048 LF group as a frame part; nothing surfaces until the last one (libhydrium.c:147-166) */
    uint32_t of_n, of_cx;           /* LF groups of the image, per row */
    uint32_t of_nsent, of_nqueued;  /* parts collected / parts sent */
    uint32_t *of_sent;              /* raster id of the k-th LF group sent */
    uint8_t *of_seen;
    uint32_t *of_len1;              /* byte length of the k-th sent LFGroup section */
    uint8_t **of_lf;                /* ... and its bytes */
    uint32_t *of_elen;              /* PassGroup section lengths, in production order */
    uint32_t of_ngroups, of_elen_cap;
    uint8_t *of_e;                  /* PassGroup sections, concatenated */
    size_t of_e_len, of_e_cap;
    uint32_t *of_hist;              /* [of_n][385]: bit count + histogram bits of each preset */
    uint32_t of_max_alpha;
};

static HYDStatusCode chunk_submit(HYDEncoder *enc);
static int api_trace(void);
static double now_ms(void);
static void release_gpu(HYDEncoder *enc);
static void segs_clear(HYDEncoder *enc);

static uint32_t env_u32(const char *name, uint32_t fallback) {
    const char *v = getenv(name);
    if (!v || !*v)
        return fallback;
    char *end = NULL;
    unsigned long x = strtoul(v, &end, 10);
    return (end && *end == 0 && x > 0 && x <= 65536) ? (uint32_t)x : fallback;
}

HYDRIUM_EXPORT HYDEncoder *hyd_encoder_new(void) { /* libhydrium.c:16-19 */
    HYDEncoder *enc = calloc(1, sizeof(*enc));
    if (!enc)
        return NULL;
    enc->device = -1;
    const char *dev = getenv("HYDRIUM_B200_DEVICE");
    if (dev && *dev)
        enc->device = atoi(dev);
    enc->batch = env_u32("HYDRIUM_B200_BATCH", 0);
    enc->depth = env_u32("HYDRIUM_B200_DEPTH", 0);
    enc->outcap_kb = env_u32("HYDRIUM_B200_OUTCAP_KB", 0);
    enc->stage_workers = hyd_stage_default_workers();
    enc->cur = -1;
    return enc;
}

static void of_reset(HYDEncoder *enc) {
    if (enc->of_lf)
        for (uint32_t k = 0; k < enc->of_n; k++)
            free(enc->of_lf[k]);
    free(enc->of_lf);
    free(enc->of_sent);
    free(enc->of_seen);
    free(enc->of_len1);
    free(enc->of_elen);
    free(enc->of_e);
    free(enc->of_hist);
    enc->of_lf = NULL;
    enc->of_sent = enc->of_len1 = enc->of_elen = enc->of_hist = NULL;
    enc->of_seen = enc->of_e = NULL;
    enc->of_n = enc->of_nsent = enc->of_nqueued = enc->of_ngroups = enc->of_elen_cap = enc->of_max_alpha = 0;
    enc->of_e_len = enc->of_e_cap = 0;
}

HYDRIUM_EXPORT HYDStatusCode hyd_encoder_destroy(HYDEncoder *enc) { /* libhydrium.c:21-44 */
    if (!enc)
        return HYD_OK;
    if (api_trace() && enc->tr_first > 0)
        fprintf(stderr, "[hydrium_b200] encoder: first tile to destroy %.2f ms (last send at %.2f); caller's thread: staging %.2f ms, submitting %.2f ms, flushing %.2f ms\n",
                now_ms() - enc->tr_first, enc->tr_last_send - enc->tr_first, enc->tr_stage, enc->tr_submit, enc->tr_flush);
    release_gpu(enc);
    segs_clear(enc);
    free(enc->segs);
    free(enc->icc);
    of_reset(enc);
    free(enc);
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_batch(HYDEncoder *enc, uint32_t tiles) {
    if (!enc || !tiles || tiles > 65536 || enc->tiles_sent) {
        if (enc)
            enc->error = "batch size must be set before the first tile";
        return HYD_API_ERROR;
    }
    enc->batch = tiles;
    return HYD_OK;
}

HYDRIUM_EXPORT void hydb_encoder_stats(const HYDEncoder *enc, uint64_t *kernel_launches, uint64_t *graph_launches) {
    if (kernel_launches)
        *kernel_launches = enc && enc->gpu.engine ? hydb_engine_launch_count(enc->gpu.engine) : 0;
    if (graph_launches)
        *graph_launches = enc && enc->gpu.engine ? hydb_engine_graph_launch_count(enc->gpu.engine) : 0;
}

HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_device(HYDEncoder *enc, int device) {
    if (!enc || enc->tiles_sent) {
        if (enc)
            enc->error = "device must be set before the first tile";
        return HYD_API_ERROR;
    }
    enc->device = device;
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hyd_set_metadata(HYDEncoder *enc, const HYDImageMetadata *md) { /* libhydrium.c:46-112 */
    if (!md->width || !md->height) {
        enc->error = "invalid zero-width or zero-height";
        return HYD_API_ERROR;
    }
    const uint64_t w = md->width, h = md->height;
    if (w > (UINT64_C(1) << 30) || h > (UINT64_C(1) << 30)) {
        enc->error = "width or height out of bounds";
        return HYD_API_ERROR;
    }
    if (w * h > (UINT64_C(1) << 40)) {
        enc->error = "width times height out of bounds";
        return HYD_API_ERROR;
    }
    if (md->tile_size_shift_x < -1 || md->tile_size_shift_x > 3 ||
        md->tile_size_shift_y < -1 || md->tile_size_shift_y > 3) {
        enc->error = "tile_size_shift_y must be between -1 and 3"; /* sic, for x too (libhydrium.c:70-77) */
        return HYD_API_ERROR;
    }
    const int one_frame = md->tile_size_shift_x < 0 || md->tile_size_shift_y < 0;
    if (one_frame && ((w + 2047) / 2048) * ((h + 2047) / 2048) > 256) {
        /* beyond 256 LF groups several of them share an HF preset and the reference defers their ANS
         * coding until the preset is complete (encoder.c:922-926): not built */
        enc->error = "one-frame mode is limited to 256 LF groups of 2048x2048 in the B200 encoder (use tile_size_shift 0..3)";
        return HYD_API_ERROR;
    }
    enc->metadata = *md;
    enc->one_frame = one_frame;
    /* libhydrium.c:99-100, encoder.c:441-446 */
    enc->tile_w = one_frame ? 2048u : (TILE << md->tile_size_shift_x);
    enc->tile_h = one_frame ? 2048u : (TILE << md->tile_size_shift_y);
    {
        const uint64_t fw = w < enc->tile_w ? w : enc->tile_w, fh = h < enc->tile_h ? h : enc->tile_h;
        enc->groups_per_tile = (uint32_t)(((fw + TILE - 1) / TILE) * ((fh + TILE - 1) / TILE));
    }
    of_reset(enc);
    if (one_frame && (w > 2048 || h > 2048)) {
        enc->of_cx = (uint32_t)((w + 2047) / 2048);
        enc->of_n = enc->of_cx * (uint32_t)((h + 2047) / 2048);
        enc->of_sent = calloc(enc->of_n, sizeof(uint32_t));
        enc->of_seen = calloc(enc->of_n, 1);
        enc->of_len1 = calloc(enc->of_n, sizeof(uint32_t));
        enc->of_lf = calloc(enc->of_n, sizeof(uint8_t *));
        enc->of_hist = calloc((size_t)enc->of_n * 385, sizeof(uint32_t));
        if (!enc->of_sent || !enc->of_seen || !enc->of_len1 || !enc->of_lf || !enc->of_hist) {
            of_reset(enc);
            return HYD_NOMEM;
        }
    }
    enc->tiles_total = ((w + enc->tile_w - 1) / enc->tile_w) * ((h + enc->tile_h - 1) / enc->tile_h);
    enc->tiles_sent = 0;
    enc->have_metadata = 1;
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hyd_provide_output_buffer(HYDEncoder *enc, uint8_t *buffer, size_t buffer_len) {
    /* libhydrium.c:114-135 */
    if (buffer_len < 64) {
        enc->error = "provided buffer must be at least 64 bytes long";
        return HYD_API_ERROR;
    }
    if (enc->out) {
        enc->error = "buffer was already provided";
        return HYD_API_ERROR;
    }
    if (!buffer) {
        enc->error = "buffer may not be null";
        return HYD_API_ERROR;
    }
    enc->out = buffer;
    enc->out_len = buffer_len;
    enc->out_pos = 0;
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hyd_release_output_buffer(HYDEncoder *enc, size_t *written) { /* libhydrium.c:137-145 */
    if (!enc->out) {
        enc->error = "buffer was never provided";
        return HYD_API_ERROR;
    }
    *written = enc->out_pos;
    enc->out = NULL;
    return HYD_OK;
}

HYDRIUM_EXPORT const char *hyd_error_message_get(HYDEncoder *enc) { return enc->error; } /* libhydrium.c:168-170 */

/* what the profile's first 128 bytes are predicted to be (libhydrium.c:205-240): the coded header is
 * the difference, so a typical display profile turns into a run of zeros */
static uint8_t icc_header_guess(const uint8_t *h, uint32_t icc_size, unsigned i) {
    if (i < 4)
        return (uint8_t)(icc_size >> (8 * (3 - i)));
    if (i == 8)
        return 4;
    if (i >= 12 && i < 24)
        return (uint8_t)"mntrRGB XYZ "[i - 12];
    if (i >= 36 && i < 40)
        return (uint8_t)"acsp"[i - 36];
    if (i >= 41 && i < 44) {
        if (h[40] == 'A')
            return (uint8_t)"PPL"[i - 41];
        if (h[40] == 'M')
            return (uint8_t)"SFT"[i - 41];
        /* "SGI " / "SUNW": the reference indexes its two-byte strings with i - 42, which is -1 for
         * i = 41 (undefined); the format's predictor expects the vendor's second letter there */
        if (h[40] == 'S' && h[41] == 'G')
            return i == 41 ? (uint8_t)'G' : (uint8_t)"I "[i - 42];
        if (h[40] == 'S' && h[41] == 'U')
            return i == 41 ? (uint8_t)'U' : (uint8_t)"NW"[i - 42];
    }
    switch (i) {
    case 70: return 246;
    case 71: return 214;
    case 73: return 1;
    case 78: return 211;
    case 79: return 45;
    default: break;
    }
    if (i >= 80 && i < 84)
        return h[i - 76];
    return 0;
}

static size_t icc_put_varint(uint8_t *dst, uint64_t v) { /* bitwriter.c:174-180 */
    size_t n = 0;
    while (v > 0x7f) {
        dst[n++] = (uint8_t)((v & 0x7f) | 0x80);
        v >>= 7;
    }
    dst[n++] = (uint8_t)v;
    return n;
}

HYDRIUM_EXPORT HYDStatusCode hyd_set_suggested_icc_profile(HYDEncoder *enc, const uint8_t *icc_data, size_t icc_size) {
    /* libhydrium.c:242-305.  The profile is rearranged here, on the host, exactly as the reference does
     * at this call: output size, command-stream size, [one "copy the rest" command], the 128-byte
     * header as prediction residuals, the remaining bytes verbatim.  Its entropy coding into the
     * image header happens on the device when the first tile is sent (k_icc_header). */
    if (!icc_data && !icc_size) {
        free(enc->icc);
        enc->icc = NULL;
        enc->icc_size = 0;
        return HYD_OK;
    }
    if (!enc->one_frame) {
        enc->error = "one-frame mode required to set the suggested ICC profile";
        return HYD_API_ERROR;
    }
    if (!icc_size || !icc_data || icc_size > UINT32_MAX) {
        enc->error = "invalid ICC size or data buffer";
        return HYD_API_ERROR;
    }
    if (icc_size > (63u << 20)) {
        enc->error = "ICC profiles above 63 MiB are not supported by the B200 encoder";
        return HYD_API_ERROR;
    }
    uint8_t *m = malloc(icc_size + 32);
    if (!m)
        return HYD_NOMEM;
    const size_t head = icc_size < 128 ? icc_size : 128, rest = icc_size - head;
    size_t n = icc_put_varint(m, icc_size);
    unsigned log2rest = 0;
    while (rest >> (log2rest + 1))
        log2rest++;
    n += icc_put_varint(m + n, rest ? 3 + log2rest / 7 : 0);
    if (rest) {
        n += icc_put_varint(m + n, 0);   /* empty tag list */
        m[n++] = 1;                      /* command 1: copy `rest` bytes */
        n += icc_put_varint(m + n, rest);
    }
    for (unsigned i = 0; i < head; i++)
        m[n + i] = (uint8_t)(icc_data[i] - icc_header_guess(icc_data, (uint32_t)icc_size, i));
    n += head;
    memcpy(m + n, icc_data + head, rest);
    n += rest;
    free(enc->icc);
    enc->icc = m;
    enc->icc_size = n;
    return HYD_OK;
}

/* HYDRIUM_B200_APITRACE=1: what each retired chunk did (stderr, milliseconds) */
static int api_trace(void) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("HYDRIUM_B200_APITRACE");
        on = e && *e && *e != '0';
    }
    return on;
}
static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

/* ---- the queue of produced bytes --------------------------------------------------------------------
 * A segment is either a heap block (ICC image header, an assembled one-frame image, spilled or re-gathered
 * output) or the page-locked output area of a finished chunk, which the GPU's compaction kernel filled
 * directly; hyd_flush copies from the head of the queue into the caller's buffer. */
static HYDStatusCode seg_push(HYDEncoder *enc, uint8_t *heap, const uint8_t *p, size_t len, int chunk) {
    if (enc->seg_head == enc->seg_tail)
        enc->seg_head = enc->seg_tail = 0;
    if (enc->seg_tail == enc->seg_cap) {
        if (enc->seg_head) {   /* slide down */
            memmove(enc->segs, enc->segs + enc->seg_head, (enc->seg_tail - enc->seg_head) * sizeof(Seg));
            enc->seg_tail -= enc->seg_head;
            enc->seg_head = 0;
        } else {
            const size_t cap = enc->seg_cap ? enc->seg_cap * 2 : 32;
            Seg *q = realloc(enc->segs, cap * sizeof(Seg));
            if (!q) {
                free(heap);
                return HYD_NOMEM;
            }
            enc->segs = q;
            enc->seg_cap = cap;
        }
    }
    Seg *sg = &enc->segs[enc->seg_tail++];
    sg->heap = heap;
    sg->p = heap ? heap : p;
    sg->len = len;
    sg->pos = 0;
    sg->chunk = chunk;
    return HYD_OK;
}

static void segs_clear(HYDEncoder *enc) {
    for (size_t i = enc->seg_head; i < enc->seg_tail; i++)
        free(enc->segs[i].heap);
    enc->seg_head = enc->seg_tail = 0;
}

static HYDStatusCode gpu_error(HYDEncoder *enc, HYDStatusCode rc) {
    const char *msg = enc->gpu.engine ? hydb_engine_error(enc->gpu.engine) : "CUDA engine unavailable";
    strncpy(enc->errbuf, msg, sizeof(enc->errbuf) - 1);
    enc->errbuf[sizeof(enc->errbuf) - 1] = 0;
    enc->error = enc->errbuf;
    return rc < HYD_ERROR_START ? rc : HYD_INTERNAL_ERROR;
}

/* ---- GPU resources: engine + chunk ring ---------------------------------------------------------------
 * A chunk is what one asynchronous engine job works on: page-locked staging for the caller's pixels, its
 * device twin, a page-locked output area and a range of workspace slots.  hyd_send_tile fills the current
 * chunk and submits it when it is full (or holds the last tile); up to `nchunks` are in flight, so the
 * staging copy of later tiles runs while the GPU encodes earlier ones, and hyd_flush only ever copies
 * finished bytes.  Up to three sets are kept alive across encoders (creating the CUDA workspace and the
 * page-locked memory costs far more than encoding an image): hyd_encoder_destroy parks its set, the next
 * encoder with the same geometry takes it back, the least recently parked one makes room.  Distinct encoders
 * may live on different threads, hence the mutex. */
#define PARKED_MAX 3   /* e.g. tile mode, one-frame mode and a second device alternate without rebuilding */
static struct {
    pthread_mutex_t lock;
    int valid[PARKED_MAX];
    uint64_t stamp[PARKED_MAX], clock;
    Gpu gpu[PARKED_MAX];
} g_parked = {PTHREAD_MUTEX_INITIALIZER, {0}, {0}, 0, {{0}}};

static void gpu_free(Gpu *g) {
    if (g->stage_host) hydb_host_free(g->stage_host);
    if (g->out_host) hydb_host_free(g->out_host);
    if (g->stage_dev) hydb_device_free(g->stage_dev);
    if (g->engine) hydb_engine_destroy(g->engine);
    free(g->frames);
    memset(g, 0, sizeof(*g));
}

static int gpu_same_shape(const Gpu *a, const Gpu *b) {
    return a->device == b->device && a->nchunks == b->nchunks && a->units == b->units &&
           a->slots_per_chunk == b->slots_per_chunk && a->chunk_cap == b->chunk_cap && a->out_cap == b->out_cap;
}

static HYDStatusCode retire(HYDEncoder *enc, int wait, int all);

/* wait for everything in flight and move finished output out of the chunks (the ring may go away) */
static HYDStatusCode drain_chunks(HYDEncoder *enc) {
    if (!enc->gpu.engine)
        return HYD_OK;
    HYDStatusCode rc = retire(enc, 1, 1);
    for (size_t i = enc->seg_head; i < enc->seg_tail; i++) {
        Seg *sg = &enc->segs[i];
        if (sg->chunk >= 0) {
            uint8_t *h = malloc(sg->len - sg->pos ? sg->len - sg->pos : 1);
            if (!h)
                return HYD_NOMEM;
            memcpy(h, sg->p + sg->pos, sg->len - sg->pos);
            enc->chunks[sg->chunk].state = CH_FREE;
            sg->heap = h;
            sg->p = h;
            sg->len -= sg->pos;
            sg->pos = 0;
            sg->chunk = -1;
        }
    }
    return rc;
}

static void release_gpu(HYDEncoder *enc) {
    if (!enc->gpu.engine)
        return;
    /* kernels may still be reading the staging memory: let them finish; output nobody fetched is dropped */
    for (uint32_t i = 0; i < enc->gpu.nchunks; i++) {
        Chunk *c = &enc->chunks[i];
        if (c->state == CH_INFLIGHT) {
            uint64_t bytes = 0;
            hydb_engine_job_poll(enc->gpu.engine, c->job, 1, &bytes);
            hydb_engine_job_release(enc->gpu.engine, c->job);
        }
        c->state = CH_FREE;
    }
    pthread_mutex_lock(&g_parked.lock);
    Gpu stale;
    memset(&stale, 0, sizeof(stale));
    int slot = -1;
    for (int i = 0; i < PARKED_MAX; i++)
        if (!g_parked.valid[i]) {
            slot = i;
            break;
        }
    if (slot < 0) {   /* all taken: the least recently parked set goes */
        slot = 0;
        for (int i = 1; i < PARKED_MAX; i++)
            if (g_parked.stamp[i] < g_parked.stamp[slot])
                slot = i;
        stale = g_parked.gpu[slot];
    }
    g_parked.gpu[slot] = enc->gpu;
    g_parked.valid[slot] = 1;
    g_parked.stamp[slot] = ++g_parked.clock;
    pthread_mutex_unlock(&g_parked.lock);
    memset(&enc->gpu, 0, sizeof(enc->gpu));
    gpu_free(&stale);
    enc->cur = -1;
    enc->ring_head = enc->ring_next = 0;
}

/* geometry for this encoder's tile shape and sample size */
static void gpu_plan(const HYDEncoder *enc, size_t item, Gpu *g) {
    memset(g, 0, sizeof(*g));
    g->device = enc->device;
    const uint32_t depth = enc->depth;
    if (enc->groups_per_tile > 1 || enc->of_n) {
        /* frames of several groups: as many per chunk as fit ~32 workspace slots (a 512x512 tile is a prefix
         * + 4 groups: six of them per job), one when the frame is an LF group of a one-frame image (its
         * sections are collected job by job) */
        const uint64_t W = enc->metadata.width, H = enc->metadata.height;
        const uint64_t fw = W < enc->tile_w ? W : enc->tile_w, fh = H < enc->tile_h ? H : enc->tile_h;
        const uint32_t spf = 1 + enc->groups_per_tile;
        uint32_t units = enc->of_n || enc->batch == 1 ? 1u : 32u / spf;
        if (enc->batch > 1 && !enc->of_n)
            units = enc->batch;
        if (units < 1)
            units = 1;
        if (units > 64)
            units = 64;
        g->units = units;
        g->slots_per_chunk = units * spf;
        g->chunk_cap = (size_t)units * (size_t)(fw * fh) * 4u * item + 4096;
        g->out_cap = (size_t)g->slots_per_chunk * (128u << 10) + (256u << 10);
        g->nchunks = enc->batch == 1 ? 1 : (depth ? depth : (g->chunk_cap > ((size_t)40 << 20) ? 2 : (g->chunk_cap > ((size_t)12 << 20) ? 4 : 8)));
    } else {
        const uint32_t units = enc->batch ? enc->batch : 32u;
        g->units = units;
        g->slots_per_chunk = units;
        g->chunk_cap = (size_t)units * TILE * TILE * 4u * item + 4096;
        g->out_cap = (size_t)units * (160u << 10) + (64u << 10);
        uint32_t n = enc->batch == 1 ? 1 : (depth ? depth : 8);
        while (n > 2 && (uint64_t)n * units > 4096)
            n--;
        g->nchunks = n;
    }
    if (g->nchunks > MAX_CHUNKS)
        g->nchunks = MAX_CHUNKS;
    if (enc->outcap_kb)
        g->out_cap = (size_t)enc->outcap_kb << 10;
}

static HYDStatusCode ensure_gpu(HYDEncoder *enc, size_t item) {
    Gpu want;
    gpu_plan(enc, item, &want);
    if (enc->gpu.engine) {
        /* hyd_set_metadata may be called again on a used encoder (the reference reallocates its group
         * state then, libhydrium.c:79-100), and a caller may switch sample formats between tiles: when
         * the live ring is too small in any dimension, finish what is in flight and rebuild it */
        if (enc->gpu.device == want.device && enc->gpu.nchunks == want.nchunks && enc->gpu.units == want.units &&
            enc->gpu.slots_per_chunk == want.slots_per_chunk && enc->gpu.chunk_cap >= want.chunk_cap &&
            enc->gpu.out_cap >= want.out_cap)
            return HYD_OK;
        if (enc->cur >= 0 && enc->chunks[enc->cur].nframes) {
            HYDStatusCode rc = chunk_submit(enc);
            if (rc < HYD_ERROR_START)
                return rc;
        }
        HYDStatusCode rc = drain_chunks(enc);
        if (rc < HYD_ERROR_START)
            return rc;
        release_gpu(enc);
    }
    pthread_mutex_lock(&g_parked.lock);
    for (int i = 0; i < PARKED_MAX; i++)
        if (g_parked.valid[i] && gpu_same_shape(&g_parked.gpu[i], &want)) {
            g_parked.valid[i] = 0;
            enc->gpu = g_parked.gpu[i];
            break;
        }
    pthread_mutex_unlock(&g_parked.lock);
    if (!enc->gpu.engine) {
        enc->gpu = want;
        Gpu *g = &enc->gpu;
        HYDStatusCode rc = hydb_engine_create(&g->engine, g->device, g->nchunks * g->slots_per_chunk);
        if (rc < HYD_ERROR_START) {
            memset(g, 0, sizeof(*g));
            enc->error = "could not create the CUDA engine (no usable GPU? this encoder has no CPU path)";
            return rc;
        }
        g->stage_host = hydb_host_alloc(g->chunk_cap * g->nchunks);
        g->stage_dev = hydb_device_alloc(g->chunk_cap * g->nchunks);
        g->out_host = hydb_host_alloc(g->out_cap * g->nchunks);
        g->frames = calloc((size_t)g->units * g->nchunks, sizeof(HydbFrame));
        if (!g->stage_host || !g->stage_dev || !g->out_host || !g->frames) {
            gpu_free(g);
            enc->error = "out of memory allocating tile staging";
            return HYD_NOMEM;
        }
    }
    for (uint32_t i = 0; i < enc->gpu.nchunks; i++) {
        Chunk *c = &enc->chunks[i];
        memset(c, 0, sizeof(*c));
        c->stage_host = enc->gpu.stage_host + (size_t)i * enc->gpu.chunk_cap;
        c->stage_dev = enc->gpu.stage_dev + (size_t)i * enc->gpu.chunk_cap;
        c->out_host = enc->gpu.out_host + (size_t)i * enc->gpu.out_cap;
        c->frames = enc->gpu.frames + (size_t)i * enc->gpu.units;
        c->slot0 = i * enc->gpu.slots_per_chunk;
        c->state = CH_FREE;
    }
    enc->cur = -1;
    enc->ring_head = enc->ring_next = 0;
    return HYD_OK;
}

/* ---- one-frame mode over several LF groups: what a finished frame part leaves behind ------------------ */
static uint32_t cllog2_u32(uint32_t v) {
    uint32_t n = 0;
    while ((1u << n) < v)
        n++;
    return n;
}

static HYDStatusCode of_collect_part(HYDEncoder *enc, Chunk *c, const uint8_t *data, uint64_t bytes) {
    const uint32_t G = c->groups, lfid = c->lfid;
    uint32_t lens[65];
    HYDStatusCode rc = hydb_engine_slot_frame_lengths(enc->gpu.engine, c->slot0, lens, 1 + G);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    uint64_t sum = 0;
    for (uint32_t i = 0; i <= G; i++)
        sum += lens[i];
    if (sum != bytes) {
        enc->error = "inconsistent section lengths";
        return HYD_INTERNAL_ERROR;
    }
    const uint32_t k = enc->of_nsent;
    uint8_t *lf = malloc(lens[0] ? lens[0] : 1);
    if (!lf)
        return HYD_NOMEM;
    if (enc->of_e_len + (size_t)(bytes - lens[0]) > enc->of_e_cap) {
        size_t cap = enc->of_e_cap ? enc->of_e_cap * 2 : (size_t)1 << 22;
        while (cap < enc->of_e_len + (size_t)(bytes - lens[0]))
            cap *= 2;
        uint8_t *p = realloc(enc->of_e, cap);
        if (!p) {
            free(lf);
            return HYD_NOMEM;
        }
        enc->of_e = p;
        enc->of_e_cap = cap;
    }
    if (enc->of_ngroups + G > enc->of_elen_cap) {
        uint32_t cap = enc->of_elen_cap ? enc->of_elen_cap * 2 : 256;
        while (cap < enc->of_ngroups + G)
            cap *= 2;
        uint32_t *p = realloc(enc->of_elen, (size_t)cap * sizeof(uint32_t));
        if (!p) {
            free(lf);
            return HYD_NOMEM;
        }
        enc->of_elen = p;
        enc->of_elen_cap = cap;
    }
    memcpy(lf, data, lens[0]);
    memcpy(enc->of_e + enc->of_e_len, data + lens[0], (size_t)(bytes - lens[0]));
    enc->of_lf[k] = lf;
    enc->of_len1[k] = lens[0];
    enc->of_e_len += (size_t)(bytes - lens[0]);
    for (uint32_t g = 0; g < G; g++)
        enc->of_elen[enc->of_ngroups + g] = lens[1 + g];
    enc->of_ngroups += G;
    uint32_t *hist = enc->of_hist + (size_t)lfid * 385;
    uint32_t alpha = 0;
    rc = hydb_engine_read_model(enc->gpu.engine, c->slot0 + 1, hist + 1, &hist[0], &alpha);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    if (alpha > enc->of_max_alpha)
        enc->of_max_alpha = alpha;
    enc->of_sent[k] = lfid;
    enc->of_nsent = k + 1;
    return HYD_OK;
}

/* every LF group is in: head and HFGlobal from the device, then  head | LFGroups | HFGlobal | PassGroups */
static HYDStatusCode of_assemble(HYDEncoder *enc) {
    if (enc->of_nsent != enc->of_n) {
        enc->error = "one-frame mode: the last tile arrived before every LF group was sent";
        return HYD_API_ERROR;
    }
    const uint32_t n = enc->of_n, NG = enc->of_ngroups;
    size_t words = 8 + 2 * (size_t)n + NG;
    for (uint32_t p = 0; p < n; p++)
        words += 1 + ((enc->of_hist[(size_t)p * 385] + 31) >> 5);
    uint32_t *info = calloc(words, sizeof(uint32_t));
    const uint32_t head_cap = 65536 + 8 * NG, hf_cap = 8192 + 2048 * n;
    uint8_t *head = malloc(head_cap), *hf = malloc(hf_cap);
    if (!info || !head || !hf) {
        free(info); free(head); free(hf);
        return HYD_NOMEM;
    }
    info[0] = (uint32_t)enc->metadata.width;
    info[1] = (uint32_t)enc->metadata.height;
    info[2] = !enc->wrote_header;
    info[3] = enc->of_max_alpha;
    info[4] = n;
    info[5] = NG;
    memcpy(info + 8, enc->of_sent, n * sizeof(uint32_t));
    memcpy(info + 8 + n, enc->of_len1, n * sizeof(uint32_t));
    memcpy(info + 8 + 2 * n, enc->of_elen, NG * sizeof(uint32_t));
    {
        uint32_t *q = info + 8 + 2 * n + NG;
        for (uint32_t p = 0; p < n; p++) {
            const uint32_t *h = enc->of_hist + (size_t)p * 385;
            const uint32_t hw = (h[0] + 31) >> 5;
            q[0] = h[0];
            memcpy(q + 1, h + 1, hw * sizeof(uint32_t));
            q += 1 + hw;
        }
    }
    uint32_t head_len = 0, hf_len = 0;
    HYDStatusCode rc = hydb_oneframe_finish(enc->gpu.engine, info, (uint32_t)words, head, head_cap, &head_len, hf, hf_cap, &hf_len);
    free(info);
    if (rc != HYD_OK) {
        free(head); free(hf);
        return gpu_error(enc, rc);
    }
    /* head | LFGroups | HFGlobal | PassGroups go to the output queue as they are: the blocks change owner,
     * nothing is concatenated (a 4096x4096 image saved a 5.7 MB copy on its critical path) */
    enc->wrote_header = 1;
    rc = seg_push(enc, head, NULL, head_len, -1);   /* seg_push frees the block itself when it fails */
    if (rc < HYD_ERROR_START) {
        free(hf);
        return rc;
    }
    for (uint32_t i = 0; i < n; i++) {
        uint8_t *lf = enc->of_lf[i];
        enc->of_lf[i] = NULL;
        rc = seg_push(enc, lf, NULL, enc->of_len1[i], -1);
        if (rc < HYD_ERROR_START) {
            free(hf);
            return rc;
        }
    }
    rc = seg_push(enc, hf, NULL, hf_len, -1);
    if (rc < HYD_ERROR_START)
        return rc;
    uint8_t *e = enc->of_e ? enc->of_e : malloc(1);
    const size_t e_len = enc->of_e_len;
    enc->of_e = NULL;
    enc->of_e_len = enc->of_e_cap = 0;
    if (!e)
        return HYD_NOMEM;
    return seg_push(enc, e, NULL, e_len, -1);
}

/* ---- chunk life cycle ------------------------------------------------------------------------------------ */
static HYDStatusCode chunk_submit(HYDEncoder *enc) {
    Chunk *c = &enc->chunks[enc->cur];
    uint32_t job = 0, slots = 0;
    c->t_submit = now_ms();
    stage_fence();
    HYDStatusCode rc = hydb_engine_submit_frames(enc->gpu.engine, c->frames, c->nframes, c->slot0, c->stage_host, c->stage_dev,
                                                 c->used, c->out_host, enc->gpu.out_cap, &job, &slots);
    enc->tr_submit += now_ms() - c->t_submit;
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    c->job = job;
    c->job_slots = slots;
    c->state = CH_INFLIGHT;
    enc->cur = -1;
    return HYD_OK;
}

/* frames that did not fit the chunk's output area: gather them again into a device buffer sized for the
 * worst case and fetch that (never seen with integer samples: the area holds 128+ KB per group) */
static HYDStatusCode chunk_regather(HYDEncoder *enc, Chunk *c, uint8_t **blk, uint64_t *len) {
    const size_t cap = (size_t)c->job_slots * TILE_OUT_BYTES;
    uint8_t *d = hydb_device_alloc(cap);
    if (!d)
        return HYD_NOMEM;
    uint64_t bytes = 0;
    HYDStatusCode rc = hydb_engine_job_regather(enc->gpu.engine, c->job, d, cap, &bytes);
    uint8_t *h = rc == HYD_OK ? malloc(bytes ? (size_t)bytes : 1) : NULL;
    if (rc == HYD_OK && !h)
        rc = HYD_NOMEM;
    if (rc == HYD_OK && hydb_memcpy_d2h(h, d, (size_t)bytes))
        rc = HYD_INTERNAL_ERROR;
    hydb_device_free(d);
    if (rc != HYD_OK) {
        free(h);
        return rc == HYD_NOMEM ? rc : gpu_error(enc, rc);
    }
    *blk = h;
    *len = bytes;
    return HYD_OK;
}

/* move finished chunks, oldest first, to the output queue.  wait: block for the oldest one in flight;
 * all: keep going (and blocking, if wait) until nothing is in flight */
static HYDStatusCode retire(HYDEncoder *enc, int wait, int all) {
    while (enc->ring_head != enc->ring_next) {
        Chunk *c = &enc->chunks[enc->ring_head % enc->gpu.nchunks];
        if (c->state == CH_FILLING)
            break;   /* the newest chunk, not submitted yet */
        if (c->state != CH_INFLIGHT) {   /* done and queued (or consumed): nothing to do for it here */
            enc->ring_head++;
            continue;
        }
        uint64_t bytes = 0;
        HYDStatusCode rc = hydb_engine_job_poll(enc->gpu.engine, c->job, wait, &bytes);
        if (rc == HYD_DEFAULT)
            return HYD_OK;   /* still running */
        if (api_trace())
            fprintf(stderr, "[hydrium_b200] chunk %u: %u frame(s), %zu bytes staged, %.2f ms from submit to retire, %llu bytes out\n",
                    (unsigned)(enc->ring_head % enc->gpu.nchunks), c->nframes, c->used, now_ms() - c->t_submit, (unsigned long long)bytes);
        uint8_t *blk = NULL;
        if (rc < HYD_ERROR_START)
            rc = gpu_error(enc, rc);   /* a tile of the job failed on the device: take the engine's message */
        if (rc == HYD_NEED_MORE_OUTPUT) {
            rc = chunk_regather(enc, c, &blk, &bytes);
            if (rc == HYD_OK && !c->is_lf_part) {
                rc = seg_push(enc, blk, NULL, (size_t)bytes, -1);
                blk = NULL;
                hydb_engine_job_release(enc->gpu.engine, c->job);
                c->state = CH_FREE;
                enc->ring_head++;
                if (rc < HYD_ERROR_START) {
                    enc->async_rc = rc;
                    return rc;
                }
                if (!all)
                    wait = 0;
                continue;
            }
        }
        if (rc == HYD_OK && c->is_lf_part) {
            const double t0 = api_trace() ? now_ms() : 0;
            rc = of_collect_part(enc, c, blk ? blk : c->out_host, bytes);
            free(blk);
            hydb_engine_job_release(enc->gpu.engine, c->job);
            c->state = CH_FREE;
            const double t1 = api_trace() ? now_ms() : 0;
            if (rc == HYD_OK && c->closes_image)
                rc = of_assemble(enc);
            if (api_trace())
                fprintf(stderr, "[hydrium_b200] LF group %u: sections collected in %.3f ms%s%.3f ms (at %.2f ms)\n", c->lfid, t1 - t0,
                        c->closes_image ? ", frame head assembled in " : ", -", now_ms() - t1, now_ms() - enc->tr_first);
        } else if (rc == HYD_OK) {
            hydb_engine_job_release(enc->gpu.engine, c->job);
            c->state = CH_DONE;
            rc = seg_push(enc, NULL, c->out_host, (size_t)bytes, (int)(enc->ring_head % enc->gpu.nchunks));
        } else {
            free(blk);
            hydb_engine_job_release(enc->gpu.engine, c->job);
            c->state = CH_FREE;
            if (rc >= HYD_ERROR_START) {
                enc->error = "internal: unexpected job status";
                rc = HYD_INTERNAL_ERROR;
            }
        }
        enc->ring_head++;
        if (rc < HYD_ERROR_START) {
            enc->async_rc = rc;
            return rc;
        }
        if (!all)
            wait = 0;   /* one blocking retirement was asked for; take whatever else is ready */
    }
    return HYD_OK;
}

/* the chunk tiles are staged into; NULL with *rc set on failure */
static Chunk *chunk_current(HYDEncoder *enc, HYDStatusCode *rc) {
    *rc = HYD_OK;
    if (enc->cur >= 0)
        return &enc->chunks[enc->cur];
    const uint32_t idx = enc->ring_next % enc->gpu.nchunks;
    Chunk *c = &enc->chunks[idx];
    if (c->state == CH_INFLIGHT) {   /* the ring is full: wait for the oldest job */
        *rc = retire(enc, 1, 0);
        if (*rc < HYD_ERROR_START)
            return NULL;
    }
    if (c->state == CH_DONE) {
        /* finished but not yet fetched by the caller: move its bytes to the heap so the chunk can go round */
        for (size_t i = enc->seg_head; i < enc->seg_tail; i++) {
            Seg *sg = &enc->segs[i];
            if (sg->chunk == (int)idx) {
                uint8_t *h = malloc(sg->len - sg->pos ? sg->len - sg->pos : 1);
                if (!h) {
                    *rc = HYD_NOMEM;
                    return NULL;
                }
                memcpy(h, sg->p + sg->pos, sg->len - sg->pos);
                sg->heap = h;
                sg->p = h;
                sg->len -= sg->pos;
                sg->pos = 0;
                sg->chunk = -1;
            }
        }
        c->state = CH_FREE;
    }
    if (c->state != CH_FREE) {
        enc->error = "internal: chunk ring out of order";
        *rc = HYD_INTERNAL_ERROR;
        return NULL;
    }
    uint8_t *sh = c->stage_host, *sd = c->stage_dev, *oh = c->out_host;
    HydbFrame *fr = c->frames;
    const uint32_t slot0 = c->slot0;
    memset(c, 0, sizeof(*c));
    c->stage_host = sh;
    c->stage_dev = sd;
    c->out_host = oh;
    c->frames = fr;
    c->slot0 = slot0;
    c->state = CH_FILLING;
    enc->cur = (int)idx;
    enc->ring_next++;
    return c;
}

/* copy one tile's samples into the chunk's staging memory; plane[] / strides describe the staged copy in
 * DEVICE memory.  Returns 0 when the tile does not fit the chunk. */
static int stage_pixels(HYDEncoder *enc, Chunk *c, uint32_t w, uint32_t h, const void **plane, int64_t *out_row_stride,
                        int64_t *out_pixel_stride, const void *const buffer[3], ptrdiff_t row_stride,
                        ptrdiff_t pixel_stride, size_t item) {
    uint8_t *dst = c->stage_host + c->used;
    uint8_t *ddst = c->stage_dev + c->used;
    const size_t room = enc->gpu.chunk_cap - c->used;
    const uint8_t *p[3] = {buffer[0], buffer[1], buffer[2]};
    const uint8_t *lo = p[0] < p[1] ? (p[0] < p[2] ? p[0] : p[2]) : (p[1] < p[2] ? p[1] : p[2]);
    const uint8_t *hi = p[0] > p[1] ? (p[0] > p[2] ? p[0] : p[2]) : (p[1] > p[2] ? p[1] : p[2]);
    size_t used;
    if (pixel_stride > 0 && pixel_stride <= 4 && (size_t)(hi - lo) < (size_t)pixel_stride * item) {
        /* interleaved (RGB / RGBA / ARGB ...): copy row spans from the lowest addressed sample to the
         * highest one of the row's last pixel -- never past it, the caller's buffer may end there */
        const size_t span = (size_t)w * (size_t)pixel_stride * item;
        const size_t need = (size_t)(w - 1) * (size_t)pixel_stride * item + (size_t)(hi - lo) + item;
        if (span * h > room)
            return 0;
        HydStageJob job = {1, h, need, {lo, NULL, NULL}, {dst, NULL, NULL}, row_stride * (ptrdiff_t)item, span, stage_copy, stage_fence};
        hyd_stage_run(&job, enc->stage_workers);
        for (int k = 0; k < 3; k++)
            plane[k] = ddst + (p[k] - lo);
        *out_row_stride = (int64_t)w * pixel_stride;
        *out_pixel_stride = pixel_stride;
        used = span * h;
    } else if (pixel_stride == 1) {
        /* planar: three contiguous planes */
        const size_t span = (size_t)w * item;
        if (3 * span * h > room)
            return 0;
        HydStageJob job = {3, h, span, {p[0], p[1], p[2]}, {dst, dst + span * h, dst + 2 * span * h},
                           row_stride * (ptrdiff_t)item, span, stage_copy, stage_fence};
        hyd_stage_run(&job, enc->stage_workers);
        for (int k = 0; k < 3; k++)
            plane[k] = ddst + (size_t)k * span * h;
        *out_row_stride = w;
        *out_pixel_stride = 1;
        used = 3 * span * h;
    } else {
        /* anything else: gather to packed RGB */
        if ((size_t)w * h * 3 * item > room)
            return 0;
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const ptrdiff_t o = ((ptrdiff_t)y * row_stride + (ptrdiff_t)x * pixel_stride) * (ptrdiff_t)item;
                uint8_t *q = dst + ((size_t)y * w + x) * 3 * item;
                for (int k = 0; k < 3; k++)
                    memcpy(q + k * item, p[k] + o, item);
            }
        for (int k = 0; k < 3; k++)
            plane[k] = ddst + k * item;
        *out_row_stride = (int64_t)w * 3;
        *out_pixel_stride = 3;
        used = (size_t)w * h * 3 * item;
    }
    c->used += (used + 255) & ~(size_t)255;
    return 1;
}

/* ICC-tagged image: the image header (with the entropy-coded profile) goes out ahead of the first
 * frame, from its own kernel; the frame paths then run as if the header had been written already */
static HYDStatusCode emit_icc_header(HYDEncoder *enc) {
    const size_t cap = enc->icc_size * 3 + 16384;
    uint8_t *blk = malloc(cap);
    if (!blk)
        return HYD_NOMEM;
    uint64_t len = 0;
    HYDStatusCode rc = hydb_engine_icc_header(enc->gpu.engine, (uint32_t)enc->metadata.width, (uint32_t)enc->metadata.height,
                                              enc->icc, (uint32_t)enc->icc_size, blk, cap, &len);
    if (rc != HYD_OK) {
        free(blk);
        return gpu_error(enc, rc);
    }
    enc->wrote_header = 1;
    return seg_push(enc, blk, NULL, (size_t)len, -1);
}

/* move queued bytes into the lent buffer, oldest first */
static void flush_copy(HYDEncoder *enc) {
    while (enc->seg_head < enc->seg_tail && enc->out_pos < enc->out_len) {
        Seg *sg = &enc->segs[enc->seg_head];
        size_t n = enc->out_len - enc->out_pos;
        if (n > sg->len - sg->pos)
            n = sg->len - sg->pos;
        memcpy(enc->out + enc->out_pos, sg->p + sg->pos, n);
        enc->out_pos += n;
        sg->pos += n;
        if (sg->pos < sg->len)
            break;
        free(sg->heap);
        if (sg->chunk >= 0)
            enc->chunks[sg->chunk].state = CH_FREE;
        enc->seg_head++;
    }
}

static int chunks_in_flight(const HYDEncoder *enc) {
    for (uint64_t k = enc->ring_head; k != enc->ring_next; k++)
        if (enc->chunks[k % enc->gpu.nchunks].state == CH_INFLIGHT)
            return 1;
    return 0;
}

HYDRIUM_EXPORT HYDStatusCode hyd_flush(HYDEncoder *enc) { /* libhydrium.c:147-166 */
    if (enc->one_frame && !enc->last_tile)
        return HYD_OK;
    if (!enc->out) {
        enc->error = "buffer was never provided";
        return HYD_API_ERROR;
    }
    if (enc->async_rc < HYD_ERROR_START)
        return enc->async_rc;
    const double t_flush0 = api_trace() ? now_ms() : 0;
    if (enc->gpu.engine) {
        /* Everything is due once the last tile is in (libhydrium.c:147-166: the caller's flush loop after
         * the last tile must surface the whole image), or when the caller asks again right after a flush
         * that had nothing more to give: no reference caller does that while it is still sending tiles, so
         * it means "I am done" -- tile subsets without an is_last are legal (libhydrium.h:235-240). */
        const int drain = enc->last_tile || enc->flush_idle;
        if (drain && enc->cur >= 0 && enc->chunks[enc->cur].nframes) {
            HYDStatusCode rc = chunk_submit(enc);
            if (rc < HYD_ERROR_START)
                return rc;
        }
        /* Hand out what has finished without waiting; when draining, block for the OLDEST job only while
         * there is room in the caller's buffer and nothing to put there, so that the caller copies chunk k
         * away while the GPU is still coding chunk k + 1 (waiting for all of them first left the whole
         * image to be copied after the last job: 0.7 ms of a 5.3 ms 4096x4096 encode). */
        for (;;) {
            HYDStatusCode rc = retire(enc, 0, 0);
            if (rc < HYD_ERROR_START)
                return rc;
            flush_copy(enc);
            if (enc->seg_head < enc->seg_tail || !drain || !chunks_in_flight(enc))
                break;
            rc = retire(enc, 1, 0);
            if (rc < HYD_ERROR_START)
                return rc;
        }
    } else {
        flush_copy(enc);
    }
    if (t_flush0 > 0)
        enc->tr_flush += now_ms() - t_flush0;
    if (enc->seg_head < enc->seg_tail)
        return HYD_NEED_MORE_OUTPUT;
    enc->flush_idle = 1;
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hyd_send_tile(HYDEncoder *enc, const void *const buffer[3], uint32_t tile_x,
                                           uint32_t tile_y, ptrdiff_t row_stride, ptrdiff_t pixel_stride,
                                           int is_last, HYDSampleFormat sample_fmt) {
    /* libhydrium.c:172-203 */
    if (sample_fmt != HYD_UINT8 && sample_fmt != HYD_UINT16 && sample_fmt != HYD_FLOAT32) {
        enc->error = "Invalid Sample Format";
        return HYD_API_ERROR;
    }
    if (!enc->have_metadata) {
        enc->error = "hyd_set_metadata must be called before hyd_send_tile";
        return HYD_API_ERROR;
    }
    /* encoder.c:437-472: bounds and the tile's real size */
    const uint64_t W = enc->metadata.width, H = enc->metadata.height;
    const uint64_t span_x = enc->tile_w, span_y = enc->tile_h;
    if (tile_x >= (W + span_x - 1) / span_x || tile_y >= (H + span_y - 1) / span_y) {
        enc->error = "tile out of bounds";
        return HYD_API_ERROR;
    }
    if (enc->async_rc < HYD_ERROR_START)
        return enc->async_rc;   /* a tile sent earlier failed on the device: the encoder is unusable, like the reference's */
    const size_t item = sample_fmt == HYD_UINT8 ? 1 : (sample_fmt == HYD_UINT16 ? 2 : 4);
    HYDStatusCode rc = ensure_gpu(enc, item);
    if (rc < HYD_ERROR_START)
        return rc;
    enc->flush_idle = 0;
    if (enc->icc && !enc->wrote_header) { /* encoder.c:490-494 with encoder.c:203-236 */
        rc = emit_icc_header(enc);
        if (rc < HYD_ERROR_START)
            return rc;
    }
    const uint32_t tw = (uint32_t)(((uint64_t)tile_x + 1) * span_x > W ? W - (uint64_t)tile_x * span_x : span_x);
    const uint32_t th = (uint32_t)(((uint64_t)tile_y + 1) * span_y > H ? H - (uint64_t)tile_y * span_y : span_y);
    /* encoder.c:482-485 */
    enc->last_tile = is_last < 0 ? (((uint64_t)tile_x + 1) * span_x >= W && ((uint64_t)tile_y + 1) * span_y >= H) : !!is_last;
    const uint32_t lfid = enc->of_n ? tile_y * enc->of_cx + tile_x : 0;
    if (enc->of_n && (enc->of_seen[lfid] || enc->of_nqueued >= enc->of_n)) {
        enc->error = "LF group sent twice in one-frame mode";
        return HYD_API_ERROR;
    }

    Chunk *c = chunk_current(enc, &rc);
    if (!c)
        return rc;
    HydbFrame *fr = &c->frames[c->nframes];
    memset(fr, 0, sizeof(*fr));
    const double t_stage0 = api_trace() ? now_ms() : 0;
    if (t_stage0 > 0 && enc->tr_first == 0)
        enc->tr_first = t_stage0;
    if (!stage_pixels(enc, c, tw, th, fr->plane, &fr->row_stride, &fr->pixel_stride, buffer, row_stride, pixel_stride, item)) {
        /* the chunk is full in bytes before it is full in tiles (wider samples than it was sized for are
         * handled by ensure_gpu; this is a mix of layouts): send it off and start the next one */
        if (!c->nframes) {
            enc->error = "internal: tile does not fit an empty staging chunk";
            return HYD_INTERNAL_ERROR;
        }
        rc = chunk_submit(enc);
        if (rc < HYD_ERROR_START)
            return rc;
        c = chunk_current(enc, &rc);
        if (!c)
            return rc;
        fr = &c->frames[c->nframes];
        memset(fr, 0, sizeof(*fr));
        if (!stage_pixels(enc, c, tw, th, fr->plane, &fr->row_stride, &fr->pixel_stride, buffer, row_stride, pixel_stride, item)) {
            enc->error = "internal: tile does not fit an empty staging chunk";
            return HYD_INTERNAL_ERROR;
        }
    }
    if (t_stage0 > 0) {
        enc->tr_last_send = now_ms();
        enc->tr_stage += enc->tr_last_send - t_stage0;
    }
    fr->width = tw;
    fr->height = th;
    fr->x0 = tile_x * enc->tile_w;
    fr->y0 = tile_y * enc->tile_h;
    fr->image_width = (uint32_t)W;
    fr->image_height = (uint32_t)H;
    fr->sample_fmt = sample_fmt;
    fr->linear_light = enc->metadata.linear_light != 0;
    fr->one_frame = enc->one_frame;
    if (enc->of_n) {
        /* one-frame mode over several LF groups: this tile is LF group (tile_x, tile_y), encoded as a
         * frame part; nothing surfaces until the last one (libhydrium.c:147-166) */
        fr->is_last = 1;
        fr->lf_part = 1;
        fr->preset = lfid;
        fr->preset_bits = cllog2_u32(enc->of_n);
        /* the running maximum of the token alphabets (entropy.c:459, 952) only matters beyond 32 tokens,
         * which integer samples never reach; float parts are encoded one at a time (below), so the value
         * is exact whenever it can matter */
        fr->alpha_floor = enc->of_max_alpha;
        fr->clusters_per_preset = enc->of_n * 9 <= 256 ? 9 : (enc->of_n * 3 <= 256 ? 3 : (enc->of_n * 2 <= 256 ? 2 : 1)); /* encoder.c:862-899 */
        c->is_lf_part = 1;
        c->lfid = lfid;
        c->groups = ((tw + TILE - 1) / TILE) * ((th + TILE - 1) / TILE);
        c->closes_image = enc->last_tile;
        enc->of_seen[lfid] = 1;
        enc->of_nqueued++;
    } else {
        fr->is_last = enc->one_frame || enc->last_tile; /* encoder.c:339 */
        fr->with_image_header = !enc->wrote_header;     /* encoder.c:490-494: the image header precedes the first frame */
        enc->wrote_header = 1;
    }
    c->nframes++;
    enc->tiles_sent++;
    /* every tile of the image has been sent once: whatever is_last said, nothing more can follow */
    const int image_complete = !enc->one_frame && enc->tiles_sent >= enc->tiles_total;
    if (image_complete)
        enc->last_tile = enc->last_tile || 1;

    const int sync = enc->batch == 1 || sample_fmt == HYD_FLOAT32;
    if (c->nframes == enc->gpu.units || enc->last_tile || sync) {
        rc = chunk_submit(enc);
        if (rc < HYD_ERROR_START)
            return rc;
        if (sync) {
            /* strict mode (hydb_encoder_set_batch(1)) and float samples: the tile is encoded before the call
             * returns, errors included ("Invalid NaN Float", format.c:123-126), exactly like the reference */
            rc = retire(enc, 1, 1);
            if (rc < HYD_ERROR_START)
                return rc;
        }
    }
    return HYD_OK; /* never HYD_NEED_MORE_OUTPUT, like the reference (libhydrium.c:195-202) */
}
