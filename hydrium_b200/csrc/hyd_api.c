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
 *   - with hydb_encoder_set_batch(n > 1) tiles are encoded n at a time; hyd_flush then returns
 *     HYD_OK with nothing written until a batch completes.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/hydrium_b200.h"

#define TILE 256u
/* worst-case staged bytes of one tile: 256 rows x 256 px x 4 samples x 2 bytes */
#define TILE_STAGE_BYTES (256u * 256u * 4u * 4u) /* worst case: RGBA float */
/* device output bytes reserved per queued tile (the engine reports overflow, never overruns) */
#define TILE_OUT_BYTES (768u * 1024u)

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

    /* bytes produced but not yet handed to the caller */
    uint8_t *pend;
    size_t pend_len, pend_cap, pend_pos;

    int wrote_header;
    int last_tile;

    double stage_ms, batch_t0;   /* HYDRIUM_B200_APITRACE accounting */

    /* suggested ICC profile, already in the form the image header codes (libhydrium.c:242-305) */
    uint8_t *icc;
    size_t icc_size;

    /* GPU side */
    int device;
    uint32_t batch;
    HydbEngine *engine;
    uint32_t slots;        /* engine workspace slots: max(batch, 1 + groups_per_tile) */
    size_t stage_cap;      /* bytes of each staging buffer */
    uint8_t *stage_host;   /* page-locked */
    uint8_t *stage_dev;
    uint8_t *out_dev;      /* batch * TILE_OUT_BYTES */
    HydbTile *tiles;
    uint32_t queued;
    size_t stage_used;
    char errbuf[256];

    /* one-frame mode over several LF groups (encoder.c:752-1011): every hyd_send_tile encodes one
     * 2048x2048 LF group as a frame part; nothing surfaces until the last one (libhydrium.c:147-166) */
    uint32_t of_n, of_cx;           /* LF groups of the image, per row */
    uint32_t of_nsent;
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

/* One engine + staging set is kept alive across encoders (creating the CUDA workspace and the
 * page-locked staging costs far more than encoding an image): hyd_encoder_destroy parks it here,
 * the next encoder with the same device / batch takes it back.  Distinct encoders may live on
 * different threads, hence the mutex. */
static struct {
    pthread_mutex_t lock;
    int valid, device;
    uint32_t batch, slots;
    size_t stage_cap;
    HydbEngine *engine;
    uint8_t *stage_host, *stage_dev, *out_dev;
    HydbTile *tiles;
} g_parked = {PTHREAD_MUTEX_INITIALIZER, 0, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL};

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
    enc->batch = env_u32("HYDRIUM_B200_BATCH", 1);
    return enc;
}

static void release_gpu(HYDEncoder *enc) {
    if (enc->engine && enc->stage_host && enc->stage_dev && enc->out_dev && enc->tiles) {
        pthread_mutex_lock(&g_parked.lock);
        if (!g_parked.valid) {
            g_parked.valid = 1;
            g_parked.device = enc->device;
            g_parked.batch = enc->batch;
            g_parked.slots = enc->slots;
            g_parked.stage_cap = enc->stage_cap;
            g_parked.engine = enc->engine;
            g_parked.stage_host = enc->stage_host;
            g_parked.stage_dev = enc->stage_dev;
            g_parked.out_dev = enc->out_dev;
            g_parked.tiles = enc->tiles;
            enc->engine = NULL;
            enc->stage_host = enc->stage_dev = enc->out_dev = NULL;
            enc->tiles = NULL;
        }
        pthread_mutex_unlock(&g_parked.lock);
    }
    if (enc->stage_host) hydb_host_free(enc->stage_host);
    if (enc->stage_dev) hydb_device_free(enc->stage_dev);
    if (enc->out_dev) hydb_device_free(enc->out_dev);
    if (enc->engine) hydb_engine_destroy(enc->engine);
    free(enc->tiles);
    enc->stage_host = enc->stage_dev = enc->out_dev = NULL;
    enc->engine = NULL;
    enc->tiles = NULL;
    enc->queued = 0;
    enc->stage_used = 0;
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
    enc->of_n = enc->of_nsent = enc->of_ngroups = enc->of_elen_cap = enc->of_max_alpha = 0;
    enc->of_e_len = enc->of_e_cap = 0;
}

HYDRIUM_EXPORT HYDStatusCode hyd_encoder_destroy(HYDEncoder *enc) { /* libhydrium.c:21-44 */
    if (!enc)
        return HYD_OK;
    release_gpu(enc);
    free(enc->pend);
    free(enc->icc);
    of_reset(enc);
    free(enc);
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_batch(HYDEncoder *enc, uint32_t tiles) {
    if (!enc || !tiles || tiles > 65536 || enc->engine) {
        if (enc)
            enc->error = "batch size must be set before the first tile";
        return HYD_API_ERROR;
    }
    enc->batch = tiles;
    return HYD_OK;
}

HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_device(HYDEncoder *enc, int device) {
    if (!enc || enc->engine) {
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

HYDRIUM_EXPORT HYDStatusCode hyd_flush(HYDEncoder *enc) { /* libhydrium.c:147-166 */
    if (enc->one_frame && !enc->last_tile)
        return HYD_OK;
    if (!enc->out) {
        enc->error = "buffer was never provided";
        return HYD_API_ERROR;
    }
    size_t n = enc->out_len - enc->out_pos;
    if (n > enc->pend_len - enc->pend_pos)
        n = enc->pend_len - enc->pend_pos;
    memcpy(enc->out + enc->out_pos, enc->pend + enc->pend_pos, n);
    enc->out_pos += n;
    enc->pend_pos += n;
    if (enc->pend_pos >= enc->pend_len) {
        enc->pend_pos = enc->pend_len = 0;
        return HYD_OK;
    }
    return HYD_NEED_MORE_OUTPUT;
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

/* HYDRIUM_B200_APITRACE=1: where a hyd_send_tile spends its time (stderr, milliseconds) */
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

static HYDStatusCode pend_reserve(HYDEncoder *enc, size_t extra) {
    if (enc->pend_len + extra <= enc->pend_cap)
        return HYD_OK;
    size_t cap = enc->pend_cap ? enc->pend_cap : 4096;
    while (cap < enc->pend_len + extra)
        cap *= 2;
    uint8_t *p = realloc(enc->pend, cap);
    if (!p)
        return HYD_NOMEM;
    enc->pend = p;
    enc->pend_cap = cap;
    return HYD_OK;
}

static HYDStatusCode gpu_error(HYDEncoder *enc, HYDStatusCode rc) {
    const char *msg = enc->engine ? hydb_engine_error(enc->engine) : "CUDA engine unavailable";
    strncpy(enc->errbuf, msg, sizeof(enc->errbuf) - 1);
    enc->errbuf[sizeof(enc->errbuf) - 1] = 0;
    enc->error = enc->errbuf;
    return rc < HYD_ERROR_START ? rc : HYD_INTERNAL_ERROR;
}

static HYDStatusCode ensure_gpu(HYDEncoder *enc) {
    if (enc->engine)
        return HYD_OK;
    pthread_mutex_lock(&g_parked.lock);
    enc->slots = enc->groups_per_tile > 1 && enc->batch < 1 + enc->groups_per_tile ? 1 + enc->groups_per_tile : enc->batch;
    enc->stage_cap = (size_t)enc->batch * TILE_STAGE_BYTES;
    if (enc->groups_per_tile > 1 && enc->stage_cap < (size_t)enc->tile_w * enc->tile_h * 16u)
        enc->stage_cap = (size_t)enc->tile_w * enc->tile_h * 16u;   /* one whole tile, RGBA float */
    HydbEngine *stale_engine = NULL;
    uint8_t *stale_host = NULL, *stale_dev = NULL, *stale_out = NULL;
    HydbTile *stale_tiles = NULL;
    if (g_parked.valid && g_parked.device == enc->device && g_parked.batch == enc->batch &&
        g_parked.slots == enc->slots && g_parked.stage_cap == enc->stage_cap) {
        g_parked.valid = 0;
        enc->engine = g_parked.engine;
        enc->stage_host = g_parked.stage_host;
        enc->stage_dev = g_parked.stage_dev;
        enc->out_dev = g_parked.out_dev;
        enc->tiles = g_parked.tiles;
    } else if (g_parked.valid) {
        /* a parked set of another shape: drop it, so that the one created now can be parked in turn
         * (the most recent geometry is the one most likely to come back) and its memory is returned */
        g_parked.valid = 0;
        stale_engine = g_parked.engine;
        stale_host = g_parked.stage_host;
        stale_dev = g_parked.stage_dev;
        stale_out = g_parked.out_dev;
        stale_tiles = g_parked.tiles;
    }
    pthread_mutex_unlock(&g_parked.lock);
    if (stale_engine) {
        hydb_host_free(stale_host);
        hydb_device_free(stale_dev);
        hydb_device_free(stale_out);
        hydb_engine_destroy(stale_engine);
        free(stale_tiles);
    }
    if (enc->engine)
        return HYD_OK;
    HYDStatusCode rc = hydb_engine_create(&enc->engine, enc->device, enc->slots);
    if (rc < HYD_ERROR_START) {
        enc->error = "could not create the CUDA engine (no usable GPU? this encoder has no CPU path)";
        return rc;
    }
    enc->stage_host = hydb_host_alloc(enc->stage_cap);
    enc->stage_dev = hydb_device_alloc(enc->stage_cap);
    enc->out_dev = hydb_device_alloc((size_t)enc->slots * TILE_OUT_BYTES);
    enc->tiles = calloc(enc->slots, sizeof(HydbTile));
    if (!enc->stage_host || !enc->stage_dev || !enc->out_dev || !enc->tiles) {
        release_gpu(enc);
        enc->error = "out of memory allocating tile staging";
        return HYD_NOMEM;
    }
    return HYD_OK;
}

/* Device output -> the pending-output queue.  The queue is ordinary (pageable, growing) memory, and a
 * device-to-pageable copy runs at a fraction of PCIe speed; the page-locked staging buffer is idle once
 * the kernels have consumed the tile pixels, so the bytes take that way when they fit. */
static int fetch_output(HYDEncoder *enc, size_t bytes) {
    uint8_t *dst = enc->pend + enc->pend_len;
    if (bytes <= enc->stage_cap) {
        if (hydb_memcpy_d2h(enc->stage_host, enc->out_dev, bytes))
            return -1;
        memcpy(dst, enc->stage_host, bytes);
        return 0;
    }
    return hydb_memcpy_d2h(dst, enc->out_dev, bytes);
}

/* encode everything queued and move the frames to the pending-output queue */
/* ICC-tagged image: the image header (with the entropy-coded profile) goes out ahead of the first
 * frame, from its own kernel; the frame paths then run as if the header had been written already */
static HYDStatusCode emit_icc_header(HYDEncoder *enc) {
    const size_t cap = enc->icc_size * 3 + 16384;
    HYDStatusCode rc = pend_reserve(enc, cap);
    if (rc < HYD_ERROR_START)
        return rc;
    uint64_t len = 0;
    rc = hydb_engine_icc_header(enc->engine, (uint32_t)enc->metadata.width, (uint32_t)enc->metadata.height, enc->icc,
                                (uint32_t)enc->icc_size, enc->pend + enc->pend_len, cap, &len);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    enc->pend_len += (size_t)len;
    enc->wrote_header = 1;
    return HYD_OK;
}

static HYDStatusCode run_batch(HYDEncoder *enc) {
    if (!enc->queued)
        return HYD_OK;
    const double t0 = now_ms();
    if (hydb_memcpy_h2d(enc->stage_dev, enc->stage_host, enc->stage_used))
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    const double t1 = now_ms();
    HYDStatusCode rc = hydb_engine_encode_tiles(enc->engine, enc->tiles, enc->queued, enc->out_dev,
                                                (uint64_t)enc->slots * TILE_OUT_BYTES, 0);
    if (rc < HYD_ERROR_START)
        return gpu_error(enc, rc);
    uint64_t bytes = 0;
    rc = hydb_engine_finish(enc->engine, &bytes);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    const double t2 = now_ms();
    rc = pend_reserve(enc, (size_t)bytes);
    if (rc < HYD_ERROR_START)
        return rc;
    if (fetch_output(enc, (size_t)bytes))
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    if (api_trace()) {
        fprintf(stderr, "[hydrium_b200] batch of %u tiles: staging %.2f (since the previous batch %.2f)  h2d %.2f  gpu %.2f  d2h %.2f ms\n",
                enc->queued, enc->stage_ms, t0 - enc->batch_t0, t1 - t0, t2 - t1, now_ms() - t2);
        enc->stage_ms = 0;
        enc->batch_t0 = now_ms();
    }
    enc->pend_len += (size_t)bytes;
    enc->queued = 0;
    enc->stage_used = 0;
    return HYD_OK;
}

/* copy one tile's samples into staging and fill its device-side descriptor */
static void stage_pixels(HYDEncoder *enc, uint32_t w, uint32_t h, const void **plane, int64_t *out_row_stride,
                         int64_t *out_pixel_stride, const void *const buffer[3], ptrdiff_t row_stride,
                         ptrdiff_t pixel_stride, size_t item) {
    struct { const void *plane[3]; int64_t row_stride, pixel_stride; } tt, *t = &tt;
    uint8_t *dst = enc->stage_host + enc->stage_used;
    uint8_t *ddst = enc->stage_dev + enc->stage_used;
    const uint8_t *p[3] = {buffer[0], buffer[1], buffer[2]};
    const uint8_t *lo = p[0] < p[1] ? (p[0] < p[2] ? p[0] : p[2]) : (p[1] < p[2] ? p[1] : p[2]);
    const uint8_t *hi = p[0] > p[1] ? (p[0] > p[2] ? p[0] : p[2]) : (p[1] > p[2] ? p[1] : p[2]);
    size_t used;
    if (pixel_stride > 0 && pixel_stride <= 4 && (size_t)(hi - lo) < (size_t)pixel_stride * item) {
        /* interleaved (RGB / RGBA ...): copy whole row spans, keep the caller's sample offsets */
        const size_t span = (size_t)w * (size_t)pixel_stride * item;
        for (uint32_t y = 0; y < h; y++)
            memcpy(dst + (size_t)y * span, lo + (ptrdiff_t)y * row_stride * (ptrdiff_t)item, span);
        for (int k = 0; k < 3; k++)
            t->plane[k] = ddst + (p[k] - lo);
        t->row_stride = (int64_t)w * pixel_stride;
        t->pixel_stride = pixel_stride;
        used = span * h;
    } else if (pixel_stride == 1) {
        /* planar: three contiguous planes */
        const size_t span = (size_t)w * item;
        for (int k = 0; k < 3; k++) {
            uint8_t *pd = dst + (size_t)k * span * h;
            for (uint32_t y = 0; y < h; y++)
                memcpy(pd + (size_t)y * span, p[k] + (ptrdiff_t)y * row_stride * (ptrdiff_t)item, span);
            t->plane[k] = ddst + (size_t)k * span * h;
        }
        t->row_stride = w;
        t->pixel_stride = 1;
        used = 3 * span * h;
    } else {
        /* anything else: gather to packed RGB */
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const ptrdiff_t o = ((ptrdiff_t)y * row_stride + (ptrdiff_t)x * pixel_stride) * (ptrdiff_t)item;
                uint8_t *q = dst + ((size_t)y * w + x) * 3 * item;
                for (int k = 0; k < 3; k++)
                    memcpy(q + k * item, p[k] + o, item);
            }
        for (int k = 0; k < 3; k++)
            t->plane[k] = ddst + k * item;
        t->row_stride = (int64_t)w * 3;
        t->pixel_stride = 3;
        used = (size_t)w * h * 3 * item;
    }
    enc->stage_used += (used + 255) & ~(size_t)255;
    for (int k = 0; k < 3; k++)
        plane[k] = tt.plane[k];
    *out_row_stride = tt.row_stride;
    *out_pixel_stride = tt.pixel_stride;
}

static void stage_tile(HYDEncoder *enc, HydbTile *t, const void *const buffer[3], ptrdiff_t row_stride,
                       ptrdiff_t pixel_stride, size_t item) {
    stage_pixels(enc, t->width, t->height, t->plane, &t->row_stride, &t->pixel_stride, buffer, row_stride, pixel_stride,
                 item);
}

/* a tile of several 256x256 groups: one frame with shared sections (k_frame.cu), encoded at once */
static HYDStatusCode run_frame(HYDEncoder *enc, HydbFrame *fr) {
    const double t0 = now_ms();
    if (hydb_memcpy_h2d(enc->stage_dev, enc->stage_host, enc->stage_used))
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    const double t1 = now_ms();
    HYDStatusCode rc = hydb_engine_encode_frames(enc->engine, fr, 1, enc->out_dev,
                                                 (uint64_t)enc->slots * TILE_OUT_BYTES, 0);
    if (rc < HYD_ERROR_START)
        return gpu_error(enc, rc);
    const double t2 = now_ms();
    uint64_t bytes = 0;
    rc = hydb_engine_finish(enc->engine, &bytes);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    const double t3 = now_ms();
    rc = pend_reserve(enc, (size_t)bytes);
    if (rc < HYD_ERROR_START)
        return rc;
    if (fetch_output(enc, (size_t)bytes))
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    if (api_trace())
        fprintf(stderr, "[hydrium_b200] frame %ux%u: h2d %.2f  launch %.2f  wait %.2f  d2h %.2f ms (%llu bytes)\n", fr->width,
                fr->height, t1 - t0, t2 - t1, t3 - t2, now_ms() - t3, (unsigned long long)bytes);
    enc->pend_len += (size_t)bytes;
    enc->stage_used = 0;
    return HYD_OK;
}

static uint32_t cllog2_u32(uint32_t v) {
    uint32_t n = 0;
    while ((1u << n) < v)
        n++;
    return n;
}

/* one LF group of a one-frame image with several of them: encode it as a frame part and keep what it
 * produced; when it is the last one, assemble the whole frame into the pending-output queue */
static HYDStatusCode run_lf_part(HYDEncoder *enc, HydbFrame *fr, uint32_t lfid) {
    if (enc->of_seen[lfid] || enc->of_nsent >= enc->of_n) {
        enc->error = "LF group sent twice in one-frame mode";
        return HYD_API_ERROR;
    }
    const uint32_t G = ((fr->width + TILE - 1) / TILE) * ((fr->height + TILE - 1) / TILE);
    fr->lf_part = 1;
    fr->preset = lfid;
    fr->preset_bits = cllog2_u32(enc->of_n);
    fr->alpha_floor = enc->of_max_alpha;
    fr->clusters_per_preset = enc->of_n * 9 <= 256 ? 9 : (enc->of_n * 3 <= 256 ? 3 : (enc->of_n * 2 <= 256 ? 2 : 1)); /* encoder.c:862-899 */
    fr->with_image_header = 0;
    if (hydb_memcpy_h2d(enc->stage_dev, enc->stage_host, enc->stage_used))
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    enc->stage_used = 0;
    HYDStatusCode rc = hydb_engine_encode_frames(enc->engine, fr, 1, enc->out_dev, (uint64_t)enc->slots * TILE_OUT_BYTES, 0);
    if (rc < HYD_ERROR_START)
        return gpu_error(enc, rc);
    uint64_t bytes = 0;
    rc = hydb_engine_finish(enc->engine, &bytes);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    uint32_t lens[65];
    rc = hydb_engine_frame_lengths(enc->engine, lens, 1 + G);
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
    if (hydb_memcpy_d2h(lf, enc->out_dev, lens[0]) ||
        hydb_memcpy_d2h(enc->of_e + enc->of_e_len, enc->out_dev + lens[0], (size_t)(bytes - lens[0]))) {
        free(lf);
        return gpu_error(enc, HYD_INTERNAL_ERROR);
    }
    enc->of_lf[k] = lf;
    enc->of_len1[k] = lens[0];
    enc->of_e_len += (size_t)(bytes - lens[0]);
    for (uint32_t g = 0; g < G; g++)
        enc->of_elen[enc->of_ngroups + g] = lens[1 + g];
    enc->of_ngroups += G;
    uint32_t *hist = enc->of_hist + (size_t)lfid * 385;
    uint32_t alpha = 0;
    rc = hydb_engine_read_model(enc->engine, 1, hist + 1, &hist[0], &alpha);
    if (rc != HYD_OK)
        return gpu_error(enc, rc);
    if (alpha > enc->of_max_alpha)
        enc->of_max_alpha = alpha;
    enc->of_sent[k] = lfid;
    enc->of_seen[lfid] = 1;
    enc->of_nsent = k + 1;
    if (!enc->last_tile)
        return HYD_OK;

    /* last LF group: head and HFGlobal from the device, then  head | LFGroups | HFGlobal | PassGroups */
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
    rc = hydb_oneframe_finish(enc->engine, info, (uint32_t)words, head, head_cap, &head_len, hf, hf_cap, &hf_len);
    free(info);
    if (rc != HYD_OK) {
        free(head); free(hf);
        return gpu_error(enc, rc);
    }
    size_t total = (size_t)head_len + hf_len + enc->of_e_len;
    for (uint32_t i = 0; i < n; i++)
        total += enc->of_len1[i];
    rc = pend_reserve(enc, total);
    if (rc < HYD_ERROR_START) {
        free(head); free(hf);
        return rc;
    }
    uint8_t *d = enc->pend + enc->pend_len;
    memcpy(d, head, head_len);
    d += head_len;
    for (uint32_t i = 0; i < n; i++) {
        memcpy(d, enc->of_lf[i], enc->of_len1[i]);
        d += enc->of_len1[i];
    }
    memcpy(d, hf, hf_len);
    d += hf_len;
    memcpy(d, enc->of_e, enc->of_e_len);
    enc->pend_len += total;
    enc->wrote_header = 1;
    free(head);
    free(hf);
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
    HYDStatusCode rc = ensure_gpu(enc);
    if (rc < HYD_ERROR_START)
        return rc;
    if (enc->icc && !enc->wrote_header) { /* encoder.c:490-494 with encoder.c:203-236 */
        rc = emit_icc_header(enc);
        if (rc < HYD_ERROR_START)
            return rc;
    }
    const uint32_t tw = (uint32_t)(((uint64_t)tile_x + 1) * span_x > W ? W - (uint64_t)tile_x * span_x : span_x);
    const uint32_t th = (uint32_t)(((uint64_t)tile_y + 1) * span_y > H ? H - (uint64_t)tile_y * span_y : span_y);
    /* encoder.c:482-485 */
    enc->last_tile = is_last < 0 ? (((uint64_t)tile_x + 1) * span_x >= W && ((uint64_t)tile_y + 1) * span_y >= H) : !!is_last;
    const size_t item = sample_fmt == HYD_UINT8 ? 1 : (sample_fmt == HYD_UINT16 ? 2 : 4);
    if (enc->of_n) {
        /* one-frame mode over several LF groups: this tile is LF group (tile_x, tile_y) */
        HydbFrame fr;
        memset(&fr, 0, sizeof(fr));
        fr.width = tw;
        fr.height = th;
        fr.x0 = tile_x * enc->tile_w;
        fr.y0 = tile_y * enc->tile_h;
        fr.image_width = (uint32_t)W;
        fr.image_height = (uint32_t)H;
        fr.is_last = 1;
        fr.sample_fmt = sample_fmt;
        fr.linear_light = enc->metadata.linear_light != 0;
        fr.one_frame = 1;
        stage_pixels(enc, tw, th, fr.plane, &fr.row_stride, &fr.pixel_stride, buffer, row_stride, pixel_stride, item);
        return run_lf_part(enc, &fr, tile_y * enc->of_cx + tile_x);
    }
    if (tw > TILE || th > TILE) {
        /* several groups in this frame: everything queued so far goes first, then the frame at once */
        rc = run_batch(enc);
        if (rc < HYD_ERROR_START)
            return rc;
        HydbFrame fr;
        memset(&fr, 0, sizeof(fr));
        fr.width = tw;
        fr.height = th;
        fr.x0 = tile_x * enc->tile_w;
        fr.y0 = tile_y * enc->tile_h;
        fr.image_width = (uint32_t)W;
        fr.image_height = (uint32_t)H;
        fr.is_last = enc->one_frame || enc->last_tile;
        fr.sample_fmt = sample_fmt;
        fr.linear_light = enc->metadata.linear_light != 0;
        fr.with_image_header = !enc->wrote_header;
        fr.one_frame = enc->one_frame;
        enc->wrote_header = 1;
        const double ts = now_ms();
        stage_pixels(enc, tw, th, fr.plane, &fr.row_stride, &fr.pixel_stride, buffer, row_stride, pixel_stride, item);
        if (api_trace())
            fprintf(stderr, "[hydrium_b200] staging %.2f ms\n", now_ms() - ts);
        return run_frame(enc, &fr);
    }

    HydbTile *t = &enc->tiles[enc->queued];
    memset(t, 0, sizeof(*t));
    t->width = tw;
    t->height = th;
    t->x0 = tile_x * enc->tile_w;
    t->y0 = tile_y * enc->tile_h;
    t->image_width = (uint32_t)W;
    t->image_height = (uint32_t)H;
    t->is_last = enc->one_frame || enc->last_tile; /* encoder.c:339 */
    t->sample_fmt = sample_fmt;
    t->linear_light = enc->metadata.linear_light != 0;
    t->with_image_header = !enc->wrote_header; /* encoder.c:490-494: the image header precedes the first frame */
    enc->wrote_header = 1;
    if (api_trace()) {
        const double ts = now_ms();
        if (!enc->queued && enc->batch_t0 == 0)
            enc->batch_t0 = ts;
        stage_tile(enc, t, buffer, row_stride, pixel_stride, item);
        enc->stage_ms += now_ms() - ts;
    } else {
        stage_tile(enc, t, buffer, row_stride, pixel_stride, item);
    }
    enc->queued++;

    if (enc->queued == enc->batch || enc->last_tile) {
        rc = run_batch(enc);
        if (rc < HYD_ERROR_START)
            return rc;
    }
    return HYD_OK; /* never HYD_NEED_MORE_OUTPUT, like the reference (libhydrium.c:195-202) */
}
