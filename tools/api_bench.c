/* tools/api_bench.c -- the reference CLI's call sequence (src/hydrium.c:402-480) over the nine hyd_* entry
 * points, in C, timed: hyd_encoder_new, hyd_set_metadata, one 1 MiB output buffer, and after every
 * hyd_send_tile the flush / release / consume / provide loop.  Links against libhydrium_b200 (or any other
 * libhydrium build); the image is the closed-form synthetic of SURVEY.md Appendix C, held in host memory.
 * Prints one JSON object.  bench.py uses it for `e2e` (the drop-in path measured without a Python loop).
 *
 *   api_bench [--width W] [--height H] [--bits 8|16] [--linear] [--rgba] [--shift S | --one-frame]
 *             [--reps N] [--warmup N] [--seed S] [--out FILE]
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <libhydrium/libhydrium.h>

static uint32_t mix32(uint32_t v) {
    v ^= v >> 16;
    v *= 0x7feb352du;
    v ^= v >> 15;
    v *= 0x846ca68bu;
    v ^= v >> 16;
    return v;
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

static void synth(void *dst, uint32_t w, uint32_t h, int bits, int channels, uint32_t seed) {
    const uint32_t maxv = bits == 8 ? 255u : 65535u;
    const uint64_t dx = w > 1 ? w - 1 : 1, dy = h > 1 ? h - 1 : 1, dxy = (uint64_t)w + h > 2 ? (uint64_t)w + h - 2 : 1;
    for (uint32_t y = 0; y < h; y++) {
        const uint32_t hy = mix32(y + 0x7F4A7C15u);
        for (uint32_t x = 0; x < w; x++) {
            const int64_t base[3] = {(int64_t)((uint64_t)x * maxv / dx), (int64_t)((uint64_t)y * maxv / dy),
                                     (int64_t)(((uint64_t)x + y) * maxv / dxy)};
            const uint32_t hx = x * 0x9E3779B1u;
            for (uint32_t c = 0; c < 3; c++) {
                const uint32_t hh = mix32(hx ^ hy ^ (c * 0x85EBCA6Bu) ^ seed);
                const int64_t n = bits == 8 ? (int64_t)((hh >> 24) & 31u) - 16 : (int64_t)((hh >> 16) & 0x1FFFu) - 4096;
                int64_t v = base[c] + n;
                v = v < 0 ? 0 : (v > (int64_t)maxv ? (int64_t)maxv : v);
                const size_t i = ((size_t)y * w + x) * (size_t)channels + c;
                if (bits == 8)
                    ((uint8_t *)dst)[i] = (uint8_t)v;
                else
                    ((uint16_t *)dst)[i] = (uint16_t)v;
            }
            if (channels == 4) {
                const size_t i = ((size_t)y * w + x) * 4u + 3u;
                if (bits == 8)
                    ((uint8_t *)dst)[i] = 255;
                else
                    ((uint16_t *)dst)[i] = 65535;
            }
        }
    }
}

typedef struct {
    uint8_t *p;
    size_t len, cap;
} Sink;

static int sink_put(Sink *s, const uint8_t *d, size_t n) {
    if (!n)
        return 0;
    if (s->len + n > s->cap) {
        size_t cap = s->cap ? s->cap : (size_t)1 << 20;
        while (cap < s->len + n)
            cap *= 2;
        uint8_t *q = realloc(s->p, cap);
        if (!q)
            return -1;
        s->p = q;
        s->cap = cap;
    }
    memcpy(s->p + s->len, d, n);
    s->len += n;
    return 0;
}

/* one image through the CLI loop; returns milliseconds or -1 */
static double encode_once(const void *pixels, uint32_t w, uint32_t h, int bits, int channels, int linear, int shift,
                          Sink *sink, uint8_t *obuf, size_t obuf_len) {
    const double t0 = now_ms();
    HYDEncoder *enc = hyd_encoder_new();
    if (!enc)
        return -1;
    HYDImageMetadata md;
    md.width = w;
    md.height = h;
    md.linear_light = linear;
    md.tile_size_shift_x = md.tile_size_shift_y = shift;
    HYDStatusCode rc = hyd_set_metadata(enc, &md);
    const uint32_t tile = shift < 0 ? 2048u : (256u << shift);
    const uint32_t ntx = (w + tile - 1) / tile, nty = (h + tile - 1) / tile;
    const size_t item = bits == 8 ? 1 : 2;
    const HYDSampleFormat fmt = bits == 8 ? HYD_UINT8 : HYD_UINT16;
    if (rc >= HYD_ERROR_START)
        rc = hyd_provide_output_buffer(enc, obuf, obuf_len);
    sink->len = 0;
    for (uint32_t ty = 0; ty < nty && rc >= HYD_ERROR_START; ty++) {
        for (uint32_t tx = 0; tx < ntx && rc >= HYD_ERROR_START; tx++) {
            const uint8_t *p = (const uint8_t *)pixels + ((size_t)ty * tile * w + (size_t)tx * tile) * (size_t)channels * item;
            const void *planes[3] = {p, p + item, p + 2 * item};
            rc = hyd_send_tile(enc, planes, tx, ty, (ptrdiff_t)w * channels, channels, -1, fmt);
            if (rc < HYD_ERROR_START)
                break;
            do {
                rc = hyd_flush(enc);
                size_t written = 0;
                if (hyd_release_output_buffer(enc, &written) < HYD_ERROR_START || sink_put(sink, obuf, written)) {
                    rc = HYD_INTERNAL_ERROR;
                    break;
                }
                if (hyd_provide_output_buffer(enc, obuf, obuf_len) < HYD_ERROR_START) {
                    rc = HYD_INTERNAL_ERROR;
                    break;
                }
            } while (rc == HYD_NEED_MORE_OUTPUT);
        }
    }
    if (rc < HYD_ERROR_START)
        fprintf(stderr, "api_bench: error %d: %s\n", (int)rc, hyd_error_message_get(enc) ? hyd_error_message_get(enc) : "?");
    hyd_encoder_destroy(enc);
    return rc < HYD_ERROR_START ? -1.0 : now_ms() - t0;
}

static int cmp_double(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int main(int argc, char **argv) {
    uint32_t w = 4096, h = 4096, seed = 0;
    int bits = 8, linear = 0, channels = 3, shift = 0, reps = 5, warmup = 2;
    const char *out_name = NULL;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const char *v = i + 1 < argc ? argv[i + 1] : "";
        if (!strcmp(a, "--width")) { w = (uint32_t)strtoul(v, NULL, 10); i++; }
        else if (!strcmp(a, "--height")) { h = (uint32_t)strtoul(v, NULL, 10); i++; }
        else if (!strcmp(a, "--bits")) { bits = atoi(v); i++; }
        else if (!strcmp(a, "--shift")) { shift = atoi(v); i++; }
        else if (!strcmp(a, "--reps")) { reps = atoi(v); i++; }
        else if (!strcmp(a, "--warmup")) { warmup = atoi(v); i++; }
        else if (!strcmp(a, "--seed")) { seed = (uint32_t)strtoul(v, NULL, 10); i++; }
        else if (!strcmp(a, "--out")) { out_name = v; i++; }
        else if (!strcmp(a, "--linear")) linear = 1;
        else if (!strcmp(a, "--rgba")) channels = 4;
        else if (!strcmp(a, "--one-frame")) shift = -1;
        else { fprintf(stderr, "api_bench: unknown option %s\n", a); return 2; }
    }
    if ((bits != 8 && bits != 16) || !w || !h || reps < 1 || reps > 1000 || warmup < 0) {
        fprintf(stderr, "api_bench: bad arguments\n");
        return 2;
    }
    const size_t item = bits == 8 ? 1 : 2, n_in = (size_t)w * h * (size_t)channels * item;
    void *pixels = malloc(n_in);
    const size_t obuf_len = (size_t)1 << 20;
    uint8_t *obuf = malloc(obuf_len);
    double *ms = malloc(sizeof(double) * (size_t)reps);
    if (!pixels || !obuf || !ms) {
        fprintf(stderr, "api_bench: not enough memory\n");
        return 1;
    }
    synth(pixels, w, h, bits, channels, seed);
    Sink sink = {NULL, 0, 0};
    const double first = encode_once(pixels, w, h, bits, channels, linear, shift, &sink, obuf, obuf_len);
    if (first < 0)
        return 1;
    for (int i = 1; i < warmup; i++)
        if (encode_once(pixels, w, h, bits, channels, linear, shift, &sink, obuf, obuf_len) < 0)
            return 1;
    double sum = 0;
    for (int i = 0; i < reps; i++) {
        ms[i] = encode_once(pixels, w, h, bits, channels, linear, shift, &sink, obuf, obuf_len);
        if (ms[i] < 0)
            return 1;
        sum += ms[i];
    }
    if (out_name) {
        FILE *f = fopen(out_name, "wb");
        if (!f || fwrite(sink.p, 1, sink.len, f) != sink.len) {
            fprintf(stderr, "api_bench: cannot write %s\n", out_name);
            return 1;
        }
        fclose(f);
    }
    if (getenv("API_BENCH_VERBOSE")) {
        fprintf(stderr, "api_bench: ms per rep:");
        for (int i = 0; i < reps; i++)
            fprintf(stderr, " %.2f", ms[i]);
        fprintf(stderr, "\n");
    }
    qsort(ms, (size_t)reps, sizeof(double), cmp_double);
    const double mean = sum / reps, mpx = (double)w * h / 1e6;
    printf("{\"width\": %u, \"height\": %u, \"bits\": %d, \"channels\": %d, \"linear\": %d, \"shift\": %d, \"reps\": %d, "
           "\"first_call_ms\": %.3f, \"ms_mean\": %.3f, \"ms_best\": %.3f, \"ms_median\": %.3f, \"ms_p90\": %.3f, \"ms_worst\": %.3f, \"mpx_per_s\": %.1f, "
           "\"bytes_in\": %zu, \"bytes_out\": %zu}\n",
           w, h, bits, channels, linear, shift, reps, first, mean, ms[0], ms[reps / 2], ms[(reps * 9) / 10 < reps ? (reps * 9) / 10 : reps - 1],
           ms[reps - 1], mpx / (mean / 1e3), n_in, sink.len);
    free(pixels);
    free(obuf);
    free(ms);
    free(sink.p);
    return 0;
}
