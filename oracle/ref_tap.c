/*
 * oracle/ref_tap.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stage taps on the UNMODIFIED reference encoder.  The reference exposes no
 * intermediate state through its public API, so this translation unit textually
 * includes the reference's encoder.c *at build time, from where it lies under
 * /root/reference* (nothing is copied into this repository) and
 *   - reaches its `static` functions (forward_dct) directly, and
 *   - intercepts three calls that hyd_encode_xyb_buffer (encoder.c:752-1017)
 *     makes into entropy.c by macro-renaming them before the include:
 *       hyd_ans_prepare_frequencies  (encoder.c:937)  -> T2/T3/T4 before, T5 after
 *       hyd_ans_write_stream_symbols (encoder.c:946)  -> T6 after
 *       hyd_ans_write_stream_header  (encoder.c:965)  -> T3b before / T5b after
 *
 * Taps:
 *   T0 XYB floats           after hyd_populate_xyb_buffer   (format.c:142-194)
 *   T1 DCT floats           after forward_dct               (encoder.c:631-668)
 *   T2 quantised ints + nz  at the prepare_frequencies hook (encoder.c:783-823)
 *   T3 LFGlobal+LFGroup bits in working_writer at that hook (encoder.c:834-843)
 *   T4 HF hybrid symbols    at that hook                    (encoder.c:689-750)
 *   T5 normalised freqs     after prepare_frequencies       (entropy.c:943-978)
 *   T6 PassGroup bit string after ans_write_stream_symbols  (entropy.c:1064-1159)
 *   T7 final bytes          via the ordinary public API (not here)
 *
 * Only tile mode with one group per frame (tile_size_shift 0/0, or a <=256x256
 * one-frame image) is supported: exactly the path SURVEY.md section 8 scopes.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bitwriter.h"
#include "encoder.h"
#include "entropy.h"
#include "format.h"
#include "internal.h"
#include "math-functions.h"
#include "memory.h"

typedef struct TapOut {
    /* caller-allocated; capacities are fixed by the caller to the 256x256 maxima */
    float    *xyb;        /* [vbh*8][vbw*8][3]                                   T0 */
    float    *dct;        /* same shape                                          T1 */
    int32_t  *quant;      /* same shape, int32 view after quantisation           T2 */
    uint8_t  *nonzeroes;  /* [1024][3]                                           T2 */
    uint32_t *hf_syms;    /* [n][4] = token, cluster, residue_bits, residue      T4 */
    uint64_t  hf_syms_cap;
    uint64_t  hf_syms_n;
    uint8_t  *lf_bits;    /* working_writer bytes incl. cache at hook 1          T3 */
    uint64_t  lf_bits_cap;
    uint64_t  lf_bitlen;
    uint32_t *freqs;      /* [9][256]                                            T5 */
    uint16_t  alphabet_sizes[16];
    uint32_t  max_alphabet_size;
    uint8_t  *ans_bits;   /* PassGroup bit string                                T6 */
    uint64_t  ans_bits_cap;
    uint64_t  ans_bitlen;
    uint8_t  *pre_bits;   /* working_writer after the ANS stream header          T5b */
    uint64_t  pre_bits_cap;
    uint64_t  pre_bitlen;
    uint32_t  vbw, vbh;
} TapOut;

static HYDEncoder *tap_encoder;
static TapOut *tap_out;

static HYDStatusCode tap_prepare_frequencies(HYDEntropyStream *stream, size_t cluster_from, size_t cluster_to,
                                             size_t symbol_from, size_t symbol_count);
static HYDStatusCode tap_write_stream_symbols(HYDEntropyStream *stream, HYDBitWriter *bw,
                                              size_t symbol_offset, size_t symbol_count);
static HYDStatusCode tap_write_stream_header(HYDEntropyStream *stream, HYDBitWriter *bw);

#define hyd_ans_prepare_frequencies tap_prepare_frequencies
#define hyd_ans_write_stream_symbols tap_write_stream_symbols
#define hyd_ans_write_stream_header tap_write_stream_header
#include REF_ENCODER_C
#undef hyd_ans_prepare_frequencies
#undef hyd_ans_write_stream_symbols
#undef hyd_ans_write_stream_header

/* copy a bit writer's drained bytes plus its pending cache bits; returns the bit length */
static uint64_t snapshot_writer(const HYDBitWriter *bw, uint8_t *dst, uint64_t cap) {
    uint64_t bits = (uint64_t)bw->buffer_pos * 8 + (uint64_t)bw->cache_bits;
    uint64_t bytes = (bits + 7) / 8;
    if (!dst || bytes > cap)
        return bits;
    memcpy(dst, bw->buffer, bw->buffer_pos);
    uint64_t cache = bw->cache;
    for (uint64_t i = bw->buffer_pos; i < bytes; i++) {
        dst[i] = cache & 0xFF;
        cache >>= 8;
    }
    return bits;
}

static HYDStatusCode tap_prepare_frequencies(HYDEntropyStream *stream, size_t cluster_from, size_t cluster_to,
                                             size_t symbol_from, size_t symbol_count) {
    HYDEncoder *enc = tap_encoder;
    TapOut *o = tap_out;
    const HYDLFGroup *lfg = &enc->lfg[0];
    size_t npix = lfg->varblock_height * lfg->varblock_width * 64;
    if (o->quant)
        memcpy(o->quant, enc->xyb, npix * sizeof(XYBEntry));
    o->lf_bitlen = snapshot_writer(&enc->working_writer, o->lf_bits, o->lf_bits_cap);
    o->hf_syms_n = stream->symbol_count;
    if (o->hf_syms && stream->symbol_count <= o->hf_syms_cap) {
        for (size_t i = 0; i < stream->symbol_count; i++) {
            o->hf_syms[4 * i + 0] = stream->symbols[i].token;
            o->hf_syms[4 * i + 1] = stream->symbols[i].cluster;
            o->hf_syms[4 * i + 2] = stream->symbols[i].residue_bits;
            o->hf_syms[4 * i + 3] = stream->symbols[i].residue;
        }
    }
    HYDStatusCode ret = hyd_ans_prepare_frequencies(stream, cluster_from, cluster_to, symbol_from, symbol_count);
    if (ret < HYD_ERROR_START)
        return ret;
    o->max_alphabet_size = stream->max_alphabet_size;
    for (size_t c = 0; c < stream->num_clusters && c < 16; c++) {
        o->alphabet_sizes[c] = stream->alphabet_sizes[c];
        if (o->freqs && c < 9) {
            for (size_t k = 0; k < stream->alphabet_sizes[c] && k < 256; k++)
                o->freqs[c * 256 + k] = stream->frequencies[c][k];
        }
    }
    return ret;
}

static HYDStatusCode tap_write_stream_symbols(HYDEntropyStream *stream, HYDBitWriter *bw,
                                              size_t symbol_offset, size_t symbol_count) {
    HYDStatusCode ret = hyd_ans_write_stream_symbols(stream, bw, symbol_offset, symbol_count);
    if (ret < HYD_ERROR_START)
        return ret;
    tap_out->ans_bitlen = snapshot_writer(bw, tap_out->ans_bits, tap_out->ans_bits_cap);
    return ret;
}

static HYDStatusCode tap_write_stream_header(HYDEntropyStream *stream, HYDBitWriter *bw) {
    HYDStatusCode ret = hyd_ans_write_stream_header(stream, bw);
    if (ret < HYD_ERROR_START)
        return ret;
    tap_out->pre_bitlen = snapshot_writer(bw, tap_out->pre_bits, tap_out->pre_bits_cap);
    return ret;
}

/*
 * Encode ONE tile with the reference and record every tap.  `out_bytes` receives the
 * bytes the public API would have produced for this call sequence (image header, if this
 * is the encoder's first tile, then the frame).  Returns a HYDStatusCode.
 */
__attribute__((visibility("default")))
int hyd_tap_encode_tile(const HYDImageMetadata *metadata, const void *const buffer[3],
                        uint32_t tile_x, uint32_t tile_y, ptrdiff_t row_stride, ptrdiff_t pixel_stride,
                        int is_last, int sample_fmt, TapOut *out,
                        uint8_t *out_bytes, uint64_t out_cap, uint64_t *out_len) {
    HYDStatusCode ret;
    HYDEncoder *enc = hyd_encoder_new();
    if (!enc)
        return HYD_NOMEM;
    tap_encoder = enc;
    tap_out = out;
    ret = hyd_set_metadata(enc, metadata);
    if (ret < HYD_ERROR_START)
        goto end;
    ret = hyd_provide_output_buffer(enc, out_bytes, out_cap);
    if (ret < HYD_ERROR_START)
        goto end;

    /* libhydrium.c:172-203, unrolled so that the statics can be tapped in between */
    ret = hyd_send_tile_pre(enc, tile_x, tile_y, is_last);
    if (ret < HYD_ERROR_START)
        goto end;
    ret = hyd_populate_xyb_buffer(enc, buffer, row_stride, pixel_stride, 0, (HYDSampleFormat)sample_fmt);
    if (ret < HYD_ERROR_START)
        goto end;
    HYDLFGroup *lfg = &enc->lfg[0];
    out->vbw = lfg->varblock_width;
    out->vbh = lfg->varblock_height;
    size_t npix = lfg->varblock_height * lfg->varblock_width * 64;
    if (out->xyb)
        memcpy(out->xyb, enc->xyb, npix * sizeof(XYBEntry));
    if (out->dct) {
        XYBEntry *saved = malloc(npix * sizeof(XYBEntry));
        if (!saved) {
            ret = HYD_NOMEM;
            goto end;
        }
        memcpy(saved, enc->xyb, npix * sizeof(XYBEntry));
        forward_dct(enc, lfg);
        memcpy(out->dct, enc->xyb, npix * sizeof(XYBEntry));
        memcpy(enc->xyb, saved, npix * sizeof(XYBEntry));
        free(saved);
    }
    if (enc->one_frame)
        enc->lfg_perm[enc->tiles_sent] = 0;
    ret = hyd_encode_xyb_buffer(enc, tile_x, tile_y);
    if (ret < HYD_ERROR_START)
        goto end;
    /* non_zeroes is a local of hyd_encode_xyb_buffer; recompute it from the quantised ints */
    if (out->nonzeroes && out->quant) {
        memset(out->nonzeroes, 0, 1024 * 3);
        const size_t stride = lfg->varblock_width * 8;
        for (size_t by = 0; by < lfg->varblock_height; by++)
            for (size_t bx = 0; bx < lfg->varblock_width; bx++)
                for (int c = 0; c < 3; c++) {
                    unsigned n = 0;
                    for (int j = 1; j < 64; j++) {
                        size_t py = by * 8 + natural_order[j].y, px = bx * 8 + natural_order[j].x;
                        n += out->quant[(py * stride + px) * 3 + c] != 0;
                    }
                    out->nonzeroes[(by * lfg->varblock_width + bx) * 3 + c] = n;
                }
    }
    {
        size_t written = 0;
        HYDStatusCode r2 = hyd_release_output_buffer(enc, &written);
        if (r2 < HYD_ERROR_START)
            ret = r2;
        *out_len = written;
    }
end:
    hyd_encoder_destroy(enc);
    tap_encoder = NULL;
    tap_out = NULL;
    return ret;
}

/* the cosine table as the reference's compiler rounded it (encoder.c:32-40) */
__attribute__((visibility("default")))
void hyd_tap_cosine_lut(float dst[56]) {
    memcpy(dst, cosine_lut, sizeof(cosine_lut));
}

/* the three runtime LUTs (format.c:58-83) for a given sample format / transfer */
__attribute__((visibility("default")))
int hyd_tap_luts(int sample_fmt, int linear_light, uint16_t *input_lut, float *bias_lut) {
    HYDEncoder *enc = hyd_encoder_new();
    if (!enc)
        return HYD_NOMEM;
    HYDImageMetadata md = {8, 8, linear_light, 0, 0};
    HYDStatusCode ret = hyd_set_metadata(enc, &md);
    static const uint16_t zeros[8 * 8 * 3];
    const void *const buf[3] = {zeros, zeros, zeros};
    uint8_t tmp[4096];
    if (ret >= HYD_ERROR_START)
        ret = hyd_provide_output_buffer(enc, tmp, sizeof(tmp));
    if (ret >= HYD_ERROR_START)
        ret = hyd_send_tile_pre(enc, 0, 0, -1);
    if (ret >= HYD_ERROR_START)
        ret = hyd_populate_xyb_buffer(enc, buf, 8, 1, 0, (HYDSampleFormat)sample_fmt);
    if (ret >= HYD_ERROR_START) {
        if (sample_fmt == HYD_UINT8)
            memcpy(input_lut, enc->input_lut8, 256 * sizeof(uint16_t));
        else
            memcpy(input_lut, enc->input_lut16, 65536 * sizeof(uint16_t));
        memcpy(bias_lut, enc->bias_cbrtf_lut, 65536 * sizeof(float));
    }
    hyd_encoder_destroy(enc);
    return ret;
}

/*
 * A prefix-coded stream straight through the reference's entropy front end
 * (entropy.c:371-425, 502-524, 1023-1034): values[i] on context ctx[i] (NULL = 0).
 * Returns the bit length, or a negative HYDStatusCode.
 */
__attribute__((visibility("default")))
int64_t hyd_tap_prefix_stream(const uint32_t *values, const uint32_t *ctx, uint64_t n,
                              const uint8_t *cluster_map, uint32_t num_dists, int custom_config,
                              int split, int msb, int lsb, uint32_t lz77_min_symbol, int modular,
                              uint8_t *dst, uint64_t cap) {
    const char *error = NULL;
    HYDEntropyStream stream;
    HYDBitWriter bw;
    static const uint8_t zeros[256];
    HYDStatusCode ret = hyd_init_bit_writer(&bw, NULL, 0, 0, 0);
    if (ret < HYD_ERROR_START)
        return ret;
    ret = hyd_entropy_init_stream(&stream, n ? n : 1, cluster_map ? cluster_map : zeros, num_dists,
                                  custom_config, lz77_min_symbol, modular, &error);
    if (ret < HYD_ERROR_START)
        goto end;
    if (custom_config)
        hyd_entropy_set_hybrid_config(&stream, 0, 0, split, msb, lsb);
    for (uint64_t i = 0; i < n; i++) {
        ret = hyd_entropy_send_symbol(&stream, ctx ? ctx[i] : 0, values[i]);
        if (ret < HYD_ERROR_START)
            goto end;
    }
    ret = hyd_prefix_finalize_stream(&stream, &bw);
    if (ret < HYD_ERROR_START)
        goto end;
    {
        int64_t bits = (int64_t)snapshot_writer(&bw, dst, cap);
        free(bw.buffer);
        return bits;
    }
end:
    free(bw.buffer);
    return ret;
}
