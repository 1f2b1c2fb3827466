/* hydrium_b200/cli/hydrium_cli.c
 *
 * Command-line front end with the reference tool's interface (reference: src/hydrium.c:27-43 options,
 * 70-504 main): PNG or PFM in, JPEG XL out, through the nine libhydrium entry points only -- so the
 * same source links against libhydrium_b200.so (the product) or any other libhydrium (the parity
 * tests link it against the reference build to check that both emit the same file).
 *
 * Call sequence kept from the reference because it shapes the bytes:
 *   - one-frame mode unless --tile-size=N is given; --tag-icc-from needs one-frame mode
 *   - a 1 MiB output buffer; after every tile  hyd_flush / release / write / provide  until the
 *     flush stops asking for more output (hydrium.c:464-476)
 *   - PNG: bit depth <= 8 is sent as packed RGB8 (pixel stride 3), 16 bit as RGBA16 (pixel stride 4),
 *     tile rows top to bottom, is_last = -1 (the library works it out)
 *   - PFM: rows are stored bottom-up, so tile rows go bottom to top with a negative row stride, and
 *     is_last is set on the right-most tile of the top row (hydrium.c:418-461)
 * The PNG decoder is our own (png_reader.c, zlib only); the reference uses libspng.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "libhydrium/libhydrium.h"
#include "png_reader.h"

typedef struct Options {
    int one_frame, pfm, linear;
    long tile_shift;            /* -1: not given */
    const char *in_name, *out_name, *icc_name;
} Options;

static void usage(const char *argv0) {
    fprintf(stderr,
            "Usage: %s [options] [--] <input.png|input.pfm> <output.jxl>\n"
            "Options:\n"
            "    --help         Print this message\n"
            "    --one-frame    Use one frame. Uses more memory but decodes faster.\n"
            "                       (default: on)\n"
            "    --tile-size=N  Use Tile Size Shift = N, valid values are 0, 1, 2, 3\n"
            "                       Tile dimensions will be 256 * 2^N\n"
            "                       Larger tiles use more memory but decode faster.\n"
            "    --pfm          Assume input is PFM (Portable FloatMap)\n"
            "    --png          Assume input is PNG (Portable Network Graphics)\n"
            "                       (default: assume PNG unless input filename ends with .pfm)\n"
            "    --linear       Assume input is in Linear Light\n"
            "                       (default: assume sRGB transfer, regardless of PNG tags)\n"
            "    --tag-icc-from=FILE.icc\n"
            "                   Use FILE as the suggested ICC profile. Input still assumed to be sRGB.\n",
            argv0);
}

/* 0: run, 1: exit with *code */
static int parse_options(int argc, const char *argv[], Options *o, int *code) {
    memset(o, 0, sizeof(*o));
    o->pfm = -1;
    o->tile_shift = -1;
    int plain_only = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (plain_only || strncmp(a, "--", 2)) {
            if (!o->in_name)
                o->in_name = a;
            else if (!o->out_name)
                o->out_name = a;
            else {
                fprintf(stderr, "Invalid trailing arg: %s\nPlease run: %s --help\n", a, argv[0]);
                *code = 2;
                return 1;
            }
        } else if (!strcmp(a, "--help")) {
            usage(argv[0]);
            *code = 0;
            return 1;
        } else if (!a[2]) {
            plain_only = 1;
        } else if (!strcmp(a, "--one-frame")) {
            o->one_frame = 1;
        } else if (!strncmp(a, "--tile-size=", 12)) {
            errno = 0;
            o->tile_shift = strtol(a + 12, NULL, 10);
            if (errno) {
                fprintf(stderr, "Invalid integer: %s\nPlease run: %s --help\n", a + 12, argv[0]);
                *code = 2;
                return 1;
            }
            if (o->tile_shift < 0 || o->tile_shift > 3) {
                fprintf(stderr, "Invalid tile size, must be 0-3: %s\nPlease run: %s --help\n", a + 12, argv[0]);
                *code = 2;
                return 1;
            }
        } else if (!strcmp(a, "--pfm")) {
            o->pfm = 1;
        } else if (!strcmp(a, "--png")) {
            o->pfm = 0;
        } else if (!strcmp(a, "--linear")) {
            o->linear = 1;
        } else if (!strncmp(a, "--tag-icc-from=", 15)) {
            o->icc_name = a + 15;
        } else {
            fprintf(stderr, "Unknown option: %s\nPlease run: %s --help\n", a, argv[0]);
            *code = 2;
            return 1;
        }
    }
    if (!o->one_frame && o->tile_shift < 0)
        o->one_frame = 1;
    if (o->one_frame && o->tile_shift >= 0) {
        fprintf(stderr, "--one-frame and --tile-size are incompatible\nPlease run: %s --help\n", argv[0]);
        *code = 2;
        return 1;
    }
    if (!o->one_frame && o->icc_name) {
        fprintf(stderr, "--tag-icc-from= requires --one-frame\n");
        *code = 2;
        return 1;
    }
    if (o->pfm < 0) {
        const size_t n = o->in_name ? strlen(o->in_name) : 0;
        o->pfm = n > 3 && !strcmp(o->in_name + n - 4, ".pfm");
    }
    return 0;
}

/* "PF\n<width> <height>\n<scale>\n": colour PFM; a negative scale means little-endian samples.
 * Returns 0 and the byte order (1 = samples need swapping on this host), or -1 with *error. */
static int read_pfm_header(FILE *f, uint64_t *width, uint64_t *height, int *swap, const char **error) {
    char sig[3];
    if (fread(sig, 1, 3, f) != 3 || memcmp(sig, "PF\n", 3)) {
        *error = "not a color PFM file";
        return -1;
    }
    uint64_t dim[2] = {0, 0};
    const int stop[2] = {' ', '\n'};
    for (int k = 0; k < 2; k++) {
        for (;;) {   /* one character at a time: nothing behind the header may be consumed */
            const int c = fgetc(f);
            if (c >= '0' && c <= '9')
                dim[k] = dim[k] * 10 + (uint64_t)(c - '0');
            else if (c == stop[k])
                break;
            else {
                *error = k ? "invalid PFM height" : "invalid PFM width";
                return -1;
            }
            if (dim[k] > (UINT64_C(1) << 30))
                break;
        }
    }
    const int file_little = fgetc(f) == '-';
    for (int n = 0;; n++) {
        const int c = fgetc(f);
        if (c == '\n')
            break;
        if (c < 0 || n > 64) {
            *error = "invalid PFM endianness";
            return -1;
        }
    }
    const uint16_t probe = 1;
    const int host_little = *(const uint8_t *)&probe;
    *swap = file_little != host_little;
    *width = dim[0];
    *height = dim[1];
    return 0;
}

static uint8_t *read_whole_file(const char *name, size_t *len, const char *argv0) {
    FILE *f = fopen(name, "rb");
    if (!f) {
        fprintf(stderr, "%s: error opening file: %s\n", argv0, name);
        return NULL;
    }
    size_t cap = 8192, n = 0;
    uint8_t *buf = malloc(cap);
    while (buf) {
        if (cap > (size_t)INT32_MAX) {   /* an endless device file, most likely */
            fprintf(stderr, "%s: that is a very big icc profile: %s\n", argv0, name);
            free(buf);
            buf = NULL;
            break;
        }
        if (cap - n < 4096) {
            uint8_t *p = realloc(buf, cap * 2);
            if (!p) {
                free(buf);
                buf = NULL;
                break;
            }
            buf = p;
            cap *= 2;
        }
        const size_t got = fread(buf + n, 1, cap - n, f);
        n += got;
        if (got == 0) {
            if (ferror(f)) {
                fprintf(stderr, "%s: error reading from icc file\n", argv0);
                free(buf);
                buf = NULL;
            }
            break;
        }
    }
    fclose(f);
    *len = n;
    return buf;
}

/* hyd_flush / release / write / provide until the encoder has nothing more for this tile */
static int drain_output(HYDEncoder *enc, uint8_t *obuf, size_t obuf_size, FILE *fout) {
    int ret;
    do {
        ret = hyd_flush(enc);
        size_t written = 0;
        int r2 = hyd_release_output_buffer(enc, &written);
        if (r2 < HYD_ERROR_START)
            return r2;
        if (written && fwrite(obuf, written, 1, fout) != 1)
            return 1;
        r2 = hyd_provide_output_buffer(enc, obuf, obuf_size);
        if (r2 < HYD_ERROR_START)
            return r2;
    } while (ret == HYD_NEED_MORE_OUTPUT);
    return ret;
}

int main(int argc, const char *argv[]) {
    fprintf(stderr, "libhydrium version %s\n", HYDRIUM_VERSION_STRING);
    if (argc < 2) {
        usage(argv[0]);
        return 1;
    }
    Options opt;
    int code = 0;
    if (parse_options(argc, argv, &opt, &code))
        return code;

    int ret = 1;
    const char *error_msg = NULL;
    FILE *fin = stdin, *fout = stdout;
    PngReader *png = NULL;
    HYDEncoder *enc = NULL;
    uint8_t *pixels = NULL, *obuf = NULL, *icc = NULL;
    uint64_t width = 0, height = 0;
    int swap = 0;

    if (opt.in_name && strcmp(opt.in_name, "-")) {
        fin = fopen(opt.in_name, "rb");
        if (!fin) {
            fprintf(stderr, "%s: error opening file: %s\n", argv[0], opt.in_name);
            goto done;
        }
    }
    if (opt.pfm) {
        if (read_pfm_header(fin, &width, &height, &swap, &error_msg))
            goto done;
    } else {
        png = png_reader_open(fin, &error_msg);
        if (!png)
            goto done;
        width = png_reader_width(png);
        height = png_reader_height(png);
    }
    if (width > (UINT64_C(1) << 30) || height > (UINT64_C(1) << 30) || width * height > (UINT64_C(1) << 40)) {
        fprintf(stderr, "%s: buffer too big\n", argv[0]);
        goto done;
    }

    const int deep = !opt.pfm && png_reader_bit_depth(png) > 8;
    const size_t row_bytes = opt.pfm ? 12 * (size_t)width : png_reader_row_bytes(png);
    HYDImageMetadata md;
    md.width = width;
    md.height = height;
    md.linear_light = opt.linear;
    md.tile_size_shift_x = md.tile_size_shift_y = opt.one_frame ? -1 : (int)opt.tile_shift;
    const uint32_t shift = opt.one_frame ? 3 : (uint32_t)opt.tile_shift;   /* one-frame mode walks 2048x2048 LF groups */
    const uint32_t tile = 256u << shift;
    const uint32_t tiles_x = (uint32_t)((width + tile - 1) / tile), tiles_y = (uint32_t)((height + tile - 1) / tile);
    const int whole = !opt.pfm && png_reader_interlaced(png);   /* Adam7 cannot be read a band at a time */

    pixels = malloc(whole ? row_bytes * (size_t)height : row_bytes * tile);
    const size_t obuf_size = (size_t)1 << 20;
    obuf = malloc(obuf_size);
    if (!pixels || !obuf) {
        fprintf(stderr, "%s: not enough memory\n", argv[0]);
        goto done;
    }
    if (whole && png_reader_read_image(png, pixels, row_bytes, &error_msg)) {
        fprintf(stderr, "%s: png error: %s\n", argv[0], error_msg);
        goto done;
    }
    /* (libhydrium_b200 pipelines the tiles behind this loop by itself: hyd_send_tile stages and returns,
     * finished bytes come out of the flush loops in send order, all of them by the one after the last tile) */
    enc = hyd_encoder_new();
    if (!enc) {
        fprintf(stderr, "%s: error allocating encoder\n", argv[0]);
        goto done;
    }
    if (opt.out_name && strcmp(opt.out_name, "-")) {
        fout = fopen(opt.out_name, "wb");
        if (!fout) {
            fprintf(stderr, "%s: error opening file for writing: %s\n", argv[0], opt.out_name);
            goto done;
        }
    }
    if (isatty(fileno(fout))) {
        fprintf(stderr, "%s: Not writing compressed data to a terminal.\n", argv[0]);
        usage(argv[0]);
        ret = 3;
        goto done;
    }
    ret = hyd_set_metadata(enc, &md);
    if (ret < HYD_ERROR_START)
        goto done;
    if (opt.icc_name && *opt.icc_name) {
        size_t icc_len = 0;
        icc = read_whole_file(opt.icc_name, &icc_len, argv[0]);
        if (!icc) {
            ret = 1;
            goto done;
        }
        if (icc_len) {
            ret = hyd_set_suggested_icc_profile(enc, icc, icc_len);
            if (ret < HYD_ERROR_START)
                goto done;
        }
    }
    ret = hyd_provide_output_buffer(enc, obuf, obuf_size);
    if (ret < HYD_ERROR_START)
        goto done;

    for (uint32_t k = 0; k < tiles_y; k++) {
        const uint32_t ty = opt.pfm ? tiles_y - 1 - k : k;   /* PFM stores the bottom row first */
        uint32_t rows = tile;
        if ((uint64_t)ty * tile + rows > height)
            rows = (uint32_t)(height - (uint64_t)ty * tile);
        const uint8_t *band = pixels;
        if (whole) {
            band = pixels + (size_t)ty * tile * row_bytes;
        } else if (!opt.pfm) {
            const long got = png_reader_read_rows(png, pixels, row_bytes, rows, &error_msg);
            if (got < 0) {
                fprintf(stderr, "%s: png error: %s\n", argv[0], error_msg);
                ret = 1;
                goto done;
            }
        } else {
            /* file order = bottom-up, so buffer row 0 is the band's LAST image row */
            for (uint32_t j = 0; j < rows; j++) {
                uint8_t *row = pixels + (size_t)j * row_bytes;
                if (fread(row, row_bytes, 1, fin) != 1) {
                    fprintf(stderr, "%s: incomplete pfm read\n", argv[0]);
                    ret = 1;
                    goto done;
                }
                if (swap)
                    for (size_t i = 0; i < row_bytes; i += 4) {
                        const uint8_t a = row[i], b = row[i + 1];
                        row[i] = row[i + 3];
                        row[i + 1] = row[i + 2];
                        row[i + 2] = b;
                        row[i + 3] = a;
                    }
            }
        }
        for (uint32_t tx = 0; tx < tiles_x; tx++) {
            const void *rgb[3];
            if (opt.pfm) {
                const float *p = (const float *)(const void *)(band + (size_t)(rows - 1) * row_bytes) + (size_t)tx * tile * 3;
                rgb[0] = p; rgb[1] = p + 1; rgb[2] = p + 2;
                ret = hyd_send_tile(enc, rgb, tx, ty, -(ptrdiff_t)(row_bytes / 4), 3, ty == 0 && tx == tiles_x - 1, HYD_FLOAT32);
            } else if (deep) {
                const uint16_t *p = (const uint16_t *)(const void *)band + (size_t)tx * tile * 4;
                rgb[0] = p; rgb[1] = p + 1; rgb[2] = p + 2;
                ret = hyd_send_tile(enc, rgb, tx, ty, (ptrdiff_t)(row_bytes / 2), 4, -1, HYD_UINT16);
            } else {
                const uint8_t *p = band + (size_t)tx * tile * 3;
                rgb[0] = p; rgb[1] = p + 1; rgb[2] = p + 2;
                ret = hyd_send_tile(enc, rgb, tx, ty, (ptrdiff_t)row_bytes, 3, -1, HYD_UINT8);
            }
            if (ret < HYD_ERROR_START)
                goto done;
            ret = drain_output(enc, obuf, obuf_size, fout);
            if (ret != HYD_OK)
                goto done;
        }
    }

done:
    if (fout)
        fclose(fout);
    if (fin)
        fclose(fin);
    png_reader_close(png);
    if (enc) {
        const char *m = hyd_error_message_get(enc);
        if (m)
            error_msg = m;
        hyd_encoder_destroy(enc);
    }
    free(pixels);
    free(obuf);
    free(icc);
    if (ret < HYD_ERROR_START)
        fprintf(stderr, "Hydrium error occurred. Error code: %d\n", ret);
    if (error_msg && *error_msg)
        fprintf(stderr, "Error message: %s\n", error_msg);
    return ret;
}
