/* hydrium_b200/cli/png_reader.c -- see png_reader.h */
#include "png_reader.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

enum { PNG_GREY = 0, PNG_RGB = 2, PNG_PALETTE = 3, PNG_GREY_ALPHA = 4, PNG_RGBA = 6 };

struct PngReader {
    FILE *f;
    uint32_t width, height;
    int depth, color, interlace;
    int channels;              /* samples per pixel in the file */
    uint8_t palette[256][3];
    uint32_t palette_size;
    /* IDAT stream */
    z_stream z;
    int z_open, z_done;
    uint32_t chunk_left;       /* bytes left in the current IDAT chunk */
    int no_more_idat;
    uint8_t inbuf[1 << 16];
    /* scanlines */
    uint8_t *line, *prev;      /* 1 filter byte + raw bytes of the widest row */
    uint32_t next_row;         /* non-interlaced progress */
};

static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static size_t raw_row_bytes(const PngReader *r, uint32_t pixels) {
    return ((size_t)pixels * (size_t)(r->channels * r->depth) + 7) / 8;
}

/* next chunk header; returns 0 at a clean end of file */
static int next_chunk(PngReader *r, uint32_t *len, char type[4]) {
    uint8_t h[8];
    if (fread(h, 1, 8, r->f) != 8)
        return 0;
    *len = be32(h);
    memcpy(type, h + 4, 4);
    return 1;
}

static int skip_bytes(FILE *f, uint64_t n) {
    uint8_t tmp[4096];
    while (n) {
        size_t k = n < sizeof(tmp) ? (size_t)n : sizeof(tmp);
        if (fread(tmp, 1, k, f) != k)
            return -1;
        n -= k;
    }
    return 0;
}

PngReader *png_reader_open(FILE *f, const char **error) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t head[8];
    if (fread(head, 1, 8, f) != 8 || memcmp(head, sig, 8)) {
        *error = "invalid signature";
        return NULL;
    }
    PngReader *r = calloc(1, sizeof(*r));
    if (!r) {
        *error = "couldn't allocate context";
        return NULL;
    }
    r->f = f;
    uint32_t len;
    char type[4];
    uint8_t ihdr[13];
    if (!next_chunk(r, &len, type) || memcmp(type, "IHDR", 4) || len != 13 || fread(ihdr, 1, 13, f) != 13 ||
        skip_bytes(f, 4)) {
        *error = "missing IHDR chunk";
        free(r);
        return NULL;
    }
    r->width = be32(ihdr);
    r->height = be32(ihdr + 4);
    r->depth = ihdr[8];
    r->color = ihdr[9];
    r->interlace = ihdr[12];
    int ok = r->width && r->height && r->width <= 0x7fffffffu && r->height <= 0x7fffffffu && !ihdr[10] && !ihdr[11] &&
             r->interlace <= 1;
    switch (r->color) {
    case PNG_GREY: r->channels = 1; ok = ok && (r->depth == 1 || r->depth == 2 || r->depth == 4 || r->depth == 8 || r->depth == 16); break;
    case PNG_RGB: r->channels = 3; ok = ok && (r->depth == 8 || r->depth == 16); break;
    case PNG_PALETTE: r->channels = 1; ok = ok && (r->depth == 1 || r->depth == 2 || r->depth == 4 || r->depth == 8); break;
    case PNG_GREY_ALPHA: r->channels = 2; ok = ok && (r->depth == 8 || r->depth == 16); break;
    case PNG_RGBA: r->channels = 4; ok = ok && (r->depth == 8 || r->depth == 16); break;
    default: ok = 0;
    }
    if (!ok) {
        *error = "invalid IHDR chunk";
        free(r);
        return NULL;
    }
    return r;
}

void png_reader_close(PngReader *r) {
    if (!r)
        return;
    if (r->z_open)
        inflateEnd(&r->z);
    free(r->line);
    free(r->prev);
    free(r);
}

uint32_t png_reader_width(const PngReader *r) { return r->width; }
uint32_t png_reader_height(const PngReader *r) { return r->height; }
int png_reader_bit_depth(const PngReader *r) { return r->depth; }
int png_reader_interlaced(const PngReader *r) { return r->interlace != 0; }
size_t png_reader_row_bytes(const PngReader *r) { return (size_t)r->width * (r->depth > 8 ? 8 : 3); }

/* walk the chunks in front of the image data: keep PLTE, stop inside the first IDAT */
static int start_image_data(PngReader *r, const char **error) {
    if (r->z_open)
        return 0;
    for (;;) {
        uint32_t len;
        char type[4];
        if (!next_chunk(r, &len, type)) {
            *error = "missing IDAT chunk";
            return -1;
        }
        if (!memcmp(type, "IDAT", 4)) {
            r->chunk_left = len;
            break;
        }
        if (!memcmp(type, "PLTE", 4)) {
            if (len % 3 || len > 768) {
                *error = "invalid PLTE chunk";
                return -1;
            }
            uint8_t buf[768];
            if (fread(buf, 1, len, r->f) != len || skip_bytes(r->f, 4)) {
                *error = "unexpected end of file";
                return -1;
            }
            r->palette_size = len / 3;
            memcpy(r->palette, buf, len);
            continue;
        }
        if (!memcmp(type, "IEND", 4)) {
            *error = "missing IDAT chunk";
            return -1;
        }
        if (skip_bytes(r->f, (uint64_t)len + 4)) {
            *error = "unexpected end of file";
            return -1;
        }
    }
    if (r->color == PNG_PALETTE && !r->palette_size) {
        *error = "missing PLTE chunk";
        return -1;
    }
    const size_t widest = 1 + raw_row_bytes(r, r->width);
    r->line = malloc(widest);
    r->prev = malloc(widest);
    if (!r->line || !r->prev || inflateInit(&r->z) != Z_OK) {
        *error = "out of memory";
        return -1;
    }
    r->z_open = 1;
    return 0;
}

/* inflate exactly n bytes of the IDAT stream into dst */
static int inflate_bytes(PngReader *r, uint8_t *dst, size_t n, const char **error) {
    r->z.next_out = dst;
    r->z.avail_out = (uInt)n;
    while (r->z.avail_out) {
        if (r->z_done) {
            *error = "image data ends early";
            return -1;
        }
        if (!r->z.avail_in) {
            while (!r->chunk_left && !r->no_more_idat) {   /* next chunk of the stream */
                uint32_t len;
                char type[4];
                if (skip_bytes(r->f, 4) || !next_chunk(r, &len, type) || memcmp(type, "IDAT", 4))
                    r->no_more_idat = 1;
                else
                    r->chunk_left = len;
            }
            if (!r->chunk_left) {
                *error = "image data ends early";
                return -1;
            }
            size_t k = r->chunk_left < sizeof(r->inbuf) ? r->chunk_left : sizeof(r->inbuf);
            if (fread(r->inbuf, 1, k, r->f) != k) {
                *error = "unexpected end of file";
                return -1;
            }
            r->chunk_left -= (uint32_t)k;
            r->z.next_in = r->inbuf;
            r->z.avail_in = (uInt)k;
        }
        const int zr = inflate(&r->z, Z_NO_FLUSH);
        if (zr == Z_STREAM_END)
            r->z_done = 1;
        else if (zr != Z_OK) {
            *error = "corrupt image data";
            return -1;
        }
    }
    return 0;
}

static int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* read + unfilter one scanline of `pixels` pixels into r->line (filter byte at [0]); `first` = no row above */
static int next_scanline(PngReader *r, uint32_t pixels, int first, const char **error) {
    const size_t n = raw_row_bytes(r, pixels);
    uint8_t *t = r->line;
    r->line = r->prev;
    r->prev = t;
    if (first)
        memset(r->prev, 0, n + 1);
    if (inflate_bytes(r, r->line, n + 1, error))
        return -1;
    const size_t bpp = (size_t)(r->channels * r->depth + 7) / 8;
    uint8_t *cur = r->line + 1;
    const uint8_t *up = r->prev + 1;
    switch (r->line[0]) {
    case 0:
        break;
    case 1:
        for (size_t i = bpp; i < n; i++)
            cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
        break;
    case 2:
        for (size_t i = 0; i < n; i++)
            cur[i] = (uint8_t)(cur[i] + up[i]);
        break;
    case 3:
        for (size_t i = 0; i < n; i++)
            cur[i] = (uint8_t)(cur[i] + (((i >= bpp ? cur[i - bpp] : 0) + up[i]) >> 1));
        break;
    case 4:
        for (size_t i = 0; i < n; i++)
            cur[i] = (uint8_t)(cur[i] + paeth(i >= bpp ? cur[i - bpp] : 0, up[i], i >= bpp ? up[i - bpp] : 0));
        break;
    default:
        *error = "invalid filter type";
        return -1;
    }
    return 0;
}

/* pixel i of the current scanline -> output pixel at dst (RGB8 or RGBA16) */
static int emit_pixel(const PngReader *r, uint32_t i, uint8_t *dst) {
    const uint8_t *raw = r->line + 1;
    if (r->depth == 16) {
        const uint8_t *p = raw + (size_t)i * 2 * r->channels;
        uint16_t s[4];
        for (int c = 0; c < r->channels; c++)
            s[c] = (uint16_t)((p[2 * c] << 8) | p[2 * c + 1]);
        uint16_t out[4];
        if (r->color == PNG_GREY || r->color == PNG_GREY_ALPHA) {
            out[0] = out[1] = out[2] = s[0];
            out[3] = r->color == PNG_GREY_ALPHA ? s[1] : 0xFFFFu;
        } else {
            out[0] = s[0]; out[1] = s[1]; out[2] = s[2];
            out[3] = r->color == PNG_RGBA ? s[3] : 0xFFFFu;
        }
        memcpy(dst, out, 8);
        return 0;
    }
    if (r->depth == 8 && r->color != PNG_PALETTE) {
        const uint8_t *p = raw + (size_t)i * r->channels;
        if (r->color == PNG_GREY || r->color == PNG_GREY_ALPHA)
            dst[0] = dst[1] = dst[2] = p[0];
        else
            dst[0] = p[0], dst[1] = p[1], dst[2] = p[2];
        return 0;
    }
    /* 1, 2, 4 or 8 bits, one sample per pixel, packed from the most significant bit down */
    const size_t bit = (size_t)i * r->depth;
    const uint32_t v = (raw[bit >> 3] >> (8 - r->depth - (int)(bit & 7))) & ((1u << r->depth) - 1u);
    if (r->color == PNG_PALETTE) {
        if (v >= r->palette_size)
            return -1;
        dst[0] = r->palette[v][0]; dst[1] = r->palette[v][1]; dst[2] = r->palette[v][2];
    } else {
        static const uint8_t scale[9] = {0, 255, 85, 0, 17, 0, 0, 0, 1};   /* bit replication up to 8 bits */
        dst[0] = dst[1] = dst[2] = (uint8_t)(v * scale[r->depth]);
    }
    return 0;
}

long png_reader_read_rows(PngReader *r, void *dst, size_t stride, uint32_t rows, const char **error) {
    if (r->interlace) {
        *error = "interlaced image: decode it whole";
        return -1;
    }
    if (start_image_data(r, error))
        return -1;
    const size_t px = r->depth > 8 ? 8 : 3;
    long done = 0;
    while (rows-- && r->next_row < r->height) {
        if (next_scanline(r, r->width, r->next_row == 0, error))
            return -1;
        uint8_t *out = (uint8_t *)dst + (size_t)done * stride;
        for (uint32_t x = 0; x < r->width; x++)
            if (emit_pixel(r, x, out + (size_t)x * px)) {
                *error = "palette index out of range";
                return -1;
            }
        r->next_row++;
        done++;
    }
    return done;
}

int png_reader_read_image(PngReader *r, void *dst, size_t stride, const char **error) {
    if (!r->interlace) {
        const long n = png_reader_read_rows(r, dst, stride, r->height, error);
        if (n < 0)
            return -1;
        if ((uint32_t)n != r->height) {
            *error = "image data ends early";
            return -1;
        }
        return 0;
    }
    if (start_image_data(r, error))
        return -1;
    static const uint8_t x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const uint8_t dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    const size_t px = r->depth > 8 ? 8 : 3;
    for (int pass = 0; pass < 7; pass++) {
        if (r->width <= x0[pass] || r->height <= y0[pass])
            continue;
        const uint32_t pw = (r->width - x0[pass] + dx[pass] - 1) / dx[pass];
        const uint32_t ph = (r->height - y0[pass] + dy[pass] - 1) / dy[pass];
        for (uint32_t j = 0; j < ph; j++) {
            if (next_scanline(r, pw, j == 0, error))
                return -1;
            uint8_t *row = (uint8_t *)dst + (size_t)(y0[pass] + j * dy[pass]) * stride;
            for (uint32_t i = 0; i < pw; i++)
                if (emit_pixel(r, i, row + (size_t)(x0[pass] + i * dx[pass]) * px)) {
                    *error = "palette index out of range";
                    return -1;
                }
        }
    }
    return 0;
}
