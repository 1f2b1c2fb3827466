/* hydrium_b200/cli/png_reader.h
 *
 * Minimal streaming PNG reader for the command-line front end (zlib for inflate, nothing else).
 * It produces what the reference CLI asks libspng for (hydrium.c:255-262, 305-313, 404-417):
 *   bit depth <= 8  ->  RGB8   (3 bytes per pixel; grey and palette expanded, low depths scaled to 8 bits,
 *                               alpha dropped)
 *   bit depth 16    ->  RGBA16 (4 host-endian uint16 per pixel; alpha 65535 when the file has none)
 * with no gamma and no tRNS processing, and it does not reject a chunk over its CRC (SPNG_CRC_USE).
 * Non-interlaced files are read a band of rows at a time; Adam7 files are decoded whole.
 */
#ifndef HYDRIUM_B200_PNG_READER_H
#define HYDRIUM_B200_PNG_READER_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

typedef struct PngReader PngReader;

/* Reads up to and including IHDR.  Returns NULL and sets *error on failure. */
PngReader *png_reader_open(FILE *f, const char **error);
void png_reader_close(PngReader *r);

uint32_t png_reader_width(const PngReader *r);
uint32_t png_reader_height(const PngReader *r);
int png_reader_bit_depth(const PngReader *r);
int png_reader_interlaced(const PngReader *r);
/* bytes of one output row: 3 * width (RGB8) or 8 * width (RGBA16) */
size_t png_reader_row_bytes(const PngReader *r);

/* Non-interlaced: decode the next `rows` rows (fewer at the end of the image) into dst, `stride`
 * bytes apart.  Returns the number of rows written, or -1 with *error set. */
long png_reader_read_rows(PngReader *r, void *dst, size_t stride, uint32_t rows, const char **error);
/* Any file: decode the whole image.  Returns 0, or -1 with *error set. */
int png_reader_read_image(PngReader *r, void *dst, size_t stride, const char **error);

#endif
