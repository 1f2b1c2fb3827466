/*
 * oracle/hyd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the hydrium tile-encode path, used
 * solely as the parity checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte
 * against the unmodified reference compiled from /root/reference (oracle/_ref, see
 * Makefile) over the known-answer set of SURVEY.md Appendix C and random tiles, and
 * against the committed fixtures in tests/golden/.
 *
 * Scope: tile mode with one 256x256 group per frame (tile_size_shift 0/0), and the
 * degenerate one-frame case of an image that fits one group; HYD_UINT8 / HYD_UINT16 input.
 */
#ifndef HYD_ORACLE_H_
#define HYD_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_NOMEM = -13, ORC_API_ERROR = -14, ORC_INTERNAL_ERROR = -15 };
enum { ORC_UINT8 = 0, ORC_UINT16 = 1, ORC_FLOAT32 = 2 };

/* One hybrid-uint coded symbol (reference: entropy.h:9-14). */
typedef struct OrcSymbol {
    uint32_t token;
    uint32_t cluster;
    uint32_t nbits;
    uint32_t residue;
} OrcSymbol;

/* Stage outputs of one tile; every pointer is optional (NULL = not wanted) and caller-owned. */
typedef struct OrcStages {
    float    *xyb;          /* [vbh*8][vbw*8][3] after colour transform + zero padding  */
    float    *dct;          /* same shape, after the forward DCT                        */
    int32_t  *quant;        /* same shape: HF ints, and the LF ints at block origins    */
    uint8_t  *nonzeroes;    /* [vbh*vbw][3] non-zero HF count per block and channel     */
    OrcSymbol *hf_syms;     /* HF symbol stream                                         */
    uint64_t  hf_syms_cap, hf_syms_n;
    uint8_t  *lf_bits;      /* LFGlobal + LFGroup bit string                            */
    uint64_t  lf_bits_cap, lf_bitlen;
    uint32_t *freqs;        /* [9][256] normalised ANS frequencies                      */
    uint16_t  alphabet_sizes[16];
    uint32_t  max_alphabet_size;
    uint8_t  *ans_bits;     /* PassGroup bit string                                     */
    uint64_t  ans_bits_cap, ans_bitlen;
    uint8_t  *pre_bits;     /* LFGlobal + LFGroup + HFGlobal + ANS stream header        */
    uint64_t  pre_bits_cap, pre_bitlen;
    uint32_t  vbw, vbh;
} OrcStages;

/* Image geometry + the tile being coded. */
typedef struct OrcTile {
    uint64_t image_width, image_height;
    int      linear_light;
    uint32_t tile_x, tile_y;      /* in units of 256 px                                   */
    int      is_last;             /* <0: lower-right tile is last; else explicit           */
    int      sample_fmt;          /* ORC_UINT8 / ORC_UINT16 / ORC_FLOAT32                 */
    const void *plane[3];         /* first R, G, B sample of the tile                      */
    ptrdiff_t row_stride, pixel_stride; /* in samples                                     */
} OrcTile;

/* Image header bytes (signature, size, metadata; with the level-10 container prefix when
 * the reference would emit it).  Returns the byte count, or <0. */
int64_t orc_image_header(uint64_t width, uint64_t height, uint8_t *dst, uint64_t cap);

/* One complete frame (frame header, TOC, payload) for a tile.  Returns the byte count, or
 * a negative ORC_* code.  `stages` may be NULL. */
int64_t orc_encode_tile(const OrcTile *tile, uint8_t *dst, uint64_t cap, OrcStages *stages);

/* Whole image, tile mode shift 0/0, tiles in raster order, header first: the exact byte
 * stream the reference produces through the CLI call sequence.  Returns bytes or <0. */
int64_t orc_encode_image(const void *pixels, uint64_t width, uint64_t height, int channels,
                         int sample_fmt, int linear_light, uint8_t *dst, uint64_t cap);

/* The runtime lookup tables (format.c:58-83).  input_lut has 256 or 65536 entries. */
void orc_build_luts(int sample_fmt, int linear_light, uint16_t *input_lut, float *bias_lut);

/* Generic prefix-coded stream (entropy.c:807-1034) for unit tests of the entropy layer:
 * values[i] is sent on context ctx[i] (ctx may be NULL = all zero). */
int64_t orc_prefix_stream(const uint32_t *values, const uint32_t *ctx, uint64_t n,
                          const uint8_t *cluster_map, uint32_t num_dists, int custom_config,
                          int split, int msb, int lsb, uint32_t lz77_min_symbol, int modular,
                          uint8_t *dst, uint64_t cap, uint64_t *bitlen);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* HYD_ORACLE_H_ */
