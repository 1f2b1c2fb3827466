/*
 * include/hydrium_b200.h -- C ABI of libhydrium_b200.so
 *
 * Part 1 is the drop-in boundary: the nine entry points, enums and the metadata struct of
 * libhydrium 0.6.0, with identical names, signatures, values and struct layout, so a program
 * compiled against the reference header links and runs against this library unchanged
 * (reference: src/include/libhydrium/libhydrium.h; each declaration cites the lines it replaces).
 * `include/libhydrium/libhydrium.h` in this tree just includes this file.
 *
 * Part 2 is additive (hydb_*): device selection, batching control, and entry points that take
 * images already resident in GPU memory.  Nothing in part 2 changes the behaviour of part 1.
 *
 * The encoder behind both parts is CUDA-only (sm_100a).  There is no CPU implementation: without
 * a usable device hyd_set_metadata / hydb_engine_create fail with HYD_INTERNAL_ERROR.
 */
#ifndef HYDRIUM_B200_H_
#define HYDRIUM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* version of the API this library is a drop-in for (libhydrium.h:17-51) */
#define HYDRIUM_VERSION_MAJOR 0
#define HYDRIUM_VERSION_MINOR 6
#define HYDRIUM_VERSION_POINT 0
#define HYDRIUM_COMPUTE_VERSION(ma, mi, po) \
    (UINT64_C(0x1000000000) | ((uint64_t)(ma) << 24) | ((uint64_t)(mi) << 12) | ((uint64_t)(po)))
#define HYDRIUM_VERSION_INT \
    HYDRIUM_COMPUTE_VERSION(HYDRIUM_VERSION_MAJOR, HYDRIUM_VERSION_MINOR, HYDRIUM_VERSION_POINT)
#define HYDRIUM_VERSION_STRING "0.6.0"

#if defined(__GNUC__) || defined(__clang__)
#define HYDRIUM_EXPORT __attribute__((visibility("default")))
#else
#define HYDRIUM_EXPORT
#endif

/* ===================================================================== part 1: libhydrium ABI */

/* libhydrium.h:67-101.  Errors are < HYD_ERROR_START. */
typedef enum HYDStatusCode {
    HYD_OK = 0,
    HYD_DEFAULT = -1,
    HYD_NEED_MORE_OUTPUT = -2,
    HYD_NEED_MORE_INPUT = -3,
    HYD_ERROR_START = -10,
    HYD_NOMEM = -13,
    HYD_API_ERROR = -14,
    HYD_INTERNAL_ERROR = -15
} HYDStatusCode;

/* libhydrium.h:103-107 */
typedef enum HYDSampleFormat {
    HYD_UINT8 = 0,
    HYD_UINT16 = 1,
    HYD_FLOAT32 = 2
} HYDSampleFormat;

/* libhydrium.h:109-155.  tile_size_shift_{x,y}: 0..3 => 256<<shift pixel tiles, -1 => one frame.
 * Every combination is encoded; the one limit is one-frame mode beyond 256 LF groups of 2048x2048
 * (more than 1 Gpx), which hyd_set_metadata refuses with HYD_API_ERROR and a message. */
typedef struct HYDImageMetadata {
    size_t width;
    size_t height;
    int linear_light;
    int tile_size_shift_x;
    int tile_size_shift_y;
} HYDImageMetadata;

typedef struct HYDEncoder HYDEncoder;

/* libhydrium.h:165 */
HYDRIUM_EXPORT HYDEncoder *hyd_encoder_new(void);
/* libhydrium.h:173 */
HYDRIUM_EXPORT HYDStatusCode hyd_encoder_destroy(HYDEncoder *encoder);
/* libhydrium.h:183 */
HYDRIUM_EXPORT HYDStatusCode hyd_set_metadata(HYDEncoder *encoder, const HYDImageMetadata *metadata);
/* libhydrium.h:193 */
HYDRIUM_EXPORT HYDStatusCode hyd_provide_output_buffer(HYDEncoder *encoder, uint8_t *buffer, size_t buffer_len);
/* libhydrium.h:260-262.  Strides are in samples; the three planes may alias (packed RGB/RGBA).
 * The sample buffers are only read during the call. */
HYDRIUM_EXPORT HYDStatusCode hyd_send_tile(HYDEncoder *encoder, const void *const buffer[3],
                                           uint32_t tile_x, uint32_t tile_y, ptrdiff_t row_stride,
                                           ptrdiff_t pixel_stride, int is_last, HYDSampleFormat sample_fmt);
/* libhydrium.h:275 */
HYDRIUM_EXPORT HYDStatusCode hyd_release_output_buffer(HYDEncoder *encoder, size_t *written);
/* libhydrium.h:288 */
HYDRIUM_EXPORT HYDStatusCode hyd_flush(HYDEncoder *encoder);
/* libhydrium.h:295 */
HYDRIUM_EXPORT const char *hyd_error_message_get(HYDEncoder *encoder);
/* libhydrium.h:296-314.  One-frame mode only, like the reference ("one-frame mode required to set the
 * suggested ICC profile" otherwise); NULL/0 clears.  The profile is copied; it is entropy coded into
 * the image header when the first tile is sent. */
HYDRIUM_EXPORT HYDStatusCode hyd_set_suggested_icc_profile(HYDEncoder *encoder, const uint8_t *icc_data,
                                                           size_t icc_size);

/* ===================================================================== part 2: additive API */

/* How hyd_send_tile feeds the GPU.  By default (no call, HYDRIUM_B200_BATCH unset) it is asynchronous: the
 * tile is copied into page-locked staging memory and the call returns; tiles are encoded a chunk (32 tiles,
 * or one multi-group frame) at a time by engine jobs that overlap the staging of later tiles, several
 * chunks in flight (HYDRIUM_B200_DEPTH, default 8 / 4); hyd_flush returns bytes in send order as chunks
 * finish, everything by the end of the flush loop after the last tile (or on a second consecutive
 * hyd_flush), and the concatenation of everything surfaced is byte-identical to the reference's.
 * 1 = every hyd_send_tile encodes synchronously and the following hyd_flush returns that tile's bytes,
 * exactly like the reference.  N > 1 = asynchronous with chunks of N tiles.
 * Must be called before the first hyd_send_tile. */
HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_batch(HYDEncoder *encoder, uint32_t tiles);
/* counters of the engine behind this encoder (0 before the first tile): kernels launched, jobs replayed as
 * CUDA graphs -- what bench.py's gpu_launches and the graph-replay test read */
HYDRIUM_EXPORT void hydb_encoder_stats(const HYDEncoder *encoder, uint64_t *kernel_launches, uint64_t *graph_launches);
/* CUDA device ordinal for this encoder.  Default: env HYDRIUM_B200_DEVICE, else the current device. */
HYDRIUM_EXPORT HYDStatusCode hydb_encoder_set_device(HYDEncoder *encoder, int device);

/* ---- engine: batch encoder over device-resident images ------------------------------------ */
typedef struct HydbEngine HydbEngine;

/* One tile to encode; pointers are DEVICE pointers (same meaning as hyd_send_tile's arguments). */
typedef struct HydbTile {
    const void *plane[3];
    int64_t row_stride;      /* samples */
    int64_t pixel_stride;    /* samples */
    uint32_t width, height;  /* tile size in pixels, <= 256 */
    uint32_t x0, y0;         /* origin in the image, multiples of 256 */
    uint32_t image_width, image_height;
    int32_t is_last;         /* 0 / 1 */
    int32_t sample_fmt;      /* HYD_UINT8 / HYD_UINT16 */
    int32_t linear_light;
    int32_t with_image_header; /* 1: this tile starts a codestream, emit the image header before its frame */
} HydbTile;

/* max_batch_tiles bounds the tiles per hydb_engine_encode_tiles call (workspace ~2.4 MB / tile). */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_create(HydbEngine **engine, int device, uint32_t max_batch_tiles);
HYDRIUM_EXPORT void hydb_engine_destroy(HydbEngine *engine);
HYDRIUM_EXPORT const char *hydb_engine_error(const HydbEngine *engine);
HYDRIUM_EXPORT uint32_t hydb_engine_max_batch(const HydbEngine *engine);
/* the cudaStream_t the engine launches on (as an integer handle), for CUDA-event timing */
HYDRIUM_EXPORT uint64_t hydb_engine_stream(const HydbEngine *engine);
/* number of kernel launches issued by the engine so far (kernels replayed from a CUDA graph count) */
HYDRIUM_EXPORT uint64_t hydb_engine_launch_count(const HydbEngine *engine);
/* how many asynchronous jobs were submitted as one cudaGraphLaunch (recurring chunk geometries) */
HYDRIUM_EXPORT uint64_t hydb_engine_graph_launch_count(const HydbEngine *engine);

/* Which rANS chain kernel the engine launches: 0 (default) = by launch size -- the table kernel (72 KB
 * inverse alias table per tile, two chains per SM, shortest step) up to 2 x SM-count tiles per launch,
 * the compact kernel (sorted alias pieces, sixteen chains per SM) beyond; 1 = always the table kernel;
 * 2 = the compact kernel wherever it applies (integer sample formats).  Same bytes either way. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_set_chain_kernel(HydbEngine *engine, int mode);

/* Encode n tiles (n <= max batch) and append their frames, in order, to d_out (device memory)
 * starting at byte d_out_pos.  Asynchronous on the engine stream; call hydb_engine_finish to
 * synchronise, collect per-tile errors and learn how many bytes the batch appended. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_encode_tiles(HydbEngine *engine, const HydbTile *tiles, uint32_t n,
                                                      uint8_t *d_out, uint64_t d_out_cap, uint64_t d_out_pos);
/* One FRAME of up to 2048x2048 pixels = up to 8x8 groups of 256x256 (what hyd_send_tile receives with
 * tile_size_shift 1..3, or in one-frame mode for an image of at most one LF group; reference:
 * libhydrium.h:198-206, encoder.c:437-472).  Pointers are DEVICE pointers.  A frame takes
 * 1 + groups workspace slots of the engine's batch (1 slot when it is a single group). */
typedef struct HydbFrame {
    const void *plane[3];
    int64_t row_stride;      /* samples */
    int64_t pixel_stride;    /* samples */
    uint32_t width, height;  /* frame size in pixels, <= 2048 */
    uint32_t x0, y0;         /* origin in the image, multiples of the tile size */
    uint32_t image_width, image_height;
    int32_t is_last;         /* 0 / 1 */
    int32_t sample_fmt;
    int32_t linear_light;
    int32_t with_image_header;
    int32_t one_frame;       /* 1: one-frame mode header flavour (no crop, always last) */
    /* one-frame mode over SEVERAL LF groups: each 2048x2048 LF group is encoded as a frame part
     * (lf_part = 1).  The part's output is  LFGroup section | PassGroup sections  without any frame
     * header; hydb_engine_frame_lengths gives the section lengths, hydb_engine_read_model the part's
     * ANS histograms, and hydb_oneframe_finish assembles the frame's head once all parts exist. */
    int32_t lf_part;
    uint32_t preset;         /* HF preset of this LF group (its raster index while the image has <= 256 of them) */
    uint32_t preset_bits;    /* ceil(log2(number of presets)) */
    uint32_t alpha_floor;    /* largest token alphabet of the parts sent before (the reference's running maximum) */
    uint32_t clusters_per_preset; /* 9, or 3 / 2 / 1 when the image has more than 28 / 85 / 128 LF groups; 0 = 9 */
} HydbFrame;
/* ANS model of the slot's group after hydb_engine_finish: the nine cluster histograms as a bit string
 * (bits_out: >= 384 words; *nbits), and the largest token alphabet seen in that frame part. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_read_model(HydbEngine *engine, uint32_t slot, uint32_t *bits_out,
                                                    uint32_t *nbits, uint32_t *max_alphabet);
/* Head of a one-frame image of several LF groups, from what the parts produced (host buffers in and out):
 * [image header] + frame header with the TOC permutation + TOC + LFGlobal section  -> head
 * HFGlobal section (presets, context map, ANS header with every preset's histograms) -> hf_global
 * info words: see hydb_oneframe_finish in csrc/engine.cu. */
HYDRIUM_EXPORT HYDStatusCode hydb_oneframe_finish(HydbEngine *engine, const uint32_t *info, uint32_t info_words,
                                                  uint8_t *head, uint32_t head_cap, uint32_t *head_len,
                                                  uint8_t *hf_global, uint32_t hf_cap, uint32_t *hf_len);
/* Encode n frames and append them, in order, to d_out at byte d_out_pos (asynchronous like
 * hydb_engine_encode_tiles; finish with hydb_engine_finish). */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_encode_frames(HydbEngine *engine, const HydbFrame *frames, uint32_t n,
                                                       uint8_t *d_out, uint64_t d_out_cap, uint64_t d_out_pos);
/* ---- asynchronous jobs: what the nine-symbol API runs on -------------------------------------------------
 * hydb_engine_submit_frames enqueues, on a stream pair of its own,
 *     [copy of h2d_bytes from h_src (page-locked) to d_dst] -> the kernels for n frames on workspace slots
 *     slot0 .. -> the compaction of their bytes into `out` -> a 16-byte result record
 * and returns at once with a job id (at most 16 jobs at a time).  `out` may be DEVICE memory or PAGE-LOCKED
 * HOST memory (hydb_host_alloc): in the second case the compaction kernel writes the codestream across PCIe
 * itself and no copy follows.  The frames' plane pointers are device pointers (normally into d_dst).  The
 * caller owns the slot ranges: concurrent jobs must not overlap, and a job's slots stay untouched until it
 * is released.  *slots = workspace slots the job took (1 per single-group frame, 1 + groups otherwise).
 * hydb_engine_job_poll: HYD_DEFAULT while running (wait = 0), HYD_OK with *bytes, HYD_NEED_MORE_OUTPUT when
 * `out` was too small (hydb_engine_job_regather gathers again, synchronously, into a device buffer), or the
 * error of the job's tiles.  hydb_engine_job_release frees the job id. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_submit_frames(HydbEngine *engine, const HydbFrame *frames, uint32_t n,
                                                       uint32_t slot0, const void *h_src, void *d_dst, size_t h2d_bytes,
                                                       uint8_t *out, uint64_t out_cap, uint32_t *job, uint32_t *slots);
HYDRIUM_EXPORT HYDStatusCode hydb_engine_job_poll(HydbEngine *engine, uint32_t job, int wait, uint64_t *bytes);
HYDRIUM_EXPORT HYDStatusCode hydb_engine_job_regather(HydbEngine *engine, uint32_t job, uint8_t *d_out, uint64_t d_out_cap,
                                                      uint64_t *bytes);
HYDRIUM_EXPORT HYDStatusCode hydb_engine_job_release(HydbEngine *engine, uint32_t job);
/* frame byte lengths of n slots starting at slot0 (of a finished job) */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_slot_frame_lengths(HydbEngine *engine, uint32_t slot0, uint32_t *dst, uint32_t n);

/* Synchronise, check per-tile error flags of the last batch, return total bytes appended by it. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_finish(HydbEngine *engine, uint64_t *batch_bytes);

/* Byte length of each frame of the last batch, in tile order (to split one batch into several
 * codestreams, e.g. one per image of a batch of frames). */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_frame_lengths(HydbEngine *engine, uint32_t *dst, uint32_t n);

/* Whole image, tile mode shift 0/0, raster tile order restricted to tile rows
 * [tile_row_begin, tile_row_end): d_pixels points at the first sample of tile row tile_row_begin
 * (interleaved, `channels` samples per pixel, row_stride samples per row).  Writes
 * [image header if with_header] + frames to d_out; *out_len = bytes.  Synchronous. */
HYDRIUM_EXPORT HYDStatusCode hydb_encode_image_device(HydbEngine *engine, const void *d_pixels, uint32_t width,
                                                      uint32_t height, uint32_t channels, int64_t row_stride,
                                                      int sample_fmt, int linear_light, uint32_t tile_row_begin,
                                                      uint32_t tile_row_end, int with_header, uint8_t *d_out,
                                                      uint64_t d_out_cap, uint64_t *out_len);
/* Same with HOST buffers: pixels are copied to the device, the codestream is copied back.
 * h_pixels / h_out should be page-locked for full PCIe speed (hydb_host_alloc). */
HYDRIUM_EXPORT HYDStatusCode hydb_encode_image_host(HydbEngine *engine, const void *h_pixels, uint32_t width,
                                                    uint32_t height, uint32_t channels, int sample_fmt,
                                                    int linear_light, uint8_t *h_out, uint64_t h_out_cap,
                                                    uint64_t *out_len);

/* Per-kernel device time, measured with CUDA events on the launching streams.  stage_ms returns and
 * resets the milliseconds accumulated since the last call:
 * [0] xyb_dct_quant [1] hf_tokens [2] ans_chain [3] ans_pack (incl. waiting for lf_group) [4] offsets+gather
 * [5] lf_group (second stream, concurrent with [1]-[2]) [6] number of batches */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_enable_timing(HydbEngine *engine, int enable);
HYDRIUM_EXPORT HYDStatusCode hydb_engine_stage_ms(HydbEngine *engine, double out[7]);

/* image header bytes (with the level-10 container prefix where the reference emits it) */
HYDRIUM_EXPORT int64_t hydb_image_header(uint32_t width, uint32_t height, uint8_t *dst, uint64_t cap);

/* Image header of a codestream tagged with an ICC profile (reference: encoder.c:203-236, the 41-context
 * prefix-coded profile stream): `mangled` is the profile as hyd_set_suggested_icc_profile prepares it
 * (libhydrium.c:242-305).  Runs k_icc_header on the engine's device; host buffers; synchronous. */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_icc_header(HydbEngine *engine, uint32_t width, uint32_t height,
                                                    const uint8_t *mangled, uint32_t mangled_size, uint8_t *dst,
                                                    uint64_t cap, uint64_t *len);

/* page-locked host memory / device memory helpers for callers without a CUDA runtime of their own */
HYDRIUM_EXPORT void *hydb_host_alloc(size_t bytes);
HYDRIUM_EXPORT void hydb_host_free(void *p);
HYDRIUM_EXPORT void *hydb_device_alloc(size_t bytes);
HYDRIUM_EXPORT void hydb_device_free(void *p);
HYDRIUM_EXPORT int hydb_memcpy_h2d(void *dst, const void *src, size_t bytes);
HYDRIUM_EXPORT int hydb_memcpy_d2h(void *dst, const void *src, size_t bytes);
HYDRIUM_EXPORT int hydb_device_count(void);

/* Multi-GPU gather over peer memory (one process per GPU): the gathering rank allocates a
 * buffer of `regions` x region_stride bytes (hydb_device_alloc) and exports it; every rank opens it and
 * passes  base + rank * region_stride + 256  as d_out to its encode call, so its frames are written
 * straight into the gathering rank's HBM over NVLink by the compaction kernel itself, then stores the
 * span length (uint64) at  base + rank * region_stride  with hydb_engine_store_u64 (stream ordered).  After a barrier the gathering rank closes the
 * gaps with hydb_engine_compact_regions: spans in rank order, contiguous, in d_out; *total = bytes. */
HYDRIUM_EXPORT int hydb_ipc_export(const void *d_ptr, uint8_t handle[64]);
HYDRIUM_EXPORT void *hydb_ipc_open(const uint8_t handle[64]);
HYDRIUM_EXPORT void hydb_ipc_close(void *p);
/* one 64-bit word written on the engine's stream (the span length in a region header) */
HYDRIUM_EXPORT HYDStatusCode hydb_engine_store_u64(HydbEngine *engine, void *d_dst, uint64_t value);
HYDRIUM_EXPORT HYDStatusCode hydb_engine_compact_regions(HydbEngine *engine, const uint8_t *d_regions, uint32_t regions,
                                                         uint64_t region_stride, uint8_t *d_out, uint64_t d_out_cap,
                                                         uint64_t *total);

/* closed-form synthetic RGB image generated directly in device memory (benchmarks) */
HYDRIUM_EXPORT int hydb_synth_fill(HydbEngine *engine, void *d_dst, uint32_t width, uint32_t height, uint32_t x0,
                                   uint32_t y0, uint32_t full_width, uint32_t full_height, int bits, uint32_t seed,
                                   int smooth);

/* stage taps for the parity tests: enable before encoding, then copy a stage of one tile of the
 * last batch to host memory.  Returns bytes copied or a negative status. */
enum {
    HYDB_TAP_XYB = 0,      /* float [256][256][3]                  */
    HYDB_TAP_DCT = 1,      /* float [256][256][3]                  */
    HYDB_TAP_COEF = 2,     /* int16 [1024][3][64] scan order       */
    HYDB_TAP_NZINFO = 3,   /* uint16 [1024][3]                     */
    HYDB_TAP_LFQ = 4,      /* int32 [3][1024]                      */
    HYDB_TAP_SYMS = 5,     /* uint32 [nsyms] packed records        */
    HYDB_TAP_FREQS = 6,    /* uint32 [9][64] normalised            */
    HYDB_TAP_LFBITS = 7,   /* uint32 words, section L              */
    HYDB_TAP_SECT = 8,     /* uint32 [4] section bit lengths       */
    HYDB_TAP_PAYLOAD = 9,  /* bytes: payload of the frame          */
    HYDB_TAP_NSYMS = 10,   /* uint32 [1]                           */
    HYDB_TAP_LFBITLEN = 11,/* uint32 [1]                           */
    HYDB_TAP_CLK = 12      /* uint32 [4] chain kernel: prologue cycles, chain cycles, SM id, chain warp */
};
HYDRIUM_EXPORT HYDStatusCode hydb_engine_enable_taps(HydbEngine *engine, int enable);
HYDRIUM_EXPORT int64_t hydb_engine_read_tap(HydbEngine *engine, int what, uint32_t tile, void *dst, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* HYDRIUM_B200_H_ */
