// hydrium_b200/csrc/common.cuh
//
// Shared constants, the device-side tile descriptor and small helpers.
// Everything marked HD compiles for the device and (for the CPU unit tests of the
// sequential entropy logic, tests/host_harness) for the host with a plain C++ compiler.
#pragma once

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

namespace hydb {

// ---- geometry ---------------------------------------------------------------------------
constexpr int kTileDim = 256;            // one JPEG XL group (reference: encoder.c:441-446)
constexpr int kBlocksPerRow = 32;        // 8x8 varblocks per group row
constexpr int kMaxBlocks = 1024;
constexpr int kMaxHfSyms = 3 * kMaxBlocks * 64;   // 196608, reference: encoder.c:689-750

// ---- entropy model ----------------------------------------------------------------------
constexpr int kHfClusters = 9;           // reference: encoder.c:862-877 (one preset)
constexpr int kHfTokens = 64;            // ANS alphabet we support (log_alphabet_size 5 or 6)
constexpr int kAnsTotal = 4096;          // 12-bit ANS precision, reference: entropy.c:278

// ---- per-tile workspace sizes (bytes / elements) ----------------------------------------
constexpr int kLfBitsWords = 4096;       // LF stream scratch (u32 words) per tile
constexpr int kSlabBytes = 768 * 1024;   // worst-case frame: header + TOC + payload
constexpr int kSlabHeaderReserve = 128;  // [image header] + frame header + TOC are right-justified before this
constexpr int kDBitsWords = 384;         // section D (ANS header tail) scratch per tile
constexpr int kTemplWords = 256;         // per-shape constant bit strings (u32 words each)

// error bits reported per tile (engine maps them to HYD_INTERNAL_ERROR + message)
enum TileError : uint32_t {
    kErrNone = 0,
    kErrAlphabet = 1u << 0,       // HF token >= 64 (cannot happen for integer sample formats)
    kErrHuffman = 1u << 1,        // reference would fail: "couldn't find target" / "VLC codes do not add up"
    kErrLfCapacity = 1u << 2,     // LF stream scratch exhausted
    kErrAnsGap = 1u << 3,         // >= 65536 symbols without a renormalisation (reference wraps a uint16)
    kErrSlab = 1u << 4,           // frame larger than the slab
    kErrAlias = 1u << 5,          // reference would fail: "empty underfull during alias table gen"
    kErrLfAlphabet = 1u << 6,     // more distinct LF tokens than the sparse coder holds
    kErrNonFinite = 1u << 7,      // NaN / Inf float sample (reference: format.c:123-126 "Invalid NaN Float")
    kErrRange = 1u << 8,          // quantised HF coefficient outside int16 (float samples far outside [0, 1])
    kErrNegative = 1u << 9,       // float samples whose opsin mix is negative (the reference's result is undefined there)
};

// Device-visible tile descriptor (mirrors what hyd_send_tile is given,
// reference: libhydrium.h:260-262 and internal.h:13-19).
struct TileDesc {
    const void *plane[3];     // first R, G, B sample of the tile (device pointers)
    int64_t row_stride;       // in samples
    int64_t pixel_stride;     // in samples
    uint32_t w, h;            // tile size in pixels (<= 256)
    uint32_t x0, y0;          // pixel origin inside the image (frame crop)
    uint32_t flags;           // kTile* below
    uint32_t shape;           // index of the (vbw, vbh) template set
    uint32_t image_w, image_h;   // only read when kTileFirst is set
    // frames with more than one 256x256 group (tile_size_shift > 0, one-frame mode): the frame takes
    // 1 + G consecutive workspace slots, a PREFIX pseudo-tile (frame header, TOC, LFGlobal, LFGroup,
    // HFGlobal) followed by its G groups in raster order; zero for classic one-group frames
    uint32_t frame_groups;    // G
    uint32_t frame_gx;        // groups per frame row
    uint32_t group_index;     // raster index of this group inside the frame
    uint32_t frame_w, frame_h;   // frame size in pixels
    uint32_t frame_x0, frame_y0; // frame origin inside the image (crop offset)
    // one-frame mode over several LF groups (each LF group is encoded as a "frame part"):
    // bits 0-7 HF preset of the LF group, 8-11 bits the preset id is written in, 12-19 largest token
    // alphabet of the LF groups sent before (the reference's running max_alphabet_size, entropy.c:459)
    uint32_t preset_info;
};
HD uint32_t tile_preset(const TileDesc &t) { return t.preset_info & 0xFFu; }
HD uint32_t tile_preset_bits(const TileDesc &t) { return (t.preset_info >> 8) & 0xFu; }
HD uint32_t tile_alpha_floor(const TileDesc &t) { return (t.preset_info >> 12) & 0xFFu; }
// bits 20-23: HF clusters per preset.  Nine normally; a one-frame image with more than 28 LF groups
// folds them so that presets x clusters stays within 256 (encoder.c:862-899): 3 = one for the
// non-zero counts + two for the coefficients by parity, 2 = counts / coefficients, 1 = everything.
HD uint32_t tile_clusters(const TileDesc &t) { const uint32_t k = (t.preset_info >> 20) & 0xFu; return k ? k : 9u; }
HD uint32_t hf_fold_cluster(uint32_t c9, uint32_t clusters_per_preset) {
    if (clusters_per_preset >= 9)
        return c9;
    if (clusters_per_preset == 3)
        return c9 < 3 ? 0u : 1u + ((c9 - 3u) & 1u);
    if (clusters_per_preset == 2)
        return c9 < 3 ? 0u : 1u;
    return 0u;
}
HD uint32_t hf_clusters_for_presets(uint32_t presets) {   // encoder.c:862, 878, 892, 897
    return presets * 9 <= 256 ? 9u : (presets * 3 <= 256 ? 3u : (presets * 2 <= 256 ? 2u : 1u));
}
enum : uint32_t {
    kTileLast = 1u << 0,      // is_last frame (reference: encoder.c:482-485)
    kTileCrop = 1u << 1,      // image larger than the tile (reference: encoder.c:340-342)
    kTileFmt16 = 1u << 2,     // HYD_UINT16 samples, else HYD_UINT8 (or float, see kTileFmtF32)
    kTileLinear = 1u << 3,    // linear-light input
    kTileFirst = 1u << 4,     // first frame of a codestream: the image header goes in front of it
    kTileFmtF32 = 1u << 5,    // HYD_FLOAT32 samples (reference: format.c:111-140)
    kTilePrefix = 1u << 6,    // pseudo-tile holding the shared sections of a multi-group frame
    kTileMulti = 1u << 7,     // group of a multi-group frame: its slab carries only the PassGroup section
    kTileOneFrame = 1u << 8,  // one-frame mode header flavour: no crop, always last (encoder.c:339-342)
    kTileLfPart = 1u << 9,    // prefix of one LF group of a larger one-frame image: LFGroup section only
};

// ---- integer helpers (reference: math-functions.h:8-88) ----------------------------------
HD int floor_log2_u32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
HD int ceil_log2_u32(uint32_t v) { return floor_log2_u32(v) + ((v & (v - 1)) != 0); }
HD uint32_t pack_signed(int32_t v) { uint32_t w = (uint32_t)v; return (w << 1) ^ (0u - (w >> 31)); }

// hybrid-uint split (reference: entropy.c:427-444).  Returns token; residue/nbits by reference.
HD uint32_t hybrid_token(uint32_t v, int split, int msb, int lsb, uint32_t &residue, uint32_t &nbits) {
    if (v < (1u << split)) {
        residue = 0;
        nbits = 0;
        return v;
    }
    const uint32_t n = (uint32_t)floor_log2_u32(v) - (uint32_t)lsb - (uint32_t)msb;
    const uint32_t low = v & ((1u << lsb) - 1u);
    v >>= lsb;
    residue = v & ((1u << n) - 1u);
    v >>= n;
    const uint32_t high = v & ((1u << msb) - 1u);
    nbits = n;
    return (1u << split) + (low | (high << lsb) | ((n - split + lsb + msb) << (msb + lsb)));
}

// HF symbol record: token:8 | cluster:4 | nbits:4 | residue:16
HD uint32_t hf_pack(uint32_t token, uint32_t cluster, uint32_t nbits, uint32_t residue) {
    return token | (cluster << 8) | (nbits << 12) | (residue << 16);
}
HD uint32_t hf_token(uint32_t s) { return s & 0xFFu; }
HD uint32_t hf_cluster(uint32_t s) { return (s >> 8) & 0xFu; }
HD uint32_t hf_nbits(uint32_t s) { return (s >> 12) & 0xFu; }
HD uint32_t hf_residue(uint32_t s) { return s >> 16; }

}  // namespace hydb
