// hydrium_b200/csrc/kernels.h
//
// Workspace layout and kernel launchers shared by the .cu translation units and engine.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"

namespace hydb {

struct LutSet {
    const uint16_t *lut8_srgb, *lut8_lin;     // 256 entries   (reference: format.c:58-71)
    const uint16_t *lut16_srgb, *lut16_lin;   // 65536 entries
    const float *bias;                        // 65536 entries (reference: format.c:73-83)
};

// Per-batch device workspace.  Everything is indexed [tile][...] with fixed strides, so kernels
// never need a global scan except for the final compaction.
struct Workspace {
    uint32_t capacity;        // tiles
    uint32_t chain_mode;      // 0 choose by launch size, 1 table kernel, 2 compact kernel (hydb_engine_set_chain_kernel)
    uint32_t chain_lpt;       // set by the launcher: chain CTA b codes tile chain_order[b] (longest chains first)
    TileDesc *tiles;          // [T]
    int16_t *coef;            // [T][1024 blocks (row stride 32)][3 channels X,Y,B][64 scan order]
    uint16_t *nzinfo;         // [T][1024][3]  nz count | last scan index << 8
    int32_t *lfq;             // [T][3][1024]  quantised LF ints
    uint32_t *syms;           // [T][kMaxHfSyms] HF symbol records (hf_pack)
    uint32_t *nsyms;          // [T]
    uint32_t *resbits;        // [T] sum of residue bits of the HF symbols
    uint32_t *hist;           // [T][9][64] raw token counts
    uint32_t *lfbits;         // [T][kLfBitsWords] LF stream bit string (section L)
    uint32_t *lfbitlen;       // [T]
    uint32_t *dbits;          // [T][kDBitsWords] section D bit string
    uint32_t *chain_out;      // [T][4] renorm word count, final state, section D bits, error bits
    uint32_t *flags;          // [T][kMaxHfSyms / 32] renormalisation flag per symbol
    uint16_t *fwords;         // [T][kMaxHfSyms] renormalisation words in chain (reverse) order
    uint8_t *slab;            // [T][kSlabBytes] frames: header+TOC right-justified before byte 64
    uint32_t *frame_off;      // [T] first byte of the frame inside its slab
    uint32_t *frame_len;      // [T]
    uint64_t *out_off;        // [T+1] exclusive scan of frame_len (compaction offsets)
    uint32_t *tile_err;       // [T] TileError bits
    uint32_t *sm_ticket;      // [256] per-SM ticket counter (spreads chain warps over sub-partitions)
    uint32_t *sm_load;        // [256][4] chain warps currently running per SM sub-partition (k_ans_chain_compact)
    uint32_t *chain_order;    // [T] tiles of the launch by descending symbol count (launches with more chains than fit at once)
    // optional stage taps for the parity tests (NULL in production)
    float *dbg_xyb, *dbg_dct; // [T][256][256][3]
    uint32_t *dbg_freqs;      // [T][9][64] normalised frequencies
    uint32_t *dbg_sect;       // [T][4] bit lengths of sections prefix(A+L+B), D, E, total
    uint32_t *dbg_clk;        // [T][4] k_ans_chain: prologue cycles, chain cycles, SM id, chain warp
};

// Shape-constant sections cached in HBM: entry 0 is section A, entry 1 + s is section B of shape s.
struct Templates {
    uint32_t *words;          // [1 + kMaxShapes][kTemplWords]
    uint32_t *bits;           // [1 + kMaxShapes]
};
constexpr int kMaxShapes = 64;

void launch_build_luts(uint16_t *lut8_srgb, uint16_t *lut8_lin, uint16_t *lut16_srgb, uint16_t *lut16_lin,
                       float *bias, cudaStream_t st);
void launch_build_templates(const Templates &t, const uint32_t *d_shape_dims /*[n][2]*/, uint32_t first_shape,
                            uint32_t n_shapes, bool build_a, cudaStream_t st);
void launch_xyb_dct_quant(const Workspace &ws, const LutSet &luts, uint32_t ntiles, cudaStream_t st);
void launch_hf_tokens(const Workspace &ws, uint32_t ntiles, cudaStream_t st);
void launch_lf_group(const Workspace &ws, uint32_t ntiles, cudaStream_t st);
// allow_compact: no HYD_FLOAT32 tile in the launch (tokens stay below 32); large launches then take
// k_ans_chain_compact (sixteen chains per SM), small ones the table kernel (two per SM, shorter steps)
// longest_first: the launch holds plain tiles (every slot has a symbol count): when it has more chains than
// the GPU keeps resident, the CTAs take the tiles by descending symbol count, so that the chains that start
// last are the short ones (the launch ends when its last chain does)
void launch_ans_chain(const Workspace &ws, uint32_t ntiles, cudaStream_t st, bool allow_compact = true,
                      uint32_t concurrent_tiles = 0, bool longest_first = false);
void launch_ans_chain_compact(const Workspace &ws, uint32_t ntiles, cudaStream_t st);
int ans_compact_smem_bytes();
void launch_ans_pack(const Workspace &ws, const Templates &t, uint32_t ntiles, cudaStream_t st);
// frames with several groups (k_frame.cu): shared ANS model, LFGroup section, frame prefix
void launch_frame_hist_sum(const Workspace &ws, uint32_t nslots, cudaStream_t st);
void launch_frame_lf(const Workspace &ws, uint32_t nslots, cudaStream_t st);
void launch_frame_finish(const Workspace &ws, uint32_t nslots, cudaStream_t st);
// one-frame image of several LF groups: head (header, TOC permutation, TOC, LFGlobal) and HFGlobal section;
// out = [head: head_cap bytes][hf_global: hf_cap bytes][4 words: head_len, hf_len, error, -]
void launch_oneframe_finish(const uint32_t *d_info, uint32_t info_words, uint32_t *d_scratch, uint32_t scratch_words,
                            uint8_t *d_out, uint32_t head_cap, uint32_t hf_cap, uint32_t *d_ctx_cache, uint32_t ctx_cache_words,
                            uint32_t ctx_cached_bits, uint32_t *d_perm_cache, uint32_t perm_cache_words, uint32_t perm_cached_bits,
                            cudaStream_t st);
// image header of an ICC-tagged image: [49-byte container prefix] signature, size, metadata, entropy-coded
// profile, byte aligned; d_res = {bytes, error}
void launch_icc_header(const uint8_t *d_icc, uint32_t n, uint32_t W, uint32_t H, uint32_t *d_bits, uint32_t bits_words,
                       uint8_t *d_out, uint32_t out_cap, uint32_t *d_res, cudaStream_t st);
// compaction: out[prefix_len + out_off[i] ...] = frame i ; total written to ws.out_off[ntiles]
void launch_gather(const Workspace &ws, uint32_t ntiles, uint8_t *out, uint64_t out_cap, uint64_t base,
                   uint32_t *d_overflow, cudaStream_t st);
void launch_job_result(const uint32_t *tile_err, uint32_t n, const uint64_t *total, const uint32_t *overflow, uint64_t *h_res,
                       cudaStream_t st);
// spans written by every rank into its region of a buffer on the gathering rank -> one contiguous stream
void launch_compact_regions(const uint8_t *regions, uint32_t nregions, uint64_t region_stride, uint8_t *out, uint64_t out_cap,
                            uint64_t *d_total, uint32_t *d_overflow, cudaStream_t st);
// one 64-bit word, stream ordered (a span length published to the gathering rank's memory)
void launch_store_u64(uint64_t *dst, uint64_t v, cudaStream_t st);
void launch_synth_fill(void *dst, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0, uint32_t full_w,
                       uint32_t full_h, int bits, uint32_t seed, int smooth, cudaStream_t st);
int ans_encode_smem_bytes();

// Every kernel of the pipeline asks for the same (largest) shared-memory carve-out: an SM only
// changes its L1 / shared split when it is idle, so kernels with different preferences cannot share
// an SM and the rANS chain CTAs of an early band (95 KB each) would otherwise wait for the front-end
// kernels of all later bands to drain.  HYDRIUM_B200_CARVEOUT=0 leaves the driver's default.
template <typename K>
inline void prefer_max_shared(K kernel) {
    static const bool on = [] { const char *e = getenv("HYDRIUM_B200_CARVEOUT"); return !(e && e[0] == '0'); }();
    if (on)
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

}  // namespace hydb
