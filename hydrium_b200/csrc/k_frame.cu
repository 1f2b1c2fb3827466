// hydrium_b200/csrc/k_frame.cu
//
// Frames with more than one 256x256 PassGroup: tile_size_shift 1..3 and one-frame mode for images of
// at most one 2048x2048 LF group (reference: encoder.c:241-325 TOC order, 327-435 frame header,
// 539-629 LF group, 752-1011 frame assembly).
//
// Such a frame occupies 1 + G consecutive workspace slots: a PREFIX pseudo-tile followed by its G
// groups in raster order.  The groups run through the ordinary per-group kernels (colour transform,
// DCT, quantisation, tokeniser, rANS chain, packer); what the frame shares is produced here:
//
//   k_frame_hist_sum   the frame's ANS model is ONE set of histograms over all groups
//                      (hyd_ans_prepare_frequencies over the whole stream, encoder.c:928-931): the
//                      groups' token counts are added up and handed back to every group, so each
//                      chain CTA builds the same model
//   k_frame_lf         LFGroup section: the LF image of the whole frame (up to 256x256 values per
//                      channel, gradient prediction across group borders) through the prefix coder,
//                      plus the constant HF-metadata image
//   k_frame_finish     LFGlobal and HFGlobal sections, TOC with one entry per section, frame header
//                      with the (identity) TOC permutation; everything lands in the prefix slot's
//                      slab so that the ordinary gather concatenates  prefix | group 0 | group 1 ...
//
// First correct version: the prefix coder runs on one thread per frame (the sequential routines of
// prefix_coder.cuh); the LF stream of a 2048x2048 frame has 196 608 values, so this is the slow
// part of a multi-group frame and the obvious next thing to parallelise.
#include "kernels.h"
#include "sections.cuh"
#include "lf_values.cuh"

namespace hydb {

constexpr uint32_t kPrefixLfOffset = 4096;     // byte offset of the LFGroup section inside the prefix slab
constexpr uint32_t kPrefixTailReserve = 4096;  // room kept for the HFGlobal section behind it

struct FrameShared {
    PrefixWork work;
    uint32_t head[256];      // frame header + TOC + LFGlobal, as words
    uint32_t s2[kDBitsWords + 128];
};

__global__ void __launch_bounds__(kHfClusters * kHfTokens)
k_frame_hist_sum(Workspace ws) {
    const uint32_t slot = blockIdx.x, tid = threadIdx.x;
    const TileDesc t = ws.tiles[slot];
    if (!(t.flags & kTilePrefix))
        return;
    constexpr uint32_t kH = kHfClusters * kHfTokens;
    __shared__ uint32_t s_sum[kH];
    uint32_t sum = 0;
    for (uint32_t g = 0; g < t.frame_groups; g++)
        sum += ws.hist[(size_t)(slot + 1 + g) * kH + tid];
    // fold the nine local clusters onto the preset's (fewer) clusters when the frame has many presets
    const uint32_t K = tile_clusters(ws.tiles[slot + 1]);
    s_sum[tid] = 0;
    __syncthreads();
    atomicAdd(&s_sum[hf_fold_cluster(tid / kHfTokens, K) * kHfTokens + tid % kHfTokens], sum);
    __syncthreads();
    for (uint32_t g = 0; g < t.frame_groups; g++)
        ws.hist[(size_t)(slot + 1 + g) * kH + tid] = s_sum[tid];
}

// quantised LF value of channel c at block (bx, by) of the frame
__device__ __forceinline__ int32_t frame_lf_at(const Workspace &ws, uint32_t slot, uint32_t gx_count, uint32_t c,
                                               uint32_t bx, uint32_t by) {
    const uint32_t g = (by >> 5) * gx_count + (bx >> 5);
    return ws.lfq[((size_t)(slot + 1 + g) * 3 + c) * kMaxBlocks + (by & 31u) * kBlocksPerRow + (bx & 31u)];
}

__global__ void __launch_bounds__(256)
k_frame_lf(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    const uint32_t slot = blockIdx.x, tid = threadIdx.x;
    const TileDesc t = ws.tiles[slot];
    if (!(t.flags & kTilePrefix))
        return;
    const uint32_t vbw = (t.frame_w + 7) >> 3, vbh = (t.frame_h + 7) >> 3, nb = vbw * vbh, total = 3 * nb;
    // ---- residuals of the whole LF image, channel order Y, X, B (encoder.c:574-592) ---------------
    uint16_t *resid = reinterpret_cast<uint16_t *>(ws.coef + (size_t)slot * kMaxBlocks * 3 * 64);   // 196 608 x u16
    uint32_t too_big = 0;
    for (uint32_t i = tid; i < total; i += 256) {
        const uint32_t ci = i / nb, r = i - ci * nb;
        const uint32_t c = ci < 2 ? 1 - ci : ci;
        const uint32_t by = r / vbw, bx = r - by * vbw;
        const int32_t v = frame_lf_at(ws, slot, t.frame_gx, c, bx, by);
        const int32_t up = by ? frame_lf_at(ws, slot, t.frame_gx, c, bx, by - 1) : 0;
        const int32_t w = bx ? frame_lf_at(ws, slot, t.frame_gx, c, bx - 1, by) : up;
        const int32_t n = by ? up : w;
        const int32_t nw = (bx && by) ? frame_lf_at(ws, slot, t.frame_gx, c, bx - 1, by - 1) : w;
        const int32_t lo = w < n ? w : n, hi = w < n ? n : w;
        int32_t pred = w + n - nw;
        pred = pred < lo ? lo : (pred > hi ? hi : pred);
        const uint32_t packed = pack_signed(v - pred);
        if (packed > 0xFFFFu)
            too_big = 1;   // would need more than 12 residue bits: outside what the LF coder holds
        resid[i] = (uint16_t)(packed > 0xFFFFu ? 0xFFFFu : packed);
    }
    if (too_big)
        atomicOr(&ws.tile_err[slot], (uint32_t)kErrLfAlphabet);
    __syncthreads();
    if (tid != 0)
        return;
    // ---- LFGroup section, one thread (see the file header) -----------------------------------------
    uint32_t *syms = ws.syms + (size_t)slot * kMaxHfSyms;
    uint8_t *slab = ws.slab + (size_t)slot * kSlabBytes;
    BitSink bw;
    bw.init(reinterpret_cast<uint32_t *>(slab + kPrefixLfOffset), (kSlabBytes - kPrefixLfOffset - kPrefixTailReserve) / 4);
    s.work.error = 0;
    put_lf_group_head(s.work, syms, bw);
    ps_encode_stream(s.work, syms, (uint32_t)kMaxHfSyms, lf_stream_params(), total, StagedValues{resid}, bw);
    put_hf_metadata(s.work, syms, (uint32_t)kMaxHfSyms, bw, vbw, vbh);
    bw.align_byte();
    bw.flush_partial();
    ws.lfbitlen[slot] = bw.bitlen() >> 3;   // bytes of the LFGroup section
    uint32_t err = s.work.error;
    if (bw.overflow)
        err |= kErrSlab;
    if (err)
        atomicOr(&ws.tile_err[slot], err);
}

__global__ void __launch_bounds__(32)
k_frame_finish(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    const uint32_t slot = blockIdx.x;
    const TileDesc t = ws.tiles[slot];
    if (!(t.flags & kTilePrefix) || threadIdx.x != 0)
        return;
    const uint32_t G = t.frame_groups;
    uint32_t *syms = ws.syms + (size_t)slot * kMaxHfSyms;
    uint8_t *slab = ws.slab + (size_t)slot * kSlabBytes;
    uint32_t err = 0;
    for (uint32_t g = 0; g < G; g++)
        err |= ws.tile_err[slot + 1 + g];
    const uint32_t len1 = ws.lfbitlen[slot];
    if (t.flags & kTileLfPart) {
        // one LF group of a larger one-frame image: only its LFGroup section; the frame's head and
        // HFGlobal are assembled when the last LF group has been sent (k_oneframe_finish)
        ws.frame_off[slot] = kPrefixLfOffset;
        ws.frame_len[slot] = err ? 0u : len1;
        if (err)
            atomicOr(&ws.tile_err[slot], err);
        return;
    }
    // ---- HFGlobal section: constants + the ANS header tail the first group's chain kernel wrote ---
    BitSink b2;
    b2.init(s.s2, kDBitsWords + 128);
    s.work.error = 0;
    put_hf_global(s.work, syms, b2, G);
    {
        const uint32_t *d = ws.dbits + (size_t)(slot + 1) * kDBitsWords;
        uint32_t n = ws.chain_out[(slot + 1) * 4 + 2] & 0xFFFFu;
        for (uint32_t i = 0; n; i++) {
            const uint32_t take = n < 32 ? n : 32;
            b2.put(take < 32 ? (d[i] & ((1u << take) - 1u)) : d[i], (int)take);
            n -= take;
        }
    }
    b2.align_byte();
    b2.flush_partial();
    const uint32_t len2 = b2.bitlen() >> 3;
    if (b2.overflow || len2 > kPrefixTailReserve || s.work.error)
        err |= kErrSlab;
    // ---- image header (first frame of a codestream), frame header, TOC, LFGlobal ---------------------
    BitSink bh;
    bh.init(s.head, 256);
    uint32_t pre_bytes = 0;
    uint8_t pre[64];
    if (t.flags & kTileFirst) {
        if (image_needs_level10(t.image_w, t.image_h))
            pre_bytes = put_level10_prefix(pre);
        put_image_header(bh, t.image_w, t.image_h);   // ends byte aligned
    }
    const bool one_frame = (t.flags & kTileOneFrame) != 0;
    const bool crop = !one_frame && (t.image_w > t.frame_w || t.image_h > t.frame_h);
    const bool last = one_frame || (t.flags & kTileLast) != 0;
    put_frame_header_multi(s.work, syms, bh, crop, t.frame_x0, t.frame_y0, t.frame_w, t.frame_h, last, 3 + G);
    bool ok = put_toc_value(bh, 16);           // LFGlobal: 126 bits
    ok = put_toc_value(bh, len1) && ok;
    ok = put_toc_value(bh, len2) && ok;
    for (uint32_t g = 0; g < G; g++)
        ok = put_toc_value(bh, ws.frame_len[slot + 1 + g]) && ok;
    bh.align_byte();
    put_lf_global(bh);
    bh.align_byte();
    bh.flush_partial();
    const uint32_t head_bytes = bh.bitlen() >> 3;
    if (!ok || bh.overflow || s.work.error || pre_bytes + head_bytes > kPrefixLfOffset)
        err |= kErrSlab;
    uint32_t foff = kPrefixLfOffset, flen = 0;
    if (!err) {
        foff = kPrefixLfOffset - head_bytes - pre_bytes;
        for (uint32_t i = 0; i < pre_bytes; i++)
            slab[foff + i] = pre[i];
        for (uint32_t i = 0; i < head_bytes; i++)
            slab[foff + pre_bytes + i] = (uint8_t)(s.head[i >> 2] >> (8 * (i & 3)));
        for (uint32_t i = 0; i < len2; i++)
            slab[kPrefixLfOffset + len1 + i] = (uint8_t)(s.s2[i >> 2] >> (8 * (i & 3)));
        flen = pre_bytes + head_bytes + len1 + len2;
    }
    ws.frame_off[slot] = foff;
    ws.frame_len[slot] = flen;
    if (err)
        atomicOr(&ws.tile_err[slot], err);
}

// ---- one-frame image of several LF groups ---------------------------------------
// Each 2048x2048 LF group went through the pipeline as a frame part (kTileLfPart) with its own HF
// preset = nine ANS clusters of its own (encoder.c:852-877, 689-692); this kernel writes what the
// whole frame shares.  Sections are written LF group by LF group in the order they were sent while
// the TOC lists them in raster order, so the permutation and its Lehmer code are real here
// (encoder.c:241-325).  info layout: hydb_oneframe_finish in engine.cu.
struct ClusterMapMtf {
    const uint16_t *idx;
    HD uint32_t operator()(uint32_t i) const { return (uint32_t)idx[i]; }
};
struct WordValues {
    const uint32_t *v;
    HD uint32_t operator()(uint32_t i) const { return v[i]; }
};

__global__ void __launch_bounds__(256)
k_oneframe_finish(const uint32_t *__restrict__ info, uint32_t info_words, uint32_t *scratch, uint32_t scratch_words,
                  uint8_t *out, uint32_t head_cap, uint32_t hf_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t W = info[0], H = info[1], with_header = info[2], max_alpha = info[3], n = info[4], G = info[5];
    const uint32_t K = hf_clusters_for_presets(n);   // HF clusters per preset
    const uint32_t *sent = info + 8, *len1 = info + 8 + n, *elen = info + 8 + 2 * n;
    const uint32_t cx = (W + 2047) >> 11;
    const uint32_t frame_gx = (W + 255) >> 8, frame_gy = (H + 255) >> 8;
    const uint32_t toc_size = 2 + frame_gx * frame_gy + n;
    uint32_t *res = reinterpret_cast<uint32_t *>(out + head_cap + hf_cap);
    // scratch: [toc: toc_size][inv: toc_size][lehmer+1: toc_size + 1][tokens ...]
    uint32_t *toc = scratch, *inv = toc + toc_size, *leh = inv + toc_size, *tokens = leh + toc_size + 1;
    const uint32_t fixed = 3 * toc_size + 1;
    const bool bad = frame_gx * frame_gy != G || fixed + 1485u * n + 4096u > scratch_words || n == 0 || n > 256;
    if (bad) {
        if (tid == 0) { res[0] = res[1] = 0; res[2] = 1; }
        return;
    }
    if (tid == 0) {   // calculate_toc_perm
        uint32_t idx = 0;
        toc[idx++] = 0;
        for (uint32_t k = 0; k < n; k++)
            toc[idx++] = 1 + sent[k];
        for (uint32_t k = 0; k < n; k++) {
            if (k == 0)
                toc[idx++] = 1 + n;
            const uint32_t lx = sent[k] % cx, ly = sent[k] / cx;
            const uint32_t lw = W - lx * 2048 < 2048 ? W - lx * 2048 : 2048, lh = H - ly * 2048 < 2048 ? H - ly * 2048 : 2048;
            const uint32_t gcx = (lw + 255) >> 8, gcy = (lh + 255) >> 8;
            for (uint32_t g = 0; g < gcx * gcy; g++)
                toc[idx++] = 2 + n + ((ly << 3) + g / gcx) * frame_gx + (lx << 3) + g % gcx;
        }
        for (uint32_t j = 0; j < toc_size; j++)
            inv[toc[j]] = j;
        leh[0] = toc_size;   // the stream's first symbol (encoder.c:413)
    }
    __syncthreads();
    // Lehmer code of inv: how many not yet used smaller elements precede each one
    for (uint32_t i = tid; i < toc_size; i += 256) {
        uint32_t smaller_before = 0;
        const uint32_t v = inv[i];
        for (uint32_t j = 0; j < i; j++)
            smaller_before += inv[j] < v ? 1u : 0u;
        leh[1 + i] = v - smaller_before;
    }
    __syncthreads();
    if (tid != 0)
        return;
    uint32_t err = 0;
    s.work.error = 0;
    // ---- HFGlobal section ------------------------------------------------------------------------------
    uint8_t *hf_bytes = out + head_cap;
    BitSink b2;
    b2.init(reinterpret_cast<uint32_t *>(hf_bytes), hf_cap / 4);
    b2.put_bool(1);                                       // default dequant matrices
    b2.put(n - 1, ceil_log2_u32(G));                      // num_presets - 1 (encoder.c:961)
    b2.put(2, 2);                                         // HF pass order
    b2.put_bool(0);                                       // ANS stream: no lz77
    {   // context map of 1485 n contexts onto K n clusters: never "simple" for n >= 2 (entropy.c:108-167)
        b2.put_bool(0);
        b2.put_bool(1);                                   // move-to-front
        uint16_t *idx = reinterpret_cast<uint16_t *>(tokens + 1485u * n / 2 + 2048u);   // upper part of the token scratch
        uint8_t mtf[256];
        for (int i = 0; i < 256; i++)
            mtf[i] = (uint8_t)i;
        for (uint32_t j = 0; j < 1485u * n; j++) {
            const uint8_t c = (uint8_t)(K * (j / 1485u) + hf_fold_cluster(hf_context_cluster(j % 1485u), K));
            int k = 0;
            while (mtf[k] != c)
                k++;
            idx[j] = (uint16_t)k;
            for (; k > 0; k--)
                mtf[k] = mtf[k - 1];
            mtf[0] = c;
        }
        PrefixParams p;
        p.num_plain_dists = 1;
        p.lz_min_symbol = 64;
        p.modular = 0;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 0;
        p.split1 = 4; p.msb1 = 1; p.lsb1 = 0;
        ps_encode_stream(s.work, tokens, 1485u * n / 2 + 2048u, p, 1485u * n, ClusterMapMtf{idx}, b2);
    }
    const int log_alpha = max_alpha > 32 ? 6 : 5;          // entropy.c:952 with tokens < 64
    if (max_alpha > 64)
        err |= kErrAlphabet;
    b2.put_bool(0);                                       // use_prefix_codes = 0
    b2.put((uint32_t)(log_alpha - 5), 2);
    for (uint32_t c = 0; c < K * n; c++)
        ps_put_hybrid_cfg(b2, 4, 1, 0, log_alpha);
    {
        const uint32_t *p = info + 8 + 2 * n + G;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t nbits = p[0];
            if ((uint32_t)(p - info) + 1 + ((nbits + 31) >> 5) > info_words) {
                err |= kErrSlab;
                break;
            }
            put_bits_from(b2, p + 1, nbits);
            p += 1 + ((nbits + 31) >> 5);
        }
    }
    b2.align_byte();
    b2.flush_partial();
    const uint32_t len2 = b2.bitlen() >> 3;
    if (b2.overflow || s.work.error)
        err |= kErrSlab;
    // ---- head: [image header] frame header + TOC permutation, TOC, LFGlobal ----------------------------
    uint32_t pre_bytes = 0;
    if (with_header && image_needs_level10(W, H))
        pre_bytes = put_level10_prefix(out);
    BitSink bh;
    bh.init(reinterpret_cast<uint32_t *>(out + 52), (head_cap - 52) / 4);   // 52: word aligned, past the 49-byte prefix
    if (with_header)
        put_image_header(bh, W, H);
    put_frame_header_fields(bh, false, 0, 0, W, H, true);
    bh.put_bool(1);                                       // permuted TOC
    {
        PrefixParams p;
        p.num_plain_dists = 8;
        p.lz_min_symbol = 0;
        p.modular = 0;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 1;
        p.split1 = 7; p.msb1 = 0; p.lsb1 = 0;
        s.work.error = 0;
        ps_encode_stream(s.work, tokens, scratch_words - fixed, p, 1 + toc_size, WordValues{leh}, bh);
    }
    bh.align_byte();
    bool ok = put_toc_value(bh, 16);
    for (uint32_t k = 0; k < n; k++)
        ok = put_toc_value(bh, len1[k]) && ok;
    ok = put_toc_value(bh, len2) && ok;
    for (uint32_t g = 0; g < G; g++)
        ok = put_toc_value(bh, elen[g]) && ok;
    bh.align_byte();
    put_lf_global(bh);
    bh.align_byte();
    bh.flush_partial();
    const uint32_t hb = bh.bitlen() >> 3;
    if (!ok || bh.overflow || s.work.error)
        err |= kErrSlab;
    // close the gap between the container prefix (49 bytes, if any) and the header bits written at byte 52
    if (!err)
        for (uint32_t i = 0; i < hb; i++)
            out[pre_bytes + i] = out[52 + i];
    res[0] = pre_bytes + hb;
    res[1] = len2;
    res[2] = err;
}

void launch_oneframe_finish(const uint32_t *d_info, uint32_t info_words, uint32_t *d_scratch, uint32_t scratch_words,
                            uint8_t *d_out, uint32_t head_cap, uint32_t hf_cap, cudaStream_t st) {
    cudaFuncSetAttribute(k_oneframe_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_oneframe_finish<<<1, 256, sizeof(FrameShared), st>>>(d_info, info_words, d_scratch, scratch_words, d_out, head_cap, hf_cap);
}

void launch_frame_hist_sum(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_hist_sum);
    k_frame_hist_sum<<<nslots, kHfClusters * kHfTokens, 0, st>>>(ws);
}

void launch_frame_lf(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_lf);
    cudaFuncSetAttribute(k_frame_lf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_frame_lf<<<nslots, 256, sizeof(FrameShared), st>>>(ws);
}

void launch_frame_finish(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_finish);
    cudaFuncSetAttribute(k_frame_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_frame_finish<<<nslots, 32, sizeof(FrameShared), st>>>(ws);
}

}  // namespace hydb
