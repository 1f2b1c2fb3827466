// hydrium_b200/csrc/sections.cuh
//
// Payload sections that do not depend on pixel data.  One tile-mode frame's single TOC section is
//
//   [A] LFGlobal | LFGroup preamble | MA tree (gradient predictor)              -- constant
//   [L] LF coefficient stream                                                   -- per tile (k_lf_group.cu)
//   [B] nb_blocks | 0x2 | MA tree (zero predictor) | HF-metadata stream
//       | HFGlobal | ANS preamble up to and including the context map            -- per (vbw, vbh) shape
//   [D] use_prefix=0 | log_alphabet | 9 hybrid configs | 9 histograms            -- per tile (k_ans.cu)
//   [E] PassGroup: final ANS state, renormalisation words, residue bits          -- per tile (k_ans.cu)
//
// (reference: encoder.c:510-629 for A/L/B, encoder.c:852-911 + entropy.c:108-167 for the context
// map, encoder.c:959-967 for HFGlobal).  A and B are produced once per engine / per shape by the
// same device prefix coder that codes L, and cached in HBM as bit strings.
#pragma once

#include "headers.cuh"
#include "prefix_coder.cuh"

namespace hydb {

constexpr int kHfContexts = 1485;   // reference: encoder.c:853
constexpr int kSectionSymCap = 3200;

// cluster of an HF context (reference: encoder.c:862-877, one preset)
HD uint32_t hf_context_cluster(uint32_t ctx) { return ctx < 111u ? ctx % 3u : 3u + (ctx - 111u) % 6u; }

// value generators for the constant streams (plain functors: usable on host and device)
struct MaTreeValues {
    uint32_t predictor;
    HD uint32_t operator()(uint32_t i) const { return i == 1 ? predictor : 0u; }
};
struct HfMetaValues {
    uint32_t zeros_pre, nb;
    HD uint32_t operator()(uint32_t i) const { return (i >= zeros_pre && i < zeros_pre + nb) ? 8u : 0u; }
};
struct StagedValues {
    const uint16_t *v;
    HD uint32_t operator()(uint32_t i) const { return (uint32_t)v[i]; }
};

// five-node MA tree: property -1 (leaf) with the given predictor (reference: encoder.c:114-116, 552-564, 600-610)
HDN inline void put_ma_tree(PrefixWork &w, uint32_t *syms, BitSink &bw, uint32_t predictor) {
    PrefixParams p;
    p.num_plain_dists = 6;
    p.lz_min_symbol = 0;
    p.modular = 0;
    p.split0 = 4; p.msb0 = 1; p.lsb0 = 1;
    p.split1 = 7; p.msb1 = 0; p.lsb1 = 0;
    ps_encode_stream(w, syms, kSectionSymCap, p, 5, MaTreeValues{predictor}, bw);
}

// section A
HDN inline void build_section_a(PrefixWork &w, uint32_t *syms, BitSink &bw) {
    put_lf_global(bw);
    bw.put(0, 2);      // extra precision
    bw.put_bool(0);    // use global tree
    bw.put_bool(1);    // wp_params all_default
    bw.put(0, 2);      // nb_transforms
    put_ma_tree(w, syms, bw, 5);
}

// tail of the LF group: nb_blocks, the zero-predictor MA tree and the constant HF-metadata image of
// an LF group of vbw x vbh varblocks (encoder.c:598-626).  `cap` = words of symbol scratch.
HDN inline void put_hf_metadata(PrefixWork &w, uint32_t *syms, uint32_t cap, BitSink &bw, uint32_t vbw, uint32_t vbh) {
    const uint32_t nb = vbw * vbh;
    bw.put(nb - 1, ceil_log2_u32(nb));
    bw.put(2, 4);
    put_ma_tree(w, syms, bw, 0);
    {   // HF metadata: cfl x/b factors + block info, all constant (encoder.c:611-626)
        const uint32_t cfl = ((vbw + 7) / 8) * ((vbh + 7) / 8);
        const uint32_t zeros_pre = 2 * cfl + nb, total = zeros_pre + 2 * nb;
        PrefixParams p;
        p.num_plain_dists = 1;
        p.lz_min_symbol = 29;
        p.modular = 1;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 1;
        p.split1 = 7; p.msb1 = 0; p.lsb1 = 0;
        // three constant stretches: zeros, (hf_mult - 1) * 2 = 8 for every block, zeros
        (void)total;
        const PsRun runs[3] = {{0u, zeros_pre}, {8u, nb}, {0u, nb}};
        ps_tokenize_runs(w, syms, cap, p, runs, 3);
        ps_put_header(w, bw, p);
        for (uint32_t i = 0; i < w.nsyms; i++)
            ps_put_symbol(w, bw, syms[i], p.lz_min_symbol);
    }
}

// HFGlobal and the head of the ANS stream header, up to and including the context map
// (encoder.c:959-964, entropy.c:108-167).  `frame_groups` = PassGroups in the frame: the preset
// count is written in ceil(log2(groups)) bits.
HDN inline void put_hf_global(PrefixWork &w, uint32_t *syms, BitSink &bw, uint32_t frame_groups);

// section B for a tile of vbw x vbh varblocks
HDN inline void build_section_b(PrefixWork &w, uint32_t *syms, BitSink &bw, uint32_t vbw, uint32_t vbh) {
    put_hf_metadata(w, syms, kSectionSymCap, bw, vbw, vbh);
    put_hf_global(w, syms, bw, 1);
}

HDN inline void put_hf_global(PrefixWork &w, uint32_t *syms, BitSink &bw, uint32_t frame_groups) {
    bw.put_bool(1);    // HFGlobal: default dequant matrices
    bw.put(0, ceil_log2_u32(frame_groups));   // num_presets - 1 = 0
    bw.put(2, 2);      // HF pass order
    bw.put_bool(0);    // ANS stream: no lz77
    bw.put_bool(0);    // context map: not simple
    bw.put_bool(1);    // move-to-front
    {   // nested prefix stream of move-to-front indices (entropy.c:125-158)
        uint8_t mtf[16];
        for (int i = 0; i < 16; i++)
            mtf[i] = (uint8_t)i;
        // MTF indices are a pure function of position here, but the list is evolved literally;
        // they are staged at the top of the symbol scratch while the tokens are written below
        uint16_t *idx = (uint16_t *)(syms + kSectionSymCap) - kHfContexts - 1;
        for (int j = 0; j < kHfContexts; j++) {
            const uint8_t c = (uint8_t)hf_context_cluster((uint32_t)j);
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
        // tokens produced (<= ~60) never reach the staged indices (upper ~750 words)
        ps_encode_stream(w, syms, kSectionSymCap - 800, p, kHfContexts, StagedValues{idx}, bw);
    }
}

// ---- frames with more than one PassGroup (tile_size_shift > 0, one-frame mode) -----------------
// TOC of such a frame: LFGlobal, LFGroup, HFGlobal, then the groups in raster order -- which is the
// order the sections are written in, so the permutation the reference encodes is the identity
// (encoder.c:241-325) and its Lehmer code is all zeros.
struct TocPermValues {
    uint32_t toc_size;
    HD uint32_t operator()(uint32_t i) const { return i == 0 ? toc_size : 0u; }
};

// frame header of a multi-section frame, byte aligned at the end (encoder.c:327-435)
HDN inline void put_frame_header_multi(PrefixWork &w, uint32_t *syms, BitSink &bw, bool crop, uint32_t x0, uint32_t y0,
                                       uint32_t fw, uint32_t fh, bool last, uint32_t toc_size) {
    put_frame_header_fields(bw, crop, x0, y0, fw, fh, last);
    bw.put_bool(1);                // permuted TOC
    PrefixParams p;
    p.num_plain_dists = 8;
    p.lz_min_symbol = 0;
    p.modular = 0;
    p.split0 = 4; p.msb0 = 1; p.lsb0 = 1;
    p.split1 = 7; p.msb1 = 0; p.lsb1 = 0;
    ps_encode_stream(w, syms, kSectionSymCap, p, 1 + toc_size, TocPermValues{toc_size}, bw);
    bw.align_byte();
}

// head of an LFGroup section: modular sub-image preamble + gradient-predictor MA tree (encoder.c:539-564)
HDN inline void put_lf_group_head(PrefixWork &w, uint32_t *syms, BitSink &bw) {
    bw.put(0, 2);      // extra precision
    bw.put_bool(0);    // use global tree
    bw.put_bool(1);    // wp_params all_default
    bw.put(0, 2);      // nb_transforms
    put_ma_tree(w, syms, bw, 5);
}

}  // namespace hydb
