// hydrium_b200/csrc/headers.cuh
//
// The small fixed-layout JPEG XL header fields of one tile-mode frame:
// image header (reference: encoder.c:164-239), frame header (encoder.c:327-435 with a single,
// unpermuted TOC entry), TOC entry (encoder.c:992-1005) and LFGlobal (encoder.c:510-537).
#pragma once

#include "bitio.cuh"
#include "common.cuh"

namespace hydb {

// SizeHeader + ImageMetadata, byte aligned at the end.  No ICC (one-frame-only feature).
HDN inline void put_image_header(BitSink &bw, uint32_t width, uint32_t height) {
    const U32Dist kSize = {{1, 1, 1, 1}, {9, 13, 18, 30}};   // encoder.c:98-101
    bw.put(0x0AFF, 17);            // signature ff 0a, div8 = 0
    put_u32(bw, kSize, height);
    bw.put(0, 3);                  // ratio = 0
    put_u32(bw, kSize, width);
    bw.put_bool(0);                // ImageMetadata all_default
    bw.put_bool(0);                // extra_fields
    bw.put_bool(0);                // float samples
    bw.put(0, 2);                  // 8-bit
    bw.put_bool(1);                // modular 16-bit buffers
    bw.put(0, 2);                  // no extra channels
    bw.put_bool(1);                // xyb_encoded
    bw.put_bool(1);                // colour encoding all_default
    put_u64(bw, 0);                // extensions
    bw.put_bool(1);                // default_matrix
    bw.align_byte();
}

// The same header for an image tagged with an ICC profile (encoder.c:203-225): colour encoding
// all_default = 0, want_icc = 1, colour space RGB; the entropy-coded profile follows default_matrix
// and the byte alignment comes after it (k_icc_header in k_frame.cu).
HDN inline void put_image_header_icc_fields(BitSink &bw, uint32_t width, uint32_t height, uint64_t icc_stream_bytes) {
    const U32Dist kSize = {{1, 1, 1, 1}, {9, 13, 18, 30}};
    bw.put(0x0AFF, 17);
    put_u32(bw, kSize, height);
    bw.put(0, 3);
    put_u32(bw, kSize, width);
    bw.put_bool(0);                // ImageMetadata all_default
    bw.put_bool(0);                // extra_fields
    bw.put_bool(0);                // float samples
    bw.put(0, 2);                  // 8-bit
    bw.put_bool(1);                // modular 16-bit buffers
    bw.put(0, 2);                  // no extra channels
    bw.put_bool(1);                // xyb_encoded
    bw.put_bool(0);                // colour encoding all_default = 0
    bw.put_bool(1);                // want_icc
    bw.put(0, 2);                  // colour space enum: RGB (U32 selector 0 = value 0)
    put_u64(bw, 0);                // extensions
    bw.put_bool(1);                // default_matrix
    put_u64(bw, icc_stream_bytes); // encoded ICC stream length
}

// level-10 container prefix (encoder.c:23-30), emitted when libhydrium.c:67-68 says so
HD bool image_needs_level10(uint64_t w, uint64_t h) { return w > (1u << 20) || h > (1u << 20) || w * h > (1u << 28); }
HDN inline uint32_t put_level10_prefix(uint8_t *dst) {
    const uint8_t k[49] = {
        0, 0, 0, 0x0c, 'J', 'X', 'L', ' ', 0x0d, 0x0a, 0x87, 0x0a, 0, 0, 0, 0x14, 'f', 't', 'y', 'p',
        'j', 'x', 'l', ' ', 0, 0, 0, 0, 'j', 'x', 'l', ' ', 0, 0, 0, 9, 'j', 'x', 'l', 'l', 0x0a,
        0, 0, 0, 0, 'j', 'x', 'l', 'c',
    };
    for (int i = 0; i < 49; i++)
        dst[i] = k[i];
    return 49;
}

// Frame header up to (not including) the TOC-permutation flag (encoder.c:327-398).
HDN inline void put_frame_header_fields(BitSink &bw, bool crop, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                                        bool last) {
    const U32Dist kFrameSize = {{0, 256, 2304, 18688}, {8, 11, 14, 30}};   // encoder.c:102-105
    bw.put(0, 1);                  // all_default = 0
    bw.put(last ? 0u : 3u, 2);     // regular frame / skip-progressive
    bw.put(0, 1);                  // VarDCT
    put_u64(bw, 0x80);             // flags: skip adaptive LF smoothing
    bw.put(0x4C, 10);              // upsampling 0, x_qm_scale 3, b_qm_scale 2, num_passes 0
    bw.put_bool(crop);
    if (crop) {
        put_u32(bw, kFrameSize, pack_signed((int32_t)x0));
        put_u32(bw, kFrameSize, pack_signed((int32_t)y0));
        put_u32(bw, kFrameSize, w);
        put_u32(bw, kFrameSize, h);
    }
    bw.put(0, 2);                  // blending: replace
    if (crop)
        bw.put(0, 2);              // blending source
    bw.put_bool(last);
    if (!last)
        bw.put(0, 2);              // save_as_reference
    bw.put(0, 2);                  // name_len
    bw.put_bool(0);                // restoration filter all_default = 0
    bw.put_bool(0);                // gab
    bw.put(0, 2);                  // epf_iters
    bw.put(0, 2);                  // extensions
    bw.put(0, 2);                  // frame header extensions
}

// Frame header of one 256x256-group frame, byte aligned at both ends.
HDN inline void put_frame_header(BitSink &bw, bool crop, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, bool last) {
    put_frame_header_fields(bw, crop, x0, y0, w, h, last);
    bw.put_bool(0);                // TOC not permuted (single section)
    bw.align_byte();
}

// one TOC entry, no alignment (multi-section frames write several back to back, encoder.c:996-1001)
HDN inline bool put_toc_value(BitSink &bw, uint32_t section_bytes) {
    const U32Dist kToc = {{0, 1024, 17408, 4211712}, {10, 14, 22, 30}};    // encoder.c:117-120
    return put_u32(bw, kToc, section_bytes);
}

HDN inline bool put_toc_entry(BitSink &bw, uint32_t payload_bytes) {
    const bool ok = put_toc_value(bw, payload_bytes);
    bw.align_byte();
    return ok;
}

// LFGlobal: 126 constant bits (encoder.c:510-537)
HDN inline void put_lf_global(BitSink &bw) {
    const U32Dist kGlobalScale = {{1, 2049, 4097, 8193}, {11, 11, 12, 16}};   // encoder.c:106-109
    const U32Dist kQuantLf = {{16, 1, 1, 1}, {0, 5, 8, 16}};                  // encoder.c:110-113
    bw.put_bool(1);                // LF dequant all_default
    put_u32(bw, kGlobalScale, 32768);
    put_u32(bw, kQuantLf, 4);
    bw.put_bool(0);                // HF block context not default
    bw.put(0, 16);                 // no lf / qf thresholds
    bw.put_bool(1);                // simple clustering
    bw.put(2, 2);                  // 2 bits per entry
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 13; j++)
            bw.put((uint32_t)i, 2);
    bw.put_bool(1);                // default LF channel correlation
    bw.put_bool(0);                // no global MA tree
}

}  // namespace hydb
