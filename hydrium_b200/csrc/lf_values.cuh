// hydrium_b200/csrc/lf_values.cuh
//
// LF (DC) coefficients as a modular sub-image: clamped-gradient prediction from the W / N / NW
// neighbours of the same channel, residual folded to unsigned (reference: encoder.c:573-592).
// The quantised LF ints are produced by k_xyb_dct.cu ( trunc(dc * {8192, 1024, 512}) ) and stored
// per channel with a fixed row stride of 32 blocks.
#pragma once

#include "common.cuh"
#include "prefix_coder.cuh"

namespace hydb {

HD uint32_t lf_residual(const int32_t *plane, uint32_t bx, uint32_t by) {
    const int32_t v = plane[by * kBlocksPerRow + bx];
    const int32_t up = by ? plane[(by - 1) * kBlocksPerRow + bx] : 0;
    const int32_t w = bx ? plane[by * kBlocksPerRow + bx - 1] : up;
    const int32_t n = by ? up : w;
    const int32_t nw = (bx && by) ? plane[(by - 1) * kBlocksPerRow + bx - 1] : w;
    const int32_t lo = w < n ? w : n, hi = w < n ? n : w;
    int32_t pred = w + n - nw;
    pred = pred < lo ? lo : (pred > hi ? hi : pred);
    return pack_signed(v - pred);
}

// i-th value of the LF stream of a vbw x vbh tile: channel order Y, X, B (encoder.c:574-575),
// raster order inside a channel.  `lfq` = three planes of kMaxBlocks ints (X, Y, B).
struct LfStreamValues {
    const int32_t *lfq;
    uint32_t vbw, nb;
    HD uint32_t operator()(uint32_t i) const {
        const uint32_t ci = i / nb, r = i - ci * nb;
        const uint32_t c = ci < 2 ? 1 - ci : ci;
        const uint32_t by = r / vbw, bx = r - by * vbw;
        return lf_residual(lfq + c * kMaxBlocks, bx, by);
    }
};

// one context, config (7,1,1) on both clusters, run-length tokens from 16384, modular
// (reference: encoder.c:567-572)
HD PrefixParams lf_stream_params() {
    PrefixParams p;
    p.num_plain_dists = 1;
    p.lz_min_symbol = 1u << 14;
    p.modular = 1;
    p.split0 = 7; p.msb0 = 1; p.lsb0 = 1;
    p.split1 = 7; p.msb1 = 1; p.lsb1 = 1;
    return p;
}

}  // namespace hydb
