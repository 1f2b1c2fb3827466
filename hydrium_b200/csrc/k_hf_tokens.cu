// hydrium_b200/csrc/k_hf_tokens.cu
//
// Stage 2: quantised HF coefficients -> the tile's ANS symbol stream + per-cluster histograms.
// Replaces initialize_hf_coeffs / hyd_entropy_send_symbol for the HF stream
// (reference: encoder.c:689-750, entropy.c:427-471, 526-544).
//
// The reference walks blocks, channels (Y, X, B) and scan positions serially and stops a block
// after its last non-zero.  Here a warp owns one (block, channel): two ballots give the 64-bit
// non-zero mask, from which every lane derives, for its own scan position j,
//   * whether the symbol exists            (j <= last non-zero)
//   * "non-zeros left" before j            nz - popc(mask below j)
//   * "previous coefficient non-zero"      bit j-1 of the mask (or nz <= 4 for the first)
// and therefore its context cluster, with no serial dependence.  Only the CLUSTER of a context
// matters for the bitstream (the context map collapses the 1485 contexts onto 9 clusters,
// encoder.c:862-877):
//   non-zero-count symbol of channel index i      -> cluster i           (ctx = 3*g(pred)+i, ctx % 3)
//   coefficient, ctx = 458 i + 111 + prev + 2 (n+f) -> 3 + prev + 2 ((i + n + f) mod 3)
// so the neighbour-predicted count of encoder.c:670-687 never influences the output.
// Symbol offsets come from one CTA-wide exclusive scan over the <= 3072 (block, channel) counts.
// Small launches put several CTAs on a tile (gridDim.y parts): each repeats the cheap scan and codes
// its share of the entries, histograms and residue-bit counts are accumulated in HBM with atomics --
// the tokeniser of one tile is issue bound on its SM, and a band of 64 tiles leaves half the SMs idle.
#include "common.cuh"
#include "kernels.h"
#include "tables.cuh"

namespace hydb {

__constant__ uint8_t c_freq_ctx[64] = {HYDB_FREQ_CTX};

constexpr int kTokThreads = 1024;

// nnz_context(left) % 3 for left = 0..63, two bits each (see tables.cuh: 0,0,31,62,62,93 x4,123 x4,152 x8,180 x12,206...)
__device__ __forceinline__ uint32_t nnz_ctx_mod3(uint32_t left) {
    // values mod 3: 0->0, 31->1, 62->2, 93->0, 123->0, 152->2, 180->0, 206->2
    if (left < 2) return 0;
    if (left < 3) return 1;
    if (left < 5) return 2;
    if (left < 13) return 0;
    if (left < 21) return 2;
    if (left < 33) return 0;
    return 2;
}

__global__ void __launch_bounds__(kTokThreads)
k_hf_tokens(const TileDesc *__restrict__ tiles, const int16_t *__restrict__ coef, const uint16_t *__restrict__ nzinfo,
            uint32_t *__restrict__ syms, uint32_t *__restrict__ nsyms, uint32_t *__restrict__ resbits,
            uint32_t *__restrict__ hist, uint32_t *__restrict__ tile_err) {
    __shared__ uint16_t s_info[3 * kMaxBlocks];      // nz | last << 8, symbol order
    __shared__ uint32_t s_off[3 * kMaxBlocks];       // exclusive symbol offsets
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_hist[kHfClusters * kHfTokens];
    __shared__ uint32_t s_resbits, s_err;

    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t part = blockIdx.y, parts = gridDim.y;
    const TileDesc t = tiles[tile];
    if (t.flags & kTilePrefix)   // pseudo-tile of a multi-group frame (k_frame.cu): no pixels
        return;
    const uint32_t vbw = (t.w + 7) >> 3, vbh = (t.h + 7) >> 3, nb = vbw * vbh, ne = 3 * nb;

    for (uint32_t i = tid; i < kHfClusters * kHfTokens; i += kTokThreads)
        s_hist[i] = 0;
    if (tid == 0) {
        s_resbits = 0;
        s_err = 0;
    }
    // entry e = (block raster index) * 3 + channel index i ; channel c = Y, X, B for i = 0, 1, 2
    for (uint32_t e = tid; e < ne; e += kTokThreads) {
        const uint32_t blk = e / 3, i = e - blk * 3, c = i < 2 ? 1 - i : i;
        const uint32_t by = blk / vbw, bx = blk - by * vbw;
        s_info[e] = nzinfo[((size_t)tile * kMaxBlocks + by * kBlocksPerRow + bx) * 3 + c];
    }
    __syncthreads();

    // ---- exclusive scan of symbol counts: 6 consecutive entries per thread --------------------
    constexpr int kPer = 3 * kMaxBlocks / kTokThreads;
    uint32_t cnt[kPer], local = 0;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        const uint32_t e = tid * kPer + k;
        uint32_t c = 0;
        if (e < ne) {
            const uint32_t info = s_info[e];
            c = 1 + ((info & 0xFF) ? (info >> 8) : 0);
        }
        cnt[k] = c;
        local += c;
    }
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d)
            incl += v;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kTokThreads / 32 ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d)
                wi += v;
        }
        s_warp[lane] = wi - w;   // exclusive warp base
        if (lane == 31 && part == 0)
            nsyms[tile] = wi;
    }
    __syncthreads();
    {
        uint32_t run = s_warp[warp] + incl - local;
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            const uint32_t e = tid * kPer + k;
            if (e < ne)
                s_off[e] = run;
            run += cnt[k];
        }
    }
    __syncthreads();

    // ---- one warp per (block, channel) ---------------------------------------------------------
    uint32_t *out = syms + (size_t)tile * kMaxHfSyms;
    const int16_t *tile_coef = coef + (size_t)tile * kMaxBlocks * 3 * 64;
    const uint32_t fc3_lo = c_freq_ctx[lane] % 3u, fc3_hi = c_freq_ctx[lane + 32] % 3u;   // per-lane constants
    uint32_t my_resbits = 0, my_err = 0;
    // The coefficients come from HBM / L2 (hundreds of cycles away) and each warp walks its entries one
    // after the other, so they are fetched kAhead entries ahead into registers.
    constexpr int kAhead = 4;
    constexpr uint32_t kStep = kTokThreads / 32;
    auto fetch = [&](uint32_t e, int &lo, int &hi) {
        lo = hi = 0;
        if (e < ne) {
            const uint32_t info = s_info[e];
            if (info & 0xFF) {
                const uint32_t blk = e / 3, i = e - blk * 3, c = i < 2 ? 1 - i : i;
                const uint32_t by = blk / vbw, bx = blk - by * vbw;
                const int16_t *q = tile_coef + ((by * kBlocksPerRow + bx) * 3 + c) * 64;
                lo = q[lane];
                if ((info >> 8) >= 32)   // warp-uniform: nothing to code in the upper half otherwise
                    hi = q[lane + 32];
            }
        }
    };
    // this CTA's share of the entries: [e_begin, e_end)
    const uint32_t e_begin = (uint32_t)(((uint64_t)ne * part) / parts), e_end = (uint32_t)(((uint64_t)ne * (part + 1)) / parts);
    auto fetch_part = [&](uint32_t e, int &lo, int &hi) {
        if (e < e_end)
            fetch(e, lo, hi);
        else
            lo = hi = 0;
    };
    int buf_lo[kAhead], buf_hi[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; k++)
        fetch_part(e_begin + warp + (uint32_t)k * kStep, buf_lo[k], buf_hi[k]);
    for (uint32_t e0 = e_begin + warp; e0 < e_end; e0 += kStep * kAhead) {
#pragma unroll
        for (int k = 0; k < kAhead; k++) {
            const uint32_t e = e0 + (uint32_t)k * kStep;
            const int q_lo = buf_lo[k], q_hi = buf_hi[k];
            fetch_part(e + kStep * kAhead, buf_lo[k], buf_hi[k]);
            if (e >= e_end)
                continue;
            const uint32_t blk = e / 3, i = e - blk * 3;
            const uint32_t info = s_info[e], nz = info & 0xFF, last = info >> 8;
            const uint32_t base = s_off[e];
            if (lane == 0) {
                // non-zero count on cluster i, hybrid config (4, 1, 0) (encoder.c:908)
                uint32_t res, nbits;
                uint32_t tok = hybrid_token(nz, 4, 1, 0, res, nbits);
                out[base] = hf_pack(tok, i, nbits, res);
                atomicAdd(&s_hist[i * kHfTokens + tok], 1u);
                my_resbits += nbits;
            }
            if (!nz)
                continue;
            const uint32_t m_lo = __ballot_sync(0xFFFFFFFFu, q_lo != 0);
            const uint32_t m_hi = last >= 32 ? __ballot_sync(0xFFFFFFFFu, q_hi != 0) : 0u;
            const uint64_t mask = (uint64_t)m_lo | ((uint64_t)m_hi << 32);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (half && last < 32)
                    break;
                const uint32_t j = lane + 32 * half;
                const int qv = half ? q_hi : q_lo;
                const bool valid = j >= 1 && j <= last;
                uint32_t key = 0xFFFFFFFFu;
                if (valid) {
                    const uint32_t below = __popcll(mask & ((1ull << j) - 1ull));
                    const uint32_t left = nz - below;
                    const uint32_t prev = j == 1 ? (nz <= 4 ? 1u : 0u) : (uint32_t)((mask >> (j - 1)) & 1ull);
                    const uint32_t cluster = 3 + prev + 2 * ((i + nnz_ctx_mod3(left) + (half ? fc3_hi : fc3_lo)) % 3);
                    uint32_t res, nbits;
                    uint32_t tok = hybrid_token(pack_signed(qv), 4, 1, 0, res, nbits);
                    if (tok >= (uint32_t)kHfTokens) {
                        my_err |= kErrAlphabet;
                        tok = kHfTokens - 1;
                    }
                    out[base + j] = hf_pack(tok, cluster, nbits, res);
                    my_resbits += nbits;
                    key = cluster * kHfTokens + tok;
                }
                // warp-aggregated histogram update
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
                if (valid && lane == (uint32_t)(__ffs(peers) - 1))
                    atomicAdd(&s_hist[key], (uint32_t)__popc(peers));
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_resbits += __shfl_down_sync(0xFFFFFFFFu, my_resbits, d);
        my_err |= __shfl_down_sync(0xFFFFFFFFu, my_err, d);
    }
    if (lane == 0) {
        atomicAdd(&s_resbits, my_resbits);
        if (my_err)
            atomicOr(&s_err, my_err);
    }
    __syncthreads();
    if (parts == 1) {
        for (uint32_t i = tid; i < kHfClusters * kHfTokens; i += kTokThreads)
            hist[(size_t)tile * kHfClusters * kHfTokens + i] = s_hist[i];
        if (tid == 0)
            resbits[tile] = s_resbits;
    } else {   // hist / resbits of the launch were zeroed beforehand (launch_hf_tokens)
        for (uint32_t i = tid; i < kHfClusters * kHfTokens; i += kTokThreads)
            if (s_hist[i])
                atomicAdd(&hist[(size_t)tile * kHfClusters * kHfTokens + i], s_hist[i]);
        if (tid == 0 && s_resbits)
            atomicAdd(&resbits[tile], s_resbits);
    }
    if (tid == 0 && s_err)
        atomicOr(&tile_err[tile], s_err);
}

void launch_hf_tokens(const Workspace &ws, uint32_t ntiles, cudaStream_t st) {
    prefer_max_shared(k_hf_tokens);
    // CTAs per tile: fill the 148 SMs when the launch has few tiles (one band of the pipeline, one frame, one tile)
    const uint32_t parts = ntiles <= 37 ? 4u : (ntiles <= 74 ? 2u : 1u);
    if (parts > 1) {
        cudaMemsetAsync(ws.hist, 0, (size_t)ntiles * kHfClusters * kHfTokens * sizeof(uint32_t), st);
        cudaMemsetAsync(ws.resbits, 0, (size_t)ntiles * sizeof(uint32_t), st);
    }
    k_hf_tokens<<<dim3(ntiles, parts), kTokThreads, 0, st>>>(ws.tiles, ws.coef, ws.nzinfo, ws.syms, ws.nsyms, ws.resbits,
                                                             ws.hist, ws.tile_err);
}

}  // namespace hydb
