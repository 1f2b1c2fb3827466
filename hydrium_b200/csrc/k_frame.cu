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
//   k_oneframe_finish  head and HFGlobal of a one-frame image of several LF groups
//   k_icc_header       image header of an ICC-tagged image (the profile as a 41-context prefix stream)
//
// k_frame_lf and k_icc_header spread the per-value work (residuals, tokens, histograms, symbol bits)
// over 1024 threads with block scans; what is inherently small and sequential (code construction,
// stream headers) runs on one thread through prefix_coder.cuh.
#include <cooperative_groups.h>

#include "kernels.h"
#include "sections.cuh"
#include "lf_values.cuh"
#include "prefix_warp.cuh"

namespace hydb {

constexpr uint32_t kPrefixLfOffset = 4096;     // byte offset of the LFGroup section inside the prefix slab
constexpr uint32_t kPrefixTailReserve = 4096;  // room kept for the HFGlobal section behind it

struct FrameShared {
    PrefixWork work;
    uint32_t head[256];      // frame header + TOC + LFGlobal, as words
    uint32_t s2[kDBitsWords + 128];
    uint16_t mapidx[1488];   // move-to-front indices of one preset's 1485-entry context map
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

// ---- block-wide scans over kLfThreads threads ------------------------------------------------------
constexpr int kLfThreads = 1024;

// exclusive scan; OP(a, b) associative, `identity` its neutral element.  Also returns the total.
template <typename Op>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t identity, Op op, uint32_t *s_warp,
                                                         uint32_t &total) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d)
            incl = op(o, incl);
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (uint32_t)d)
                wi = op(o, wi);
        }
        s_warp[32 + lane] = wi;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_excl = warp ? s_warp[32 + warp - 1] : identity;
    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    const uint32_t excl = op(warp_excl, lane ? prev : identity);
    total = s_warp[63];
    __syncthreads();
    return excl;
}

// what position i of the LF value sequence emits (the reference's run-length front end,
// entropy.c:473-524, as a per-position rule): s / e = first position and end of its maximal run
struct LfEmit {
    uint32_t n;          // 0, 1 or 2 symbols
    uint32_t sym[2];
};
__device__ __forceinline__ LfEmit lf_position_symbols(uint32_t v, uint32_t i, uint32_t s, uint32_t e,
                                                      const PrefixParams &p) {
    LfEmit out;
    out.n = 0;
    const uint32_t o = i - s, r = o & 127u, L = (e - s) - (o & ~127u) < 128u ? (e - s) - (o & ~127u) : 128u;
    uint32_t res, nb;
    if (r == 0 || L - 1 <= 3 || !p.lz_min_symbol) {   // the chunk's literal, or one of its (at most three) repeats
        const uint32_t tok = hybrid_token(v, p.split0, p.msb0, p.lsb0, res, nb);
        out.sym[0] = ps_pack(tok, 0, nb, res);
        out.n = 1;
    } else if (r == 1) {          // length token + distance symbol
        out.sym[0] = ps_pack(p.lz_min_symbol + (L - 1 - 3), 0, 0, 0);
        const uint32_t tok = hybrid_token(p.modular ? 1u : 0u, p.split1, p.msb1, p.lsb1, res, nb);
        out.sym[1] = ps_pack(tok, 1, nb, res);
        out.n = 2;
    }
    return out;
}

// One prefix-coded stream by the whole CTA (kLfThreads threads): the values value_at(0 .. total) with the
// stream parameters `prm`, appended to the bit string in `outw` that thread 0's `bw` has written so
// far.  The run-length front end is applied as a per-position rule, so symbols are counted, placed
// (block scans) and histogrammed by all threads; the first warp finds the code lengths, thread 0
// writes the stream header, all threads pack the symbol bits (scanned bit offsets, atomicOr at word
// boundaries).  On return thread 0's `bw` continues behind the stream.  Returns error flags (uniform).
// `syms`: symbol scratch of sym_cap words; s_first: kLfThreads words, s_warp: 64 words of shared memory.
template <typename ValueAt>
__device__ uint32_t block_prefix_stream(PrefixWork &w, BitSink &bw, uint32_t *outw, uint32_t out_words, uint32_t *syms,
                                        uint32_t sym_cap, const PrefixParams &prm, uint32_t total, ValueAt value_at,
                                        uint32_t *s_warp, uint32_t *s_first) {
    __shared__ uint32_t s_bitpos, s_err;
    const uint32_t tid = threadIdx.x;
    const uint32_t lz = prm.lz_min_symbol;
    __syncthreads();
    for (uint32_t i = tid; i < (uint32_t)kAllBins; i += kLfThreads)
        w.freq[i] = 0;
    if (tid == 0) {
        w.alpha0 = w.alpha1 = 0;
        s_err = 0;
    }
    __syncthreads();
    uint32_t my_err = 0;
    // ---- symbols: thread = contiguous range of positions -------------------------------------------
    const uint32_t per = (total + kLfThreads - 1) / kLfThreads;
    const uint32_t a = tid * per < total ? tid * per : total, b = a + per < total ? a + per : total;
    // without run-length mode every position is its own run
    auto starts_run = [&](uint32_t i) -> bool { return !lz || i == 0 || value_at(i) != value_at(i - 1); };
    // last run start inside [a, b) (+1; 0 = none) and first run start inside it (total = none)
    uint32_t last_start = 0, first_start = total;
    for (uint32_t i = a; i < b; i++) {
        if (starts_run(i)) {
            last_start = i + 1;
            if (first_start == total)
                first_start = i;
        }
    }
    uint32_t dummy;
    const uint32_t carry_start = block_exclusive_scan(last_start, 0u, [](uint32_t x, uint32_t y) { return x > y ? x : y; },
                                                      s_warp, dummy);   // start (+1) of the run reaching into a
    // next run start at or after b: exclusive min-scan from the right, done on mirrored thread order
    uint32_t carry_end;
    {
        s_first[kLfThreads - 1 - tid] = first_start;
        __syncthreads();
        const uint32_t mirrored = s_first[tid];
        const uint32_t ex = block_exclusive_scan(mirrored, total, [](uint32_t x, uint32_t y) { return x < y ? x : y; },
                                                 s_warp, dummy);
        s_first[tid] = ex;
        __syncthreads();
        carry_end = s_first[kLfThreads - 1 - tid];
        __syncthreads();
    }
    auto run_end = [&](uint32_t i) -> uint32_t {   // end of the run containing position i (i in [a, b))
        uint32_t e = i + 1;
        while (e < b && !starts_run(e))
            e++;
        return e < b ? e : carry_end;
    };
    const uint32_t sr0 = (a < b && !starts_run(a)) ? carry_start - 1 : a;
    // pass 1: count
    uint32_t count = 0;
    {
        uint32_t i = a, sr = sr0;
        while (i < b) {
            const uint32_t e = run_end(i);
            const uint32_t v = value_at(i);
            const uint32_t stop = e < b ? e : b;
            for (; i < stop; i++)
                count += lf_position_symbols(v, i, sr, e, prm).n;
            sr = i;
        }
    }
    uint32_t nsyms_total;
    uint32_t off = block_exclusive_scan(count, 0u, [](uint32_t x, uint32_t y) { return x + y; }, s_warp, nsyms_total);
    if (nsyms_total > sym_cap)
        my_err |= kErrLfCapacity;
    // pass 2: emit + histogram
    if (nsyms_total <= sym_cap) {
        uint32_t i = a, sr = sr0;
        while (i < b) {
            const uint32_t e = run_end(i);
            const uint32_t v = value_at(i);
            const uint32_t stop = e < b ? e : b;
            for (; i < stop; i++) {
                const LfEmit em = lf_position_symbols(v, i, sr, e, prm);
                for (uint32_t k = 0; k < em.n; k++) {
                    const uint32_t sym = em.sym[k];
                    const uint32_t token = sym & 0x7FFFu, cluster = (sym >> 15) & 1u, nbits = (sym >> 16) & 0xFu;
                    bool fits;
                    if (cluster)
                        fits = token < (uint32_t)kDistBins;
                    else if (lz && token >= lz)
                        fits = token - lz < (uint32_t)kLzBins;
                    else
                        fits = token < (uint32_t)kLitBins;
                    if (!fits || nbits > 12) {
                        my_err |= kErrLfAlphabet;
                        syms[off++] = ps_pack(0, 0, 0, 0);
                        continue;
                    }
                    syms[off++] = sym;
                    atomicAdd(&w.freq[ps_bin(token, cluster, lz)], 1u);
                    if (cluster)
                        atomicMax(&w.alpha1, token + 1);
                    else
                        atomicMax(&w.alpha0, token + 1);
                }
            }
            sr = i;
        }
    }
    if (my_err)
        atomicOr(&s_err, my_err);
    __syncthreads();
    // ---- code lengths of the literal / length cluster by the first warp, then the stream header --------
    const bool warp_lengths = w.alpha0 > 1;
    if (tid < 32 && warp_lengths)
        warp_code_lengths(w, w.alpha0, 15, lz, tid);
    __syncthreads();
    if (tid == 0) {
        w.nsyms = nsyms_total;
        ps_put_header(w, bw, prm, warp_lengths);
        bw.flush_partial();
        s_bitpos = bw.bitlen();
        if (bw.overflow)
            s_err |= kErrSlab;
        if (w.error)
            s_err |= w.error;
    }
    __syncthreads();
    const uint32_t p0 = s_bitpos;
    // ---- symbol bits, in parallel: thread = contiguous range of symbols -----------------------------
    uint32_t total_bits = 0;
    if (!s_err) {
        const uint32_t sper = (nsyms_total + kLfThreads - 1) / kLfThreads;
        const uint32_t sa = tid * sper < nsyms_total ? tid * sper : nsyms_total;
        const uint32_t sb = sa + sper < nsyms_total ? sa + sper : nsyms_total;
        uint32_t bits = 0;
        for (uint32_t i = sa; i < sb; i++) {
            const uint32_t sym = syms[i];
            bits += w.len[ps_bin(sym & 0x7FFFu, (sym >> 15) & 1u, lz)] + ((sym >> 16) & 0xFu);
        }
        const uint32_t boff = block_exclusive_scan(bits, 0u, [](uint32_t x, uint32_t y) { return x + y; }, s_warp, total_bits);
        const uint64_t endbit = (uint64_t)p0 + total_bits;
        if (endbit + 64 > (uint64_t)out_words * 32) {
            __syncthreads();
            if (tid == 0)
                s_err |= kErrSlab;
        } else {
            // flush_partial() wrote word p0 >> 5 only when the header left a partial word there
            for (uint32_t wd = ((p0 + 31) >> 5) + tid; wd <= (uint32_t)(endbit >> 5) + 1; wd += kLfThreads)
                outw[wd] = 0;
            __syncthreads();
            uint64_t pos = (uint64_t)p0 + boff;
            uint32_t wpos = (uint32_t)(pos >> 5), nacc = (uint32_t)(pos & 31u);
            uint64_t acc = 0;
            auto put = [&](uint32_t v, uint32_t n) {
                acc |= (uint64_t)v << nacc;
                nacc += n;
                if (nacc >= 32) {
                    atomicOr(&outw[wpos], (uint32_t)acc);
                    wpos++;
                    acc >>= 32;
                    nacc -= 32;
                }
            };
            for (uint32_t i = sa; i < sb; i++) {
                const uint32_t sym = syms[i];
                const uint32_t bin = ps_bin(sym & 0x7FFFu, (sym >> 15) & 1u, lz);
                const uint32_t nbits = (sym >> 16) & 0xFu;
                if (w.len[bin])
                    put(w.code[bin], w.len[bin]);
                if (nbits)
                    put(sym >> 20, nbits);
            }
            if (nacc && (uint32_t)acc)
                atomicOr(&outw[wpos], (uint32_t)acc);
        }
    }
    __syncthreads();
    const uint32_t err = s_err;
    if (tid == 0 && !err)
        bw.resume(outw, out_words, p0 + total_bits);
    __syncthreads();
    return err;
}

struct ResidValues {
    const uint16_t *v;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return (uint32_t)v[i]; }
};

// ---- the LF stream of a frame by a thread-block CLUSTER -------------------------------------------------
// Up to 3 x 65 536 values go through one prefix-coded stream (one histogram, one header).  One CTA of 1024
// threads walked 192 positions per thread, every access a trip to L2 (the kernel runs with the L1 carve-out
// at maximum shared memory): 3.15 ms, longer than the 64 rANS chains it runs beside.  Here the stream
// belongs to a cluster of kLfCluster CTAs that are scheduled together: 8192 threads take 24 positions each,
// and what the single-CTA version kept in its shared memory -- scan carries, the histogram, the code
// table, error flags -- travels between the CTAs through distributed shared memory
// (cluster.map_shared_rank) around cluster-wide barriers.  CTA 0 is the master: it owns the bit sink,
// builds the code and writes the stream header.
namespace cg = cooperative_groups;
constexpr int kLfCluster = 8;

// exclusive scan over all threads of the cluster, in (rank, thread) order; `s_x`: one shared word per CTA
template <typename Op>
__device__ __forceinline__ uint32_t cluster_exclusive_scan(cg::cluster_group &cl, uint32_t v, uint32_t identity, Op op,
                                                           uint32_t *s_warp, uint32_t *s_x, uint32_t &total) {
    uint32_t btotal;
    const uint32_t excl = block_exclusive_scan(v, identity, op, s_warp, btotal);
    if (threadIdx.x == 0)
        *s_x = btotal;
    cl.sync();
    uint32_t prefix = identity, all = identity;
    const uint32_t r = cl.block_rank();
    for (uint32_t k = 0; k < (uint32_t)kLfCluster; k++) {
        const uint32_t t = *cl.map_shared_rank(s_x, k);
        if (k < r)
            prefix = op(prefix, t);
        all = op(all, t);
    }
    cl.sync();   // everyone has read the slots: they may be written again
    total = all;
    return op(prefix, excl);
}

// block_prefix_stream (above) for a cluster.  Every CTA passes its own PrefixWork / scratch; on return
// thread 0 of CTA 0 holds the bit sink behind the stream.  Returns error flags (uniform over the cluster).
template <typename ValueAt>
__device__ uint32_t cluster_prefix_stream(cg::cluster_group &cl, PrefixWork &w, BitSink &bw, uint32_t *outw, uint32_t out_words,
                                          uint32_t *syms, uint32_t sym_cap, const PrefixParams &prm, uint32_t total,
                                          ValueAt value_at, uint32_t *s_warp, uint32_t *s_first, uint32_t err_in,
                                          long long *stamps = nullptr) {
    __shared__ uint32_t s_bitpos, s_err, s_x;
    const uint32_t tid = threadIdx.x, rank = cl.block_rank();
    const uint32_t gid = rank * kLfThreads + tid, gthreads = kLfCluster * kLfThreads;
    const uint32_t lz = prm.lz_min_symbol;
    __syncthreads();
    for (uint32_t i = tid; i < (uint32_t)kAllBins; i += kLfThreads)
        w.freq[i] = 0;
    if (tid == 0) {
        w.alpha0 = w.alpha1 = 0;
        s_err = err_in;
    }
    cl.sync();   // nobody adds into CTA 0's histogram before it has been cleared
    uint32_t my_err = 0;
    // ---- symbols: thread = contiguous range of positions -------------------------------------------
    const uint32_t per = (total + gthreads - 1) / gthreads;
    const uint32_t a = gid * per < total ? gid * per : total, b = a + per < total ? a + per : total;
    auto starts_run = [&](uint32_t i) -> bool { return !lz || i == 0 || value_at(i) != value_at(i - 1); };
    uint32_t last_start = 0, first_start = total;
    for (uint32_t i = a; i < b; i++) {
        if (starts_run(i)) {
            last_start = i + 1;
            if (first_start == total)
                first_start = i;
        }
    }
    uint32_t dummy;
    const uint32_t carry_start = cluster_exclusive_scan(cl, last_start, 0u, [](uint32_t x, uint32_t y) { return x > y ? x : y; },
                                                        s_warp, &s_x, dummy);   // start (+1) of the run reaching into a
    // next run start at or after b: exclusive min-scan from the right (mirrored thread order inside the
    // CTA, then the minima of the CTAs to the right)
    uint32_t carry_end;
    {
        s_first[kLfThreads - 1 - tid] = first_start;
        __syncthreads();
        const uint32_t mirrored = s_first[tid];
        uint32_t bmin;
        const uint32_t ex = block_exclusive_scan(mirrored, total, [](uint32_t x, uint32_t y) { return x < y ? x : y; }, s_warp, bmin);
        s_first[tid] = ex;
        if (tid == 0)
            s_x = bmin;
        __syncthreads();
        carry_end = s_first[kLfThreads - 1 - tid];
        cl.sync();
        for (uint32_t k = rank + 1; k < (uint32_t)kLfCluster; k++) {
            const uint32_t t = *cl.map_shared_rank(&s_x, k);
            carry_end = t < carry_end ? t : carry_end;
        }
        cl.sync();
    }
    auto run_end = [&](uint32_t i) -> uint32_t {   // end of the run containing position i (i in [a, b))
        uint32_t e = i + 1;
        while (e < b && !starts_run(e))
            e++;
        return e < b ? e : carry_end;
    };
    const uint32_t sr0 = (a < b && !starts_run(a)) ? carry_start - 1 : a;
    // pass 1: count
    uint32_t count = 0;
    {
        uint32_t i = a, sr = sr0;
        while (i < b) {
            const uint32_t e = run_end(i);
            const uint32_t v = value_at(i);
            const uint32_t stop = e < b ? e : b;
            for (; i < stop; i++)
                count += lf_position_symbols(v, i, sr, e, prm).n;
            sr = i;
        }
    }
    uint32_t nsyms_total;
    uint32_t off = cluster_exclusive_scan(cl, count, 0u, [](uint32_t x, uint32_t y) { return x + y; }, s_warp, &s_x, nsyms_total);
    if (nsyms_total > sym_cap)
        my_err |= kErrLfCapacity;
    // pass 2: emit + histogram (into this CTA's own bins)
    if (nsyms_total <= sym_cap) {
        uint32_t i = a, sr = sr0;
        while (i < b) {
            const uint32_t e = run_end(i);
            const uint32_t v = value_at(i);
            const uint32_t stop = e < b ? e : b;
            for (; i < stop; i++) {
                const LfEmit em = lf_position_symbols(v, i, sr, e, prm);
                for (uint32_t k = 0; k < em.n; k++) {
                    const uint32_t sym = em.sym[k];
                    const uint32_t token = sym & 0x7FFFu, cluster = (sym >> 15) & 1u, nbits = (sym >> 16) & 0xFu;
                    bool fits;
                    if (cluster)
                        fits = token < (uint32_t)kDistBins;
                    else if (lz && token >= lz)
                        fits = token - lz < (uint32_t)kLzBins;
                    else
                        fits = token < (uint32_t)kLitBins;
                    if (!fits || nbits > 12) {
                        my_err |= kErrLfAlphabet;
                        syms[off++] = ps_pack(0, 0, 0, 0);
                        continue;
                    }
                    syms[off++] = sym;
                    atomicAdd(&w.freq[ps_bin(token, cluster, lz)], 1u);
                    if (cluster)
                        atomicMax(&w.alpha1, token + 1);
                    else
                        atomicMax(&w.alpha0, token + 1);
                }
            }
            sr = i;
        }
    }
    if (my_err)
        atomicOr(&s_err, my_err);
    __syncthreads();
    // ---- the CTAs' histograms, alphabet bounds and error flags meet in CTA 0 ---------------------------
    if (rank != 0) {
        PrefixWork *w0 = cl.map_shared_rank(&w, 0);
        for (uint32_t i = tid; i < (uint32_t)kAllBins; i += kLfThreads)
            if (w.freq[i])
                atomicAdd(&w0->freq[i], w.freq[i]);
        if (tid == 0) {
            atomicMax(&w0->alpha0, w.alpha0);
            atomicMax(&w0->alpha1, w.alpha1);
            if (s_err)
                atomicOr(cl.map_shared_rank(&s_err, 0), s_err);
        }
    }
    cl.sync();
    if (stamps)
        stamps[0] = clock64();   // symbols emitted, histograms merged
    // ---- CTA 0: code lengths of the literal / length cluster by its first warp, then the stream header ---
    if (rank == 0) {
        const bool warp_lengths = w.alpha0 > 1;
        if (tid < 32 && warp_lengths)
            warp_code_lengths(w, w.alpha0, 15, lz, tid);
        __syncthreads();
        if (tid == 0) {
            w.nsyms = nsyms_total;
            ps_put_header(w, bw, prm, warp_lengths);
            bw.flush_partial();
            s_bitpos = bw.bitlen();
            if (bw.overflow)
                s_err |= kErrSlab;
            if (w.error)
                s_err |= w.error;
        }
    }
    cl.sync();
    if (stamps)
        stamps[1] = clock64();   // code built, stream header written
    // ---- everyone takes the code table, the header's end and the verdict from CTA 0 ------------------------
    if (rank != 0) {
        const PrefixWork *w0 = cl.map_shared_rank(&w, 0);
        for (uint32_t i = tid; i < (uint32_t)kAllBins; i += kLfThreads) {
            w.len[i] = w0->len[i];
            w.code[i] = w0->code[i];
        }
        if (tid == 0) {
            s_bitpos = *cl.map_shared_rank(&s_bitpos, 0);
            s_err = *cl.map_shared_rank(&s_err, 0);
        }
    }
    cl.sync();   // (also: CTA 0's table is not touched again before everyone has copied it)
    const uint32_t p0 = s_bitpos;
    const uint32_t err_now = s_err;
    // ---- symbol bits, in parallel: thread = contiguous range of symbols -----------------------------
    uint32_t total_bits = 0;
    uint32_t late_err = 0;
    if (!err_now) {
        const uint32_t sper = (nsyms_total + gthreads - 1) / gthreads;
        const uint32_t sa = gid * sper < nsyms_total ? gid * sper : nsyms_total;
        const uint32_t sb = sa + sper < nsyms_total ? sa + sper : nsyms_total;
        uint32_t bits = 0;
        for (uint32_t i = sa; i < sb; i++) {
            const uint32_t sym = syms[i];
            bits += w.len[ps_bin(sym & 0x7FFFu, (sym >> 15) & 1u, lz)] + ((sym >> 16) & 0xFu);
        }
        const uint32_t boff = cluster_exclusive_scan(cl, bits, 0u, [](uint32_t x, uint32_t y) { return x + y; }, s_warp, &s_x, total_bits);
        const uint64_t endbit = (uint64_t)p0 + total_bits;
        if (endbit + 64 > (uint64_t)out_words * 32) {
            late_err = kErrSlab;   // uniform: every thread sees the same totals
        } else {
            // flush_partial() wrote word p0 >> 5 only when the header left a partial word there
            for (uint32_t wd = ((p0 + 31) >> 5) + gid; wd <= (uint32_t)(endbit >> 5) + 1; wd += gthreads)
                outw[wd] = 0;
            cl.sync();   // the words are clear before any CTA ORs bits into them
            uint64_t pos = (uint64_t)p0 + boff;
            uint32_t wpos = (uint32_t)(pos >> 5), nacc = (uint32_t)(pos & 31u);
            uint64_t acc = 0;
            auto put = [&](uint32_t v, uint32_t n) {
                acc |= (uint64_t)v << nacc;
                nacc += n;
                if (nacc >= 32) {
                    atomicOr(&outw[wpos], (uint32_t)acc);
                    wpos++;
                    acc >>= 32;
                    nacc -= 32;
                }
            };
            for (uint32_t i = sa; i < sb; i++) {
                const uint32_t sym = syms[i];
                const uint32_t bin = ps_bin(sym & 0x7FFFu, (sym >> 15) & 1u, lz);
                const uint32_t nbits = (sym >> 16) & 0xFu;
                if (w.len[bin])
                    put(w.code[bin], w.len[bin]);
                if (nbits)
                    put(sym >> 20, nbits);
            }
            if (nacc && (uint32_t)acc)
                atomicOr(&outw[wpos], (uint32_t)acc);
        }
    }
    cl.sync();   // all bits are in place before CTA 0's thread 0 continues behind them
    const uint32_t err = err_now | late_err;
    if (rank == 0 && tid == 0 && !err)
        bw.resume(outw, out_words, p0 + total_bits);
    return err;
}

__global__ void __cluster_dims__(kLfCluster, 1, 1) __launch_bounds__(kLfThreads)
k_frame_lf(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    __shared__ uint32_t s_warp[64];
    __shared__ uint32_t s_first[kLfThreads];
    __shared__ uint32_t s_resid_err;
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t slot = blockIdx.x / kLfCluster, tid = threadIdx.x, rank = cl.block_rank();
    const uint32_t gid = rank * kLfThreads + tid, gthreads = kLfCluster * kLfThreads;
    const TileDesc t = ws.tiles[slot];
    if (!(t.flags & kTilePrefix))
        return;   // (the whole cluster leaves: its CTAs share the slot)
    const long long clk0 = clock64();
    long long stamps[2] = {0, 0};
    PrefixWork &w = s.work;
    const uint32_t vbw = (t.frame_w + 7) >> 3, vbh = (t.frame_h + 7) >> 3, nb = vbw * vbh, total = 3 * nb;
    uint32_t *syms = ws.syms + (size_t)slot * kMaxHfSyms;
    uint8_t *slab = ws.slab + (size_t)slot * kSlabBytes;
    uint32_t *outw = reinterpret_cast<uint32_t *>(slab + kPrefixLfOffset);
    constexpr uint32_t kOutWords = (kSlabBytes - kPrefixLfOffset - kPrefixTailReserve) / 4;
    if (tid == 0)
        s_resid_err = 0;
    __syncthreads();
    // ---- residuals of the whole LF image, channel order Y, X, B (encoder.c:574-592) ---------------
    uint16_t *resid = reinterpret_cast<uint16_t *>(ws.coef + (size_t)slot * kMaxBlocks * 3 * 64);   // 196 608 x u16
    for (uint32_t i = gid; i < total; i += gthreads) {
        const uint32_t ci = i / nb, r = i - ci * nb;
        const uint32_t c = ci < 2 ? 1 - ci : ci;
        const uint32_t by = r / vbw, bx = r - by * vbw;
        const int32_t v = frame_lf_at(ws, slot, t.frame_gx, c, bx, by);
        const int32_t up = by ? frame_lf_at(ws, slot, t.frame_gx, c, bx, by - 1) : 0;
        const int32_t wv = bx ? frame_lf_at(ws, slot, t.frame_gx, c, bx - 1, by) : up;
        const int32_t n = by ? up : wv;
        const int32_t nw = (bx && by) ? frame_lf_at(ws, slot, t.frame_gx, c, bx - 1, by - 1) : wv;
        const int32_t lo = wv < n ? wv : n, hi = wv < n ? n : wv;
        int32_t pred = wv + n - nw;
        pred = pred < lo ? lo : (pred > hi ? hi : pred);
        const uint32_t packed = pack_signed(v - pred);
        if (packed > 0xFFFFu)
            s_resid_err = kErrLfAlphabet;   // would need more than 12 residue bits: outside what the LF coder holds
        resid[i] = (uint16_t)(packed > 0xFFFFu ? 0xFFFFu : packed);
    }
    // ---- section head (one thread of the master CTA; uses the prefix work area for the small MA-tree stream)
    BitSink bw;
    if (rank == 0 && tid == 0) {
        w.error = 0;
        bw.init(outw, kOutWords);
        put_lf_group_head(w, syms, bw);
    }
    __syncthreads();
    const uint32_t resid_err = s_resid_err;
    cl.sync();   // every CTA's residuals are in HBM before any CTA reads its range of them
    const long long clk1 = clock64();
    // ---- the LF stream ---------------------------------------------------------------------------------
    uint32_t err = cluster_prefix_stream(cl, w, bw, outw, kOutWords, syms, (uint32_t)kMaxHfSyms, lf_stream_params(), total,
                                         ResidValues{resid}, s_warp, s_first, resid_err, ws.dbg_clk ? stamps : nullptr);
    if (rank != 0 || tid != 0)
        return;
    const long long clk2 = clock64();
    // ---- the constant HF-metadata image behind it, then the section is closed ------------------------
    err |= w.error;
    if (!err) {
        put_hf_metadata(w, syms, (uint32_t)kMaxHfSyms, bw, vbw, vbh);
        bw.align_byte();
        bw.flush_partial();
        ws.lfbitlen[slot] = bw.bitlen() >> 3;   // bytes of the LFGroup section
        err |= w.error;
        if (bw.overflow)
            err |= kErrSlab;
    } else {
        ws.lfbitlen[slot] = 0;
    }
    if (err)
        atomicOr(&ws.tile_err[slot], err);
    if (ws.dbg_clk) {   // stage-tap builds: cycles of residuals + section head, symbols, code + header, (bits, HF metadata)
        ws.dbg_clk[slot * 4 + 0] = (uint32_t)(clk1 - clk0);
        ws.dbg_clk[slot * 4 + 1] = (uint32_t)(stamps[0] - clk1);
        ws.dbg_clk[slot * 4 + 2] = (uint32_t)(stamps[1] - stamps[0]);
        ws.dbg_clk[slot * 4 + 3] = (uint32_t)(clk2 - stamps[1]) | ((uint32_t)((clock64() - clk2) >> 8) << 20);
    }
}

struct ClusterMapMtf {
    const uint16_t *idx;
    HD uint32_t operator()(uint32_t i) const { return (uint32_t)idx[i]; }
};
struct WordValues {
    const uint32_t *v;
    HD uint32_t operator()(uint32_t i) const { return v[i]; }
};

// Move-to-front index of entry j of the HF context map (1485 contexts per preset, K clusters per
// preset numbered K * preset + local; reference: entropy.c:136-151 evolves the list entry by entry).
// Closed form: a cluster's previous occurrence is at most six entries back inside its preset, and the
// index is the number of distinct clusters in between; at its first occurrence every cluster seen so
// far is smaller, so it still sits at its initial place, index = its own number.
__device__ __forceinline__ uint32_t hf_map_mtf_index(uint32_t j, uint32_t K) {
    const uint32_t preset = j / 1485u, ctx = j - preset * 1485u;
    const uint32_t c = hf_fold_cluster(hf_context_cluster(ctx), K);
    uint32_t between = 0;
    for (uint32_t d = 1; d <= 8 && d <= ctx; d++) {
        const uint32_t cc = hf_fold_cluster(hf_context_cluster(ctx - d), K);
        if (cc == c)
            return (uint32_t)__popc(between);
        between |= 1u << cc;
    }
    return K * preset + c;
}

__global__ void __launch_bounds__(kLfThreads)
k_frame_finish(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    __shared__ uint32_t s_warp[64];
    __shared__ uint32_t s_first[kLfThreads];
    const uint32_t slot = blockIdx.x, tid = threadIdx.x;
    const TileDesc t = ws.tiles[slot];
    if (!(t.flags & kTilePrefix))
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
        if (tid == 0) {
            ws.frame_off[slot] = kPrefixLfOffset;
            ws.frame_len[slot] = err ? 0u : len1;
            if (err)
                atomicOr(&ws.tile_err[slot], err);
        }
        return;
    }
    // ---- HFGlobal section: constants + the ANS header tail the first group's chain kernel wrote ---
    // (sections.cuh::put_hf_global, with the context map's nested stream coded by the whole CTA)
    BitSink b2;
    if (tid == 0) {
        b2.init(s.s2, kDBitsWords + 128);
        s.work.error = 0;
        b2.put_bool(1);    // HFGlobal: default dequant matrices
        b2.put(0, ceil_log2_u32(G));   // num_presets - 1 = 0
        b2.put(2, 2);      // HF pass order
        b2.put_bool(0);    // ANS stream: no lz77
        b2.put_bool(0);    // context map: not simple
        b2.put_bool(1);    // move-to-front
    }
    for (uint32_t j = tid; j < (uint32_t)kHfContexts; j += kLfThreads)
        s.mapidx[j] = (uint16_t)hf_map_mtf_index(j, 9);
    {
        PrefixParams p;
        p.num_plain_dists = 1;
        p.lz_min_symbol = 64;
        p.modular = 0;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 0;
        p.split1 = 4; p.msb1 = 1; p.lsb1 = 0;
        err |= block_prefix_stream(s.work, b2, s.s2, kDBitsWords + 128, syms, (uint32_t)kSectionSymCap, p,
                                   (uint32_t)kHfContexts, ClusterMapMtf{s.mapidx}, s_warp, s_first);
    }
    if (tid != 0)
        return;
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

__global__ void __launch_bounds__(kLfThreads)
k_oneframe_finish(const uint32_t *__restrict__ info, uint32_t info_words, uint32_t *scratch, uint32_t scratch_words,
                  uint8_t *out, uint32_t head_cap, uint32_t hf_cap, uint32_t *ctx_cache, uint32_t ctx_cache_words,
                  uint32_t ctx_cached_bits, uint32_t *perm_cache, uint32_t perm_cache_words, uint32_t perm_cached_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    __shared__ uint32_t s_warp[64];
    __shared__ uint32_t s_first[kLfThreads];
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
    const long long tap0 = clock64();
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
    for (uint32_t i = tid; i < toc_size; i += kLfThreads) {
        uint32_t smaller_before = 0;
        const uint32_t v = inv[i];
        for (uint32_t j = 0; j < i; j++)
            smaller_before += inv[j] < v ? 1u : 0u;
        leh[1 + i] = v - smaller_before;
    }
    __syncthreads();
    uint32_t err = 0;
    const long long tap1 = clock64();
    // ---- HFGlobal section ------------------------------------------------------------------------------
    uint8_t *hf_bytes = out + head_cap;
    BitSink b2;
    if (tid == 0) {
        s.work.error = 0;
        b2.init(reinterpret_cast<uint32_t *>(hf_bytes), hf_cap / 4);
        b2.put_bool(1);                                       // default dequant matrices
        b2.put(n - 1, ceil_log2_u32(G));                      // num_presets - 1 (encoder.c:961)
        b2.put(2, 2);                                         // HF pass order
        b2.put_bool(0);                                       // ANS stream: no lz77
        // context map of 1485 n contexts onto K n clusters: never "simple" for n >= 2 (entropy.c:108-167)
        b2.put_bool(0);
        b2.put_bool(1);                                       // move-to-front
    }
    // The context map's stream depends on nothing but the number of presets: the engine keeps its bits from the
    // first image of a geometry (ctx_cache) and later images splice them in (ctx_cached_bits != 0) instead of
    // coding 1485 n entries again -- half of this kernel's time, on the tail of every one-frame image.
    __shared__ uint32_t s_ctx_span[2];
    if (ctx_cached_bits) {
        if (tid == 0)
            put_bits_from(b2, ctx_cache, ctx_cached_bits);
    } else {
        if (tid == 0)
            s_ctx_span[0] = b2.bitlen();
        uint16_t *idx = reinterpret_cast<uint16_t *>(tokens + 1485u * n / 2 + 2048u);   // upper part of the token scratch
        for (uint32_t j = tid; j < 1485u * n; j += kLfThreads)
            idx[j] = (uint16_t)hf_map_mtf_index(j, K);
        PrefixParams p;
        p.num_plain_dists = 1;
        p.lz_min_symbol = 64;
        p.modular = 0;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 0;
        p.split1 = 4; p.msb1 = 1; p.lsb1 = 0;
        err |= block_prefix_stream(s.work, b2, reinterpret_cast<uint32_t *>(hf_bytes), hf_cap / 4, tokens,
                                   1485u * n / 2 + 2048u, p, 1485u * n, ClusterMapMtf{idx}, s_warp, s_first);
        if (tid == 0)
            s_ctx_span[1] = b2.bitlen();
        __syncthreads();
        // keep the stream's bits, shifted down to bit 0, for the next image of this geometry
        const uint32_t from = s_ctx_span[0], bits = s_ctx_span[1] - s_ctx_span[0];
        const uint32_t words = (bits + 31) >> 5, sh = from & 31u;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(hf_bytes) + (from >> 5);
        if (!err && words <= ctx_cache_words) {
            for (uint32_t i = tid; i < words; i += kLfThreads) {
                const uint32_t lo = src[i] >> sh, hi = sh ? src[i + 1] << (32u - sh) : 0u;   // src[i + 1]: at most the partial last word + 1, inside hf_cap
                ctx_cache[i] = lo | hi;
            }
            if (tid == 0)
                res[8] = bits;
        } else if (tid == 0) {
            res[8] = 0;
        }
    }
    const long long tap2 = clock64();
    // ---- head, part 1: [image header] frame header + TOC permutation (all threads) ----------------------
    uint32_t pre_bytes = 0;
    BitSink bh;
    if (tid == 0) {
        if (with_header && image_needs_level10(W, H))
            pre_bytes = put_level10_prefix(out);
        bh.init(reinterpret_cast<uint32_t *>(out + 52), (head_cap - 52) / 4);   // 52: word aligned, past the 49-byte prefix
        if (with_header)
            put_image_header(bh, W, H);
        put_frame_header_fields(bh, false, 0, 0, W, H, true);
        bh.put_bool(1);                                       // permuted TOC
        s.work.error = 0;
    }
    // the TOC permutation's stream depends on the geometry and the order the LF groups were sent in: the engine
    // keeps its bits too (perm_cache; the host compares the order) -- an image sent like the one before splices them
    if (perm_cached_bits) {
        if (tid == 0)
            put_bits_from(bh, perm_cache, perm_cached_bits);
    } else {
        if (tid == 0)
            s_ctx_span[0] = bh.bitlen();
        PrefixParams p;
        p.num_plain_dists = 8;
        p.lz_min_symbol = 0;
        p.modular = 0;
        p.split0 = 4; p.msb0 = 1; p.lsb0 = 1;
        p.split1 = 7; p.msb1 = 0; p.lsb1 = 0;
        err |= block_prefix_stream(s.work, bh, reinterpret_cast<uint32_t *>(out + 52), (head_cap - 52) / 4, tokens,
                                   scratch_words - fixed, p, 1 + toc_size, WordValues{leh}, s_warp, s_first);
        if (tid == 0)
            s_ctx_span[1] = bh.bitlen();
        __syncthreads();
        const uint32_t from = s_ctx_span[0], bits = s_ctx_span[1] - s_ctx_span[0];
        const uint32_t words = (bits + 31) >> 5, sh = from & 31u;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(out + 52) + (from >> 5);
        if (!err && words <= perm_cache_words) {
            for (uint32_t i = tid; i < words; i += kLfThreads) {
                const uint32_t lo = src[i] >> sh, hi = sh ? src[i + 1] << (32u - sh) : 0u;
                perm_cache[i] = lo | hi;
            }
            if (tid == 0)
                res[9] = bits;
        } else if (tid == 0) {
            res[9] = 0;
        }
    }
    if (tid != 0)
        return;
    const long long tap3 = clock64();
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
    // ---- head, part 2: TOC, LFGlobal ---------------------------------------------------------------------
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
    // the container prefix (49 bytes, if any) sits at byte 0, the header bits at byte 52: the host closes the gap
    // when it copies the head out (a single thread moving ~1 KB byte by byte through global memory was most of
    // this kernel's 0.69 ms)
    res[0] = pre_bytes + hb;
    res[1] = len2;
    res[2] = err;
    res[3] = pre_bytes;
    // phase tap (cycles): TOC permutation + Lehmer code, context-map stream, permutation stream, single-thread tail
    res[4] = (uint32_t)(tap1 - tap0);
    res[5] = (uint32_t)(tap2 - tap1);
    res[6] = (uint32_t)(tap3 - tap2);
    res[7] = (uint32_t)(clock64() - tap3);
}

// ---- image header of an ICC-tagged image (hyd_set_suggested_icc_profile) ----------------------------
// reference: encoder.c:122-162 (context model), 203-236 (header fields + the profile as a prefix-coded
// stream over 41 contexts in 9 clusters, default hybrid config 4/1/1, no lz77).  `icc` is the profile
// as hyd_set_suggested_icc_profile mangles it (libhydrium.c:242-305; host code, hyd_api.c).
// One CTA: the histograms and the symbol bits are done by all threads, the code construction by one.
__device__ __forceinline__ uint32_t icc_context_of(uint32_t i, uint32_t b1, uint32_t b2) {
    if (i <= 128)
        return 0;
    auto alpha = [](uint32_t b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z'); };
    auto digit = [](uint32_t b) { return (b >= '0' && b <= '9') || b == '.' || b == ','; };
    uint32_t p1, p2;
    if (alpha(b1)) p1 = 0;
    else if (digit(b1)) p1 = 1;
    else if (b1 <= 1) p1 = b1 + 2;
    else if (b1 < 16) p1 = 4;
    else if (b1 > 240 && b1 < 255) p1 = 5;
    else if (b1 == 255) p1 = 6;
    else p1 = 7;
    if (alpha(b2)) p2 = 0;
    else if (digit(b2)) p2 = 1;
    else if (b2 < 16) p2 = 2;
    else if (b2 > 240) p2 = 3;
    else p2 = 4;
    return 1 + p1 + p2 * 8;
}
// contexts 1 + p1 + 8 * p2 share cluster 1 + p1; context 0 (the 128-byte header) is cluster 0
__device__ __forceinline__ uint32_t icc_cluster_of(uint32_t ctx) { return ctx ? 1u + ((ctx - 1u) & 7u) : 0u; }
constexpr uint32_t kIccClusters = 9, kIccContexts = 41, kIccBins = 32;   // byte tokens under 4/1/1 are < 32

__device__ __forceinline__ uint32_t icc_symbol(const uint8_t *__restrict__ icc, uint32_t i, uint32_t &cluster,
                                               uint32_t &res, uint32_t &nb) {
    const uint32_t b1 = i >= 1 ? icc[i - 1] : 0u, b2 = i >= 2 ? icc[i - 2] : 0u;
    cluster = icc_cluster_of(icc_context_of(i, b1, b2));
    return hybrid_token(icc[i], 4, 1, 1, res, nb);
}

__global__ void __launch_bounds__(kLfThreads)
k_icc_header(const uint8_t *__restrict__ icc, uint32_t n, uint32_t W, uint32_t H, uint32_t *bits, uint32_t bits_words,
             uint8_t *out, uint32_t out_cap, uint32_t *res_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared &s = *reinterpret_cast<FrameShared *>(smem_raw);
    __shared__ uint32_t s_warp[64];
    __shared__ uint32_t s_alpha[kIccClusters];
    __shared__ uint32_t s_bitpos, s_err;
    __shared__ uint16_t s_idx[kIccContexts];
    PrefixWork &w = s.work;
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < (uint32_t)kAllBins; i += kLfThreads)
        w.freq[i] = 0, w.len[i] = 0, w.code[i] = 0;
    if (tid < kIccClusters)
        s_alpha[tid] = 0;
    BitSink bw;
    if (tid == 0) {
        s_err = 0;
        w.error = 0;
        bw.init(bits, bits_words);
        put_image_header_icc_fields(bw, W, H, n);
        bw.put_bool(0);    // no lz77
        bw.put_bool(0);    // cluster map: 9 clusters need 4 bits per entry, so not the simple form ...
        bw.put_bool(1);    // ... but move-to-front + a nested prefix stream (entropy.c:125-158)
        uint8_t mtf[kIccClusters];
        for (uint32_t k = 0; k < kIccClusters; k++)
            mtf[k] = (uint8_t)k;
        for (uint32_t j = 0; j < kIccContexts; j++) {
            const uint8_t c = (uint8_t)icc_cluster_of(j);
            int k = 0;
            while (mtf[k] != c)
                k++;
            s_idx[j] = (uint16_t)k;
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
        ps_encode_stream(w, s.s2, 128, p, kIccContexts, ClusterMapMtf{s_idx}, bw);
        if (w.error)
            s_err |= kErrSlab;
        for (int i = 0; i < kAllBins; i++)
            w.freq[i] = 0, w.len[i] = 0, w.code[i] = 0;
    }
    __syncthreads();
    // ---- histograms: bin = cluster * 32 + token ------------------------------------------------------
    for (uint32_t i = tid; i < n; i += kLfThreads) {
        uint32_t c, r, nb;
        const uint32_t tok = icc_symbol(icc, i, c, r, nb);
        atomicAdd(&w.freq[c * kIccBins + tok], 1u);
        atomicMax(&s_alpha[c], tok + 1);
    }
    __syncthreads();
    // ---- stream header: prefix codes, hybrid configs, alphabet sizes, the nine codes ----------------
    if (tid == 0) {
        bw.put_bool(1);   // use prefix codes
        for (uint32_t c = 0; c < kIccClusters; c++)
            ps_put_hybrid_cfg(bw, 4, 1, 1, 15);
        for (uint32_t c = 0; c < kIccClusters; c++) {   // entropy.c:835-844
            if (s_alpha[c] <= 1) {
                bw.put_bool(0);
                continue;
            }
            bw.put_bool(1);
            const int nb = floor_log2_u32(s_alpha[c] - 1);
            bw.put((uint32_t)nb, 4);
            bw.put(s_alpha[c] - 1, nb);
        }
        for (uint32_t c = 0; c < kIccClusters; c++)
            if (s_alpha[c] > 1)
                ps_put_cluster_code(w, bw, c * kIccBins, kIccBins, s_alpha[c], 0, true);
        bw.flush_partial();
        s_bitpos = bw.bitlen();
        if (bw.overflow || w.error)
            s_err |= kErrSlab;
    }
    __syncthreads();
    const uint32_t p0 = s_bitpos;
    // ---- symbol bits: thread = contiguous range of profile bytes --------------------------------------
    uint32_t total_bits = 0;
    if (!s_err) {
        const uint32_t per = (n + kLfThreads - 1) / kLfThreads;
        const uint32_t a = tid * per < n ? tid * per : n, b = a + per < n ? a + per : n;
        uint32_t cnt = 0;
        for (uint32_t i = a; i < b; i++) {
            uint32_t c, r, nb;
            const uint32_t tok = icc_symbol(icc, i, c, r, nb);
            cnt += w.len[c * kIccBins + tok] + nb;
        }
        const uint32_t boff = block_exclusive_scan(cnt, 0u, [](uint32_t x, uint32_t y) { return x + y; }, s_warp, total_bits);
        const uint64_t endbit = (uint64_t)p0 + total_bits;
        if (endbit + 64 > (uint64_t)bits_words * 32) {
            if (tid == 0)
                s_err |= kErrSlab;
        } else {
            for (uint32_t wd = ((p0 + 31) >> 5) + tid; wd <= (uint32_t)(endbit >> 5) + 1; wd += kLfThreads)
                bits[wd] = 0;
            __syncthreads();
            uint64_t pos = (uint64_t)p0 + boff;
            uint32_t wpos = (uint32_t)(pos >> 5), nacc = (uint32_t)(pos & 31u);
            uint64_t acc = 0;
            auto put = [&](uint32_t v, uint32_t nbit) {
                acc |= (uint64_t)v << nacc;
                nacc += nbit;
                if (nacc >= 32) {
                    atomicOr(&bits[wpos], (uint32_t)acc);
                    wpos++;
                    acc >>= 32;
                    nacc -= 32;
                }
            };
            for (uint32_t i = a; i < b; i++) {
                uint32_t c, r, nb;
                const uint32_t tok = icc_symbol(icc, i, c, r, nb);
                const uint32_t bi = c * kIccBins + tok;
                if (w.len[bi])
                    put(w.code[bi], w.len[bi]);
                if (nb)
                    put(r, nb);
            }
            if (nacc && (uint32_t)acc)
                atomicOr(&bits[wpos], (uint32_t)acc);
        }
    }
    __syncthreads();
    if (tid == 0 && !s_err) {
        bw.resume(bits, bits_words, p0 + total_bits);
        bw.align_byte();
        bw.flush_partial();
        s_bitpos = bw.bitlen() >> 3;
        if (bw.overflow)
            s_err |= kErrSlab;
    }
    __syncthreads();
    const uint32_t nbytes = s_bitpos;
    uint32_t pre = 0;
    if (image_needs_level10(W, H))
        pre = 49;
    if (!s_err && pre + nbytes <= out_cap) {
        if (pre && tid == 0)
            put_level10_prefix(out);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(bits);
        for (uint32_t i = tid; i < nbytes; i += kLfThreads)
            out[pre + i] = src[i];
    } else if (tid == 0) {
        s_err |= kErrSlab;
    }
    __syncthreads();
    if (tid == 0) {
        res_out[0] = pre + nbytes;
        res_out[1] = s_err;
    }
}

void launch_icc_header(const uint8_t *d_icc, uint32_t n, uint32_t W, uint32_t H, uint32_t *d_bits, uint32_t bits_words,
                       uint8_t *d_out, uint32_t out_cap, uint32_t *d_res, cudaStream_t st) {
    cudaFuncSetAttribute(k_icc_header, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_icc_header<<<1, kLfThreads, sizeof(FrameShared), st>>>(d_icc, n, W, H, d_bits, bits_words, d_out, out_cap, d_res);
}

void launch_oneframe_finish(const uint32_t *d_info, uint32_t info_words, uint32_t *d_scratch, uint32_t scratch_words,
                            uint8_t *d_out, uint32_t head_cap, uint32_t hf_cap, uint32_t *d_ctx_cache, uint32_t ctx_cache_words,
                            uint32_t ctx_cached_bits, uint32_t *d_perm_cache, uint32_t perm_cache_words, uint32_t perm_cached_bits,
                            cudaStream_t st) {
    cudaFuncSetAttribute(k_oneframe_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_oneframe_finish<<<1, kLfThreads, sizeof(FrameShared), st>>>(d_info, info_words, d_scratch, scratch_words, d_out, head_cap, hf_cap,
                                                                  d_ctx_cache, ctx_cache_words, ctx_cached_bits, d_perm_cache,
                                                                  perm_cache_words, perm_cached_bits);
}

void launch_frame_hist_sum(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_hist_sum);
    k_frame_hist_sum<<<nslots, kHfClusters * kHfTokens, 0, st>>>(ws);
}

void launch_frame_lf(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_lf);
    cudaFuncSetAttribute(k_frame_lf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_frame_lf<<<nslots * kLfCluster, kLfThreads, sizeof(FrameShared), st>>>(ws);   // one cluster per slot
}

void launch_frame_finish(const Workspace &ws, uint32_t nslots, cudaStream_t st) {
    prefer_max_shared(k_frame_finish);
    cudaFuncSetAttribute(k_frame_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameShared));
    k_frame_finish<<<nslots, kLfThreads, sizeof(FrameShared), st>>>(ws);
}

}  // namespace hydb
