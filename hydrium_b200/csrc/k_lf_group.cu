// hydrium_b200/csrc/k_lf_group.cu
//
// Stage 3: the LF coefficient stream of each tile (section L) and the shape-constant sections
// A and B (see sections.cuh).  Replaces write_lf_group's modular sub-image coding
// (reference: encoder.c:539-629) including the whole prefix-code back end
// (entropy.c:502-524, 546-941, 1003-1034).
//
// One warp per tile, every phase warp-parallel except the few hundred header fields:
//   1. residuals        lane = position; clamped-gradient prediction, folded sign
//   2. run structure    ballots of "value differs from predecessor" per 32 positions; a backward
//                       sweep gives every position the end of its run, a forward sweep the start,
//                       so the reference's sequential run-length state machine (one literal + up
//                       to 127 repeats per chunk, entropy.c:473-524) becomes a per-position rule
//   3. histogram        shared-memory atomics on the sparse bins
//   4. code lengths     the reference's slot-swapping tree builder (entropy.c:592-662) on live
//                       nodes only, the two minima found by warp reductions on a packed 64-bit key
//   5. header           lane 0, sequential (prefix_coder.cuh), written straight to HBM
//   6. symbol bits      per-position bit counts -> warp scans -> atomicOr into the bit string
// ~26 KB of shared memory per tile, so the kernel co-resides with the rANS chain CTAs.
#include "kernels.h"
#include "lf_values.cuh"
#include "sections.cuh"
#include "prefix_warp.cuh"

namespace hydb {

constexpr uint32_t FULL = 0xFFFFFFFFu;

struct LfShared {
    PrefixWork work;
    uint16_t vals[3 * kMaxBlocks];
    uint16_t next_start[3 * kMaxBlocks];
    uint32_t starts[3 * kMaxBlocks / 32];
    uint32_t lencount[16], firstcode[16];
};

// canonical codes (ps_assign_codes) with lane l handling code length l
__device__ void warp_assign_codes(LfShared &s, uint32_t lane) {
    PrefixWork &w = s.work;
    if (lane < 16)
        s.lencount[lane] = 0;
    __syncwarp();
    for (uint32_t b = lane; b < (uint32_t)kBins; b += 32)
        if (w.len[b])
            atomicAdd(&s.lencount[w.len[b]], 1u);
    __syncwarp();
    if (lane == 0) {
        uint32_t next = 0;   // left-aligned in 16 bits
        for (int l = 1; l <= 15; l++) {
            s.firstcode[l] = next;
            next += s.lencount[l] << (16 - l);
        }
        if (next && next != (1u << 16))
            w.error |= kErrHuffman;   // reference: "VLC codes do not add up"
    }
    __syncwarp();
    if (lane >= 1 && lane <= 15 && s.lencount[lane]) {
        uint32_t next = s.firstcode[lane];
        for (uint32_t b = 0; b < (uint32_t)kBins; b++)
            if (w.len[b] == lane) {
                w.code[b] = (uint16_t)(__brev(next >> (16 - lane)) >> (32 - lane));
                next += 1u << (16 - lane);
            }
    }
    __syncwarp();
}

// what the chunk starting at position i emits: literal count, optional run-length pair
struct ChunkPlan {
    uint32_t v, lit_tok, lit_nbits, lit_res;
    uint32_t lits;      // 1 + repeats coded as literals
    uint32_t rep_tok;   // 0, or lz_min + (rep - 3)
};

__device__ __forceinline__ bool chunk_plan(const LfShared &s, uint32_t i, uint32_t run_start, uint32_t lz_min, ChunkPlan &c) {
    if (((i - run_start) & 127u) != 0)
        return false;
    const uint32_t remaining = (uint32_t)s.next_start[i] - i;
    const uint32_t rep = (remaining < 128u ? remaining : 128u) - 1u;
    c.v = s.vals[i];
    c.lit_tok = hybrid_token(c.v, 7, 1, 1, c.lit_res, c.lit_nbits);
    c.lits = 1 + (rep <= 3 ? rep : 0);
    c.rep_tok = rep > 3 ? lz_min + (rep - 3) : 0;
    return true;
}

__global__ void __launch_bounds__(32)
k_lf_group(const TileDesc *__restrict__ tiles, const int32_t *__restrict__ lfq, uint32_t *__restrict__ lfbits,
           uint32_t *__restrict__ lfbitlen, uint32_t *__restrict__ tile_err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LfShared &s = *reinterpret_cast<LfShared *>(smem_raw);
    PrefixWork &w = s.work;
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    const TileDesc t = tiles[tile];
    if (t.flags & (kTilePrefix | kTileMulti))   // multi-group frames code ONE LF image per frame (k_frame.cu)
        return;
    const uint32_t vbw = (t.w + 7) >> 3, vbh = (t.h + 7) >> 3, nb = vbw * vbh, n = 3 * nb;
    const uint32_t rounds = (n + 31) >> 5;
    const PrefixParams prm = lf_stream_params();
    const uint32_t lz_min = prm.lz_min_symbol;
    uint32_t *out = lfbits + (size_t)tile * kLfBitsWords;
    uint32_t err = 0;

    // ---- 1. residuals ---------------------------------------------------------------------------
    {
        const LfStreamValues values{lfq + (size_t)tile * 3 * kMaxBlocks, vbw, nb};
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t v = values(i);
            if (v > 0xFFFFu)
                err |= kErrLfAlphabet;
            s.vals[i] = (uint16_t)v;
        }
        for (uint32_t b = lane; b < (uint32_t)kAllBins; b += 32) {
            w.freq[b] = 0;
            w.len[b] = 0;    // stays zero when a cluster needs no code (alphabet <= 1)
            w.code[b] = 0;
        }
        if (lane == 0) {
            w.error = 0;
            w.alpha0 = w.alpha1 = 0;
        }
    }
    __syncwarp();
    // ---- 2. run structure -----------------------------------------------------------------------
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t i = r * 32 + lane;
        const bool start = i < n && (i == 0 || s.vals[i] != s.vals[i - 1]);
        const uint32_t m = __ballot_sync(FULL, start);
        if (lane == 0)
            s.starts[r] = m;
    }
    __syncwarp();
    {
        uint32_t carry = n;   // first run start after the current round
        for (int r = (int)rounds - 1; r >= 0; r--) {
            const uint32_t i = (uint32_t)r * 32 + lane, m = s.starts[r];
            const uint32_t above = lane == 31 ? 0u : (m & ~((2u << lane) - 1u));
            if (i < n)
                s.next_start[i] = (uint16_t)(above ? (uint32_t)r * 32 + __ffs(above) - 1 : carry);
            if (m)
                carry = (uint32_t)r * 32 + __ffs(m) - 1;
        }
    }
    __syncwarp();
    // ---- 3. histogram ---------------------------------------------------------------------------
    uint32_t max_tok = 0, any_rep = 0;
    {
        uint32_t carry = 0;   // start of the run that is open at the beginning of the round
        for (uint32_t r = 0; r < rounds; r++) {
            const uint32_t i = r * 32 + lane, m = s.starts[r];
            const uint32_t upto = m & ((2u << lane) - 1u);   // lane 31: (2u << 31) - 1 = all ones
            const uint32_t run_start = upto ? r * 32 + 31 - __clz(upto) : carry;
            ChunkPlan c;
            if (i < n && chunk_plan(s, i, run_start, lz_min, c)) {
                if (c.lit_tok >= (uint32_t)kLitBins || c.lit_nbits > 12)
                    err |= kErrLfAlphabet;
                else
                    atomicAdd(&w.freq[c.lit_tok], c.lits);
                max_tok = max_tok > c.lit_tok ? max_tok : c.lit_tok;
                if (c.rep_tok) {
                    atomicAdd(&w.freq[kLitBins + (c.rep_tok - lz_min)], 1u);
                    atomicAdd(&w.freq[kBins + 1], 1u);   // distance symbol: value 1 -> token 1 (config 7,1,1)
                    max_tok = max_tok > c.rep_tok ? max_tok : c.rep_tok;
                    any_rep = 1;
                }
            }
            if (m)
                carry = r * 32 + 31 - __clz(m);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint32_t o = __shfl_xor_sync(FULL, max_tok, d);
            max_tok = o > max_tok ? o : max_tok;
            any_rep |= __shfl_xor_sync(FULL, any_rep, d);
            err |= __shfl_xor_sync(FULL, err, d);
        }
        if (lane == 0) {
            w.alpha0 = max_tok + 1;
            w.alpha1 = any_rep ? 2 : 0;
        }
    }
    __syncwarp();
    // ---- 4./5. code lengths, codes, header --------------------------------------------------------
    if (w.alpha0 > 1)
        warp_code_lengths(w, w.alpha0, 15, lz_min, lane);
    uint32_t hdr_bits = 0;
    {
        BitSink bw;
        bw.init(out, kLfBitsWords);
        if (lane == 0) {
            // stream preamble + code descriptions; mirrors ps_put_header but with the lengths of
            // cluster 0 already computed by the warp
            const U32Dist kMinSymbol = {{224, 512, 4096, 8}, {0, 0, 0, 15}};
            const U32Dist kMinLength = {{3, 4, 5, 9}, {0, 0, 2, 8}};
            bw.put_bool(1);
            put_u32(bw, kMinSymbol, lz_min);
            put_u32(bw, kMinLength, 3);
            ps_put_hybrid_cfg(bw, 7, 0, 0, 8);
            bw.put_bool(1);   // simple cluster map, 1 bit per context: contexts {0, lz77} -> clusters {0, 1}
            bw.put(1, 2);
            bw.put(0, 1);
            bw.put(1, 1);
            bw.put_bool(1);   // prefix codes
            ps_put_hybrid_cfg(bw, prm.split0, prm.msb0, prm.lsb0, 15);
            ps_put_hybrid_cfg(bw, prm.split1, prm.msb1, prm.lsb1, 15);
            const uint32_t alpha[2] = {w.alpha0, w.alpha1};
            for (int c = 0; c < 2; c++) {
                if (alpha[c] <= 1) {
                    bw.put_bool(0);
                    continue;
                }
                bw.put_bool(1);
                const int nbv = floor_log2_u32(alpha[c] - 1);
                bw.put((uint32_t)nbv, 4);
                bw.put(alpha[c] - 1, nbv);
            }
        }
        __syncwarp();
        // cluster 0 code description needs the canonical codes only afterwards; lengths are ready
        if (lane == 0 && w.alpha0 > 1) {
            uint32_t used = 0, fsym[4] = {0, 0, 0, 0}, flen[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; b < (uint32_t)kBins; b++) {
                if (!w.len[b])
                    continue;
                if (used < 4) {
                    fsym[used] = ps_bin_token(b, lz_min);
                    flen[used] = w.len[b];
                }
                if (++used > 4)
                    break;
            }
            if (used > 4) {
                ps_put_complex_code(w, bw, 0, kBins, w.alpha0, lz_min);
            } else {
                if (!used) {
                    used = 1;
                    fsym[0] = w.alpha0 - 1;
                }
                bw.put(1, 2);
                bw.put(used - 1, 2);
#define HYDB_SWAP_FEW(a, b) do { uint32_t ts = fsym[a], tl = flen[a]; fsym[a] = fsym[b]; flen[a] = flen[b]; \
                                 fsym[b] = ts; flen[b] = tl; } while (0)
                if (used == 3 && flen[0] != 1) {
                    if (flen[1] == 1) HYDB_SWAP_FEW(0, 1); else HYDB_SWAP_FEW(0, 2);
                }
                int select = 0;
                if (used == 4) {
                    for (int q = 0; q < 4; q++)
                        if (flen[q] != 2) { select = 1; break; }
                    if (select && flen[0] != 1) {
                        if (flen[1] == 1) HYDB_SWAP_FEW(0, 1);
                        else if (flen[2] == 1) HYDB_SWAP_FEW(0, 2);
                        else HYDB_SWAP_FEW(0, 3);
                    }
                    if (select && flen[1] != 2) {
                        if (flen[2] == 2) HYDB_SWAP_FEW(1, 2); else HYDB_SWAP_FEW(1, 3);
                    }
                }
#undef HYDB_SWAP_FEW
                const int width = ceil_log2_u32(w.alpha0);
                for (uint32_t q = 0; q < used; q++)
                    bw.put(fsym[q], width);
                if (used == 4)
                    bw.put_bool(select);
            }
        }
        // cluster 1 (distance symbols): a single used token -> "simple code, 1 symbol" (entropy.c:879-886)
        if (lane == 0) {
            w.len[kBins] = w.len[kBins + 1] = 0;
            w.code[kBins] = w.code[kBins + 1] = 0;
            if (w.alpha1 > 1) {
                bw.put(1, 2);
                bw.put(0, 2);
                bw.put(w.alpha1 - 1, ceil_log2_u32(w.alpha1));
            }
            bw.flush_partial();
            if (bw.overflow)
                w.error |= kErrLfCapacity;
        }
        hdr_bits = __shfl_sync(FULL, bw.bitlen(), 0);
    }
    __syncwarp();
    warp_assign_codes(s, lane);
    // ---- 6. symbol bits -------------------------------------------------------------------------
    // sweep 0 sizes the string (so the tail can be zeroed), sweep 1 writes it
    uint32_t total_bits = 0;
    for (int sweep = 0; sweep < 2; sweep++) {
        uint32_t carry = 0, pos = hdr_bits;
        for (uint32_t r = 0; r < rounds; r++) {
            const uint32_t i = r * 32 + lane, m = s.starts[r];
            const uint32_t upto = m & ((2u << lane) - 1u);
            const uint32_t run_start = upto ? r * 32 + 31 - __clz(upto) : carry;
            ChunkPlan c;
            const bool head = i < n && chunk_plan(s, i, run_start, lz_min, c);
            uint32_t lit_len = 0, lit_bits = 0, rep_bits = 0, mine = 0;
            if (head) {
                lit_len = w.len[c.lit_tok];
                lit_bits = lit_len + c.lit_nbits;
                rep_bits = c.rep_tok ? (uint32_t)w.len[kLitBins + (c.rep_tok - lz_min)] + w.len[kBins + 1] : 0;
                mine = c.lits * lit_bits + rep_bits;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= (uint32_t)d)
                    incl += o;
            }
            if (sweep == 1 && head && mine) {
                uint64_t at = (uint64_t)pos + incl - mine;
                auto put = [&](uint32_t v, uint32_t nbits) {
                    if (!nbits)
                        return;
                    const uint32_t wd = (uint32_t)(at >> 5), sh = (uint32_t)(at & 31);
                    if (wd + 1 < (uint32_t)kLfBitsWords) {
                        atomicOr(&out[wd], v << sh);
                        if (sh + nbits > 32)
                            atomicOr(&out[wd + 1], v >> (32 - sh));
                    }
                    at += nbits;
                };
                const uint32_t lit_code = w.code[c.lit_tok];
                for (uint32_t q = 0; q < c.lits; q++) {
                    put(lit_code, lit_len);
                    put(c.lit_res, c.lit_nbits);
                }
                if (c.rep_tok) {
                    const uint32_t b = kLitBins + (c.rep_tok - lz_min);
                    put(w.code[b], w.len[b]);
                    put(w.code[kBins + 1], w.len[kBins + 1]);
                }
            }
            pos += __shfl_sync(FULL, incl, 31);
            if (m)
                carry = r * 32 + 31 - __clz(m);
        }
        if (sweep == 0) {
            total_bits = pos;
            // zero everything after the header's last (partial) word, which lane 0 already wrote
            const uint32_t w0 = (hdr_bits + 31) >> 5, w1 = (total_bits + 31) >> 5;
            for (uint32_t q = w0 + lane; q <= w1 && q < (uint32_t)kLfBitsWords; q += 32)
                out[q] = 0;
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t e = err | w.error;
        if (((total_bits + 31) >> 5) + 1 >= (uint32_t)kLfBitsWords)
            e |= kErrLfCapacity;
        if (e)
            atomicOr(&tile_err[tile], e);
        lfbitlen[tile] = e ? 0 : total_bits;
    }
}

struct TemplShared {
    PrefixWork work;
    uint32_t syms[kSectionSymCap];
};

// one thread per section: block 0 builds section A (if asked), block 1 + k builds section B of shape k
__global__ void __launch_bounds__(32)
k_build_templates(Templates t, const uint32_t *__restrict__ shape_dims, uint32_t first_shape, int build_a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TemplShared &s = *reinterpret_cast<TemplShared *>(smem_raw);
    if (threadIdx.x != 0)
        return;
    s.work.error = 0;
    BitSink bw;
    if (blockIdx.x == 0) {
        if (!build_a)
            return;
        bw.init(t.words, kTemplWords);
        build_section_a(s.work, s.syms, bw);
        bw.flush_partial();
        t.bits[0] = (bw.overflow || s.work.error) ? 0xFFFFFFFFu : bw.bitlen();
    } else {
        const uint32_t k = blockIdx.x - 1, slot = 1 + first_shape + k;
        bw.init(t.words + (size_t)slot * kTemplWords, kTemplWords);
        build_section_b(s.work, s.syms, bw, shape_dims[2 * k], shape_dims[2 * k + 1]);
        bw.flush_partial();
        t.bits[slot] = (bw.overflow || s.work.error) ? 0xFFFFFFFFu : bw.bitlen();
    }
}

void launch_lf_group(const Workspace &ws, uint32_t ntiles, cudaStream_t st) {
    cudaFuncSetAttribute(k_lf_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LfShared));
    prefer_max_shared(k_lf_group);
    k_lf_group<<<ntiles, 32, sizeof(LfShared), st>>>(ws.tiles, ws.lfq, ws.lfbits, ws.lfbitlen, ws.tile_err);
}

void launch_build_templates(const Templates &t, const uint32_t *d_shape_dims, uint32_t first_shape, uint32_t n_shapes,
                            bool build_a, cudaStream_t st) {
    cudaFuncSetAttribute(k_build_templates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TemplShared));
    k_build_templates<<<1 + n_shapes, 32, sizeof(TemplShared), st>>>(t, d_shape_dims, first_shape, build_a ? 1 : 0);
}

}  // namespace hydb
