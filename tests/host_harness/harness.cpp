// tests/host_harness/harness.cpp -- TEST TOOL, not a product path.
//
// Compiles the product's host/device headers (hydrium_b200/csrc/*.cuh, the sequential entropy
// logic the CUDA kernels run in one thread per tile) with a plain C++ compiler so that the very
// same source can be checked against the oracle on a machine without a GPU.  The product never
// runs this code on the CPU: libhydrium_b200.so contains no host implementation of the encoder.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ans_chain.cuh"
#include "headers.cuh"
#include "lf_values.cuh"
#include "sections.cuh"

using namespace hydb;

extern "C" {

// generic prefix stream over one context
uint32_t hh_prefix_stream(const uint32_t *values, uint32_t n, uint32_t plain_dists, uint32_t lz_min, uint32_t modular,
                          const int *cfg0, const int *cfg1, uint32_t *out, uint32_t cap_words, uint32_t *bitlen) {
    PrefixWork *w = new PrefixWork();
    memset(w, 0, sizeof(*w));
    std::vector<uint32_t> syms(n + 16);
    PrefixParams p;
    p.num_plain_dists = plain_dists;
    p.lz_min_symbol = lz_min;
    p.modular = modular;
    p.split0 = (uint8_t)cfg0[0]; p.msb0 = (uint8_t)cfg0[1]; p.lsb0 = (uint8_t)cfg0[2];
    p.split1 = (uint8_t)cfg1[0]; p.msb1 = (uint8_t)cfg1[1]; p.lsb1 = (uint8_t)cfg1[2];
    BitSink bw;
    bw.init(out, cap_words);
    struct V { const uint32_t *v; uint32_t operator()(uint32_t i) const { return v[i]; } } va{values};
    ps_encode_stream(*w, syms.data(), (uint32_t)syms.size(), p, n, va, bw);
    bw.flush_partial();
    *bitlen = bw.bitlen();
    uint32_t err = w->error | (bw.overflow ? (uint32_t)kErrLfCapacity : 0u);
    delete w;
    return err;
}

uint32_t hh_section_a(uint32_t *out, uint32_t cap_words, uint32_t *bitlen) {
    PrefixWork *w = new PrefixWork();
    memset(w, 0, sizeof(*w));
    std::vector<uint32_t> syms(kSectionSymCap);
    BitSink bw;
    bw.init(out, cap_words);
    build_section_a(*w, syms.data(), bw);
    bw.flush_partial();
    *bitlen = bw.bitlen();
    uint32_t err = w->error | (bw.overflow ? (uint32_t)kErrLfCapacity : 0u);
    delete w;
    return err;
}

uint32_t hh_section_b(uint32_t vbw, uint32_t vbh, uint32_t *out, uint32_t cap_words, uint32_t *bitlen) {
    PrefixWork *w = new PrefixWork();
    memset(w, 0, sizeof(*w));
    std::vector<uint32_t> syms(kSectionSymCap);
    BitSink bw;
    bw.init(out, cap_words);
    build_section_b(*w, syms.data(), bw, vbw, vbh);
    bw.flush_partial();
    *bitlen = bw.bitlen();
    uint32_t err = w->error | (bw.overflow ? (uint32_t)kErrLfCapacity : 0u);
    delete w;
    return err;
}

// LF stream from the three planes of quantised LF ints (X, Y, B; row stride 32)
uint32_t hh_lf_stream(const int32_t *lfq, uint32_t vbw, uint32_t vbh, uint32_t *out, uint32_t cap_words, uint32_t *bitlen) {
    PrefixWork *w = new PrefixWork();
    memset(w, 0, sizeof(*w));
    std::vector<uint32_t> syms(3 * kMaxBlocks + 16);
    BitSink bw;
    bw.init(out, cap_words);
    LfStreamValues va{lfq, vbw, vbw * vbh};
    ps_encode_stream(*w, syms.data(), (uint32_t)syms.size(), lf_stream_params(), 3 * vbw * vbh, va, bw);
    bw.flush_partial();
    *bitlen = bw.bitlen();
    uint32_t err = w->error | (bw.overflow ? (uint32_t)kErrLfCapacity : 0u);
    delete w;
    return err;
}

// 1: hh_ans_encode runs the chain over the COMPACT inverse alias map (sorted pieces, 32 emulated lanes,
// maximum of the lane candidates) the way k_ans_chain_compact does; 0: the direct table of k_ans_chain
static int g_compact = 0;
void hh_set_compact(int on) { g_compact = on; }

// HF symbols (packed hf_pack records) -> normalised frequencies, section D bits, section E bits.
// Mirrors what k_ans.cu does per tile, sequentially.
uint32_t hh_ans_encode(const uint32_t *syms, uint32_t n, uint32_t *freqs_out /*[9][64]*/, uint32_t *alpha_out /*[9]*/,
                       uint32_t *d_out, uint32_t d_cap, uint32_t *d_bitlen,
                       uint32_t *e_out, uint32_t e_cap, uint32_t *e_bitlen) {
    uint32_t err = 0;
    static uint32_t hist[kHfClusters][kHfTokens];
    uint32_t alpha[kHfClusters] = {0}, max_alpha = 0;
    memset(hist, 0, sizeof(hist));
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t t = hf_token(syms[i]), c = hf_cluster(syms[i]);
        if (t >= (uint32_t)kHfTokens || c >= (uint32_t)kHfClusters)
            return kErrAlphabet;
        hist[c][t]++;
        if (t + 1 > alpha[c]) alpha[c] = t + 1;
        if (t + 1 > max_alpha) max_alpha = t + 1;
    }
    int log_alpha = ceil_log2_u32(max_alpha);
    if (log_alpha < 5) log_alpha = 5;
    static AnsCluster cl[kHfClusters];
    static uint16_t inv[kHfClusters][kAnsTotal];
    static AnsSymInfo info[kHfClusters][kHfTokens];
    memset(cl, 0, sizeof(cl));
    for (int c = 0; c < kHfClusters; c++) {
        alpha_out[c] = alpha[c];
        if (!alpha[c])
            continue;
        const int single = ans_normalise(hist[c], alpha[c]);
        if (single < 0)
            return kErrAlias;
        if (!ans_build_alias(cl[c], hist[c], alpha[c], log_alpha, single != 0))
            return kErrAlias;
        for (uint32_t k = 0; k < (uint32_t)kHfTokens; k++) {
            freqs_out[c * kHfTokens + k] = cl[c].freq[k];
            info[c][k] = ans_sym_info(cl[c].freq[k], (uint32_t)c * kAnsTotal + cl[c].cum[k]);
        }
        for (uint32_t s = 0; s < (uint32_t)kAnsTotal; s++) {
            uint32_t sym, off;
            ans_slot_symbol(cl[c], s, log_alpha, sym, off);
            inv[c][cl[c].cum[sym] + off] = (uint16_t)s;
        }
    }
    // section D (reference: entropy.c:563-572, 988-992)
    {
        BitSink bw;
        bw.init(d_out, d_cap);
        bw.put_bool(0);
        bw.put((uint32_t)(log_alpha - 5), 2);
        for (int c = 0; c < kHfClusters; c++)
            ps_put_hybrid_cfg(bw, 4, 1, 0, log_alpha);
        for (int c = 0; c < kHfClusters; c++)
            ans_put_histogram(bw, cl[c].freq, alpha[c]);
        bw.flush_partial();
        *d_bitlen = bw.bitlen();
        if (bw.overflow) err |= kErrSlab;
    }
    // chain, last symbol first, the way k_ans_chain runs it: every step yields the pre-renormalisation
    // state s'; the flag / word of symbol p are derived from the s' of step p + 1 and f(p)
    std::vector<uint8_t> flag(n + 1, 0);
    std::vector<uint16_t> words;
    uint32_t x = 0;
    if (n) {
        std::vector<uint32_t> sprime(n + 1, 0);
        sprime[n] = kAnsInitState;   // the virtual step before the last symbol
        auto freq_of = [&](uint32_t p) { return asi_freq(info[hf_cluster(syms[p])][hf_token(syms[p])]); };
        const uint8_t *inv_bytes = (const uint8_t *)&inv[0][0];
        auto info_of = [&](uint32_t p) -> const AnsSymInfo & { return info[hf_cluster(syms[p])][hf_token(syms[p])]; };
        AnsCarry carry;
        if (g_compact) {
            if (log_alpha != 5)
                return kErrAlphabet;
            static AnsPieceLane lanes[kHfClusters][32];
            for (int c = 0; c < kHfClusters; c++) {
                uint32_t lo[kAnsPieces];
                int32_t delta[kAnsPieces];
                uint64_t covered = 0;
                for (uint32_t p = 0; p < (uint32_t)kAnsPieces; p++) {
                    const uint32_t len = alpha[c] ? ans_piece(cl[c], p, lo[p], delta[p]) : 0u;
                    if (!len)
                        lo[p] = kAnsPieceNone;
                    covered += len;
                }
                if (alpha[c] && covered != (uint64_t)kAnsTotal)
                    return kErrAlias;
                for (uint32_t p = 0; p < (uint32_t)kAnsPieces; p++) {
                    const uint32_t rank = ans_piece_rank(lo, p), L = rank >> 1;
                    uint32_t *pl = reinterpret_cast<uint32_t *>(&lanes[c][L]);
                    pl[rank & 1u] = lo[p];
                    pl[2u + (rank & 1u)] = (uint32_t)delta[p] + (L << 13);
                }
                // the pieces must reproduce the direct table everywhere
                for (uint32_t g = 0; alpha[c] && g < (uint32_t)kAnsTotal; g++) {
                    uint32_t best = 0;
                    for (int L = 0; L < 32; L++) {
                        const uint32_t cand = ans_piece_candidate(lanes[c][L], g);
                        best = cand > best ? cand : best;
                    }
                    if ((best & 0xFFFu) != inv[c][g])
                        return kErrAlias | 0x80000000u;
                }
            }
            auto rec_of = [&](uint32_t p) {
                const uint32_t c = hf_cluster(syms[p]), k = hf_token(syms[p]);
                return ans_rec_c(cl[c].freq[k], cl[c].cum[k]);
            };
            ans_chain_begin_c(carry, rec_of(n - 1));
            for (uint32_t r = 0; r < n; r++) {
                const uint32_t p = n - 1 - r, c = hf_cluster(syms[p]);
                const AnsRecC own = rec_of(p), nx = p ? rec_of(p - 1) : own;
                ans_step_c(carry, own, p ? &nx : nullptr,
                           [&](uint32_t g) {
                               uint32_t best = 0;
                               for (int L = 0; L < 32; L++) {
                                   const uint32_t cand = ans_piece_candidate(lanes[c][L], g);
                                   best = cand > best ? cand : best;
                               }
                               return best & 0xFFFu;
                           },
                           sprime[p]);
            }
        } else {
        ans_chain_begin(carry, info_of(n - 1), 0u);
        for (uint32_t r = 0; r < n; r++) {
            const uint32_t p = n - 1 - r;
            ans_step(carry, info_of(p), p ? &info_of(p - 1) : nullptr, 0u,
                     [inv_bytes](uint32_t off) { uint16_t v; memcpy(&v, inv_bytes + off, 2); return (uint32_t)v; },
                     sprime[p]);
        }
        }
        x = sprime[0];   // the state left by the last step is final (no renormalisation follows)
        for (uint32_t p = n; p-- > 0;) {   // descending, like the chain emits them
            if ((sprime[p + 1] >> 20) >= freq_of(p)) {
                flag[p] = 1;
                words.push_back((uint16_t)(sprime[p + 1] & 0xFFFF));
            }
        }
    }
    // forward emission: final state, then per symbol [word][residue]
    {
        BitSink bw;
        bw.init(e_out, e_cap);
        if (n) {
            bw.put(x, 32);
            size_t wi = words.size();
            uint32_t last = 0;
            bool first = true;
            for (uint32_t p = 0; p < n; p++) {
                if (flag[p]) {
                    if ((first ? p : p - last) >= 65536u)
                        err |= kErrAnsGap;
                    first = false;
                    last = p;
                    bw.put(words[--wi], 16);
                }
                bw.put(hf_residue(syms[p]), (int)hf_nbits(syms[p]));
            }
        }
        bw.flush_partial();
        *e_bitlen = bw.bitlen();
        if (bw.overflow) err |= kErrSlab;
    }
    return err;
}

uint32_t hh_image_header(uint32_t width, uint32_t height, uint32_t *out, uint32_t cap_words) {
    BitSink bw;
    bw.init(out, cap_words);
    put_image_header(bw, width, height);
    bw.flush_partial();
    return bw.bitlen();
}

uint32_t hh_frame_header(int crop, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, int last, uint32_t payload_bytes,
                         uint32_t *out, uint32_t cap_words) {
    BitSink bw;
    bw.init(out, cap_words);
    put_frame_header(bw, crop != 0, x0, y0, w, h, last != 0);
    put_toc_entry(bw, payload_bytes);
    bw.flush_partial();
    return bw.bitlen();
}

// Exactness of the corrected-reciprocal division of ans_chain.cuh: for every frequency, states
// x = a + v at the range boundaries and pseudo-random ones must give q = x / f and the table index
// x % f.  Returns the number of mismatches.
uint64_t hh_div_check(uint32_t f_lo, uint32_t f_hi, uint32_t samples) {
    uint64_t bad = 0;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    for (uint32_t f = f_lo; f <= f_hi; f++) {
        const AnsSymInfo si = ans_sym_info(f, 0);
        const uint64_t lim = (uint64_t)f << 20;   // x < f * 2^20
        auto check1 = [&](uint32_t a, uint32_t v, bool counts) {
            AnsCarry c;
            c.v = v;
            ans_prepare(c, a, counts, si.mc, si.ne, 0x12340u);
            const uint32_t q = ans_hi32((uint64_t)c.v * c.meff + c.R);
            const uint32_t addr = q * si.nf2 + (c.v * c.k + c.c0);
            const uint32_t x = a + (counts ? v : 0u);
            if (q != x / f || addr != 0x12340u + 2u * (x % f)) bad++;
        };
        auto check = [&](uint64_t x) {
            if (x >= lim || x > 0xFFFFFFFFull || x < 65536) return;
            const uint32_t xx = (uint32_t)x;
            check1(xx & ~0xFFFu, xx & 0xFFFu, true);   // not renormalised: a = q_prev << 12, v = slot
            if (xx < 65536u * 16u) {                   // renormalised: a is the whole state, v is ignored
                check1(xx, 0xFFFu, false);
                check1(xx, (uint32_t)(rng >> 52), false);
            }
        };
        for (uint64_t q = 0; q < 64; q++) {
            check(q * f); check(q * f + f - 1); if (q) check(q * f - 1);
            const uint64_t qq = (1ull << 20) - 1 - q;
            check(qq * f); check(qq * f + f - 1); check(qq * f - 1);
        }
        for (int b = 1; b < 32; b++) {
            const uint64_t x = 1ull << b;
            check(x); check(x - 1); check(x + 1);
            const uint64_t qv = x / f;
            check(qv * f); check(qv * f + f - 1); if (qv) check(qv * f - 1);
            for (uint64_t d = 0; d < 4096; d += 37) { check((x & ~0xFFFull) + d); check((qv * f & ~0xFFFull) + d); }
        }
        for (uint32_t i = 0; i < samples; i++) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            const uint64_t qv = rng % (1ull << 20);
            check(qv * f); check(qv * f + f - 1); if (qv) check(qv * f - 1); check(qv * f + (rng >> 40) % f);
            check(65536 + (rng >> 44));   // small states, as after a renormalisation
        }
    }
    return bad;
}

}  // extern "C"
