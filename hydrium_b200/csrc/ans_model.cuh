// hydrium_b200/csrc/ans_model.cuh
//
// Per-cluster rANS model for the HF coefficient stream: histogram normalisation to 4096,
// alias-table construction, histogram header, and the constants the encoder chain needs.
// (reference: entropy.c:267-301 normalisation, 184-265 alias mapping, 303-369 histogram header,
// 1083-1120 encoder step).
//
// Formulation differences from the reference (results identical):
//   * the alias table is kept in its per-bucket form (owner / cut / offset) and inverted into a
//     direct slot table  inv[cum[sym] + offset] = (bucket << log_bucket) | pos  -- the alias map
//     is a bijection on [0, 4096), so the reference's per-symbol entry search (entropy.c:1106-1113)
//     becomes one shared-memory load in the serial chain;
//   * state / freq is split into two exact multiply-high divisions (ans_chain.cuh).
#pragma once

#include "bitio.cuh"
#include "common.cuh"

namespace hydb {

template <int NT>
struct AnsClusterT {
    static constexpr int kTokens = NT;
    uint16_t freq[NT];   // normalised frequencies (0 for unused)
    uint16_t cum[NT];    // exclusive prefix sums
    uint8_t owner[NT];   // per bucket: symbol that owns the part above `cut`
    uint16_t cut[NT];    // per bucket: size of the bucket's own symbol share
    int16_t off[NT];     // per bucket: symbol offset of position 0 of the foreign part
    uint32_t alpha;      // alphabet size (max token + 1), 0 = cluster unused
    uint32_t single;     // 1: one symbol with frequency 4096
};
using AnsCluster = AnsClusterT<kHfTokens>;   // log_alphabet_size 5 or 6
using AnsCluster32 = AnsClusterT<32>;        // log_alphabet_size 5 only (k_ans_chain_compact)

// counts -> 12-bit frequencies (reference: entropy.c:267-301).  Returns -1 if all zero, else
// whether the LAST symbol took everything.
HDN inline int ans_normalise(uint32_t *f, uint32_t n) {
    uint64_t total = 0;
    for (uint32_t k = 0; k < n; k++)
        total += f[k];
    if (!total)
        return -1;
    uint32_t sum = 0;
    const bool narrow = total < (1ull << 20);   // then f << 12 fits 32 bits: avoid the slow 64-bit divide
    for (uint32_t k = 0; k < n; k++) {
        if (!f[k])
            continue;
        uint32_t v = narrow ? (((f[k] << 12) / (uint32_t)total) & 0xFFFFu)
                            : (uint32_t)((((uint64_t)f[k] << 12) / total) & 0xFFFFu);
        if (!v)
            v = 1;
        f[k] = v;
        sum += v;
    }
    uint32_t j = n - 1;
    while (sum > (uint32_t)kAnsTotal) {
        const uint32_t excess = sum - kAnsTotal;
        if (excess < f[j]) {
            f[j] -= excess;
            sum -= excess;
            break;
        } else if (f[j] > 1) {
            sum -= f[j] - 1;
            f[j] = 1;
        }
        j--;
    }
    f[0] += (uint32_t)kAnsTotal - sum;
    return f[n - 1] == (uint32_t)kAnsTotal;
}

// Vose-style alias split with the reference's LIFO work lists (entropy.c:184-242).
// counts[] holds the NORMALISED frequencies.  Returns false where the reference errors out.
template <class Cluster>
HDN inline bool ans_build_alias(Cluster &c, const uint32_t *normalised, uint32_t alpha, int log_alpha, bool single) {
    const uint32_t bucket = 1u << (12 - log_alpha), slots = 1u << log_alpha;
    c.alpha = alpha;
    c.single = single ? 1u : 0u;
    uint32_t run = 0;
    for (uint32_t s = 0; s < (uint32_t)Cluster::kTokens; s++) {
        const uint32_t f = s < alpha ? normalised[s] : 0;
        c.freq[s] = (uint16_t)f;
        c.cum[s] = (uint16_t)run;
        run += f;
        c.owner[s] = (uint8_t)s;
        c.cut[s] = 0;
        c.off[s] = 0;
    }
    if (single) {
        // every bucket belongs to the one symbol; slot == symbol offset (entropy.c:195-200)
        for (uint32_t i = 0; i < slots; i++) {
            c.owner[i] = (uint8_t)(alpha - 1);
            c.cut[i] = 0;
            c.off[i] = (int16_t)(i * bucket);
        }
        return true;
    }
    uint32_t cut[kHfTokens], off[kHfTokens];
    uint8_t small_[kHfTokens], large_[kHfTokens];
    uint32_t ns = 0, nl = 0;
    for (uint32_t p = 0; p < slots; p++) {
        cut[p] = p < alpha ? normalised[p] : 0;
        off[p] = 0;
    }
    for (uint32_t p = 0; p < alpha; p++) {
        if (cut[p] < bucket) small_[ns++] = (uint8_t)p;
        else if (cut[p] > bucket) large_[nl++] = (uint8_t)p;
    }
    for (uint32_t i = alpha; i < slots; i++)
        small_[ns++] = (uint8_t)i;
    while (nl) {
        if (!ns)
            return false;   // reference: "empty underfull during alias table gen"
        const uint32_t u = small_[--ns], o = large_[--nl];
        const uint32_t by = bucket - cut[u];
        cut[o] -= by;
        off[u] = cut[o];
        c.owner[u] = (uint8_t)o;
        if (cut[o] < bucket) small_[ns++] = (uint8_t)o;
        else if (cut[o] > bucket) large_[nl++] = (uint8_t)o;
    }
    for (uint32_t s = 0; s < slots; s++) {
        if (cut[s] == bucket) {
            c.owner[s] = (uint8_t)s;
            c.cut[s] = 0;
            c.off[s] = 0;
        } else {
            c.cut[s] = (uint16_t)cut[s];
            c.off[s] = (int16_t)((int32_t)off[s] - (int32_t)cut[s]);
        }
    }
    return true;
}

// slot s of the alias table decodes to (symbol, offset) (reference: entropy.c:233-262 read backwards)
template <class Cluster>
HD void ans_slot_symbol(const Cluster &c, uint32_t s, int log_alpha, uint32_t &sym, uint32_t &offset) {
    const uint32_t lb = 12u - (uint32_t)log_alpha;
    const uint32_t i = s >> lb, pos = s & ((1u << lb) - 1u);
    if (c.single) {
        sym = c.alpha - 1;
        offset = s;
    } else if (pos < c.cut[i] || (c.cut[i] == 0 && c.owner[i] == i)) {
        // own share; a completely full bucket is recorded as cut 0 / owner self / off 0
        sym = i;
        offset = pos;
    } else {
        sym = c.owner[i];
        offset = (uint32_t)((int32_t)c.off[i] + (int32_t)pos);
    }
}

// ---- compact form of the inverse alias map (log_alpha 5: 32 buckets of 128 slots) ----------------
// The alias map is a bijection between g = cum[sym] + offset in [0, 4096) and the slots.  Every
// bucket i contributes at most two PIECES on which  slot = g + delta  holds:
//   own share      sym i          offsets [0, own)                 slots [128 i, 128 i + own)
//   foreign share  sym owner[i]   offsets [off + cut, off + 128)   slots [128 i + cut, 128 i + 128)
// (reference: entropy.c:233-262 read backwards), so <= 64 pieces partition [0, 4096).  Sorted by their
// first g they are dealt two per lane; lane L keeps {lo0, lo1, delta0 + (L << 13), delta1 + (L << 13)}
// and the slot of g is found by the whole warp at once:
//     cand = g >= lo0 ? g + (g >= lo1 ? e1 : e0) : 0          slot = max over lanes (cand) & 0xFFF
// The owner is the highest lane with lo0 <= g; lower lanes see g ABOVE their pieces, so their
// g + delta stays in [0, 8190] and never reaches the lane tag in bits 13+.  4.6 KB per tile instead of
// the 72 KB direct table: the form k_ans_chain_compact uses when many chains must share an SM.
constexpr int kAnsPieces = 64;
constexpr uint32_t kAnsPieceNone = 4096;   // first g of an unused piece: never reached
// piece p of the cluster: p < 32 own share of bucket p, p >= 32 foreign share of bucket p - 32.
// Returns its length (0 = unused).
template <class Cluster>
HD uint32_t ans_piece(const Cluster &c, uint32_t p, uint32_t &lo, int32_t &delta) {
    const uint32_t bucket = 128u, i = p & 31u;
    lo = kAnsPieceNone;
    delta = 0;
    if (c.single) {   // one symbol owns everything and slot == offset (entropy.c:195-200)
        if (p)
            return 0;
        lo = 0;
        return (uint32_t)kAnsTotal;
    }
    const bool full = c.cut[i] == 0 && c.owner[i] == i;   // the bucket holds nothing but its own symbol
    if (p < 32u) {
        const uint32_t own = full ? bucket : c.cut[i];
        if (!own || !c.freq[i])
            return 0;
        lo = c.cum[i];
        delta = (int32_t)(i * bucket) - (int32_t)c.cum[i];
        return own;
    }
    if (full)
        return 0;
    const uint32_t o = c.owner[i];
    lo = (uint32_t)((int32_t)c.cum[o] + (int32_t)c.off[i] + (int32_t)c.cut[i]);
    delta = (int32_t)(i * bucket) - (int32_t)c.off[i] - (int32_t)c.cum[o];
    return bucket - c.cut[i];
}
// position of piece p in the sorted order (used pieces by first g, then the unused ones): O(64)
HD uint32_t ans_piece_rank(const uint32_t *lo /*[64]*/, uint32_t p) {
    uint32_t r = 0;
    const uint32_t mine = lo[p];
    for (uint32_t q = 0; q < (uint32_t)kAnsPieces; q++)
        r += (lo[q] < mine || (lo[q] == mine && q < p)) ? 1u : 0u;
    return r;
}
struct AnsPieceLane {
    uint32_t lo0, lo1, e0, e1;
};
// what one lane contributes to the warp-wide maximum
HD uint32_t ans_piece_candidate(const AnsPieceLane &pl, uint32_t g) {
    return g >= pl.lo0 ? g + (g >= pl.lo1 ? pl.e1 : pl.e0) : 0u;
}

HD void ans_put_u8(BitSink &bw, uint32_t b) {   // reference: entropy.c:71-78
    bw.put_bool(b != 0);
    if (!b)
        return;
    const int l = floor_log2_u32(b);
    bw.put((uint32_t)l, 3);
    bw.put(b, l);
}

// histogram header of one cluster (reference: entropy.c:303-369)
HDN inline void ans_put_histogram(BitSink &bw, const uint16_t *f, uint32_t alpha) {
    static const uint8_t kLcBits[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};   // entropy.c:35-38
    static const uint8_t kLcLen[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
    if (!alpha) {
        bw.put(1, 2);
        ans_put_u8(bw, 0);
        return;
    }
    int a = -1, b = -1, seen = 0;
    for (uint32_t k = 0; k < alpha; k++) {
        if (f[k] == kAnsTotal) {
            bw.put(1, 2);
            ans_put_u8(bw, k);
            return;
        }
        if (!f[k])
            continue;
        if (++seen > 2)
            break;
        if (a < 0) {
            a = (int)k;
        } else if ((uint32_t)f[a] + f[k] == (uint32_t)kAnsTotal) {
            b = (int)k;
            break;
        }
    }
    if (a >= 0 && b >= 0) {
        bw.put(3, 2);
        ans_put_u8(bw, (uint32_t)a);
        ans_put_u8(bw, (uint32_t)b);
        bw.put(f[a], 12);
        return;
    }
    bw.put(0, 2);
    bw.put(7, 3);
    bw.put(6, 3);
    ans_put_u8(bw, alpha - 3);
    uint32_t omit = 0;
    int omit_log = 0;
    for (uint32_t k = 0; k < alpha; k++) {
        const int lc = f[k] ? 1 + floor_log2_u32(f[k]) : 0;
        bw.put(kLcBits[lc], kLcLen[lc]);
        if (lc > omit_log) {
            omit_log = lc;
            omit = k;
        }
    }
    for (uint32_t k = 0; k < alpha; k++) {
        const int lc = f[k] ? 1 + floor_log2_u32(f[k]) : 0;
        if (k == omit || lc <= 1)
            continue;
        bw.put(f[k], lc - 1);
    }
}

}  // namespace hydb
