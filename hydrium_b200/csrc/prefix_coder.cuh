// hydrium_b200/csrc/prefix_coder.cuh
//
// Prefix-coded (Huffman) entropy streams as the reference writes them for the LF coefficients,
// the modular HF-metadata image, the MA trees and the nested cluster map
// (reference: entropy.c:371-524 front end, 546-575 preamble, 577-941 code construction,
// 1003-1034 symbol emission).
//
// B200-side formulation, not a translation:
//   * alphabets are SPARSE.  The LF stream uses lz77_min_symbol = 16384 (encoder.c:567), so the
//     reference builds 2*16509-slot arrays and scans them O(alphabet x symbols).  Here only the
//     used tokens exist, as `bins` (<= 384), and the reference's slot-swapping tree builder
//     (entropy.c:592-662) is simulated on slot *positions* of live nodes only.
//   * the run-length front end is a pure function of the value sequence (chunks of one literal
//     plus up to 127 repeats), so no encoder state survives between calls.
// The routines are sequential (one thread drives them over shared-memory scratch) because every
// stream here is <= 3.1k symbols; the bulk HF data takes the parallel ANS path (k_ans.cu).
#pragma once

#include "bitio.cuh"
#include "common.cuh"

namespace hydb {

constexpr int kLitBins = 256;                 // literal tokens 0..255
constexpr int kLzBins = 128;                  // lz77 length tokens min_symbol .. min_symbol+127
constexpr int kBins = kLitBins + kLzBins;     // cluster-0 bins
constexpr int kDistBins = 2;                  // cluster-1 (lz77 distance) tokens 0..1
constexpr int kAllBins = kBins + kDistBins;
constexpr int kMaxLeaves = kBins;
constexpr int kMaxNodes = 2 * kMaxLeaves;

struct PrefixParams {
    uint32_t num_plain_dists;   // contexts, all clustered to 0 (the lz77 context is added on top)
    uint32_t lz_min_symbol;     // 0: run-length mode off
    uint32_t modular;           // lz77 distance symbol value (1 for modular streams, entropy.c:486)
    uint8_t split0, msb0, lsb0; // hybrid config of cluster 0
    uint8_t split1, msb1, lsb1; // hybrid config of the lz77 distance cluster
};

// prefix symbol record: token:15 | cluster:1 | nbits:4 | residue:12
HD uint32_t ps_pack(uint32_t token, uint32_t cluster, uint32_t nbits, uint32_t residue) {
    return token | (cluster << 15) | (nbits << 16) | (residue << 20);
}

// scratch for one stream; lives in shared memory on the device (~15 KB)
struct PrefixWork {
    uint32_t freq[kAllBins];
    uint8_t len[kAllBins];
    uint16_t code[kAllBins];        // bit-reversed canonical code
    // tree builder
    uint32_t weight[kMaxNodes];
    int32_t symp1[kMaxNodes];       // dense token + 1 for leaves, 0 for internal nodes
    uint16_t pos[kMaxNodes];        // slot position in the reference's array
    uint8_t reach[kMaxNodes];       // deepest leaf depth below the node
    int16_t parent[kMaxNodes];
    uint16_t live[kMaxLeaves];
    uint16_t leaf_bin[kMaxLeaves];
    uint32_t nsyms;
    uint32_t alpha0, alpha1;
    uint32_t error;
};

HD uint32_t ps_bin(uint32_t token, uint32_t cluster, uint32_t lz_min) {
    if (cluster)
        return kBins + token;
    if (lz_min && token >= lz_min)
        return kLitBins + (token - lz_min);
    return token;
}
HD uint32_t ps_bin_token(uint32_t bin, uint32_t lz_min) {
    if (bin >= (uint32_t)kBins)
        return bin - kBins;
    return bin < (uint32_t)kLitBins ? bin : lz_min + (bin - kLitBins);
}

// ---- front end: values -> hybrid symbols with run-length chunks -------------------------
// (reference: entropy.c:466-524, flush at header time entropy.c:549-551).
// A maximal run of equal values is cut into chunks of <= 128; each chunk is one literal plus r
// repeats, coded as r more literals (r <= 3) or a length token + a distance symbol (r > 3).
HDN inline void ps_emit(PrefixWork &w, uint32_t *syms, uint32_t cap, uint32_t token, uint32_t cluster,
                        uint32_t nbits, uint32_t residue, uint32_t lz_min) {
    bool fits;
    if (cluster)
        fits = token < (uint32_t)kDistBins;
    else if (lz_min && token >= lz_min)
        fits = token - lz_min < (uint32_t)kLzBins;
    else
        fits = token < (uint32_t)kLitBins;
    if (!fits || nbits > 12 || token >= (1u << 15)) {
        w.error |= kErrLfAlphabet;
        return;
    }
    if (w.nsyms >= cap) {
        w.error |= kErrLfCapacity;
        return;
    }
    syms[w.nsyms++] = ps_pack(token, cluster, nbits, residue);
    w.freq[ps_bin(token, cluster, lz_min)]++;
    if (cluster) {
        if (token + 1 > w.alpha1) w.alpha1 = token + 1;
    } else {
        if (token + 1 > w.alpha0) w.alpha0 = token + 1;
    }
}

HDN inline void ps_literal(PrefixWork &w, uint32_t *syms, uint32_t cap, const PrefixParams &p, uint32_t v) {
    uint32_t res, nb;
    const uint32_t tok = hybrid_token(v, p.split0, p.msb0, p.lsb0, res, nb);
    ps_emit(w, syms, cap, tok, 0, nb, res, p.lz_min_symbol);
}

template <typename ValueAt>
HDN inline void ps_tokenize(PrefixWork &w, uint32_t *syms, uint32_t cap, const PrefixParams &p,
                            uint32_t n, ValueAt value_at) {
    for (int i = 0; i < kAllBins; i++)
        w.freq[i] = 0;
    w.nsyms = 0;
    w.alpha0 = w.alpha1 = 0;
    uint32_t i = 0;
    while (i < n) {
        const uint32_t v = value_at(i);
        uint32_t run = 1;
        if (p.lz_min_symbol) {
            while (run < 128 && i + run < n && value_at(i + run) == v)
                run++;
        }
        ps_literal(w, syms, cap, p, v);
        const uint32_t rep = run - 1;
        if (rep > 3) {
            // length token (config 7,0,0 => plain value, entropy.c:40, 478-483) + distance symbol
            ps_emit(w, syms, cap, p.lz_min_symbol + (rep - 3), 0, 0, 0, p.lz_min_symbol);
            uint32_t res, nb;
            const uint32_t tok = hybrid_token(p.modular ? 1u : 0u, p.split1, p.msb1, p.lsb1, res, nb);
            ps_emit(w, syms, cap, tok, 1, nb, res, p.lz_min_symbol);
        } else {
            for (uint32_t k = 0; k < rep; k++)
                ps_literal(w, syms, cap, p, v);
        }
        i += run;
    }
}

// the same front end for a value sequence given as (value, length) runs of DISTINCT neighbouring
// values: O(symbols) instead of O(values) for the long constant stretches of the HF-metadata image
struct PsRun {
    uint32_t value, length;
};
HDN inline void ps_tokenize_runs(PrefixWork &w, uint32_t *syms, uint32_t cap, const PrefixParams &p,
                                 const PsRun *runs, uint32_t nruns) {
    for (int i = 0; i < kAllBins; i++)
        w.freq[i] = 0;
    w.nsyms = 0;
    w.alpha0 = w.alpha1 = 0;
    for (uint32_t k = 0; k < nruns; k++) {
        uint32_t left = runs[k].length;
        const uint32_t v = runs[k].value;
        while (left) {
            const uint32_t run = (p.lz_min_symbol && left > 128) ? 128 : (p.lz_min_symbol ? left : 1);
            ps_literal(w, syms, cap, p, v);
            const uint32_t rep = run - 1;
            if (rep > 3) {
                ps_emit(w, syms, cap, p.lz_min_symbol + (rep - 3), 0, 0, 0, p.lz_min_symbol);
                uint32_t res, nb;
                const uint32_t tok = hybrid_token(p.modular ? 1u : 0u, p.split1, p.msb1, p.lsb1, res, nb);
                ps_emit(w, syms, cap, tok, 1, nb, res, p.lz_min_symbol);
            } else {
                for (uint32_t j = 0; j < rep; j++)
                    ps_literal(w, syms, cap, p, v);
            }
            left -= run;
        }
    }
}

// ---- length-limited code lengths on live nodes only ---------------------------------------
// Simulates reference entropy.c:592-662: pass k takes the two cheapest eligible nodes from array
// slots [2k, n+k), swaps them into slots 2k / 2k+1 and parks the parent in slot n+k.  Order:
// weight, leaves before internal nodes, leaves by token, and between equal-weight internal nodes
// the one in the HIGHER slot first (the reference's scan lets a later candidate displace an
// internal incumbent on ties, entropy.c:577-581, 627-632).
HD bool ps_node_before(const PrefixWork &w, int a, int b) {
    if (w.weight[a] != w.weight[b])
        return w.weight[a] < w.weight[b];
    const int32_t sa = w.symp1[a], sb = w.symp1[b];
    if (sa && sb)
        return sa < sb;
    if (sa || sb)
        return sa != 0;
    return w.pos[a] > w.pos[b];
}

// leaves: bins first_bin .. first_bin+nbins-1 with non-zero freq; `alphabet` = dense alphabet size.
// Writes w.len[] for those bins.  Returns number of leaves.
HDN inline uint32_t ps_code_lengths(PrefixWork &w, uint32_t first_bin, uint32_t nbins, uint32_t alphabet,
                                    int limit, uint32_t lz_min, bool dist_cluster) {
    uint32_t nz = 0;
    for (uint32_t b = first_bin; b < first_bin + nbins; b++) {
        w.len[b] = 0;
        if (!w.freq[b])
            continue;
        const uint32_t tok = dist_cluster ? b - first_bin : ps_bin_token(b, lz_min);
        w.weight[nz] = w.freq[b];
        w.symp1[nz] = (int32_t)tok + 1;
        w.pos[nz] = (uint16_t)tok;
        w.reach[nz] = 0;
        w.parent[nz] = -1;
        w.live[nz] = (uint16_t)nz;
        w.leaf_bin[nz] = (uint16_t)b;
        nz++;
    }
    if (!nz) {
        w.error |= kErrHuffman;   // reference: "No nonzero frequencies"
        return 0;
    }
    uint32_t nlive = nz, nnodes = nz;
    for (uint32_t k = 0; k + 1 < alphabet; k++) {
        const int bound = limit - ceil_log2_u32(nlive) + 1;
        int best = -1, next = -1, at0 = -1, at1 = -1;
        int best_i = -1, next_i = -1;
        for (uint32_t i = 0; i < nlive; i++) {
            const int id = w.live[i];
            if (w.pos[id] == 2 * k) at0 = id;
            else if (w.pos[id] == 2 * k + 1) at1 = id;
            if ((int)w.reach[id] >= bound)
                continue;
            if (best < 0 || ps_node_before(w, id, best)) {
                next = best; next_i = best_i;
                best = id; best_i = (int)i;
            } else if (next < 0 || ps_node_before(w, id, next)) {
                next = id; next_i = (int)i;
            }
        }
        if (best < 0) {
            w.error |= kErrHuffman;   // reference: "couldn't find target"
            break;
        }
        const uint16_t ps = w.pos[best];
        if (at0 >= 0 && at0 != best)
            w.pos[at0] = ps;
        w.pos[best] = (uint16_t)(2 * k);
        if (next < 0)
            break;
        const int y = (ps == 2 * k + 1) ? (at0 != best ? at0 : -1) : (at1 != best ? at1 : -1);
        const uint16_t pt = w.pos[next];
        if (y >= 0 && y != next)
            w.pos[y] = pt;
        w.pos[next] = (uint16_t)(2 * k + 1);
        const int id = (int)nnodes++;
        w.weight[id] = w.weight[best] + w.weight[next];
        w.symp1[id] = 0;
        w.pos[id] = (uint16_t)(alphabet + k);
        w.reach[id] = (uint8_t)(1 + (w.reach[best] > w.reach[next] ? w.reach[best] : w.reach[next]));
        w.parent[id] = -1;
        w.parent[best] = (int16_t)id;
        w.parent[next] = (int16_t)id;
        // drop best/next from the live list, add the parent
        const int hi = best_i > next_i ? best_i : next_i, lo = best_i > next_i ? next_i : best_i;
        w.live[hi] = w.live[nlive - 1];
        nlive--;
        w.live[lo] = (uint16_t)id;
    }
    for (uint32_t i = 0; i < nz; i++) {
        uint32_t d = 0;
        for (int j = w.parent[i]; j >= 0; j = w.parent[j])
            d++;
        w.len[w.leaf_bin[i]] = (uint8_t)d;
    }
    return nz;
}

HD uint32_t ps_reverse_bits(uint32_t v, int n) {
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++)
        r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// canonical codes by (length, token), bit-reversed (reference: entropy.c:664-707).
// Bins are ordered by token, so a scan per length reproduces the reference's stable sort.
HDN inline void ps_assign_codes(PrefixWork &w, uint32_t first_bin, uint32_t nbins) {
    uint32_t next = 0;   // code value left-aligned in 16 bits
    uint32_t total = 0;
    for (int l = 1; l <= 15; l++) {
        for (uint32_t b = first_bin; b < first_bin + nbins; b++) {
            if (w.len[b] != l)
                continue;
            w.code[b] = (uint16_t)ps_reverse_bits(next >> (16 - l), l);
            next += 1u << (16 - l);
            total += 1u << (16 - l);
        }
    }
    if (total && total != (1u << 16))
        w.error |= kErrHuffman;   // reference: "VLC codes do not add up"
}

// ---- code-length code ("complex" prefix header, reference: entropy.c:709-805) --------------
struct ClcWork {
    uint32_t freq[18];
    uint8_t len[18];
    uint16_t code[18];
};

HD void ps_zero_run_count(uint32_t zeros, uint32_t *freq) {
    if (zeros >= 3) {
        while (zeros > 10) {
            freq[17]++;
            zeros = (zeros + 13) / 8;
        }
        freq[17]++;
    } else {
        freq[0] += zeros;
    }
}

HDN inline void ps_zero_run_put(BitSink &bw, const ClcWork &c, uint32_t zeros) {
    if (zeros >= 3) {
        uint32_t part[8];
        int k = 0;
        while (zeros > 10) {
            const uint32_t up = (zeros + 13) / 8;
            part[k++] = zeros - 8 * up + 16;
            zeros = up;
        }
        part[k++] = zeros;
        while (k--) {
            bw.put(c.code[17], c.len[17]);
            bw.put(part[k] - 3, 3);
        }
    } else {
        while (zeros--)
            bw.put(c.code[0], c.len[0]);
    }
}

// small dense version of the tree builder for the 18-symbol code-length alphabet, reusing the
// same scratch (positions are dense there, but the sparse routine handles that too)
HDN inline void ps_clc_lengths(PrefixWork &w, ClcWork &c) {
    // borrow node arrays; leaves indexed by their token
    uint32_t nz = 0;
    uint16_t leaf_tok[18];
    for (int t = 0; t < 18; t++) {
        c.len[t] = 0;
        if (!c.freq[t])
            continue;
        w.weight[nz] = c.freq[t];
        w.symp1[nz] = t + 1;
        w.pos[nz] = (uint16_t)t;
        w.reach[nz] = 0;
        w.parent[nz] = -1;
        w.live[nz] = (uint16_t)nz;
        leaf_tok[nz] = (uint16_t)t;
        nz++;
    }
    if (!nz) {
        w.error |= kErrHuffman;
        return;
    }
    uint32_t nlive = nz, nnodes = nz;
    for (uint32_t k = 0; k + 1 < 18; k++) {
        const int bound = 5 - ceil_log2_u32(nlive) + 1;
        int best = -1, next = -1, at0 = -1, at1 = -1, best_i = -1, next_i = -1;
        for (uint32_t i = 0; i < nlive; i++) {
            const int id = w.live[i];
            if (w.pos[id] == 2 * k) at0 = id;
            else if (w.pos[id] == 2 * k + 1) at1 = id;
            if ((int)w.reach[id] >= bound)
                continue;
            if (best < 0 || ps_node_before(w, id, best)) {
                next = best; next_i = best_i;
                best = id; best_i = (int)i;
            } else if (next < 0 || ps_node_before(w, id, next)) {
                next = id; next_i = (int)i;
            }
        }
        if (best < 0) {
            w.error |= kErrHuffman;
            break;
        }
        const uint16_t ps = w.pos[best];
        if (at0 >= 0 && at0 != best)
            w.pos[at0] = ps;
        w.pos[best] = (uint16_t)(2 * k);
        if (next < 0)
            break;
        const int y = (ps == 2 * k + 1) ? (at0 != best ? at0 : -1) : (at1 != best ? at1 : -1);
        const uint16_t pt = w.pos[next];
        if (y >= 0 && y != next)
            w.pos[y] = pt;
        w.pos[next] = (uint16_t)(2 * k + 1);
        const int id = (int)nnodes++;
        w.weight[id] = w.weight[best] + w.weight[next];
        w.symp1[id] = 0;
        w.pos[id] = (uint16_t)(18 + k);
        w.reach[id] = (uint8_t)(1 + (w.reach[best] > w.reach[next] ? w.reach[best] : w.reach[next]));
        w.parent[id] = -1;
        w.parent[best] = (int16_t)id;
        w.parent[next] = (int16_t)id;
        const int hi = best_i > next_i ? best_i : next_i, lo = best_i > next_i ? next_i : best_i;
        w.live[hi] = w.live[nlive - 1];
        nlive--;
        w.live[lo] = (uint16_t)id;
    }
    for (uint32_t i = 0; i < nz; i++) {
        uint32_t d = 0;
        for (int j = w.parent[i]; j >= 0; j = w.parent[j])
            d++;
        c.len[leaf_tok[i]] = (uint8_t)d;
    }
    uint32_t next = 0, total = 0;
    for (int l = 1; l <= 5; l++)
        for (int t = 0; t < 18; t++) {
            if (c.len[t] != l)
                continue;
            c.code[t] = (uint16_t)ps_reverse_bits(next >> (16 - l), l);
            next += 1u << (16 - l);
            total += 1u << (16 - l);
        }
    if (total && total != (1u << 16))
        w.error |= kErrHuffman;
}

// `dense`: the cluster's tokens are first_bin-relative bin numbers (distance cluster, ICC clusters)
HDN inline void ps_put_complex_code(PrefixWork &w, BitSink &bw, uint32_t first_bin, uint32_t nbins,
                                    uint32_t alphabet, uint32_t lz_min, bool dense = false) {
    static const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};  // entropy.c:42
    static const uint8_t kL0Bits[6] = {0, 7, 3, 2, 1, 15};                                             // entropy.c:44-46
    static const uint8_t kL0Len[6] = {2, 4, 3, 2, 2, 4};
    ClcWork c;
    bw.put(0, 2);   // hskip = 0
    for (int t = 0; t < 18; t++)
        c.freq[t] = 0, c.len[t] = 0, c.code[t] = 0;
    // pass 1: code-length statistics over the dense alphabet, zeros as gaps between used bins
    uint32_t prev_tok_p1 = 0;   // token after the last non-zero-length symbol
    for (uint32_t b = first_bin; b < first_bin + nbins; b++) {
        if (!w.len[b])
            continue;
        const uint32_t tok = dense ? b - first_bin : ps_bin_token(b, lz_min);
        ps_zero_run_count(tok - prev_tok_p1, c.freq);
        c.freq[w.len[b]]++;
        prev_tok_p1 = tok + 1;
    }
    ps_clc_lengths(w, c);
    uint32_t space = 0;
    for (int j = 0; j < 18; j++) {
        const uint32_t l = c.len[kOrder[j]];
        bw.put(kL0Bits[l], kL0Len[l]);
        if (l)
            space += 32u >> l;
        if (space >= 32)
            break;
    }
    if (space && space != 32)
        w.error |= kErrHuffman;   // reference: "level1 code total mismatch"
    // pass 2: the lengths themselves
    space = 0;
    prev_tok_p1 = 0;
    bool done = false;
    for (uint32_t b = first_bin; b < first_bin + nbins; b++) {
        if (!w.len[b])
            continue;
        const uint32_t tok = dense ? b - first_bin : ps_bin_token(b, lz_min);
        ps_zero_run_put(bw, c, tok - prev_tok_p1);
        bw.put(c.code[w.len[b]], c.len[w.len[b]]);
        prev_tok_p1 = tok + 1;
        space += 32768u >> w.len[b];
        if (space == 32768) {
            done = true;
            break;
        }
    }
    if (!done)
        ps_zero_run_put(bw, c, alphabet - prev_tok_p1);
}

// per-cluster code description + canonical code assignment (reference: entropy.c:846-927)
// `lengths_ready`: w.len[] of the cluster's bins was filled in already (warp_code_lengths)
HDN inline void ps_put_cluster_code(PrefixWork &w, BitSink &bw, uint32_t first_bin, uint32_t nbins,
                                    uint32_t alphabet, uint32_t lz_min, bool dist_cluster, bool lengths_ready = false) {
    if (!lengths_ready)
        ps_code_lengths(w, first_bin, nbins, alphabet, 15, lz_min, dist_cluster);
    uint32_t used = 0;
    uint32_t fsym[4] = {0, 0, 0, 0}, flen[4] = {0, 0, 0, 0};
    for (uint32_t b = first_bin; b < first_bin + nbins; b++) {
        if (!w.len[b])
            continue;
        if (used < 4) {
            fsym[used] = dist_cluster ? b - first_bin : ps_bin_token(b, lz_min);
            flen[used] = w.len[b];
        }
        if (++used > 4)
            break;
    }
    if (used > 4) {
        ps_put_complex_code(w, bw, first_bin, nbins, alphabet, lz_min, dist_cluster);
        ps_assign_codes(w, first_bin, nbins);
        return;
    }
    if (!used) {
        used = 1;
        fsym[0] = alphabet - 1;
    }
    bw.put(1, 2);           // hskip = 1: simple code
    bw.put(used - 1, 2);
#define HYDB_SWAP_FEW(a, b) do { uint32_t ts = fsym[a], tl = flen[a]; fsym[a] = fsym[b]; flen[a] = flen[b]; \
                                 fsym[b] = ts; flen[b] = tl; } while (0)
    if (used == 3 && flen[0] != 1) {
        if (flen[1] == 1) HYDB_SWAP_FEW(0, 1); else HYDB_SWAP_FEW(0, 2);
    }
    int select = 0;
    if (used == 4) {
        for (int i = 0; i < 4; i++)
            if (flen[i] != 2) { select = 1; break; }
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
    const int width = ceil_log2_u32(alphabet);
    for (uint32_t i = 0; i < used; i++)
        bw.put(fsym[i], width);
    if (used == 4)
        bw.put_bool(select);
    ps_assign_codes(w, first_bin, nbins);
}

HD void ps_put_hybrid_cfg(BitSink &bw, int split, int msb, int lsb, int log_alpha) {   // entropy.c:169-182
    bw.put((uint32_t)split, ceil_log2_u32(1u + (uint32_t)log_alpha));
    if (split == log_alpha)
        return;
    bw.put((uint32_t)msb, ceil_log2_u32(1u + (uint32_t)split));
    bw.put((uint32_t)lsb, ceil_log2_u32(1u + (uint32_t)(split - msb)));
}

// stream preamble + code descriptions (reference: entropy.c:546-575, 807-931).  After this,
// w.code/w.len hold the per-bin codes for ps_put_symbols().
HDN inline void ps_put_header(PrefixWork &w, BitSink &bw, const PrefixParams &p, bool lengths0_ready = false) {
    const U32Dist kMinSymbol = {{224, 512, 4096, 8}, {0, 0, 0, 15}};    // entropy.c:48-51
    const U32Dist kMinLength = {{3, 4, 5, 9}, {0, 0, 2, 8}};            // entropy.c:52-55
    const uint32_t lz = p.lz_min_symbol;
    bw.put_bool(lz != 0);
    if (lz) {
        put_u32(bw, kMinSymbol, lz);
        put_u32(bw, kMinLength, 3);
        ps_put_hybrid_cfg(bw, 7, 0, 0, 8);
    }
    const uint32_t num_dists = p.num_plain_dists + (lz ? 1u : 0u);
    const uint32_t num_clusters = lz ? 2u : 1u;
    if (num_dists != 1) {   // simple cluster map (entropy.c:114-123); nbits is 0 or 1 here
        const int nbits = ceil_log2_u32(num_clusters);
        bw.put_bool(1);
        bw.put((uint32_t)nbits, 2);
        for (uint32_t i = 0; i < num_dists; i++)
            bw.put(lz && i == num_dists - 1 ? 1u : 0u, nbits);
    }
    bw.put_bool(1);   // use prefix codes
    ps_put_hybrid_cfg(bw, p.split0, p.msb0, p.lsb0, 15);
    if (lz)
        ps_put_hybrid_cfg(bw, p.split1, p.msb1, p.lsb1, 15);
    // alphabet sizes (entropy.c:835-844)
    const uint32_t alpha[2] = {w.alpha0, w.alpha1};
    for (uint32_t c = 0; c < num_clusters; c++) {
        if (alpha[c] <= 1) {
            bw.put_bool(0);
            continue;
        }
        bw.put_bool(1);
        const int nb = floor_log2_u32(alpha[c] - 1);
        bw.put((uint32_t)nb, 4);
        bw.put(alpha[c] - 1, nb);
    }
    for (int b = lengths0_ready ? kBins : 0; b < kAllBins; b++)
        w.len[b] = 0, w.code[b] = 0;
    if (w.alpha0 > 1)
        ps_put_cluster_code(w, bw, 0, kBins, w.alpha0, lz, false, lengths0_ready);
    if (lz && w.alpha1 > 1)
        ps_put_cluster_code(w, bw, kBins, kDistBins, w.alpha1, lz, true);
}

// one symbol's code + residue bits (reference: entropy.c:1012-1018)
HD void ps_put_symbol(const PrefixWork &w, BitSink &bw, uint32_t sym, uint32_t lz_min) {
    const uint32_t token = sym & 0x7FFFu, cluster = (sym >> 15) & 1u;
    const uint32_t nbits = (sym >> 16) & 0xFu, residue = sym >> 20;
    const uint32_t b = ps_bin(token, cluster, lz_min);
    bw.put(w.code[b], w.len[b]);
    bw.put(residue, (int)nbits);
}

// whole stream, sequentially: tokenise, header, symbols
template <typename ValueAt>
HDN inline void ps_encode_stream(PrefixWork &w, uint32_t *syms, uint32_t cap, const PrefixParams &p,
                                 uint32_t n, ValueAt value_at, BitSink &bw) {
    ps_tokenize(w, syms, cap, p, n, value_at);
    ps_put_header(w, bw, p);
    for (uint32_t i = 0; i < w.nsyms; i++)
        ps_put_symbol(w, bw, syms[i], p.lz_min_symbol);
}

}  // namespace hydb
