// hydrium_b200/csrc/ans_chain.cuh
//
// One step of the reverse rANS state chain (reference: entropy.c:1087-1120), arranged so that
// the dependent path per symbol is  multiply-high -> shift -> multiply-subtract -> one shared
// load -> one logic op:
//
//   x        state after renormalisation for this symbol            (x < f * 2^20)
//   q        = floor(x / f)                 (exact reciprocal, ans_model.cuh)
//   slot     = inv[cum + (x - q*f)]         (inverse alias table)
//   s'       = (q << 12) | slot             state after coding the symbol
//   renormalise for the NEXT symbol iff (s' >> 20) >= f_next  <=>  (q >> 8) >= f_next,
//   which does not wait for the table load; the emitted word is s' & 0xFFFF and the carried
//   state is then s' >> 16 = q >> 4.
#pragma once

#include "ans_model.cuh"

namespace hydb {

constexpr uint32_t kAnsInitState = 0x130000u;   // reference: entropy.c:1083

// per (cluster, token) constants for the chain (16 bytes, one LDS.128)
struct AnsSymInfo {
    uint32_t m;     // reciprocal multiplier (ans_div_consts)
    uint32_t w1;    // shift (0..12) in bits 0..7, frequency in bits 8..
    uint32_t nf2;   // -2 * frequency (mod 2^32)
    uint32_t b2;    // byte offset of the symbol's first slot in the flat uint16 inverse table
};
// `base` = cluster * 4096 + cumulative frequency
HD AnsSymInfo ans_sym_info(uint32_t f, uint32_t base) {
    AnsSymInfo s;
    s.m = 0;
    s.w1 = 0u | (1u << 8);
    s.nf2 = 0;
    s.b2 = 0;
    if (!f)
        return s;
    uint32_t sh;
    ans_div_consts(f, s.m, sh);
    s.w1 = sh | (f << 8);
    s.nf2 = 0u - 2u * f;
    s.b2 = 2u * base;
    return s;
}
HD uint32_t asi_freq(const AnsSymInfo &s) { return s.w1 >> 8; }
constexpr uint32_t kAnsNoNext = 0xFFFFFFu;   // "no further symbol": never triggers a renormalisation

// Code one symbol.  `w1n` = this symbol's shift in bits 0..7 and, in bits 8.., the frequency of
// the symbol that will be coded NEXT (the previous one in stream order; kAnsNoNext for none).
// `lookup(byte_offset)` reads the uint16 slot table.  Out: x for the next step, whether that
// step's renormalisation fires (p), and the 16-bit word it would emit.
template <typename Lookup>
HD void ans_step(uint32_t &x, uint32_t m, uint32_t w1n, uint32_t nf2, uint32_t b2, Lookup lookup,
                 uint32_t &p, uint32_t &word) {
    const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
    const uint32_t q = (uint32_t)(t >> 32) >> (w1n & 31u);   // total shift 32 + sh: only the high word matters
    const uint32_t slot = lookup(q * nf2 + (2u * x + b2));   // 2 * (base + x - q * f), mod 2^32
    p = (q >> 8) >= (w1n >> 8) ? 1u : 0u;
    const uint32_t a = p ? (q >> 4) : (q << 12);
    const uint32_t keep = p ? 0u : 0xFFFFu;
    word = ((q << 12) | slot) & 0xFFFFu;
    x = a | (slot & keep);
}


// Leaner form used by the kernel: returns the state s' BEFORE the next renormalisation instead of
// (flag, word).  Both are recovered later, off the dependent path, by whoever knows the next
// symbol's frequency f:   flag = (s' >> 20) >= f,  word = s' & 0xFFFF   (entropy.c:1092-1100).
// `thr` = (next symbol's frequency << 8) | shift: the renormalisation test
// (q >> 8) >= f_next is evaluated as (q | 0xFF) >= thr, and the shifter uses the low five bits.
template <typename Lookup>
HD void ans_step_state(uint32_t &x, uint32_t m, uint32_t thr, uint32_t nf2, uint32_t b2, Lookup lookup,
                       uint32_t &s_out) {
    const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
    const uint32_t q = (uint32_t)(t >> 32) >> (thr & 31u);   // total shift 32 + sh: only the high word matters
    const uint32_t slot = lookup(q * nf2 + (2u * x + b2));
    const bool p = (q | 0xFFu) >= thr;                       // renormalise for the next symbol?
    const uint32_t a = p ? (q >> 4) : (q << 12);             // s' >> 16 == q >> 4 (slot < 4096)
    const uint32_t keep = p ? 0u : 0xFFFFu;
    s_out = (q << 12) | slot;
    x = a | (slot & keep);
}

}  // namespace hydb
