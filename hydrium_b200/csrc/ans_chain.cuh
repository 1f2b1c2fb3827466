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

// per (cluster, token) constants for the chain
struct AnsSymInfo {
    uint32_t m;        // reciprocal multiplier
    uint32_t packed;   // (freq - 1):12 | shift:4 << 12 | table base:16 << 16
};
// `base` = cluster * 4096 + cumulative frequency: index of the symbol's first slot in the flat
// inverse table inv[9 * 4096]
HD AnsSymInfo ans_sym_info(uint32_t f, uint32_t base) {
    AnsSymInfo s;
    uint32_t sh;
    if (!f) {
        s.m = 0;
        s.packed = 0;
        return s;
    }
    ans_div_consts(f, s.m, sh);
    s.packed = (f - 1u) | (sh << 12) | (base << 16);
    return s;
}
HD uint32_t asi_freq(uint32_t packed) { return (packed & 0xFFFu) + 1u; }
HD uint32_t asi_shift(uint32_t packed) { return (packed >> 12) & 0xFu; }
HD uint32_t asi_base(uint32_t packed) { return packed >> 16; }

// Code one symbol.  In: x (renormalised state), this symbol's constants, the next symbol's
// frequency (the one that will be coded after this one, i.e. the PREVIOUS symbol in stream
// order; pass 0x7FFFFFFF for "none").  Out: x for the next step, whether that step's
// renormalisation fires, and the 16-bit word it emits.
HD void ans_step(uint32_t &x, uint32_t m, uint32_t packed, const uint16_t *inv,
                 uint32_t f_next, bool &flush, uint32_t &word) {
    const uint32_t f = asi_freq(packed);
    const uint32_t q = ans_div(x, m, asi_shift(packed));
    const uint32_t idx = asi_base(packed) + (x - q * f);
    const uint32_t slot = inv[idx];
    const uint32_t s = (q << 12) | slot;
    flush = (q >> 8) >= f_next;
    word = s & 0xFFFFu;
    x = flush ? (q >> 4) : s;
}

}  // namespace hydb
