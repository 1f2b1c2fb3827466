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
    uint32_t packed;   // freq:13 | shift:4 << 13 | cum:12 << 17
};
HD AnsSymInfo ans_sym_info(uint32_t f, uint32_t cum) {
    AnsSymInfo s;
    uint32_t sh;
    if (!f) {
        s.m = 0;
        s.packed = 0;
        return s;
    }
    ans_div_consts(f, s.m, sh);
    s.packed = f | (sh << 13) | (cum << 17);
    return s;
}
HD uint32_t asi_freq(uint32_t packed) { return packed & 0x1FFFu; }
HD uint32_t asi_shift(uint32_t packed) { return (packed >> 13) & 0xFu; }
HD uint32_t asi_cum(uint32_t packed) { return packed >> 17; }

// Code one symbol.  In: x (renormalised state), this symbol's constants, the next symbol's
// frequency (the one that will be coded after this one, i.e. the PREVIOUS symbol in stream
// order; pass 0x7FFFFFFF for "none").  Out: x for the next step, whether that step's
// renormalisation fires, and the 16-bit word it emits.
HD void ans_step(uint32_t &x, uint32_t m, uint32_t packed, const uint16_t *inv_cluster,
                 uint32_t f_next, bool &flush, uint32_t &word) {
    const uint32_t f = asi_freq(packed);
    const uint32_t q = ans_div(x, m, asi_shift(packed));
    const uint32_t idx = asi_cum(packed) + (x - q * f);
    const uint32_t slot = inv_cluster[idx];
    const uint32_t s = (q << 12) | slot;
    flush = (q >> 8) >= f_next;
    word = s & 0xFFFFu;
    x = flush ? (q >> 4) : s;
}

}  // namespace hydb
