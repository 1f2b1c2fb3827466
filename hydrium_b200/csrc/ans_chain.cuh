// hydrium_b200/csrc/ans_chain.cuh
//
// One step of the reverse rANS state chain (reference: entropy.c:1087-1120), arranged so that the
// only operations between two consecutive slot-table loads are ONE multiply-high and ONE
// multiply-add.  The reference's step for a symbol with frequency f is
//
//     x   state after renormalisation for this symbol                       (2^16 <= x < f * 2^20)
//     q = x / f,  r = x % f,  slot = inv[base + r],  s' = (q << 12) | slot
//     renormalise for the NEXT symbol iff (s' >> 20) >= f_next  <=>  q >= f_next << 8:
//     the 16-bit word s' & 0xFFFF is emitted and the state becomes s' >> 16 = q >> 4.
//
// The state entering a step is x = a + v with
//     a = q_prev << 12 (no renormalisation) or q_prev >> 4 (renormalised)    known BEFORE the previous
//                                                                            step's table load returns
//     v = slot_prev (no renormalisation) or 0                                the load result.
// Division by multiplication, exact (round 2, second form).  M = ceil(2^32 / f) - 1 (for every f >= 1: the
// floor for a non-power of two, one less than the quotient for a power of two, 2^32 - 1 for f = 1) and
// e = 2^32 - f * M, so 1 <= e <= f and f * M + e = 2^32.  For 0 <= y < 2^14
//     hi32((y + 1) * M) = y / f :   (y + 1) * M / 2^32 = (y + 1) / f - d  with  0 < d = (y + 1) e / (f 2^32) <= 2^-18,
//     which lies in [k, k + 1) for y = k f + r because (r + 1) / f >= 2^-12 > d, and strictly below k + 1 when
//     r = f - 1 because d > 0.
// While the previous load is in flight the "shadow" computes, for the coming symbol,
//     w  = (a + 1) * M (64 bit),  qa = hi32(w)       qa in {a/f - 1, a/f}: (a + 1) M / 2^32 = (a + 1)/f - (a + 1) e / (f 2^32)
//                                                    and the last term is at most 1; so y = x - qa f < 2f + 4096 <= 2^14 - 1
//     R  = w + qa * e = (a + 1 - qa f) * M + qa * 2^32,       C0 = base2 + 2 * a
// and once v arrives the dependent path is
//     q    = hi32(v * M + R) = qa + hi32((y + 1) * M) = qa + y / f = x / f       multiply-high, 64-bit addend
//     addr = C0 + 2 * v - 2 * f * q  = base2 + 2 * (x - q * f)                   multiply-add -> next load
// When the previous step renormalised, v must not count: the shadow replaces M by 0 and the factor
// 2 by 0, so the two path instructions are the same in both cases.
// (The first form used mc = ceil(2^32 / f) with a signed correction R = a mc - (hi32(a mc) - 1) e: the "- 1"
// was one more dependent instruction between the two wide multiplies of the state recurrence.)
// tests/test_host_logic.py checks the identity exhaustively at the range boundaries and the whole
// formulation against the oracle.
//
// Why this shape: measured on B200 (tools/ubench), a single warp pays ~8 issue cycles per IMAD.WIDE,
// ~6 per IMAD.HI, ~8 per LDS.128 and ~2.6 per IMAD, so the step is bound by instruction issue as
// much as by latency; this form needs two wide multiplies, one multiply-high, two multiply-adds and
// one 16-byte record per symbol.
#pragma once

#include "ans_model.cuh"

namespace hydb {

constexpr uint32_t kAnsInitState = 0x130000u;   // reference: entropy.c:1083
// per (cluster, token) constants for the chain: also the 16-byte record staged per symbol
struct AnsSymInfo {
    uint32_t mc;    // M = ceil(2^32 / f) - 1; 0 for an unused symbol
    uint32_t ne;    // e = 2^32 - f * M, 1 <= e <= f
    uint32_t nf2;   // -2f (mod 2^32)
    uint32_t b2;    // byte offset of the symbol's first slot in the flat uint16 inverse table
};
// `base` = cluster * 4096 + cumulative frequency
HD AnsSymInfo ans_sym_info(uint32_t f, uint32_t base) {
    AnsSymInfo s;
    s.mc = s.ne = s.nf2 = s.b2 = 0;
    if (!f)
        return s;
    s.mc = (uint32_t)(((1ull << 32) + f - 1u) / f - 1u);
    s.ne = (0u - f) * s.mc;
    s.nf2 = 0u - 2u * f;
    s.b2 = 2u * base;
    return s;
}
HD uint32_t asi_freq(const AnsSymInfo &s) { return (0u - s.nf2) >> 1; }
// renormalisation threshold f << 8 of the symbol coded NEXT
HD uint32_t asi_threshold(const AnsSymInfo &s) { return (0u - s.nf2) << 7; }
constexpr uint32_t kAnsNoNext = 0xFFFFFFFFu;   // threshold for "no further symbol": never renormalises

// The live values carried from step to step.
struct AnsCarry {
    uint32_t v;        // previous table load result
    uint32_t meff;     // mc of the symbol being coded, or 0 when v must not count
    uint32_t k;        // 2, or 0 when v must not count
    uint32_t c0;       // table byte address of "remainder" a
    uint64_t R;        // a * mc - qa * e
};

HD uint32_t ans_hi32(uint64_t w) {
#if defined(__CUDA_ARCH__)
    uint32_t h;   // taken as a 32-bit value so the signed multiply below stays a single IMAD.WIDE
    asm("{ .reg .b32 lo; mov.b64 {lo, %0}, %1; }" : "=r"(h) : "l"(w));
    return h;
#else
    return (uint32_t)(w >> 32);
#endif
}

// shadow part: given the state's known part `a` and the record of the symbol to code
HD void ans_prepare(AnsCarry &c, uint32_t a, bool v_counts, uint32_t mc, uint32_t ne, uint32_t b2) {
    c.meff = v_counts ? mc : 0u;
    c.k = v_counts ? 2u : 0u;
    const uint64_t w = (uint64_t)a * mc + mc;
    const uint32_t qa = ans_hi32(w);
    c.c0 = 2u * a + b2;
    c.R = w + (uint64_t)qa * ne;
}

// set-up before the first step: the initial state renormalised for the first coded symbol
// (reference: entropy.c:1083, 1092-1100).  `table_addr` is added to every table offset.
HD void ans_chain_begin(AnsCarry &c, const AnsSymInfo &first, uint32_t table_addr) {
    const uint32_t f = asi_freq(first);
    const uint32_t x = ((kAnsInitState >> 20) >= f) ? (kAnsInitState >> 16) : kAnsInitState;
    c.v = 0;
    ans_prepare(c, x, false, first.mc, first.ne, first.b2 + table_addr);
}

// Code one symbol (`own`); `next` is the symbol coded in the following step (the previous one in
// stream order) or NULL.  `lookup(byte_address)` reads the uint16 slot table.  Leaves in `s_out`
// the state s' BEFORE the next renormalisation; flag and word are recovered later, off the
// dependent path, by whoever knows the next symbol's frequency f:  flag = (s' >> 20) >= f,
// word = s' & 0xFFFF  (entropy.c:1092-1100).  After the last step s_out is the final state.
// (k_ans_chain spells the same sequence out by hand to control the instruction order.)
template <typename Lookup>
HD void ans_step(AnsCarry &c, const AnsSymInfo &own, const AnsSymInfo *next, uint32_t table_addr, Lookup lookup,
                 uint32_t &s_out) {
    const uint32_t q = ans_hi32((uint64_t)c.v * c.meff + c.R);
    const uint32_t slot = lookup(q * own.nf2 + (c.v * c.k + c.c0));
    const bool p = next && q >= asi_threshold(*next);   // renormalise for the next symbol?
    const uint32_t q12 = q << 12;
    const uint32_t a = p ? (q >> 4) : q12;              // s' >> 16 == q >> 4 (slot < 4096)
    if (next)
        ans_prepare(c, a, !p, next->mc, next->ne, next->b2 + table_addr);
    c.v = slot;
    s_out = q12 | slot;
}

// ---- the same step over the COMPACT inverse map (ans_model.cuh) ---------------------------------------
// Element units instead of table byte addresses: the step yields  g = cum + x % f  and the caller turns
// g into the slot (a warp-wide search over the cluster's sorted pieces).  Same quotient, same exactness
// argument; k is 1 / 0 instead of 2 / 0 and c0 = a + cum.
struct AnsRecC {
    uint32_t mc;    // M, as AnsSymInfo
    uint32_t ne;    // e
    uint32_t nf;    // -f (mod 2^32)
    uint32_t cum;   // cumulative frequency of the symbol inside its cluster
};
HD AnsRecC ans_rec_c(uint32_t f, uint32_t cum) {
    const AnsSymInfo s = ans_sym_info(f, 0);
    AnsRecC r;
    r.mc = s.mc;
    r.ne = s.ne;
    r.nf = f ? 0u - f : 0u;
    r.cum = cum;
    return r;
}
HD void ans_prepare_c(AnsCarry &c, uint32_t a, bool v_counts, uint32_t mc, uint32_t ne, uint32_t cum) {
    c.meff = v_counts ? mc : 0u;
    c.k = v_counts ? 1u : 0u;
    const uint64_t w = (uint64_t)a * mc + mc;
    const uint32_t qa = ans_hi32(w);
    c.c0 = a + cum;
    c.R = w + (uint64_t)qa * ne;
}
HD void ans_chain_begin_c(AnsCarry &c, const AnsRecC &first) {
    const uint32_t f = 0u - first.nf;
    const uint32_t x = ((kAnsInitState >> 20) >= f) ? (kAnsInitState >> 16) : kAnsInitState;
    c.v = 0;
    ans_prepare_c(c, x, false, first.mc, first.ne, first.cum);
}
// `slot_of(g)` maps g = cum + remainder to the alias slot of the symbol's cluster
template <typename SlotOf>
HD void ans_step_c(AnsCarry &c, const AnsRecC &own, const AnsRecC *next, SlotOf slot_of, uint32_t &s_out) {
    const uint32_t q = ans_hi32((uint64_t)c.v * c.meff + c.R);
    const uint32_t slot = slot_of(q * own.nf + (c.v * c.k + c.c0));
    const bool p = next && q >= ((0u - next->nf) << 8);
    const uint32_t q12 = q << 12;
    const uint32_t a = p ? (q >> 4) : q12;
    if (next)
        ans_prepare_c(c, a, !p, next->mc, next->ne, next->cum);
    c.v = slot;
    s_out = q12 | slot;
}

}  // namespace hydb
