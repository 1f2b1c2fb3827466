// hydrium_b200/csrc/bitio.cuh
//
// LSB-first bit sink over 32-bit words (the role of reference bitwriter.c:110-172, without its
// byte cache / realloc machinery: capacity is fixed by the caller and overflow is a flag).
// Used by ONE thread at a time for the small sequential headers; bulk symbol bits are packed in
// parallel elsewhere (see k_ans.cu) and spliced with append_bits().
#pragma once

#include "common.cuh"

namespace hydb {

struct BitSink {
    uint32_t *words;
    uint32_t cap_words;
    uint32_t wpos;
    uint32_t nacc;
    uint64_t acc;
    uint32_t overflow;

    HD void init(uint32_t *w, uint32_t cap) {
        words = w;
        cap_words = cap;
        wpos = 0;
        nacc = 0;
        acc = 0;
        overflow = 0;
    }
    // continue a bit string whose first `bitpos` bits are already in w (the partial last word too)
    HD void resume(uint32_t *w, uint32_t cap, uint32_t bitpos) {
        words = w;
        cap_words = cap;
        wpos = bitpos >> 5;
        nacc = bitpos & 31u;
        acc = nacc && wpos < cap ? (uint64_t)(w[wpos] & ((1u << nacc) - 1u)) : 0;
        overflow = 0;
    }
    // append the low n bits of v, n in [0, 32]
    HD void put(uint32_t v, int n) {
        if (n <= 0)
            return;
        const uint64_t m = n >= 32 ? 0xFFFFFFFFull : ((1ull << n) - 1ull);
        acc |= ((uint64_t)v & m) << nacc;
        nacc += (uint32_t)n;
        if (nacc >= 32) {
            if (wpos < cap_words)
                words[wpos] = (uint32_t)acc;
            else
                overflow = 1;
            wpos++;
            acc >>= 32;
            nacc -= 32;
        }
    }
    HD void put_bool(int f) { put(f ? 1u : 0u, 1); }
    HD uint32_t bitlen() const { return wpos * 32u + nacc; }
    // zero-pad to a byte boundary (reference: bitwriter.c:126-128)
    HD void align_byte() { put(0, (int)((8u - (bitlen() & 7u)) & 7u)); }
    // write out the partial last word (upper bits zero); bitlen() is unchanged
    HD void flush_partial() {
        if (nacc) {
            if (wpos < cap_words)
                words[wpos] = (uint32_t)acc;
            else
                overflow = 1;
        }
    }
};

struct U32Dist {
    uint32_t c[4];
    uint32_t u[4];
};

// JPEG XL U32() field (reference: bitwriter.c:134-142); returns false if unrepresentable
HD bool put_u32(BitSink &bw, const U32Dist &d, uint32_t v) {
    for (int i = 0; i < 4; i++) {
        const uint64_t lim = (1ull << d.u[i]) - 1ull;
        const uint64_t x = (uint64_t)(uint32_t)(v - d.c[i]);
        if (x <= lim) {
            const uint64_t field = (x << 2) | (uint64_t)i;
            const int n = (int)d.u[i] + 2;
            bw.put((uint32_t)field, n > 32 ? 32 : n);
            if (n > 32)
                bw.put((uint32_t)(field >> 32), n - 32);
            return true;
        }
    }
    return false;
}

// JPEG XL U64() field (reference: bitwriter.c:152-172)
HD void put_u64(BitSink &bw, uint64_t v) {
    if (!v) {
        bw.put(0, 2);
        return;
    }
    if (v < 17) {
        bw.put((uint32_t)(((v - 1) << 2) | 1), 6);
        return;
    }
    if (v < 273) {
        bw.put((uint32_t)(((v - 17) << 2) | 2), 10);
        return;
    }
    bw.put((uint32_t)(((v & 0xFFF) << 2) | 3), 14);
    for (int shift = 12;; shift += 8) {
        const uint64_t rest = v >> shift;
        if (!rest) {
            bw.put(0, 1);
            return;
        }
        if (shift == 60) {
            bw.put((uint32_t)(((rest & 0xF) << 1) | 1), 5);
            return;
        }
        bw.put((uint32_t)(((rest & 0xFF) << 1) | 1), 9);
    }
}

// Sequential bit-granular append of `nbits` from src (bit 0 aligned) to the sink.
HD void put_bits_from(BitSink &bw, const uint32_t *src, uint32_t nbits) {
    uint32_t i = 0;
    for (; i + 32 <= nbits; i += 32)
        bw.put(src[i >> 5], 32);
    if (i < nbits)
        bw.put(src[i >> 5], (int)(nbits - i));
}

}  // namespace hydb
