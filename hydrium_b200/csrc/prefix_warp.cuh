// hydrium_b200/csrc/prefix_warp.cuh
//
// Warp-parallel twin of ps_code_lengths (prefix_coder.cuh): the reference's slot-swapping tree builder
// (entropy.c:592-662) on live nodes only, the two cheapest eligible nodes of every pass found by warp
// reductions on a packed 64-bit key.  Device only; used by k_lf_group (single-group frames) and
// k_frame_lf (frames of several groups).
#pragma once

#include "prefix_coder.cuh"

namespace hydb {

constexpr uint32_t kWarpFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint64_t o = __shfl_xor_sync(kWarpFull, v, d);
        v = o < v ? o : v;
    }
    return v;
}

// ordering key of the tree builder: weight, then leaves (by token) before internal nodes, and
// between internal nodes the higher slot first (see ps_node_before)
__device__ __forceinline__ uint64_t node_key(const PrefixWork &w, int id) {
    const uint32_t lo = w.symp1[id] ? (uint32_t)w.symp1[id] : (0x80000000u | (0xFFFFu - w.pos[id]));
    return ((uint64_t)w.weight[id] << 32) | lo;
}

// warp-parallel version of ps_code_lengths for cluster 0 (same results; tests/test_gpu_parity.py
// compares the LF bit strings against the oracle, tests/test_host_logic.py the sequential twin)
static __device__ void warp_code_lengths(PrefixWork &w, uint32_t alphabet, int limit, uint32_t lz_min, uint32_t lane) {
    // gather leaves (bins are few: sequential compaction by ballot)
    uint32_t nz = 0;
    for (uint32_t b0 = 0; b0 < (uint32_t)kBins; b0 += 32) {
        const uint32_t b = b0 + lane;
        const bool used = b < (uint32_t)kBins && w.freq[b] != 0;
        const uint32_t m = __ballot_sync(kWarpFull, used);
        if (b < (uint32_t)kBins)
            w.len[b] = 0;
        if (used) {
            const uint32_t i = nz + __popc(m & ((1u << lane) - 1u));
            const uint32_t tok = ps_bin_token(b, lz_min);
            w.weight[i] = w.freq[b];
            w.symp1[i] = (int32_t)tok + 1;
            w.pos[i] = (uint16_t)tok;
            w.reach[i] = 0;
            w.parent[i] = -1;
            w.live[i] = (uint16_t)i;
            w.leaf_bin[i] = (uint16_t)b;
        }
        nz += __popc(m);
    }
    __syncwarp();
    if (!nz) {
        if (lane == 0)
            w.error |= kErrHuffman;
        return;
    }
    uint32_t nlive = nz, nnodes = nz;
    for (uint32_t k = 0; k + 1 < alphabet; k++) {
        const int bound = limit - ceil_log2_u32(nlive) + 1;
        uint64_t k1 = ~0ull, k2 = ~0ull;   // local best / second
        int i1 = -1, i2 = -1;              // their indices in the live list
        int at0 = -1, at1 = -1;
        for (uint32_t i = lane; i < nlive; i += 32) {
            const int id = w.live[i];
            const uint32_t p = w.pos[id];
            if (p == 2 * k) at0 = id;
            else if (p == 2 * k + 1) at1 = id;
            if ((int)w.reach[id] >= bound)
                continue;
            const uint64_t key = node_key(w, id);
            if (key < k1) {
                k2 = k1; i2 = i1;
                k1 = key; i1 = (int)i;
            } else if (key < k2) {
                k2 = key; i2 = (int)i;
            }
        }
        // who sits in slots 2k / 2k+1 (at most one lane each)
        {
            const uint32_t m0 = __ballot_sync(kWarpFull, at0 >= 0), m1 = __ballot_sync(kWarpFull, at1 >= 0);
            at0 = m0 ? __shfl_sync(kWarpFull, at0, __ffs(m0) - 1) : -1;
            at1 = m1 ? __shfl_sync(kWarpFull, at1, __ffs(m1) - 1) : -1;
        }
        const uint64_t gbest = warp_min_u64(k1);
        if (gbest == ~0ull) {
            if (lane == 0)
                w.error |= kErrHuffman;   // reference: "couldn't find target"
            break;
        }
        const uint32_t owner = __ffs(__ballot_sync(kWarpFull, k1 == gbest)) - 1;
        const int best_i = __shfl_sync(kWarpFull, i1, owner);
        // the owner lane's runner-up competes with everyone else's best
        const uint64_t cand = lane == owner ? k2 : k1;
        const int cand_i = lane == owner ? i2 : i1;
        const uint64_t gnext = warp_min_u64(cand);
        const int best = w.live[best_i];
        int next = -1, next_i = -1;
        if (gnext != ~0ull) {
            const uint32_t owner2 = __ffs(__ballot_sync(kWarpFull, cand == gnext)) - 1;
            next_i = __shfl_sync(kWarpFull, cand_i, owner2);
            next = w.live[next_i];
        }
        __syncwarp();
        if (lane == 0) {
            const uint16_t ps = w.pos[best];
            if (at0 >= 0 && at0 != best)
                w.pos[at0] = ps;
            w.pos[best] = (uint16_t)(2 * k);
            if (next >= 0) {
                const int y = (ps == 2 * k + 1) ? (at0 != best ? at0 : -1) : (at1 != best ? at1 : -1);
                const uint16_t pt = w.pos[next];
                if (y >= 0 && y != next)
                    w.pos[y] = pt;
                w.pos[next] = (uint16_t)(2 * k + 1);
                const int id = (int)nnodes;
                w.weight[id] = w.weight[best] + w.weight[next];
                w.symp1[id] = 0;
                w.pos[id] = (uint16_t)(alphabet + k);
                w.reach[id] = (uint8_t)(1 + (w.reach[best] > w.reach[next] ? w.reach[best] : w.reach[next]));
                w.parent[id] = -1;
                w.parent[best] = (int16_t)id;
                w.parent[next] = (int16_t)id;
                const int hi = best_i > next_i ? best_i : next_i, lo = best_i > next_i ? next_i : best_i;
                w.live[hi] = w.live[nlive - 1];
                w.live[lo] = (uint16_t)id;
            }
        }
        __syncwarp();
        if (next < 0)
            break;
        nnodes++;
        nlive--;
    }
    for (uint32_t i = lane; i < nz; i += 32) {
        uint32_t d = 0;
        for (int j = w.parent[i]; j >= 0; j = w.parent[j])
            d++;
        w.len[w.leaf_bin[i]] = (uint8_t)d;
    }
    __syncwarp();
}

}  // namespace hydb
