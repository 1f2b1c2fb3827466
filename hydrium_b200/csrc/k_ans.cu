// hydrium_b200/csrc/k_ans.cu
//
// Stage 4: everything after tokenisation, two kernels, one CTA per tile each.
//
// k_ans_chain (needs only the HF symbols + histograms):
//   1. ANS model from the tile's histograms: normalisation, alias split, inverse slot table,
//      exact reciprocals                                   (reference: entropy.c:943-978, 184-301)
//   2. section D (ANS stream header tail) -> HBM           (entropy.c:563-572, 303-369, 980-1001)
//   3. the reverse rANS state chain                        (entropy.c:1083-1120)
// k_ans_pack (also needs the LF stream of k_lf_group, which ran concurrently with the chain):
//   4. payload prefix  A | L | B | D  spliced bit-exactly into the tile's output slab
//                                                          (encoder.c:834-843, 959-967, bitwriter.c:80-108)
//   5. forward bit packing of [state][renorm words + residue bits]   (entropy.c:1122-1147)
//   6. frame header + TOC entry                            (encoder.c:327-435, 992-1005)
//
// The chain (3) is the one inherently serial part of the codec: one 32-bit state threads
// through every symbol of the group.  A CHAIN warp runs nothing but the recurrence, every lane
// carrying the same state (the slot-table read is a shared-memory broadcast, nothing diverges);
// a HELPER warp looks up each symbol's 16-byte record {reciprocal, correction, -2f, table address}
// one to four batches ahead, stages it in a shared-memory ring and later turns the states the chain
// leaves behind into renormalisation flags and words; the other warps only build the tables and
// retire.  The 32 steps of a batch are straight-line code whose records roll through registers
// four steps ahead.  ans_chain.cuh has the per-step maths: between two table loads there is one
// multiply-high and one multiply-add; the rest of the step (about fifteen instructions) executes
// in the load's shadow.  Measured with ncu source-level sampling (profiles/r01_chain_source_*.txt),
// the step is bound by in-order issue of those instructions -- wide integer multiplies occupy the
// multiply pipe for ~8 cycles each -- rather than by the load latency.
#include "ans_chain.cuh"
#include "chain_util.cuh"
#include "headers.cuh"
#include "kernels.h"
#include "prefix_coder.cuh"

namespace hydb {

constexpr int kAnsThreads = 256;    // k_ans_chain
constexpr int kPackThreads = 1024;  // k_ans_pack: one thread per ~3 chunks of 32 symbols, so its global loads overlap
constexpr uint32_t kRecAddrBias = 2u;   // see HYDB_ANS_STEP
constexpr uint32_t kSlotBias = 1u;      // added to every entry of the inverse alias table
constexpr int kRing = 4;                 // batches in flight between the helper and the chain warp
constexpr int kBarFull = 1, kBarEmpty = 1 + kRing;   // named barrier ids (0 is __syncthreads)

struct AnsShared {
    uint16_t inv[kHfClusters * kAnsTotal];              // 73,728 B inverse alias table
    uint4 info4[kHfClusters * kHfTokens];               //  9,216 B per-symbol chain constants (AnsSymInfo)
    uint4 stage[kRing][32];                             //  2,048 B ring of staged batch records (AnsSymInfo + table address)
    uint32_t cap[kRing][32];                            //  pre-renormalisation state left by each step
    uint32_t fring[kRing][32];                          //  frequencies of the symbols of each ring slot
    uint32_t chain_warp, ticket;
    AnsCluster cl[kHfClusters];                         //  5,256 B
    uint32_t hist[kHfClusters * kHfTokens];             //  2,304 B
    uint32_t dbits[kDBitsWords];                        //  1,536 B
    uint32_t alpha[kHfClusters];
    uint32_t own_alpha, dhist_off;
    int log_alpha;
    uint32_t dbitlen, err;
};

int ans_encode_smem_bytes() { return (int)sizeof(AnsShared); }

// 32 bits of `src` starting at bit `sb` (may be negative / run past the end); bits outside
// [0, nbits) read as zero
__device__ __forceinline__ uint32_t fetch32(const uint32_t *__restrict__ src, uint32_t nbits, int64_t sb) {
    const int64_t i = sb >> 5;
    const uint32_t r = (uint32_t)(sb & 31);
    const int64_t nw = ((int64_t)nbits + 31) >> 5;
    auto word = [&](int64_t k) -> uint32_t {
        if (k < 0 || k >= nw)
            return 0u;
        uint32_t v = src[k];
        if (k == nw - 1 && (nbits & 31))
            v &= (1u << (nbits & 31)) - 1u;
        return v;
    };
    const uint32_t a = word(i), b = word(i + 1);
    return r ? ((a >> r) | (b << (32 - r))) : a;
}

// CTA-cooperative append of a bit string at an arbitrary bit offset.  Destination words fully
// covered by this string are stored, boundary words are OR-ed (destination pre-zeroed).
__device__ __forceinline__ void append_bits(uint32_t *__restrict__ dst, uint64_t dbit, const uint32_t *__restrict__ src,
                                            uint32_t nbits, uint32_t tid, uint32_t nthreads) {
    if (!nbits)
        return;
    const uint64_t first = dbit >> 5, last = (dbit + nbits - 1) >> 5;
    for (uint64_t w = first + tid; w <= last; w += nthreads) {
        const uint32_t v = fetch32(src, nbits, (int64_t)(w * 32) - (int64_t)dbit);
        const bool full = w * 32 >= dbit && w * 32 + 32 <= dbit + nbits;
        if (full)
            dst[w] = v;
        else if (v)
            atomicOr(&dst[w], v);
    }
}

// block-wide exclusive scan of two values per thread (kPackThreads threads)
__device__ __forceinline__ void block_scan2(uint32_t &a, uint32_t &b, uint32_t *sa, uint32_t *sb, uint32_t tid,
                                            uint32_t &total_a, uint32_t &total_b) {
    const uint32_t lane = tid & 31, warp = tid >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t va = __shfl_up_sync(0xFFFFFFFFu, ia, d), vb = __shfl_up_sync(0xFFFFFFFFu, ib, d);
        if (lane >= (uint32_t)d) {
            ia += va;
            ib += vb;
        }
    }
    if (lane == 31) {
        sa[warp] = ia;
        sb[warp] = ib;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wa = lane < kPackThreads / 32 ? sa[lane] : 0, wb = lane < kPackThreads / 32 ? sb[lane] : 0;
        uint32_t xa = wa, xb = wb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t va = __shfl_up_sync(0xFFFFFFFFu, xa, d), vb = __shfl_up_sync(0xFFFFFFFFu, xb, d);
            if (lane >= (uint32_t)d) {
                xa += va;
                xb += vb;
            }
        }
        if (lane < kPackThreads / 32) {
            sa[lane] = xa - wa;
            sb[lane] = xb - wb;
        }
        if (lane == kPackThreads / 32 - 1) {
            sa[32] = xa;
            sb[32] = xb;
        }
    }
    __syncthreads();
    const uint32_t ea = sa[warp] + ia - a, eb = sb[warp] + ib - b;
    total_a = sa[32];
    total_b = sb[32];
    a = ea;
    b = eb;
    __syncthreads();
}

__global__ void __launch_bounds__(kAnsThreads)
k_ans_chain(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AnsShared &s = *reinterpret_cast<AnsShared *>(smem_raw);
    const uint32_t tile = ws.chain_lpt ? ws.chain_order[blockIdx.x] : blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ws.tiles[tile].flags & kTilePrefix)   // pseudo-tile of a multi-group frame (k_frame.cu)
        return;
    const uint32_t N = ws.nsyms[tile];
    const uint32_t *__restrict__ sy = ws.syms + (size_t)tile * kMaxHfSyms;
    uint32_t *__restrict__ flags = ws.flags + (size_t)tile * (kMaxHfSyms / 32);
    uint16_t *__restrict__ fwords = ws.fwords + (size_t)tile * kMaxHfSyms;

    const long long clk0 = clock64();
    // ---- 1. model ---------------------------------------------------------------------------
    for (uint32_t i = tid; i < kHfClusters * kHfTokens; i += kAnsThreads)
        s.hist[i] = ws.hist[(size_t)tile * kHfClusters * kHfTokens + i];
    if (tid == 0)
        s.err = 0;
    __syncthreads();
    if (tid < (uint32_t)kHfClusters) {
        uint32_t a = 0;
        for (uint32_t k = 0; k < (uint32_t)kHfTokens; k++)
            if (s.hist[tid * kHfTokens + k])
                a = k + 1;
        s.alpha[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t mx = 0;
        for (int c = 0; c < kHfClusters; c++)
            mx = s.alpha[c] > mx ? s.alpha[c] : mx;
        s.own_alpha = mx;
        // one-frame mode over several LF groups: the alphabet bound is the largest seen so far in the
        // frame's stream, not this LF group's alone (entropy.c:459, 952)
        const uint32_t floor_alpha = tile_alpha_floor(ws.tiles[tile]);
        mx = floor_alpha > mx ? floor_alpha : mx;
        int la = mx ? ceil_log2_u32(mx) : 0;
        s.log_alpha = la < 5 ? 5 : la;   // reference: entropy.c:952 (<= 6 because tokens < 64)
    }
    __syncthreads();
    const int log_alpha = s.log_alpha;
    if (tid < (uint32_t)kHfClusters) {
        const uint32_t c = tid, a = s.alpha[c];
        AnsCluster &cl = s.cl[c];
        if (!a) {
            cl.alpha = 0;
            cl.single = 0;
            for (int k = 0; k < kHfTokens; k++) {
                cl.freq[k] = 0;
                cl.cum[k] = 0;
            }
        } else {
            const int single = ans_normalise(&s.hist[c * kHfTokens], a);
            if (single < 0 || !ans_build_alias(cl, &s.hist[c * kHfTokens], a, log_alpha, single > 0))
                atomicOr(&s.err, (uint32_t)kErrAlias);
        }
    }
    __syncthreads();
    // per-symbol chain constants (one 64-bit division each): all threads
    // symbols carry one of nine local clusters; with folded presets (many LF groups in one frame)
    // several of them share a model (the histograms were folded by k_frame_hist_sum)
    const uint32_t kclusters = tile_clusters(ws.tiles[tile]);
    for (uint32_t idx = tid; idx < kHfClusters * kHfTokens; idx += kAnsThreads) {
        const uint32_t c9 = idx / kHfTokens, k = idx - c9 * kHfTokens, c = hf_fold_cluster(c9, kclusters);
        const AnsCluster &cl = s.cl[c];
        const AnsSymInfo si = ans_sym_info(cl.freq[k], c * kAnsTotal + cl.cum[k]);
        s.info4[idx] = make_uint4(si.mc, si.ne, si.nf2, si.b2);
        if (ws.dbg_freqs)
            ws.dbg_freqs[(size_t)tile * kHfClusters * kHfTokens + idx] = cl.freq[k];
    }
    for (uint32_t idx = tid; idx < kHfClusters * kAnsTotal; idx += kAnsThreads) {
        const uint32_t c = idx >> 12, slot = idx & (kAnsTotal - 1);
        if (s.alpha[c]) {
            uint32_t sym, off;
            ans_slot_symbol(s.cl[c], slot, log_alpha, sym, off);
            s.inv[c * kAnsTotal + s.cl[c].cum[sym] + off] = (uint16_t)(slot + kSlotBias);
        }
    }
    // (section D is written further down by a third warp, while the chain is already running)
    __syncthreads();
    const bool sane = !s.err && N > 0 && log_alpha <= 6;
    // Two warps stay: the CHAIN warp runs nothing but the state recurrence; a HELPER warp feeds it
    // staged per-symbol records through a small shared-memory ring and drains the states it leaves
    // behind (renormalisation flags / words).  Co-resident CTAs put their chain warps on different
    // SM sub-partitions: each CTA takes the next ticket of its SM and picks the warp whose hardware
    // slot (%warpid) lives in sub-partition ticket % 4.
    if (tid == 0) {
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        s.chain_warp = 0xFFFFFFFFu;
        s.ticket = atomicAdd(&ws.sm_ticket[smid & 255u], 1u) & 3u;
    }
    __syncthreads();
    if (lane == 0) {
        uint32_t hw;
        asm("mov.u32 %0, %%warpid;" : "=r"(hw));
        if ((hw & 3u) == s.ticket)
            atomicMin(&s.chain_warp, warp);
    }
    __syncthreads();
    if (s.chain_warp == 0xFFFFFFFFu && tid == 0)
        s.chain_warp = 0;
    __syncthreads();
    const uint32_t chain_warp = s.chain_warp, helper_warp = chain_warp ^ 1u, header_warp = chain_warp ^ 2u;
    if (warp == header_warp) {
        // ---- 2. section D: only the packer needs it, so it is written off the chain's path ---------
        if (lane == 0) {
            BitSink bw;
            bw.init(s.dbits, kDBitsWords);
            bw.put_bool(0);                               // use_prefix_codes = 0
            bw.put((uint32_t)(log_alpha - 5), 2);
            for (uint32_t c = 0; c < kclusters; c++)
                ps_put_hybrid_cfg(bw, 4, 1, 0, log_alpha);
            s.dhist_off = bw.bitlen();
            for (uint32_t c = 0; c < kclusters; c++)
                ans_put_histogram(bw, s.cl[c].freq, s.alpha[c]);
            bw.flush_partial();
            s.dbitlen = bw.bitlen();
            if (bw.overflow)
                atomicOr(&ws.tile_err[tile], (uint32_t)kErrSlab);
        }
        __syncwarp();
        const uint32_t words = (s.dbitlen + 31) >> 5;
        for (uint32_t i = lane; i < words && i < (uint32_t)kDBitsWords; i += 32)
            ws.dbits[(size_t)tile * kDBitsWords + i] = s.dbits[i];
        if (lane == 0)   // bits 0-15 length of D, 16-23 where its histograms start, 24-31 this tile's largest alphabet
            ws.chain_out[tile * 4 + 2] = s.dbitlen | (s.dhist_off << 16) | (s.own_alpha << 24);
        return;
    }
    if (warp != chain_warp && warp != helper_warp)
        return;   // the remaining warps have nothing to do while the chain runs
    if (!sane) {
        if (warp == chain_warp && lane == 0) {
            ws.chain_out[tile * 4 + 0] = 0;
            ws.chain_out[tile * 4 + 1] = 0;
            ws.chain_out[tile * 4 + 3] = s.err ? s.err : (uint32_t)kErrAlphabet;
        }
        return;
    }

    // ---- 3. the chain ------------------------------------------------------------------------------
    // Batches of 32 symbols, last batch first; batch number `seq` uses ring slot seq % kRing.
    //   FULL[slot]  : helper arrives after staging a batch, chain waits before coding it
    //   EMPTY[slot] : chain arrives after coding it, helper waits before draining / restaging it
    const uint32_t FULL = 0xFFFFFFFFu;
    const int nbatch = (int)((N + 31) >> 5);
    const uint32_t inv_base = (uint32_t)__cvta_generic_to_shared(s.inv);
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(s.stage);
    const uint32_t cap_base = (uint32_t)__cvta_generic_to_shared(s.cap);
    const long long clk1 = clock64();
    if (warp == helper_warp) {
        auto load_sym = [&](int bi) -> uint32_t {
            const uint32_t p = (uint32_t)bi * 32u + lane;
            return (bi >= 0 && p < N) ? sy[p] : 0xFFFFFFFFu;
        };
        auto info_of = [&](uint32_t sym) -> uint4 {   // unused lanes: frequency 0 (never flags)
            if (sym == 0xFFFFFFFFu)
                return make_uint4(0u, 0u, 0u, 0u);
            return s.info4[hf_cluster(sym) * kHfTokens + hf_token(sym)];
        };
        uint32_t cnt = 0, lowest_flag = 0xFFFFFFFFu, gap_err = 0;
        uint32_t carry_s = kAnsInitState;   // the virtual step in front of the last symbol left 0x130000
        // drain batch `seq`: lane L owns symbol base + L; its flag / word come from the state left
        // by step L + 1, or by the previous batch's step 0 for the top lane
        auto drain = [&](int seq) {
            const int slot = seq % kRing, bi = nbatch - 1 - seq;
            const uint32_t base = (uint32_t)bi * 32u;
            const int jtop = (int)((N - 1 - base) < 31u ? (N - 1 - base) : 31u);
            const uint32_t capb = cap_base + (uint32_t)slot * 32u * 4u;
            const uint32_t sprev = (int)lane >= jtop ? carry_s : lds32(capb + (lane + 1u) * 4u);
            const bool flagged = (int)lane <= jtop && (sprev >> 20) >= s.fring[slot][lane];
            const uint32_t mask = __ballot_sync(FULL, flagged);
            carry_s = lds32(capb);
            if (lane == 0)
                flags[bi] = mask;
            if (flagged) {
                const uint32_t above = __popc(mask & ~((2u << lane) - 1u));
                fwords[cnt + above] = (uint16_t)(sprev & 0xFFFFu);   // chain order = descending position
            }
            if (mask) {
                const uint32_t hi = base + 31u - (uint32_t)__clz(mask), lo = base + (uint32_t)__ffs(mask) - 1u;
                if (lowest_flag != 0xFFFFFFFFu && lowest_flag - hi >= 65536u)
                    gap_err = 1;
                lowest_flag = lo;
            }
            cnt += __popc(mask);
        };
        uint4 inf_cur = info_of(load_sym(nbatch - 1));
        uint32_t sym_nxt = load_sym(nbatch - 2);
        for (int seq = 0; seq < nbatch; seq++) {
            const int slot = seq % kRing, bi = nbatch - 1 - seq;
            const uint32_t sym_nn = load_sym(bi - 2);   // consumed one iteration later
            const uint4 inf_nxt = info_of(sym_nxt);
            if (seq >= kRing) {
                bar_sync(kBarEmpty + slot, 64);
                drain(seq - kRing);
            }
            // record of lane L = the chain constants of symbol base + L, table address folded in;
            // unused lanes of the last batch carry frequency 0 and are never read by the chain
            s.stage[slot][lane] = make_uint4(inf_cur.x, inf_cur.y, inf_cur.z, inf_cur.w + inv_base - kRecAddrBias);
            s.fring[slot][lane] = (0u - inf_cur.z) >> 1;
            bar_arrive(kBarFull + slot, 64);
            inf_cur = inf_nxt;
            sym_nxt = sym_nn;
        }
        for (int seq = nbatch > kRing ? nbatch - kRing : 0; seq < nbatch; seq++) {
            bar_sync(kBarEmpty + seq % kRing, 64);
            drain(seq);
        }
        if (lowest_flag != 0xFFFFFFFFu && lowest_flag >= 65536u)
            gap_err = 1;   // the reference keeps this distance in a uint16_t (entropy.c:16, 1094, 1123)
        if (lane == 0) {
            ws.chain_out[tile * 4 + 0] = cnt;
            ws.chain_out[tile * 4 + 3] = gap_err ? (uint32_t)kErrAnsGap : 0u;
        }
    } else {
        // Steps run in chain order n = 0 .. N-1 (symbol N-1-n); a step needs the record of its own
        // symbol and, in the table load's shadow, that of the next one.  Only the first batch can be
        // partial; it runs through a plain loop.  Full batches are 32 straight-line steps whose records
        // roll through three register sets loaded two steps ahead -- across the batch boundary too,
        // which is why the FULL barrier of batch seq + 1 is taken inside batch seq.
        AnsCarry c;
        uint32_t x = 0, q12_prev = 0;
        auto rec_at = [&](int bi, int j) -> uint32_t {   // shared address of the record of symbol bi * 32 + j
            return stage_base + (uint32_t)((((nbatch - 1 - bi) % kRing) * 32 + j) * 16);
        };
        // One step, spelled out in the order the instructions should issue.  `own` / `nxt` = {M, e, -2f, table
        // address - 2}.  This is the third form of ans_chain.cuh's division: the table holds slot + 1
        // (kSlotBias), so the dependent multiply is (slot + 1) * M and the shadow's first wide multiply is a * M
        // without a "+ M" addend: R = a M + qa e with qa = hi32(a M), which is a/f - 2 .. a/f, so
        // y = x - qa f < 3f + 4095 < 2^14 still holds and hi32((v + 1) M + R) = qa + hi32((y + 1) M) = x / f.
        // After a renormalisation v does not count and the + 1 moves into the known part: a' = (q >> 4) + 1,
        // R = a' M + qa e.  Both cases address the table 2 bytes high; the helper folds - 2 into the records'
        // table address (kRecAddrBias).  Order: the table load first, then, in its shadow, the state export and
        // the record prefetch (two steps ahead), then the state recurrence q -> p -> a -> a M -> R.
        // The state a step leaves, (q << 12) | slot, is complete in the FOLLOWING step (once the table load has
        // returned), which stores it for the helper.  Wider stores were measured and are slower: {q << 12,
        // slot + 1} as 8 bytes per step (no combining add) 50.3 cycles per symbol, two steps' pairs as one
        // 16-byte store every other step 53.9, against 49.1 for this 4-byte store.
#define HYDB_ANS_STEP(own, nxt, thr_n, EXPORT, cap_addr, PREFETCH)                                             \
        {                                                                                                      \
            const uint32_t vprev = c.v;                                                                        \
            const uint32_t q = ans_hi32((uint64_t)vprev * c.meff + c.R);                                       \
            const uint32_t cv = vprev * c.k + c.c0;                                                            \
            const uint32_t slotv = lds16(q * (own).z + cv);                                                    \
            if (EXPORT)                                                                                        \
                sts32((cap_addr), q12_prev + vprev - kSlotBias);                                               \
            PREFETCH;                                                                                          \
            const bool p = q >= (thr_n);                                                                       \
            const uint32_t q4 = (q >> 4) + 1u;                                                                 \
            q12_prev = q << 12;                                                                                \
            const uint32_t a = p ? q4 : q12_prev;                                                              \
            c.meff = p ? 0u : (nxt).x;                                                                         \
            c.k = p ? 0u : 2u;                                                                                 \
            const uint64_t w = (uint64_t)a * (nxt).x;                                                          \
            const uint32_t qa = ans_hi32(w);                                                                   \
            c.c0 = 2u * a + (nxt).w;                                                                           \
            c.R = w + (uint64_t)qa * (nxt).y;                                                                  \
            c.v = slotv;                                                                                       \
        }
        bar_sync(kBarFull + 0, 64);
        {
            const uint4 fr = lds128(rec_at(nbatch - 1, (int)((N - 1) & 31u)));
            AnsSymInfo first;
            first.mc = fr.x; first.ne = fr.y; first.nf2 = fr.z; first.b2 = fr.w;
            ans_chain_begin(c, first, kRecAddrBias);
        }
        uint4 r0, r1, r2;
        r0 = r1 = r2 = make_uint4(0u, 0u, 0u, 0u);
        uint32_t prev_cap0 = 0;   // where the state left by the previous batch's last step goes
        int prev_slot = -1;
        for (int seq = 0; seq < nbatch; seq++) {
            const int slot = seq % kRing, bi = nbatch - 1 - seq;
            const uint32_t base = (uint32_t)bi * 32u;
            const int jtop = (int)((N - 1 - base) < 31u ? (N - 1 - base) : 31u);
            const uint32_t stg = stage_base + (uint32_t)slot * 32u * 16u, capb = cap_base + (uint32_t)slot * 32u * 4u;
            const uint32_t nstg = stage_base + (uint32_t)((seq + 1) % kRing) * 32u * 16u;
            if (jtop == 31) {
                if (seq == 0 || (seq == 1 && ((N - 1) & 31u) != 31u)) {   // first straight-line batch: fill the pipeline
                    r0 = lds128(stg + 31 * 16);
                    r1 = lds128(stg + 30 * 16);
                }
#pragma unroll
                for (int j = 31; j >= 0; --j) {
                    const uint4 own = r0, nxt = r1;
                    if (j == 6 && seq + 1 < nbatch)   // the next batch's records are read from step 1 on
                        bar_sync(kBarFull + (seq + 1) % kRing, 64);
                    const uint32_t pf = j >= 2 ? stg + (uint32_t)(j - 2) * 16u : nstg + (uint32_t)(30 + j) * 16u;
                    const uint32_t thr_n = (j == 0 && bi == 0) ? kAnsNoNext : (0u - nxt.z) << 7;
                    if (j == 31) {   // the previous batch's last state is complete now: store it and release the batch
                        HYDB_ANS_STEP(own, nxt, thr_n, prev_slot >= 0, prev_cap0, r2 = lds128(pf));
                        if (prev_slot >= 0)
                            bar_arrive(kBarEmpty + prev_slot, 64);
                    } else {
                        HYDB_ANS_STEP(own, nxt, thr_n, true, capb + (uint32_t)(j + 1) * 4u, r2 = lds128(pf));
                    }
                    r0 = r1;
                    r1 = r2;
                }
            } else {
                if (seq + 1 < nbatch)
                    bar_sync(kBarFull + (seq + 1) % kRing, 64);
                for (int j = jtop; j >= 0; --j) {
                    const uint4 own = lds128(stg + (uint32_t)j * 16u);
                    const uint4 nxt = j ? lds128(stg + (uint32_t)(j - 1) * 16u) : lds128(nstg + 31u * 16u);
                    const uint32_t thr_n = (j == 0 && bi == 0) ? kAnsNoNext : (0u - nxt.z) << 7;
                    HYDB_ANS_STEP(own, nxt, thr_n, j < jtop, capb + (uint32_t)(j + 1) * 4u, (void)0);
                }
            }
            prev_cap0 = capb;
            prev_slot = slot;
        }
        x = q12_prev + c.v - kSlotBias;   // final state: what the last step leaves (no renormalisation follows)
        sts32(prev_cap0, x);
        bar_arrive(kBarEmpty + prev_slot, 64);
#undef HYDB_ANS_STEP
        if (lane == 0) {
            ws.chain_out[tile * 4 + 1] = x;
            if (ws.dbg_clk) {   // stage-tap builds only: cycles spent in the prologue and in the chain
                uint32_t smid;
                asm("mov.u32 %0, %%smid;" : "=r"(smid));
                ws.dbg_clk[tile * 4 + 0] = (uint32_t)(clk1 - clk0);
                ws.dbg_clk[tile * 4 + 1] = (uint32_t)(clock64() - clk1);
                ws.dbg_clk[tile * 4 + 2] = smid;
                ws.dbg_clk[tile * 4 + 3] = chain_warp;
            }
        }
    }
}

struct PackShared {
    uint32_t scan_a[kPackThreads], scan_b[kPackThreads];
    uint32_t err, ebits_total;
};

__global__ void __launch_bounds__(kPackThreads)
k_ans_pack(Workspace ws, Templates tp) {
    __shared__ PackShared s;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const TileDesc t = ws.tiles[tile];
    if (t.flags & kTilePrefix)   // pseudo-tile of a multi-group frame: filled by k_frame_finish
        return;
    const bool multi = (t.flags & kTileMulti) != 0;   // group of a multi-group frame: PassGroup section only
    const uint32_t N = ws.nsyms[tile];
    const uint32_t *__restrict__ sy = ws.syms + (size_t)tile * kMaxHfSyms;
    const uint32_t *__restrict__ flags = ws.flags + (size_t)tile * (kMaxHfSyms / 32);
    const uint16_t *__restrict__ fwords = ws.fwords + (size_t)tile * kMaxHfSyms;
    uint8_t *slab = ws.slab + (size_t)tile * kSlabBytes;
    uint32_t *payload = reinterpret_cast<uint32_t *>(slab + kSlabHeaderReserve);
    constexpr uint32_t kPayloadCapBits = (uint32_t)(kSlabBytes - kSlabHeaderReserve) * 8u - 64u;
    const uint32_t W = ws.chain_out[tile * 4 + 0], final_state = ws.chain_out[tile * 4 + 1];
    const uint32_t ld = ws.chain_out[tile * 4 + 2] & 0xFFFFu, chain_err = ws.chain_out[tile * 4 + 3];
    if (tid == 0) {
        s.err = chain_err;
        s.ebits_total = 0;
    }

    // ---- 4. payload prefix A | L | B | D ---------------------------------------------------------
    const uint32_t la = multi ? 0u : tp.bits[0], lb = multi ? 0u : tp.bits[1 + t.shape];
    const uint32_t ll = multi ? 0u : ws.lfbitlen[tile];
    // multi: A / L / B / D live in the frame's prefix; the section opens with the group's HF preset id
    // (zero bits wide unless the frame has several LF groups, encoder.c:937)
    const uint32_t e_start = multi ? tile_preset_bits(t) : la + ll + lb + ld;
    const bool sane = la != 0xFFFFFFFFu && lb != 0xFFFFFFFFu && !chain_err && N > 0 && !ws.tile_err[tile];
    for (uint32_t w = tid; w < (e_start >> 5) + 3; w += kPackThreads)
        payload[w] = 0;
    __syncthreads();
    if (sane && multi && tid == 0 && tile_preset_bits(t))
        atomicOr(&payload[0], tile_preset(t));
    if (sane && !multi) {
        append_bits(payload, 0, tp.words, la, tid, kPackThreads);
        append_bits(payload, la, ws.lfbits + (size_t)tile * kLfBitsWords, ll, tid, kPackThreads);
        append_bits(payload, (uint64_t)la + ll, tp.words + (size_t)(1 + t.shape) * kTemplWords, lb, tid, kPackThreads);
        append_bits(payload, (uint64_t)la + ll + lb, ws.dbits + (size_t)tile * kDBitsWords, ld, tid, kPackThreads);
    }

    // ---- 5. forward packing of section E -------------------------------------------------------------
    const uint64_t total_bits64 = (uint64_t)e_start + 32u + ws.resbits[tile] + 16ull * W;
    const bool fits = sane && total_bits64 <= kPayloadCapBits;
    if (tid == 0 && sane && !fits)
        s.err |= kErrSlab;
    const uint32_t total_bits = fits ? (uint32_t)total_bits64 : 0;
    if (fits) {
        for (uint32_t w = (e_start >> 5) + 3 + tid; w <= (total_bits >> 5) + 1; w += kPackThreads)
            payload[w] = 0;
    }
    __syncthreads();
    {
        const uint32_t nchunks = (N + 31) >> 5, cpt = (nchunks + kPackThreads - 1) / kPackThreads;
        const uint32_t c0 = tid * cpt < nchunks ? tid * cpt : nchunks;
        const uint32_t c1 = c0 + cpt < nchunks ? c0 + cpt : nchunks;
        uint32_t bits = 0, nfl = 0;
        if (fits) {
            for (uint32_t c = c0; c < c1; c++) {
                const uint32_t fl = flags[c];
                nfl += __popc(fl);
                const uint4 *q = reinterpret_cast<const uint4 *>(sy + (size_t)c * 32);
                uint4 vv[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    vv[k] = q[k];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = vv[k];
                    const uint32_t p = c * 32 + k * 4;
                    bits += (p + 0 < N ? hf_nbits(v.x) : 0) + (p + 1 < N ? hf_nbits(v.y) : 0) +
                            (p + 2 < N ? hf_nbits(v.z) : 0) + (p + 3 < N ? hf_nbits(v.w) : 0);
                }
            }
            bits += 16 * nfl;
        }
        uint32_t tot_bits, tot_fl;
        block_scan2(bits, nfl, s.scan_a, s.scan_b, tid, tot_bits, tot_fl);
        if (fits) {
            if (tid == 0) {
                s.ebits_total = 32u + tot_bits;
                if (tot_fl != W || (uint64_t)e_start + 32u + tot_bits != total_bits64)
                    s.err |= kErrSlab;   // internal consistency check
                // final state, low half first (entropy.c:1127-1130)
                const uint32_t st = final_state, sh = e_start & 31u, w0 = e_start >> 5;
                atomicOr(&payload[w0], st << sh);
                if (sh)
                    atomicOr(&payload[w0 + 1], st >> (32 - sh));
            }
            const uint64_t start = (uint64_t)e_start + 32u + bits;
            uint32_t wpos = (uint32_t)(start >> 5);
            uint32_t nacc = (uint32_t)(start & 31u);
            uint64_t acc = 0;
            bool first = nacc != 0;
            int widx = (int)W - 1 - (int)nfl;   // next renormalisation word in forward order
            auto put = [&](uint32_t v, uint32_t n) {
                acc |= (uint64_t)v << nacc;
                nacc += n;
                if (nacc >= 32) {
                    if (first)
                        atomicOr(&payload[wpos], (uint32_t)acc);
                    else
                        payload[wpos] = (uint32_t)acc;
                    first = false;
                    wpos++;
                    acc >>= 32;
                    nacc -= 32;
                }
            };
            for (uint32_t c = c0; c < c1; c++) {
                const uint32_t fl = flags[c];
                const uint4 *q = reinterpret_cast<const uint4 *>(sy + (size_t)c * 32);
                uint4 vv[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    vv[k] = q[k];
                // this chunk's renormalisation words, fetched together (forward order = descending index)
                const int nw = __popc(fl);
                uint16_t wbuf[32];
#pragma unroll
                for (int k = 0; k < 32; k++)
                    wbuf[k] = k < nw ? fwords[widx - k] : (uint16_t)0;
                int wk = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = vv[k];
                    const uint32_t sv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t p = c * 32 + k * 4 + u;
                        if (p < N) {
                            if ((fl >> (k * 4 + u)) & 1u)
                                put(wbuf[wk++], 16);
                            const uint32_t nb = hf_nbits(sv[u]);
                            if (nb)
                                put(hf_residue(sv[u]), nb);
                        }
                    }
                }
                widx -= nw;
            }
            if (nacc && (uint32_t)acc)
                atomicOr(&payload[wpos], (uint32_t)acc);
        }
    }
    __syncthreads();

    // ---- 6. frame header + TOC, lengths ------------------------------------------------------------
    if (tid == 0) {
        uint32_t err = s.err;
        if (!sane && !err)
            err |= kErrSlab;
        uint32_t flen = 0, foff = kSlabHeaderReserve;
        if (fits && !err && multi) {
            flen = (total_bits + 7) >> 3;   // the section, byte aligned (encoder.c:976-981)
        } else if (fits && !err) {
            const uint32_t payload_bytes = (total_bits + 7) >> 3;
            uint8_t hb[kSlabHeaderReserve];
            uint32_t n = 0;
            uint32_t hw[12];
            BitSink bw;
            if (t.flags & kTileFirst) {   // image header in front of the codestream's first frame
                if (image_needs_level10(t.image_w, t.image_h))
                    n += put_level10_prefix(hb);
                bw.init(hw, 12);
                put_image_header(bw, t.image_w, t.image_h);
                bw.flush_partial();
                for (uint32_t i = 0; i < (bw.bitlen() >> 3); i++)
                    hb[n++] = (uint8_t)(hw[i >> 2] >> (8 * (i & 3)));
            }
            bw.init(hw, 12);
            put_frame_header(bw, (t.flags & kTileCrop) != 0, t.x0, t.y0, t.w, t.h, (t.flags & kTileLast) != 0);
            const bool ok = put_toc_entry(bw, payload_bytes);
            bw.flush_partial();
            const uint32_t fb = bw.bitlen() >> 3;
            if (!ok || bw.overflow || n + fb > (uint32_t)kSlabHeaderReserve) {
                err |= kErrSlab;
            } else {
                for (uint32_t i = 0; i < fb; i++)
                    hb[n++] = (uint8_t)(hw[i >> 2] >> (8 * (i & 3)));
                foff = kSlabHeaderReserve - n;
                for (uint32_t i = 0; i < n; i++)
                    slab[foff + i] = hb[i];
                flen = n + payload_bytes;
            }
        }
        ws.frame_off[tile] = foff;
        ws.frame_len[tile] = flen;
        if (err)
            atomicOr(&ws.tile_err[tile], err);
        if (ws.dbg_sect) {
            ws.dbg_sect[tile * 4 + 0] = la + ll + lb;
            ws.dbg_sect[tile * 4 + 1] = ld;
            ws.dbg_sect[tile * 4 + 2] = fits ? s.ebits_total : 0;
            ws.dbg_sect[tile * 4 + 3] = total_bits;
        }
    }
}

// HYDRIUM_B200_CHAIN=table|compact forces one kernel (tests, measurements).  Otherwise by launch size:
// the table kernel runs 2 chains per SM at ~55 cycles per symbol, the compact kernel 16 per SM at
// ~150 under full load (~90 when an SM holds few), i.e. twice the throughput but a longer critical
// path.  With chain lengths spread by ~1.3x around their mean the two meet near 4.3 tiles per SM
// (profiles/r02_config_shares_*.txt: 512 tiles 9.5 ms table / 10.9 ms compact, 8192 tiles 85 / 50 ms).
static int chain_mode_override() {
    static const int mode = [] {
        const char *e = getenv("HYDRIUM_B200_CHAIN");
        return !e ? 0 : (e[0] == 't' ? 1 : (e[0] == 'c' ? 2 : 0));
    }();
    return mode;
}
static uint32_t chain_table_limit() {
    static const uint32_t lim = [] {
        const char *e = getenv("HYDRIUM_B200_CHAIN_TABLE_MAX");
        if (e && atoi(e) > 0)
            return (uint32_t)atoi(e);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (uint32_t)(sms * 13 / 3);   // 641 tiles on a B200
    }();
    return lim;
}

// rank of every tile by symbol count (descending, ties by index): chain_order[rank] = tile
__global__ void __launch_bounds__(256)
k_chain_order(Workspace ws, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t mine = ws.nsyms[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t o = ws.nsyms[j];
        rank += (o > mine || (o == mine && j < i)) ? 1u : 0u;
    }
    ws.chain_order[rank] = i;
}

void launch_ans_chain(const Workspace &ws, uint32_t ntiles, cudaStream_t st, bool allow_compact, uint32_t concurrent_tiles,
                      bool longest_first) {
    const int mode = ws.chain_mode ? (int)ws.chain_mode : chain_mode_override();
    // what counts is how many chains compete for the GPU at once: a band of a larger batch is launched next
    // to its sibling bands on other streams
    const uint32_t competing = concurrent_tiles > ntiles ? concurrent_tiles : ntiles;
    const bool compact = allow_compact && (mode == 2 || (mode == 0 && competing > chain_table_limit()));
    static const int sms = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    static const bool lpt_on = [] { const char *e = getenv("HYDRIUM_B200_CHAIN_LPT"); return !(e && e[0] == '0'); }();
    Workspace w = ws;
    w.chain_lpt = 0;
    if (longest_first && lpt_on && competing > (uint32_t)(compact ? 16 * sms : 2 * sms)) {
        k_chain_order<<<(ntiles + 255) / 256, 256, 0, st>>>(ws, ntiles);
        w.chain_lpt = 1;
    }
    if (compact) {
        launch_ans_chain_compact(w, ntiles, st);
        return;
    }
    cudaFuncSetAttribute(k_ans_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AnsShared));
    prefer_max_shared(k_ans_chain);
    k_ans_chain<<<ntiles, kAnsThreads, sizeof(AnsShared), st>>>(w);
}

void launch_ans_pack(const Workspace &ws, const Templates &t, uint32_t ntiles, cudaStream_t st) {
    prefer_max_shared(k_ans_pack);
    k_ans_pack<<<ntiles, kPackThreads, 0, st>>>(ws, t);
}

}  // namespace hydb
