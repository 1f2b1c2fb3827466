// hydrium_b200/csrc/k_ans_compact.cu
//
// k_ans_chain_compact: the reverse rANS state chain (reference: entropy.c:943-978 model,
// 1083-1120 chain) for launches with MANY tiles.  Same inputs and outputs as k_ans_chain (k_ans.cu),
// different trade: that kernel keeps a 72 KB direct inverse-alias table per tile, so only two chains
// fit an SM and a large launch runs at 2 x 148 chains of ~55 cycles per symbol; this one keeps the
// alias map as <= 64 sorted linear pieces (ans_model.cuh), 4.6 KB per tile, finds the slot with a
// warp-wide compare + REDUX.MAX instead of a table load, and runs in 64-thread CTAs of ~13 KB, so
// sixteen chains share an SM.  A step is slower (the reduction costs ~45 cycles against ~25 for the
// load, tools/ubench/lat4.cu) but four chain warps per SM sub-partition interleave, and the step
// becomes issue-bound instead of latency-bound.  launch_ans_chain (k_ans.cu) picks the kernel by
// launch size; tokens >= 32 (only float samples reach them) stay with the table kernel.
#include "ans_chain.cuh"
#include "chain_util.cuh"
#include "kernels.h"
#include "prefix_coder.cuh"

namespace hydb {

constexpr int kCThreads = 64;            // chain warp + helper warp
constexpr int kCRing = 4;                // batches in flight between them
constexpr int kCTok = 32;                // tokens per cluster (log_alphabet_size 5)

struct CompactBuild {                    // prologue only
    uint32_t hist[kHfClusters * kCTok];
    AnsCluster32 cl[kHfClusters];
    uint32_t lo[kAnsPieces];
    int32_t delta[kAnsPieces];
};
struct CompactRun {                      // while the chain runs
    uint4 rec_a[kCRing][32];             // per symbol, read as "next":  {mc, -e, cum, f << 8}
    uint2 rec_b[kCRing][32];             // per symbol, read as "own":   {-f, byte offset of its cluster's pieces}
    uint2 cap[kCRing][32];               // what each step leaves: {q << 12, reduction result}; state = x | (y & 0xFFF)
    uint32_t fring[kCRing][32];          // frequencies, for the drain
};
struct CompactShared {
    AnsPieceLane pieces[kHfClusters][32];        // 4,608 B
    uint2 info[kHfClusters * kCTok];             // 2,304 B  {mc, f | cum << 13 | cluster << 25}
    union {
        CompactBuild b;
        CompactRun r;
    };
    // FULL[slot]: the helper's 32 lanes arrive once a batch is staged; EMPTY[slot]: the chain warp's 32
    // lanes arrive once it is coded.  Phase of batch `seq` on its slot: (seq / kCRing) & 1.
    uint64_t bar_full[kCRing], bar_empty[kCRing];
    uint32_t alpha[kHfClusters];
    uint32_t own_alpha, err, chain_warp;
    int log_alpha;
};

int ans_compact_smem_bytes() { return (int)sizeof(CompactShared); }

__global__ void __launch_bounds__(kCThreads, 16)
k_ans_chain_compact(Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CompactShared &s = *reinterpret_cast<CompactShared *>(smem_raw);
    const uint32_t tile = ws.chain_lpt ? ws.chain_order[blockIdx.x] : blockIdx.x;   // longest chains first
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileDesc &td = ws.tiles[tile];
    if (td.flags & kTilePrefix)
        return;
    const uint32_t N = ws.nsyms[tile];
    const uint32_t *__restrict__ sy = ws.syms + (size_t)tile * kMaxHfSyms;
    uint32_t *__restrict__ flags = ws.flags + (size_t)tile * (kMaxHfSyms / 32);
    uint16_t *__restrict__ fwords = ws.fwords + (size_t)tile * kMaxHfSyms;
    const long long clk0 = clock64();

    // ---- 1. model (the same steps as k_ans_chain, by 64 threads) --------------------------------
    if (tid == 0)
        s.err = 0;
    if (tid < (uint32_t)kCRing) {
        mbar_init((uint32_t)__cvta_generic_to_shared(&s.bar_full[tid]), 32);
        mbar_init((uint32_t)__cvta_generic_to_shared(&s.bar_empty[tid]), 32);
    }
    __syncthreads();
    {
        const uint32_t *gh = ws.hist + (size_t)tile * kHfClusters * kHfTokens;
        uint32_t high = 0;
        for (uint32_t i = tid; i < (uint32_t)(kHfClusters * kCTok); i += kCThreads) {
            const uint32_t c = i / kCTok, k = i - c * kCTok;
            s.b.hist[i] = gh[c * kHfTokens + k];
            high |= gh[c * kHfTokens + kCTok + k];
        }
        if (high)
            atomicOr(&s.err, (uint32_t)kErrAlphabet);   // the launcher keeps float tiles away from this kernel
    }
    __syncthreads();
    if (tid < (uint32_t)kHfClusters) {
        uint32_t a = 0;
        for (uint32_t k = 0; k < (uint32_t)kCTok; k++)
            if (s.b.hist[tid * kCTok + k])
                a = k + 1;
        s.alpha[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t mx = 0;
        for (int c = 0; c < kHfClusters; c++)
            mx = s.alpha[c] > mx ? s.alpha[c] : mx;
        s.own_alpha = mx;
        const uint32_t floor_alpha = tile_alpha_floor(td);   // entropy.c:459, 952 (one-frame mode)
        mx = floor_alpha > mx ? floor_alpha : mx;
        int la = mx ? ceil_log2_u32(mx) : 0;
        s.log_alpha = la < 5 ? 5 : la;
        if (s.log_alpha != 5)
            s.err |= kErrAlphabet;
    }
    __syncthreads();
    const int log_alpha = 5;
    if (tid < (uint32_t)kHfClusters && !s.err) {
        const uint32_t c = tid, a = s.alpha[c];
        AnsCluster32 &cl = s.b.cl[c];
        if (!a) {
            cl.alpha = 0;
            cl.single = 0;
            for (int k = 0; k < kCTok; k++) {
                cl.freq[k] = 0;
                cl.cum[k] = 0;
            }
        } else {
            const int single = ans_normalise(&s.b.hist[c * kCTok], a);
            if (single < 0 || !ans_build_alias(cl, &s.b.hist[c * kCTok], a, log_alpha, single > 0))
                atomicOr(&s.err, (uint32_t)kErrAlias);
        }
    }
    __syncthreads();
    const uint32_t kclusters = tile_clusters(td);
    if (!s.err) {
        for (uint32_t idx = tid; idx < (uint32_t)(kHfClusters * kCTok); idx += kCThreads) {
            const uint32_t c9 = idx / kCTok, k = idx - c9 * kCTok, c = hf_fold_cluster(c9, kclusters);
            const AnsCluster32 &cl = s.b.cl[c];
            const uint32_t f = cl.freq[k];
            s.info[idx] = make_uint2(ans_sym_info(f, 0).mc, f | ((uint32_t)cl.cum[k] << 13) | (c << 25));
            if (ws.dbg_freqs) {
                ws.dbg_freqs[(size_t)tile * kHfClusters * kHfTokens + c9 * kHfTokens + k] = f;
                ws.dbg_freqs[(size_t)tile * kHfClusters * kHfTokens + c9 * kHfTokens + kCTok + k] = 0;
            }
        }
        // sorted pieces, one cluster at a time: thread p owns piece p, its rank is its place
        for (uint32_t c = 0; c < (uint32_t)kHfClusters; c++) {
            uint32_t lo;
            int32_t delta;
            const uint32_t len = s.alpha[c] ? ans_piece(s.b.cl[c], tid, lo, delta) : 0u;
            s.b.lo[tid] = len ? lo : kAnsPieceNone;
            s.b.delta[tid] = delta;
            __syncthreads();
            const uint32_t rank = ans_piece_rank(s.b.lo, tid), L = rank >> 1;
            uint32_t *pl = reinterpret_cast<uint32_t *>(&s.pieces[c][L]);
            pl[rank & 1u] = s.b.lo[tid];
            pl[2u + (rank & 1u)] = (uint32_t)s.b.delta[tid] + 1u + (L << 13);   // + 1: the chain works on slot + 1
            __syncthreads();
        }
    }
    // ---- 2. section D straight into HBM (entropy.c:563-572, 303-369, 980-1001) ------------------------
    if (tid == 0) {
        uint32_t dbitlen = 0, dhist_off = 0;
        if (!s.err) {
            BitSink bw;
            bw.init(ws.dbits + (size_t)tile * kDBitsWords, kDBitsWords);
            bw.put_bool(0);
            bw.put((uint32_t)(log_alpha - 5), 2);
            for (uint32_t c = 0; c < kclusters; c++)
                ps_put_hybrid_cfg(bw, 4, 1, 0, log_alpha);
            dhist_off = bw.bitlen();
            for (uint32_t c = 0; c < kclusters; c++)
                ans_put_histogram(bw, s.b.cl[c].freq, s.alpha[c]);
            bw.flush_partial();
            dbitlen = bw.bitlen();
            if (bw.overflow)
                atomicOr(&ws.tile_err[tile], (uint32_t)kErrSlab);
        }
        ws.chain_out[tile * 4 + 2] = dbitlen | (dhist_off << 16) | (s.own_alpha << 24);
        // which of the two warps runs the chain: the one on the SM sub-partition (%warpid & 3) that
        // currently carries fewer chain warps of co-resident CTAs
        s.chain_warp = 0xFFFFFFFFu;
    }
    __syncthreads();   // the build area is dead from here on: the ring takes its place
    uint32_t smid, hw;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%warpid;" : "=r"(hw));
    uint32_t *load = ws.sm_load + (smid & 255u) * 4u;
    if (lane == 0)
        s.r.fring[0][warp] = *reinterpret_cast<volatile uint32_t *>(load + (hw & 3u));
    __syncthreads();
    if (tid == 0)
        s.chain_warp = s.r.fring[0][1] < s.r.fring[0][0] ? 1u : 0u;
    __syncthreads();
    const uint32_t chain_warp = s.chain_warp;
    const bool sane = !s.err && N > 0;
    if (!sane) {
        if (tid == 0) {
            ws.chain_out[tile * 4 + 0] = 0;
            ws.chain_out[tile * 4 + 1] = 0;
            ws.chain_out[tile * 4 + 3] = s.err ? s.err : (uint32_t)kErrAlphabet;
        }
        return;
    }
    if (warp == chain_warp && lane == 0)
        atomicAdd(load + (hw & 3u), 1u);

    // ---- 3. the chain --------------------------------------------------------------------------------
    const uint32_t FULLM = 0xFFFFFFFFu;
    const int nbatch = (int)((N + 31) >> 5);
    const uint32_t reca_base = (uint32_t)__cvta_generic_to_shared(s.r.rec_a);
    const uint32_t recb_base = (uint32_t)__cvta_generic_to_shared(s.r.rec_b);
    const uint32_t cap_base = (uint32_t)__cvta_generic_to_shared(s.r.cap);
    const uint32_t piece_lane = (uint32_t)__cvta_generic_to_shared(s.pieces) + lane * 16u;
    const uint32_t full_base = (uint32_t)__cvta_generic_to_shared(s.bar_full);
    const uint32_t empty_base = (uint32_t)__cvta_generic_to_shared(s.bar_empty);
    auto wait_full = [&](int seq) { mbar_wait(full_base + (uint32_t)(seq % kCRing) * 8u, (uint32_t)(seq / kCRing) & 1u); };
    auto wait_empty = [&](int seq) { mbar_wait(empty_base + (uint32_t)(seq % kCRing) * 8u, (uint32_t)(seq / kCRing) & 1u); };
    auto arrive_full = [&](int seq) { mbar_arrive(full_base + (uint32_t)(seq % kCRing) * 8u); };
    auto arrive_empty = [&](int seq) { mbar_arrive(empty_base + (uint32_t)(seq % kCRing) * 8u); };
    const long long clk1 = clock64();
    if (warp != chain_warp) {
        // ---- helper: stages the records of each batch, drains the states the chain leaves -------------
        auto load_sym = [&](int bi) -> uint32_t {
            const uint32_t p = (uint32_t)bi * 32u + lane;
            return (bi >= 0 && p < N) ? sy[p] : 0xFFFFFFFFu;
        };
        auto info_of = [&](uint32_t sym) -> uint2 {
            if (sym == 0xFFFFFFFFu)
                return make_uint2(0u, 0u);
            return s.info[hf_cluster(sym) * kCTok + (hf_token(sym) & (kCTok - 1))];
        };
        uint32_t cnt = 0, lowest_flag = 0xFFFFFFFFu, gap_err = 0;
        uint32_t carry_s = kAnsInitState;
        auto drain = [&](int seq) {
            const int slot = seq % kCRing, bi = nbatch - 1 - seq;
            const uint32_t base = (uint32_t)bi * 32u;
            const int jtop = (int)((N - 1 - base) < 31u ? (N - 1 - base) : 31u);
            const uint32_t capb = cap_base + (uint32_t)slot * 32u * 8u;
            auto state_at = [&](uint32_t addr) {
                uint32_t hi, lo;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hi), "=r"(lo) : "r"(addr));
                return hi + (lo & 0x1FFFu) - 1u;   // the reduction yields slot + 1 (third form, ans_chain.cuh)
            };
            const uint32_t sprev = (int)lane >= jtop ? carry_s : state_at(capb + (lane + 1u) * 8u);
            const bool flagged = (int)lane <= jtop && (sprev >> 20) >= s.r.fring[slot][lane];
            const uint32_t mask = __ballot_sync(FULLM, flagged);
            carry_s = state_at(capb);
            if (lane == 0)
                flags[bi] = mask;
            if (flagged) {
                const uint32_t above = __popc(mask & ~((2u << lane) - 1u));
                fwords[cnt + above] = (uint16_t)(sprev & 0xFFFFu);
            }
            if (mask) {
                const uint32_t hi = base + 31u - (uint32_t)__clz(mask), lo = base + (uint32_t)__ffs(mask) - 1u;
                if (lowest_flag != 0xFFFFFFFFu && lowest_flag - hi >= 65536u)
                    gap_err = 1;
                lowest_flag = lo;
            }
            cnt += __popc(mask);
        };
        uint2 inf_cur = info_of(load_sym(nbatch - 1));
        uint32_t sym_nxt = load_sym(nbatch - 2);
        for (int seq = 0; seq < nbatch; seq++) {
            const int slot = seq % kCRing, bi = nbatch - 1 - seq;
            const uint32_t sym_nn = load_sym(bi - 2);
            const uint2 inf_nxt = info_of(sym_nxt);
            if (seq >= kCRing) {
                wait_empty(seq - kCRing);
                drain(seq - kCRing);
            }
            const uint32_t f = inf_cur.y & 0x1FFFu, cum = (inf_cur.y >> 13) & 0xFFFu, cl = inf_cur.y >> 25;
            const uint32_t ne = (0u - f) * inf_cur.x;   // AnsSymInfo::ne = 2^32 - f * M
            s.r.rec_a[slot][lane] = make_uint4(inf_cur.x, ne, cum - 1u, f << 8);   // - 1: see HYDB_ANS_STEP_C
            s.r.rec_b[slot][lane] = make_uint2(0u - f, cl * (uint32_t)(32 * sizeof(AnsPieceLane)));
            s.r.fring[slot][lane] = f;
            arrive_full(seq);
            inf_cur = inf_nxt;
            sym_nxt = sym_nn;
        }
        for (int seq = nbatch > kCRing ? nbatch - kCRing : 0; seq < nbatch; seq++) {
            wait_empty(seq);
            drain(seq);
        }
        if (lowest_flag != 0xFFFFFFFFu && lowest_flag >= 65536u)
            gap_err = 1;
        if (lane == 0) {
            ws.chain_out[tile * 4 + 0] = cnt;
            ws.chain_out[tile * 4 + 3] = gap_err ? (uint32_t)kErrAnsGap : 0u;
        }
        return;
    }

    // ---- chain warp: every lane carries the same state; lane L also holds pieces 2L, 2L+1 of the
    //      cluster of the symbol being coded ----------------------------------------------------------
    // Carried from step to step: R, a + cum (as two addends), mc of the symbol being coded, vraw (the
    // last reduction result = slot + 1, lane tag included), vmask (0x1FFF, or 0 when the previous step
    // renormalised: then v must not count), q12_prev.  The state a step leaves ({q << 12, vraw}) is
    // stored by the FOLLOWING step, once its reduction has returned.  Full batches are 32 straight-line
    // steps: every ring address is an immediate and nothing but the step itself is issued.
    uint64_t R;
    uint32_t a_prev, cum, mc, vraw = 0, vmask = 0, q12_prev = 0;
    wait_full(0);
    uint4 nxt;   // record A of the symbol coded next
    uint2 own_b, nb;
    uint4 pc;
    {
        const int j0 = (int)((N - 1) & 31u);
        const uint4 fa = lds128(reca_base + (uint32_t)j0 * 16u);
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(own_b.x), "=r"(own_b.y) : "r"(recb_base + (uint32_t)j0 * 8u));
        AnsRecC first;
        first.mc = fa.x; first.ne = fa.y; first.nf = own_b.x; first.cum = fa.z;
        AnsCarry c;
        ans_chain_begin_c(c, first);
        R = c.R;
        a_prev = c.c0 - fa.z + 1u;   // the first step's + 1 sits in the known part (fa.z = cum - 1)
        cum = fa.z;
        mc = fa.x;
        pc = lds128(piece_lane + own_b.y);
    }
    uint32_t prev_cap = 0;    // where the state left by the previous step goes (0 = nowhere: first step)
    int prev_seq = -1;
    // `store` / `cap_addr`: whether and where the previous step's state is stored.  `nxt_a` = record A of
    // the symbol coded next, `thr_n` = its renormalisation threshold f << 8.
#define HYDB_ANS_STEP_C(nxt_a, thr_n, store, cap_addr)                                                         \
    {                                                                                                          \
        const uint32_t v = vraw & vmask;                                                                       \
        const uint32_t q = ans_hi32((uint64_t)v * mc + R);                                                     \
        const uint32_t cv = v + a_prev + cum;                                                                  \
        if (store)                                                                                             \
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(cap_addr), "r"(q12_prev), "r"(vraw));        \
        const uint32_t g = q * own_b.x + cv;                                                                   \
        const uint32_t cand = g >= pc.x ? g + (g >= pc.y ? pc.w : pc.z) : 0u;                                  \
        vraw = __reduce_max_sync(FULLM, cand);                                                                 \
        const bool p = q >= (thr_n);                                                                           \
        q12_prev = q << 12;                                                                                    \
        a_prev = p ? (q >> 4) + 1u : q12_prev;                                                                 \
        vmask = p ? 0u : 0x1FFFu;                                                                              \
        mc = (nxt_a).x;                                                                                        \
        cum = (nxt_a).z;                                                                                       \
        const uint64_t w = (uint64_t)a_prev * mc;                                                              \
        const uint32_t qa = ans_hi32(w);                                                                       \
        R = w + (uint64_t)qa * (nxt_a).y;                                                                      \
    }
#define HYDB_LDS64(dst, addr) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"((dst).x), "=r"((dst).y) : "r"(addr))
    for (int seq = 0; seq < nbatch; seq++) {
        const int slot = seq % kCRing, bi = nbatch - 1 - seq;
        const uint32_t base = (uint32_t)bi * 32u;
        const int jtop = (int)((N - 1 - base) < 31u ? (N - 1 - base) : 31u);
        const uint32_t ra = reca_base + (uint32_t)slot * 32u * 16u, rb = recb_base + (uint32_t)slot * 32u * 8u;
        const uint32_t capb = cap_base + (uint32_t)slot * 32u * 8u;
        const int nslot = (seq + 1) % kCRing;
        if (seq + 1 < nbatch)
            wait_full(seq + 1);   // the last step of this batch reads the next batch's top record
        if (jtop == 31) {
#pragma unroll
            for (int j = 31; j >= 1; --j) {
                nxt = lds128(ra + (uint32_t)(j - 1) * 16u);
                HYDB_LDS64(nb, rb + (uint32_t)(j - 1) * 8u);
                if (j == 31) {
                    HYDB_ANS_STEP_C(nxt, nxt.w, prev_cap != 0u, prev_cap);
                    if (prev_seq >= 0)
                        arrive_empty(prev_seq);   // its last state was stored by the step above
                } else {
                    HYDB_ANS_STEP_C(nxt, nxt.w, true, capb + (uint32_t)(j + 1) * 8u);
                }
                own_b = nb;
                pc = lds128(piece_lane + nb.y);
            }
            prev_cap = capb + 8u;
        } else {
            for (int j = jtop; j >= 1; --j) {
                nxt = lds128(ra + (uint32_t)(j - 1) * 16u);
                HYDB_LDS64(nb, rb + (uint32_t)(j - 1) * 8u);
                HYDB_ANS_STEP_C(nxt, nxt.w, prev_cap != 0u, prev_cap);
                if (j == jtop && prev_seq >= 0)
                    arrive_empty(prev_seq);
                prev_cap = capb + (uint32_t)j * 8u;
                own_b = nb;
                pc = lds128(piece_lane + nb.y);
            }
        }
        // step 0 of the batch: its successor's record sits on top of the next ring slot
        if (bi > 0) {
            nxt = lds128(reca_base + (uint32_t)nslot * 32u * 16u + 31u * 16u);
            HYDB_LDS64(nb, recb_base + (uint32_t)nslot * 32u * 8u + 31u * 8u);
            HYDB_ANS_STEP_C(nxt, nxt.w, prev_cap != 0u, prev_cap);
            if (jtop == 0 && prev_seq >= 0)
                arrive_empty(prev_seq);
            own_b = nb;
            pc = lds128(piece_lane + nb.y);
        } else {
            nxt = make_uint4(0u, 0u, 0u, 0u);
            HYDB_ANS_STEP_C(nxt, kAnsNoNext, prev_cap != 0u, prev_cap);
            if (jtop == 0 && prev_seq >= 0)
                arrive_empty(prev_seq);
        }
        prev_cap = capb;
        prev_seq = seq;
    }
#undef HYDB_ANS_STEP_C
#undef HYDB_LDS64
    const uint32_t x = q12_prev + (vraw & 0x1FFFu) - 1u;   // final state: what the last step leaves
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(prev_cap), "r"(q12_prev), "r"(vraw));
    arrive_empty(prev_seq);
    if (lane == 0) {
        atomicSub(load + (hw & 3u), 1u);
        ws.chain_out[tile * 4 + 1] = x;
        if (ws.dbg_clk) {
            ws.dbg_clk[tile * 4 + 0] = (uint32_t)(clk1 - clk0);
            ws.dbg_clk[tile * 4 + 1] = (uint32_t)(clock64() - clk1);
            ws.dbg_clk[tile * 4 + 2] = smid;
            ws.dbg_clk[tile * 4 + 3] = chain_warp | ((hw & 3u) << 8);
        }
    }
}

void launch_ans_chain_compact(const Workspace &ws, uint32_t ntiles, cudaStream_t st) {
    cudaFuncSetAttribute(k_ans_chain_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompactShared));
    prefer_max_shared(k_ans_chain_compact);
    k_ans_chain_compact<<<ntiles, kCThreads, sizeof(CompactShared), st>>>(ws);
}

}  // namespace hydb
