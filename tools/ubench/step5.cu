// Replica of k_ans_chain's chain-warp loop (two-part division formulation, ans_chain.cuh) over a
// static ring, plus variants that cut one of the two recurrences, to see which one binds.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o step5 step5.cu && ./step5
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hydrium_b200/csrc/ans_chain.cuh"
using namespace hydb;

__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds16v(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

__device__ __forceinline__ uint64_t madwide_u(uint32_t a, uint32_t b, uint64_t c) { uint64_t d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t madwide_s(uint32_t a, uint32_t b, uint64_t c) { uint64_t d; asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
__device__ __forceinline__ uint32_t hi_of(uint64_t w) { uint32_t h; asm("{ .reg .b32 lo; mov.b64 {lo, %0}, %1; }" : "=r"(h) : "l"(w)); return h; }
struct Sh { uint16_t inv[9 * 4096]; uint4 stage[4][32][2]; uint32_t cap[4][32]; };


// k9: fraction-word formulation (ans_chain.cuh, as in k_ans_chain).  W: 0 full; 1 shadow cut (a opaque);
// 2 table-load cut (v opaque); 3 both cut (issue-bound floor); 4 full without record prefetch/STS
template <int W>
__global__ void k9(uint64_t *out, int nbatch) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sh &s = *reinterpret_cast<Sh *>(raw);
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    for (int i = lane; i < 9 * 4096; i += 32) s.inv[i] = (uint16_t)((i * 2654435761u >> 7) & 4095);
    const uint32_t inv_base = (uint32_t)__cvta_generic_to_shared(s.inv);
    auto fq = [](uint32_t l, int rr) { return 300u + ((l * 97u + (uint32_t)rr * 31u) % 3000u); };
    auto info = [&](uint32_t l, int rr) { const uint32_t f = fq(l, rr); return ans_sym_info(f, 4096 * (l % 9) + (l * 13 % (4096 - f))); };
    for (int r = 0; r < 4; r++) {
        const AnsSymInfo own = info(lane, r), nx = lane ? info(lane - 1, r) : info(31, (r + 1) & 3);
        AnsStepRec rec = ans_step_rec(own, &nx, inv_base);
        s.stage[r][lane][0] = make_uint4(rec.f2, rec.thr_n, rec.zero, rec.b2);
        s.stage[r][lane][1] = make_uint4(rec.mf_n, rec.mc_n, rec.ne_n, 0);
    }
    __syncwarp();
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(s.stage), cap_base = (uint32_t)__cvta_generic_to_shared(s.cap);
    AnsCarry c;
    ans_chain_begin(c, info(31, 0));
    uint32_t lows = 0, q12_prev = 0, junk = 0;
    uint32_t opaque = 0x130000, opaque_v = 77;
    asm volatile("" : "+r"(opaque), "+r"(opaque_v));
    long long t0 = clock64();
    for (int seq = 0; seq < nbatch; seq++) {
        const int slot = seq & 3;
        const uint32_t stg = stage_base + slot * 1024, capb = cap_base + slot * 128;
        uint4 a0 = lds128(stg + 31 * 32), b0 = lds128(stg + 31 * 32 + 16);
        uint4 a1 = lds128(stg + 30 * 32), b1 = lds128(stg + 30 * 32 + 16);
        uint4 a2 = lds128(stg + 29 * 32), b2 = lds128(stg + 29 * 32 + 16);
#pragma unroll
        for (int j = 31; j >= 0; --j) {
            const uint4 ra = a0, rb = b0;
            a0 = a1; b0 = b1; a1 = a2; b1 = b2;
            if (W != 4 && j >= 3) { a2 = lds128(stg + (j - 3) * 32); b2 = lds128(stg + (j - 3) * 32 + 16); }
            const uint32_t vprev = c.v;
            const uint64_t t = (uint64_t)vprev * c.meff + c.R;
            const uint32_t addr = (uint32_t)(((uint64_t)(uint32_t)t * ra.x + (((uint64_t)ra.w << 32) | ra.z)) >> 32);
            const uint32_t sp = q12_prev | vprev;
            if (W != 4 && j < 31) sts32(capb + (j + 1) * 4, sp); else junk ^= sp;
            uint32_t slotv = lds16(addr);
            if (W == 2 || W == 3) { junk ^= slotv; slotv = opaque_v; }
            const uint32_t q = (uint32_t)(t >> 32);
            const bool p = q >= ra.y;
            const uint32_t q4 = q >> 4;
            q12_prev = q << 12;
            uint32_t a = p ? q4 : q12_prev;
            if (W == 1 || W == 3) { junk ^= a; a = opaque; }
            c.meff = p ? 0u : rb.y;
            const uint32_t qa = __umulhi(a, rb.x);
            const uint64_t w = (uint64_t)a * rb.y;
            lows ^= (uint32_t)t;
            c.R = w + (uint64_t)((int64_t)(int32_t)qa * (int64_t)(int32_t)rb.z);
            c.v = slotv;
        }
        sts32(capb, q12_prev | c.v);
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = (q12_prev | c.v) ^ lows ^ junk; }
}

// k10: quotient-based address, single 16-byte record per symbol {mc, -e, 2f, table offset}, qa from the
// high word of a*mc.  W: 0 full; 3 both recurrences cut; 5 state stores packed (one STS.64 per 4 steps)
template <int W>
__global__ void k10(uint64_t *out, int nbatch) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sh &s = *reinterpret_cast<Sh *>(raw);
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    for (int i = lane; i < 9 * 4096; i += 32) s.inv[i] = (uint16_t)((i * 2654435761u >> 7) & 4095);
    const uint32_t inv_base = (uint32_t)__cvta_generic_to_shared(s.inv);
    auto fq = [](uint32_t l, int rr) { return 300u + ((l * 97u + (uint32_t)rr * 31u) % 3000u); };
    auto info = [&](uint32_t l, int rr) { const uint32_t f = fq(l, rr); return ans_sym_info(f, 4096 * (l % 9) + (l * 13 % (4096 - f))); };
    uint4 *rec = reinterpret_cast<uint4 *>(s.stage);   // [4][32] records
    for (int r = 0; r < 4; r++) {
        const AnsSymInfo own = info(lane, r);
        rec[r * 32 + lane] = make_uint4(own.mc, asi_ne(own), 2u * own.nf, own.b2 + inv_base);
    }
    __syncwarp();
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(s.stage), cap_base = (uint32_t)__cvta_generic_to_shared(s.cap);
    // carry: v, meff, k, c0, R
    uint32_t v = 0, meff = 0, k = 0, c0, lows = 0, q12_prev = 0, junk = 0;
    uint64_t R;
    uint32_t opaque = 0x130000, opaque_v = 77;
    asm volatile("" : "+r"(opaque), "+r"(opaque_v));
    {
        const AnsSymInfo f0 = info(31, 0);
        const uint32_t a = ((kAnsInitState >> 20) >= asi_freq(f0)) ? (kAnsInitState >> 16) : kAnsInitState;
        const uint64_t w = (uint64_t)a * f0.mc;
        const uint32_t qa = hi_of(w) - 1u;
        R = w + (uint64_t)((int64_t)(int32_t)qa * (int64_t)(int32_t)asi_ne(f0));
        c0 = 2u * a + f0.b2 + inv_base;
    }
    uint32_t pk0 = 0, pk1 = 0;
    long long t0 = clock64();
    uint4 r0 = lds128(stage_base + 31 * 16), r1 = lds128(stage_base + 30 * 16), r2 = lds128(stage_base + 29 * 16), r3 = lds128(stage_base + 28 * 16);
    uint32_t thr1 = (0u - r1.z) << 7;
    for (int seq = 0; seq < nbatch; seq++) {
        const int slot = seq & 3;
        const uint32_t stg = stage_base + slot * 512, nstg = stage_base + ((slot + 1) & 3) * 512, capb = cap_base + slot * 128;
#pragma unroll
        for (int j = 31; j >= 0; --j) {
            const uint4 own = r0, nx = r1;
            const uint32_t thr_n = thr1;
            r0 = r1; r1 = r2; r2 = r3;
            thr1 = (0u - r1.z) << 7;
            r3 = j >= 4 ? lds128(stg + (j - 4) * 16) : lds128(nstg + (28 + j) * 16);
            const uint32_t vprev = v;
            const uint64_t tt = (uint64_t)vprev * meff + R;
            const uint32_t q = (uint32_t)(tt >> 32);
            if (W == 6 || W == 7) lows ^= (uint32_t)tt;
            const uint32_t cv = W == 7 ? c0 + ((vprev + vprev) & k) : vprev * k + c0;
            const uint32_t sp = q12_prev | vprev;
            if (W == 5) {
                if ((j & 3) == 2) pk0 = sp & 0xFFFFu; else if ((j & 3) == 1) pk0 |= sp << 16; else if ((j & 3) == 0) pk1 = sp & 0xFFFFu; else { pk1 |= sp << 16; if (j < 31) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(capb + ((j + 1) >> 2) * 8), "r"(pk0), "r"(pk1)); }
            } else if (j < 31) sts32(capb + (j + 1) * 4, sp);
            const uint32_t addr = q * own.z + cv;
            uint32_t slotv = lds16(addr);
            if (W == 3) { junk ^= slotv; slotv = opaque_v; }
            const bool p = q >= thr_n;
            const uint32_t q4 = q >> 4;
            q12_prev = q << 12;
            uint32_t a = p ? q4 : q12_prev;
            if (W == 3) { junk ^= a; a = opaque; }
            meff = p ? 0u : nx.x;
            k = W == 7 ? (p ? 0u : 0xFFFFFFFFu) : (p ? 0u : 2u);
            const uint64_t w = (uint64_t)a * nx.x;
            const uint32_t qa = hi_of(w) - 1u;
            c0 = 2u * a + nx.w;
            R = w + (uint64_t)((int64_t)(int32_t)qa * (int64_t)(int32_t)nx.y);
            v = slotv;
        }
        sts32(capb, q12_prev | v);
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = (q12_prev | v) ^ lows ^ junk ^ pk0 ^ pk1; }
}

// raw dependent-chain latencies of single SASS ops (inline PTX so nothing is folded)
template <int M>
__global__ void lat(uint64_t *out, uint32_t seed, uint32_t mulc, uint32_t thr) {
    __shared__ uint32_t tab[4096];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = (base + ((i * 4 + 64) & 0x3FFC)) & 0xFFFF;
    __syncthreads();
    uint32_t x = M == 0 ? base : seed + threadIdx.x * 0x10001u, y = seed + 7 + threadIdx.x, z = mulc ^ threadIdx.x;
    uint32_t lo = seed, hi = threadIdx.x;
    constexpr int N = 4096;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (M == 0) asm volatile("ld.shared.u16 %0, [%0];" : "+r"(x));
        else if (M == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
        else if (M == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(z), "r"(y));
        else if (M == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(z));
        else if (M == 4) asm volatile("{ .reg .b64 t, c; mov.b64 c, {%1, %2}; mad.wide.u32 t, %0, %3, c; mov.b64 {%1, %0}, t; }" : "+r"(x), "+r"(lo) : "r"(hi), "r"(z));   // hi32(x*z + C) -> x
        else if (M == 5) asm volatile("{ .reg .b64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t; }" : "+r"(x), "+r"(lo) : "r"(z));       // lo word chain of IMAD.WIDE
        else if (M == 6) asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %3, p; }" : "+r"(x) : "r"(thr), "r"(y), "r"(z));
        else if (M == 7) asm volatile("shr.u32 %0, %0, %1;" : "+r"(x) : "r"(thr));
        else if (M == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        else if (M == 9) asm volatile("{ .reg .b64 t, c; mov.b64 c, {%1, %2}; mad.wide.s32 t, %0, %3, c; mov.b64 {%0, %1}, t; }" : "+r"(x), "+r"(lo) : "r"(hi), "r"(z));   // lo32 chain signed wide
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x + lo; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const int nb = 3000;
    {
        auto run9 = [&](auto kern, const char *name) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            for (int rep = 0; rep < 2; rep++) { kern<<<1, 64, sizeof(Sh)>>>(d, nb); cudaDeviceSynchronize(); }
            uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-44s %.2f cycles/step (%s)\n", name, (double)h[0] / (nb * 32.0), cudaGetErrorString(cudaGetLastError()));
        };
        run9(k10<0>, "K10 quotient address, 1 record, full");
        run9(k10<3>, "K10 both cut");
        run9(k10<5>, "K10 full, packed state stores");
        run9(k10<6>, "K10 full, wide path multiply");
        run9(k10<7>, "K10 wide path, cv by add/and");
        run9(k9<0>, "K9 fraction-word step, full");
        run9(k9<1>, "K9 shadow cut");
        run9(k9<2>, "K9 table-load cut");
        run9(k9<3>, "K9 both cut");
    }
    const char *ln[] = {"LDS.U16 pure chase", "IADD", "IMAD", "IMAD.HI", "IMAD.HI/WIDE + 64b addend -> hi", "IMAD.WIDE -> lo", "ISETP + SEL", "SHF", "LOP3", "IMAD.WIDE signed + 64b addend -> lo"};
    for (int m = 0; m < 10; m++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
            case 0: lat<0><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 1: lat<1><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 2: lat<2><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 3: lat<3><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 4: lat<4><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 5: lat<5><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 6: lat<6><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 7: lat<7><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 8: lat<8><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            case 9: lat<9><<<1, 32>>>(d, 4, 0x9E3779B1u, 3); break;
            }
            cudaDeviceSynchronize();
        }
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-40s %.2f cycles/iter\n", ln[m], (double)h[0] / 4096);
    }
    return 0;
}
