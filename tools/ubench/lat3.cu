// Dependent-path microbenchmarks for the rANS step: (multiply-high) -> (multiply-add) -> LDS chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int M>
__global__ void k(uint64_t *out, const volatile uint32_t *prm) {
    uint32_t m = prm[threadIdx.x & 3 ? 0 : 0], c = prm[1 + (threadIdx.x >> 7)], rlo = prm[2 + (threadIdx.x >> 7)], rhi = prm[3 + (threadIdx.x >> 7)];
    m = prm[threadIdx.x >> 7];
    __shared__ uint32_t tab[4096];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = (i * 2654435761u >> 9) & 0x0FFF0FFF;
    __syncthreads();
    uint32_t x = threadIdx.x + 5, acc = 0, lo, hi, addr;
    uint64_t R = ((uint64_t)rhi << 32) | rlo;
    asm volatile("" : "+l"(R));
    constexpr int N = 4096;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        if (M == 0) {        // WIDE(+64b addend).hi -> IMAD -> LDS.U16
            { const uint64_t t = (uint64_t)x * m + R; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); asm volatile("" : "+r"(lo), "+r"(hi)); }
            acc ^= lo;
            asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(hi), "r"(c), "r"(base));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(addr));
        } else if (M == 1) { // mul.hi -> IMAD -> LDS.U16
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x), "r"(m));
            asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(hi), "r"(c), "r"(base));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(addr));
        } else if (M == 2) { // IMAD -> LDS.U16
            asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(x), "r"(c), "r"(base));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(addr));
        } else if (M == 3) { // WIDE.hi -> LDS.U16 (hi used as address offset: c*0)
            { const uint64_t t = (uint64_t)x * m + R; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); asm volatile("" : "+r"(lo), "+r"(hi)); }
            acc ^= lo;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(hi));
        } else if (M == 4) { // WIDE.hi -> IMAD -> LDS.32
            { const uint64_t t = (uint64_t)x * m + R; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); asm volatile("" : "+r"(lo), "+r"(hi)); }
            acc ^= lo;
            asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(hi), "r"(c), "r"(base));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(addr));
        } else if (M == 5) { // WIDE.hi only chain (hi feeds next multiplicand)
            { const uint64_t t = (uint64_t)x * m + R; lo = (uint32_t)t; x = (uint32_t)(t >> 32); asm volatile("" : "+r"(lo), "+r"(x)); }
            acc ^= lo;
        } else if (M == 6) { // WIDE.lo only chain
            { const uint64_t t = (uint64_t)x * m + R; x = (uint32_t)t; hi = (uint32_t)(t >> 32); asm volatile("" : "+r"(x), "+r"(hi)); }
            acc ^= hi;
        } else if (M == 7) { // setp + selp chain
            asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %3, p; }" : "+r"(x) : "r"(c), "r"(m), "r"(rlo));
        } else if (M == 8) { // LDS.U16 -> WIDE.hi (no IMAD): x = hi + base (add)
            { const uint64_t t = (uint64_t)x * m + R; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); asm volatile("" : "+r"(lo), "+r"(hi)); }
            acc ^= lo;
            asm volatile("add.u32 %0, %1, %2;" : "=r"(addr) : "r"(hi), "r"(base));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(addr));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x + acc; }
}
int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    uint32_t *dp; cudaMalloc(&dp, 64);
    auto setp = [&](uint32_t a, uint32_t b, uint32_t c2, uint32_t e) { uint32_t h[8] = {a, b, c2, e, 0, 0, 0, 0}; cudaMemcpy(dp, h, 32, cudaMemcpyHostToDevice); };
    const char *names[] = {"WIDE.hi -> IMAD -> LDS.U16", "IMAD.HI -> IMAD -> LDS.U16", "IMAD -> LDS.U16", "WIDE.hi -> LDS.U16", "WIDE.hi -> IMAD -> LDS.32", "WIDE.hi chain", "WIDE.lo chain", "ISETP+SEL chain", "WIDE.hi -> IADD -> LDS.U16"};
    for (int v = 0; v < 9; v++) {
        for (int rep = 0; rep < 2; rep++) {
            // m small so hi stays 0 with rhi = 0: addresses valid.  c = 0 for M0/1/4 (address = base), M2: c = 2 (x < 4096 -> in table)
            switch (v) {
            case 0: setp(3, 0, 1, 0); k<0><<<1, 32>>>(d, dp); break;
            case 1: setp(3, 0, 1, 0); k<1><<<1, 32>>>(d, dp); break;
            case 2: setp(3, 2, 1, 0); k<2><<<1, 32>>>(d, dp); break;
            case 3: setp(3, 0, 1, 0); k<3><<<1, 32>>>(d, dp); break;   // hi = 0 -> address 0 of shared window (valid: static smem starts at 0? use with care)
            case 4: setp(3, 0, 1, 0); k<4><<<1, 32>>>(d, dp); break;
            case 5: setp(0x9E3779B1u, 0, 1, 7); k<5><<<1, 32>>>(d, dp); break;
            case 6: setp(0x9E3779B1u, 0, 1, 7); k<6><<<1, 32>>>(d, dp); break;
            case 7: setp(0x9E3779B1u, 100, 1, 7); k<7><<<1, 32>>>(d, dp); break;
            case 8: setp(3, 0, 1, 0); k<8><<<1, 32>>>(d, dp); break;
            }
            cudaDeviceSynchronize();
        }
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-32s %.2f cycles/iter (%s)\n", names[v], (double)h[0] / 4096, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
