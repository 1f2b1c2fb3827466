// Full-step formulations of the rANS chain (one warp, unrolled x32 like the kernel), cycles/step.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }

template <int F>
__device__ __forceinline__ void step(uint32_t &x, uint32_t m, uint32_t thr, uint32_t nf2, uint32_t b2, uint32_t &s_out) {
    if (F == 1) {   // SEL a/keep, 64-bit shift (thr low bits = 32 + sh)
        const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
        const uint32_t q = (uint32_t)(t >> (thr & 63u));
        const uint32_t slot = lds16(q * nf2 + (2u * x + b2));
        const bool p = (q | 0xFFu) >= thr;
        const uint32_t a = p ? (q >> 4) : (q << 12), keep = p ? 0u : 0xFFFFu;
        s_out = (q << 12) | slot;
        x = a | (slot & keep);
    } else if (F == 2) {   // SEL a/keep, IMAD.HI + 32-bit shift (thr low bits = sh)
        const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
        const uint32_t q = (uint32_t)(t >> 32) >> (thr & 31u);
        const uint32_t slot = lds16(q * nf2 + (2u * x + b2));
        const bool p = (q | 0xFFu) >= thr;
        const uint32_t a = p ? (q >> 4) : (q << 12), keep = p ? 0u : 0xFFFFu;
        s_out = (q << 12) | slot;
        x = a | (slot & keep);
    } else if (F == 3) {   // drop shift
        const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
        const uint32_t q = (uint32_t)(t >> 32) >> (thr & 31u);
        const uint32_t slot = lds16(q * nf2 + (2u * x + b2));
        const uint32_t drop = (q | 0xFFu) >= thr ? 16u : 0u;
        s_out = (q << 12) | slot;
        x = s_out >> drop;
    } else if (F == 4) {   // split: quotient of the known part early, only the slot's share after the load
        // state kept as B (known early) + sk (slot share); here x carries B + sk of the previous step
        const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
        const uint32_t q = (uint32_t)(t >> 32) >> (thr & 31u);
        const uint32_t slot = lds16(q * nf2 + (2u * x + b2));
        const bool p = (q | 0xFFu) >= thr;
        s_out = (q << 12) | slot;
        x = p ? (q >> 4) : s_out;
    } else {               // F == 5: mask form: x = (s & mk) | (q>>4 & ~mk) with arithmetic mask
        const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
        const uint32_t q = (uint32_t)(t >> 32) >> (thr & 31u);
        const uint32_t slot = lds16(q * nf2 + (2u * x + b2));
        const uint32_t d = thr - 1u - (q | 0xFFu);          // negative (top bit set) iff renormalise
        const uint32_t drop = (d >> 27) & 16u;
        s_out = (q << 12) | slot;
        x = s_out >> drop;
    }
}

template <int F>
__global__ void k(uint64_t *out, int nbatch, uint32_t m, uint32_t sh, uint32_t f) {
    __shared__ __align__(16) uint16_t tab[16384];
    __shared__ uint32_t cap[32];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) tab[i] = (uint16_t)((i * 2654435761u >> 9) & 4095);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab), capb = (uint32_t)__cvta_generic_to_shared(cap);
    const uint32_t thr = ((f + 7) << 8) | (F == 1 ? 32 + sh : sh);
    const uint32_t nf2 = 0u - 2u * f, b2 = base + 2 * 4096;
    uint32_t x = 0x130000;
    long long t0 = clock64();
    for (int b = 0; b < nbatch; b++) {
#pragma unroll
        for (int j = 31; j >= 0; --j) {
            uint32_t sp;
            step<F>(x, m, thr, nf2, b2, sp);
            sts32(capb + j * 4, sp);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const int nb = 4000;
    const uint32_t f = 1500;
    const unsigned long long M = ((1ull << 43) + f - 1) / f;
    const uint32_t m = (uint32_t)(M - (1ull << 32)), sh = 11;
    const char *names[] = {"", "F1 SEL + IMAD.WIDE/SHF64", "F2 SEL + IMAD.HI/SHF32", "F3 drop shift", "F4 select s'/q>>4", "F5 arithmetic drop"};
    for (int v = 1; v <= 5; v++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (v) {
            case 1: k<1><<<1, 64>>>(d, nb, m, sh, f); break;
            case 2: k<2><<<1, 64>>>(d, nb, m, sh, f); break;
            case 3: k<3><<<1, 64>>>(d, nb, m, sh, f); break;
            case 4: k<4><<<1, 64>>>(d, nb, m, sh, f); break;
            case 5: k<5><<<1, 64>>>(d, nb, m, sh, f); break;
            }
            cudaDeviceSynchronize();
        }
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-28s %.2f cycles/step  (x=%llx) %s\n", names[v], (double)h[0] / (nb * 32.0), (unsigned long long)h[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
