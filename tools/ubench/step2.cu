// Variants of the rANS step's dependent chain, one warp, to find the fastest formulation on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <int V>
__global__ void k(uint64_t *out, int n, uint32_t m, uint32_t sh, uint32_t f, uint32_t thr) {
    __shared__ __align__(16) uint32_t tab[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (i * 2654435761u >> 9) & 4095;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t nf2 = 0u - 2u * f, nf4 = 0u - 4u * f, nf = 0u - f;
    uint32_t x = 0x130000;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; i++) {
        uint32_t q, slot;
        if (V == 0) {          // IMAD.HI with 64-bit addend (current)
            const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
            q = (uint32_t)(t >> 32) >> (sh & 31);
            slot = lds16(base + ((q * nf2 + 2u * x) & 0x7FFE));
        } else if (V == 1) {   // separate mulhi + add
            q = (__umulhi(x, m) + x) >> (sh & 31);
            slot = lds16(base + ((q * nf2 + 2u * x) & 0x7FFE));
        } else if (V == 2) {   // 32-bit reciprocal only (no + x term)
            q = __umulhi(x, m) >> (sh & 31);
            slot = lds16(base + ((q * nf2 + 2u * x) & 0x7FFE));
        } else if (V == 3) {   // u32 table entries
            const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
            q = (uint32_t)(t >> 32) >> (sh & 31);
            slot = lds32(base + ((q * nf4 + 4u * x) & 0x7FFC));
        } else if (V == 4) {   // no lookup at all (arithmetic only)
            const uint64_t t = (uint64_t)x * m + ((uint64_t)x << 32);
            q = (uint32_t)(t >> 32) >> (sh & 31);
            slot = (q * nf + x) & 4095;
        } else if (V == 5) {   // lookup only: address straight from x
            q = x >> 12;
            slot = lds16(base + ((x * 2u) & 0x3FFE));
        } else {               // float reciprocal estimate + exact fix-up
            q = (uint32_t)(__uint2float_rz(x) * __uint_as_float(m));
            uint32_t r = x - q * f;
            if (r >= f) { q++; }
            slot = lds16(base + ((q * nf2 + 2u * x) & 0x7FFE));
        }
        const uint32_t s = (q << 12) | (slot & 4095);
        const uint32_t drop = (q | 0xFFu) >= thr ? 16u : 0u;
        x = s >> drop;
        if (V == 5) x |= 0x20000;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const int n = 200000;
    const uint32_t f = 1500;
    // l = 10, k = 43: M = ceil(2^43 / 1500) = 2^32 + m
    const unsigned long long M = ((1ull << 43) + f - 1) / f;
    const uint32_t m = (uint32_t)(M - (1ull << 32)), sh = 11;
    const uint32_t thr = (f << 8) | sh;
    const char *names[] = {"IMAD.HI+addend (current)", "mulhi + add", "mulhi only (32-bit recip)", "u32 table", "no lookup", "lookup only", "float estimate + fixup"};
    for (int v = 0; v < 7; v++) {
        float rf = 1.0f / f; uint32_t mf; memcpy(&mf, &rf, 4);
        for (int rep = 0; rep < 2; rep++) {
            switch (v) {
            case 0: k<0><<<1, 64>>>(d, n, m, sh, f, thr); break;
            case 1: k<1><<<1, 64>>>(d, n, m, sh, f, thr); break;
            case 2: k<2><<<1, 64>>>(d, n, 2863312u, 0, f, thr); break;
            case 3: k<3><<<1, 64>>>(d, n, m, sh, f, thr); break;
            case 4: k<4><<<1, 64>>>(d, n, m, sh, f, thr); break;
            case 5: k<5><<<1, 64>>>(d, n, m, sh, f, thr); break;
            case 6: k<6><<<1, 64>>>(d, n, mf, sh, f, thr); break;
            }
            cudaDeviceSynchronize();
        }
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-28s %.2f cycles/step  (x=%llx) %s\n", names[v], (double)h[0] / n, (unsigned long long)h[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
