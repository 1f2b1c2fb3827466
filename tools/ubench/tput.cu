// Single-warp issue throughput of the instruction classes in the rANS step (independent streams).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int M, int ONE>
__global__ void k(uint64_t *out, const volatile uint32_t *prm) {
    __shared__ __align__(16) uint32_t tab[4096];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i;
    __syncthreads();
    uint32_t m = prm[threadIdx.x >> 7], c = prm[1 + (threadIdx.x >> 7)];
    uint32_t x[8];
    uint64_t R[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i * 77 + prm[2]; R[i] = ((uint64_t)prm[3] << 32) + i; }
    uint32_t acc = 0;
    constexpr int N = 2048;
    if (ONE && threadIdx.x != 0) return;
    long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < N; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (M == 0) { const uint64_t t = (uint64_t)x[i] * m + R[i]; x[i] = (uint32_t)(t >> 32) ^ (uint32_t)t; }   // WIDE + LOP3
            else if (M == 1) { x[i] = __umulhi(x[i], m) + c; }                                                   // HI + IADD
            else if (M == 2) { x[i] = x[i] * m + c; }                                                            // IMAD
            else if (M == 3) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + i * 64)); acc ^= v.x ^ v.w; }
            else if (M == 4) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(base + i * 64)); acc ^= v; }
            else if (M == 5) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + i * 64), "r"(acc)); }
            else if (M == 6) { x[i] = (x[i] ^ m) + c; asm volatile("" : "+r"(x[i])); }                             // LOP3 + IADD (alu)
            else if (M == 7) { const uint64_t t = (uint64_t)x[i] * m + R[i]; R[i] = t; }                           // WIDE only (accumulate)
            else if (M == 9) { const uint64_t t = (uint64_t)x[i] * m; x[i] = (uint32_t)(t >> 32) ^ (uint32_t)t; }              // WIDE (no addend) + LOP3
            else if (M == 10) { const uint64_t t = (uint64_t)x[i] * m + R[i]; x[i] = (uint32_t)(t >> 32) ^ (uint32_t)t; x[i] = (x[i] ^ c) + m; asm volatile("" : "+r"(x[i])); x[i] = (x[i] ^ m) + c; }   // WIDE + 5 alu
            else if (M == 11) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + i * 64)); acc ^= v.x ^ v.y; }
            else if (M == 12) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + i * 64)); acc ^= v; }
            else if (M == 13) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + i * 64)); acc ^= v.x ^ v.w; x[i] = x[i] * m + c; x[i] = (x[i] ^ m) + c; }   // LDS.128 + IMAD + 2 alu
            else if (M == 14) { x[i] = __umulhi(x[i], m) ^ c; x[i] = (x[i] ^ m) + c; asm volatile("" : "+r"(x[i])); x[i] = (x[i] ^ c) + m; }   // HI + 5 alu
            else if (M == 8) { x[i] = (x[i] >= c) ? (x[i] >> 4) : (x[i] << 12); asm volatile("" : "+r"(x[i])); }   // ISETP SHF SHL SEL
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i] ^ (uint32_t)R[i] ^ (uint32_t)(R[i] >> 32);
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = acc; }
}
int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    uint32_t *dp; cudaMalloc(&dp, 64);
    uint32_t h[8] = {0x9E3779B1u, 12345u, 3u, 9u, 0, 0, 0, 0}; cudaMemcpy(dp, h, 32, cudaMemcpyHostToDevice);
    const char *names[] = {"IMAD.WIDE(+64b) + LOP3", "IMAD.HI + IADD", "IMAD", "LDS.128", "LDS.U16", "STS.32", "LOP3 + IADD", "-", "ISETP+SHF+SHL+SEL", "IMAD.WIDE(no addend) + LOP3", "IMAD.WIDE(+64b) + 5 alu", "LDS.64", "LDS.32", "LDS.128 + IMAD + 2 alu", "IMAD.HI + 5 alu"};
    for (int one = 0; one < 2; one++)
    for (int v = 0; v < 15; v++) {
        if (v == 7) continue;
        for (int rep = 0; rep < 2; rep++) {
#define RUN(V) case V: if (one) k<V, 1><<<1, 32>>>(d, dp); else k<V, 0><<<1, 32>>>(d, dp); break;
            switch (v) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) }
            cudaDeviceSynchronize();
        }
        uint64_t r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
        printf("%s %-28s %.2f cycles per group (%s)\n", one ? "[1 lane ]" : "[32 lanes]", names[v], (double)r[0] / (2048.0 * 8), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
