// Dependent-path microbenchmarks for the COMPACT rANS step: the slot lookup is a warp-parallel search
// over <= 64 sorted pieces (two per lane) instead of one load from a 72 KB inverse table.
//   M0: IMAD.HI -> IMAD -> LDS.U16                               (the table form, for reference)
//   M1: IMAD.HI -> 4 x IMAD -> 2 x ISETP -> SEL -> SEL -> REDUX.OR
//   M2: IMAD.HI -> IMAD -> ISETP -> VOTE -> POPC -> SHFL.IDX
//   M3: REDUX.OR chain alone      M4: SHFL.IDX chain alone      M5: VOTE.BALLOT + POPC chain alone
//   M6: IMAD.HI -> IMAD -> IADD, ISETP, ... (unfolded form of M1: one IMAD then adds)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int M>
__global__ void k(uint64_t *out, const volatile uint32_t *prm, int nwarps_active) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m = prm[0], nf = prm[1];
    __shared__ uint32_t tab[4096];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = (i * 2654435761u >> 9) & 0x0FFF0FFF;
    __syncthreads();
    // sorted pieces: lane l owns [128 l, 128 l + 128), split at +64
    const uint32_t lo0 = lane * 128u, lo1 = lo0 + 64u, span = 128u;
    uint32_t d0 = prm[2] + lane, d1 = prm[3] + lane;
    uint32_t x = prm[4] + 5, acc = 0;
    constexpr int N = 4096;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        if (M == 0) {
            uint32_t hi, addr;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x), "r"(m));
            asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(hi), "r"(nf), "r"(base));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(addr));
        } else if (M == 1) {
            uint32_t q;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(m));
            // four multiply-adds off q (their addends are known before q)
            const uint32_t t = q * nf + (x - lo0), u = q * nf + (x - lo1), s0 = q * nf + (x + d0), s1 = q * nf + (x + d1);
            const uint32_t s = (int32_t)u >= 0 ? s1 : s0;
            const uint32_t cand = t < span ? s : 0u;
            x = __reduce_or_sync(0xFFFFFFFFu, cand) & 4095u;
        } else if (M == 2) {
            uint32_t q;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(m));
            const uint32_t g = q * nf + x;
            const uint32_t s0 = g + d0, s1 = g + d1;
            const uint32_t s = g >= lo1 ? s1 : s0;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, g >= lo0);
            const uint32_t idx = __popc(b) - 1u;
            x = __shfl_sync(0xFFFFFFFFu, s, idx) & 4095u;
        } else if (M == 3) {
            x = __reduce_or_sync(0xFFFFFFFFu, x ^ d0) & 4095u;
        } else if (M == 4) {
            x = __shfl_sync(0xFFFFFFFFu, x + d0, x & 31u) & 4095u;
        } else if (M == 5) {
            x = __popc(__ballot_sync(0xFFFFFFFFu, x + lane >= lo0)) + d0;
        } else if (M == 6) {
            uint32_t q;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(m));
            const uint32_t g = q * nf + x;
            const uint32_t t = g - lo0, s0 = g + d0, s1 = g + d1;
            const uint32_t s = g >= lo1 ? s1 : s0;
            const uint32_t cand = t < span ? s : 0u;
            x = __reduce_or_sync(0xFFFFFFFFu, cand) & 4095u;
        } else if (M == 7) {   // match-based: not applicable, placeholder for REDUX.MAX form
            uint32_t q;
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(m));
            const uint32_t g = q * nf + x;
            const uint32_t s0 = g + d0, s1 = g + d1;
            const uint32_t s = g >= lo1 ? s1 : s0;
            // lanes below the owner contribute nothing: max over (g >= lo0 ? (lane << 13) | s : 0)
            const uint32_t cand = g >= lo0 ? ((lane << 13) | (s & 8191u)) : 0u;
            x = __reduce_max_sync(0xFFFFFFFFu, cand) & 4095u;
        }
        acc += x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x + acc; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    uint32_t *dp; cudaMalloc(&dp, 64);
    uint32_t h[8] = {3, 0, 1, 2, 0, 0, 0, 0};
    cudaMemcpy(dp, h, 32, cudaMemcpyHostToDevice);
    const char *names[] = {"IMAD.HI -> IMAD -> LDS.U16", "IMAD.HI -> 4 IMAD -> ISETP -> SEL -> SEL -> REDUX.OR",
                           "IMAD.HI -> IMAD -> ISETP -> VOTE -> POPC -> SHFL", "REDUX.OR chain", "SHFL.IDX chain", "VOTE+POPC chain",
                           "IMAD.HI -> IMAD -> IADD/ISETP/SEL/SEL -> REDUX.OR", "IMAD.HI -> IMAD -> ... -> REDUX.MAX"};
    for (int nw = 1; nw <= 8; nw *= 2) {
        printf("--- %d warp(s) per CTA, 1 CTA ---\n", nw);
        for (int v = 0; v < 8; v++) {
            for (int rep = 0; rep < 2; rep++) {
                switch (v) {
                case 0: k<0><<<1, 32 * nw>>>(d, dp, nw); break;
                case 1: k<1><<<1, 32 * nw>>>(d, dp, nw); break;
                case 2: k<2><<<1, 32 * nw>>>(d, dp, nw); break;
                case 3: k<3><<<1, 32 * nw>>>(d, dp, nw); break;
                case 4: k<4><<<1, 32 * nw>>>(d, dp, nw); break;
                case 5: k<5><<<1, 32 * nw>>>(d, dp, nw); break;
                case 6: k<6><<<1, 32 * nw>>>(d, dp, nw); break;
                case 7: k<7><<<1, 32 * nw>>>(d, dp, nw); break;
                }
                cudaDeviceSynchronize();
            }
            uint64_t r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
            printf("%-56s %.2f cycles/iter (%s)\n", names[v], (double)r[0] / 4096, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
