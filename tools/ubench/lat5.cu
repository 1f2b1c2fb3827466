// Can two half-warps run independent redux.sync.max at once (different member masks in one instruction),
// what does it cost, and what is the issue rate of REDUX / CREDUX per SM sub-partition?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int M>
__global__ void k(uint64_t *out, const volatile uint32_t *prm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t half_mask = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
    uint32_t x = prm[4] + lane * 7u + (threadIdx.x >> 5), acc = 0, bad = 0;
    constexpr int N = 4096;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        if (M == 0) {          // full-warp CREDUX chain
            x = __reduce_max_sync(0xFFFFFFFFu, (x * 2654435761u) >> 19) + lane;
        } else if (M == 1) {   // half-warp REDUX chain, both halves at once
            const uint32_t v = (x * 2654435761u) >> 19;
            const uint32_t r = __reduce_max_sync(half_mask, v);
            // check against a shuffle butterfly inside the half
            uint32_t m = v;
            for (int d = 8; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, m, d); m = o > m ? o : m; }
            bad += (m != r);
            x = r + lane;
        } else if (M == 2) {   // half-warp REDUX chain without the check (latency)
            x = __reduce_max_sync(half_mask, (x * 2654435761u) >> 19) + lane;
        } else if (M == 3) {   // 16-lane butterfly by shuffles (latency)
            uint32_t m = (x * 2654435761u) >> 19;
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, m, d); m = o > m ? o : m; }
            x = m + lane;
        } else if (M == 4) {   // throughput: 4 independent CREDUX per iteration
            const uint32_t a = __reduce_max_sync(0xFFFFFFFFu, x), b = __reduce_max_sync(0xFFFFFFFFu, x ^ 1u),
                           c = __reduce_max_sync(0xFFFFFFFFu, x ^ 2u), d = __reduce_max_sync(0xFFFFFFFFu, x ^ 3u);
            acc += a + b + c + d;
            x += 5;
        } else if (M == 5) {   // throughput: 4 independent half-mask REDUX per iteration
            const uint32_t a = __reduce_max_sync(half_mask, x), b = __reduce_max_sync(half_mask, x ^ 1u),
                           c = __reduce_max_sync(half_mask, x ^ 2u), d = __reduce_max_sync(half_mask, x ^ 3u);
            acc += a + b + c + d;
            x += 5;
        }
    }
    long long t1 = clock64();
    if (lane == 0 || bad) { out[0] = (uint64_t)(t1 - t0); out[1] = x + acc; }
    if (bad) out[2] = bad;
}

int main() {
    uint64_t *d; cudaMalloc(&d, 32);
    uint32_t *dp; cudaMalloc(&dp, 64);
    uint32_t h[8] = {3, 0, 1, 2, 11, 0, 0, 0};
    cudaMemcpy(dp, h, 32, cudaMemcpyHostToDevice);
    const char *names[] = {"CREDUX.MAX full warp, dependent", "REDUX.MAX two half-warps at once + check", "REDUX.MAX two half-warps, dependent",
                           "16-lane SHFL butterfly max, dependent", "CREDUX.MAX x4 independent (per REDUX)", "REDUX.MAX half masks x4 independent (per REDUX)"};
    for (int nw = 1; nw <= 16; nw *= 4) {
        printf("--- %d warp(s) in one CTA (one SM) ---\n", nw);
        for (int v = 0; v < 6; v++) {
            cudaMemset(d, 0, 32);
            for (int rep = 0; rep < 2; rep++) {
                switch (v) {
                case 0: k<0><<<1, 32 * nw>>>(d, dp); break;
                case 1: k<1><<<1, 32 * nw>>>(d, dp); break;
                case 2: k<2><<<1, 32 * nw>>>(d, dp); break;
                case 3: k<3><<<1, 32 * nw>>>(d, dp); break;
                case 4: k<4><<<1, 32 * nw>>>(d, dp); break;
                case 5: k<5><<<1, 32 * nw>>>(d, dp); break;
                }
                cudaDeviceSynchronize();
            }
            uint64_t r[4]; cudaMemcpy(r, d, 32, cudaMemcpyDeviceToHost);
            const double per = (double)r[0] / 4096 / (v >= 4 ? 4 : 1);
            printf("%-52s %.2f cycles  mismatches %llu (%s)\n", names[v], per, (unsigned long long)r[2], cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
