// Microbenchmark of the rANS chain step loop in isolation (one warp), toggling its side ops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hydrium_b200/csrc/ans_chain.cuh"
using namespace hydb;

__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

struct Sh { uint16_t inv[9 * 4096]; uint4 stage[32]; uint32_t cap[32]; };

template <int MODE>
__global__ void k(uint64_t *out, int nbatch) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sh &s = *reinterpret_cast<Sh *>(raw);
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    for (int i = lane; i < 9 * 4096; i += 32) s.inv[i] = (uint16_t)((i * 2654435761u >> 7) & 4095);
    // a plausible symbol: f = 1500, base = 4096*3 + 100
    AnsSymInfo si = ans_sym_info(1500 + lane * 37 % 700, 4096 * 3 + 100);
    const uint32_t inv_base = (uint32_t)__cvta_generic_to_shared(s.inv);
    s.stage[lane] = make_uint4(si.m, (si.w1 & 0xFF) | ((si.w1 >> 8) << 8), si.nf2, si.b2 + inv_base);
    __syncwarp();
    const uint32_t stg = (uint32_t)__cvta_generic_to_shared(s.stage), cap = (uint32_t)__cvta_generic_to_shared(s.cap);
    auto lookup = [](uint32_t a) -> uint32_t { return lds16(a); };
    uint32_t x = 0x130000;
    const uint4 fixed = s.stage[3];
    long long t0 = clock64();
    for (int b = 0; b < nbatch; b++) {
        uint4 r0 = lds128(stg + 31 * 16), r1 = lds128(stg + 30 * 16), r2 = lds128(stg + 29 * 16);
#pragma unroll
        for (int j = 31; j >= 0; --j) {
            uint4 st;
            if (MODE & 1) { st = r0; r0 = r1; r1 = r2; if (j >= 3) r2 = lds128(stg + (uint32_t)(j - 3) * 16u); }
            else st = fixed;
            uint32_t sp;
            ans_step_state(x, st.x, st.y, st.z, st.w, lookup, sp);
            if (x < 0x10000) x |= 0x10000;   // keep the state in range for this synthetic stream
            if (MODE & 2) sts32(cap + j * 4, sp);
        }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const int nb = 2000;
    for (int m = 0; m < 4; m++) {
        for (int smem : {(int)sizeof(Sh), 93 * 1024}) {
            auto run = [&](auto kern) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                for (int rep = 0; rep < 2; rep++) { kern<<<1, 256, smem>>>(d, nb); cudaDeviceSynchronize(); }
            };
            if (m == 0) run(k<0>); if (m == 1) run(k<1>); if (m == 2) run(k<2>); if (m == 3) run(k<3>);
            uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("prefetch=%d sts=%d smem=%dKB: %.2f cycles/step (err %s)\n", m & 1, (m >> 1) & 1, smem / 1024, (double)h[0] / (nb * 32.0), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
