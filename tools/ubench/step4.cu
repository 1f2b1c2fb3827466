// Faithful replica of k_ans_chain's chain-warp loop over a static ring (no helper), for trying
// source arrangements quickly.  Prints cycles/step per variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hydrium_b200/csrc/ans_chain.cuh"
using namespace hydb;

__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128nv(uint32_t a) { uint4 v; asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

struct Sh { uint16_t inv[9 * 4096]; uint4 stage[4][32]; uint32_t cap[4][32]; };

template <int V>
__global__ void k(uint64_t *out, int nbatch) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sh &s = *reinterpret_cast<Sh *>(raw);
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    for (int i = lane; i < 9 * 4096; i += 32) s.inv[i] = (uint16_t)((i * 2654435761u >> 7) & 4095);
    const uint32_t inv_base = (uint32_t)__cvta_generic_to_shared(s.inv);
    for (int r = 0; r < 4; r++) {
        auto fq = [](uint32_t l, int rr) { return 300u + ((l * 97u + (uint32_t)rr * 31u) % 3000u); };
        const uint32_t f = fq(lane, r), fn = lane ? fq(lane - 1, r) : fq(31, (r + 1) & 3);
        AnsSymInfo si = ans_sym_info(f, 4096 * (lane % 9) + (lane * 13 % (4096 - f)));
        s.stage[r][lane] = make_uint4(si.m, (si.w1 & 0xFF) | (fn << 8), si.nf2, si.b2 + inv_base);
    }
    __syncwarp();
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(s.stage), cap_base = (uint32_t)__cvta_generic_to_shared(s.cap);
    auto lookup = [](uint32_t a) -> uint32_t { return lds16(a); };
    uint32_t x = 0x13;
    long long t0 = clock64();
    for (int seq = 0; seq < nbatch; seq++) {
        const int slot = seq & 3;
        const uint32_t stg = stage_base + slot * 512, capb = cap_base + slot * 128;
        if (V == 0) {          // as in the kernel: volatile prefetch, depth 3
            uint4 r0 = lds128(stg + 31 * 16), r1 = lds128(stg + 30 * 16), r2 = lds128(stg + 29 * 16);
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                const uint4 st = r0; r0 = r1; r1 = r2;
                if (j >= 3) r2 = lds128(stg + (j - 3) * 16);
                uint32_t sp; ans_step_state(x, st.x, st.y, st.z, st.w, lookup, sp);
                sts32(capb + j * 4, sp);
            }
        } else if (V == 1) {   // non-volatile loads, compiler decides
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                const uint4 st = lds128nv(stg + j * 16);
                uint32_t sp; ans_step_state(x, st.x, st.y, st.z, st.w, lookup, sp);
                sts32(capb + j * 4, sp);
            }
        } else if (V == 2) {   // all 32 records loaded up front (register heavy)
            uint4 rec[32];
#pragma unroll
            for (int j = 0; j < 32; j++) rec[j] = lds128(stg + j * 16);
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                uint32_t sp; ans_step_state(x, rec[j].x, rec[j].y, rec[j].z, rec[j].w, lookup, sp);
                sts32(capb + j * 4, sp);
            }
        } else if (V == 3) {   // as V0 but no STS (upper bound on its cost)
            uint4 r0 = lds128(stg + 31 * 16), r1 = lds128(stg + 30 * 16), r2 = lds128(stg + 29 * 16);
            uint32_t acc = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                const uint4 st = r0; r0 = r1; r1 = r2;
                if (j >= 3) r2 = lds128(stg + (j - 3) * 16);
                uint32_t sp; ans_step_state(x, st.x, st.y, st.z, st.w, lookup, sp);
                acc ^= sp;
            }
            sts32(capb, acc);
        } else if (V == 5 || V == 6 || V == 7) {   // G records up front, states kept in registers, stored after the group
            constexpr int G = V == 5 ? 16 : (V == 6 ? 8 : 32);
#pragma unroll
            for (int h = 32 / G - 1; h >= 0; --h) {
                uint4 rec[G];
                uint32_t sp[G];
#pragma unroll
                for (int j = 0; j < G; j++) rec[j] = lds128(stg + (h * G + j) * 16);
#pragma unroll
                for (int j = G - 1; j >= 0; --j)
                    ans_step_state(x, rec[j].x, rec[j].y, rec[j].z, rec[j].w, lookup, sp[j]);
#pragma unroll
                for (int j = 0; j < G; j++) sts32(capb + (h * G + j) * 4, sp[j]);
            }
        } else if (V == 4) {   // 8 records ahead in two halves of 16
#pragma unroll
            for (int h = 1; h >= 0; --h) {
                uint4 rec[16];
#pragma unroll
                for (int j = 0; j < 16; j++) rec[j] = lds128(stg + (h * 16 + j) * 16);
#pragma unroll
                for (int j = 15; j >= 0; --j) {
                    uint32_t sp; ans_step_state(x, rec[j].x, rec[j].y, rec[j].z, rec[j].w, lookup, sp);
                    sts32(capb + (h * 16 + j) * 4, sp);
                }
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = x; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const int nb = 3000;
    const char *names[] = {"V0 kernel (volatile prefetch x3)", "V1 non-volatile per-step load", "V2 all 32 records up front", "V3 V0 without STS", "V4 two halves of 16 up front", "V5 groups of 16, states in regs", "V6 groups of 8, states in regs", "V7 group of 32, states in regs"};
    for (int v = 0; v < 8; v++) {
        if (v == 1 || v == 2) continue;
        auto run = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            for (int rep = 0; rep < 2; rep++) { kern<<<1, 64, sizeof(Sh)>>>(d, nb); cudaDeviceSynchronize(); }
        };
        if (v == 0) run(k<0>); if (v == 1) run(k<1>); if (v == 2) run(k<2>); if (v == 3) run(k<3>); if (v == 4) run(k<4>); if (v == 5) run(k<5>); if (v == 6) run(k<6>); if (v == 7) run(k<7>);
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-36s %.2f cycles/step (%s)\n", names[v], (double)h[0] / (nb * 32.0), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
