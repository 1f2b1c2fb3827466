// Dependent-chain latency microbenchmarks for the ops on the rANS critical path (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <int MODE>
__global__ void k(uint64_t *out, uint32_t seed, uint32_t mulc, uint32_t shc) {
    __shared__ uint32_t tab[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (MODE == 0 || MODE == 1) ? ((i * 4) & 0xFFFC) | (((i * 4 + 2) & 0xFFFC) << 16) : (i * 4 & 0x7FFC);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    uint32_t x = seed;
    uint64_t acc = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (MODE == 0) {            // LDS.U16 pointer chase (value = next byte offset)
            x = lds16(base + (x & 0x3FFE));
        } else if (MODE == 1) {     // LDS.32 pointer chase
            x = lds32(base + (x & 0x3FFC)) & 0xFFFF;
        } else if (MODE == 2) {     // IMAD.WIDE.U32 + take hi
            uint64_t t = (uint64_t)x * mulc + ((uint64_t)x << 32);
            x = (uint32_t)(t >> 32) | 1;
        } else if (MODE == 3) {     // IMAD.WIDE + variable 64-bit shift
            uint64_t t = (uint64_t)x * mulc + ((uint64_t)x << 32);
            x = (uint32_t)(t >> (shc & 63)) | 0x10000;
        } else if (MODE == 4) {     // IMAD (32-bit)
            x = x * mulc + seed;
        } else if (MODE == 5) {     // LOP3
            x = (x ^ mulc) | (x & seed);
            asm volatile("" : "+r"(x));
        } else if (MODE == 6) {     // the full step shape: LOP3 -> IMAD.WIDE -> SHF64 -> IMAD -> LDS.U16
            uint64_t t = (uint64_t)x * mulc + ((uint64_t)x << 32);
            uint32_t q = (uint32_t)(t >> (shc & 63));
            uint32_t slot = lds16(base + ((q * 6u + 2u * x) & 0x3FFE));
            x = (q << 12) | slot | 0x10000;
        } else if (MODE == 7) {     // same with 32-bit table load
            uint64_t t = (uint64_t)x * mulc + ((uint64_t)x << 32);
            uint32_t q = (uint32_t)(t >> (shc & 63));
            uint32_t slot = lds32(base + ((q * 6u + 2u * x) & 0x3FFC)) & 0xFFF;
            x = (q << 12) | slot | 0x10000;
        } else if (MODE == 8) {     // mul.hi + shift (32-bit reciprocal, no wide add)
            uint32_t q = __umulhi(x, mulc) >> (shc & 31);
            x = q | 0x10000;
        }
        acc += x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint64_t)(t1 - t0); out[1] = acc; }
}

int main() {
    uint64_t *d; cudaMalloc(&d, 16);
    const char *names[] = {"LDS.U16 chase", "LDS.32 chase (+LOP)", "IMAD.WIDE(+LOP)", "IMAD.WIDE+SHF.R.U64(+LOP)", "IMAD", "LOP3 x2", "full step (u16 table)", "full step (u32 table)", "IMAD.HI+SHF(+LOP)"};
    for (int m = 0; m < 9; m++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
            case 0: k<0><<<1, 32>>>(d, 4, 0x9E3779B1u, 37); break;
            case 1: k<1><<<1, 32>>>(d, 4, 0x9E3779B1u, 37); break;
            case 2: k<2><<<1, 32>>>(d, 12345, 0x9E3779B1u, 37); break;
            case 3: k<3><<<1, 32>>>(d, 12345, 0x9E3779B1u, 37); break;
            case 4: k<4><<<1, 32>>>(d, 12345, 0x9E3779B1u, 37); break;
            case 5: k<5><<<1, 32>>>(d, 12345, 0x9E3779B1u, 37); break;
            case 6: k<6><<<1, 32>>>(d, 0x130000, 0x9E3779B1u, 40); break;
            case 7: k<7><<<1, 32>>>(d, 0x130000, 0x9E3779B1u, 40); break;
            case 8: k<8><<<1, 32>>>(d, 0x130000, 0x9E3779B1u, 5); break;
            }
            cudaDeviceSynchronize();
        }
        uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-30s %.2f cycles/iter\n", names[m], (double)h[0] / N);
    }
    return 0;
}
