// hydrium_b200/csrc/k_lf_group.cu
//
// Stage 3: the LF coefficient stream of each tile (section L) and the shape-constant sections
// A and B (see sections.cuh).  Replaces write_lf_group's modular sub-image coding
// (reference: encoder.c:539-629) including the whole prefix-code back end
// (entropy.c:502-524, 546-941, 1003-1034).
//
// One warp-sized CTA per tile: the 3 x nb quantised LF ints are staged in shared memory by all
// lanes; the residual / run-length / code construction / bit emission chain is <= 3072 symbols and
// inherently ordered, so lane 0 drives it over shared-memory scratch (prefix_coder.cuh) and the
// warp copies the finished bit string out.  Many such CTAs are resident per SM (~56 KB each), and
// the whole kernel runs concurrently with the HF tokeniser on another stream.
#include "kernels.h"
#include "lf_values.cuh"
#include "sections.cuh"

namespace hydb {

struct LfShared {
    PrefixWork work;
    uint32_t syms[3 * kMaxBlocks + 16];
    int32_t lfq[3 * kMaxBlocks];
    uint32_t bits[kLfBitsWords];
    uint32_t bitlen;
};

__global__ void __launch_bounds__(32)
k_lf_group(const TileDesc *__restrict__ tiles, const int32_t *__restrict__ lfq, uint32_t *__restrict__ lfbits,
           uint32_t *__restrict__ lfbitlen, uint32_t *__restrict__ tile_err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LfShared &s = *reinterpret_cast<LfShared *>(smem_raw);
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    const TileDesc t = tiles[tile];
    const uint32_t vbw = (t.w + 7) >> 3, vbh = (t.h + 7) >> 3;

    for (uint32_t i = lane; i < 3 * kMaxBlocks; i += 32)
        s.lfq[i] = lfq[(size_t)tile * 3 * kMaxBlocks + i];
    __syncwarp();
    if (lane == 0) {
        s.work.error = 0;
        BitSink bw;
        bw.init(s.bits, kLfBitsWords);
        LfStreamValues values{s.lfq, vbw, vbw * vbh};
        ps_encode_stream(s.work, s.syms, 3 * kMaxBlocks + 16, lf_stream_params(), 3 * vbw * vbh, values, bw);
        bw.flush_partial();
        s.bitlen = bw.bitlen();
        uint32_t err = s.work.error | (bw.overflow ? (uint32_t)kErrLfCapacity : 0u);
        if (err)
            atomicOr(&tile_err[tile], err);
        lfbitlen[tile] = bw.overflow ? 0 : s.bitlen;
    }
    __syncwarp();
    const uint32_t words = (s.bitlen + 31) >> 5;
    for (uint32_t i = lane; i < words && i < (uint32_t)kLfBitsWords; i += 32)
        lfbits[(size_t)tile * kLfBitsWords + i] = s.bits[i];
}

struct TemplShared {
    PrefixWork work;
    uint32_t syms[kSectionSymCap];
};

// one thread per section: block 0 builds section A (if asked), block 1 + k builds section B of shape k
__global__ void __launch_bounds__(32)
k_build_templates(Templates t, const uint32_t *__restrict__ shape_dims, uint32_t first_shape, int build_a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TemplShared &s = *reinterpret_cast<TemplShared *>(smem_raw);
    if (threadIdx.x != 0)
        return;
    s.work.error = 0;
    BitSink bw;
    if (blockIdx.x == 0) {
        if (!build_a)
            return;
        bw.init(t.words, kTemplWords);
        build_section_a(s.work, s.syms, bw);
        bw.flush_partial();
        t.bits[0] = (bw.overflow || s.work.error) ? 0xFFFFFFFFu : bw.bitlen();
    } else {
        const uint32_t k = blockIdx.x - 1, slot = 1 + first_shape + k;
        bw.init(t.words + (size_t)slot * kTemplWords, kTemplWords);
        build_section_b(s.work, s.syms, bw, shape_dims[2 * k], shape_dims[2 * k + 1]);
        bw.flush_partial();
        t.bits[slot] = (bw.overflow || s.work.error) ? 0xFFFFFFFFu : bw.bitlen();
    }
}

void launch_lf_group(const Workspace &ws, uint32_t ntiles, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_lf_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LfShared));
        configured = true;
    }
    k_lf_group<<<ntiles, 32, sizeof(LfShared), st>>>(ws.tiles, ws.lfq, ws.lfbits, ws.lfbitlen, ws.tile_err);
}

void launch_build_templates(const Templates &t, const uint32_t *d_shape_dims, uint32_t first_shape, uint32_t n_shapes,
                            bool build_a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_build_templates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TemplShared));
        configured = true;
    }
    k_build_templates<<<1 + n_shapes, 32, sizeof(TemplShared), st>>>(t, d_shape_dims, first_shape, build_a ? 1 : 0);
}

}  // namespace hydb
