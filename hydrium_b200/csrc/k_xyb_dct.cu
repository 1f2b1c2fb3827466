// hydrium_b200/csrc/k_xyb_dct.cu
//
// Stage 1 of the tile pipeline: RGB samples -> XYB -> 8x8 forward DCT -> quantised coefficients.
// Replaces hyd_populate_xyb_buffer (reference: format.c:142-194), forward_dct (encoder.c:631-668),
// the HF quantisation loop (encoder.c:783-823) and the LF quantisation (encoder.c:573, 582).
//
// B200 mapping: one CTA per row of 32 varblocks of a tile (grid = 32 x tiles, 256 threads).
// Thread (block b, row r) converts its 8 pixels and runs the three row transforms in registers,
// rows are exchanged through a conflict-free padded shared-memory tile, thread (b, column t) runs
// the three column transforms, quantises, and the CTA writes 12 KB of int16 coefficients in scan
// order with coalesced 16-byte stores.  Nothing is a dense contraction that could use tensor
// cores without changing the result: the summation ORDER is part of the format (SURVEY.md
// Appendix B), every product and sum is a separate IEEE round-to-nearest op (no FMA).
#include "chain_util.cuh"
#include "common.cuh"
#include "kernels.h"
#include "tables.cuh"

namespace hydb {

__constant__ uint32_t c_cos_bits[7][8] = {HYDB_COS_BITS};
__constant__ uint8_t c_scan_index[64] = {HYDB_SCAN_INDEX};
__constant__ uint16_t c_hf_weights[3][64] = {HYDB_HF_WEIGHTS};

// ---- lookup tables (reference: format.c:15-36, 58-83) -------------------------------------
__device__ __forceinline__ float srgb_to_linear(float x) {
    if (x <= 0.0404482362771082f)
        return __fmul_rn(0.07739938080495357f, x);
    float t = __fadd_rn(0.72007737769f, __fmul_rn(0.2852804880f, x));
    t = __fadd_rn(-0.009982599f, __fmul_rn(x, t));
    return __fadd_rn(0.003094300919832f, __fmul_rn(x, t));
}

__device__ __forceinline__ float fast_cbrt(float x) {
    uint32_t zi = 0x548c39cbu - __float_as_uint(x) / 3u;
    float z = __uint_as_float(zi);
    // z *= c0 - c1 * x * z * z * z, evaluated left to right as ((((c1*x)*z)*z)*z)
    float t = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.534850249f, x), z), z), z);
    z = __fmul_rn(z, __fsub_rn(1.5015480449f, t));
    t = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.33333333f, x), z), z), z);
    z = __fmul_rn(z, __fsub_rn(1.333333985f, t));
    return __fdiv_rn(1.0f, z);
}

// reference: format.c:29-31
__device__ __forceinline__ float opsin_bias(float x) {
    return __fsub_rn(fast_cbrt(__fadd_rn(x, 0.0037930732552754493f)), 0.155954f);
}

// entry `idx` of the reference's bias table, computed instead of looked up (format.c:73-83)
__device__ __forceinline__ float bias_entry(uint32_t idx) {
    const float step16 = __fdiv_rn(1.0f, __fsub_rn(65536.0f, 1.0f));   // folded at compile time
    return opsin_bias(__fmul_rn((float)idx, step16));
}

__global__ void k_build_luts(uint16_t *lut8_srgb, uint16_t *lut8_lin, uint16_t *lut16_srgb, uint16_t *lut16_lin,
                             float *bias) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536)
        return;
    const float step16 = __fdiv_rn(1.0f, __fsub_rn(65536.0f, 1.0f));
    const float f16 = __fmul_rn((float)i, step16);
    {
        bias[i] = bias_entry(i);   // kept for the table tap of the parity tests
    }
    auto to_u16 = [](float x) -> uint16_t {
        int v = __float2int_rz(__fadd_rn(__fmul_rn(x, 65535.f), 0.5f));
        return (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
    };
    lut16_lin[i] = to_u16(f16);
    lut16_srgb[i] = to_u16(srgb_to_linear(f16));
    if (i < 256) {
        const float step8 = __fdiv_rn(1.0f, __fsub_rn(256.0f, 1.0f));
        const float f8 = __fmul_rn((float)i, step8);
        lut8_lin[i] = to_u16(f8);
        lut8_srgb[i] = to_u16(srgb_to_linear(f8));
    }
}

// ---- 8-point transform, reference summation order (encoder.c:641-648) ----------------------
__device__ __forceinline__ void dct8(const float (&v)[8], float (&o)[8]) {
    float dc = v[0];
#pragma unroll
    for (int n = 1; n < 8; n++)
        dc = __fadd_rn(dc, v[n]);
    o[0] = __fmul_rn(dc, 0.125f);
#pragma unroll
    for (int k = 1; k < 8; k++) {
        // the reference starts from +0.0f; 0.0f + p differs from p only in the sign of a zero,
        // which no consumer can observe (every consumer truncates to int)
        float acc = __fmul_rn(v[0], __uint_as_float(c_cos_bits[k - 1][0]));
#pragma unroll
        for (int n = 1; n < 8; n++)
            acc = __fadd_rn(acc, __fmul_rn(v[n], __uint_as_float(c_cos_bits[k - 1][n])));
        o[k] = acc;
    }
}

constexpr int kRowPad = 9;                  // 8 + 1: conflict-free row/column exchange
constexpr int kBlkPad = 8 * kRowPad;        // 72 floats per block and channel

template <typename Sample>
__device__ __forceinline__ void load_xyb(const TileDesc &t, const uint16_t *in_lut, const float *__restrict__ bias,
                                         uint32_t px0, uint32_t y, float (&X)[8], float (&Y)[8], float (&B)[8]) {
    const Sample *p0 = (const Sample *)t.plane[0];
    const Sample *p1 = (const Sample *)t.plane[1];
    const Sample *p2 = (const Sample *)t.plane[2];
    // packed 8-bit RGB whose 24-byte block rows are word aligned (the usual case): six 32-bit loads
    // per thread instead of twenty-four byte loads
    uint32_t pk[6];
    bool packed = false;
    if (sizeof(Sample) == 1) {
        packed = t.pixel_stride == 3 && p1 == p0 + 1 && p2 == p0 + 2 && px0 + 8 <= t.w && y < t.h &&
                 (((uintptr_t)p0 | (uintptr_t)t.row_stride) & 3u) == 0;
        if (packed) {
            const uint32_t *q = (const uint32_t *)(p0 + (int64_t)y * t.row_stride + (int64_t)px0 * 3);
#pragma unroll
            for (int k = 0; k < 6; k++)
                pk[k] = __ldg(q + k);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t px = px0 + i;
        float x = 0.0f, yy = 0.0f, b = 0.0f;   // zero padding of partial blocks (format.c:182-191)
        if (px < t.w && y < t.h) {
            uint32_t sr, sg, sb;
            if (sizeof(Sample) == 1 && packed) {
                sr = (pk[(3 * i) >> 2] >> (8 * ((3 * i) & 3))) & 0xFFu;
                sg = (pk[(3 * i + 1) >> 2] >> (8 * ((3 * i + 1) & 3))) & 0xFFu;
                sb = (pk[(3 * i + 2) >> 2] >> (8 * ((3 * i + 2) & 3))) & 0xFFu;
            } else {
                const int64_t o = (int64_t)y * t.row_stride + (int64_t)px * t.pixel_stride;
                sr = __ldg(p0 + o);
                sg = __ldg(p1 + o);
                sb = __ldg(p2 + o);
            }
            const uint32_t r = in_lut[sr];
            const uint32_t g = in_lut[sg];
            const uint32_t bl = in_lut[sb];
            // format.c:48-56.  The reference reads bias_lut[idx]; entry idx of that table is
            // opsin_bias((float)idx * (1 / 65535)) (format.c:73-83), which is recomputed here in
            // registers: about thirty FP32 instructions instead of a 32-way scattered gather from a
            // 256 KB table, which was half of this kernel's time (L1 wavefronts).
            const float l = bias_entry(((19661u * r + 40761u * g + 5112u * bl) >> 16) & 0xFFFFu);
            const float m = bias_entry(((15073u * r + 45350u * g + 5112u * bl) >> 16) & 0xFFFFu);
            const float s = bias_entry(((15953u * r + 13419u * g + 36163u * bl) >> 16) & 0xFFFFu);
            yy = __fmul_rn(__fadd_rn(l, m), 0.5f);
            x = __fsub_rn(yy, m);
            b = __fsub_rn(s, yy);
        }
        X[i] = x;
        Y[i] = yy;
        B[i] = b;
    }
}

// ---- bulk-asynchronous staging of the CTA's 8 pixel rows (TMA, cp.async.bulk -> SASS UBLKCP) -------------
// One thread arms an mbarrier with the byte count and issues one bulk copy per pixel row (global -> shared,
// completion counted on the mbarrier); the copy engine moves the rows while the CTA loads its tables, and
// every thread then picks its samples out of shared memory with one LDS each (no packed-word unpacking).
// Applies to interleaved u8 / u16 tiles whose rows are 16-byte aligned and a multiple of 16 bytes long
// (every tile of an image whose row pitch is a multiple of 16 bytes: 768 / 1536 / 2048-byte tile rows);
// everything else takes the direct-load path above.
constexpr int kStageRowMax = 2048;                 // RGBA16
constexpr int kStagePitch = kStageRowMax + 16;     // rows start on different banks

__device__ __forceinline__ bool stage_applicable(const TileDesc &t, uint32_t item, uint32_t &row_bytes) {
    const uint8_t *p0 = (const uint8_t *)t.plane[0];
    row_bytes = t.w * (uint32_t)t.pixel_stride * item;
    return (t.pixel_stride == 3 || t.pixel_stride == 4) && (const uint8_t *)t.plane[1] == p0 + item &&
           (const uint8_t *)t.plane[2] == p0 + 2 * item && t.row_stride > 0 && row_bytes <= (uint32_t)kStageRowMax &&
           (row_bytes & 15u) == 0 && (((uintptr_t)p0 | (uintptr_t)(t.row_stride * (int64_t)item)) & 15u) == 0;
}

__device__ __forceinline__ void stage_issue(const TileDesc &t, uint32_t item, uint32_t row_bytes, uint32_t y0, uint32_t rows,
                                            uint8_t *s_in, uint64_t *mbar) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows * row_bytes) : "memory");
    const uint8_t *src = (const uint8_t *)t.plane[0] + (int64_t)y0 * t.row_stride * (int64_t)item;
    for (uint32_t r = 0; r < rows; r++) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_in + r * kStagePitch);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src + (int64_t)r * t.row_stride * (int64_t)item), "r"(row_bytes), "r"(bar) : "memory");
    }
}

template <typename Sample>
__device__ __forceinline__ void load_xyb_staged(const TileDesc &t, const uint16_t *in_lut, const uint8_t *s_in, uint32_t px0,
                                                uint32_t r, uint32_t y, float (&X)[8], float (&Y)[8], float (&B)[8]) {
    const Sample *row = (const Sample *)(s_in + r * kStagePitch);
    const uint32_t ps = (uint32_t)t.pixel_stride;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t px = px0 + i;
        float x = 0.0f, yy = 0.0f, b = 0.0f;   // zero padding of partial blocks (format.c:182-191)
        if (px < t.w && y < t.h) {
            const uint32_t rr = in_lut[row[px * ps]], g = in_lut[row[px * ps + 1]], bl = in_lut[row[px * ps + 2]];
            const float l = bias_entry(((19661u * rr + 40761u * g + 5112u * bl) >> 16) & 0xFFFFu);   // format.c:48-56
            const float m = bias_entry(((15073u * rr + 45350u * g + 5112u * bl) >> 16) & 0xFFFFu);
            const float s = bias_entry(((15953u * rr + 13419u * g + 36163u * bl) >> 16) & 0xFFFFu);
            yy = __fmul_rn(__fadd_rn(l, m), 0.5f);
            x = __fsub_rn(yy, m);
            b = __fsub_rn(s, yy);
        }
        X[i] = x;
        Y[i] = yy;
        B[i] = b;
    }
}

// HYD_FLOAT32 samples: no tables, the transfer curve and the opsin mix are evaluated per pixel in the
// reference's operation order (format.c:38-46, 111-140).  Returns 0, or kErrNonFinite for a NaN / Inf sample,
// or kErrNegative when an opsin mix (plus bias) is negative: the reference's bit-hack cube root then yields
// NaN and its conversion to int is undefined in C (INT_MIN on x86) -- there is nothing to be bit-exact with.
__device__ __forceinline__ uint32_t load_xyb_f32(const TileDesc &t, bool linear, uint32_t px0, uint32_t y, float (&X)[8],
                                                 float (&Y)[8], float (&B)[8]) {
    const float *p0 = (const float *)t.plane[0];
    const float *p1 = (const float *)t.plane[1];
    const float *p2 = (const float *)t.plane[2];
    bool finite = true, nonneg = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t px = px0 + i;
        float x = 0.0f, yy = 0.0f, b = 0.0f;
        if (px < t.w && y < t.h) {
            const int64_t o = (int64_t)y * t.row_stride + (int64_t)px * t.pixel_stride;
            float r = __ldg(p0 + o), g = __ldg(p1 + o), bl = __ldg(p2 + o);
            finite = finite && isfinite(r) && isfinite(g) && isfinite(bl);
            if (!linear) {
                r = srgb_to_linear(r);
                g = srgb_to_linear(g);
                bl = srgb_to_linear(bl);
            }
            // (c0 * r + c1 * g) + c2 * b, every product and sum rounded separately
            auto mix = [&](float c0, float c1, float c2) {
                return __fadd_rn(__fadd_rn(__fmul_rn(c0, r), __fmul_rn(c1, g)), __fmul_rn(c2, bl));
            };
            const float ml = mix(0.3f, 0.622f, 0.078f), mm = mix(0.23f, 0.692f, 0.078f), ms = mix(0.243423f, 0.204767f, 0.55181f);
            const float kBias = 0.0037930732552754493f;   // what opsin_bias adds before the cube root
            nonneg = nonneg && !(__fadd_rn(ml, kBias) < 0.0f) && !(__fadd_rn(mm, kBias) < 0.0f) && !(__fadd_rn(ms, kBias) < 0.0f);
            const float l = opsin_bias(ml);
            const float m = opsin_bias(mm);
            const float s = opsin_bias(ms);
            yy = __fmul_rn(__fadd_rn(l, m), 0.5f);
            x = __fsub_rn(yy, m);
            b = __fsub_rn(s, yy);
        }
        X[i] = x;
        Y[i] = yy;
        B[i] = b;
    }
    return (finite ? 0u : (uint32_t)kErrNonFinite) | ((nonneg || !finite) ? 0u : (uint32_t)kErrNegative);
}

__global__ void __launch_bounds__(256)
k_xyb_dct_quant(const TileDesc *__restrict__ tiles, LutSet luts, int16_t *__restrict__ coef,
                uint16_t *__restrict__ nzinfo, int32_t *__restrict__ lfq, uint32_t *__restrict__ tile_err,
                float *__restrict__ dbg_xyb, float *__restrict__ dbg_dct, bool use_tma) {
    __shared__ __align__(16) float s_rows[3 * 32 * kBlkPad];   // 27,648 B; first the staged input rows (8 x 2,064 B)
    __shared__ __align__(16) int16_t s_q[32 * 3 * 64];    // 12,288 B
    __shared__ __align__(8) uint64_t s_mbar;
    static_assert(8 * kStagePitch <= (int)sizeof(float) * 3 * 32 * kBlkPad, "staged rows alias the row-exchange tile");
    __shared__ uint16_t s_lut8[256];
    __shared__ float s_w[3 * 64];
    __shared__ uint8_t s_scan[64];

    const uint32_t tile = blockIdx.y, by = blockIdx.x;
    const TileDesc t = tiles[tile];
    const uint32_t vbw = (t.w + 7) >> 3, vbh = (t.h + 7) >> 3;
    if (by >= vbh || (t.flags & kTilePrefix))
        return;
    const uint32_t tid = threadIdx.x, b = tid >> 3, r = tid & 7;
    const bool fmt16 = (t.flags & kTileFmt16) != 0, fmt32 = (t.flags & kTileFmtF32) != 0;
    const bool linear = (t.flags & kTileLinear) != 0;

    // bulk-asynchronous staging of this CTA's pixel rows (see stage_applicable); HYDRIUM_B200_TMA=0 turns it off
    uint8_t *s_in = reinterpret_cast<uint8_t *>(s_rows);
    uint32_t row_bytes = 0;
    const uint32_t item = fmt16 ? 2u : 1u;
    const bool staged = use_tma && !fmt32 && stage_applicable(t, item, row_bytes);
    if (staged && tid == 0) {
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t rows = t.h - by * 8 < 8 ? t.h - by * 8 : 8;
        stage_issue(t, item, row_bytes, by * 8, rows, s_in, &s_mbar);
    }
    if (!fmt16 && !fmt32)
        s_lut8[tid] = (linear ? luts.lut8_lin : luts.lut8_srgb)[tid];
    if (tid < 192)
        s_w[tid] = (float)c_hf_weights[tid >> 6][tid & 63];
    if (tid < 64)
        s_scan[tid] = c_scan_index[tid];
    __syncthreads();

    // ---- colour transform + row pass --------------------------------------------------------
    float v[3][8];
    if (b < vbw) {
        if (fmt32) {
            const uint32_t bad = load_xyb_f32(t, linear, b * 8, by * 8 + r, v[0], v[1], v[2]);
            if (bad)
                atomicOr(&tile_err[tile], bad);
        } else if (staged) {
            mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar), 0);
            if (fmt16)
                load_xyb_staged<uint16_t>(t, linear ? luts.lut16_lin : luts.lut16_srgb, s_in, b * 8, r, by * 8 + r, v[0], v[1], v[2]);
            else
                load_xyb_staged<uint8_t>(t, s_lut8, s_in, b * 8, r, by * 8 + r, v[0], v[1], v[2]);
        } else if (fmt16)
            load_xyb<uint16_t>(t, linear ? luts.lut16_lin : luts.lut16_srgb, luts.bias, b * 8, by * 8 + r, v[0], v[1], v[2]);
        else
            load_xyb<uint8_t>(t, s_lut8, luts.bias, b * 8, by * 8 + r, v[0], v[1], v[2]);
        if (dbg_xyb) {
            float *d = dbg_xyb + ((size_t)tile * 65536 + (size_t)(by * 8 + r) * 256 + b * 8) * 3;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                d[i * 3 + 0] = v[0][i];
                d[i * 3 + 1] = v[1][i];
                d[i * 3 + 2] = v[2][i];
            }
        }
    }
    if (staged)
        __syncthreads();   // the staged rows live where the row-exchange tile goes: everyone has read its pixels
    if (b < vbw) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float o[8];
            dct8(v[c], o);
            float *dst = s_rows + (c * 32 + b) * kBlkPad + r * kRowPad;
#pragma unroll
            for (int k = 0; k < 8; k++)
                dst[k] = o[k];
        }
    }
    __syncthreads();

    // ---- column pass + quantisation ---------------------------------------------------------
    // thread (b, t): horizontal frequency kh = t, produces vertical frequencies kv = 0..7
    if (b < vbw) {
        const uint32_t kh = r;
        bool in_range = true;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float v[8], o[8];
            const float *src = s_rows + (c * 32 + b) * kBlkPad + kh;
#pragma unroll
            for (int n = 0; n < 8; n++)
                v[n] = src[n * kRowPad];
            dct8(v, o);
            if (dbg_dct) {
                // stored transposed like the reference: position (row = kh, col = kv)
                float *d = dbg_dct + ((size_t)tile * 65536 + (size_t)(by * 8 + kh) * 256 + b * 8) * 3 + c;
#pragma unroll
                for (int kv = 0; kv < 8; kv++)
                    d[kv * 3] = o[kv];
            }
            int16_t *qd = s_q + (b * 3 + c) * 64;
#pragma unroll
            for (int kv = 0; kv < 8; kv++) {
                const uint32_t j = s_scan[kv * 8 + kh];
                if (j == 0) {
                    // LF: trunc(dc * {8192, 1024, 512}) (encoder.c:573, 582)
                    const float scale = c == 0 ? 8192.f : (c == 1 ? 1024.f : 512.f);
                    const float lf = __fmul_rn(o[kv], scale);
                    in_range = in_range && fabsf(lf) < 2147483648.0f;   // beyond it the reference's cast is undefined
                    lfq[((size_t)tile * 3 + c) * kMaxBlocks + by * kBlocksPerRow + b] = __float2int_rz(lf);
                    qd[0] = 0;
                } else {
                    // HF: trunc((f * w) * 5), dead zone |q| < 2 -> 0 (encoder.c:808-810)
                    int q = __float2int_rz(__fmul_rn(__fmul_rn(o[kv], s_w[c * 64 + j]), 5.0f));
                    if (q > -2 && q < 2)
                        q = 0;
                    // the reference keeps 32 bits here (encoder.c:808); coefficients, residues and tokens
                    // downstream are sized for 16 (every u8 / u16 image, floats within a few thousand of
                    // [0, 1]).  Anything larger is refused loudly instead of wrapping.
                    in_range = in_range && q >= -32768 && q <= 32767;
                    qd[j] = (int16_t)q;
                }
            }
        }
        if (!in_range)
            atomicOr(&tile_err[tile], (uint32_t)kErrRange);
    } else {
        // blocks beyond the tile edge: keep the staging buffer defined
        for (int i = r; i < 3 * 64; i += 8)
            s_q[b * 3 * 64 + i] = 0;
    }
    __syncthreads();

    // ---- per block and channel: number of non-zeros and scan index of the last one -----------
    {
        const uint32_t warp = tid >> 5, lane = tid & 31;
        for (uint32_t bc = warp; bc < 96; bc += 8) {
            const int16_t *qd = s_q + bc * 64;
            const uint32_t lo = __ballot_sync(0xFFFFFFFFu, qd[lane] != 0);
            const uint32_t hi = __ballot_sync(0xFFFFFFFFu, qd[lane + 32] != 0);
            if (lane == 0) {
                const uint32_t nz = __popc(lo) + __popc(hi);
                const uint32_t last = hi ? 63 - __clz(hi) : (lo ? 31 - __clz(lo) : 0);
                const uint32_t blk = bc / 3, c = bc - blk * 3;
                if (blk < vbw)
                    nzinfo[((size_t)tile * kMaxBlocks + by * kBlocksPerRow + blk) * 3 + c] = (uint16_t)(nz | (last << 8));
            }
        }
    }
    // ---- coalesced store of the coefficient rows ----------------------------------------------
    {
        const uint4 *src = (const uint4 *)s_q;
        uint4 *dst = (uint4 *)(coef + ((size_t)tile * kMaxBlocks + by * kBlocksPerRow) * 3 * 64);
        for (uint32_t i = tid; i < 32 * 3 * 64 * 2 / 16; i += 256)
            dst[i] = src[i];
    }
}

void launch_build_luts(uint16_t *lut8_srgb, uint16_t *lut8_lin, uint16_t *lut16_srgb, uint16_t *lut16_lin, float *bias,
                       cudaStream_t st) {
    k_build_luts<<<65536 / 256, 256, 0, st>>>(lut8_srgb, lut8_lin, lut16_srgb, lut16_lin, bias);
}

void launch_xyb_dct_quant(const Workspace &ws, const LutSet &luts, uint32_t ntiles, cudaStream_t st) {
    prefer_max_shared(k_xyb_dct_quant);
    static const bool use_tma = [] { const char *e = getenv("HYDRIUM_B200_TMA"); return !(e && e[0] == '0'); }();
    k_xyb_dct_quant<<<dim3(kBlocksPerRow, ntiles), 256, 0, st>>>(ws.tiles, luts, ws.coef, ws.nzinfo, ws.lfq, ws.tile_err,
                                                                 ws.dbg_xyb, ws.dbg_dct, use_tma);
}

}  // namespace hydb
