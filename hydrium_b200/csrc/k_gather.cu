// hydrium_b200/csrc/k_gather.cu
//
// Stage 5: compaction of the per-tile frames into one contiguous, ordered codestream in HBM
// (the role of hyd_flush's memcpy, reference: libhydrium.c:147-166, for a whole batch), plus the
// closed-form synthetic image generator used by the benchmarks.
#include "kernels.h"

namespace hydb {

// exclusive scan of frame_len[0..n) into out_off[0..n], single CTA (n <= 65536 per batch)
__global__ void __launch_bounds__(1024)
k_frame_offsets(const uint32_t *__restrict__ frame_len, uint64_t *__restrict__ out_off, uint32_t n) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + tid;
        const uint64_t v = i < n ? frame_len[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d)
                incl += u;
        }
        if (lane == 31)
            s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint64_t w = s_warp[lane];
            uint64_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= (uint32_t)d)
                    wi += u;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        const uint64_t excl = carry + s_warp[warp] + incl - v;
        if (i < n)
            out_off[i] = excl;
        __syncthreads();
        if (tid == 1023)
            s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0)
        out_off[n] = s_carry;
}

// one CTA per frame: byte copy with arbitrary source / destination alignment
__global__ void __launch_bounds__(256)
k_gather_frames(const uint8_t *__restrict__ slab, const uint32_t *__restrict__ frame_off,
                const uint32_t *__restrict__ frame_len, const uint64_t *__restrict__ out_off,
                uint8_t *__restrict__ out, uint64_t out_cap, uint64_t base, uint32_t *overflow) {
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t len = frame_len[tile];
    if (!len)
        return;
    const uint64_t d0 = base + out_off[tile];
    if (d0 + len > out_cap) {
        if (tid == 0)
            atomicOr(overflow, 1u);
        return;
    }
    const uint8_t *src = slab + (size_t)tile * kSlabBytes + frame_off[tile];
    uint8_t *dst = out + d0;
    // head: bring dst to 4-byte alignment
    const uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    const uint32_t h = head < len ? head : len;
    if (tid < h)
        dst[tid] = src[tid];
    const uint32_t body = (len - h) >> 2;   // words
    const uint8_t *s2 = src + h;
    uint32_t *d2 = reinterpret_cast<uint32_t *>(dst + h);
    const uint32_t mis = (uint32_t)((uintptr_t)s2 & 3);
    const uint32_t *sa = reinterpret_cast<const uint32_t *>(s2 - mis);
    for (uint32_t i = tid; i < body; i += 256) {
        uint32_t v;
        if (mis == 0) {
            v = sa[i];
        } else {
            const uint32_t lo = sa[i], hi = sa[i + 1];   // hi stays inside the slab: frames end before its last word
            v = __funnelshift_r(lo, hi, mis * 8);
        }
        d2[i] = v;
    }
    const uint32_t done = h + body * 4;
    if (tid < len - done)
        dst[done + tid] = src[done + tid];
}

void launch_gather(const Workspace &ws, uint32_t ntiles, uint8_t *out, uint64_t out_cap, uint64_t base,
                   uint32_t *d_overflow, cudaStream_t st) {
    prefer_max_shared(k_frame_offsets);
    prefer_max_shared(k_gather_frames);
    k_frame_offsets<<<1, 1024, 0, st>>>(ws.frame_len, ws.out_off, ntiles);
    k_gather_frames<<<ntiles, 256, 0, st>>>(ws.slab, ws.frame_off, ws.frame_len, ws.out_off, out, out_cap, base, d_overflow);
}

// last kernel of an asynchronous job: {bytes gathered, OR of the tiles' error bits | overflow << 31} into
// the job's page-locked host record (written by the device, read by the host once the job's event fired)
__global__ void __launch_bounds__(256)
k_job_result(const uint32_t *__restrict__ tile_err, uint32_t n, const uint64_t *__restrict__ total,
             const uint32_t *__restrict__ overflow, uint64_t *__restrict__ h_res) {
    __shared__ uint32_t s_err;
    if (threadIdx.x == 0)
        s_err = 0;
    __syncthreads();
    uint32_t e = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 256)
        e |= tile_err[i];
    if (e)
        atomicOr(&s_err, e);
    __syncthreads();
    if (threadIdx.x == 0) {
        h_res[0] = *total;
        h_res[1] = (uint64_t)((s_err & 0x7FFFFFFFu) | (*overflow ? 0x80000000u : 0u));
    }
}
void launch_job_result(const uint32_t *tile_err, uint32_t n, const uint64_t *total, const uint32_t *overflow, uint64_t *h_res,
                       cudaStream_t st) {
    k_job_result<<<1, 256, 0, st>>>(tile_err, n, total, overflow, h_res);
}

// ---- multi-GPU gather over peer memory ------------------------------------------------------------
// Every rank's k_gather_frames writes its span straight into its REGION of a buffer that lives on the
// gathering rank (peer pointer over NVLink, opened with CUDA IPC): region r = [8-byte length, pad to
// 256][span bytes].  Once all ranks are done this kernel closes the gaps: spans in rank order into one
// contiguous stream.  grid = (chunks, regions); each CTA re-derives its region's offset from the
// headers (regions <= 64).
constexpr uint32_t kRegionHeader = 256;
__global__ void __launch_bounds__(256)
k_compact_regions(const uint8_t *__restrict__ regions, uint32_t nregions, uint64_t region_stride, uint8_t *__restrict__ out,
                  uint64_t out_cap, uint64_t *__restrict__ total_out, uint32_t *overflow) {
    const uint32_t r = blockIdx.y, tid = threadIdx.x;
    uint64_t off = 0, len = 0, total = 0;
    for (uint32_t k = 0; k < nregions; k++) {
        const uint64_t l = *reinterpret_cast<const uint64_t *>(regions + (uint64_t)k * region_stride);
        if (k < r)
            off += l;
        if (k == r)
            len = l;
        total += l;
    }
    if (r == 0 && blockIdx.x == 0 && tid == 0)
        *total_out = total;
    if (len + 16 > region_stride - kRegionHeader || off + len > out_cap) {
        if (tid == 0)
            atomicOr(overflow, 1u);
        return;
    }
    const uint8_t *src = regions + (uint64_t)r * region_stride + kRegionHeader;   // 16-byte aligned
    uint8_t *dst = out + off;
    // destination words are stored whole; the source is read as aligned words and funnel-shifted
    const uint32_t h = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    const uint64_t head = h < len ? h : len;
    if (blockIdx.x == 0 && tid < head)
        dst[tid] = src[tid];
    const uint64_t nw = (len - head) >> 2;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
    const uint64_t per = (nw + gridDim.x - 1) / gridDim.x;
    const uint64_t k0 = (uint64_t)blockIdx.x * per, k1 = k0 + per < nw ? k0 + per : nw;
    if (h == 0) {
        for (uint64_t k = k0 + tid; k < k1; k += 256)
            dw[k] = sw[k];
    } else {
        for (uint64_t k = k0 + tid; k < k1; k += 256)   // sw[k + 1] stays inside the region: len + 16 <= its capacity
            dw[k] = __funnelshift_r(sw[k], sw[k + 1], h * 8);
    }
    const uint64_t done = head + nw * 4;
    if (blockIdx.x == gridDim.x - 1 && tid < len - done)
        dst[done + tid] = src[done + tid];
}

__global__ void k_store_u64(uint64_t *dst, uint64_t v) { *dst = v; }
void launch_store_u64(uint64_t *dst, uint64_t v, cudaStream_t st) { k_store_u64<<<1, 1, 0, st>>>(dst, v); }

void launch_compact_regions(const uint8_t *regions, uint32_t nregions, uint64_t region_stride, uint8_t *out, uint64_t out_cap,
                            uint64_t *d_total, uint32_t *d_overflow, cudaStream_t st) {
    k_compact_regions<<<dim3(148, nregions), 256, 0, st>>>(regions, nregions, region_stride, out, out_cap, d_total, d_overflow);
}

// ---- synthetic input (hydrium_b200/synth.py is the numpy twin) ----------
__device__ __forceinline__ uint32_t mix32(uint32_t v) {
    v ^= v >> 16;
    v *= 0x7feb352du;
    v ^= v >> 15;
    v *= 0x846ca68bu;
    v ^= v >> 16;
    return v;
}

template <typename Sample>
__global__ void k_synth_fill(Sample *dst, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0,
                             uint32_t full_w, uint32_t full_h, uint32_t seed, int smooth) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)width * height;
    if (idx >= total)
        return;
    const uint32_t lx = (uint32_t)(idx % width), ly = (uint32_t)(idx / width);
    const uint32_t x = lx + x0, y = ly + y0;
    const uint32_t maxv = sizeof(Sample) == 1 ? 255u : 65535u;
    const uint64_t dx = full_w > 1 ? full_w - 1 : 1, dy = full_h > 1 ? full_h - 1 : 1;
    const uint64_t dxy = (uint64_t)full_w + full_h > 2 ? (uint64_t)full_w + full_h - 2 : 1;
    const int64_t basev[3] = {(int64_t)((uint64_t)x * maxv / dx), (int64_t)((uint64_t)y * maxv / dy),
                              (int64_t)(((uint64_t)x + y) * maxv / dxy)};
    const uint32_t hx = x * 0x9E3779B1u, hy = mix32(y + 0x7F4A7C15u);
#pragma unroll
    for (uint32_t c = 0; c < 3; c++) {
        const uint32_t h = mix32(hx ^ hy ^ (c * 0x85EBCA6Bu) ^ seed);
        int64_t n = sizeof(Sample) == 1 ? (int64_t)((h >> 24) & 31u) - 16 : (int64_t)((h >> 16) & 0x1FFFu) - 4096;
        if (smooth)
            n = n >= 0 ? n / 8 : -((-n + 7) / 8);   // floor division, as numpy's //
        int64_t v = basev[c] + n;
        v = v < 0 ? 0 : (v > (int64_t)maxv ? (int64_t)maxv : v);
        dst[idx * 3 + c] = (Sample)v;
    }
}

void launch_synth_fill(void *dst, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0, uint32_t full_w,
                       uint32_t full_h, int bits, uint32_t seed, int smooth, cudaStream_t st) {
    const uint64_t total = (uint64_t)width * height;
    const uint32_t blocks = (uint32_t)((total + 255) / 256);
    if (bits == 8)
        k_synth_fill<uint8_t><<<blocks, 256, 0, st>>>((uint8_t *)dst, width, height, x0, y0, full_w, full_h, seed, smooth);
    else
        k_synth_fill<uint16_t><<<blocks, 256, 0, st>>>((uint16_t *)dst, width, height, x0, y0, full_w, full_h, seed, smooth);
}

}  // namespace hydb
