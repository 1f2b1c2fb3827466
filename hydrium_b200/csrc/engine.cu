// hydrium_b200/csrc/engine.cu
//
// The thin C-ABI CUDA layer: device workspace, kernel sequencing on CUDA streams, and the hydb_*
// entry points of include/hydrium_b200.h.  The nine libhydrium entry points live in hyd_api.c
// (portable C) and call into this file.
//
// Pipeline per batch of tiles (all asynchronous on the engine stream `st`; the LF coder runs on a
// second stream, concurrent with the HF tokeniser, and is joined before the ANS kernel):
//
//   st : [descs H2D] -> k_xyb_dct_quant -+-> k_hf_tokens ------+-> k_ans_encode -> k_frame_offsets -> k_gather_frames
//   st2:                                  +-> k_lf_group -------+
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/hydrium_b200.h"
#include "headers.cuh"
#include "kernels.h"
#include "sections.cuh"

using namespace hydb;

// NVTX range over the host-side enqueue of one stage (nsys / ncu --nvtx show the pipeline by name)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct HydbEngine {
    int device = 0;
    cudaStream_t st = nullptr, st2 = nullptr;
    cudaEvent_t ev_front = nullptr, ev_lf = nullptr;
    uint32_t max_batch = 0;
    Workspace ws{};
    LutSet luts{};
    uint16_t *lut8_srgb = nullptr, *lut8_lin = nullptr, *lut16_srgb = nullptr, *lut16_lin = nullptr;
    float *bias = nullptr;
    Templates templ{};
    std::vector<uint32_t> shapes;   // (vbw << 16) | vbh, index = shape id
    uint32_t *d_shape_dims = nullptr;
    uint32_t *d_overflow = nullptr;
    uint64_t *d_small = nullptr;    // [4] scratch for single-result kernels (compact_regions)
    uint64_t *h_small = nullptr;    // pinned twin
    uint32_t *h_err = nullptr;      // pinned [max_batch + 1]; last entry = gather overflow flag
    uint64_t *h_total = nullptr;    // pinned [1]
    uint32_t last_n = 0;
    uint64_t last_base = 0;
    bool taps = false;
    uint64_t launches = 0;
    std::string error;
    std::vector<TileDesc> h_tiles;
    // band pipeline of the host path: H2D of band k+1 overlaps the kernels of band k
    static constexpr int kBands = 4;
    cudaStream_t band_st[kBands] = {nullptr, nullptr, nullptr, nullptr}, band_st2[kBands] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t band_front[kBands] = {nullptr, nullptr, nullptr, nullptr}, band_lf[kBands] = {nullptr, nullptr, nullptr, nullptr},
                band_done[kBands] = {nullptr, nullptr, nullptr, nullptr}, band_h2d[kBands] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_st = nullptr;   // host path: the bands' pixels cross PCIe on ONE stream, in band order
    cudaEvent_t ev_desc = nullptr;
    // grow-only device buffers behind hydb_encode_image_host
    void *host_in = nullptr;
    uint8_t *host_out = nullptr;
    size_t host_in_cap = 0, host_out_cap = 0;
    // one-frame head assembly (hydb_oneframe_finish): device arena {info | scratch | out} and its page-locked mirror, grow-only
    void *of_dev = nullptr;
    size_t of_dev_cap = 0;
    uint8_t *of_host = nullptr;
    size_t of_host_cap = 0;
    // bits of the HF context map's stream for of_ctx_n presets (k_oneframe_finish): it depends on nothing else
    uint32_t *of_ctx = nullptr;
    uint32_t of_ctx_cap_words = 0, of_ctx_n = 0, of_ctx_bits = 0;
    // ... and of the TOC permutation's stream for the geometry and send order in of_perm_key
    uint32_t *of_perm = nullptr;
    uint32_t of_perm_cap_words = 0, of_perm_bits = 0;
    std::vector<uint32_t> of_perm_key;
    // optional per-stage timing with CUDA events on the launching streams
    bool timing = false, timed_pending = false;
    cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, lev[2] = {nullptr, nullptr};
    double stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};   // xyb_dct, hf_tokens, ans_chain, ans_pack(+LF wait), gather, lf_group, batches
    // asynchronous jobs (hydb_engine_submit_*): each owns a stream pair, a range of workspace slots
    // and a page-locked result record the last kernel of the job writes directly
    struct Job {
        cudaStream_t st = nullptr, st2 = nullptr;
        cudaEvent_t ev_front = nullptr, ev_lf = nullptr, ev_done = nullptr;
        uint32_t slot0 = 0, n = 0;
        bool busy = false;
        uint64_t *h_res = nullptr;   // page-locked [2]: bytes gathered, error bits | overflow << 31
        cudaEvent_t trace_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // HYDRIUM_B200_JOBTRACE
        bool traced = false;
        uint32_t *d_ovf = nullptr;
    };
    static constexpr int kJobs = 16;
    Job jobs[kJobs];
    uint64_t *h_job_res = nullptr;   // page-locked [kJobs][2]
    uint32_t *d_job_ovf = nullptr;   // [kJobs]
    TileDesc *h_job_tiles = nullptr; // page-locked [max_batch]: descriptors of asynchronous jobs, by slot
    // Jobs of classic tiles that recur with the same geometry (same job, slots, staging and output
    // addresses: a chunk of the nine-symbol API's ring) are replayed as CUDA graphs: one cudaGraphLaunch
    // instead of ~19 stream calls (56 -> ~15 us of host time per chunk).  A key is captured the second
    // time it is seen, so one-off encodes never pay for a capture.
    struct JobGraphKey {
        int job;
        uint32_t slot0, slots, chain_mode;
        const void *h_src;
        void *d_dst;
        size_t h2d_bytes;
        uint8_t *out;
        uint64_t out_cap;
        bool allow_compact;
        bool operator==(const JobGraphKey &o) const {
            return job == o.job && slot0 == o.slot0 && slots == o.slots && chain_mode == o.chain_mode && h_src == o.h_src &&
                   d_dst == o.d_dst && h2d_bytes == o.h2d_bytes && out == o.out && out_cap == o.out_cap &&
                   allow_compact == o.allow_compact;
        }
    };
    struct JobGraph {
        JobGraphKey key;
        cudaGraphExec_t exec = nullptr;
        bool failed = false;     // capture or instantiation failed once: submit this key directly from now on
        uint64_t last_use = 0;
        uint32_t kernels = 0;    // kernel launches one replay stands for
    };
    std::vector<JobGraph> job_graphs;
    uint64_t job_graph_clock = 0;
    uint64_t graph_launches = 0;
};

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            eng->error = std::string(#call) + ": " + cudaGetErrorString(e_);            \
            return HYD_INTERNAL_ERROR;                                                  \
        }                                                                               \
    } while (0)

// captured job graphs hold the workspace's pointers and launch choices as they were: drop them when those change
static void drop_job_graphs(HydbEngine *eng) {
    if (eng->job_graphs.empty())
        return;
    for (HydbEngine::Job &jb : eng->jobs)   // replays in flight finish first
        if (jb.st) cudaStreamSynchronize(jb.st);
    for (HydbEngine::JobGraph &g : eng->job_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    eng->job_graphs.clear();
}

template <typename T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }

static size_t sample_item_bytes(int sample_fmt) {
    return sample_fmt == HYD_UINT8 ? 1 : (sample_fmt == HYD_UINT16 ? 2 : 4);
}

static const char *tile_error_text(uint32_t bits) {
    if (bits & kErrNonFinite) return "Invalid NaN Float";                 // reference: format.c:124
    if (bits & kErrNegative) return "float samples with a negative opsin mix cannot be encoded (the reference's cube root is undefined there)";
    if (bits & kErrRange) return "HF coefficient exceeds the 16-bit range of the B200 encoder (float samples far outside [0, 1])";
    if (bits & kErrAlphabet) return "HF token alphabet exceeds 64 symbols";
    if (bits & kErrHuffman) return "couldn't find target";               // reference: entropy.c:635
    if (bits & kErrAlias) return "empty underfull during alias table gen";   // reference: entropy.c:219
    if (bits & kErrAnsGap) return "rANS renormalisation gap exceeds 65535 symbols";
    if (bits & kErrLfCapacity) return "LF stream scratch exhausted";
    if (bits & kErrLfAlphabet) return "LF token alphabet exceeds the sparse coder";
    if (bits & kErrSlab) return "encoded tile exceeds the output slab";
    return "unknown device-side error";
}

extern "C" {

HYDStatusCode hydb_engine_create(HydbEngine **out, int device, uint32_t max_batch_tiles) {
    if (!out || !max_batch_tiles)
        return HYD_API_ERROR;
    *out = nullptr;
    HydbEngine *eng = new (std::nothrow) HydbEngine();
    if (!eng)
        return HYD_NOMEM;
    auto fail = [&](HYDStatusCode rc) {
        fprintf(stderr, "hydrium_b200: engine creation failed: %s\n", eng->error.c_str());
        hydb_engine_destroy(eng);
        return rc;
    };
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        eng->error = "no CUDA device available (the B200 encoder has no CPU path)";
        return fail(HYD_INTERNAL_ERROR);
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess)
            device = 0;
    }
    eng->device = device;
    eng->max_batch = max_batch_tiles;
    const size_t T = max_batch_tiles;
    cudaError_t e = cudaSetDevice(device);
    Workspace &w = eng->ws;
    w.capacity = max_batch_tiles;
#define A(call) if (e == cudaSuccess) e = (call)
    A(cudaStreamCreateWithFlags(&eng->st, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&eng->st2, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&eng->ev_front, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&eng->ev_lf, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&eng->ev_desc, cudaEventDisableTiming));
    // earlier bands get higher stream priority: when an SM frees a slot, the tokeniser / chain CTAs of
    // band 0 go before the front-end CTAs of later bands, so the first (3 ms long) chains start early
    int prio_least = 0, prio_greatest = 0;
    A(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    A(cudaStreamCreateWithPriority(&eng->copy_st, cudaStreamNonBlocking, prio_greatest));
    for (int b = 0; b < HydbEngine::kBands; b++) {
        const int prio = prio_greatest + b < prio_least ? prio_greatest + b : prio_least;
        A(cudaStreamCreateWithPriority(&eng->band_st[b], cudaStreamNonBlocking, prio));
        A(cudaStreamCreateWithPriority(&eng->band_st2[b], cudaStreamNonBlocking, prio_least));
        A(cudaEventCreateWithFlags(&eng->band_front[b], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&eng->band_lf[b], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&eng->band_done[b], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&eng->band_h2d[b], cudaEventDisableTiming));
    }
    A(dalloc(&w.tiles, T));
    A(dalloc(&w.coef, T * kMaxBlocks * 3 * 64));
    A(dalloc(&w.nzinfo, T * kMaxBlocks * 3));
    A(dalloc(&w.lfq, T * 3 * kMaxBlocks));
    A(dalloc(&w.syms, T * kMaxHfSyms));
    A(dalloc(&w.nsyms, T));
    A(dalloc(&w.resbits, T));
    A(dalloc(&w.hist, T * kHfClusters * kHfTokens));
    A(dalloc(&w.lfbits, T * kLfBitsWords));
    A(dalloc(&w.lfbitlen, T));
    A(dalloc(&w.dbits, T * kDBitsWords));
    A(dalloc(&w.chain_out, T * 4));
    A(dalloc(&w.flags, T * (kMaxHfSyms / 32)));
    A(dalloc(&w.fwords, T * kMaxHfSyms));
    A(dalloc(&w.slab, T * kSlabBytes));
    A(dalloc(&w.frame_off, T));
    A(dalloc(&w.frame_len, T));
    A(dalloc(&w.out_off, 2 * T + 2));   // a view starting at slot f uses entries [2 f, 2 f + n]
    A(dalloc(&w.tile_err, T));
    A(dalloc(&w.sm_ticket, 256));
    A(cudaMemsetAsync(w.sm_ticket, 0, 256 * sizeof(uint32_t), eng->st));
    A(dalloc(&w.chain_order, T));
    A(dalloc(&w.sm_load, 1024));
    A(cudaMemsetAsync(w.sm_load, 0, 1024 * sizeof(uint32_t), eng->st));
    A(dalloc(&eng->lut8_srgb, 256));
    A(dalloc(&eng->lut8_lin, 256));
    A(dalloc(&eng->lut16_srgb, 65536));
    A(dalloc(&eng->lut16_lin, 65536));
    A(dalloc(&eng->bias, 65536));
    A(dalloc(&eng->templ.words, (size_t)(1 + kMaxShapes) * kTemplWords));
    A(dalloc(&eng->templ.bits, 1 + kMaxShapes));
    A(dalloc(&eng->d_shape_dims, 2 * kMaxShapes));
    A(dalloc(&eng->d_overflow, 1));
    A(dalloc(&eng->d_small, 4));
    A(cudaMallocHost((void **)&eng->h_small, 4 * sizeof(uint64_t)));
    A(cudaMallocHost((void **)&eng->h_err, (T + 1) * sizeof(uint32_t)));
    A(cudaMallocHost((void **)&eng->h_total, sizeof(uint64_t)));
    A(cudaMallocHost((void **)&eng->h_job_res, HydbEngine::kJobs * 2 * sizeof(uint64_t)));
    A(dalloc(&eng->d_job_ovf, HydbEngine::kJobs));
    A(cudaMallocHost((void **)&eng->h_job_tiles, T * sizeof(TileDesc)));
    for (int j = 0; j < HydbEngine::kJobs && e == cudaSuccess; j++) {
        HydbEngine::Job &jb = eng->jobs[j];
        A(cudaStreamCreateWithFlags(&jb.st, cudaStreamNonBlocking));
        A(cudaStreamCreateWithFlags(&jb.st2, cudaStreamNonBlocking));
        A(cudaEventCreateWithFlags(&jb.ev_front, cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&jb.ev_lf, cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&jb.ev_done, cudaEventDisableTiming));
        jb.h_res = eng->h_job_res + 2 * j;
        jb.d_ovf = eng->d_job_ovf + j;
    }
    A(cudaMemsetAsync(eng->templ.words, 0, (size_t)(1 + kMaxShapes) * kTemplWords * sizeof(uint32_t), eng->st));
    A(cudaMemsetAsync(w.lfbits, 0, T * kLfBitsWords * sizeof(uint32_t), eng->st));
    A(cudaMemsetAsync(w.slab, 0, T * kSlabBytes, eng->st));
#undef A
    if (e != cudaSuccess) {
        eng->error = std::string("CUDA allocation failed: ") + cudaGetErrorString(e);
        return fail(e == cudaErrorMemoryAllocation ? HYD_NOMEM : HYD_INTERNAL_ERROR);
    }
    eng->luts = LutSet{eng->lut8_srgb, eng->lut8_lin, eng->lut16_srgb, eng->lut16_lin, eng->bias};
    launch_build_luts(eng->lut8_srgb, eng->lut8_lin, eng->lut16_srgb, eng->lut16_lin, eng->bias, eng->st);
    launch_build_templates(eng->templ, eng->d_shape_dims, 0, 0, true, eng->st);
    eng->launches += 2;
    e = cudaStreamSynchronize(eng->st);
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess) {
        eng->error = std::string("CUDA initialisation kernels failed: ") + cudaGetErrorString(e);
        return fail(HYD_INTERNAL_ERROR);
    }
    *out = eng;
    return HYD_OK;
}

void hydb_engine_destroy(HydbEngine *eng) {
    if (!eng)
        return;
    cudaSetDevice(eng->device);
    if (eng->st) cudaStreamSynchronize(eng->st);
    if (eng->st2) cudaStreamSynchronize(eng->st2);
    Workspace &w = eng->ws;
    void *dev[] = {w.tiles, w.coef, w.nzinfo, w.lfq, w.syms, w.nsyms, w.resbits, w.hist, w.lfbits, w.lfbitlen, w.flags, w.dbits, w.chain_out,
                   w.fwords, w.slab, w.frame_off, w.frame_len, w.out_off, w.tile_err, w.sm_ticket, w.sm_load, w.chain_order, w.dbg_xyb, w.dbg_dct,
                   w.dbg_freqs, w.dbg_sect, w.dbg_clk, eng->lut8_srgb, eng->lut8_lin, eng->lut16_srgb, eng->lut16_lin, eng->bias,
                   eng->templ.words, eng->templ.bits, eng->d_shape_dims, eng->d_overflow};
    for (void *p : dev)
        if (p) cudaFree(p);
    for (int b = 0; b < HydbEngine::kBands; b++) {
        if (b == 0 && eng->copy_st) cudaStreamDestroy(eng->copy_st);
        if (eng->band_st[b]) cudaStreamDestroy(eng->band_st[b]);
        if (eng->band_st2[b]) cudaStreamDestroy(eng->band_st2[b]);
        if (eng->band_front[b]) cudaEventDestroy(eng->band_front[b]);
        if (eng->band_lf[b]) cudaEventDestroy(eng->band_lf[b]);
        if (eng->band_done[b]) cudaEventDestroy(eng->band_done[b]);
        if (eng->band_h2d[b]) cudaEventDestroy(eng->band_h2d[b]);
    }
    drop_job_graphs(eng);
    for (HydbEngine::Job &jb : eng->jobs) {
        if (jb.st) { cudaStreamSynchronize(jb.st); cudaStreamDestroy(jb.st); }
        if (jb.st2) { cudaStreamSynchronize(jb.st2); cudaStreamDestroy(jb.st2); }
        if (jb.ev_front) cudaEventDestroy(jb.ev_front);
        if (jb.ev_lf) cudaEventDestroy(jb.ev_lf);
        if (jb.ev_done) cudaEventDestroy(jb.ev_done);
        for (cudaEvent_t e : jb.trace_ev)
            if (e) cudaEventDestroy(e);
    }
    if (eng->h_job_res) cudaFreeHost(eng->h_job_res);
    if (eng->d_job_ovf) cudaFree(eng->d_job_ovf);
    if (eng->h_job_tiles) cudaFreeHost(eng->h_job_tiles);
    if (eng->ev_desc) cudaEventDestroy(eng->ev_desc);
    if (eng->of_dev) cudaFree(eng->of_dev);
    if (eng->of_host) cudaFreeHost(eng->of_host);
    if (eng->of_ctx) cudaFree(eng->of_ctx);
    if (eng->of_perm) cudaFree(eng->of_perm);
    if (eng->host_in) cudaFree(eng->host_in);
    if (eng->host_out) cudaFree(eng->host_out);
    for (cudaEvent_t ev : eng->tev) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : eng->lev) if (ev) cudaEventDestroy(ev);
    if (eng->h_err) cudaFreeHost(eng->h_err);
    if (eng->h_small) cudaFreeHost(eng->h_small);
    cudaFree(eng->d_small);
    if (eng->h_total) cudaFreeHost(eng->h_total);
    if (eng->ev_front) cudaEventDestroy(eng->ev_front);
    if (eng->ev_lf) cudaEventDestroy(eng->ev_lf);
    if (eng->st) cudaStreamDestroy(eng->st);
    if (eng->st2) cudaStreamDestroy(eng->st2);
    delete eng;
}

const char *hydb_engine_error(const HydbEngine *eng) { return eng ? eng->error.c_str() : "no engine"; }
uint32_t hydb_engine_max_batch(const HydbEngine *eng) { return eng ? eng->max_batch : 0; }
uint64_t hydb_engine_stream(const HydbEngine *eng) { return eng ? (uint64_t)(uintptr_t)eng->st : 0; }
uint64_t hydb_engine_launch_count(const HydbEngine *eng) { return eng ? eng->launches : 0; }
uint64_t hydb_engine_graph_launch_count(const HydbEngine *eng) { return eng ? eng->graph_launches : 0; }

HYDStatusCode hydb_engine_set_chain_kernel(HydbEngine *eng, int mode) {
    if (!eng || mode < 0 || mode > 2)
        return HYD_API_ERROR;
    if (eng->ws.chain_mode != (uint32_t)mode)
        drop_job_graphs(eng);
    eng->ws.chain_mode = (uint32_t)mode;
    return HYD_OK;
}

HYDStatusCode hydb_engine_enable_taps(HydbEngine *eng, int enable) {
    if (!eng)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    Workspace &w = eng->ws;
    const size_t T = eng->max_batch;
    if (!!enable != eng->taps)
        drop_job_graphs(eng);
    if (enable && !eng->taps) {
        CK(dalloc(&w.dbg_xyb, T * 65536 * 3));
        CK(dalloc(&w.dbg_dct, T * 65536 * 3));
        CK(dalloc(&w.dbg_freqs, T * kHfClusters * kHfTokens));
        CK(dalloc(&w.dbg_sect, T * 4));
        CK(dalloc(&w.dbg_clk, T * 4));
        CK(cudaMemset(w.dbg_xyb, 0, T * 65536 * 3 * sizeof(float)));
        CK(cudaMemset(w.dbg_dct, 0, T * 65536 * 3 * sizeof(float)));
        eng->taps = true;
    } else if (!enable && eng->taps) {
        CK(cudaStreamSynchronize(eng->st));
        cudaFree(w.dbg_xyb); cudaFree(w.dbg_dct); cudaFree(w.dbg_freqs); cudaFree(w.dbg_sect); cudaFree(w.dbg_clk);
        w.dbg_xyb = w.dbg_dct = nullptr;
        w.dbg_freqs = w.dbg_sect = w.dbg_clk = nullptr;
        eng->taps = false;
    }
    return HYD_OK;
}

// shape id of a (vbw, vbh) tile; builds its section-B template on first use
static int shape_of(HydbEngine *eng, uint32_t vbw, uint32_t vbh, std::vector<uint32_t> &fresh) {
    const uint32_t key = (vbw << 16) | vbh;
    for (size_t i = 0; i < eng->shapes.size(); i++)
        if (eng->shapes[i] == key)
            return (int)i;
    if (eng->shapes.size() >= (size_t)kMaxShapes)
        return -1;
    eng->shapes.push_back(key);
    fresh.push_back(key);
    return (int)eng->shapes.size() - 1;
}

// workspace seen by a sub-range of the batch starting at tile `first` (all arrays are [tile][...])
static Workspace ws_view(const Workspace &w, uint32_t first) {
    Workspace v = w;
    const size_t f = first;
    v.tiles += f;
    v.coef += f * kMaxBlocks * 3 * 64;
    v.nzinfo += f * kMaxBlocks * 3;
    v.lfq += f * 3 * kMaxBlocks;
    v.syms += f * kMaxHfSyms;
    v.nsyms += f;
    v.resbits += f;
    v.hist += f * kHfClusters * kHfTokens;
    v.lfbits += f * kLfBitsWords;
    v.lfbitlen += f;
    v.dbits += f * kDBitsWords;
    v.chain_out += f * 4;
    v.flags += f * (kMaxHfSyms / 32);
    v.fwords += f * kMaxHfSyms;
    v.slab += f * kSlabBytes;
    v.frame_off += f;
    v.frame_len += f;
    v.out_off += 2 * f;
    v.tile_err += f;
    v.chain_order += f;
    if (v.dbg_xyb) v.dbg_xyb += f * 65536 * 3;
    if (v.dbg_dct) v.dbg_dct += f * 65536 * 3;
    if (v.dbg_freqs) v.dbg_freqs += f * kHfClusters * kHfTokens;
    if (v.dbg_sect) v.dbg_sect += f * 4;
    if (v.dbg_clk) v.dbg_clk += f * 4;
    return v;
}

// validate + translate the caller's tiles, build templates for new shapes, upload the descriptors
// per-slot additions for multi-group frames (see TileDesc in common.cuh)
struct SlotExtra {
    uint32_t flags;
    uint32_t frame_groups, frame_gx, group_index, frame_w, frame_h, frame_x0, frame_y0, preset_info;
};

enum { kPrepHost = 1, kPrepUpload = 2 };   // prepare_tiles phases: descriptors + new shapes / the asynchronous uploads
static HYDStatusCode prepare_tiles(HydbEngine *eng, const HydbTile *tiles, uint32_t n, cudaStream_t st,
                                   const SlotExtra *extra = nullptr, uint32_t slot0 = 0, uint32_t *d_ovf = nullptr,
                                   bool *any_float = nullptr, bool job = false, int phases = kPrepHost | kPrepUpload) {
    if (!(phases & kPrepHost)) {
        TileDesc *h = job ? eng->h_job_tiles + slot0 : eng->h_tiles.data();
        CK(cudaMemcpyAsync(eng->ws.tiles + slot0, h, n * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(eng->ws.tile_err + slot0, 0, n * sizeof(uint32_t), st));
        CK(cudaMemsetAsync(d_ovf ? d_ovf : eng->d_overflow, 0, sizeof(uint32_t), st));
        return HYD_OK;
    }
    std::vector<uint32_t> fresh;
    if (any_float)
        *any_float = false;
    const uint32_t first_fresh = (uint32_t)eng->shapes.size();
    // jobs keep their descriptors in page-locked memory (by slot) so that the upload is asynchronous
    if (!job)
        eng->h_tiles.resize(n);
    TileDesc *h_tiles = job ? eng->h_job_tiles + slot0 : eng->h_tiles.data();
    for (uint32_t i = 0; i < n; i++) {
        const HydbTile &s = tiles[i];
        if (!s.width || !s.height || s.width > 256 || s.height > 256 || (s.x0 & 255) || (s.y0 & 255) ||
            (s.sample_fmt != HYD_UINT8 && s.sample_fmt != HYD_UINT16 && s.sample_fmt != HYD_FLOAT32) || !s.plane[0] || !s.plane[1] || !s.plane[2]) {
            eng->error = "invalid tile descriptor";
            return HYD_API_ERROR;
        }
        if (any_float && s.sample_fmt == HYD_FLOAT32)
            *any_float = true;
        const int shape = shape_of(eng, (s.width + 7) >> 3, (s.height + 7) >> 3, fresh);
        if (shape < 0) {
            eng->error = "too many distinct tile shapes in one engine";
            return HYD_API_ERROR;
        }
        TileDesc &d = h_tiles[i];
        d.plane[0] = s.plane[0];
        d.plane[1] = s.plane[1];
        d.plane[2] = s.plane[2];
        d.row_stride = s.row_stride;
        d.pixel_stride = s.pixel_stride;
        d.w = s.width;
        d.h = s.height;
        d.x0 = s.x0;
        d.y0 = s.y0;
        d.shape = (uint32_t)shape;
        d.image_w = s.image_width;
        d.image_h = s.image_height;
        d.flags = (s.is_last ? kTileLast : 0u) | (s.with_image_header ? kTileFirst : 0u) |
                  ((s.image_width > s.width || s.image_height > s.height) ? kTileCrop : 0u) |
                  (s.sample_fmt == HYD_UINT16 ? kTileFmt16 : 0u) | (s.sample_fmt == HYD_FLOAT32 ? kTileFmtF32 : 0u) |
                  (s.linear_light ? kTileLinear : 0u);
        d.frame_groups = d.frame_gx = d.group_index = d.frame_w = d.frame_h = d.frame_x0 = d.frame_y0 = d.preset_info = 0;
        if (extra) {
            const SlotExtra &e = extra[i];
            d.flags |= e.flags;
            d.frame_groups = e.frame_groups;
            d.frame_gx = e.frame_gx;
            d.group_index = e.group_index;
            d.frame_w = e.frame_w;
            d.frame_h = e.frame_h;
            d.frame_x0 = e.frame_x0;
            d.frame_y0 = e.frame_y0;
            d.preset_info = e.preset_info;
        }
    }
    if (!fresh.empty()) {
        // Section B of a new tile shape (nb_blocks, the zero-predictor MA tree, the constant HF-metadata image:
        // encoder.c:598-626) depends on nothing but the shape.  It is a strictly sequential bit string of a
        // few hundred bits: one GPU thread needed 0.8-1.4 ms for it (k_build_templates, round 1), the host
        // runs the same source (sections.cuh, compiled for both sides) in microseconds and uploads the words.
        std::vector<uint32_t> words((size_t)fresh.size() * kTemplWords, 0u), bits(fresh.size(), 0u);
        std::vector<uint32_t> syms(kSectionSymCap);
        PrefixWork *work = new (std::nothrow) PrefixWork();
        if (!work) {
            eng->error = "out of memory";
            return HYD_NOMEM;
        }
        for (size_t k = 0; k < fresh.size(); k++) {
            memset(work, 0, sizeof(*work));
            BitSink bw;
            bw.init(words.data() + k * kTemplWords, kTemplWords);
            build_section_b(*work, syms.data(), bw, fresh[k] >> 16, fresh[k] & 0xFFFF);
            bw.flush_partial();
            bits[k] = (bw.overflow || work->error) ? 0xFFFFFFFFu : bw.bitlen();
        }
        delete work;
        CK(cudaMemcpyAsync(eng->templ.words + (size_t)(1 + first_fresh) * kTemplWords, words.data(), words.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(eng->templ.bits + 1 + first_fresh, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // other streams (jobs, bands) may use the shape from now on
    }
    if (!(phases & kPrepUpload))
        return HYD_OK;
    // pageable source: the runtime stages it before returning, so h_tiles may be reused immediately
    CK(cudaMemcpyAsync(eng->ws.tiles + slot0, h_tiles, n * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(eng->ws.tile_err + slot0, 0, n * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(d_ovf ? d_ovf : eng->d_overflow, 0, sizeof(uint32_t), st));
    return HYD_OK;
}

// queue the result read-back of a batch of n tiles whose frames were gathered at d_out_pos
static HYDStatusCode queue_readback(HydbEngine *eng, uint32_t n, uint64_t d_out_pos) {
    cudaStream_t st = eng->st;
    CK(cudaMemcpyAsync(eng->h_err, eng->ws.tile_err, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(eng->h_err + eng->max_batch, eng->d_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(eng->h_total, eng->ws.out_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    eng->last_n = n;
    eng->last_base = d_out_pos;
    return HYD_OK;
}

// the kernel sequence of a batch of classic (one-group) tiles whose descriptors are in place:
// `v` = workspace view of the batch, st / st2 = its stream pair
static HYDStatusCode enqueue_tile_kernels(HydbEngine *eng, const Workspace &v, uint32_t n, cudaStream_t st, cudaStream_t st2,
                                          cudaEvent_t ev_front, cudaEvent_t ev_lf, bool allow_compact, uint8_t *out,
                                          uint64_t out_cap, uint64_t out_pos, uint32_t *d_ovf, bool tm) {
    NvtxRange range("hydb:tile_batch");
    if (tm) CK(cudaEventRecord(eng->tev[0], st));
    { NvtxRange r("hydb:xyb_dct_quant"); launch_xyb_dct_quant(v, eng->luts, n, st); }
    if (tm) CK(cudaEventRecord(eng->tev[1], st));
    CK(cudaEventRecord(ev_front, st));
    CK(cudaStreamWaitEvent(st2, ev_front, 0));
    if (tm) CK(cudaEventRecord(eng->lev[0], st2));
    { NvtxRange r("hydb:lf_group"); launch_lf_group(v, n, st2); }
    if (tm) CK(cudaEventRecord(eng->lev[1], st2));
    CK(cudaEventRecord(ev_lf, st2));
    { NvtxRange r("hydb:hf_tokens"); launch_hf_tokens(v, n, st); }
    if (tm) CK(cudaEventRecord(eng->tev[2], st));
    { NvtxRange r("hydb:ans_chain"); launch_ans_chain(v, n, st, allow_compact, 0, true); }
    if (tm) CK(cudaEventRecord(eng->tev[3], st));
    CK(cudaStreamWaitEvent(st, ev_lf, 0));   // the LF stream is first needed by the packer
    { NvtxRange r("hydb:ans_pack"); launch_ans_pack(v, eng->templ, n, st); }
    if (tm) CK(cudaEventRecord(eng->tev[4], st));
    { NvtxRange r("hydb:gather"); launch_gather(v, n, out, out_cap, out_pos, d_ovf, st); }
    if (tm) CK(cudaEventRecord(eng->tev[5], st));
    eng->launches += 7;
    return HYD_OK;
}

HYDStatusCode hydb_engine_encode_tiles(HydbEngine *eng, const HydbTile *tiles, uint32_t n, uint8_t *d_out,
                                       uint64_t d_out_cap, uint64_t d_out_pos) {
    if (!eng || !tiles || !n || n > eng->max_batch || !d_out) {
        if (eng) eng->error = "invalid arguments to hydb_engine_encode_tiles";
        return HYD_API_ERROR;
    }
    CK(cudaSetDevice(eng->device));
    cudaStream_t st = eng->st;
    bool any_float = false;
    {
        const HYDStatusCode rc = prepare_tiles(eng, tiles, n, st, nullptr, 0, nullptr, &any_float);
        if (rc != HYD_OK)
            return rc;
    }
    const bool tm = eng->timing;
    const HYDStatusCode rc = enqueue_tile_kernels(eng, eng->ws, n, st, eng->st2, eng->ev_front, eng->ev_lf, !any_float, d_out,
                                                  d_out_cap, d_out_pos, eng->d_overflow, tm);
    if (rc != HYD_OK)
        return rc;
    eng->timed_pending = tm;
    return queue_readback(eng, n, d_out_pos);
}

// Frames of up to 8 x 8 groups.  A frame of one group is an ordinary tile; a larger one becomes a
// prefix pseudo-tile plus its groups (k_frame.cu).
// frames -> workspace slots (a frame of several groups = one prefix pseudo-tile + its groups)
static HYDStatusCode expand_frames(HydbEngine *eng, const HydbFrame *frames, uint32_t n, std::vector<HydbTile> &tiles,
                                   std::vector<SlotExtra> &extra, bool &any_multi) {
    any_multi = false;
    for (uint32_t f = 0; f < n; f++) {
        const HydbFrame &fr = frames[f];
        if (!fr.width || !fr.height || fr.width > 2048 || fr.height > 2048 || (fr.x0 & 255) || (fr.y0 & 255) ||
            (fr.sample_fmt != HYD_UINT8 && fr.sample_fmt != HYD_UINT16 && fr.sample_fmt != HYD_FLOAT32) ||
            !fr.plane[0] || !fr.plane[1] || !fr.plane[2]) {
            eng->error = "invalid frame descriptor";
            return HYD_API_ERROR;
        }
        const uint32_t gx = (fr.width + 255) >> 8, gy = (fr.height + 255) >> 8, G = gx * gy;
        const size_t item = sample_item_bytes(fr.sample_fmt);
        HydbTile t;
        memset(&t, 0, sizeof(t));
        t.row_stride = fr.row_stride;
        t.pixel_stride = fr.pixel_stride;
        t.image_width = fr.image_width;
        t.image_height = fr.image_height;
        t.is_last = fr.is_last;
        t.sample_fmt = fr.sample_fmt;
        t.linear_light = fr.linear_light;
        SlotExtra e;
        memset(&e, 0, sizeof(e));
        if (G == 1 && !fr.lf_part) {   // single section: the classic path (encoder.c:992-1004)
            for (int k = 0; k < 3; k++) t.plane[k] = fr.plane[k];
            t.width = fr.width;
            t.height = fr.height;
            t.x0 = fr.x0;
            t.y0 = fr.y0;
            t.with_image_header = fr.with_image_header;
            if (fr.one_frame) {
                // one-frame mode, image inside one group: no crop, last (what hyd_api.c did so far)
                t.is_last = 1;
            }
            tiles.push_back(t);
            extra.push_back(e);
            continue;
        }
        any_multi = true;
        e.frame_groups = G;
        e.frame_gx = gx;
        e.frame_w = fr.width;
        e.frame_h = fr.height;
        e.frame_x0 = fr.x0;
        e.frame_y0 = fr.y0;
        if (fr.lf_part)
            e.preset_info = (fr.preset & 0xFFu) | ((fr.preset_bits & 0xFu) << 8) | ((fr.alpha_floor & 0xFFu) << 12) |
                            ((fr.clusters_per_preset & 0xFu) << 20);
        // prefix pseudo-tile
        HydbTile p = t;
        for (int k = 0; k < 3; k++) p.plane[k] = fr.plane[k];
        p.width = p.height = 8;
        p.x0 = p.y0 = 0;
        p.with_image_header = fr.with_image_header;
        SlotExtra pe = e;
        pe.flags = kTilePrefix | (fr.one_frame ? kTileOneFrame : 0u) | (fr.lf_part ? kTileLfPart : 0u);
        pe.group_index = 0xFFFFFFFFu;
        tiles.push_back(p);
        extra.push_back(pe);
        for (uint32_t g = 0; g < G; g++) {
            const uint32_t cx = g % gx, cy = g / gx;
            HydbTile q = t;
            for (int k = 0; k < 3; k++)
                q.plane[k] = (const uint8_t *)fr.plane[k] +
                             ((int64_t)cy * 256 * fr.row_stride + (int64_t)cx * 256 * fr.pixel_stride) * (int64_t)item;
            q.width = fr.width - cx * 256 < 256 ? fr.width - cx * 256 : 256;
            q.height = fr.height - cy * 256 < 256 ? fr.height - cy * 256 : 256;
            q.x0 = fr.x0 + cx * 256;
            q.y0 = fr.y0 + cy * 256;
            q.with_image_header = 0;
            SlotExtra ge = e;
            ge.flags = kTileMulti;
            ge.group_index = g;
            tiles.push_back(q);
            extra.push_back(ge);
        }
    }
    return HYD_OK;
}

// the kernel sequence of a batch that holds multi-group frames (k_frame.cu), descriptors in place
static HYDStatusCode enqueue_frame_kernels(HydbEngine *eng, const Workspace &v, uint32_t slots, cudaStream_t st,
                                           cudaStream_t st2, cudaEvent_t ev_front, cudaEvent_t ev_lf, bool allow_compact,
                                           uint8_t *out, uint64_t out_cap, uint64_t out_pos, uint32_t *d_ovf) {
    NvtxRange range("hydb:frame_batch");
    launch_xyb_dct_quant(v, eng->luts, slots, st);
    CK(cudaEventRecord(ev_front, st));
    CK(cudaStreamWaitEvent(st2, ev_front, 0));
    launch_lf_group(v, slots, st2);   // classic single-group frames in the same batch
    launch_frame_lf(v, slots, st2);
    CK(cudaEventRecord(ev_lf, st2));
    launch_hf_tokens(v, slots, st);
    launch_frame_hist_sum(v, slots, st);
    launch_ans_chain(v, slots, st, allow_compact);
    CK(cudaStreamWaitEvent(st, ev_lf, 0));
    launch_ans_pack(v, eng->templ, slots, st);
    launch_frame_finish(v, slots, st);
    launch_gather(v, slots, out, out_cap, out_pos, d_ovf, st);
    eng->launches += 10;
    return HYD_OK;
}

HYDStatusCode hydb_engine_encode_frames(HydbEngine *eng, const HydbFrame *frames, uint32_t n, uint8_t *d_out,
                                        uint64_t d_out_cap, uint64_t d_out_pos) {
    if (!eng || !frames || !n || !d_out) {
        if (eng) eng->error = "invalid arguments to hydb_engine_encode_frames";
        return HYD_API_ERROR;
    }
    std::vector<HydbTile> tiles;
    std::vector<SlotExtra> extra;
    bool any_multi = false;
    {
        const HYDStatusCode rc = expand_frames(eng, frames, n, tiles, extra, any_multi);
        if (rc != HYD_OK)
            return rc;
    }
    const uint32_t slots = (uint32_t)tiles.size();
    if (slots > eng->max_batch) {
        eng->error = "frames need more workspace slots than the engine's batch size";
        return HYD_API_ERROR;
    }
    if (!any_multi)
        return hydb_engine_encode_tiles(eng, tiles.data(), slots, d_out, d_out_cap, d_out_pos);
    CK(cudaSetDevice(eng->device));
    cudaStream_t st = eng->st;
    bool any_float = false;
    {
        const HYDStatusCode rc = prepare_tiles(eng, tiles.data(), slots, st, extra.data(), 0, nullptr, &any_float);
        if (rc != HYD_OK)
            return rc;
    }
    const HYDStatusCode rc = enqueue_frame_kernels(eng, eng->ws, slots, st, eng->st2, eng->ev_front, eng->ev_lf, !any_float,
                                                   d_out, d_out_cap, d_out_pos, eng->d_overflow);
    if (rc != HYD_OK)
        return rc;
    eng->timed_pending = false;
    return queue_readback(eng, slots, d_out_pos);
}

// ---- asynchronous jobs ---------------------------------------------------------------------------------
// A job = n frames (or tiles) on workspace slots [slot0, slot0 + slots) with its own stream pair:
// [h2d copy of the staged pixels] -> kernels -> gather straight into `out`, which may be page-locked
// HOST memory (the compaction kernel then writes the codestream across PCIe itself and nothing is
// copied back afterwards) -> k_job_result writes {bytes, error bits} into the job's page-locked record.
// Several jobs run concurrently; the caller owns the slot ranges and polls jobs in any order.
HYDStatusCode hydb_engine_submit_frames(HydbEngine *eng, const HydbFrame *frames, uint32_t n, uint32_t slot0,
                                        const void *h_src, void *d_dst, size_t h2d_bytes, uint8_t *out, uint64_t out_cap,
                                        uint32_t *job_out, uint32_t *slots_out) {
    if (!eng || !frames || !n || !out || !job_out) {
        if (eng) eng->error = "invalid arguments to hydb_engine_submit_frames";
        return HYD_API_ERROR;
    }
    int j = -1;
    for (int k = 0; k < HydbEngine::kJobs; k++)
        if (!eng->jobs[k].busy) {
            j = k;
            break;
        }
    if (j < 0) {
        eng->error = "no free job: poll and release a finished one first";
        return HYD_API_ERROR;
    }
    NvtxRange range("hydb:submit_job");
    std::vector<HydbTile> tiles;
    std::vector<SlotExtra> extra;
    bool any_multi = false;
    {
        const HYDStatusCode rc = expand_frames(eng, frames, n, tiles, extra, any_multi);
        if (rc != HYD_OK)
            return rc;
    }
    const uint32_t slots = (uint32_t)tiles.size();
    if ((uint64_t)slot0 + slots > eng->max_batch) {
        eng->error = "job exceeds the engine's workspace slots";
        return HYD_API_ERROR;
    }
    CK(cudaSetDevice(eng->device));
    HydbEngine::Job &jb = eng->jobs[j];
    bool any_float = false;
    {   // descriptors into the job's page-locked slots; sections of new tile shapes (synchronous, rare)
        const HYDStatusCode rc = prepare_tiles(eng, tiles.data(), slots, jb.st, extra.data(), slot0, jb.d_ovf, &any_float, true, kPrepHost);
        if (rc != HYD_OK)
            return rc;
    }
    const Workspace v = ws_view(eng->ws, slot0);
    // everything the job enqueues on its stream pair, in order
    auto enqueue = [&]() -> HYDStatusCode {
        if (h_src && h2d_bytes)
            CK(cudaMemcpyAsync(d_dst, h_src, h2d_bytes, cudaMemcpyHostToDevice, jb.st));
        HYDStatusCode rc = prepare_tiles(eng, tiles.data(), slots, jb.st, extra.data(), slot0, jb.d_ovf, nullptr, true, kPrepUpload);
        if (rc != HYD_OK)
            return rc;
        rc = any_multi ? enqueue_frame_kernels(eng, v, slots, jb.st, jb.st2, jb.ev_front, jb.ev_lf, !any_float, out, out_cap, 0, jb.d_ovf)
                       : enqueue_tile_kernels(eng, v, slots, jb.st, jb.st2, jb.ev_front, jb.ev_lf, !any_float, out, out_cap, 0, jb.d_ovf,
                                              false);
        if (rc != HYD_OK)
            return rc;
        launch_job_result(v.tile_err, slots, v.out_off + slots, jb.d_ovf, jb.h_res, jb.st);
        eng->launches++;
        return HYD_OK;
    };
    // HYDRIUM_B200_JOBTRACE=1: timestamps between the stages of every classic-tile job (no graphs then),
    // printed when the job is polled done -- where a job's latency goes while other jobs' chains fill the GPU
    static const bool job_trace = [] { const char *e = getenv("HYDRIUM_B200_JOBTRACE"); return e && *e && *e != '0'; }();
    static const bool graphs_on = [] { const char *e = getenv("HYDRIUM_B200_GRAPHS"); return !(e && e[0] == '0'); }();
    bool done = false;
    jb.traced = false;
    if (job_trace && !any_multi) {
        for (cudaEvent_t &e : jb.trace_ev)
            if (!e) CK(cudaEventCreate(&e));
        cudaStream_t st = jb.st;
        CK(cudaEventRecord(jb.trace_ev[0], st));
        if (h_src && h2d_bytes)
            CK(cudaMemcpyAsync(d_dst, h_src, h2d_bytes, cudaMemcpyHostToDevice, st));
        HYDStatusCode rc = prepare_tiles(eng, tiles.data(), slots, st, extra.data(), slot0, jb.d_ovf, nullptr, true, kPrepUpload);
        if (rc != HYD_OK)
            return rc;
        CK(cudaEventRecord(jb.trace_ev[1], st));
        launch_xyb_dct_quant(v, eng->luts, slots, st);
        CK(cudaEventRecord(jb.trace_ev[2], st));
        CK(cudaEventRecord(jb.ev_front, st));
        CK(cudaStreamWaitEvent(jb.st2, jb.ev_front, 0));
        launch_lf_group(v, slots, jb.st2);
        CK(cudaEventRecord(jb.ev_lf, jb.st2));
        launch_hf_tokens(v, slots, st);
        CK(cudaEventRecord(jb.trace_ev[3], st));
        launch_ans_chain(v, slots, st, !any_float, 0, true);
        CK(cudaEventRecord(jb.trace_ev[4], st));
        CK(cudaStreamWaitEvent(st, jb.ev_lf, 0));
        CK(cudaEventRecord(jb.trace_ev[5], st));
        launch_ans_pack(v, eng->templ, slots, st);
        CK(cudaEventRecord(jb.trace_ev[6], st));
        launch_gather(v, slots, out, out_cap, 0, jb.d_ovf, st);
        launch_job_result(v.tile_err, slots, v.out_off + slots, jb.d_ovf, jb.h_res, st);
        CK(cudaEventRecord(jb.trace_ev[7], st));
        eng->launches += 8;
        jb.traced = true;
        done = true;
    }
    if (!done && graphs_on && !any_multi) {
        const HydbEngine::JobGraphKey key{j, slot0, slots, eng->ws.chain_mode, h_src, d_dst, h_src ? h2d_bytes : 0, out, out_cap, !any_float};
        HydbEngine::JobGraph *g = nullptr;
        for (HydbEngine::JobGraph &c : eng->job_graphs)
            if (c.key == key) {
                g = &c;
                break;
            }
        if (!g) {   // first sighting: remember the key, submit directly
            if (eng->job_graphs.size() >= 64) {
                size_t lru = 0;
                for (size_t i = 1; i < eng->job_graphs.size(); i++)
                    if (eng->job_graphs[i].last_use < eng->job_graphs[lru].last_use)
                        lru = i;
                if (eng->job_graphs[lru].exec)
                    cudaGraphExecDestroy(eng->job_graphs[lru].exec);
                eng->job_graphs.erase(eng->job_graphs.begin() + (long)lru);
            }
            HydbEngine::JobGraph fresh;
            fresh.key = key;
            fresh.last_use = ++eng->job_graph_clock;
            eng->job_graphs.push_back(fresh);
        } else {
            g->last_use = ++eng->job_graph_clock;
            if (!g->exec && !g->failed) {   // second sighting: record the sequence instead of running it
                const uint64_t launches_before = eng->launches;
                cudaGraph_t graph = nullptr;
                bool ok = cudaStreamBeginCapture(jb.st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    const HYDStatusCode rc = enqueue();
                    const cudaError_t e = cudaStreamEndCapture(jb.st, &graph);
                    ok = rc == HYD_OK && e == cudaSuccess && graph;
                }
                if (ok)
                    ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
                if (graph)
                    cudaGraphDestroy(graph);
                g->kernels = (uint32_t)(eng->launches - launches_before);
                eng->launches = launches_before;
                if (!ok) {
                    cudaGetLastError();   // a failed capture leaves a (non-sticky) error behind
                    g->exec = nullptr;
                    g->failed = true;
                }
            }
            if (g->exec) {
                CK(cudaGraphLaunch(g->exec, jb.st));
                eng->graph_launches++;
                eng->launches += g->kernels;   // the counter counts kernels, launched one by one or replayed
                done = true;
            }
        }
    }
    if (!done) {
        const HYDStatusCode rc = enqueue();
        if (rc != HYD_OK)
            return rc;
    }
    CK(cudaEventRecord(jb.ev_done, jb.st));
    CK(cudaGetLastError());
    jb.busy = true;
    jb.slot0 = slot0;
    jb.n = slots;
    *job_out = (uint32_t)j;
    if (slots_out)
        *slots_out = slots;
    return HYD_OK;
}

// wait = 0: HYD_DEFAULT while the job is still running.  HYD_OK: *bytes were gathered.
// HYD_NEED_MORE_OUTPUT: the frames did not fit `out` (hydb_engine_job_regather them into a larger buffer).
// Errors of the job's tiles come back as HYD_INTERNAL_ERROR / HYD_API_ERROR with hydb_engine_error set.
HYDStatusCode hydb_engine_job_poll(HydbEngine *eng, uint32_t job, int wait, uint64_t *bytes) {
    if (!eng || job >= (uint32_t)HydbEngine::kJobs || !eng->jobs[job].busy)
        return HYD_API_ERROR;
    HydbEngine::Job &jb = eng->jobs[job];
    if (wait) {
        CK(cudaEventSynchronize(jb.ev_done));
    } else {
        const cudaError_t q = cudaEventQuery(jb.ev_done);
        if (q == cudaErrorNotReady)
            return HYD_DEFAULT;
        CK(q);
    }
    if (jb.traced) {
        jb.traced = false;
        float ms[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 7; i++)
            cudaEventElapsedTime(&ms[i], jb.trace_ev[i], jb.trace_ev[i + 1]);
        fprintf(stderr, "[hydrium_b200] job %u (%u tiles): copies %.3f, xyb_dct_quant %.3f, hf_tokens %.3f, ans_chain %.3f, wait for lf_group %.3f, "
                "ans_pack %.3f, gather + result %.3f ms\n", job, jb.n, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5], ms[6]);
    }
    const uint64_t res = reinterpret_cast<volatile uint64_t *>(jb.h_res)[1];
    if (bytes)
        *bytes = reinterpret_cast<volatile uint64_t *>(jb.h_res)[0];
    const uint32_t err = (uint32_t)res & 0x7FFFFFFFu;
    if (err) {
        eng->error = tile_error_text(err);
        return (err & (kErrNonFinite | kErrNegative)) ? HYD_API_ERROR : HYD_INTERNAL_ERROR;
    }
    return (res >> 31) & 1u ? HYD_NEED_MORE_OUTPUT : HYD_OK;
}

// gather a finished job's frames again, into device memory, synchronously (after HYD_NEED_MORE_OUTPUT)
HYDStatusCode hydb_engine_job_regather(HydbEngine *eng, uint32_t job, uint8_t *d_out, uint64_t d_out_cap, uint64_t *bytes) {
    if (!eng || job >= (uint32_t)HydbEngine::kJobs || !eng->jobs[job].busy || !d_out || !bytes)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    HydbEngine::Job &jb = eng->jobs[job];
    const Workspace v = ws_view(eng->ws, jb.slot0);
    CK(cudaMemsetAsync(jb.d_ovf, 0, sizeof(uint32_t), jb.st));
    launch_gather(v, jb.n, d_out, d_out_cap, 0, jb.d_ovf, jb.st);
    launch_job_result(v.tile_err, jb.n, v.out_off + jb.n, jb.d_ovf, jb.h_res, jb.st);
    eng->launches += 3;
    CK(cudaStreamSynchronize(jb.st));
    const uint64_t res = reinterpret_cast<volatile uint64_t *>(jb.h_res)[1];
    *bytes = reinterpret_cast<volatile uint64_t *>(jb.h_res)[0];
    if ((res >> 31) & 1u) {
        eng->error = "device output buffer too small";
        return HYD_NEED_MORE_OUTPUT;
    }
    return HYD_OK;
}

HYDStatusCode hydb_engine_job_release(HydbEngine *eng, uint32_t job) {
    if (!eng || job >= (uint32_t)HydbEngine::kJobs)
        return HYD_API_ERROR;
    eng->jobs[job].busy = false;
    return HYD_OK;
}

HYDStatusCode hydb_engine_read_model(HydbEngine *eng, uint32_t slot, uint32_t *bits_out, uint32_t *nbits,
                                     uint32_t *max_alphabet) {
    if (!eng || !bits_out || !nbits || !max_alphabet || slot >= eng->max_batch)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    CK(cudaStreamSynchronize(eng->st));   // (jobs: the caller has polled the job that owns the slot)
    uint32_t meta = 0;
    std::vector<uint32_t> d(kDBitsWords);
    CK(cudaMemcpy(&meta, eng->ws.chain_out + (size_t)slot * 4 + 2, sizeof(meta), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d.data(), eng->ws.dbits + (size_t)slot * kDBitsWords, kDBitsWords * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    const uint32_t len = meta & 0xFFFFu, off = (meta >> 16) & 0xFFu;
    *max_alphabet = meta >> 24;
    *nbits = len - off;
    // histograms start `off` bits into section D: shift them down to bit 0
    for (uint32_t i = 0; i < kDBitsWords; i++) {
        const uint32_t b = off + 32u * i, w = b >> 5, r = b & 31u;
        const uint32_t lo = w < (uint32_t)kDBitsWords ? d[w] : 0u, hi = w + 1 < (uint32_t)kDBitsWords ? d[w + 1] : 0u;
        bits_out[i] = r ? ((lo >> r) | (hi << (32u - r))) : lo;
    }
    return HYD_OK;
}

// info words (all uint32):
//   [0] image width  [1] image height  [2] with_image_header  [3] largest token alphabet of the frame
//   [4] n = LF groups of the image (all sent exactly once)     [5] total PassGroups
//   [8 .. 8+n)            lfid of the k-th LF group sent (raster id)
//   [8+n .. 8+2n)         byte length of the k-th sent LF group's LFGroup section
//   [8+2n .. 8+2n+G)      byte length of every PassGroup section, in the order the parts produced them
//   then per preset p = 0..n-1:  nbits[p], followed by ceil(nbits/32) words of histogram bits
HYDStatusCode hydb_oneframe_finish(HydbEngine *eng, const uint32_t *info, uint32_t info_words, uint8_t *head,
                                   uint32_t head_cap, uint32_t *head_len, uint8_t *hf_global, uint32_t hf_cap,
                                   uint32_t *hf_len) {
    if (!eng || !info || info_words < 8 || !head || !head_len || !hf_global || !hf_len)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    // One device arena {info | scratch | out} and one page-locked mirror {info | out}, kept by the engine: this
    // call sits on the tail of every one-frame image (nothing surfaces before it), where three cudaMalloc /
    // cudaFree pairs (each cudaFree synchronises the device) and two pageable copies cost more than the kernel.
    const uint32_t n = info[4];
    const size_t scratch_words = (size_t)2 * 1485 * n + 16384 + (size_t)4 * (8 + n + info[5]);
    const size_t info_bytes = ((size_t)info_words * 4 + 255) & ~(size_t)255, scratch_bytes = (scratch_words * 4 + 255) & ~(size_t)255;
    const size_t out_bytes = (size_t)head_cap + hf_cap + 64;
    if (eng->of_dev_cap < info_bytes + scratch_bytes + out_bytes) {
        if (eng->of_dev) cudaFree(eng->of_dev);
        eng->of_dev = nullptr;
        eng->of_dev_cap = 0;
        const size_t want = (info_bytes + scratch_bytes + out_bytes) * 2;
        if (cudaMalloc(&eng->of_dev, want) != cudaSuccess) {
            eng->error = "device allocation failed";
            return HYD_NOMEM;
        }
        eng->of_dev_cap = want;
    }
    if (eng->of_host_cap < info_bytes + out_bytes) {
        if (eng->of_host) cudaFreeHost(eng->of_host);
        eng->of_host = nullptr;
        eng->of_host_cap = 0;
        const size_t want = (info_bytes + out_bytes) * 2;
        if (cudaMallocHost((void **)&eng->of_host, want) != cudaSuccess) {
            eng->error = "page-locked allocation failed";
            return HYD_NOMEM;
        }
        eng->of_host_cap = want;
    }
    uint8_t *dev = static_cast<uint8_t *>(eng->of_dev);
    uint32_t *d_info = reinterpret_cast<uint32_t *>(dev), *d_scratch = reinterpret_cast<uint32_t *>(dev + info_bytes);
    uint8_t *d_out = dev + info_bytes + scratch_bytes, *h_out = eng->of_host + info_bytes;
    memcpy(eng->of_host, info, (size_t)info_words * 4);
    CK(cudaMemcpyAsync(d_info, eng->of_host, (size_t)info_words * 4, cudaMemcpyHostToDevice, eng->st));
    if (eng->of_ctx_n != n || !eng->of_ctx) {   // another geometry: the cached stream is of no use
        eng->of_ctx_bits = 0;
        eng->of_ctx_n = n;
        const uint32_t want_words = 1485u * n + 1024u;
        if (eng->of_ctx_cap_words < want_words) {
            if (eng->of_ctx) cudaFree(eng->of_ctx);
            eng->of_ctx = nullptr;
            eng->of_ctx_cap_words = 0;
            if (cudaMalloc(&eng->of_ctx, (size_t)want_words * 4) != cudaSuccess) {
                eng->error = "device allocation failed";
                return HYD_NOMEM;
            }
            eng->of_ctx_cap_words = want_words;
        }
    }
    const bool ctx_hit = eng->of_ctx_bits != 0;
    {   // key of the permutation stream: image size, whether the image header precedes, LF groups and their send order
        std::vector<uint32_t> key;
        if (info_words >= 8 + n) {
            key.assign({info[0], info[1], info[4], info[5]});
            key.insert(key.end(), info + 8, info + 8 + n);
        }
        const uint32_t want_words = 2u + info[5] + n + 1024u;
        if (key.empty() || key != eng->of_perm_key || !eng->of_perm) {
            eng->of_perm_bits = 0;
            eng->of_perm_key = key;
            if (eng->of_perm_cap_words < want_words) {
                if (eng->of_perm) cudaFree(eng->of_perm);
                eng->of_perm = nullptr;
                eng->of_perm_cap_words = 0;
                if (cudaMalloc(&eng->of_perm, (size_t)want_words * 4) != cudaSuccess) {
                    eng->error = "device allocation failed";
                    return HYD_NOMEM;
                }
                eng->of_perm_cap_words = want_words;
            }
        }
    }
    const bool perm_hit = eng->of_perm_bits != 0;
    launch_oneframe_finish(d_info, info_words, d_scratch, (uint32_t)scratch_words, d_out, head_cap, hf_cap, eng->of_ctx,
                           eng->of_ctx_cap_words, eng->of_ctx_bits, eng->of_perm, eng->of_perm_cap_words, eng->of_perm_bits, eng->st);
    eng->launches++;
    CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, eng->st));
    CK(cudaStreamSynchronize(eng->st));
    uint32_t res[10];
    memcpy(res, h_out + head_cap + hf_cap, 40);
    if (!ctx_hit && !res[2])
        eng->of_ctx_bits = res[8];   // 0 when the stream did not fit the cache: coded again next time
    if (!perm_hit && !res[2] && !eng->of_perm_key.empty())
        eng->of_perm_bits = res[9];
    {
        static const bool trace = [] { const char *e = getenv("HYDRIUM_B200_APITRACE"); return e && *e && *e != '0'; }();
        if (trace)
            fprintf(stderr, "[hydrium_b200] k_oneframe_finish phases (cycles): TOC permutation %u, context map stream %u, permutation stream %u, "
                    "single-thread tail %u\n", res[4], res[5], res[6], res[7]);
    }
    if (res[2] || res[0] > head_cap || res[1] > hf_cap || res[3] > 52 || res[3] > res[0] || 52 + (res[0] - res[3]) > head_cap) {
        eng->error = "one-frame head does not fit its buffer";
        return HYD_INTERNAL_ERROR;
    }
    *head_len = res[0];
    *hf_len = res[1];
    memcpy(head, h_out, res[3]);                              // container prefix, if any
    memcpy(head + res[3], h_out + 52, res[0] - res[3]);      // image / frame header, TOC, LFGlobal
    memcpy(hf_global, h_out + head_cap, res[1]);
    return HYD_OK;
}

// Image header of an ICC-tagged codestream (k_icc_header): `icc` is the profile as
// hyd_set_suggested_icc_profile mangles it.  Synchronous; dst receives [container prefix] + header bytes.
HYDStatusCode hydb_engine_icc_header(HydbEngine *eng, uint32_t width, uint32_t height, const uint8_t *icc,
                                     uint32_t icc_size, uint8_t *dst, uint64_t cap, uint64_t *len) {
    if (!eng || !icc || !icc_size || !dst || !len || icc_size > (64u << 20))
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    const size_t bits_words = (size_t)icc_size * 21 / 32 + 2048;
    const size_t out_cap = bits_words * 4 + 64;
    uint8_t *d_icc = nullptr, *d_out = nullptr;
    uint32_t *d_bits = nullptr, *d_res = nullptr;
    if (cudaMalloc(&d_icc, icc_size) != cudaSuccess || cudaMalloc(&d_bits, bits_words * 4) != cudaSuccess ||
        cudaMalloc(&d_out, out_cap) != cudaSuccess || cudaMalloc(&d_res, 16) != cudaSuccess) {
        cudaFree(d_icc); cudaFree(d_bits); cudaFree(d_out); cudaFree(d_res);
        eng->error = "device allocation failed";
        return HYD_NOMEM;
    }
    HYDStatusCode rc = HYD_OK;
    uint32_t res[2] = {0, 0};
    if (cudaMemcpyAsync(d_icc, icc, icc_size, cudaMemcpyHostToDevice, eng->st) != cudaSuccess)
        rc = HYD_INTERNAL_ERROR;
    if (rc == HYD_OK) {
        launch_icc_header(d_icc, icc_size, width, height, d_bits, (uint32_t)bits_words, d_out, (uint32_t)out_cap, d_res, eng->st);
        eng->launches++;
        if (cudaMemcpyAsync(res, d_res, 8, cudaMemcpyDeviceToHost, eng->st) != cudaSuccess ||
            cudaStreamSynchronize(eng->st) != cudaSuccess)
            rc = HYD_INTERNAL_ERROR;
    }
    if (rc == HYD_OK && (res[1] || res[0] > cap)) {
        eng->error = res[1] ? "ICC profile could not be entropy coded" : "image header does not fit its buffer";
        rc = HYD_INTERNAL_ERROR;
    }
    if (rc == HYD_OK) {
        *len = res[0];
        if (cudaMemcpy(dst, d_out, res[0], cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = HYD_INTERNAL_ERROR;
    }
    cudaFree(d_icc); cudaFree(d_bits); cudaFree(d_out); cudaFree(d_res);
    if (rc != HYD_OK && eng->error.empty())
        eng->error = "CUDA failure while writing the ICC image header";
    return rc;
}

HYDStatusCode hydb_engine_finish(HydbEngine *eng, uint64_t *batch_bytes) {
    if (!eng)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    CK(cudaStreamSynchronize(eng->st));
    CK(cudaGetLastError());
    if (eng->timed_pending) {
        CK(cudaStreamSynchronize(eng->st2));
        float ms = 0;
        for (int k = 0; k < 5; k++) {
            CK(cudaEventElapsedTime(&ms, eng->tev[k], eng->tev[k + 1]));
            eng->stage_ms[k] += ms;
        }
        CK(cudaEventElapsedTime(&ms, eng->lev[0], eng->lev[1]));
        eng->stage_ms[5] += ms;
        eng->stage_ms[6] += 1;
        eng->timed_pending = false;
    }
    for (uint32_t i = 0; i < eng->last_n; i++) {
        if (eng->h_err[i]) {
            eng->error = tile_error_text(eng->h_err[i]);
            return (eng->h_err[i] & (kErrNonFinite | kErrNegative)) ? HYD_API_ERROR : HYD_INTERNAL_ERROR;
        }
    }
    if (eng->last_n && eng->h_err[eng->max_batch]) {
        eng->error = "device output buffer too small";
        return HYD_NEED_MORE_OUTPUT;
    }
    if (batch_bytes)
        *batch_bytes = eng->last_n ? *eng->h_total : 0;
    return HYD_OK;
}

// byte length of every frame of the last batch, in tile order (for callers that split a batch
// into several codestreams, e.g. one per image)
HYDStatusCode hydb_engine_frame_lengths(HydbEngine *eng, uint32_t *dst, uint32_t n) {
    if (!eng || !dst || n > eng->last_n)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    CK(cudaStreamSynchronize(eng->st));
    CK(cudaMemcpy(dst, eng->ws.frame_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return HYD_OK;
}
// the same for the slots of a finished job
HYDStatusCode hydb_engine_slot_frame_lengths(HydbEngine *eng, uint32_t slot0, uint32_t *dst, uint32_t n) {
    if (!eng || !dst || (uint64_t)slot0 + n > eng->max_batch)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    CK(cudaMemcpy(dst, eng->ws.frame_len + slot0, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return HYD_OK;
}

int64_t hydb_image_header(uint32_t width, uint32_t height, uint8_t *dst, uint64_t cap) {
    uint64_t n = 0;
    if (image_needs_level10(width, height)) {
        if (cap < 49)
            return HYD_API_ERROR;
        n = put_level10_prefix(dst);
    }
    uint32_t words[8] = {0};
    BitSink bw;
    bw.init(words, 8);
    put_image_header(bw, width, height);
    bw.flush_partial();
    const uint32_t bytes = bw.bitlen() >> 3;
    if (bw.overflow || n + bytes > cap)
        return HYD_API_ERROR;
    for (uint32_t i = 0; i < bytes; i++)
        dst[n + i] = (uint8_t)(words[i >> 2] >> (8 * (i & 3)));
    return (int64_t)(n + bytes);
}

// Whole-image tile list (raster order) for pixels at `base` (device memory).
static void image_tiles(std::vector<HydbTile> &tiles, const void *base, uint32_t width, uint32_t height, uint32_t channels,
                        int64_t row_stride, int sample_fmt, int linear_light, uint32_t row_begin, uint32_t row_end,
                        int with_header) {
    const uint32_t tiles_x = (width + 255) >> 8, tiles_y = (height + 255) >> 8;
    const size_t item = sample_item_bytes(sample_fmt);
    tiles.resize((size_t)tiles_x * (row_end - row_begin));
    for (size_t idx = 0; idx < tiles.size(); idx++) {
        const uint32_t tx = (uint32_t)(idx % tiles_x), ty = row_begin + (uint32_t)(idx / tiles_x);
        HydbTile &t = tiles[idx];
        memset(&t, 0, sizeof(t));
        const uint8_t *p = (const uint8_t *)base +
                           ((int64_t)(ty - row_begin) * 256 * row_stride + (int64_t)tx * 256 * channels) * (int64_t)item;
        t.plane[0] = p;
        t.plane[1] = p + item;
        t.plane[2] = p + 2 * item;
        t.row_stride = row_stride;
        t.pixel_stride = channels;
        t.x0 = tx * 256;
        t.y0 = ty * 256;
        t.width = width - t.x0 < 256 ? width - t.x0 : 256;
        t.height = height - t.y0 < 256 ? height - t.y0 : 256;
        t.image_width = width;
        t.image_height = height;
        t.is_last = (tx + 1 == tiles_x && ty + 1 == tiles_y) ? 1 : 0;   // encoder.c:482-485
        t.sample_fmt = sample_fmt;
        t.linear_light = linear_light;
        t.with_image_header = (with_header && idx == 0) ? 1 : 0;
    }
}

// Band pipeline for a batch that holds `rows` tile rows of `tiles_x` tiles (descriptors already
// uploaded): the rows are cut into up to kBands bands, each with its own stream pair, so that the
// rANS chains of early bands start while later bands are still in their front-end kernels (and,
// on the host path, while their pixels are still crossing PCIe: `h_src` != NULL copies each
// band's rows to `d_dst` on the band's stream first).  Frames are position- but not
// order-dependent; the caller gathers all tiles afterwards on eng->st, which waits for every band.
static HYDStatusCode launch_bands(HydbEngine *eng, uint32_t tiles_x, uint32_t rows, const void *h_src, void *d_dst,
                                  size_t row_bytes, uint32_t pixel_rows, bool any_float) {
    NvtxRange range("hydb:band_pipeline");
    // HYDRIUM_B200_BANDTRACE=1: print when each band's stages started / ended (development aid; synchronises)
    static const bool trace = [] { const char *e = getenv("HYDRIUM_B200_BANDTRACE"); return e && e[0] == '1'; }();
    static cudaEvent_t tr[1 + HydbEngine::kBands * 5];
    if (trace && !tr[0])
        for (cudaEvent_t &e : tr) cudaEventCreate(&e);
    if (trace) cudaEventRecord(tr[0], eng->st);
    CK(cudaEventRecord(eng->ev_desc, eng->st));
    // HYDRIUM_B200_BANDS=1..4 for experiments (tools/bands_sweep.py: 3.80 / 3.70 / 3.61 / 3.60 ms for 1 / 2 / 3 / 4 bands
    // on config 2; five and more were slower still)
    static const uint32_t band_limit = [] { const char *e = getenv("HYDRIUM_B200_BANDS"); int v = e ? atoi(e) : 0;
                                            return (uint32_t)(v >= 1 && v <= HydbEngine::kBands ? v : HydbEngine::kBands); }();
    const uint32_t nbands = rows < band_limit ? rows : band_limit;
    if (h_src) {
        // all copies first, on one stream: they reach the device in band order at full PCIe rate, and
        // every band's kernels wait only for their own rows (separate per-band copies were served in
        // an order of the copy engine's choosing, the heavy first bands not first)
        CK(cudaStreamWaitEvent(eng->copy_st, eng->ev_desc, 0));
        for (uint32_t b = 0; b < nbands; b++) {
            const uint32_t r0 = (uint32_t)((uint64_t)rows * b / nbands), r1 = (uint32_t)((uint64_t)rows * (b + 1) / nbands);
            const size_t y0 = (size_t)r0 * 256, y1 = (size_t)r1 * 256 < pixel_rows ? (size_t)r1 * 256 : pixel_rows;
            CK(cudaMemcpyAsync((uint8_t *)d_dst + y0 * row_bytes, (const uint8_t *)h_src + y0 * row_bytes,
                               (y1 - y0) * row_bytes, cudaMemcpyHostToDevice, eng->copy_st));
            CK(cudaEventRecord(eng->band_h2d[b], eng->copy_st));
        }
    }
    for (uint32_t b = 0; b < nbands; b++) {
        const uint32_t r0 = (uint32_t)((uint64_t)rows * b / nbands), r1 = (uint32_t)((uint64_t)rows * (b + 1) / nbands);
        const uint32_t first = r0 * tiles_x, n = (r1 - r0) * tiles_x;
        cudaStream_t sb = eng->band_st[b], sb2 = eng->band_st2[b];
        CK(cudaStreamWaitEvent(sb, eng->ev_desc, 0));
        if (h_src)
            CK(cudaStreamWaitEvent(sb, eng->band_h2d[b], 0));
        const Workspace v = ws_view(eng->ws, first);
        if (trace) cudaEventRecord(tr[1 + b * 5 + 0], sb);
        launch_xyb_dct_quant(v, eng->luts, n, sb);
        if (trace) cudaEventRecord(tr[1 + b * 5 + 1], sb);
        CK(cudaEventRecord(eng->band_front[b], sb));
        CK(cudaStreamWaitEvent(sb2, eng->band_front[b], 0));
        launch_lf_group(v, n, sb2);
        CK(cudaEventRecord(eng->band_lf[b], sb2));
        launch_hf_tokens(v, n, sb);
        if (trace) cudaEventRecord(tr[1 + b * 5 + 2], sb);
        launch_ans_chain(v, n, sb, !any_float, rows * tiles_x, true);
        if (trace) cudaEventRecord(tr[1 + b * 5 + 3], sb);
        CK(cudaStreamWaitEvent(sb, eng->band_lf[b], 0));
        launch_ans_pack(v, eng->templ, n, sb);
        if (trace) cudaEventRecord(tr[1 + b * 5 + 4], sb);
        CK(cudaEventRecord(eng->band_done[b], sb));
        CK(cudaStreamWaitEvent(eng->st, eng->band_done[b], 0));
        eng->launches += 5;
    }
    if (trace) {
        cudaDeviceSynchronize();
        for (uint32_t b = 0; b < nbands; b++) {
            float t[5];
            for (int k = 0; k < 5; k++) cudaEventElapsedTime(&t[k], tr[0], tr[1 + b * 5 + k]);
            fprintf(stderr, "[bandtrace] band %u: start %.3f  xyb done %.3f  tokens done %.3f  chain done %.3f  pack done %.3f ms\n",
                    b, t[0], t[1], t[2], t[3], t[4]);
        }
    }
    return HYD_OK;
}

HYDStatusCode hydb_encode_image_device(HydbEngine *eng, const void *d_pixels, uint32_t width, uint32_t height,
                                       uint32_t channels, int64_t row_stride, int sample_fmt, int linear_light,
                                       uint32_t tile_row_begin, uint32_t tile_row_end, int with_header, uint8_t *d_out,
                                       uint64_t d_out_cap, uint64_t *out_len) {
    if (!eng || !d_pixels || !d_out || !out_len || !width || !height || channels < 3 ||
        (sample_fmt != HYD_UINT8 && sample_fmt != HYD_UINT16 && sample_fmt != HYD_FLOAT32)) {
        if (eng) eng->error = "invalid arguments to hydb_encode_image_device";
        return HYD_API_ERROR;
    }
    const uint32_t tiles_x = (width + 255) >> 8, tiles_y = (height + 255) >> 8;
    if (tile_row_end > tiles_y)
        tile_row_end = tiles_y;
    if (tile_row_begin >= tile_row_end) {
        eng->error = "empty tile row range";
        return HYD_API_ERROR;
    }
    CK(cudaSetDevice(eng->device));
    uint64_t pos = 0;
    const uint64_t ntiles = (uint64_t)tiles_x * (tile_row_end - tile_row_begin);
    std::vector<HydbTile> tiles;
    // Band pipelining pays while the chains are few and long (the table kernel's regime: a band's chains start
    // while later bands are still in their front-end kernels); a launch of thousands of tiles is bound by
    // chain throughput and every band adds a tail, so it runs as one plain sequence instead (8192 tiles of
    // config 5: 4096-tile batches 49.6 ms plain, 55.5 ms banded).  HYDRIUM_B200_BAND_TILES moves the limit.
    static const uint32_t band_tiles = [] { const char *e = getenv("HYDRIUM_B200_BAND_TILES"); int v = e ? atoi(e) : 0;
                                            return (uint32_t)(v > 0 ? v : 1024); }();
    uint32_t rows_per_batch = (eng->max_batch < band_tiles ? eng->max_batch : band_tiles) / tiles_x;
    if (ntiles <= eng->max_batch && ntiles <= band_tiles)
        rows_per_batch = tile_row_end - tile_row_begin;
    if (rows_per_batch >= 2 && tile_row_end - tile_row_begin >= 2 && !eng->timing && ntiles <= band_tiles) {
        // whole tile rows per batch, each batch band-pipelined (see launch_bands).  With per-kernel timing
        // on, the plain single-stream sequence below is used so that the events bracket whole kernels.
        const size_t item = sample_item_bytes(sample_fmt);
        for (uint32_t r0 = tile_row_begin, r1; r0 < tile_row_end; r0 = r1) {
            r1 = tile_row_end - r0 > rows_per_batch ? r0 + rows_per_batch : tile_row_end;
            const uint32_t n = (r1 - r0) * tiles_x;
            image_tiles(tiles, (const uint8_t *)d_pixels + (int64_t)(r0 - tile_row_begin) * 256 * row_stride * (int64_t)item, width,
                        height, channels, row_stride, sample_fmt, linear_light, r0, r1, with_header && r0 == tile_row_begin);
            HYDStatusCode rc = prepare_tiles(eng, tiles.data(), n, eng->st);
            if (rc == HYD_OK)
                rc = launch_bands(eng, tiles_x, r1 - r0, nullptr, nullptr, 0, 0, sample_fmt == HYD_FLOAT32);
            if (rc != HYD_OK)
                return rc;
            launch_gather(eng->ws, n, d_out, d_out_cap, pos, eng->d_overflow, eng->st);
            eng->launches += 2;
            rc = queue_readback(eng, n, pos);
            if (rc != HYD_OK)
                return rc;
            uint64_t total = 0;
            rc = hydb_engine_finish(eng, &total);
            if (rc != HYD_OK)
                return rc;
            pos += total;
        }
        *out_len = pos;
        return HYD_OK;
    }
    for (uint64_t first = 0; first < ntiles; first += eng->max_batch) {
        const uint32_t n = (uint32_t)((ntiles - first) < eng->max_batch ? (ntiles - first) : eng->max_batch);
        const uint32_t r0 = tile_row_begin + (uint32_t)(first / tiles_x);
        // batches are cut at arbitrary tiles: build the descriptors of whole rows and slice
        const uint32_t r1 = tile_row_begin + (uint32_t)((first + n + tiles_x - 1) / tiles_x);
        const size_t item = sample_item_bytes(sample_fmt);
        image_tiles(tiles, (const uint8_t *)d_pixels + (int64_t)(r0 - tile_row_begin) * 256 * row_stride * (int64_t)item, width,
                    height, channels, row_stride, sample_fmt, linear_light, r0, r1, 0);
        const size_t skip = (size_t)(first - (uint64_t)(r0 - tile_row_begin) * tiles_x);
        if (with_header && first == 0)
            tiles[0].with_image_header = 1;
        HYDStatusCode rc = hydb_engine_encode_tiles(eng, tiles.data() + skip, n, d_out, d_out_cap, pos);
        if (rc < HYD_ERROR_START)
            return rc;
        uint64_t total = 0;
        rc = hydb_engine_finish(eng, &total);
        if (rc != HYD_OK)
            return rc;
        pos += total;
    }
    *out_len = pos;
    return HYD_OK;
}

void *hydb_host_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
void hydb_host_free(void *p) { if (p) cudaFreeHost(p); }
void *hydb_device_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
}
void hydb_device_free(void *p) { if (p) cudaFree(p); }
int hydb_memcpy_h2d(void *dst, const void *src, size_t bytes) {
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}
int hydb_memcpy_d2h(void *dst, const void *src, size_t bytes) {
    return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
// ---- peer-memory gather (one process per GPU): CUDA IPC plumbing + the compaction launch ---------------
int hydb_ipc_export(const void *d_ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)) != cudaSuccess)
        return -1;
    memcpy(handle, &h, 64);
    return 0;
}
void *hydb_ipc_open(const uint8_t handle[64]) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void hydb_ipc_close(void *p) {
    if (p)
        cudaIpcCloseMemHandle(p);
}

// Stream-ordered store of one 64-bit word into device (or mapped peer) memory.  A host-side cudaMemcpy
// from pageable memory may return before the bytes have landed, which is not good enough for a
// length that another GPU reads right after the barrier.
HYDStatusCode hydb_engine_store_u64(HydbEngine *eng, void *d_dst, uint64_t value) {
    if (!eng || !d_dst || ((uintptr_t)d_dst & 7))
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    launch_store_u64(static_cast<uint64_t *>(d_dst), value, eng->st);
    eng->launches++;
    CK(cudaGetLastError());
    return HYD_OK;
}

HYDStatusCode hydb_engine_compact_regions(HydbEngine *eng, const uint8_t *d_regions, uint32_t nregions, uint64_t region_stride,
                                          uint8_t *d_out, uint64_t d_out_cap, uint64_t *total) {
    if (!eng || !d_regions || !nregions || nregions > 64 || region_stride < 512 || (region_stride & 15) || !d_out || !total) {
        if (eng) eng->error = "invalid arguments to hydb_engine_compact_regions";
        return HYD_API_ERROR;
    }
    CK(cudaSetDevice(eng->device));
    uint64_t *d_total = eng->d_small;
    uint32_t *d_ovf = reinterpret_cast<uint32_t *>(d_total + 1);
    volatile uint64_t *res = eng->h_small;
    HYDStatusCode rc = HYD_OK;
    if (cudaMemsetAsync(d_total, 0, 16, eng->st) != cudaSuccess)
        rc = HYD_INTERNAL_ERROR;
    if (rc == HYD_OK) {
        launch_compact_regions(d_regions, nregions, region_stride, d_out, d_out_cap, d_total, d_ovf, eng->st);
        eng->launches++;
        if (cudaMemcpyAsync(eng->h_small, d_total, 16, cudaMemcpyDeviceToHost, eng->st) != cudaSuccess ||
            cudaStreamSynchronize(eng->st) != cudaSuccess)
            rc = HYD_INTERNAL_ERROR;
    }
    if (rc == HYD_OK && (uint32_t)res[1]) {
        eng->error = "gathered codestream does not fit the output buffer";
        rc = HYD_INTERNAL_ERROR;
    }
    if (rc != HYD_OK && eng->error.empty())
        eng->error = "CUDA failure while compacting the gathered spans";
    *total = rc == HYD_OK ? res[0] : 0;
    return rc;
}

int hydb_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

static HYDStatusCode grow_device(HydbEngine *eng, void **p, size_t *cap, size_t need) {
    if (*cap >= need)
        return HYD_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (cudaMalloc(p, need) != cudaSuccess) {
        eng->error = "device allocation failed";
        return HYD_NOMEM;
    }
    *cap = need;
    return HYD_OK;
}

HYDStatusCode hydb_encode_image_host(HydbEngine *eng, const void *h_pixels, uint32_t width, uint32_t height,
                                     uint32_t channels, int sample_fmt, int linear_light, uint8_t *h_out,
                                     uint64_t h_out_cap, uint64_t *out_len) {
    if (!eng || !h_pixels || !h_out || !out_len || !width || !height || channels < 3 ||
        (sample_fmt != HYD_UINT8 && sample_fmt != HYD_UINT16 && sample_fmt != HYD_FLOAT32)) {
        if (eng) eng->error = "invalid arguments to hydb_encode_image_host";
        return HYD_API_ERROR;
    }
    if ((uint64_t)width * height > (UINT64_C(1) << 40) / channels) {
        eng->error = "image too large";
        return HYD_API_ERROR;
    }
    CK(cudaSetDevice(eng->device));
    const size_t item = sample_item_bytes(sample_fmt);
    const size_t in_bytes = (size_t)width * height * channels * item;
    HYDStatusCode rc = grow_device(eng, &eng->host_in, &eng->host_in_cap, in_bytes);
    if (rc == HYD_OK)
        rc = grow_device(eng, (void **)&eng->host_out, &eng->host_out_cap, (size_t)h_out_cap);
    if (rc != HYD_OK)
        return rc;
    const uint32_t tiles_x = (width + 255) >> 8, tiles_y = (height + 255) >> 8;
    const uint64_t ntiles = (uint64_t)tiles_x * tiles_y;
    if (ntiles > eng->max_batch || tiles_y < 2) {
        // several launches anyway: plain copy, then the device path
        CK(cudaMemcpyAsync(eng->host_in, h_pixels, in_bytes, cudaMemcpyHostToDevice, eng->st));
        rc = hydb_encode_image_device(eng, eng->host_in, width, height, channels, (int64_t)width * channels, sample_fmt,
                                      linear_light, 0, tiles_y, 1, eng->host_out, h_out_cap, out_len);
        if (rc != HYD_OK)
            return rc;
        CK(cudaMemcpyAsync(h_out, eng->host_out, *out_len, cudaMemcpyDeviceToHost, eng->st));
        CK(cudaStreamSynchronize(eng->st));
        return HYD_OK;
    }
    std::vector<HydbTile> tiles;
    image_tiles(tiles, eng->host_in, width, height, channels, (int64_t)width * channels, sample_fmt, linear_light, 0,
                tiles_y, 1);
    rc = prepare_tiles(eng, tiles.data(), (uint32_t)ntiles, eng->st);
    if (rc != HYD_OK)
        return rc;
    rc = launch_bands(eng, tiles_x, tiles_y, h_pixels, eng->host_in, (size_t)width * channels * item, height, sample_fmt == HYD_FLOAT32);
    if (rc != HYD_OK)
        return rc;
    launch_gather(eng->ws, (uint32_t)ntiles, eng->host_out, h_out_cap, 0, eng->d_overflow, eng->st);
    eng->launches += 2;
    rc = queue_readback(eng, (uint32_t)ntiles, 0);
    if (rc != HYD_OK)
        return rc;
    uint64_t total = 0;
    rc = hydb_engine_finish(eng, &total);
    if (rc != HYD_OK)
        return rc;
    *out_len = total;
    CK(cudaMemcpyAsync(h_out, eng->host_out, total, cudaMemcpyDeviceToHost, eng->st));
    CK(cudaStreamSynchronize(eng->st));
    return HYD_OK;
}

HYDStatusCode hydb_engine_enable_timing(HydbEngine *eng, int enable) {
    if (!eng)
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    if (enable && !eng->tev[0]) {
        for (cudaEvent_t &ev : eng->tev) CK(cudaEventCreate(&ev));
        for (cudaEvent_t &ev : eng->lev) CK(cudaEventCreate(&ev));
    }
    eng->timing = enable != 0;
    for (double &v : eng->stage_ms) v = 0;
    return HYD_OK;
}

// accumulated device milliseconds since the last call:
// xyb_dct, hf_tokens, ans_chain, ans_pack (incl. any wait for lf_group), gather, lf_group, #batches
HYDStatusCode hydb_engine_stage_ms(HydbEngine *eng, double out[7]) {
    if (!eng || !out)
        return HYD_API_ERROR;
    for (int k = 0; k < 7; k++) {
        out[k] = eng->stage_ms[k];
        eng->stage_ms[k] = 0;
    }
    return HYD_OK;
}

int hydb_synth_fill(HydbEngine *eng, void *d_dst, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0,
                    uint32_t full_width, uint32_t full_height, int bits, uint32_t seed, int smooth) {
    if (!eng || !d_dst || (bits != 8 && bits != 16))
        return HYD_API_ERROR;
    CK(cudaSetDevice(eng->device));
    launch_synth_fill(d_dst, width, height, x0, y0, full_width, full_height, bits, seed, smooth, eng->st);
    CK(cudaStreamSynchronize(eng->st));
    CK(cudaGetLastError());
    return HYD_OK;
}

int64_t hydb_engine_read_tap(HydbEngine *eng, int what, uint32_t tile, void *dst, uint64_t cap) {
    if (!eng || tile >= eng->last_n || !dst)
        return HYD_API_ERROR;
    if (cudaSetDevice(eng->device) != cudaSuccess || cudaStreamSynchronize(eng->st) != cudaSuccess)
        return HYD_INTERNAL_ERROR;
    const Workspace &w = eng->ws;
    const void *src = nullptr;
    uint64_t bytes = 0;
    uint32_t tmp = 0;
    auto read_u32 = [&](const uint32_t *p) {
        return cudaMemcpy(&tmp, p, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    };
    switch (what) {
    case HYDB_TAP_XYB: src = w.dbg_xyb ? w.dbg_xyb + (size_t)tile * 65536 * 3 : nullptr; bytes = 65536 * 3 * 4; break;
    case HYDB_TAP_DCT: src = w.dbg_dct ? w.dbg_dct + (size_t)tile * 65536 * 3 : nullptr; bytes = 65536 * 3 * 4; break;
    case HYDB_TAP_COEF: src = w.coef + (size_t)tile * kMaxBlocks * 3 * 64; bytes = kMaxBlocks * 3 * 64 * 2; break;
    case HYDB_TAP_NZINFO: src = w.nzinfo + (size_t)tile * kMaxBlocks * 3; bytes = kMaxBlocks * 3 * 2; break;
    case HYDB_TAP_LFQ: src = w.lfq + (size_t)tile * 3 * kMaxBlocks; bytes = 3 * kMaxBlocks * 4; break;
    case HYDB_TAP_SYMS:
        if (!read_u32(w.nsyms + tile)) return HYD_INTERNAL_ERROR;
        src = w.syms + (size_t)tile * kMaxHfSyms; bytes = (uint64_t)tmp * 4; break;
    case HYDB_TAP_FREQS: src = w.dbg_freqs ? w.dbg_freqs + (size_t)tile * kHfClusters * kHfTokens : nullptr;
        bytes = kHfClusters * kHfTokens * 4; break;
    case HYDB_TAP_LFBITS:
        if (!read_u32(w.lfbitlen + tile)) return HYD_INTERNAL_ERROR;
        src = w.lfbits + (size_t)tile * kLfBitsWords; bytes = (((uint64_t)tmp + 31) / 32) * 4; break;
    case HYDB_TAP_CLK: src = w.dbg_clk ? w.dbg_clk + (size_t)tile * 4 : nullptr; bytes = 16; break;
    case HYDB_TAP_SECT: src = w.dbg_sect ? w.dbg_sect + (size_t)tile * 4 : nullptr; bytes = 16; break;
    case HYDB_TAP_PAYLOAD: {
        uint32_t off = 0;
        if (!read_u32(w.frame_off + tile)) return HYD_INTERNAL_ERROR;
        off = tmp;
        if (!read_u32(w.frame_len + tile)) return HYD_INTERNAL_ERROR;
        if (tmp < kSlabHeaderReserve - off) return HYD_INTERNAL_ERROR;
        src = w.slab + (size_t)tile * kSlabBytes + kSlabHeaderReserve; bytes = tmp - (kSlabHeaderReserve - off); break;
    }
    case HYDB_TAP_NSYMS: src = w.nsyms + tile; bytes = 4; break;
    case HYDB_TAP_LFBITLEN: src = w.lfbitlen + tile; bytes = 4; break;
    default: return HYD_API_ERROR;
    }
    if (!src)
        return HYD_API_ERROR;
    if (bytes > cap)
        return HYD_NEED_MORE_OUTPUT;
    if (bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
        return HYD_INTERNAL_ERROR;
    return (int64_t)bytes;
}

}  // extern "C"

