/*
 * oracle/hyd_oracle.c -- TEST INFRASTRUCTURE ONLY (see hyd_oracle.h).
 *
 * CPU restatement of the reference tile encoder, one function per pipeline stage, each
 * citing the reference lines it follows (paths relative to /root/reference/src/libhydrium
 * unless noted).  Written from the algorithm description in SURVEY.md; organised for
 * stage-by-stage comparison with the CUDA kernels rather than like the reference.
 *
 * Float arithmetic must be plain IEEE binary32 with no contraction: build with
 * -ffp-contract=off on baseline x86-64.
 */
#include "hyd_oracle.h"

#include <stdlib.h>
#include <math.h>
#include <string.h>

static const char *g_err = NULL;
const char *orc_last_error(void) { return g_err; }
#define FAIL(code, msg) do { g_err = (msg); return (code); } while (0)

/* ------------------------------------------------------------------ math-functions.h:8-88 */
static int floor_log2(uint64_t v) { return 63 - __builtin_clzll(v); }
static int ceil_log2(uint64_t v) { return floor_log2(v) + ((v & (v - 1)) != 0); }
static uint32_t zigzag_sign(int32_t v) { uint32_t w = (uint32_t)v; return (w << 1) ^ (0u - (w >> 31)); }

/* ------------------------------------------------------------------ bit sink
 * LSB-first bit string (bitwriter.c:110-124).  The reference's 64-bit cache, realloc and
 * overflow machinery have no effect on the bits; only their order is restated. */
typedef struct Bits {
    uint8_t *data;
    uint64_t cap_bytes;
    uint64_t nbits;
    int oom;
} Bits;

static void bits_init(Bits *b) { memset(b, 0, sizeof(*b)); }
static void bits_free(Bits *b) { free(b->data); memset(b, 0, sizeof(*b)); }

static void bits_put(Bits *b, uint64_t value, int n) {
    if (n <= 0 || b->oom)
        return;
    uint64_t need = (b->nbits + (uint64_t)n + 7) / 8 + 8;
    if (need > b->cap_bytes) {
        uint64_t ncap = b->cap_bytes ? b->cap_bytes * 2 : 4096;
        while (ncap < need)
            ncap *= 2;
        uint8_t *np = realloc(b->data, ncap);
        if (!np) { b->oom = 1; return; }
        memset(np + b->cap_bytes, 0, ncap - b->cap_bytes);
        b->data = np;
        b->cap_bytes = ncap;
    }
    if (n < 64)
        value &= (UINT64_C(1) << n) - 1;
    while (n > 0) {
        unsigned off = (unsigned)(b->nbits & 7);
        int take = 8 - (int)off;
        if (take > n)
            take = n;
        b->data[b->nbits >> 3] |= (uint8_t)((value & ((1u << take) - 1)) << off);
        value >>= take;
        b->nbits += (uint64_t)take;
        n -= take;
    }
}
static void bits_bool(Bits *b, int f) { bits_put(b, f ? 1 : 0, 1); }
static void bits_align(Bits *b) { bits_put(b, 0, (int)((8 - (b->nbits & 7)) & 7)); }       /* bitwriter.c:126-128 */
static void bits_append(Bits *dst, const Bits *src) {                                     /* bitwriter.c:80-108 */
    uint64_t i = 0;
    for (; i + 32 <= src->nbits; i += 32) {
        uint32_t w;
        memcpy(&w, src->data + (i >> 3), 4);
        bits_put(dst, w, 32);
    }
    for (; i < src->nbits; i++)
        bits_put(dst, (src->data[i >> 3] >> (i & 7)) & 1, 1);
}

typedef struct U32Dist { uint32_t c[4]; uint32_t u[4]; } U32Dist;
static int bits_u32(Bits *b, const U32Dist *d, uint32_t v) {                               /* bitwriter.c:134-142 */
    for (int i = 0; i < 4; i++) {
        uint64_t lim = (UINT64_C(1) << d->u[i]) - 1;
        uint64_t x = (uint64_t)(uint32_t)(v - d->c[i]);
        if (x <= lim) {
            bits_put(b, (x << 2) | (uint64_t)i, (int)d->u[i] + 2);
            return 0;
        }
    }
    return -1;
}
static void bits_u64(Bits *b, uint64_t v) {                                                /* bitwriter.c:152-172 */
    if (!v) { bits_put(b, 0, 2); return; }
    if (v < 17) { bits_put(b, ((v - 1) << 2) | 1, 6); return; }
    if (v < 273) { bits_put(b, ((v - 17) << 2) | 2, 10); return; }
    bits_put(b, ((v & 0xFFF) << 2) | 3, 14);
    for (int shift = 12;; shift += 8) {
        uint64_t rest = v >> shift;
        if (!rest) { bits_put(b, 0, 1); return; }
        if (shift == 60) { bits_put(b, ((rest & 0xF) << 1) | 1, 5); return; }
        bits_put(b, ((rest & 0xFF) << 1) | 1, 9);
    }
}

/* ------------------------------------------------------------------ constant tables */
static const uint8_t k_level10_prefix[49] = {                                              /* encoder.c:23-30 */
    0, 0, 0, 0x0c, 'J', 'X', 'L', ' ', 0x0d, 0x0a, 0x87, 0x0a, 0, 0, 0, 0x14, 'f', 't', 'y', 'p',
    'j', 'x', 'l', ' ', 0, 0, 0, 0, 'j', 'x', 'l', ' ', 0, 0, 0, 9, 'j', 'x', 'l', 'l', 0x0a,
    0, 0, 0, 0, 'j', 'x', 'l', 'c',
};
/* encoder.c:32-40 as IEEE bit patterns */
static const uint32_t k_cos_bits[7][8] = {
    {0x3e318a87, 0x3e1682f9, 0x3dc92352, 0x3d0d42a9, 0xbd0d42a9, 0xbdc92352, 0xbe1682f9, 0xbe318a87},
    {0x3e273d5c, 0x3d8a8bd2, 0xbd8a8bd2, 0xbe273d5c, 0xbe273d5c, 0xbd8a8bd2, 0x3d8a8bd2, 0x3e273d5c},
    {0x3e1682f9, 0xbd0d42a9, 0xbe318a87, 0xbdc92352, 0x3dc92352, 0x3e318a87, 0x3d0d42a9, 0xbe1682f9},
    {0x3e000000, 0xbe000000, 0xbe000000, 0x3e000000, 0x3e000000, 0xbe000000, 0xbe000000, 0x3e000000},
    {0x3dc92352, 0xbe318a87, 0x3d0d42a9, 0x3e1682f9, 0xbe1682f9, 0xbd0d42a9, 0x3e318a87, 0xbdc92352},
    {0x3d8a8bd2, 0xbe273d5c, 0x3e273d5c, 0xbd8a8bd2, 0xbd8a8bd2, 0x3e273d5c, 0xbe273d5c, 0x3d8a8bd2},
    {0x3d0d42a9, 0xbdc92352, 0x3e1682f9, 0xbe318a87, 0x3e318a87, 0xbe1682f9, 0x3dc92352, 0xbd0d42a9},
};
static float cos_entry(int k, int n) { float f; memcpy(&f, &k_cos_bits[k][n], 4); return f; }

/* encoder.c:42-51: scan order as (column, row) pairs inside the stored (transposed) block */
static const uint8_t k_scan_x[64] = {
    0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4,
    3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 5, 6, 7, 7, 6, 7,
};
static const uint8_t k_scan_y[64] = {
    0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3,
    4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 4, 5, 6, 7, 7, 6, 5, 6, 7, 7,
};
static const uint8_t k_freq_ctx[64] = {                                                    /* encoder.c:53-58 */
    0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20,
    21, 21, 22, 22, 23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27,
    28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30,
};
static uint32_t nnz_ctx(uint32_t nz) {                                                     /* encoder.c:60-66 */
    static const uint8_t first[9] = {0, 0, 31, 62, 62, 93, 93, 93, 93};
    if (nz < 9) return first[nz];
    if (nz < 13) return 123;
    if (nz < 21) return 152;
    if (nz < 33) return 180;
    return 206;
}
static const uint16_t k_hf_weights[3][64] = {                                              /* encoder.c:74-93 */
    {1969, 1969, 1969, 1962, 1969, 1962, 1655, 1885, 1885, 1655, 1397, 1610, 1704, 1610, 1397, 1178,
     1368, 1494, 1494, 1368, 1178, 994, 1159, 1289, 1340, 1289, 1159, 994, 839, 980, 1104, 1178,
     1178, 1104, 980, 839, 829, 941, 1023, 1054, 1023, 941, 829, 800, 881, 928, 928, 881,
     800, 755, 809, 829, 809, 755, 663, 731, 731, 663, 491, 524, 491, 349, 349, 239},
    {280, 280, 280, 279, 280, 279, 245, 271, 271, 245, 214, 239, 250, 239, 214, 188,
     211, 226, 226, 211, 188, 164, 185, 201, 207, 201, 185, 164, 144, 163, 178, 188,
     188, 178, 163, 144, 143, 157, 168, 172, 168, 157, 143, 139, 150, 156, 156, 150,
     139, 133, 140, 143, 140, 133, 125, 129, 129, 125, 116, 118, 116, 107, 107, 98},
    {256, 147, 147, 85, 117, 85, 60, 78, 78, 60, 43, 56, 63, 56, 43, 43,
     43, 48, 48, 43, 43, 42, 43, 43, 43, 43, 43, 42, 29, 41, 43, 43,
     43, 43, 41, 29, 29, 37, 43, 43, 43, 37, 29, 27, 33, 36, 36, 33,
     27, 24, 27, 29, 27, 24, 20, 22, 22, 20, 15, 16, 15, 10, 10, 7},
};
static const U32Dist k_size_hdr = {{1, 1, 1, 1}, {9, 13, 18, 30}};                        /* encoder.c:98-101 */
static const U32Dist k_frame_size = {{0, 256, 2304, 18688}, {8, 11, 14, 30}};             /* encoder.c:102-105 */
static const U32Dist k_global_scale = {{1, 2049, 4097, 8193}, {11, 11, 12, 16}};          /* encoder.c:106-109 */
static const U32Dist k_quant_lf = {{16, 1, 1, 1}, {0, 5, 8, 16}};                         /* encoder.c:110-113 */
static const U32Dist k_toc = {{0, 1024, 17408, 4211712}, {10, 14, 22, 30}};               /* encoder.c:117-120 */
static const U32Dist k_lz_min_symbol = {{224, 512, 4096, 8}, {0, 0, 0, 15}};              /* entropy.c:48-51 */
static const U32Dist k_lz_min_length = {{3, 4, 5, 9}, {0, 0, 2, 8}};                      /* entropy.c:52-55 */

/* ------------------------------------------------------------------ hybrid integers */
typedef struct HybridCfg { uint8_t split, msb, lsb; } HybridCfg;

static OrcSymbol hybrid_split(uint32_t v, HybridCfg c) {                                  /* entropy.c:427-444 */
    OrcSymbol s = {0, 0, 0, 0};
    if (v < (1u << c.split)) {
        s.token = v;
        return s;
    }
    uint32_t n = (uint32_t)floor_log2(v) - c.lsb - c.msb;
    uint32_t low = v & ((1u << c.lsb) - 1);
    v >>= c.lsb;
    s.residue = v & ((1u << n) - 1);
    v >>= n;
    uint32_t high = v & ((1u << c.msb) - 1);
    s.nbits = n;
    s.token = (1u << c.split) + (low | (high << c.lsb) | ((n - c.split + c.lsb + c.msb) << (c.msb + c.lsb)));
    return s;
}

/* ------------------------------------------------------------------ entropy stream (front end) */
typedef struct Stream {
    uint32_t num_dists;
    uint8_t *cmap;
    uint32_t num_clusters;
    OrcSymbol *syms;
    uint64_t n, cap;
    uint32_t alpha[256];
    uint32_t max_alpha;
    HybridCfg cfg[256];
    uint32_t lz_min_symbol, lz_min_length;
    uint32_t last_plus1, last_dist, run;
    int modular;
    uint32_t *freq[256];
    int oom;
} Stream;

static void stream_free(Stream *s) {
    for (int i = 0; i < 256; i++)
        free(s->freq[i]);
    free(s->cmap);
    free(s->syms);
    memset(s, 0, sizeof(*s));
}

static void stream_set_cfg(Stream *s, uint32_t from, uint32_t to, int split, int msb, int lsb) { /* entropy.c:91-106 */
    for (uint32_t j = from; (!to || j < to) && j < s->num_clusters; j++) {
        s->cfg[j].split = (uint8_t)split;
        s->cfg[j].msb = (uint8_t)msb;
        s->cfg[j].lsb = (uint8_t)lsb;
    }
}

static int stream_init(Stream *s, const uint8_t *cmap, uint32_t num_dists, int custom_cfg,         /* entropy.c:371-425 */
                       uint32_t lz_min_symbol, int modular) {
    memset(s, 0, sizeof(*s));
    uint32_t plain = num_dists;
    if (lz_min_symbol) {
        num_dists++;
        s->lz_min_length = 3;
        s->lz_min_symbol = lz_min_symbol;
    }
    s->num_dists = num_dists;
    s->modular = modular;
    s->cmap = calloc(num_dists, 1);
    if (!s->cmap)
        return ORC_NOMEM;
    for (uint32_t i = 0; i < plain; i++) {
        s->cmap[i] = cmap ? cmap[i] : 0;
        if (s->cmap[i] >= s->num_clusters)
            s->num_clusters = s->cmap[i] + 1u;
    }
    if (lz_min_symbol)
        s->cmap[num_dists - 1] = (uint8_t)s->num_clusters++;
    if (!custom_cfg) {
        stream_set_cfg(s, 0, s->num_clusters - (lz_min_symbol ? 1 : 0), 4, 1, 1);
        if (lz_min_symbol)
            stream_set_cfg(s, s->num_clusters - 1, s->num_clusters, 7, 0, 0);
    }
    return ORC_OK;
}

static void stream_push(Stream *s, OrcSymbol sym) {                                       /* entropy.c:446-464 */
    if (s->n == s->cap) {
        uint64_t ncap = s->cap ? s->cap * 2 : 1024;
        OrcSymbol *np = realloc(s->syms, ncap * sizeof(OrcSymbol));
        if (!np) { s->oom = 1; return; }
        s->syms = np;
        s->cap = ncap;
    }
    s->syms[s->n++] = sym;
    if (sym.token + 1 > s->max_alpha)
        s->max_alpha = sym.token + 1;
    if (sym.token + 1 > s->alpha[sym.cluster])
        s->alpha[sym.cluster] = sym.token + 1;
}

static void stream_emit(Stream *s, uint32_t dist, uint32_t value) {                       /* entropy.c:466-471 */
    uint32_t cl = s->cmap[dist];
    OrcSymbol sym = hybrid_split(value, s->cfg[cl]);
    sym.cluster = cl;
    stream_push(s, sym);
}

static void stream_end_run(Stream *s) {                                                   /* entropy.c:473-500 */
    if (s->run > s->lz_min_length) {
        const HybridCfg len_cfg = {7, 0, 0};
        OrcSymbol sym = hybrid_split(s->run - s->lz_min_length, len_cfg);
        sym.cluster = s->cmap[s->last_dist];
        sym.token += s->lz_min_symbol;
        stream_push(s, sym);
        stream_emit(s, s->num_dists - 1, s->modular ? 1 : 0);
    } else if (s->last_plus1 && s->run) {
        for (uint32_t k = 0; k < s->run; k++)
            stream_emit(s, s->last_dist, s->last_plus1 - 1);
    }
    s->run = 0;
}

static void stream_send(Stream *s, uint32_t dist, uint32_t value) {                       /* entropy.c:502-524 */
    if (!s->lz_min_symbol) {
        stream_emit(s, dist, value);
        return;
    }
    if (s->last_plus1 == value + 1 && s->cmap[s->last_dist] == s->cmap[dist] && s->run < 127) {
        s->run++;
        return;
    }
    stream_end_run(s);
    s->last_plus1 = value + 1;
    s->last_dist = dist;
    stream_emit(s, dist, value);
}

static int stream_count(Stream *s) {                                                      /* entropy.c:526-544 */
    for (uint32_t c = 0; c < s->num_clusters; c++) {
        free(s->freq[c]);
        s->freq[c] = NULL;
        if (!s->alpha[c])
            continue;
        s->freq[c] = calloc(s->alpha[c], sizeof(uint32_t));
        if (!s->freq[c])
            return ORC_NOMEM;
    }
    for (uint64_t i = 0; i < s->n; i++)
        s->freq[s->syms[i].cluster][s->syms[i].token]++;
    return ORC_OK;
}

/* ------------------------------------------------------------------ stream headers */
static int prefix_stream_finish(Stream *s, Bits *bw);

static void put_hybrid_cfg(Bits *bw, HybridCfg c, int log_alpha) {                        /* entropy.c:169-182 */
    bits_put(bw, c.split, ceil_log2(1 + (uint64_t)log_alpha));
    if (c.split == log_alpha)
        return;
    bits_put(bw, c.msb, ceil_log2(1 + (uint64_t)c.split));
    bits_put(bw, c.lsb, ceil_log2(1 + (uint64_t)c.split - c.msb));
}

static int put_cluster_map(const uint8_t *cmap, uint32_t num_dists, uint32_t num_clusters, Bits *bw) { /* entropy.c:108-167 */
    if (num_dists == 1)
        return ORC_OK;
    int nbits = ceil_log2(num_clusters);
    if (nbits <= 3 && (uint64_t)num_dists * (uint64_t)nbits <= 32) {
        bits_bool(bw, 1);
        bits_put(bw, (uint64_t)nbits, 2);
        for (uint32_t i = 0; i < num_dists; i++)
            bits_put(bw, cmap[i], nbits);
        return ORC_OK;
    }
    bits_bool(bw, 0);
    bits_bool(bw, 1); /* move-to-front */
    Stream nested;
    int ret = stream_init(&nested, NULL, 1, 1, 64, 0);
    if (ret < 0)
        return ret;
    stream_set_cfg(&nested, 0, 0, 4, 1, 0);
    uint8_t mtf[256];
    for (int i = 0; i < 256; i++)
        mtf[i] = (uint8_t)i;
    for (uint32_t j = 0; j < num_dists; j++) {
        int idx = 0;
        while (mtf[idx] != cmap[j])
            idx++;
        stream_send(&nested, 0, (uint32_t)idx);
        if (idx) {
            uint8_t v = mtf[idx];
            memmove(mtf + 1, mtf, (size_t)idx);
            mtf[0] = v;
        }
    }
    ret = prefix_stream_finish(&nested, bw);
    stream_free(&nested);
    return ret;
}

static int put_stream_preamble(Stream *s, Bits *bw, int log_alpha) {                      /* entropy.c:546-575 */
    bits_bool(bw, s->lz_min_symbol != 0);
    if (s->lz_min_symbol) {
        const HybridCfg len_cfg = {7, 0, 0};
        stream_end_run(s);
        if (bits_u32(bw, &k_lz_min_symbol, s->lz_min_symbol) || bits_u32(bw, &k_lz_min_length, s->lz_min_length))
            FAIL(ORC_INTERNAL_ERROR, "lz77 parameter out of range");
        put_hybrid_cfg(bw, len_cfg, 8);
    }
    int ret = put_cluster_map(s->cmap, s->num_dists, s->num_clusters, bw);
    if (ret < 0)
        return ret;
    bits_bool(bw, log_alpha == 0);
    if (log_alpha)
        bits_put(bw, (uint64_t)(log_alpha - 5), 2);
    for (uint32_t i = 0; i < s->num_clusters; i++)
        put_hybrid_cfg(bw, s->cfg[i], log_alpha ? log_alpha : 15);
    return ORC_OK;
}

/* ------------------------------------------------------------------ length-limited code lengths
 * entropy.c:577-662.  A 2n-1 slot array; pass k moves the two cheapest eligible live nodes
 * to slots 2k, 2k+1 and parks their parent in slot n+k.  Slot positions decide ties between
 * equal-weight internal nodes, so the array and its swaps are simulated literally. */
typedef struct HNode {
    uint32_t weight;
    int32_t symbol_plus1;   /* 0 for internal nodes */
    int32_t depth, reach;   /* reach = deepest depth below (and including) this node */
    int32_t kid0, kid1;     /* slot indices, -1 if none */
} HNode;

static int hnode_before(const HNode *a, const HNode *b) {                                 /* entropy.c:577-581 */
    if (a->weight != b->weight)
        return a->weight < b->weight;
    if (!b->symbol_plus1)
        return 1;
    if (!a->symbol_plus1)
        return 0;
    return a->symbol_plus1 < b->symbol_plus1;
}

static int32_t hnode_deepen(HNode *t, int32_t i) {                                        /* entropy.c:583-590 */
    if (i < 0)
        return 0;
    int32_t self = ++t[i].depth;
    int32_t a = hnode_deepen(t, t[i].kid0);
    int32_t b = hnode_deepen(t, t[i].kid1);
    int32_t m = self > a ? self : a;
    return t[i].reach = (m > b ? m : b);
}

static int code_lengths(const uint32_t *weights, uint32_t *lengths, uint32_t n, int32_t limit) { /* entropy.c:592-662 */
    HNode *t = calloc(2 * (size_t)n - 1, sizeof(HNode));
    if (!t)
        FAIL(ORC_NOMEM, "out of memory");
    uint32_t live = 0;
    for (uint32_t i = 0; i < 2 * n - 1; i++)
        t[i].kid0 = t[i].kid1 = -1;
    for (uint32_t i = 0; i < n; i++) {
        t[i].weight = weights[i];
        t[i].symbol_plus1 = (int32_t)i + 1;
        live += weights[i] != 0;
    }
    if (!live) {
        free(t);
        FAIL(ORC_INTERNAL_ERROR, "No nonzero frequencies");
    }
    for (uint32_t k = 0; k + 1 < n; k++, live--) {
        int32_t best = -1, next = -1;
        int32_t bound = limit - ceil_log2(live) + 1;
        for (uint32_t j = 2 * k; j < n + k; j++) {
            if (!t[j].weight || t[j].reach >= bound)
                continue;
            if (best < 0 || hnode_before(&t[j], &t[best])) {
                next = best;
                best = (int32_t)j;
            } else if (next < 0 || hnode_before(&t[j], &t[next])) {
                next = (int32_t)j;
            }
        }
        if (best < 0) {
            free(t);
            FAIL(ORC_INTERNAL_ERROR, "couldn't find target");
        }
        HNode tmp = t[best]; t[best] = t[2 * k]; t[2 * k] = tmp;
        if (next < 0)
            break;
        if ((uint32_t)next == 2 * k)
            next = best;
        tmp = t[next]; t[next] = t[2 * k + 1]; t[2 * k + 1] = tmp;
        HNode *p = &t[n + k];
        p->weight = t[2 * k].weight + t[2 * k + 1].weight;
        p->kid0 = (int32_t)(2 * k);
        p->kid1 = (int32_t)(2 * k + 1);
        hnode_deepen(t, (int32_t)(n + k));
    }
    for (uint32_t j = 0; j < 2 * n - 1; j++)
        if (t[j].symbol_plus1)
            lengths[t[j].symbol_plus1 - 1] = (uint32_t)t[j].depth;
    free(t);
    return ORC_OK;
}

static uint32_t reverse_bits32(uint32_t v) {                                              /* entropy.c:60-69 */
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}

typedef struct Code { uint32_t bits, len; } Code;

/* canonical codes by (length, symbol), stored bit-reversed for LSB-first output (entropy.c:664-707) */
static int assign_codes(Code *table, const uint32_t *lengths, uint32_t n) {
    uint64_t next = 0;
    for (uint32_t len = 1; len <= 32; len++) {
        for (uint32_t s = 0; s < n; s++) {
            if (lengths[s] != len)
                continue;
            table[s].bits = reverse_bits32((uint32_t)next);
            table[s].len = len;
            next += UINT64_C(1) << (32 - len);
        }
    }
    if (next && next != (UINT64_C(1) << 32))
        FAIL(ORC_INTERNAL_ERROR, "VLC codes do not add up");
    return ORC_OK;
}

static const uint8_t k_clc_order[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15}; /* entropy.c:42 */
static const Code k_clc_code[6] = {{0, 2}, {7, 4}, {3, 3}, {2, 2}, {1, 2}, {15, 4}};                   /* entropy.c:44-46 */

static void put_zero_run(Bits *bw, const Code *l1, uint32_t zeros) {                      /* entropy.c:709-728 */
    if (zeros >= 3) {
        uint32_t part[8];
        int k = 0;
        while (zeros > 10) {
            uint32_t up = (zeros + 13) / 8;
            part[k++] = zeros - 8 * up + 16;
            zeros = up;
        }
        part[k++] = zeros;
        while (k--) {
            bits_put(bw, l1[17].bits, (int)l1[17].len);
            bits_put(bw, part[k] - 3, 3);
        }
    } else {
        while (zeros--)
            bits_put(bw, l1[0].bits, (int)l1[0].len);
    }
}

static int put_complex_code(Bits *bw, uint32_t n, const uint32_t *lengths) {              /* entropy.c:730-805 */
    bits_put(bw, 0, 2);
    uint32_t l1w[18] = {0};
    uint32_t zeros = 0;
    for (uint32_t j = 0; j < n; j++) {
        if (!lengths[j]) {
            zeros++;
            continue;
        }
        if (zeros >= 3) {
            while (zeros > 10) {
                l1w[17]++;
                zeros = (zeros + 13) / 8;
            }
            l1w[17]++;
        } else {
            l1w[0] += zeros;
        }
        zeros = 0;
        l1w[lengths[j]]++;
    }
    uint32_t l1len[18] = {0};
    int ret = code_lengths(l1w, l1len, 18, 5);
    if (ret < 0)
        return ret;
    uint32_t space = 0;
    for (int j = 0; j < 18; j++) {
        uint32_t len = l1len[k_clc_order[j]];
        bits_put(bw, k_clc_code[len].bits, (int)k_clc_code[len].len);
        if (len)
            space += 32u >> len;
        if (space >= 32)
            break;
    }
    if (space && space != 32)
        FAIL(ORC_INTERNAL_ERROR, "level1 code total mismatch");
    Code l1[18];
    memset(l1, 0, sizeof(l1));
    ret = assign_codes(l1, l1len, 18);
    if (ret < 0)
        return ret;
    space = 0;
    zeros = 0;
    for (uint32_t j = 0; j < n; j++) {
        uint32_t len = lengths[j];
        if (!len) {
            zeros++;
            continue;
        }
        put_zero_run(bw, l1, zeros);
        zeros = 0;
        bits_put(bw, l1[len].bits, (int)l1[len].len);
        space += 32768u >> len;
        if (space == 32768)
            break;
    }
    put_zero_run(bw, l1, zeros);
    return ORC_OK;
}

/* header + symbols of a prefix-coded stream (entropy.c:807-941, 1003-1034) */
static int prefix_stream_finish(Stream *s, Bits *bw) {
    int ret = put_stream_preamble(s, bw, 0);
    if (ret < 0)
        return ret;
    if (s->oom)
        FAIL(ORC_NOMEM, "out of memory");
    ret = stream_count(s);
    if (ret < 0)
        return ret;
    Code *codes[256] = {0};
    uint32_t *lengths = calloc(s->max_alpha ? s->max_alpha : 1, sizeof(uint32_t));
    if (!lengths)
        FAIL(ORC_NOMEM, "out of memory");
    for (uint32_t c = 0; c < s->num_clusters; c++) {
        if (s->alpha[c] <= 1) {
            bits_bool(bw, 0);
            continue;
        }
        bits_bool(bw, 1);
        int nb = floor_log2(s->alpha[c] - 1);
        bits_put(bw, (uint64_t)nb, 4);
        bits_put(bw, s->alpha[c] - 1, nb);
    }
    for (uint32_t c = 0; c < s->num_clusters && ret >= 0; c++) {
        uint32_t n = s->alpha[c];
        codes[c] = calloc(n ? n : 1, sizeof(Code));
        if (!codes[c]) { ret = ORC_NOMEM; break; }
        if (n <= 1)
            continue;
        memset(lengths, 0, s->max_alpha * sizeof(uint32_t));
        ret = code_lengths(s->freq[c], lengths, n, 15);
        if (ret < 0)
            break;
        uint32_t used = 0;
        struct { uint32_t sym, len; } few[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}}, sw;
        for (uint32_t j = 0; j < n; j++) {
            if (!lengths[j])
                continue;
            if (used < 4) {
                few[used].sym = j;
                few[used].len = lengths[j];
            }
            if (++used > 4)
                break;
        }
        if (used > 4) {
            ret = put_complex_code(bw, n, lengths);
            if (ret >= 0)
                ret = assign_codes(codes[c], lengths, n);
            continue;
        }
        if (!used) {
            used = 1;
            few[0].sym = n - 1;
        }
        bits_put(bw, 1, 2);
        bits_put(bw, used - 1, 2);
#define SWAP_FEW(a, b) do { sw = few[a]; few[a] = few[b]; few[b] = sw; } while (0)
        if (used == 3 && few[0].len != 1) {                                               /* entropy.c:888-894 */
            if (few[1].len == 1) SWAP_FEW(0, 1); else SWAP_FEW(0, 2);
        }
        int select = 0;
        if (used == 4) {                                                                  /* entropy.c:896-919 */
            for (int i = 0; i < 4; i++)
                if (few[i].len != 2) { select = 1; break; }
            if (select && few[0].len != 1) {
                if (few[1].len == 1) SWAP_FEW(0, 1);
                else if (few[2].len == 1) SWAP_FEW(0, 2);
                else SWAP_FEW(0, 3);
            }
            if (select && few[1].len != 2) {
                if (few[2].len == 2) SWAP_FEW(1, 2); else SWAP_FEW(1, 3);
            }
        }
#undef SWAP_FEW
        int width = ceil_log2(n);
        for (uint32_t i = 0; i < used; i++)
            bits_put(bw, few[i].sym, width);
        if (used == 4)
            bits_bool(bw, select);
        ret = assign_codes(codes[c], lengths, n);
    }
    if (ret >= 0) {
        for (uint64_t i = 0; i < s->n; i++) {                                             /* entropy.c:1003-1021 */
            const OrcSymbol *y = &s->syms[i];
            bits_put(bw, codes[y->cluster][y->token].bits, (int)codes[y->cluster][y->token].len);
            bits_put(bw, y->residue, (int)y->nbits);
        }
    }
    for (int c = 0; c < 256; c++)
        free(codes[c]);
    free(lengths);
    if (ret == ORC_NOMEM)
        FAIL(ORC_NOMEM, "out of memory");
    return ret;
}

/* ------------------------------------------------------------------ ANS back end */
static int ans_normalise(uint32_t *f, uint32_t n) {                                       /* entropy.c:267-301 */
    uint64_t total = 0;
    for (uint32_t k = 0; k < n; k++)
        total += f[k];
    if (!total)
        return -1;
    uint64_t sum = 0;
    for (uint32_t k = 0; k < n; k++) {
        if (!f[k])
            continue;
        f[k] = (uint32_t)((((uint64_t)f[k] << 12) / total) & 0xFFFF);
        if (!f[k])
            f[k] = 1;
        sum += f[k];
    }
    uint32_t j = n - 1;
    while (sum > 4096) {
        uint64_t excess = sum - 4096;
        if (excess < f[j]) {
            f[j] -= (uint32_t)excess;
            sum -= excess;
            break;
        } else if (f[j] > 1) {
            sum -= f[j] - 1;
            f[j] = 1;
        }
        j--;
    }
    f[0] += (uint32_t)(4096 - sum);
    return f[n - 1] == 4096;
}

/* Alias table in the reference's per-symbol "entry list" form (entropy.c:184-265). */
typedef struct AliasRow {
    uint32_t count;
    int32_t cut[258], off[258], bucket[258];
} AliasRow;

static int ans_alias(const uint32_t *f, uint32_t alpha, int log_alpha, int32_t single, AliasRow *rows) {
    const uint32_t bucket = 1u << (12 - log_alpha), slots = 1u << log_alpha;
    uint32_t owner[256] = {0}, cut[256] = {0}, off[256] = {0};
    if (single >= 0) {
        for (uint32_t i = 0; i < slots; i++) {
            owner[i] = (uint32_t)single;
            off[i] = i * bucket;
        }
        rows[single].count = slots;
    } else {
        uint8_t small[256], large[256];
        uint32_t ns = 0, nl = 0;
        for (uint32_t p = 0; p < alpha; p++) {
            cut[p] = f[p];
            if (cut[p] < bucket) small[ns++] = (uint8_t)p;
            else if (cut[p] > bucket) large[nl++] = (uint8_t)p;
        }
        for (uint32_t i = alpha; i < slots; i++)
            small[ns++] = (uint8_t)i;
        while (nl) {
            if (!ns)
                FAIL(ORC_INTERNAL_ERROR, "empty underfull during alias table gen");
            uint8_t u = small[--ns], o = large[--nl];
            uint32_t by = bucket - cut[u];
            off[u] = (cut[o] -= by);
            owner[u] = o;
            if (cut[o] < bucket) small[ns++] = o;
            else if (cut[o] > bucket) large[nl++] = o;
        }
        for (uint32_t s = 0; s < slots; s++) {
            if (cut[s] == bucket) {
                owner[s] = s;
                cut[s] = off[s] = 0;
            } else {
                off[s] -= cut[s];
            }
            rows[owner[s]].count++;
        }
    }
    for (uint32_t s = 0; s < alpha; s++) {
        memset(rows[s].cut, -1, sizeof(rows[s].cut));
        memset(rows[s].off, -1, sizeof(rows[s].off));
        memset(rows[s].bucket, -1, sizeof(rows[s].bucket));
        rows[s].off[0] = 0;
        rows[s].cut[0] = (int32_t)cut[s];
        rows[s].bucket[0] = (int32_t)s;
    }
    for (uint32_t i = 0; i < slots; i++) {
        AliasRow *r = &rows[owner[i]];
        int j = 1;
        while (r->cut[j] >= 0)
            j++;
        r->cut[j] = (int32_t)cut[i];
        r->off[j] = (int32_t)off[i];
        r->bucket[j] = (int32_t)i;
    }
    return ORC_OK;
}

static void put_ans_u8(Bits *bw, uint32_t b) {                                            /* entropy.c:71-78 */
    bits_bool(bw, b != 0);
    if (!b)
        return;
    int l = floor_log2(b);
    bits_put(bw, (uint64_t)l, 3);
    bits_put(bw, b, l);
}

static void put_ans_histogram(Bits *bw, const uint32_t *f, uint32_t alpha) {              /* entropy.c:303-369 */
    static const Code logcount_code[14] = {                                               /* entropy.c:35-38 */
        {17, 5}, {11, 4}, {15, 4}, {3, 4}, {9, 4}, {7, 4}, {4, 3},
        {2, 3}, {5, 3}, {6, 3}, {0, 3}, {33, 6}, {1, 7}, {65, 7},
    };
    if (!alpha) {
        bits_put(bw, 1, 2);
        put_ans_u8(bw, 0);
        return;
    }
    int32_t a = -1, b = -1, seen = 0;
    for (uint32_t k = 0; k < alpha; k++) {
        if (f[k] == 4096) {
            bits_put(bw, 1, 2);
            put_ans_u8(bw, k);
            return;
        }
        if (!f[k])
            continue;
        if (++seen > 2)
            break;
        if (a < 0) {
            a = (int32_t)k;
        } else if (f[a] + f[k] == 4096) {
            b = (int32_t)k;
            break;
        }
    }
    if (a >= 0 && b >= 0) {
        bits_put(bw, 3, 2);
        put_ans_u8(bw, (uint32_t)a);
        put_ans_u8(bw, (uint32_t)b);
        bits_put(bw, f[a], 12);
        return;
    }
    bits_put(bw, 0, 2);
    bits_put(bw, 7, 3);
    bits_put(bw, 6, 3);
    put_ans_u8(bw, alpha - 3);
    int logc[256];
    uint32_t omit = 0;
    int omit_log = 0;
    for (uint32_t k = 0; k < alpha; k++) {
        logc[k] = f[k] ? 1 + floor_log2(f[k]) : 0;
        bits_put(bw, logcount_code[logc[k]].bits, (int)logcount_code[logc[k]].len);
        if (logc[k] > omit_log) {
            omit_log = logc[k];
            omit = k;
        }
    }
    for (uint32_t k = 0; k < alpha; k++) {
        if (k == omit || logc[k] <= 1)
            continue;
        bits_put(bw, f[k], logc[k] - 1);
    }
}

typedef struct AnsModel {
    int log_alpha;
    AliasRow *rows[256];
} AnsModel;

static void ans_model_free(AnsModel *m) {
    for (int i = 0; i < 256; i++)
        free(m->rows[i]);
}

static int ans_prepare(Stream *s, AnsModel *m) {                                          /* entropy.c:943-978 */
    memset(m, 0, sizeof(*m));
    int ret = stream_count(s);
    if (ret < 0)
        FAIL(ORC_NOMEM, "out of memory");
    int l = ceil_log2(s->max_alpha);
    m->log_alpha = l > 5 ? l : 5;
    for (uint32_t c = 0; c < s->num_clusters; c++) {
        if (!s->alpha[c])
            continue;
        int single = ans_normalise(s->freq[c], s->alpha[c]);
        if (single < 0)
            FAIL(ORC_INTERNAL_ERROR, "all-zero ANS frequencies");
        m->rows[c] = calloc(s->alpha[c], sizeof(AliasRow));
        if (!m->rows[c])
            FAIL(ORC_NOMEM, "out of memory");
        ret = ans_alias(s->freq[c], s->alpha[c], m->log_alpha, single ? (int32_t)s->alpha[c] - 1 : -1, m->rows[c]);
        if (ret < 0)
            return ret;
    }
    return ORC_OK;
}

static int ans_put_header(Stream *s, const AnsModel *m, Bits *bw) {                       /* entropy.c:980-1001 */
    int ret = put_stream_preamble(s, bw, m->log_alpha);
    if (ret < 0)
        return ret;
    for (uint32_t c = 0; c < s->num_clusters; c++)
        put_ans_histogram(bw, s->freq[c], s->alpha[c]);
    return ORC_OK;
}

/* Reverse rANS pass, then forward emission (entropy.c:1064-1159).  The reference stores the
 * distance between consecutive renormalisation points in a uint16_t (entropy.c:16, 1094), so
 * gaps >= 65536 symbols wrap; `gap16` reproduces that. */
static int ans_put_symbols(const Stream *s, const AnsModel *m, Bits *bw, uint64_t first, uint64_t count) {
    typedef struct Flush { uint16_t gap16, word; } Flush;
    const uint32_t log_bucket = 12u - (uint32_t)m->log_alpha;
    const uint32_t pos_mask = (1u << log_bucket) - 1;
    Flush *fl = malloc((count + 3) * sizeof(Flush));
    if (!fl)
        FAIL(ORC_NOMEM, "out of memory");
    uint64_t nfl = 0;
    uint32_t state = 0x130000u;
    const OrcSymbol *sy = s->syms + first;
    uint64_t last_push = count;
    uint16_t last_word = 0;
    for (uint64_t r = 0; r < count; r++) {
        const uint64_t p = count - 1 - r;
        const uint32_t tok = sy[p].token & 0xFF;
        const uint32_t cl = sy[p].cluster;
        const uint32_t fr = s->freq[cl][tok];
        if ((state >> 20) >= fr) {
            if (last_push != count) {
                fl[nfl].gap16 = (uint16_t)(last_push - p);
                fl[nfl++].word = last_word;
            }
            last_push = p;
            last_word = (uint16_t)(state & 0xFFFF);
            state >>= 16;
        }
        const uint32_t q = state / fr, rem = state - q * fr;
        const AliasRow *row = &m->rows[cl][tok];
        uint32_t j, pos = 0, slot = 0;
        for (j = 0; j <= row->count; j++) {
            pos = rem - (uint32_t)row->off[j];
            int32_t k = (int32_t)(pos - (uint32_t)row->cut[j]);
            if (!(pos & ~pos_mask) && (j > 0 ? k >= 0 : k < 0)) {
                slot = (uint32_t)row->bucket[j];
                break;
            }
        }
        if (j > row->count) {
            free(fl);
            FAIL(ORC_INTERNAL_ERROR, "alias table lookup failed");
        }
        state = (q << 12) | (slot << log_bucket) | pos;
    }
    if (last_push != count) {
        fl[nfl].gap16 = (uint16_t)last_push;
        fl[nfl++].word = last_word;
    }
    fl[nfl].gap16 = 0; fl[nfl++].word = (uint16_t)(state >> 16);
    fl[nfl].gap16 = 0; fl[nfl++].word = (uint16_t)(state & 0xFFFF);
    uint64_t last_pop = 0;
    for (uint64_t p = 0; p < count; p++) {
        while (nfl && p - last_pop >= fl[nfl - 1].gap16) {
            bits_put(bw, fl[nfl - 1].word, 16);
            last_pop = p;
            nfl--;
        }
        bits_put(bw, sy[p].residue, (int)sy[p].nbits);
    }
    free(fl);
    return ORC_OK;
}

/* ------------------------------------------------------------------ colour transform */
static float srgb_to_linear(float x) {                                                    /* format.c:15-19 */
    if (x <= 0.0404482362771082f)
        return 0.07739938080495357f * x;
    return 0.003094300919832f + x * (-0.009982599f + x * (0.72007737769f + 0.2852804880f * x));
}
static float fast_cbrt(float x) {                                                         /* format.c:21-27 */
    union { float f; uint32_t i; } z;
    z.f = x;
    z.i = 0x548c39cbu - z.i / 3u;
    z.f *= 1.5015480449f - 0.534850249f * x * z.f * z.f * z.f;
    z.f *= 1.333333985f - 0.33333333f * x * z.f * z.f * z.f;
    return 1.0f / z.f;
}
static float opsin_bias(float x) { return fast_cbrt(x + 0.0037930732552754493f) - 0.155954f; } /* format.c:29-31 */

void orc_build_luts(int sample_fmt, int linear_light, uint16_t *input_lut, float *bias_lut) { /* format.c:58-83 */
    const size_t n = sample_fmt == ORC_UINT8 ? 256 : 65536;
    const float step = 1.0f / (n - 1.0f);
    for (size_t i = 0; i < n; i++) {
        float f = i * step;
        if (!linear_light)
            f = srgb_to_linear(f);
        int32_t v = (int32_t)(f * 65535.f + 0.5f);                                        /* format.c:33-36 */
        input_lut[i] = (uint16_t)(v < 0 ? 0 : v > 65535 ? 65535 : v);
    }
    const float step16 = 1.0f / (65536 - 1.0f);
    for (size_t i = 0; i < 65536; i++)
        bias_lut[i] = opsin_bias(i * step16);
}

/* pixel -> XYB with zero padding of partial blocks (format.c:48-56, 85-109, 182-191);
 * float samples skip both tables (format.c:38-46, 111-140).  Returns 0, or -1 for a non-finite sample. */
static int stage_xyb(const OrcTile *t, uint32_t w, uint32_t h, uint32_t stride, uint32_t rows,
                     const uint16_t *in_lut, const float *bias_lut, float *xyb) {
    memset(xyb, 0, (size_t)stride * rows * 3 * sizeof(float));
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            ptrdiff_t o = (ptrdiff_t)y * t->row_stride + (ptrdiff_t)x * t->pixel_stride;
            if (t->sample_fmt == ORC_FLOAT32) {
                float fr = ((const float *)t->plane[0])[o], fg = ((const float *)t->plane[1])[o],
                      fb = ((const float *)t->plane[2])[o];
                /* the reference sets "Invalid NaN Float" but drops the status (format.c:123-126, 168-172)
                 * and goes on with a partly stale buffer; both the oracle and the product refuse instead */
                if (!isfinite(fr) || !isfinite(fg) || !isfinite(fb))
                    return -1;
                if (!t->linear_light) {
                    fr = srgb_to_linear(fr);
                    fg = srgb_to_linear(fg);
                    fb = srgb_to_linear(fb);
                }
                const float l = opsin_bias(0.3f * fr + 0.622f * fg + 0.078f * fb);
                const float m = opsin_bias(0.23f * fr + 0.692f * fg + 0.078f * fb);
                const float s = opsin_bias(0.243423f * fr + 0.204767f * fg + 0.55181f * fb);
                const float Y = (l + m) * 0.5f;
                float *px = xyb + ((size_t)y * stride + x) * 3;
                px[0] = Y - m;
                px[1] = Y;
                px[2] = s - Y;
                continue;
            }
            uint32_t r, g, b;
            if (t->sample_fmt == ORC_UINT8) {
                r = in_lut[((const uint8_t *)t->plane[0])[o]];
                g = in_lut[((const uint8_t *)t->plane[1])[o]];
                b = in_lut[((const uint8_t *)t->plane[2])[o]];
            } else {
                r = in_lut[((const uint16_t *)t->plane[0])[o]];
                g = in_lut[((const uint16_t *)t->plane[1])[o]];
                b = in_lut[((const uint16_t *)t->plane[2])[o]];
            }
            const float l = bias_lut[((19661u * r + 40761u * g + 5112u * b) >> 16) & 0xFFFFu];
            const float m = bias_lut[((15073u * r + 45350u * g + 5112u * b) >> 16) & 0xFFFFu];
            const float s = bias_lut[((15953u * r + 13419u * g + 36163u * b) >> 16) & 0xFFFFu];
            const float Y = (l + m) * 0.5f;
            float *px = xyb + ((size_t)y * stride + x) * 3;
            px[0] = Y - m;
            px[1] = Y;
            px[2] = s - Y;
        }
    }
    return 0;
}

/* 8-point transform in the reference's summation order (encoder.c:641-648) */
static void dct8(const float in[8], float out[8]) {
    float dc = in[0];
    for (int n = 1; n < 8; n++)
        dc += in[n];
    out[0] = dc * 0.125f;
    for (int k = 1; k < 8; k++) {
        float acc = 0.0f;
        for (int n = 0; n < 8; n++)
            acc += in[n] * cos_entry(k - 1, n);
        out[k] = acc;
    }
}

/* rows, then columns, result stored transposed (encoder.c:631-668) */
static void stage_dct(float *xyb, uint32_t vbw, uint32_t vbh) {
    const size_t stride = (size_t)vbw * 8;
    for (int c = 0; c < 3; c++)
        for (uint32_t by = 0; by < vbh; by++)
            for (uint32_t bx = 0; bx < vbw; bx++) {
                float rowpass[8][8], colpass[8][8], v[8], o[8];
                for (int y = 0; y < 8; y++) {
                    for (int x = 0; x < 8; x++)
                        v[x] = xyb[(((size_t)by * 8 + y) * stride + bx * 8 + x) * 3 + c];
                    dct8(v, rowpass[y]);
                }
                for (int x = 0; x < 8; x++) {
                    for (int y = 0; y < 8; y++)
                        v[y] = rowpass[y][x];
                    dct8(v, o);
                    for (int k = 0; k < 8; k++)
                        colpass[k][x] = o[k];
                }
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++)
                        xyb[(((size_t)by * 8 + y) * stride + bx * 8 + x) * 3 + c] = colpass[x][y];
            }
}

/* HF quantisation with dead zone, in place float -> int (encoder.c:783-823) */
static void stage_quant(float *xyb, uint32_t vbw, uint32_t vbh, uint8_t *nonzeroes) {
    const size_t stride = (size_t)vbw * 8;
    int32_t *ints = (int32_t *)xyb;
    memset(nonzeroes, 0, (size_t)vbw * vbh * 3);
    for (uint32_t by = 0; by < vbh; by++)
        for (uint32_t bx = 0; bx < vbw; bx++)
            for (int c = 0; c < 3; c++)
                for (int j = 1; j < 64; j++) {
                    size_t at = (((size_t)by * 8 + k_scan_y[j]) * stride + bx * 8 + k_scan_x[j]) * 3 + c;
                    int32_t q = (int32_t)(xyb[at] * k_hf_weights[c][j] * 5);
                    if (q > -2 && q < 2)
                        q = 0;
                    else
                        nonzeroes[((size_t)by * vbw + bx) * 3 + c]++;
                    ints[at] = q;
                }
}

/* ------------------------------------------------------------------ payload sections */
static void put_lf_global(Bits *bw) {                                                     /* encoder.c:510-537 */
    bits_bool(bw, 1);
    bits_u32(bw, &k_global_scale, 32768);
    bits_u32(bw, &k_quant_lf, 4);
    bits_bool(bw, 0);
    bits_put(bw, 0, 16);
    bits_bool(bw, 1);
    bits_put(bw, 2, 2);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 13; j++)
            bits_put(bw, (uint64_t)i, 2);
    bits_bool(bw, 1);
    bits_bool(bw, 0);
}

/* fixed five-node MA tree stream; `predictor` is 5 (gradient) for LF, 0 for HF metadata */
static int put_ma_tree(Bits *bw, uint32_t predictor) {                                    /* encoder.c:552-564, 600-610 */
    Stream st;
    int ret = stream_init(&st, NULL, 6, 0, 0, 0);
    if (ret < 0)
        FAIL(ORC_NOMEM, "out of memory");
    stream_send(&st, 1, 0);
    stream_send(&st, 2, predictor);
    stream_send(&st, 3, 0);
    stream_send(&st, 4, 0);
    stream_send(&st, 5, 0);
    ret = prefix_stream_finish(&st, bw);
    stream_free(&st);
    return ret;
}

/* LFGroup: LF coefficients (modular, gradient predictor) + HF metadata (encoder.c:539-629).
 * Converts the DC slot of every block from float to int in place, like the reference. */
static int put_lf_group(Bits *bw, float *xyb, uint32_t vbw, uint32_t vbh) {
    const size_t stride = (size_t)vbw * 8;
    int32_t *ints = (int32_t *)xyb;
    const uint32_t blocks = vbw * vbh;
    bits_put(bw, 0, 2);
    bits_bool(bw, 0);
    bits_bool(bw, 1);
    bits_put(bw, 0, 2);
    int ret = put_ma_tree(bw, 5);
    if (ret < 0)
        return ret;
    Stream st;
    ret = stream_init(&st, NULL, 1, 1, 1u << 14, 1);
    if (ret < 0)
        FAIL(ORC_NOMEM, "out of memory");
    stream_set_cfg(&st, 0, 0, 7, 1, 1);
    static const float scale[3] = {8192.f, 1024.f, 512.f};
    static const int order[3] = {1, 0, 2};
    for (int i = 0; i < 3; i++) {
        const int c = order[i];
        for (uint32_t by = 0; by < vbh; by++)
            for (uint32_t bx = 0; bx < vbw; bx++) {
                const size_t at = ((size_t)by * 8 * stride + (size_t)bx * 8) * 3 + c;
                ints[at] = (int32_t)(xyb[at] * scale[c]);
                const int32_t up = by ? ints[at - 8 * stride * 3] : 0;
                const int32_t w = bx ? ints[at - 8 * 3] : up;
                const int32_t n = by ? up : w;
                const int32_t nw = (bx && by) ? ints[at - 8 * stride * 3 - 8 * 3] : w;
                const int32_t lo = w < n ? w : n, hi = w < n ? n : w;
                int32_t pred = w + n - nw;
                pred = pred < lo ? lo : pred > hi ? hi : pred;
                stream_send(&st, 0, zigzag_sign(ints[at] - pred));
            }
    }
    ret = prefix_stream_finish(&st, bw);
    stream_free(&st);
    if (ret < 0)
        return ret;
    bits_put(bw, blocks - 1, ceil_log2(blocks));
    bits_put(bw, 2, 4);
    ret = put_ma_tree(bw, 0);
    if (ret < 0)
        return ret;
    const uint32_t cfl = ((vbw + 7) / 8) * ((vbh + 7) / 8);
    ret = stream_init(&st, NULL, 1, 0, 29, 1);
    if (ret < 0)
        FAIL(ORC_NOMEM, "out of memory");
    for (uint32_t i = 0; i < 2 * cfl + blocks; i++)
        stream_send(&st, 0, 0);
    for (uint32_t i = 0; i < blocks; i++)
        stream_send(&st, 0, (5 - 1) * 2);
    for (uint32_t i = 0; i < blocks; i++)
        stream_send(&st, 0, 0);
    ret = prefix_stream_finish(&st, bw);
    stream_free(&st);
    return ret;
}

/* HF symbol stream: per block Y, X, B: non-zero count, then coefficients (encoder.c:670-750) */
static void stage_hf_symbols(Stream *st, const int32_t *ints, const uint8_t *nonzeroes, uint32_t vbw, uint32_t vbh) {
    const size_t stride = (size_t)vbw * 8;
    static const int order[3] = {1, 0, 2};
    for (uint32_t by = 0; by < vbh; by++)
        for (uint32_t bx = 0; bx < vbw; bx++)
            for (uint32_t i = 0; i < 3; i++) {
                const int c = order[i];
                uint32_t pred;
                if (!bx && !by) pred = 32;
                else if (!bx) pred = nonzeroes[((size_t)(by - 1) * vbw) * 3 + c];
                else if (!by) pred = nonzeroes[((size_t)bx - 1) * 3 + c];
                else pred = (nonzeroes[((size_t)(by - 1) * vbw + bx) * 3 + c] +
                             (uint32_t)nonzeroes[((size_t)by * vbw + bx - 1) * 3 + c] + 1) >> 1;
                const uint32_t pctx = pred < 8 ? pred : 4 + ((pred > 64 ? 64 : pred) >> 1);
                uint32_t left = nonzeroes[((size_t)by * vbw + bx) * 3 + c];
                stream_send(st, 3 * pctx + i, left);
                if (!left)
                    continue;
                const uint32_t base = 458 * i + 111;
                for (int k = 0; k < 63; k++) {
                    const size_t prev_at = (((size_t)by * 8 + k_scan_y[k]) * stride + bx * 8 + k_scan_x[k]) * 3 + c;
                    const size_t at = (((size_t)by * 8 + k_scan_y[k + 1]) * stride + bx * 8 + k_scan_x[k + 1]) * 3 + c;
                    const uint32_t prev = k ? ints[prev_at] != 0 : left <= 4;
                    const uint32_t ctx = base + prev + ((nnz_ctx(left) + k_freq_ctx[k + 1]) << 1);
                    const uint32_t v = zigzag_sign(ints[at]);
                    stream_send(st, ctx, v);
                    if (v && !--left)
                        break;
                }
            }
}

static void hf_cluster_map(uint8_t map[1485]) {                                           /* encoder.c:862-877 */
    for (int j = 0; j < 111; j++)
        map[j] = (uint8_t)(j % 3);
    for (int j = 111; j < 1485; j++)
        map[j] = (uint8_t)(3 + (j - 111) % 6);
}

/* ------------------------------------------------------------------ headers */
int64_t orc_image_header(uint64_t width, uint64_t height, uint8_t *dst, uint64_t cap) {   /* encoder.c:164-239 */
    Bits bw;
    bits_init(&bw);
    uint64_t n = 0;
    if (width > (1u << 20) || height > (1u << 20) || width * height > (1u << 28)) {       /* libhydrium.c:67-68 */
        if (cap < sizeof(k_level10_prefix))
            FAIL(ORC_API_ERROR, "output too small");
        memcpy(dst, k_level10_prefix, sizeof(k_level10_prefix));
        n = sizeof(k_level10_prefix);
    }
    bits_put(&bw, 0x0AFF, 17);
    bits_u32(&bw, &k_size_hdr, (uint32_t)height);
    bits_put(&bw, 0, 3);
    bits_u32(&bw, &k_size_hdr, (uint32_t)width);
    bits_bool(&bw, 0);
    bits_bool(&bw, 0);
    bits_bool(&bw, 0);
    bits_put(&bw, 0, 2);
    bits_bool(&bw, 1);
    bits_put(&bw, 0, 2);
    bits_bool(&bw, 1);
    bits_bool(&bw, 1);
    bits_u64(&bw, 0);
    bits_bool(&bw, 1);
    bits_align(&bw);
    uint64_t bytes = bw.nbits / 8;
    if (bw.oom || n + bytes > cap) {
        bits_free(&bw);
        FAIL(ORC_API_ERROR, "output too small");
    }
    memcpy(dst + n, bw.data, bytes);
    bits_free(&bw);
    return (int64_t)(n + bytes);
}

static void put_frame_header(Bits *bw, const OrcTile *t, uint32_t w, uint32_t h, int last) { /* encoder.c:327-435 */
    const int crop = !(t->image_width <= w && t->image_height <= h);
    bits_put(bw, 0, 1);
    bits_put(bw, last ? 0 : 3, 2);
    bits_put(bw, 0, 1);
    bits_u64(bw, 0x80);
    bits_put(bw, 0x4C, 10);
    bits_bool(bw, crop);
    if (crop) {
        bits_u32(bw, &k_frame_size, zigzag_sign((int32_t)(t->tile_x * 256u)));
        bits_u32(bw, &k_frame_size, zigzag_sign((int32_t)(t->tile_y * 256u)));
        bits_u32(bw, &k_frame_size, w);
        bits_u32(bw, &k_frame_size, h);
    }
    bits_put(bw, 0, 2);
    if (crop)
        bits_put(bw, 0, 2);
    bits_bool(bw, last);
    if (!last)
        bits_put(bw, 0, 2);
    bits_put(bw, 0, 2);
    bits_bool(bw, 0);
    bits_bool(bw, 0);
    bits_put(bw, 0, 2);
    bits_put(bw, 0, 2);
    bits_put(bw, 0, 2);
    bits_bool(bw, 0); /* single TOC entry: not permuted */
    bits_align(bw);
}

/* ------------------------------------------------------------------ one tile */
static void snapshot(const Bits *bw, uint8_t *dst, uint64_t cap, uint64_t *bitlen) {
    *bitlen = bw->nbits;
    if (dst && (bw->nbits + 7) / 8 <= cap)
        memcpy(dst, bw->data, (bw->nbits + 7) / 8);
}

int64_t orc_encode_tile(const OrcTile *t, uint8_t *dst, uint64_t cap, OrcStages *stages) {
    g_err = NULL;
    if (t->sample_fmt != ORC_UINT8 && t->sample_fmt != ORC_UINT16 && t->sample_fmt != ORC_FLOAT32)
        FAIL(ORC_API_ERROR, "Invalid Sample Format");
    const uint64_t tiles_x = (t->image_width + 255) / 256, tiles_y = (t->image_height + 255) / 256;
    if (t->tile_x >= tiles_x || t->tile_y >= tiles_y)                                     /* encoder.c:448-451 */
        FAIL(ORC_API_ERROR, "tile out of bounds");
    const uint32_t w = (uint32_t)((t->tile_x + 1) * 256u > t->image_width ? t->image_width - t->tile_x * 256u : 256);
    const uint32_t h = (uint32_t)((t->tile_y + 1) * 256u > t->image_height ? t->image_height - t->tile_y * 256u : 256);
    const uint32_t vbw = (w + 7) / 8, vbh = (h + 7) / 8, stride = vbw * 8, rows = vbh * 8;
    const int last = t->is_last < 0 ? (t->tile_x + 1 == tiles_x && t->tile_y + 1 == tiles_y) : !!t->is_last; /* encoder.c:482-485 */
    const size_t nfloats = (size_t)stride * rows * 3;

    int64_t result = ORC_NOMEM;
    uint16_t *in_lut = malloc(65536 * sizeof(uint16_t));
    float *bias_lut = malloc(65536 * sizeof(float));
    float *xyb = malloc(nfloats * sizeof(float));
    uint8_t *nonzeroes = malloc((size_t)vbw * vbh * 3);
    Bits payload, frame;
    Stream hf;
    AnsModel model;
    bits_init(&payload);
    bits_init(&frame);
    memset(&hf, 0, sizeof(hf));
    memset(&model, 0, sizeof(model));
    if (!in_lut || !bias_lut || !xyb || !nonzeroes) {
        g_err = "out of memory";
        goto done;
    }
    if (t->sample_fmt != ORC_FLOAT32)
        orc_build_luts(t->sample_fmt, t->linear_light, in_lut, bias_lut);

    if (stage_xyb(t, w, h, stride, rows, in_lut, bias_lut, xyb) < 0) {
        g_err = "Invalid NaN Float";
        result = ORC_API_ERROR;
        goto done;
    }
    if (stages) {
        stages->vbw = vbw;
        stages->vbh = vbh;
        if (stages->xyb) memcpy(stages->xyb, xyb, nfloats * sizeof(float));
    }
    stage_dct(xyb, vbw, vbh);
    if (stages && stages->dct) memcpy(stages->dct, xyb, nfloats * sizeof(float));
    stage_quant(xyb, vbw, vbh, nonzeroes);

    put_lf_global(&payload);
    result = put_lf_group(&payload, xyb, vbw, vbh);
    if (result < 0)
        goto done;
    if (stages) {
        if (stages->quant) memcpy(stages->quant, xyb, nfloats * sizeof(float));
        if (stages->nonzeroes) memcpy(stages->nonzeroes, nonzeroes, (size_t)vbw * vbh * 3);
        snapshot(&payload, stages->lf_bits, stages->lf_bits_cap, &stages->lf_bitlen);
    }

    uint8_t cmap[1485];
    hf_cluster_map(cmap);
    result = stream_init(&hf, cmap, 1485, 1, 0, 0);                                       /* encoder.c:903-910 */
    if (result < 0)
        goto done;
    stream_set_cfg(&hf, 0, 0, 4, 1, 0);
    stage_hf_symbols(&hf, (const int32_t *)xyb, nonzeroes, vbw, vbh);
    if (hf.oom) {
        result = ORC_NOMEM;
        goto done;
    }
    if (stages) {
        stages->hf_syms_n = hf.n;
        if (stages->hf_syms && hf.n <= stages->hf_syms_cap)
            memcpy(stages->hf_syms, hf.syms, hf.n * sizeof(OrcSymbol));
    }
    result = ans_prepare(&hf, &model);
    if (result < 0)
        goto done;
    if (stages) {
        stages->max_alphabet_size = hf.max_alpha;
        for (uint32_t c = 0; c < hf.num_clusters && c < 16; c++) {
            stages->alphabet_sizes[c] = (uint16_t)hf.alpha[c];
            if (stages->freqs && c < 9)
                for (uint32_t k = 0; k < hf.alpha[c] && k < 256; k++)
                    stages->freqs[c * 256 + k] = hf.freq[c][k];
        }
    }
    Bits pass;
    bits_init(&pass);
    result = ans_put_symbols(&hf, &model, &pass, 0, hf.n);                                /* encoder.c:942-950 */
    if (result >= 0 && stages)
        snapshot(&pass, stages->ans_bits, stages->ans_bits_cap, &stages->ans_bitlen);
    if (result >= 0) {
        bits_bool(&payload, 1);                                                           /* encoder.c:959-967 */
        bits_put(&payload, 2, 2);
        result = ans_put_header(&hf, &model, &payload);
    }
    if (result >= 0) {
        if (stages)
            snapshot(&payload, stages->pre_bits, stages->pre_bits_cap, &stages->pre_bitlen);
        bits_append(&payload, &pass);                                                     /* encoder.c:973-981 */
        bits_align(&payload);                                                             /* encoder.c:984 */
    }
    bits_free(&pass);
    if (result < 0)
        goto done;

    put_frame_header(&frame, t, w, h, last);
    if (bits_u32(&frame, &k_toc, (uint32_t)(payload.nbits / 8))) {                        /* encoder.c:1002 */
        result = ORC_INTERNAL_ERROR;
        g_err = "TOC entry out of range";
        goto done;
    }
    bits_align(&frame);
    if (payload.oom || frame.oom) {
        result = ORC_NOMEM;
        g_err = "out of memory";
        goto done;
    }
    {
        const uint64_t fb = frame.nbits / 8, pb = payload.nbits / 8;
        if (fb + pb > cap) {
            result = ORC_API_ERROR;
            g_err = "output too small";
            goto done;
        }
        memcpy(dst, frame.data, fb);
        memcpy(dst + fb, payload.data, pb);
        result = (int64_t)(fb + pb);
    }
done:
    ans_model_free(&model);
    stream_free(&hf);
    bits_free(&payload);
    bits_free(&frame);
    free(in_lut);
    free(bias_lut);
    free(xyb);
    free(nonzeroes);
    return result;
}

int64_t orc_encode_image(const void *pixels, uint64_t width, uint64_t height, int channels,
                         int sample_fmt, int linear_light, uint8_t *dst, uint64_t cap) {
    g_err = NULL;
    if (!width || !height)                                                                /* libhydrium.c:48-51 */
        FAIL(ORC_API_ERROR, "invalid zero-width or zero-height");
    int64_t n = orc_image_header(width, height, dst, cap);
    if (n < 0)
        return n;
    const size_t item = sample_fmt == ORC_UINT8 ? 1 : (sample_fmt == ORC_UINT16 ? 2 : 4);
    const uint64_t tiles_x = (width + 255) / 256, tiles_y = (height + 255) / 256;
    for (uint64_t ty = 0; ty < tiles_y; ty++)
        for (uint64_t tx = 0; tx < tiles_x; tx++) {
            OrcTile t;
            memset(&t, 0, sizeof(t));
            t.image_width = width;
            t.image_height = height;
            t.linear_light = linear_light;
            t.tile_x = (uint32_t)tx;
            t.tile_y = (uint32_t)ty;
            t.is_last = -1;
            t.sample_fmt = sample_fmt;
            t.row_stride = (ptrdiff_t)(width * (uint64_t)channels);
            t.pixel_stride = channels;
            const uint8_t *p = (const uint8_t *)pixels + ((ty * 256) * width + tx * 256) * (uint64_t)channels * item;
            t.plane[0] = p;
            t.plane[1] = p + item;
            t.plane[2] = p + 2 * item;
            int64_t got = orc_encode_tile(&t, dst + n, cap - (uint64_t)n, NULL);
            if (got < 0)
                return got;
            n += got;
        }
    return n;
}

int64_t orc_prefix_stream(const uint32_t *values, const uint32_t *ctx, uint64_t n,
                          const uint8_t *cluster_map, uint32_t num_dists, int custom_config,
                          int split, int msb, int lsb, uint32_t lz77_min_symbol, int modular,
                          uint8_t *dst, uint64_t cap, uint64_t *bitlen) {
    g_err = NULL;
    Stream st;
    Bits bw;
    bits_init(&bw);
    int ret = stream_init(&st, cluster_map, num_dists, custom_config, lz77_min_symbol, modular);
    if (ret < 0)
        FAIL(ORC_NOMEM, "out of memory");
    if (custom_config)
        stream_set_cfg(&st, 0, 0, split, msb, lsb);
    for (uint64_t i = 0; i < n; i++)
        stream_send(&st, ctx ? ctx[i] : 0, values[i]);
    ret = prefix_stream_finish(&st, &bw);
    stream_free(&st);
    int64_t out = ret;
    if (ret >= 0) {
        *bitlen = bw.nbits;
        const uint64_t bytes = (bw.nbits + 7) / 8;
        if (bw.oom || bytes > cap) {
            out = ORC_API_ERROR;
            g_err = "output too small";
        } else {
            memcpy(dst, bw.data, bytes);
            out = (int64_t)bytes;
        }
    }
    bits_free(&bw);
    return out;
}
