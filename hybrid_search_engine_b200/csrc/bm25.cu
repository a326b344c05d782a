// K1 -- BM25 scoring over a device-resident CSR inverted index (replaces bm25.py:83-127
// BM25.score / score_batch, called from pipelines.py:271,328,485).
//
// Design (doc-range tiles, no global atomics, no score buffer round trip through HBM during
// accumulation).  Two kernels implement it -- bm25_batch_kernel (default: persistent, several queries per tile,
// register-pipelined posting stream, impact table in shared memory; described where it is defined) and
// bm25_tile_kernel (one CTA per (tile, query); kept for A/B runs, HS_BM25_IMPL=tile) -- with the same arithmetic:
//   * a CTA (or a 256-thread group of it) owns kTileDocs consecutive docs and keeps their float64
//     partial scores in shared memory; the tile's doc lengths are staged into shared memory once
//     (coalesced) -- the "document-length table staged in shared memory" of the north star.
//   * for each query token IN QUERY ORDER (bm25.py:99, duplicates included) the CTA locates the
//     slice of that term's posting list that falls in its doc range (warp-wide 32-ary search on the
//     ascending doc ids), then streams the slice with coalesced 8-byte (doc_id, tf) loads.  A doc
//     occurs at most once per posting list, so `acc[doc] += contribution` is a plain shared-memory
//     read-modify-write; __syncthreads() between tokens keeps the reference's float64 summation order.
//   * arithmetic is the reference's, float64, no FMA contraction (bm25.py:104-110):
//         num = tf * (k1 + 1); den = tf + k1 * (1 - b + b * (dl / avgdl)); score += idf * (num / den)
//     the fraction num / den depends on the small integers (dl, tf) only and is gathered from a float64
//     table built once per index with the same instructions (hs_bm25_impact_table); pairs outside
//     the table are computed inline.
//   * epilogue: single rounding to float32 (bm25.py:124-126), coalesced store, tile max folded into
//     stats slot HS_STAT_MAX_B (pipelines.py:332).
//
// Algorithmic bytes per query: 8 * P(q) postings + 4 * N doc lengths + 4 * N score store.
#include "common.cuh"

#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace {

constexpr int kTileDocs = 4096;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTermGroup = 32;
constexpr int kUnroll = 4;        // postings in flight per thread in the streaming part   // query tokens resolved per search round

struct Bm25Params {
    const int64_t* indptr;
    const uint2* postings;
    const uint32_t* dl;
    const double* impact_table;
    uint32_t max_dl, tf_cap;
    double k1, one_minus_b, b, avgdl, k1p1;
    const int32_t* q_terms;
    const double* q_idf;
    const int32_t* q_off;
    int64_t n_docs, n_terms;
    float* scores;        // [B, n]
    __half* scores_h;     // batched kernel only: the scores as binary16 [B, ld_h] instead (scores == nullptr) -- the SCREEN
    int64_t ld_h;         // of the verified mode; the max folded into stats is still that of the float32 values
    uint32_t* stats;      // [B, 4] or null
    int64_t* ranges;      // [n_tokens, n_tiles + 1] posting offsets at every tile boundary (workspace)
    int n_tiles;
    double delta;         // BM25Plus only (bm25.py:150-179)
    // hot terms (df >= half the shard): their per-doc contribution idf * frac(dl, tf) as a DENSE float64 vector, built
    // once per index (hs_bm25_build_hot).  The batched kernel streams it like a posting chunk and adds it slot by slot
    const double* hot_c;         // [n_hot, n_docs]; 0.0 where the doc does not hold the term
    const int32_t* hot_of_term;  // [n_terms] -> row of hot_c, or -1
    int pf;                      // batched kernel: L2 prefetch distance in chunks
    int skip_hot;                // ranges kernel: hot terms need no posting offsets (the batched kernel streams hot_c)
};

// (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * (dl / avgdl))), or 0 where the reference adds 0 (den <= 0):
// float64, the reference's operation order, no contraction (bm25.py:107-110)
__device__ __noinline__ double bm25_frac_compute(double k1, double one_minus_b, double b, double avgdl,
                                                 double k1p1, uint32_t tf_u, uint32_t dl) {
    const double tf = (double)tf_u;
    const double kd = __dmul_rn(k1, __dadd_rn(one_minus_b, __dmul_rn(b, __ddiv_rn((double)dl, avgdl))));
    const double num = __dmul_rn(tf, k1p1);
    const double den = __dadd_rn(tf, kd);
    return den > 0.0 ? __ddiv_rn(num, den) : 0.0;
}
// The fraction depends on the small integers (dl, tf) only, so it is tabulated once per index with the
// same instructions (hs_bm25_impact_table): the hot loop replaces two float64 divisions by one gather.
__device__ __forceinline__ double bm25_frac(const Bm25Params& p, uint32_t tf, uint32_t dl) {
    if (p.impact_table != nullptr && dl <= p.max_dl && tf <= p.tf_cap)
        return __ldg(p.impact_table + (size_t)dl * (p.tf_cap + 1) + tf);
    return bm25_frac_compute(p.k1, p.one_minus_b, p.b, p.avgdl, p.k1p1, tf, dl);
}

// first index in [lo, hi) whose doc id >= target; whole warp cooperates (32-ary search)
__device__ __forceinline__ int64_t warp_lower_bound(const uint2* __restrict__ post, int64_t lo, int64_t hi,
                                                    uint32_t target, int lane) {
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 31) / 32;
        // probe positions lo + (lane+1)*step - 1, clamped
        int64_t pos = lo + (int64_t)(lane + 1) * step - 1;
        if (pos >= hi) pos = hi - 1;
        const bool ge = __ldg(&post[pos].x) >= target;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ge);
        if (m == 0) {
            lo = hi;   // every probe (the last is hi-1) is < target
            break;
        }
        const int f = __ffs(m) - 1;   // first probe that is >= target
        int64_t new_hi = lo + (int64_t)(f + 1) * step - 1;
        if (new_hi >= hi) new_hi = hi - 1;
        const int64_t new_lo = lo + (int64_t)f * step;   // everything before probe f-1 (incl.) is < target
        lo = new_lo;
        hi = new_hi + 1;   // answer is in [new_lo, new_hi]
        if (hi - lo <= 32) break;
    }
    // final: <= 32 candidates
    const int64_t pos = lo + lane;
    const bool ge = (pos < hi) ? (__ldg(&post[pos].x) >= target) : true;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, ge);
    if (m == 0) return hi;            // exactly 32 candidates, all below the target
    const int64_t r = lo + (__ffs(m) - 1);
    return r < hi ? r : hi;
}

// postings [lo, hi) of one token, kUnroll per thread in flight: loads, then impact gathers, then the
// shared-memory updates (docs within one list are distinct, so the updates of a batch never collide)
__device__ __forceinline__ void bm25_stream_slice(const Bm25Params& p, double* acc, const uint32_t* sdl, int64_t lo,
                                                  int64_t hi, double idf, uint32_t d_lo, int tid) {
    if (lo >= hi) return;
    const uint2* pp = p.postings + lo;
    const int len = (int)(hi - lo);
    for (int i0 = tid; i0 < len; i0 += kThreads * kUnroll) {
        uint2 pt[kUnroll];
        double fr[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int i = i0 + u * kThreads;
            pt[u] = (i < len) ? __ldg(pp + i) : make_uint2(0xFFFFFFFFu, 0u);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            fr[u] = (pt[u].x != 0xFFFFFFFFu) ? bm25_frac(p, pt[u].y, sdl[pt[u].x - d_lo]) : 0.0;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (pt[u].x != 0xFFFFFFFFu) {
                const int j = (int)(pt[u].x - d_lo);
                acc[j] = __dadd_rn(acc[j], __dmul_rn(idf, fr[u]));
            }
        }
    }
}

// kPreTerms x kPreDepth postings per thread are fetched into registers for ALL leading tokens at once,
// before the ordered accumulation starts: the CTA then pays one DRAM round trip for its postings instead
// of one per token (the accumulation itself only touches shared memory and the small impact table).
constexpr int kPreTerms = 4;
constexpr int kPreDepth = 4;

__global__ void __launch_bounds__(kThreads, 3) bm25_tile_kernel(const Bm25Params p) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    double* acc = reinterpret_cast<double*>(bm25_smem);                       // [kTileDocs] float64 partial scores
    uint32_t* sdl = reinterpret_cast<uint32_t*>(acc + kTileDocs);             // [kTileDocs] staged doc lengths
    __shared__ int64_t rng_lo[kTermGroup], rng_hi[kTermGroup];
    __shared__ double s_idf[kTermGroup];
    __shared__ float warp_max[kWarps];
    __shared__ int s_any;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // queries are the fastest grid index: the B CTAs of one doc tile run together and share its doc
    // lengths through L2
    const int b = blockIdx.x;
    const int tile = blockIdx.y;
    const int64_t d_lo = (int64_t)tile * kTileDocs;
    const int64_t d_hi = (d_lo + kTileDocs < p.n_docs) ? d_lo + kTileDocs : p.n_docs;
    const int ndoc = (int)(d_hi - d_lo);
    const int t_begin = p.q_off[b], t_end = p.q_off[b + 1];

    for (int j = tid; j < ndoc; j += kThreads) acc[j] = 0.0;
    bool dl_staged = false;

    for (int g0 = t_begin; g0 < t_end; g0 += kTermGroup) {
        const int gn = (t_end - g0 < kTermGroup) ? (t_end - g0) : kTermGroup;
        if (tid == 0) s_any = 0;
        __syncthreads();
        // ---- each token's posting slice for this doc range was located by bm25_ranges_kernel
        if (tid < gn) {
            const int64_t* r = p.ranges + (int64_t)(g0 + tid) * (p.n_tiles + 1) + tile;
            const int64_t lo = r[0], hi = r[1];
            rng_lo[tid] = lo;
            rng_hi[tid] = hi;
            s_idf[tid] = p.q_idf[g0 + tid];
            if (hi > lo) s_any = 1;
        }
        __syncthreads();
        if (!s_any) continue;
        // ---- prefetch: postings of the leading tokens (registers) + the tile's doc lengths (smem)
        uint2 pre[kPreTerms][kPreDepth];
#pragma unroll
        for (int t = 0; t < kPreTerms; ++t) {
            const uint2* pp = p.postings + rng_lo[t < gn ? t : 0];
            const int len = (t < gn) ? (int)(rng_hi[t] - rng_lo[t]) : 0;      // a slice never exceeds the tile size
#pragma unroll
            for (int u = 0; u < kPreDepth; ++u) {
                const int i = tid + u * kThreads;
                pre[t][u] = (i < len) ? __ldg(pp + i) : make_uint2(0xFFFFFFFFu, 0u);
            }
        }
        if (!dl_staged) {
            for (int j = tid; j < ndoc; j += kThreads) sdl[j] = p.dl[d_lo + j];
            dl_staged = true;
            __syncthreads();
        }
        // ---- accumulate token by token, in query order (bm25.py:99).  Leading tokens: compile-time
        // indexed register batches first, then whatever is left of a long slice streams in batches of kUnroll.
#pragma unroll
        for (int t = 0; t < kPreTerms; ++t) {
            if (t < gn) {                                           // block-uniform
                const double idf = s_idf[t];
                double fr[kPreDepth];
#pragma unroll
                for (int u = 0; u < kPreDepth; ++u)
                    fr[u] = (pre[t][u].x != 0xFFFFFFFFu) ? bm25_frac(p, pre[t][u].y, sdl[pre[t][u].x - (uint32_t)d_lo]) : 0.0;
#pragma unroll
                for (int u = 0; u < kPreDepth; ++u) {
                    if (pre[t][u].x != 0xFFFFFFFFu) {
                        const int j = (int)(pre[t][u].x - (uint32_t)d_lo);
                        acc[j] = __dadd_rn(acc[j], __dmul_rn(idf, fr[u]));
                    }
                }
                bm25_stream_slice(p, acc, sdl, rng_lo[t] + kThreads * kPreDepth, rng_hi[t], idf, (uint32_t)d_lo, tid);
                __syncthreads();
            }
        }
        for (int t = kPreTerms; t < gn; ++t) {
            bm25_stream_slice(p, acc, sdl, rng_lo[t], rng_hi[t], s_idf[t], (uint32_t)d_lo, tid);
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- epilogue: float64 -> float32 once, store, tile max
    float mx = 0.0f;
    float* out = p.scores + (int64_t)b * p.n_docs + d_lo;
    for (int j = tid; j < ndoc; j += kThreads) {
        const float s = __double2float_rn(acc[j]);
        out[j] = s;
        mx = fmaxf(mx, s);
    }
    if (p.stats != nullptr) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
        if (lane == 0) warp_max[warp] = mx;
        __syncthreads();
        if (tid == 0) {
            float v = warp_max[0];
            for (int w = 1; w < kWarps; ++w) v = fmaxf(v, warp_max[w]);
            atomicMax(&p.stats[b * 4 + HS_STAT_MAX_B], hs_enc_f32(v));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Batched tile kernel (default).  One CTA owns a doc tile for up to kMaxQpc queries of the batch, one after
// the other, which amortises the fixed costs of a tile over the queries: the slice bounds of ALL its query
// tokens are fetched in one round trip, the doc lengths are staged once (already multiplied into impact-table
// row offsets), and the float64 tile is re-zeroed by the epilogue that drains it.  The posting stream is
// software-pipelined through two register buffers of kDepth postings per thread: while chunk i is gathered
// and accumulated, the loads of chunk i+1 -- the next token's, or the next query's first token -- are in
// flight, so DRAM latency is paid once per CTA instead of once per token.  Arithmetic, accumulation order
// (query order per doc, one doc at most once per token) and rounding are the tile kernel's: identical bits.
constexpr int kBThreads = 256;                 // threads per group of the batched kernel
constexpr int kBWarps = kBThreads / 32;
constexpr int kDepth = 8;                      // postings per thread per chunk
constexpr int kChunkB = kBThreads * kDepth;     // 2048 postings
constexpr int kMaxQpc = 8;                     // queries per work item
constexpr int kTokWindow = 32;                 // query tokens whose slice bounds are staged at once (one per lane)

constexpr int kDenseFlag = 0x100;              // ChunkDesc.tok bit: the chunk is a slice of a hot term's dense vector
struct __align__(16) ChunkDesc {
    int64_t off;                               // first posting of the chunk (dense: first element of hot_c)
    int32_t len;                               // 1 .. kChunkB
    int16_t tok, bq;                           // token (window relative; | kDenseFlag | chunk index << 9 when dense)
};                                             // and query (CTA relative) it belongs to

struct BatchSmem {
    double acc[kTileDocs + 2];                 // + a dummy slot that absorbs the padding of partial chunks
    uint32_t row[kTileDocs + 4];               // what the impact lookup needs per doc (see frac_fast)
    ChunkDesc chunk[2 * kTokWindow];           // a slice holds <= kTileDocs = 2 * kChunkB postings
    double idf[kTokWindow];
    int qoff[kMaxQpc + 1];
    float wmax[kBWarps][kMaxQpc];
    int n_chunks;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// MODE 0: the fraction is computed per posting; 1: gathered from the impact table in global memory (L1);
// 2: gathered from a copy of the table in shared memory.  A gather of 32 unrelated table entries costs the L1 one
// tag lookup per cache line (32 cycles per warp instruction, which made mode 1 L1-bound at ~0.35 of the HBM
// peak) but only ~5 bank-conflict wavefronts in shared memory -- mode 2 is the default whenever the table fits.
// The kernel is persistent, one CTA per SM: kGroups independent 256-thread groups (named barriers), each with
// its own tile buffers, share the one table copy; each takes a contiguous run of (tile, query group) work items.
constexpr int kGroups = 3;

template <int MODE>
__global__ void __launch_bounds__(kBThreads* kGroups, 1)
    bm25_batch_kernel(const Bm25Params p, int B, int qpc, int n_qgroups, int table_rows) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    const int grp = threadIdx.x / kBThreads;
    const int tid = threadIdx.x - grp * kBThreads, lane = tid & 31, warp = tid >> 5;
    BatchSmem& sm = reinterpret_cast<BatchSmem*>(bm25_smem)[grp];
    double* s_table = reinterpret_cast<double*>(bm25_smem + kGroups * sizeof(BatchSmem));
    // row stride of the shared copy: odd, so that entries of equal tf in different rows fall into different banks
    const uint32_t width = p.tf_cap + 1;
    const uint32_t stride = MODE == 2 ? (width | 1u) : width;
    if (MODE == 2) {
        for (int i = threadIdx.x; i < table_rows * (int)width; i += kBThreads * kGroups) {
            const uint32_t r = (uint32_t)i / width, c = (uint32_t)i - r * width;
            s_table[r * stride + c] = p.impact_table[i];
        }
    }
    for (int j = tid; j < kTileDocs + 2; j += kBThreads) sm.acc[j] = 0.0;      // incl. the dummy slot
    __syncthreads();
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kBThreads) : "memory"); };
    const uint32_t tbl_s = (uint32_t)__cvta_generic_to_shared(s_table);
    const int64_t n_items = (int64_t)p.n_tiles * n_qgroups;

    // every group takes a contiguous run of items (tile-major): consecutive items mostly share the doc tile, so
    // its doc lengths are staged once per tile, and the runs differ by at most one item
    const int64_t n_workers = (int64_t)gridDim.x * kGroups, me = (int64_t)blockIdx.x * kGroups + grp;
    const int64_t item_lo = n_items * me / n_workers, item_hi = n_items * (me + 1) / n_workers;
    int staged_tile = -1;
    for (int64_t item = item_lo; item < item_hi;) {
        // merge the run's consecutive items of one tile into a single pass over up to kMaxQpc queries
        const int tile = (int)(item / n_qgroups);
        const int qg = (int)(item - (int64_t)tile * n_qgroups);
        int qg_end = (item_hi - item < (int64_t)(n_qgroups - qg)) ? qg + (int)(item_hi - item) : n_qgroups;
        if ((qg_end - qg) * qpc > kMaxQpc) qg_end = qg + kMaxQpc / qpc;
        const int b0 = qg * qpc;
        const int nb = (B < qg_end * qpc ? B : qg_end * qpc) - b0;
        item += qg_end - qg;
        const int64_t d_lo = (int64_t)tile * kTileDocs;
        const int ndoc = (int)((d_lo + kTileDocs < p.n_docs ? d_lo + kTileDocs : p.n_docs) - d_lo);
        // shared-memory addresses pre-biased by the tile's first doc id: slot of doc x = base + x * size (mod 2^32)
        const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(sm.acc) - (uint32_t)d_lo * 8u;
        const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(sm.row) - (uint32_t)d_lo * 4u;

        if (tile != staged_tile) {
            for (int j = tid; j < ndoc; j += kBThreads) {
                const uint32_t dl = p.dl[d_lo + j];
                sm.row[j] = MODE == 2 ? tbl_s + dl * stride * 8u : (MODE == 1 ? dl * stride : dl);
            }
            if (tid == 0) sm.row[kTileDocs] = MODE == 2 ? tbl_s : 0u;      // dummy doc: length 0
            staged_tile = tile;
        }
        if (tid <= nb) sm.qoff[tid] = p.q_off[b0 + tid];
        if (tid < kMaxQpc)
            for (int w = 0; w < kBWarps; ++w) sm.wmax[w][tid] = 0.0f;
        gsync();
        const int T0 = sm.qoff[0], T1 = sm.qoff[nb];

        int bq = 0;                                // query (relative to b0) whose partial scores are in the tile
        // drains the tile into the score row of query bq: float64 -> float32 once, coalesced store, re-zero, max
        auto epilogue = [&]() {
            float mx = 0.0f;
            if (p.scores_h != nullptr) {                     // binary16 screen (row stride even, tile start a multiple of 4096)
                __half* oh = p.scores_h + (int64_t)(b0 + bq) * p.ld_h + d_lo;
                if (ndoc == kTileDocs) {
                    double2* a2 = reinterpret_cast<double2*>(sm.acc) + tid;
                    __half2* o2 = reinterpret_cast<__half2*>(oh) + tid;
#pragma unroll
                    for (int r = 0; r < kTileDocs / (2 * kBThreads); ++r) {
                        const double2 a = a2[r * kBThreads];
                        a2[r * kBThreads] = make_double2(0.0, 0.0);
                        const float s0 = __double2float_rn(a.x), s1 = __double2float_rn(a.y);
                        o2[r * kBThreads] = __floats2half2_rn(s0, s1);
                        mx = fmaxf(mx, fmaxf(s0, s1));
                    }
                } else {
                    for (int j = tid; j < ndoc; j += kBThreads) {
                        const float sc = __double2float_rn(sm.acc[j]);
                        sm.acc[j] = 0.0;
                        oh[j] = __float2half_rn(sc);
                        mx = fmaxf(mx, sc);
                    }
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
                if (lane == 0) sm.wmax[warp][bq] = mx;
                gsync();
                return;
            }
            float* o = p.scores + (int64_t)(b0 + bq) * p.n_docs + d_lo;
            if (ndoc == kTileDocs && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {     // group-uniform: full tile
                double2* a2 = reinterpret_cast<double2*>(sm.acc) + tid;
                float2* o2 = reinterpret_cast<float2*>(o) + tid;
#pragma unroll
                for (int r = 0; r < kTileDocs / (2 * kBThreads); ++r) {
                    const double2 a = a2[r * kBThreads];
                    a2[r * kBThreads] = make_double2(0.0, 0.0);
                    const float s0 = __double2float_rn(a.x), s1 = __double2float_rn(a.y);
                    o2[r * kBThreads] = make_float2(s0, s1);
                    mx = fmaxf(mx, fmaxf(s0, s1));
                }
            } else {
                for (int j = tid; j < ndoc; j += kBThreads) {
                    const float sc = __double2float_rn(sm.acc[j]);
                    sm.acc[j] = 0.0;
                    o[j] = sc;
                    mx = fmaxf(mx, sc);
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
            if (lane == 0) sm.wmax[warp][bq] = mx;
            gsync();
        };
        // A chunk is loaded and consumed by one of three group-uniform code paths: full (kChunkB postings, no
        // predicates at all), short (<= kBThreads postings: one predicated slot) or general.
        auto load_chunk = [&](const ChunkDesc& cd, uint2 (&buf)[kDepth]) {
            const int64_t off = cd.off;
            const int len = cd.len;
            // a dense chunk is 8-byte elements too: the same register pipeline carries postings and contributions
            const uint2* pp = ((cd.tok & kDenseFlag) ? reinterpret_cast<const uint2*>(p.hot_c) : p.postings) + off + tid;
            if (len == kChunkB) {
#pragma unroll
                for (int u = 0; u < kDepth; ++u) buf[u] = __ldg(pp + u * kBThreads);
            } else if (len <= kBThreads) {
                if (tid < len) buf[0] = __ldg(pp);
            } else {
                const uint2 pad = make_uint2((uint32_t)d_lo + kTileDocs, 0u);           // dummy doc, tf = 0
#pragma unroll
                for (int u = 0; u < kDepth; ++u) buf[u] = (tid + u * kBThreads < len) ? __ldg(pp + u * kBThreads) : pad;
            }
        };
        // one thread asks the L2 for the chunk after next (the 16-byte aligned part of its byte range), so that the
        // register loads issued one chunk ahead find their lines on chip
        auto prefetch_chunk = [&](int ci) {
            if (tid == 0) {
                const ChunkDesc d = sm.chunk[ci];
                const uint2* base = (d.tok & kDenseFlag) ? reinterpret_cast<const uint2*>(p.hot_c) : p.postings;
                const uintptr_t lo = (reinterpret_cast<uintptr_t>(base + d.off) + 15) & ~(uintptr_t)15;
                const uintptr_t hi = reinterpret_cast<uintptr_t>(base + d.off + d.len) & ~(uintptr_t)15;
                if (hi > lo)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo))
                                 : "memory");
            }
        };
        // r: what sm.row holds for the doc (MODE 2: shared-memory address of its table row, MODE 1: row index
        // times the row stride, MODE 0: the doc length)
        auto frac_fast = [&](uint32_t r, uint32_t tf) -> double {                 // requires tf <= tf_cap
            return MODE == 2 ? lds_f64(r + tf * 8u) : __ldg(p.impact_table + (r + tf));
        };
        auto frac_any = [&](uint32_t r, uint32_t tf) -> double {
            if (MODE != 0 && tf <= p.tf_cap) return frac_fast(r, tf);
            const uint32_t dl = MODE == 2 ? (r - tbl_s) / (stride * 8u) : (MODE == 1 ? r / stride : r);
            return bm25_frac_compute(p.k1, p.one_minus_b, p.b, p.avgdl, p.k1p1, tf, dl);
        };
        auto consume = [&](int ci, const uint2 (&buf)[kDepth]) {
            const ChunkDesc d = sm.chunk[ci];
            while (bq < d.bq) {                                                  // earlier queries are complete
                epilogue();
                ++bq;
            }
            if (d.tok & kDenseFlag) {
                // hot term: buf holds float64 contributions of consecutive docs -- slot = chunk base + tid + u * 256,
                // conflict-free read-modify-write, no row lookup, no impact gather (x + 0.0 == x: docs without the
                // term keep their bits)
                const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sm.acc) + (uint32_t)(((d.tok >> 9) & 1) * kChunkB + tid) * 8u;
#pragma unroll
                for (int u = 0; u < kDepth; ++u) {
                    if (tid + u * kBThreads < d.len) {
                        const uint32_t a = a0 + (uint32_t)(u * kBThreads) * 8u;
                        const double c = __hiloint2double((int)buf[u].y, (int)buf[u].x);
                        sts_f64(a, __dadd_rn(lds_f64(a), c));
                    }
                }
                gsync();
                return;
            }
            const double idf = sm.idf[d.tok & 31];
            if (d.len > kBThreads) {
                // full chunks and padded partial ones share this path: no per-posting predicates; one slow-path
                // test per thread (the OR of the tfs bounds their maximum)
                uint32_t tf_or = 0;
#pragma unroll
                for (int u = 0; u < kDepth; ++u) tf_or |= buf[u].y;
                const bool fast = MODE != 0 && tf_or <= p.tf_cap;
#pragma unroll
                for (int h = 0; h < kDepth; h += 4) {
                    if (h * kBThreads < d.len) {                                     // group-uniform
                        double fr[4], cur[4];
                        uint32_t r[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) r[u] = lds_u32(row_s + buf[h + u].x * 4u);
                        if (fast) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) fr[u] = frac_fast(r[u], buf[h + u].y);
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u) fr[u] = frac_any(r[u], buf[h + u].y);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) cur[u] = lds_f64(acc_s + buf[h + u].x * 8u);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            sts_f64(acc_s + buf[h + u].x * 8u, __dadd_rn(cur[u], __dmul_rn(idf, fr[u])));
                    }
                }
            } else if (tid < d.len) {
                const double f = frac_any(lds_u32(row_s + buf[0].x * 4u), buf[0].y);
                const uint32_t a = acc_s + buf[0].x * 8u;
                sts_f64(a, __dadd_rn(lds_f64(a), __dmul_rn(idf, f)));
            }
            gsync();                                         // token order: these updates land before the next token's
        };

        for (int w0 = T0; w0 < T1; w0 += kTokWindow) {
            // ---- warp 0 turns the next <= 32 query tokens into a list of posting chunks for this doc tile
            if (warp == 0) {
                const int t = w0 + lane;
                int64_t lo = 0, hi = 0;
                int dense = 0;
                if (t < T1) {
                    const int64_t* r = p.ranges + (int64_t)t * (p.n_tiles + 1) + tile;
                    lo = r[0];
                    hi = r[1];
                    sm.idf[lane] = p.q_idf[t];
                    if (p.hot_of_term != nullptr && hi > lo) {
                        const int term = p.q_terms[t];
                        const int h = (term >= 0 && term < p.n_terms) ? __ldg(p.hot_of_term + term) : -1;
                        if (h >= 0) {                            // the tile's slice of the term's dense vector
                            dense = kDenseFlag;
                            lo = (int64_t)h * p.n_docs + d_lo;
                            hi = lo + ndoc;
                        }
                    }
                }
                const int len = (int)(hi - lo);                  // <= kTileDocs: a doc occurs once per posting list
                const int n = (len + kChunkB - 1) / kChunkB < 2 ? (len + kChunkB - 1) / kChunkB : 2;   // chunk[] holds 2 per token
                int incl = n;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, dlt);
                    if (lane >= dlt) incl += v;
                }
                int q = 0;
                for (int i = 1; i < nb; ++i) q += (t >= sm.qoff[i]);
                for (int c = 0; c < n; ++c) {
                    ChunkDesc d;
                    d.off = lo + (int64_t)c * kChunkB;
                    d.len = len - c * kChunkB < kChunkB ? len - c * kChunkB : kChunkB;
                    d.tok = (int16_t)(lane | dense | (dense ? (c << 9) : 0));
                    d.bq = (int16_t)q;
                    sm.chunk[incl - n + c] = d;
                }
                if (lane == 31) sm.n_chunks = incl;
            }
            gsync();
            const int nC = sm.n_chunks;
            uint2 A[kDepth], Bf[kDepth];
            if (nC > 0) load_chunk(sm.chunk[0], A);
            // L2 prefetch distance p.pf (chunks ahead of the one being consumed; HS_BM25_PF, default 2)
            const int kPf = p.pf;
            for (int c = 1; c < kPf + 1 && c < nC; ++c) prefetch_chunk(c);
            for (int i = 0; i < nC; i += 2) {
                if (i + kPf < nC) prefetch_chunk(i + kPf);
                if (i + 1 < nC) load_chunk(sm.chunk[i + 1], Bf);
                consume(i, A);
                if (i + 1 >= nC) break;
                if (i + kPf + 1 < nC) prefetch_chunk(i + kPf + 1);
                if (i + 2 < nC) load_chunk(sm.chunk[i + 2], A);
                consume(i + 1, Bf);
            }
            gsync();     // the window's chunk list and idf values may be overwritten
        }
        for (; bq < nb; ++bq) epilogue();                    // the last query, and queries without known tokens
        if (p.stats != nullptr && tid < nb) {
            float v = sm.wmax[0][tid];
            for (int w = 1; w < kBWarps; ++w) v = fmaxf(v, sm.wmax[w][tid]);
            atomicMax(&p.stats[(b0 + tid) * 4 + HS_STAT_MAX_B], hs_enc_f32(v));
        }
        gsync();         // wmax / qoff / row are rewritten by the next item
    }
}

// BM25Plus (bm25.py:150-179): every doc receives idf * (num / den + delta) for every known query token,
// tf = 0 included (num / den = 0 there), so the variant is dense.  Same tiling; per token the posting slice
// is first scattered into a second float64 tile (sentinel -1 = no posting), then ALL docs of the tile are
// updated in one ordered pass.  Requires k1 >= 0, 0 <= b <= 1 (checked on the host) so that den > 0 wherever
// k1 > 0 and the doc is not (b == 1, dl == 0).
__global__ void __launch_bounds__(kThreads, 2) bm25plus_tile_kernel(const Bm25Params p) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    double* acc = reinterpret_cast<double*>(bm25_smem);
    double* tmp = acc + kTileDocs;
    uint32_t* sdl = reinterpret_cast<uint32_t*>(tmp + kTileDocs);
    __shared__ float warp_max[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, tile = blockIdx.y;
    const int64_t d_lo = (int64_t)tile * kTileDocs;
    const int64_t d_hi = (d_lo + kTileDocs < p.n_docs) ? d_lo + kTileDocs : p.n_docs;
    const int ndoc = (int)(d_hi - d_lo);
    for (int j = tid; j < ndoc; j += kThreads) {
        acc[j] = 0.0;
        tmp[j] = -1.0;
        sdl[j] = p.dl[d_lo + j];
    }
    __syncthreads();
    for (int t = p.q_off[b]; t < p.q_off[b + 1]; ++t) {
        const int64_t* r = p.ranges + (int64_t)t * (p.n_tiles + 1) + tile;
        const int64_t lo = r[0], hi = r[1];
        const double idf = p.q_idf[t];
        for (int64_t i = lo + tid; i < hi; i += kThreads) {
            const uint2 pt = __ldg(&p.postings[i]);
            const int j = (int)(pt.x - (uint32_t)d_lo);
            tmp[j] = bm25_frac(p, pt.y, sdl[j]);
        }
        __syncthreads();
        for (int j = tid; j < ndoc; j += kThreads) {
            const double x = tmp[j];
            if (x >= 0.0) {
                acc[j] = __dadd_rn(acc[j], __dmul_rn(idf, __dadd_rn(x, p.delta)));
                tmp[j] = -1.0;
            } else if (p.k1 > 0.0 && !(p.b == 1.0 && sdl[j] == 0)) {     // den = k1 * (...) > 0 with tf = 0
                acc[j] = __dadd_rn(acc[j], __dmul_rn(idf, __dadd_rn(0.0, p.delta)));
            }
        }
        __syncthreads();
    }
    float mx = 0.0f;
    float* out = p.scores + (int64_t)b * p.n_docs + d_lo;
    for (int j = tid; j < ndoc; j += kThreads) {
        const float s = __double2float_rn(acc[j]);
        out[j] = s;
        mx = fmaxf(mx, s);
    }
    if (p.stats != nullptr) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
        if (lane == 0) warp_max[warp] = mx;
        __syncthreads();
        if (tid == 0) {
            float v = warp_max[0];
            for (int w = 1; w < kWarps; ++w) v = fmaxf(v, warp_max[w]);
            atomicMax(&p.stats[b * 4 + HS_STAT_MAX_B], hs_enc_f32(v));
        }
    }
}

// One warp per (query token, tile boundary): offset of the first posting with doc id >= boundary * kTileDocs.
// All searches of a batch run concurrently (one ~5-round latency chain in total) instead of serially
// at the head of every tile CTA.
__global__ void bm25_ranges_kernel(const Bm25Params p, int n_tokens) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t per = p.n_tiles + 1;
    if (w >= (int64_t)n_tokens * per) return;
    const int tok = (int)(w / per);
    const int j = (int)(w - (int64_t)tok * per);
    const int term = p.q_terms[tok];
    int64_t r = 0;
    if (term >= 0 && term < p.n_terms) {
        if (p.skip_hot && __ldg(p.hot_of_term + term) >= 0) {
            // a hot term: the batched kernel only tests "slice not empty" before it takes the dense vector's slice (a
            // tile without a posting of the term adds zeros: same bits) -- no search, offsets j, j + 1, ...
            if (lane == 0) p.ranges[w] = j;
            return;
        }
        const int64_t pl = p.indptr[term], ph = p.indptr[term + 1];
        if (j == 0) r = pl;
        else if (j == p.n_tiles) r = ph;
        else r = warp_lower_bound(p.postings, pl, ph, (uint32_t)((int64_t)j * kTileDocs), lane);
    }
    if (lane == 0) p.ranges[w] = r;
}

__global__ void impact_table_kernel(double avgdl, double k1, double one_minus_b, double b, double k1p1,
                                    uint32_t max_dl, uint32_t tf_cap, double* table) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t w = tf_cap + 1;
    if (i < (uint64_t)(max_dl + 1) * w)
        table[i] = bm25_frac_compute(k1, one_minus_b, b, avgdl, k1p1, (uint32_t)(i % w), (uint32_t)(i / w));
}

// hot_c[h, doc] = idf_h * frac(dl[doc], tf) for every posting of hot term h (the row was zeroed first): exactly the
// product the streaming path adds per posting (bm25.py:110), so both paths accumulate the same float64 values
__global__ void bm25_hot_build_kernel(const Bm25Params p, const int32_t* __restrict__ hot_terms,
                                      const double* __restrict__ hot_idf, double* __restrict__ hot_c) {
    const int h = blockIdx.y;
    const int term = hot_terms[h];
    if (term < 0 || term >= p.n_terms) return;
    const int64_t lo = p.indptr[term], hi = p.indptr[term + 1];
    const double idf = hot_idf[h];
    double* row = hot_c + (int64_t)h * p.n_docs;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 pt = __ldg(&p.postings[i]);
        if (pt.x < (uint64_t)p.n_docs) row[pt.x] = __dmul_rn(idf, bm25_frac(p, pt.y, p.dl[pt.x]));
    }
}

// BM25.score for selected docs (multi_stage stage 2, pipelines.py:485): one warp per (query, candidate);
// per token a 32-ary search of the posting list for the doc, float64 accumulation in query order.
// PLUS: BM25Plus.score (bm25.py:161-179) -- idf * (num / den + delta) for every known query token, tf = 0 included.
template <bool PLUS>
__global__ void bm25_docs_kernel(const Bm25Params p, const int64_t* __restrict__ doc_ids, int C,
                                 double* __restrict__ out, int B) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)B * C) return;
    const int b = (int)(w / C);
    const int64_t doc = doc_ids[w];
    double score = 0.0;
    if (doc >= 0 && doc < p.n_docs) {
        const uint32_t dl = p.dl[doc];
        for (int t = p.q_off[b]; t < p.q_off[b + 1]; ++t) {
            const int term = p.q_terms[t];
            if (term < 0 || term >= p.n_terms) continue;
            const int64_t pl = p.indptr[term], ph = p.indptr[term + 1];
            const int64_t pos = warp_lower_bound(p.postings, pl, ph, (uint32_t)doc, lane);
            bool hit = false;
            if (pos < ph) {
                const uint2 pt = __ldg(&p.postings[pos]);
                if (pt.x == (uint32_t)doc) {
                    hit = true;
                    const double frac = bm25_frac(p, pt.y, dl);
                    score = __dadd_rn(score, __dmul_rn(p.q_idf[t], PLUS ? __dadd_rn(frac, p.delta) : frac));
                }
            }
            // tf = 0: numerator 0, denominator k1 * (...) -- contributes idf * (0.0 + delta) where it is > 0
            if (PLUS && !hit && p.k1 > 0.0 && !(p.b == 1.0 && dl == 0))
                score = __dadd_rn(score, __dmul_rn(p.q_idf[t], __dadd_rn(0.0, p.delta)));
        }
    }
    if (lane == 0) out[w] = score;
}

int fill_params(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                Bm25Params& p, const char* who) {
    if (idx->indptr == nullptr || idx->dl == nullptr) {
        hs_set_error("%s: index has no CSR / doc stats (call hs_index_set_csr and hs_index_set_doc_stats)", who);
        return HS_ERR_STATE;
    }
    HS_REQUIRE(q_terms != nullptr && q_idf != nullptr && q_off != nullptr, "%s: null query arrays", who);
    p.indptr = idx->indptr;
    p.postings = idx->postings;
    p.dl = idx->dl;
    p.impact_table = idx->impact_table;
    p.max_dl = idx->max_dl;
    p.tf_cap = idx->tf_cap;
    p.k1 = idx->k1;
    p.b = idx->b;
    p.one_minus_b = 1 - idx->b;          // python: 1 - self.b
    p.k1p1 = idx->k1 + 1;                // python: self.k1 + 1
    p.avgdl = idx->avgdl;
    p.q_terms = q_terms;
    p.q_idf = q_idf;
    p.q_off = q_off;
    p.n_docs = idx->n_docs;
    p.n_terms = idx->n_terms;
    p.scores = nullptr;
    p.scores_h = nullptr;
    p.ld_h = 0;
    static const int pf_env = getenv("HS_BM25_PF") != nullptr ? atoi(getenv("HS_BM25_PF")) : 2;
    p.pf = pf_env < 1 ? 1 : (pf_env > 16 ? 16 : pf_env);
    p.skip_hot = 0;
    p.stats = nullptr;
    p.ranges = nullptr;
    p.n_tiles = (int)((idx->n_docs + kTileDocs - 1) / kTileDocs);
    p.delta = 0.0;
    p.hot_c = idx->hot_c;
    p.hot_of_term = idx->hot_c != nullptr ? idx->hot_of_term : nullptr;
    return HS_OK;
}

// HS_BM25_IMPL=tile selects the one-query-per-CTA tile kernel (A/B measurements); both give the same bits.
// 0 compute, 1 table in global memory, 2 table in shared memory (HS_BM25_TABLE=global forces 1 for A/B runs)
int bm25_table_mode(bool have_table, size_t smem_with_table) {
    static const bool force_global = [] {
        const char* e = getenv("HS_BM25_TABLE");
        return e != nullptr && strcmp(e, "global") == 0;
    }();
    if (!have_table) return 0;
    return (!force_global && smem_with_table <= 227 * 1024) ? 2 : 1;
}

bool bm25_use_batch() {
    static const bool tile = [] {
        const char* e = getenv("HS_BM25_IMPL");
        return e != nullptr && strcmp(e, "tile") == 0;
    }();
    return !tile;
}

}  // namespace

extern "C" {

int hs_bm25_impact_table(double avgdl, double k1, double b, uint32_t max_dl, uint32_t tf_cap, double* table,
                         void* stream) {
    HS_REQUIRE(table != nullptr, "hs_bm25_impact_table: table is null");
    const uint64_t n = (uint64_t)(max_dl + 1) * (tf_cap + 1);
    HS_REQUIRE(n <= (1ull << 31), "hs_bm25_impact_table: table too large");
    impact_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(avgdl, k1, 1 - b, b, k1 + 1,
                                                                                       max_dl, tf_cap, table);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_bm25_build_hot(hs_index* idx, const int32_t* hot_terms, const double* hot_idf, int32_t n_hot,
                      const int32_t* hot_of_term, double* hot_c, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_bm25_build_hot: idx is null");
    idx->hot_c = nullptr;
    idx->hot_of_term = nullptr;
    idx->n_hot = 0;
    if (n_hot == 0 || idx->n_docs == 0) return HS_OK;
    HS_REQUIRE(n_hot > 0 && n_hot <= 32767 / 2 && hot_terms != nullptr && hot_idf != nullptr && hot_of_term != nullptr &&
                   hot_c != nullptr, "hs_bm25_build_hot: bad arguments");
    HS_REQUIRE(((uintptr_t)hot_c & 15) == 0, "hs_bm25_build_hot: hot_c must be 16-byte aligned");
    static const int32_t dummy_q[1] = {0};
    Bm25Params p;
    int rc = fill_params(idx, dummy_q, (const double*)dummy_q, dummy_q, p, "hs_bm25_build_hot");
    if (rc != HS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    HS_CUDA(cudaMemsetAsync(hot_c, 0, (size_t)n_hot * idx->n_docs * sizeof(double), st));
    dim3 grid(1024, (unsigned)n_hot);
    bm25_hot_build_kernel<<<grid, 256, 0, st>>>(p, hot_terms, hot_idf, hot_c);
    HS_LAUNCH_CHECK();
    idx->hot_c = hot_c;
    idx->hot_of_term = hot_of_term;
    idx->n_hot = n_hot;
    return HS_OK;
}

size_t hs_bm25_workspace_bytes(int64_t n_docs, int32_t n_tokens) {
    if (n_docs <= 0 || n_tokens <= 0) return 0;
    return (size_t)n_tokens * (size_t)((n_docs + kTileDocs - 1) / kTileDocs + 1) * sizeof(int64_t);
}

static int bm25_score_impl(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                           int32_t B, int32_t n_tokens, void* workspace, size_t workspace_bytes, float* scores,
                           uint32_t* stats_enc, void* stream, bool plus, double delta, uint16_t* scores_f16 = nullptr,
                           int64_t ld_f16 = 0) {
    HS_REQUIRE(idx != nullptr, "hs_bm25_score: idx is null");
    if (idx->n_docs == 0 || B == 0) return HS_OK;
    HS_REQUIRE(B > 0 && B <= 65535 && (scores != nullptr || scores_f16 != nullptr) && n_tokens >= 0,
               "hs_bm25_score: bad arguments (B=%d)", B);
    HS_REQUIRE(scores_f16 == nullptr || (!plus && bm25_use_batch() && ld_f16 >= idx->n_docs && (ld_f16 & 7) == 0 &&
                                         ((uintptr_t)scores_f16 & 15) == 0),
               "hs_bm25_score_f16: needs the batched kernel, a 16-byte aligned array and a row stride >= n_docs that is a "
               "multiple of 8");
    Bm25Params p;
    int rc = fill_params(idx, q_terms, q_idf, q_off, p, "hs_bm25_score");
    if (rc != HS_OK) return rc;
    p.scores = scores;
    p.scores_h = (__half*)scores_f16;
    p.ld_h = ld_f16;
    p.stats = stats_enc;
    if (n_tokens > 0) {
        HS_REQUIRE(workspace != nullptr && workspace_bytes >= hs_bm25_workspace_bytes(idx->n_docs, n_tokens),
                   "hs_bm25_score: workspace too small");
        p.ranges = (int64_t*)workspace;
        p.skip_hot = (!plus && bm25_use_batch() && p.hot_of_term != nullptr) ? 1 : 0;
        const int64_t warps = (int64_t)n_tokens * (p.n_tiles + 1);
        bm25_ranges_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n_tokens);
        HS_LAUNCH_CHECK();
    }
    HS_REQUIRE(p.n_tiles <= 65535, "hs_bm25_score: shard has more than 65535 doc tiles (%d)", p.n_tiles);
    dim3 grid((unsigned)B, (unsigned)p.n_tiles);
    if (plus) {
        HS_REQUIRE(idx->k1 >= 0.0 && idx->b >= 0.0 && idx->b <= 1.0 && idx->avgdl > 0.0,
                   "hs_bm25plus_score: needs k1 >= 0, 0 <= b <= 1 and a non-empty corpus");
        p.delta = delta;
        const size_t smem = (size_t)kTileDocs * (2 * sizeof(double) + sizeof(uint32_t));
        static size_t smem_set[16] = {0};
        HS_CUDA(hs_smem_limit(bm25plus_tile_kernel, smem, smem_set));
        bm25plus_tile_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    } else if (bm25_use_batch()) {
        // queries per work item: as many as possible (<= kMaxQpc) while there are a few items per group
        const int sms = idx->num_sms;
        int qpc = B < kMaxQpc ? B : kMaxQpc;
        while (qpc > 1 && (int64_t)p.n_tiles * ((B + qpc - 1) / qpc) < (int64_t)16 * kGroups * sms) qpc = qpc / 2;
        const int nqg = (B + qpc - 1) / qpc;
        const int64_t n_items = (int64_t)p.n_tiles * nqg;
        const int64_t rows = p.impact_table != nullptr ? (int64_t)p.max_dl + 1 : 0;
        const size_t tbl_bytes = (size_t)rows * ((p.tf_cap + 1) | 1u) * sizeof(double);
        const size_t base = kGroups * sizeof(BatchSmem);
        const int mode = bm25_table_mode(p.impact_table != nullptr, base + tbl_bytes);
        const size_t smem = base + (mode == 2 ? tbl_bytes : 0);
        const unsigned grid = (unsigned)((n_items + kGroups - 1) / kGroups < sms ? (n_items + kGroups - 1) / kGroups : sms);
        static size_t smem_set[3][16] = {{0}};
        auto launch = [&](auto kern) -> int {
            HS_CUDA(hs_smem_limit(kern, smem, smem_set[mode]));
            kern<<<grid, kBThreads * kGroups, smem, (cudaStream_t)stream>>>(p, B, qpc, nqg, (int)rows);
            return HS_OK;
        };
        rc = mode == 2 ? launch(bm25_batch_kernel<2>) : (mode == 1 ? launch(bm25_batch_kernel<1>) : launch(bm25_batch_kernel<0>));
        if (rc != HS_OK) return rc;
    } else {
        const size_t smem = (size_t)kTileDocs * (sizeof(double) + sizeof(uint32_t));
        static size_t smem_set[16] = {0};
        HS_CUDA(hs_smem_limit(bm25_tile_kernel, smem, smem_set));
        bm25_tile_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    }
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_bm25_score(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                  int32_t B, int32_t n_tokens, void* workspace, size_t workspace_bytes, float* scores,
                  uint32_t* stats_enc, void* stream) {
    return bm25_score_impl(idx, q_terms, q_idf, q_off, B, n_tokens, workspace, workspace_bytes, scores, stats_enc, stream,
                           false, 0.0);
}

int hs_bm25_score_f16(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                      int32_t B, int32_t n_tokens, void* workspace, size_t workspace_bytes, uint16_t* scores_f16,
                      int64_t ld, uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(scores_f16 != nullptr, "hs_bm25_score_f16: scores_f16 is null");
    return bm25_score_impl(idx, q_terms, q_idf, q_off, B, n_tokens, workspace, workspace_bytes, nullptr, stats_enc, stream,
                           false, 0.0, scores_f16, ld);
}

int hs_bm25plus_score(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                      int32_t B, int32_t n_tokens, double delta, void* workspace, size_t workspace_bytes,
                      float* scores, uint32_t* stats_enc, void* stream) {
    return bm25_score_impl(idx, q_terms, q_idf, q_off, B, n_tokens, workspace, workspace_bytes, scores, stats_enc, stream,
                           true, delta);
}

static int bm25_score_docs_impl(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                                int32_t B, const int64_t* doc_ids, int32_t C, double* out, void* stream, bool plus,
                                double delta) {
    HS_REQUIRE(idx != nullptr, "hs_bm25_score_docs: idx is null");
    if (B == 0 || C == 0) return HS_OK;
    HS_REQUIRE(B > 0 && C > 0 && doc_ids != nullptr && out != nullptr, "hs_bm25_score_docs: bad arguments");
    Bm25Params p;
    int rc = fill_params(idx, q_terms, q_idf, q_off, p, "hs_bm25_score_docs");
    if (rc != HS_OK) return rc;
    const int64_t warps = (int64_t)B * C;
    const int64_t blocks = (warps * 32 + 255) / 256;
    if (plus) {
        HS_REQUIRE(idx->k1 >= 0.0 && idx->b >= 0.0 && idx->b <= 1.0 && idx->avgdl > 0.0,
                   "hs_bm25plus_score_docs: needs k1 >= 0, 0 <= b <= 1 and a non-empty corpus");
        p.delta = delta;
        bm25_docs_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, doc_ids, C, out, B);
    } else {
        bm25_docs_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, doc_ids, C, out, B);
    }
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_bm25_score_docs(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                       int32_t B, const int64_t* doc_ids, int32_t C, double* out, void* stream) {
    return bm25_score_docs_impl(idx, q_terms, q_idf, q_off, B, doc_ids, C, out, stream, false, 0.0);
}

int hs_bm25plus_score_docs(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                           int32_t B, const int64_t* doc_ids, int32_t C, double delta, double* out, void* stream) {
    return bm25_score_docs_impl(idx, q_terms, q_idf, q_off, B, doc_ids, C, out, stream, true, delta);
}

}  // extern "C"
