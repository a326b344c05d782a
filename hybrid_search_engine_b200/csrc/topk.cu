// K3 + K4 -- min-max normalisation and weighted fusion fused into the top-k select, plus the list
// merge used for the final per-query result and for the multi-GPU merge (C1).
//
// Replaces: normalize_scores (utils.py:57-71) x2 + weighted sum (core.py:264-268), hybrid_bm25's
// max-normalise + fuse loop (pipelines.py:331-340), and every "full sort then slice"
// (core.py:271, bm25.py:141, utils.py:86, pipelines.py:342-343).
//
// The fused score of doc i is computed on the fly from a[b,i] / b[b,i] and the per-query stats with
// exactly the reference's float32/float64 rounding steps (IEEE _rn intrinsics, no contraction) and is
// never written to memory: it goes straight into a 64-bit ranking key
//     key = ordered_u32(score) << 32 | (0xFFFFFFFF - doc_id)
// whose unsigned order is the canonical total order (score desc, doc_id asc) -- the reference's own
// order wherever the reference is deterministic (stable sort over ascending i, pipelines.py:342).
//
// Select: every CTA streams a contiguous doc range and keeps a shared-memory candidate buffer of
// 2*KP keys (KP = pow2 >= k).  Keys above the running threshold are appended with one shared atomic;
// when the buffer fills it is bitonic-sorted and cut back to the best KP, which raises the threshold.
// For random input the number of appended keys is O(k log(n/k)), so the pass is a pure read stream.
// Per-CTA winners go to the workspace; a single-CTA-per-query merge kernel applies the same filter
// to the candidate lists.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 4;
constexpr int kChunkDocs = 4096;
constexpr int kMaxChunks = 296;

template <int CAP>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* a) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= CAP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < CAP / 2; i += kThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t x = a[lo], y = a[hi];
                if ((x < y) == desc) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}

// Streaming top-k state in shared memory.
template <int KP>
struct Selector {
    static constexpr int CAP = 2 * KP;
    uint64_t buf[CAP];
    uint64_t thr;   // keys <= thr cannot be in the top k any more
    int cnt;        // next free slot (>= KP: slots [0, KP) hold the current best)

    __device__ void init() {
        for (int i = threadIdx.x; i < CAP; i += kThreads) buf[i] = 0;
        if (threadIdx.x == 0) {
            thr = 0;
            cnt = KP;
        }
        __syncthreads();
    }
    // all threads call with their kItems keys (0 = nothing).  Collective.
    __device__ void push(uint64_t (&key)[kItems], int k) {
        unsigned pend = 0;
#pragma unroll
        for (int j = 0; j < kItems; ++j)
            if (key[j] != 0) pend |= 1u << j;
        while (true) {
            const uint64_t t = thr;
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                if (pend & (1u << j)) {
                    if (key[j] > t) {
                        const int pos = atomicAdd(&cnt, 1);
                        if (pos < CAP) {
                            buf[pos] = key[j];
                            pend &= ~(1u << j);
                        }
                    } else {
                        pend &= ~(1u << j);
                    }
                }
            }
            // a key stays pending only if the buffer overflowed under it, so "someone is pending"
            // is the (block-uniform) prune condition
            const int more = __syncthreads_or(pend != 0);
            if (!more) break;
            prune(k);
        }
    }
    // collective: sort, keep the best KP, raise the threshold to the k-th best
    __device__ void prune(int k) {
        bitonic_sort_desc<CAP>(buf);
        if (threadIdx.x == 0) {
            thr = buf[k - 1];
            cnt = KP;
        }
        __syncthreads();
    }
    __device__ void finish(int k) { prune(k); }
};

struct FuseParams {
    const float* a;
    const float* b;
    const uint32_t* stats;
    const uint64_t* below;
    uint64_t* cand;       // [B, n_chunks, k]
    int64_t n, doc_base;
    int mode, k, n_chunks;
    float wa32, wb32;
    double wa64;
};

struct FuseConsts {
    float min_a, range_a, max_b_div, min_b, range_b;
    bool const_a, const_b;
};

__device__ __forceinline__ FuseConsts load_consts(const FuseParams& p, int b) {
    FuseConsts c;
    c.min_a = c.range_a = c.max_b_div = c.min_b = c.range_b = 0.f;
    c.const_a = c.const_b = false;
    if (p.mode == HS_FUSE_RAW) return c;
    const uint32_t* s = p.stats + b * 4;
    const float mn = hs_dec_f32(s[HS_STAT_MIN_A]), mx = hs_dec_f32(s[HS_STAT_MAX_A]);
    c.min_a = mn;
    c.range_a = __fsub_rn(mx, mn);                 // utils.py:69  max_score - min_score
    c.const_a = (c.range_a == 0.0f);
    if (p.mode == HS_FUSE_HYBRID_BM25) {
        const float mb = hs_dec_f32(s[HS_STAT_MAX_B]);
        c.max_b_div = (mb > 0.0f) ? mb : 1.0f;     // pipelines.py:332
    } else if (p.b != nullptr) {
        const float mnb = hs_dec_f32(s[HS_STAT_MIN_B]), mxb = hs_dec_f32(s[HS_STAT_MAX_B]);
        c.min_b = mnb;
        c.range_b = __fsub_rn(mxb, mnb);
        c.const_b = (c.range_b == 0.0f);
    }
    return c;
}

__device__ __forceinline__ float fuse_score(const FuseParams& p, const FuseConsts& c, float a, float b) {
    if (p.mode == HS_FUSE_RAW) return a;
    // utils.py:69-71: constant vector -> ones, else (x - min) / (max - min)
    const float an = c.const_a ? 1.0f : __fdiv_rn(__fsub_rn(a, c.min_a), c.range_a);
    if (p.mode == HS_FUSE_HYBRID_BM25) {
        // pipelines.py:337-339: f32(f64(sem)/1.0 * ws) + f32(bm/max_bm) * f32(wb)
        const float t1 = __double2float_rn(__dmul_rn((double)an, p.wa64));
        const float t2 = __fmul_rn(__fdiv_rn(b, c.max_b_div), p.wb32);
        return __fadd_rn(t1, t2);
    }
    // core.py:268: (sem * sw) + (lex * lw) in float32
    const float t1 = __fmul_rn(an, p.wa32);
    float t2 = 0.0f;
    if (p.b != nullptr) {
        const float bn = c.const_b ? 1.0f : __fdiv_rn(__fsub_rn(b, c.min_b), c.range_b);
        t2 = __fmul_rn(bn, p.wb32);
    }
    return __fadd_rn(t1, t2);
}

template <int KP>
__global__ void __launch_bounds__(kThreads) fuse_topk_kernel(const FuseParams p) {
    __shared__ Selector<KP> sel;
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int64_t span = ((p.n + p.n_chunks - 1) / p.n_chunks + kThreads * kItems - 1) / (kThreads * kItems) *
                         (kThreads * kItems);
    const int64_t lo = (int64_t)chunk * span;
    const int64_t hi = (lo + span < p.n) ? lo + span : p.n;
    const FuseConsts c = load_consts(p, b);
    const uint64_t below = p.below ? p.below[b] : ~0ull;
    const float* pa = p.a + (int64_t)b * p.n;
    const float* pb = p.b ? p.b + (int64_t)b * p.n : nullptr;
    sel.init();
    for (int64_t base = lo; base < hi; base += kThreads * kItems) {
        uint64_t key[kItems];
        float av[kItems], bv[kItems];
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const int64_t i = base + j * kThreads + tid;
            av[j] = (i < hi) ? __ldg(pa + i) : 0.f;
            bv[j] = (pb != nullptr && i < hi) ? __ldg(pb + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const int64_t i = base + j * kThreads + tid;
            key[j] = 0;
            if (i < hi) {
                const uint64_t kk = hs_make_key(fuse_score(p, c, av[j], bv[j]), (uint32_t)(p.doc_base + i));
                if (kk < below) key[j] = kk;
            }
        }
        sel.push(key, p.k);
    }
    sel.finish(p.k);
    uint64_t* out = p.cand + ((int64_t)b * p.n_chunks + chunk) * p.k;
    for (int i = tid; i < p.k; i += kThreads) out[i] = sel.buf[i];
}

// keys: [n_lists, B, k] -> out [B, k]; one CTA per query
template <int KP>
__global__ void __launch_bounds__(kThreads) topk_merge_kernel(const uint64_t* __restrict__ keys, int n_lists,
                                                              int B, int k, int64_t list_stride,
                                                              int64_t query_stride, uint64_t* __restrict__ out) {
    __shared__ Selector<KP> sel;
    const int b = blockIdx.x, tid = threadIdx.x;
    sel.init();
    const int64_t total = (int64_t)n_lists * k;
    for (int64_t base = 0; base < total; base += kThreads * kItems) {
        uint64_t key[kItems];
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const int64_t i = base + j * kThreads + tid;
            key[j] = 0;
            if (i < total) {
                const int64_t l = i / k, r = i - l * k;
                key[j] = keys[l * list_stride + (int64_t)b * query_stride + r];
            }
        }
        sel.push(key, k);
    }
    sel.finish(k);
    for (int i = tid; i < k; i += kThreads) out[(int64_t)b * k + i] = sel.buf[i];
}

int n_chunks_for(int64_t n) {
    int64_t c = (n + kChunkDocs - 1) / kChunkDocs;
    if (c < 1) c = 1;
    if (c > kMaxChunks) c = kMaxChunks;
    return (int)c;
}

template <int KP>
int launch_merge(const uint64_t* keys, int n_lists, int B, int k, int64_t list_stride, int64_t query_stride,
                 uint64_t* out, cudaStream_t st) {
    topk_merge_kernel<KP><<<B, kThreads, 0, st>>>(keys, n_lists, B, k, list_stride, query_stride, out);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int merge_dispatch(const uint64_t* keys, int n_lists, int B, int k, int64_t list_stride, int64_t query_stride,
                   uint64_t* out, cudaStream_t st) {
    if (k <= 128) return launch_merge<128>(keys, n_lists, B, k, list_stride, query_stride, out, st);
    if (k <= 512) return launch_merge<512>(keys, n_lists, B, k, list_stride, query_stride, out, st);
    return launch_merge<2048>(keys, n_lists, B, k, list_stride, query_stride, out, st);
}

}  // namespace

extern "C" {

size_t hs_fuse_topk_workspace_bytes(int64_t n_docs, int32_t B, int32_t k) {
    if (n_docs <= 0 || B <= 0 || k <= 0) return 0;
    return (size_t)B * n_chunks_for(n_docs) * (size_t)k * sizeof(uint64_t);
}

int hs_fuse_topk(const hs_index* idx, int32_t fuse_mode, const float* a, const float* b,
                 const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k,
                 const uint64_t* below_key, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                 void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_fuse_topk: idx is null");
    HS_REQUIRE(B > 0 && B <= 65535 && k > 0 && k <= HS_TOPK_MAX, "hs_fuse_topk: B=%d k=%d out of range (k <= %d)", B,
               k, HS_TOPK_MAX);
    HS_REQUIRE(out_keys != nullptr, "hs_fuse_topk: out_keys is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (idx->n_docs == 0) {
        HS_CUDA(cudaMemsetAsync(out_keys, 0, (size_t)B * k * sizeof(uint64_t), st));
        return HS_OK;
    }
    HS_REQUIRE(a != nullptr, "hs_fuse_topk: a is null");
    HS_REQUIRE(fuse_mode == HS_FUSE_RAW || fuse_mode == HS_FUSE_SEARCHER || fuse_mode == HS_FUSE_HYBRID_BM25,
               "hs_fuse_topk: unknown fuse_mode %d", fuse_mode);
    HS_REQUIRE(fuse_mode == HS_FUSE_RAW || stats_enc != nullptr, "hs_fuse_topk: stats_enc is null");
    HS_REQUIRE(fuse_mode != HS_FUSE_HYBRID_BM25 || b != nullptr, "hs_fuse_topk: hybrid_bm25 needs b");
    HS_REQUIRE(fuse_mode != HS_FUSE_SEARCHER || b != nullptr || w_b == 0.0,
               "hs_fuse_topk: searcher fusion with w_b != 0 needs b");
    const size_t need = hs_fuse_topk_workspace_bytes(idx->n_docs, B, k);
    HS_REQUIRE(workspace != nullptr && workspace_bytes >= need, "hs_fuse_topk: workspace too small (%zu < %zu)",
               workspace_bytes, need);
    FuseParams p;
    p.a = a;
    p.b = b;
    p.stats = stats_enc;
    p.below = below_key;
    p.cand = (uint64_t*)workspace;
    p.n = idx->n_docs;
    p.doc_base = idx->doc_base;
    p.mode = fuse_mode;
    p.k = k;
    p.n_chunks = n_chunks_for(idx->n_docs);
    p.wa32 = (float)w_a;   // numpy: float32 array * python float -> float32(w)
    p.wb32 = (float)w_b;
    p.wa64 = w_a;
    dim3 grid((unsigned)p.n_chunks, (unsigned)B);
    if (k <= 128)
        fuse_topk_kernel<128><<<grid, kThreads, 0, st>>>(p);
    else if (k <= 512)
        fuse_topk_kernel<512><<<grid, kThreads, 0, st>>>(p);
    else
        fuse_topk_kernel<2048><<<grid, kThreads, 0, st>>>(p);
    HS_LAUNCH_CHECK();
    // candidate layout [B, n_chunks, k]: list stride k, query stride n_chunks * k
    return merge_dispatch(p.cand, p.n_chunks, B, k, (int64_t)k, (int64_t)p.n_chunks * k, out_keys, st);
}

int hs_topk_merge(const uint64_t* keys, int32_t n_lists, int32_t B, int32_t k, uint64_t* out_keys, void* stream) {
    HS_REQUIRE(keys != nullptr && out_keys != nullptr, "hs_topk_merge: null pointer");
    HS_REQUIRE(n_lists > 0 && B > 0 && k > 0 && k <= HS_TOPK_MAX, "hs_topk_merge: bad sizes");
    // layout [n_lists, B, k]: list stride B * k, query stride k
    return merge_dispatch(keys, n_lists, B, k, (int64_t)B * k, (int64_t)k, out_keys, (cudaStream_t)stream);
}

}  // extern "C"
