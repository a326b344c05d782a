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
#include <stdlib.h>

#include "common.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

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

// same network over the first n (power of two) entries only
__device__ __forceinline__ void bitonic_sort_desc_n(uint64_t* a, int n) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < n / 2; i += kThreads) {
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
//
// `shared_thr` (optional, global memory, one per query) lets all CTAs working on the same query pool
// their knowledge: whenever a CTA has k real keys, its k-th best is a lower bound on the GLOBAL k-th
// best, so it is published with atomicMax and every CTA filters with max(own, published).  CTAs that
// start late then append almost nothing; their output is "every key of my range above the published
// bound", still a superset of (global top k) restricted to the range, which is all the merge needs.
template <int KP>
struct Selector {
    static constexpr int CAP = (KP <= 128) ? 8 * KP : (KP <= 256 ? 4 * KP : 2 * KP);
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
    // after init(), for a pass that only appends (no prune): no slots reserved for a "current best", every slot beyond
    // cnt stays zero.  Collective.
    __device__ void start_empty() {
        if (threadIdx.x == 0) cnt = 0;
        __syncthreads();
    }
    __device__ __forceinline__ uint64_t bound(const unsigned long long* shared_thr) const {
        uint64_t t = thr;
        if (shared_thr != nullptr) {
            const uint64_t g = *reinterpret_cast<const volatile unsigned long long*>(shared_thr);
            if (g > t) t = g;
        }
        return t;
    }
    // all threads call with their kItems keys (0 = nothing).  Collective.
    __device__ void push(uint64_t (&key)[kItems], int k, unsigned long long* shared_thr = nullptr) {
        const int lane = threadIdx.x & 31;
        unsigned pend = 0;
#pragma unroll
        for (int j = 0; j < kItems; ++j)
            if (key[j] != 0) pend |= 1u << j;
        while (true) {
            const uint64_t t = bound(shared_thr);
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const bool live = (pend >> j) & 1u;
                const bool want = live && key[j] > t;
                if (live && !want) pend &= ~(1u << j);
                // warp-aggregated append: one shared atomic per warp and item slot
                const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
                if (m != 0) {
                    const int leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(&cnt, __popc(m));
                    base = __shfl_sync(0xFFFFFFFFu, base, leader);
                    if (want) {
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < CAP) {
                            buf[pos] = key[j];
                            pend &= ~(1u << j);
                        }
                    }
                }
            }
            // a key stays pending only if the buffer overflowed under it, so "someone is pending"
            // is the (block-uniform) prune condition
            const int more = __syncthreads_or(pend != 0);
            if (!more) break;
            prune(k, shared_thr);
        }
    }
    // collective: sort, keep the best KP, raise the threshold to the k-th best
    __device__ void prune(int k, unsigned long long* shared_thr = nullptr) {
        bitonic_sort_desc<CAP>(buf);
        if (threadIdx.x == 0) {
            const uint64_t kth = buf[k - 1];
            thr = kth;
            cnt = KP;
            if (shared_thr != nullptr && kth != 0) atomicMax(shared_thr, (unsigned long long)kth);
        }
        __syncthreads();
    }
    // final ordering: only slots [0, cnt) can hold keys that matter (slots beyond were discarded by an
    // earlier prune), so sort the smallest power of two covering them
    // (A CTA that never pruned may hold just a handful of keys -- start_empty() -- and sorts 32 slots instead of 2 KP:
    // with a starting bound the final sort, not the scan, was most of what the select kernel executed.)
    __device__ void finish(int k) {
        int used = cnt < CAP ? cnt : CAP;
        int n = 32;
        while (n < used) n <<= 1;
        for (int i = used + threadIdx.x; i < n; i += kThreads) buf[i] = 0;   // stale keys below the cut
        __syncthreads();
        bitonic_sort_desc_n(buf, n);
    }
};

struct FuseParams {
    const float* a;
    const __half* a16;    // the a array as IEEE binary16 (the verified mode's SCREEN scores) when a == nullptr
    const float* b;
    const __half* b16;    // the b array as binary16 [B, ld_a] (hs_bm25_score_f16) when b == nullptr and a16 is set
    const uint32_t* stats;
    const uint64_t* below;
    uint64_t* cand;       // [B, n_chunks, k]
    unsigned long long* gthr;   // [B] published lower bounds on the k-th best key (zeroed per call)
    uint64_t* blockmax;         // [B, kBoundBlocks] best key of each sampled block (workspace)
    int64_t n, doc_base;
    int64_t ld;                 // row stride of b (elements)
    int64_t ld_a;               // row stride of a / a16 (elements)
    int mode, k, n_chunks;
    int n_bound;                // blocks sampled for the starting bound (0 = none)
    float wa32, wb32;
    double wa64;
};

// One row of the a / b array: float32, or binary16 for the screen scores of the verified mode (half the bytes of the
// producer's write and of this pass's read; the rounding is covered by the verification's eps).  load<KD> fetches KD
// consecutive elements with one 8- or 16-byte instruction.
__device__ __forceinline__ void hs_unpack_h2(uint32_t w, float& x, float& y) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    x = f.x;
    y = f.y;
}
template <bool H>
struct FRow;
template <>
struct FRow<false> {
    const float* p;
    __device__ __forceinline__ FRow(const float* base, const __half*, int64_t ld, int b)
        : p(base != nullptr ? base + (int64_t)b * ld : nullptr) {}
    __device__ __forceinline__ bool present() const { return p != nullptr; }
    __device__ __forceinline__ float at(int64_t i) const { return __ldg(p + i); }
    template <int KD>
    __device__ __forceinline__ void load(int64_t i, float (&o)[KD]) const {
        static_assert(KD == 4, "float32 rows are read four at a time");
        if (p == nullptr) {
            o[0] = o[1] = o[2] = o[3] = 0.f;
            return;
        }
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
        o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
    }
    template <int KD>
    __device__ __forceinline__ bool aligned(int64_t i) const {
        return p == nullptr || (reinterpret_cast<uintptr_t>(p + i) & 15) == 0;
    }
};
template <>
struct FRow<true> {
    const __half* p;
    __device__ __forceinline__ FRow(const float*, const __half* base, int64_t ld, int b) : p(base + (int64_t)b * ld) {}
    __device__ __forceinline__ bool present() const { return true; }
    __device__ __forceinline__ float at(int64_t i) const { return __half2float(__ldg(p + i)); }
    template <int KD>
    __device__ __forceinline__ void load(int64_t i, float (&o)[KD]) const {
        static_assert(KD == 4 || KD == 8, "binary16 rows are read four or eight at a time");
        if constexpr (KD == 4) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p + i));
            hs_unpack_h2(u.x, o[0], o[1]);
            hs_unpack_h2(u.y, o[2], o[3]);
        } else {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + i));
            hs_unpack_h2(u.x, o[0], o[1]);
            hs_unpack_h2(u.y, o[2], o[3]);
            hs_unpack_h2(u.z, o[4], o[5]);
            hs_unpack_h2(u.w, o[6], o[7]);
        }
    }
    template <int KD>
    __device__ __forceinline__ bool aligned(int64_t i) const {
        return (reinterpret_cast<uintptr_t>(p + i) & (KD * 2 - 1)) == 0;
    }
};
// SRC 0: a, b float32; 1: a binary16, b float32 (or absent); 2: a and b binary16
template <int SRC>
using ARow = FRow<(SRC >= 1)>;
template <int SRC>
using BRow = FRow<(SRC == 2)>;

struct FuseConsts {
    float min_a, range_a, max_b_div, min_b, range_b;
    bool const_a, const_b;
    // cheap upper-bound path: reciprocals for a multiply-only estimate of the fused score, and the
    // margin that covers its distance from the exactly rounded value
    float rcp_a, rcp_b, delta;
};

__device__ __forceinline__ FuseConsts load_consts(const FuseParams& p, int b) {
    FuseConsts c;
    c.min_a = c.range_a = c.max_b_div = c.min_b = c.range_b = 0.f;
    c.const_a = c.const_b = false;
    c.rcp_a = c.rcp_b = c.delta = 0.f;
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
    c.rcp_a = c.const_a ? 0.f : 1.0f / c.range_a;
    if (p.mode == HS_FUSE_HYBRID_BM25) c.rcp_b = 1.0f / c.max_b_div;
    else c.rcp_b = (p.b == nullptr || c.const_b) ? 0.f : 1.0f / c.range_b;
    // both normalised terms lie in [0, 1]; each estimate is within a few ulp of the exact value
    c.delta = 8e-6f * (fabsf(p.wa32) + fabsf(p.wb32)) + 1e-30f;
    return c;
}

// Multiply-only estimate of fuse_score: |estimate - exact| < c.delta.  Used only to REJECT elements that
// cannot reach the current threshold; everything else goes through the exact path.
__device__ __forceinline__ float fuse_score_estimate(const FuseParams& p, const FuseConsts& c, float a, float b) {
    const float an = c.const_a ? 1.0f : (a - c.min_a) * c.rcp_a;
    float bn;
    if (p.mode == HS_FUSE_HYBRID_BM25) bn = b * c.rcp_b;
    else bn = (p.b == nullptr) ? 0.f : (c.const_b ? 1.0f : (b - c.min_b) * c.rcp_b);
    return an * p.wa32 + bn * p.wb32;
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

// ---- starting bound from block maxima -------------------------------------------------------------
// kBoundBlocks blocks of kBoundDocs docs are spread evenly over the shard; one warp computes the best
// exact key of each block.  The k-th largest of those maxima is attained by k DISTINCT docs, hence it is
// a valid lower bound on the k-th best key of the whole shard -- and for k = 100 of 1024 maxima it sits
// around the top 0.1 % of all docs, so the main pass rejects ~99.9 % of the elements with two FMAs.
constexpr int kBoundBlocks = 1024;             // at most; smaller shards sample 512 or 256 blocks.  4096 blocks were
                                               // measured: the larger sampling pass costs more than the tighter bound saves
constexpr int kBoundDocs = 128;

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)v, m);
        const uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), m);
        const uint64_t o = ((uint64_t)hi << 32) | lo;
        if (o > v) v = o;
    }
    return v;
}

template <int SRC>
__global__ void __launch_bounds__(kThreads) fuse_blockmax_kernel(const FuseParams p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (blk >= p.n_bound) return;
    const int64_t stride = (p.n / p.n_bound) / kBoundDocs * kBoundDocs;        // >= kBoundDocs by the host check
    const int64_t start = (int64_t)blk * stride;
    const FuseConsts c = load_consts(p, b);
    const uint64_t below = p.below ? p.below[b] : ~0ull;
    const ARow<SRC> pa(p.a, p.a16, p.ld_a, b);
    const BRow<SRC> pb(p.b, p.b16, SRC == 2 ? p.ld_a : p.ld, b);
    uint64_t best = 0;
#pragma unroll
    for (int u = 0; u < kBoundDocs / 32; ++u) {
        const int64_t i = start + u * 32 + lane;
        const float a = pa.at(i);
        const float bb = pb.present() ? pb.at(i) : 0.f;
        const uint64_t kk = hs_make_key(fuse_score(p, c, a, bb), (uint32_t)(p.doc_base + i));
        if (kk < below && kk > best) best = kk;
    }
    best = warp_max_u64(best);
    if (lane == 0) p.blockmax[(int64_t)b * kBoundBlocks + blk] = best;
}

__global__ void __launch_bounds__(kThreads) fuse_bound_kernel(const FuseParams p) {
    __shared__ uint64_t m[kBoundBlocks];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < p.n_bound; i += kThreads) m[i] = p.blockmax[(int64_t)b * kBoundBlocks + i];
    __syncthreads();
    bitonic_sort_desc_n(m, p.n_bound);
    // selectors drop keys <= bound, and the doc that attains m[k-1] may itself be the k-th best: publish
    // one less (keys are unique integers).  0 if fewer than k blocks had a key: no bound.
    if (threadIdx.x == 0) p.gthr[b] = m[p.k - 1] > 0 ? m[p.k - 1] - 1 : 0;
}

// ---- main pass ------------------------------------------------------------------------------------------
// Optimistic streaming: with a starting bound almost nothing is appended, so the CTA first runs WITHOUT
// any block-wide synchronisation in the loop (loads of consecutive iterations overlap freely); an append
// that finds the candidate buffer full only raises a flag.  If the flag is up at the end the chunk is
// redone with the synchronised selector (always correct, used from the start when there is no bound).
template <int KP, int SRC>
__global__ void __launch_bounds__(kThreads) fuse_topk_kernel(const FuseParams p) {
    __shared__ Selector<KP> sel;
    __shared__ int overflow;
    constexpr int CAP = Selector<KP>::CAP;
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    // docs per CTA: a multiple of the widest vector step (4096 docs: binary16 a and b, 8 per load, 2 loads in flight), so
    // that only the shard's last chunk has a ragged tail for the element-wise path
    constexpr int64_t kSpanUnit = kThreads * 16;
    const int64_t span = ((p.n + p.n_chunks - 1) / p.n_chunks + kSpanUnit - 1) / kSpanUnit * kSpanUnit;
    const int64_t lo = (int64_t)chunk * span;
    const int64_t hi = (lo + span < p.n) ? lo + span : p.n;
    const FuseConsts c = load_consts(p, b);
    const uint64_t below = p.below ? p.below[b] : ~0ull;
    const ARow<SRC> pa(p.a, p.a16, p.ld_a, b);
    const BRow<SRC> pb(p.b, p.b16, SRC == 2 ? p.ld_a : p.ld, b);
    unsigned long long* gthr = p.gthr + b;
    __shared__ uint64_t t0_shared;
    if (tid == 0) overflow = 0;
    sel.init();
    // ONE read of the published bound per CTA: other CTAs of the query publish concurrently (atomicMax in prune), and threads
    // that read different values would split between the two paths below, whose barriers differ (an intermittent
    // cudaErrorIllegalInstruction for k > 512, where no bound kernel publishes before this kernel starts)
    if (tid == 0) t0_shared = sel.bound(gthr);
    __syncthreads();
    const uint64_t t0 = t0_shared;
    bool redo = (t0 == 0);                                   // no bound: go straight to the safe path
    if (!redo) {
        sel.start_empty();
        const float thr_f = hs_dec_f32((uint32_t)(t0 >> 32));
        // per-thread pre-filter: the estimate of fuse_score_estimate as one subtract and two FMAs per element,
        // (a - sa) * ca + ((b - sb) * cb + c0) with the margin folded into c0.  A thread whose 8 elements all
        // stay below the bound (all but ~0.6 % of them) does nothing else in the iteration.
        float sa = 0.f, ca = 1.f, sb = 0.f, cb = 0.f, c0 = 0.f;
        if (p.mode != HS_FUSE_RAW) {
            sa = c.const_a ? 0.f : c.min_a;
            ca = c.const_a ? 0.f : c.rcp_a * p.wa32;
            c0 = c.delta + (c.const_a ? p.wa32 : 0.f);
            if (p.mode == HS_FUSE_HYBRID_BM25) {
                cb = c.rcp_b * p.wb32;
            } else if (pb.present()) {
                sb = c.const_b ? 0.f : c.min_b;
                cb = c.const_b ? 0.f : c.rcp_b * p.wb32;
                c0 += c.const_b ? p.wb32 : 0.f;
            }
        }
        // exact key of one element, appended to the candidate buffer by a warp-aggregated atomic
        auto consider = [&](float av, float bv, int64_t i, bool valid) {
            uint64_t kk = 0;
            if (valid && !(p.mode != HS_FUSE_RAW && fuse_score_estimate(p, c, av, bv) + c.delta < thr_f)) {
                kk = hs_make_key(fuse_score(p, c, av, bv), (uint32_t)(p.doc_base + i));
                if (kk >= below || kk <= t0) kk = 0;
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, kk != 0);
            if (m != 0) {
                const int leader = __ffs(m) - 1;
                int pos0 = 0;
                if (lane == leader) pos0 = atomicAdd(&sel.cnt, __popc(m));
                pos0 = __shfl_sync(0xFFFFFFFFu, pos0, leader);
                if (kk != 0) {
                    const int pos = pos0 + __popc(m & ((1u << lane) - 1u));
                    if (pos < CAP) sel.buf[pos] = kk;
                    else overflow = 1;
                }
            }
        };
        constexpr int kD = (SRC == 2) ? 8 : 4;                  // docs per vector load
        constexpr int kV = 2;                                   // vector loads per array, thread and iteration
        constexpr int kStep = kThreads * kD * kV;               // 2048 / 4096 docs per iteration
        const bool vec_ok = pa.template aligned<kD>(lo) && pb.template aligned<kD>(lo);
        for (int64_t base = lo; base < hi; base += kStep) {
            if (vec_ok && base + kStep <= hi) {                 // block-uniform
                float ax[kV][kD], bx[kV][kD];
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    const int64_t i = base + v * (kThreads * kD) + tid * kD;
                    pa.template load<kD>(i, ax[v]);
                    pb.template load<kD>(i, bx[v]);
                }
                bool cand = false;
#pragma unroll
                for (int v = 0; v < kV; ++v)
#pragma unroll
                    for (int e = 0; e < kD; ++e)
                        cand |= !(__fmaf_rn(__fsub_rn(ax[v][e], sa), ca, __fmaf_rn(__fsub_rn(bx[v][e], sb), cb, c0)) < thr_f);
                if (!cand) continue;
                // rare (about one thread in a thousand): exact keys of this thread's survivors, appended one by one
#pragma unroll
                for (int v = 0; v < kV; ++v) {
#pragma unroll
                    for (int e = 0; e < kD; ++e) {
                        if (__fmaf_rn(__fsub_rn(ax[v][e], sa), ca, __fmaf_rn(__fsub_rn(bx[v][e], sb), cb, c0)) < thr_f) continue;
                        const int64_t i = base + v * (kThreads * kD) + tid * kD + e;
                        const uint64_t kk = hs_make_key(fuse_score(p, c, ax[v][e], bx[v][e]), (uint32_t)(p.doc_base + i));
                        if (kk >= below || kk <= t0) continue;
                        const int pos = atomicAdd(&sel.cnt, 1);
                        if (pos < CAP) sel.buf[pos] = kk;
                        else overflow = 1;
                    }
                }
            } else {                                            // unaligned rows and the ragged tail
                for (int64_t i0 = base + tid; i0 < base + kStep; i0 += kThreads) {
                    const bool valid = i0 < hi;
                    consider(valid ? pa.at(i0) : 0.f, (valid && pb.present()) ? pb.at(i0) : 0.f, i0, valid);
                }
            }
        }
        __syncthreads();
        redo = overflow != 0;
        if (redo) sel.init();
    }
    if (redo) {
        for (int64_t base = lo; base < hi; base += kThreads * kItems) {
            uint64_t key[kItems];
            float av[kItems], bv[kItems];
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int64_t i = base + j * kThreads + tid;
                av[j] = (i < hi) ? pa.at(i) : 0.f;
                bv[j] = (pb.present() && i < hi) ? pb.at(i) : 0.f;
            }
            // score the current bound stands for (bound 0 = nothing known yet -> -inf)
            const uint64_t t = sel.bound(gthr);
            const float thr_f = (t == 0) ? __int_as_float(0xff800000) : hs_dec_f32((uint32_t)(t >> 32));
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int64_t i = base + j * kThreads + tid;
                key[j] = 0;
                if (i < hi) {
                    if (p.mode != HS_FUSE_RAW && fuse_score_estimate(p, c, av[j], bv[j]) + c.delta < thr_f) continue;
                    const uint64_t kk = hs_make_key(fuse_score(p, c, av[j], bv[j]), (uint32_t)(p.doc_base + i));
                    if (kk < below) key[j] = kk;
                }
            }
            sel.push(key, p.k, gthr);
        }
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

// Same contract as topk_merge_kernel for MANY lists (the per-CTA candidate lists of fuse_topk_kernel):
// every list is sorted descending, so thread t walks list t (and t + 256, ...) from the top and stops
// at the first key that is not above the running threshold.  The number of rounds is the deepest any
// list reaches into the final top k (a handful for random data), not n_lists * k / 1024.
template <int KP>
__global__ void __launch_bounds__(kThreads) topk_merge_walk_kernel(const uint64_t* __restrict__ keys, int n_lists,
                                                                   int B, int k, int64_t list_stride,
                                                                   int64_t query_stride, uint64_t* __restrict__ out) {
    __shared__ Selector<KP> sel;
    const int b = blockIdx.x, tid = threadIdx.x;
    sel.init();
    int pos[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) pos[j] = (j * kThreads + tid < n_lists) ? 0 : k;
    // requires n_lists <= kItems * kThreads (the dispatcher falls back to the flat kernel otherwise)
    while (true) {
        uint64_t key[kItems];
        const uint64_t t = sel.thr;
        int any = 0;
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            key[j] = 0;
            if (pos[j] < k) {
                const int l = j * kThreads + tid;
                const uint64_t kk = keys[l * list_stride + (int64_t)b * query_stride + pos[j]];
                if (kk > t) {
                    key[j] = kk;
                    ++pos[j];
                    any = 1;
                } else {
                    pos[j] = k;      // sorted: nothing further down this list can qualify
                }
            }
        }
        if (!__syncthreads_or(any)) break;
        sel.push(key, k);
    }
    sel.finish(k);
    for (int i = tid; i < k; i += kThreads) out[(int64_t)b * k + i] = sel.buf[i];
}

// thr[b] = score of the kth best key of query b (-inf when the list is shorter: no bound)
// shard-local doc id of every key (-1 for an empty slot or a doc of another shard): the candidate list handed to
// hs_bm25_score_docs by the verified mode
__global__ void keys_local_docs_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t doc_base, int64_t n_docs,
                                       int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    const int64_t d = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu)) - doc_base;
    out[i] = (key != 0 && d >= 0 && d < n_docs) ? d : -1;
}

__global__ void keys_kth_score_kernel(const uint64_t* __restrict__ keys, int B, int k, int kth, float* __restrict__ thr) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint64_t key = keys[(int64_t)b * k + kth - 1];
    thr[b] = key != 0 ? hs_dec_f32((uint32_t)(key >> 32)) : __int_as_float(0xff800000);
}

// Candidate lists of the FILTER epilogue of the tensor-core scan (dense_gemm.cu) -> final ranking.  One CTA per
// query: the best k_sel RAW keys (score = cosine) among the min(cnt, cap) appended candidates and n_extra keys from
// elsewhere (the sample pass that produced the threshold), then -- HS_FUSE_SEARCHER -- every survivor is re-keyed with
// the fused score f32(minmax(cos) * f32(w_a)) (core.py:264-268 with lexical weight 0) under the FINAL stats and the
// list is sorted again; the best k_out go out.  k_sel > k_out covers fused-score ties at the cut (the map cos -> fused
// score is monotone but not injective).  cnt > cap raises *overflow: the caller must redo the batch unfiltered.
template <int KP>
__global__ void __launch_bounds__(kThreads) cand_select_kernel(const uint64_t* __restrict__ cand,
                                                               const uint32_t* __restrict__ cand_cnt, int n_seg, int cap,
                                                               const uint64_t* __restrict__ extra, int n_extra,
                                                               const FuseParams p, int k_sel, int k_out,
                                                               uint64_t* __restrict__ out, int32_t* overflow) {
    __shared__ Selector<KP> sel;
    __shared__ int s_pre[1025];                     // exclusive prefix of the segments' fill counts (n_seg <= 1024)
    const int b = blockIdx.x, tid = threadIdx.x;
    sel.init();
    int over = 0;
    for (int s = tid; s < n_seg; s += kThreads) {
        uint32_t c = cand_cnt[(int64_t)b * n_seg + s];
        if (c > (uint32_t)cap) {
            over = 1;
            c = (uint32_t)cap;
        }
        s_pre[s + 1] = (int)c;
    }
    if (__syncthreads_or(over) && tid == 0) atomicOr(overflow, 1);
    if (tid == 0) {
        s_pre[0] = 0;
        for (int s = 0; s < n_seg; ++s) s_pre[s + 1] += s_pre[s];
    }
    __syncthreads();
    // flattened index i over the FILLED slots only: segment = last s with s_pre[s] <= i
    const int filled = s_pre[n_seg];
    const int total = filled + n_extra;
    for (int base = 0; base < total; base += kThreads * kItems) {
        uint64_t key[kItems];
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const int i = base + j * kThreads + tid;
            key[j] = 0;
            if (i < filled) {
                int lo = 0, hi = n_seg;                       // invariant: s_pre[lo] <= i < s_pre[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_pre[mid] <= i) lo = mid;
                    else hi = mid;
                }
                key[j] = cand[((int64_t)b * n_seg + lo) * cap + (i - s_pre[lo])];
            } else if (i < total) {
                key[j] = extra[(int64_t)b * n_extra + (i - filled)];
            }
        }
        sel.push(key, k_sel);
    }
    sel.finish(k_sel);
    if (p.mode != HS_FUSE_RAW) {
        const FuseConsts c = load_consts(p, b);
        for (int i = tid; i < KP; i += kThreads) {
            const uint64_t key = sel.buf[i];
            uint64_t nk = 0;
            if (i < k_sel && key != 0) {
                const float cosv = hs_dec_f32((uint32_t)(key >> 32));
                nk = hs_make_key(fuse_score(p, c, cosv, 0.f), 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu));
            }
            sel.buf[i] = nk;
        }
        __syncthreads();
        bitonic_sort_desc_n(sel.buf, KP);
    }
    for (int i = tid; i < k_out; i += kThreads) out[(int64_t)b * k_out + i] = sel.buf[i];
}

// ---- exact verification of an approximate dense scan ("screen with the tensor cores, verify in the conformance order") ---------
// The bf16 GEMM gives every cosine within eps of the exact (float64-accumulated) value.  Two small kernels turn the
// approximate hybrid ranking into the EXACT one:
//   verify_stats_kernel  the exact global min / max cosine: only docs whose approximate score lies within 2 eps of the
//                        approximate extreme can attain it; the GEMM epilogue listed (a superset of) them per segment
//   verify_topk_kernel   the approximate select returned k' > k keys; their cosines are recomputed exactly, the fused score
//                        re-evaluated with the reference's rounding steps and the list re-sorted.  A doc outside the list
//                        has an exact fused score <= (k'-th approximate score) + delta, delta = |w_a| eps / range: if the
//                        k-th exact score clears that bound the top k is provably the exact one, else the query is flagged
//                        and the caller redoes it in the exact mode.
struct VerifyParams {
    const float* v;          // [n, ld] float32 rows
    const float* vnorm;      // [n]
    const float* q;          // [B, ld_q]
    int64_t n, ld, ld_q;
    int dim;
    uint32_t doc_base;
};

constexpr int kVerifyList = 512;

__global__ void __launch_bounds__(kThreads) verify_stats_kernel(const VerifyParams vp, const uint64_t* __restrict__ ext,
                                                                const uint32_t* __restrict__ ext_cnt, int n_seg, int cap,
                                                                float eps2, uint32_t* __restrict__ stats,
                                                                int32_t* __restrict__ flags) {
    __shared__ uint32_t lst[2][kVerifyList];         // shard-local docs that can attain the max (0) / min (1)
    __shared__ int cnt[2];
    __shared__ uint32_t best[2];                     // encoded exact max / min
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) {
        cnt[tid] = 0;
        best[tid] = tid == 0 ? 0u : 0xFFFFFFFFu;
    }
    __syncthreads();
    const float amx = hs_dec_f32(stats[b * 4 + HS_STAT_MAX_A]), amn = hs_dec_f32(stats[b * 4 + HS_STAT_MIN_A]);
    if (!(amn <= amx)) return;                       // the shard holds no doc: nothing to verify
    int over = 0;
    for (int s = tid; s < n_seg * 2; s += kThreads) {
        const int side = s & 1;
        uint32_t c = ext_cnt[(int64_t)b * n_seg * 2 + s];
        if (c > (uint32_t)cap) {
            over = 1;
            c = (uint32_t)cap;
        }
        const uint64_t* e = ext + ((int64_t)b * n_seg * 2 + s) * cap;
        for (uint32_t i = 0; i < c; ++i) {
            const uint64_t key = e[i];
            const float sc = hs_dec_f32((uint32_t)(key >> 32));
            if (side == 0 ? sc >= amx - eps2 : sc <= amn + eps2) {
                const int pos = atomicAdd(&cnt[side], 1);
                if (pos < kVerifyList) lst[side][pos] = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
                else over = 1;
            }
        }
    }
    over = __syncthreads_or(over);
    const float* q = vp.q + (int64_t)b * vp.ld_q;
    const float qn = hs_exact_norm_warp(q, vp.dim, lane);
    for (int side = 0; side < 2; ++side) {
        const int m = cnt[side] < kVerifyList ? cnt[side] : kVerifyList;
        for (int i = warp; i < m; i += kThreads / 32) {
            const uint32_t d = lst[side][i];
            const float c = hs_exact_cos_warp(q, qn, vp.v + (int64_t)d * vp.ld, vp.vnorm[d], vp.dim, vp.ld, lane);
            if (lane == 0) {
                if (side == 0) atomicMax(&best[0], hs_enc_f32(c));
                else atomicMin(&best[1], hs_enc_f32(c));
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (over || cnt[0] == 0 || cnt[1] == 0) {
            atomicOr(flags + b, 1);                  // a list overflowed: the exact extremes are not guaranteed
        } else {
            stats[b * 4 + HS_STAT_MAX_A] = best[0];
            stats[b * 4 + HS_STAT_MIN_A] = best[1];
        }
    }
}

// VT threads: the re-scoring is one warp per candidate and latency-bound (a 1.5 KB row + a float64 reduction each), so the
// lists of <= 512 keys run 32 warps per query; threads >= kThreads fall through the sort loops (n / 2 <= kThreads there).
template <int KP, int VT>
__global__ void __launch_bounds__(VT) verify_topk_kernel(const VerifyParams vp, const FuseParams p,
                                                         const uint64_t* __restrict__ approx, int k_sel, int k_out,
                                                         float eps, const double* __restrict__ b_cand, float eps_b,
                                                         uint64_t* __restrict__ out, int32_t* __restrict__ flags) {
    static_assert(VT == kThreads || KP / 2 <= kThreads, "bitonic_sort_desc_n strides by kThreads");
    __shared__ uint64_t keys[KP];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FuseConsts c = load_consts(p, b);
    const float* q = vp.q + (int64_t)b * vp.ld_q;
    const float qn = hs_exact_norm_warp(q, vp.dim, lane);
    for (int i = tid; i < KP; i += VT) keys[i] = 0;
    __syncthreads();
    for (int i = warp; i < k_sel; i += VT / 32) {
        const uint64_t key = approx[(int64_t)b * k_sel + i];
        if (key == 0) continue;                                                    // warp-uniform
        const uint32_t gid = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
        const uint32_t d = gid - vp.doc_base;
        const float cs = hs_exact_cos_warp(q, qn, vp.v + (int64_t)d * vp.ld, vp.vnorm[d], vp.dim, vp.ld, lane);
        if (lane == 0) {
            // the b value: re-computed exactly for the candidate (b_cand, float64 -> the float32 the reference holds), or
            // read from the exact float32 array
            const float bv = b_cand != nullptr ? __double2float_rn(b_cand[(int64_t)b * k_sel + i])
                                               : (p.b != nullptr ? p.b[(int64_t)b * p.ld + d] : 0.f);
            keys[i] = hs_make_key(fuse_score(p, c, cs, bv), gid);
        }
    }
    __syncthreads();
    bitonic_sort_desc_n(keys, KP);
    for (int i = tid; i < k_out; i += VT) out[(int64_t)b * k_out + i] = keys[i];
    if (tid == 0) {
        // soundness: every doc outside the approximate list scores at most (its last key) + delta exactly
        const uint64_t last = approx[(int64_t)b * k_sel + k_sel - 1];
        if (last != 0) {                                    // the list is full: there are docs outside it
            const float f_last = hs_dec_f32((uint32_t)(last >> 32));
            // + the b term's own screen error (binary16 storage of b / max_b: <= eps_b each)
            const float delta = (c.const_a ? 0.f : fabsf(p.wa32) * eps / c.range_a * 1.001f) + fabsf(p.wb32) * eps_b * 1.001f + 1e-6f;
            const uint64_t kth = keys[k_out - 1];
            const float f_k = kth != 0 ? hs_dec_f32((uint32_t)(kth >> 32)) : -1e30f;
            if (!(f_k > f_last + delta)) atomicOr(flags + b, 2);
        }
    }
}

// CTAs (= candidate lists of k keys) per query.  Every list costs k keys of output, a sort and a slot in the merge,
// whatever its doc range: on a small shard with a large k (doc-sharded runs, the 256 / 512-key lists of the verified
// mode) 296 lists of a few thousand docs made the select chain a FIXED cost.  So: at least 512 k docs per list (measured
// at 10 M docs x 256 queries, k = 256: 64 k -> 128 k -> 512 k docs per list = 0.62 -> 0.80 -> 0.88 of HBM for the main
// pass), but at least 8 x 296 CTAs over the batch, at most 296 lists, at least 4096 docs each.
// HS_TOPK_LIST_MULT / HS_TOPK_FILL override the two factors (A/B runs).
int n_chunks_for(int64_t n, int B, int k) {
    int64_t cap = (n + kChunkDocs - 1) / kChunkDocs;
    if (cap > kMaxChunks) cap = kMaxChunks;
    if (cap < 1) cap = 1;
    static const int64_t mult = getenv("HS_TOPK_LIST_MULT") != nullptr ? atoll(getenv("HS_TOPK_LIST_MULT")) : 512;
    int64_t c = n / ((mult > 0 ? mult : 512) * (k > 0 ? k : 1));
    static const int64_t fillx = getenv("HS_TOPK_FILL") != nullptr ? atoll(getenv("HS_TOPK_FILL")) : 8;
    const int64_t fill = ((fillx > 0 ? fillx : 8) * kMaxChunks + (B > 0 ? B : 1) - 1) / (B > 0 ? B : 1);
    if (c < fill) c = fill;
    if (c > cap) c = cap;
    if (c < 1) c = 1;
    return (int)c;
}

template <int KP>
int launch_merge(const uint64_t* keys, int n_lists, int B, int k, int64_t list_stride, int64_t query_stride,
                 uint64_t* out, cudaStream_t st) {
    if (n_lists >= 32 && n_lists <= kItems * kThreads)
        topk_merge_walk_kernel<KP><<<B, kThreads, 0, st>>>(keys, n_lists, B, k, list_stride, query_stride, out);
    else
        topk_merge_kernel<KP><<<B, kThreads, 0, st>>>(keys, n_lists, B, k, list_stride, query_stride, out);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int merge_dispatch(const uint64_t* keys, int n_lists, int B, int k, int64_t list_stride, int64_t query_stride,
                   uint64_t* out, cudaStream_t st) {
    if (k <= 128) return launch_merge<128>(keys, n_lists, B, k, list_stride, query_stride, out, st);
    if (k <= 256) return launch_merge<256>(keys, n_lists, B, k, list_stride, query_stride, out, st);
    if (k <= 512) return launch_merge<512>(keys, n_lists, B, k, list_stride, query_stride, out, st);
    return launch_merge<2048>(keys, n_lists, B, k, list_stride, query_stride, out, st);
}

template <int SRC>
static void launch_fuse_select(const FuseParams& p, dim3 grid, int B, int k, cudaStream_t st) {
    if (p.n_bound > 0) {
        dim3 bg((unsigned)(p.n_bound / (kThreads / 32)), (unsigned)B);
        fuse_blockmax_kernel<SRC><<<bg, kThreads, 0, st>>>(p);
        fuse_bound_kernel<<<B, kThreads, 0, st>>>(p);
    }
    if (k <= 128)
        fuse_topk_kernel<128, SRC><<<grid, kThreads, 0, st>>>(p);
    else if (k <= 256)
        fuse_topk_kernel<256, SRC><<<grid, kThreads, 0, st>>>(p);
    else if (k <= 512)
        fuse_topk_kernel<512, SRC><<<grid, kThreads, 0, st>>>(p);
    else
        fuse_topk_kernel<2048, SRC><<<grid, kThreads, 0, st>>>(p);
}

}  // namespace

extern "C" {

size_t hs_fuse_topk_workspace_bytes(int64_t n_docs, int32_t B, int32_t k) {
    if (n_docs <= 0 || B <= 0 || k <= 0) return 0;
    return ((size_t)B * n_chunks_for(n_docs, B, k) * (size_t)k + (size_t)B + (size_t)B * kBoundBlocks) * sizeof(uint64_t);
}

// a: float32 [B, ld], or binary16 [B, ld] when a_half; b: float32 [B, n_docs], or binary16 [B, ld] when b_half (a_half too)
static int fuse_topk_impl(int64_t n_docs, int64_t doc_base, int64_t ld, int32_t fuse_mode, const void* a, bool a_half,
                          const void* b, bool b_half, const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k,
                          const uint64_t* below_key, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                          cudaStream_t st) {
    HS_REQUIRE(B > 0 && B <= 65535 && k > 0 && k <= HS_TOPK_MAX, "hs_fuse_topk: B=%d k=%d out of range (k <= %d)", B,
               k, HS_TOPK_MAX);
    HS_REQUIRE(out_keys != nullptr, "hs_fuse_topk: out_keys is null");
    if (n_docs == 0) {
        HS_CUDA(cudaMemsetAsync(out_keys, 0, (size_t)B * k * sizeof(uint64_t), st));
        return HS_OK;
    }
    HS_REQUIRE(a != nullptr && ld >= n_docs, "hs_fuse_topk: a is null or row stride < n_docs");
    HS_REQUIRE(fuse_mode == HS_FUSE_RAW || fuse_mode == HS_FUSE_SEARCHER || fuse_mode == HS_FUSE_HYBRID_BM25,
               "hs_fuse_topk: unknown fuse_mode %d", fuse_mode);
    HS_REQUIRE(fuse_mode == HS_FUSE_RAW || stats_enc != nullptr, "hs_fuse_topk: stats_enc is null");
    HS_REQUIRE(fuse_mode != HS_FUSE_HYBRID_BM25 || b != nullptr, "hs_fuse_topk: hybrid_bm25 needs b");
    HS_REQUIRE(fuse_mode != HS_FUSE_SEARCHER || b != nullptr || w_b == 0.0,
               "hs_fuse_topk: searcher fusion with w_b != 0 needs b");
    const size_t need = hs_fuse_topk_workspace_bytes(n_docs, B, k);
    HS_REQUIRE(workspace != nullptr && workspace_bytes >= need, "hs_fuse_topk: workspace too small (%zu < %zu)",
               workspace_bytes, need);
    FuseParams p;
    p.a = a_half ? nullptr : (const float*)a;
    p.a16 = a_half ? (const __half*)a : nullptr;
    p.b = b_half ? nullptr : (const float*)b;
    p.b16 = b_half ? (const __half*)b : nullptr;
    p.stats = stats_enc;
    p.below = below_key;
    p.gthr = (unsigned long long*)workspace;
    p.blockmax = (uint64_t*)workspace + B;
    p.cand = (uint64_t*)workspace + B + (size_t)B * kBoundBlocks;
    HS_CUDA(cudaMemsetAsync(p.gthr, 0, (size_t)B * sizeof(uint64_t), st));
    p.n = n_docs;
    p.ld = a_half ? n_docs : ld;          // b rows are a shard's [B, n_docs] array whenever a is the binary16 screen
    p.ld_a = ld;
    p.doc_base = doc_base;
    p.mode = fuse_mode;
    p.k = k;
    p.n_chunks = n_chunks_for(n_docs, B, k);
    p.wa32 = (float)w_a;   // numpy: float32 array * python float -> float32(w)
    p.wb32 = (float)w_b;
    p.wa64 = w_a;
    dim3 grid((unsigned)p.n_chunks, (unsigned)B);
    // starting bound from block maxima when the shard is large enough for it to pay (see above)
    static const bool no_bound = getenv("HS_NO_BOUND") != nullptr;      // A/B switch, read once
    // as many sampled blocks (1024 / 512 / 256) as leave >= 2 strides of 128 docs per block and hold k twice
    p.n_bound = 0;
    for (int nb = kBoundBlocks; nb >= 256 && !no_bound; nb >>= 1)
        if (n_docs >= (int64_t)nb * kBoundDocs * 2 && k <= nb / 2) {
            p.n_bound = nb;
            break;
        }
    if (a_half && b_half) launch_fuse_select<2>(p, grid, B, k, st);
    else if (a_half) launch_fuse_select<1>(p, grid, B, k, st);
    else launch_fuse_select<0>(p, grid, B, k, st);
    HS_LAUNCH_CHECK();
    // candidate layout [B, n_chunks, k]: list stride k, query stride n_chunks * k
    return merge_dispatch(p.cand, p.n_chunks, B, k, (int64_t)k, (int64_t)p.n_chunks * k, out_keys, st);
}

int hs_fuse_topk(const hs_index* idx, int32_t fuse_mode, const float* a, const float* b,
                 const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k,
                 const uint64_t* below_key, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                 void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_fuse_topk: idx is null");
    return fuse_topk_impl(idx->n_docs, idx->doc_base, idx->n_docs, fuse_mode, a, false, b, false, stats_enc, w_a, w_b, B, k,
                          below_key, workspace, workspace_bytes, out_keys, (cudaStream_t)stream);
}

int hs_fuse_topk_f16(const hs_index* idx, int32_t fuse_mode, const uint16_t* a_f16, const void* b, int32_t b_is_f16,
                     int64_t ld, const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k, void* workspace,
                     size_t workspace_bytes, uint64_t* out_keys, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_fuse_topk_f16: idx is null");
    HS_REQUIRE(fuse_mode != HS_FUSE_RAW, "hs_fuse_topk_f16: screen scores are only fused (SEARCHER / HYBRID_BM25)");
    HS_REQUIRE(!b_is_f16 || (fuse_mode == HS_FUSE_HYBRID_BM25 && b != nullptr),
               "hs_fuse_topk_f16: a binary16 b array is the BM25 screen of HS_FUSE_HYBRID_BM25");
    return fuse_topk_impl(idx->n_docs, idx->doc_base, ld, fuse_mode, a_f16, true, b, b_is_f16 != 0, stats_enc, w_a, w_b, B, k,
                          nullptr, workspace, workspace_bytes, out_keys, (cudaStream_t)stream);
}

int hs_keys_local_docs(const uint64_t* keys, int64_t n, int64_t doc_base, int64_t n_docs, int64_t* local_ids, void* stream) {
    HS_REQUIRE(keys != nullptr && local_ids != nullptr && n >= 0, "hs_keys_local_docs: bad arguments");
    if (n == 0) return HS_OK;
    keys_local_docs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys, n, doc_base, n_docs, local_ids);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_topk_select(const float* x, int64_t n, int64_t ld, int64_t doc_base, int32_t B, int32_t k, void* workspace,
                   size_t workspace_bytes, uint64_t* out_keys, void* stream) {
    HS_REQUIRE(n >= 0 && doc_base >= 0 && n + doc_base <= 0xFFFFFFFFll, "hs_topk_select: doc ids must fit uint32");
    return fuse_topk_impl(n, doc_base, ld, HS_FUSE_RAW, x, false, nullptr, false, nullptr, 1.0, 0.0, B, k, nullptr, workspace,
                          workspace_bytes, out_keys, (cudaStream_t)stream);
}

int hs_keys_kth_score(const uint64_t* keys, int32_t B, int32_t k, int32_t kth, float* thr, void* stream) {
    HS_REQUIRE(keys != nullptr && thr != nullptr && B > 0 && kth >= 1 && kth <= k, "hs_keys_kth_score: bad arguments");
    keys_kth_score_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(keys, B, k, kth, thr);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_cand_select(const uint64_t* cand, const uint32_t* cand_cnt, int32_t n_seg, int32_t cand_cap,
                   const uint64_t* extra_keys, int32_t n_extra, int32_t fuse_mode, const uint32_t* stats_enc, double w_a,
                   int32_t B, int32_t k_sel, int32_t k_out, uint64_t* out_keys, int32_t* overflow, void* stream) {
    HS_REQUIRE(cand != nullptr && cand_cnt != nullptr && cand_cap > 0 && n_seg > 0 && n_seg <= 1024 && out_keys != nullptr &&
                   overflow != nullptr, "hs_cand_select: null pointer or n_seg out of range");
    HS_REQUIRE(B > 0 && k_out > 0 && k_out <= k_sel && k_sel <= HS_TOPK_MAX && n_extra >= 0 &&
                   (n_extra == 0 || extra_keys != nullptr),
               "hs_cand_select: bad sizes (k_out=%d k_sel=%d)", k_out, k_sel);
    HS_REQUIRE(fuse_mode == HS_FUSE_RAW || (fuse_mode == HS_FUSE_SEARCHER && stats_enc != nullptr),
               "hs_cand_select: fuse_mode must be RAW or SEARCHER (with stats)");
    FuseParams p;
    memset(&p, 0, sizeof(p));
    p.stats = stats_enc;
    p.mode = fuse_mode;
    p.wa32 = (float)w_a;
    p.wa64 = w_a;
    cudaStream_t st = (cudaStream_t)stream;
    if (k_sel <= 128)
        cand_select_kernel<128><<<B, kThreads, 0, st>>>(cand, cand_cnt, n_seg, cand_cap, extra_keys, n_extra, p, k_sel, k_out, out_keys, overflow);
    else if (k_sel <= 512)
        cand_select_kernel<512><<<B, kThreads, 0, st>>>(cand, cand_cnt, n_seg, cand_cap, extra_keys, n_extra, p, k_sel, k_out, out_keys, overflow);
    else
        cand_select_kernel<2048><<<B, kThreads, 0, st>>>(cand, cand_cnt, n_seg, cand_cap, extra_keys, n_extra, p, k_sel, k_out, out_keys, overflow);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

static int fill_verify(const hs_index* idx, const float* queries, int64_t ld_q, VerifyParams& vp, const char* who) {
    HS_REQUIRE(idx != nullptr && queries != nullptr, "%s: null pointer", who);
    if (idx->vectors == nullptr) {
        hs_set_error("%s: index has no dense matrix", who);
        return HS_ERR_STATE;
    }
    HS_REQUIRE(ld_q >= idx->dim, "%s: ld_q < dim", who);
    vp.v = idx->vectors;
    vp.vnorm = idx->vnorm;
    vp.q = queries;
    vp.n = idx->n_docs;
    vp.ld = idx->ld;
    vp.ld_q = ld_q;
    vp.dim = idx->dim;
    vp.doc_base = (uint32_t)idx->doc_base;
    return HS_OK;
}

int hs_verify_stats(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, const uint64_t* ext,
                    const uint32_t* ext_cnt, int32_t n_seg, int32_t ext_cap, double eps, uint32_t* stats_enc, int32_t* flags,
                    void* stream) {
    VerifyParams vp;
    int rc = fill_verify(idx, queries, ld_q, vp, "hs_verify_stats");
    if (rc != HS_OK) return rc;
    if (B == 0 || idx->n_docs == 0) return HS_OK;
    HS_REQUIRE(B > 0 && ext != nullptr && ext_cnt != nullptr && n_seg > 0 && ext_cap > 0 && stats_enc != nullptr &&
                   flags != nullptr && eps >= 0.0, "hs_verify_stats: bad arguments");
    verify_stats_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(vp, ext, ext_cnt, n_seg, ext_cap, (float)(2.0 * eps),
                                                                   stats_enc, flags);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

static int verify_topk_impl(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t fuse_mode,
                            const float* b, const double* b_cand, double eps_b, const uint32_t* stats_enc, double w_a,
                            double w_b, const uint64_t* approx_keys, int32_t k_sel, int32_t k_out, double eps,
                            uint64_t* out_keys, int32_t* flags, void* stream) {
    VerifyParams vp;
    int rc = fill_verify(idx, queries, ld_q, vp, "hs_verify_topk");
    if (rc != HS_OK) return rc;
    if (B == 0) return HS_OK;
    HS_REQUIRE(B > 0 && approx_keys != nullptr && out_keys != nullptr && flags != nullptr && stats_enc != nullptr &&
                   k_out > 0 && k_out <= k_sel && k_sel <= 2048 && eps >= 0.0, "hs_verify_topk: bad arguments");
    HS_REQUIRE(fuse_mode == HS_FUSE_SEARCHER || (fuse_mode == HS_FUSE_HYBRID_BM25 && (b != nullptr || b_cand != nullptr)),
               "hs_verify_topk: fuse_mode must be SEARCHER or HYBRID_BM25 (with b)");
    FuseParams p;
    memset(&p, 0, sizeof(p));
    p.b = b;
    p.stats = stats_enc;
    p.n = idx->n_docs;
    p.ld = idx->n_docs;
    p.doc_base = idx->doc_base;
    p.mode = fuse_mode;
    p.wa32 = (float)w_a;
    p.wb32 = (float)w_b;
    p.wa64 = w_a;
    cudaStream_t st = (cudaStream_t)stream;
    if (k_sel <= 128)
        verify_topk_kernel<128, 1024><<<B, 1024, 0, st>>>(vp, p, approx_keys, k_sel, k_out, (float)eps, b_cand, (float)eps_b,
                                                      out_keys, flags);
    else if (k_sel <= 256)
        verify_topk_kernel<256, 1024><<<B, 1024, 0, st>>>(vp, p, approx_keys, k_sel, k_out, (float)eps, b_cand, (float)eps_b,
                                                      out_keys, flags);
    else if (k_sel <= 512)
        verify_topk_kernel<512, 1024><<<B, 1024, 0, st>>>(vp, p, approx_keys, k_sel, k_out, (float)eps, b_cand, (float)eps_b,
                                                      out_keys, flags);
    else
        verify_topk_kernel<2048, kThreads><<<B, kThreads, 0, st>>>(vp, p, approx_keys, k_sel, k_out, (float)eps, b_cand,
                                                                   (float)eps_b, out_keys, flags);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_verify_topk(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t fuse_mode, const float* b,
                   const uint32_t* stats_enc, double w_a, double w_b, const uint64_t* approx_keys, int32_t k_sel,
                   int32_t k_out, double eps, uint64_t* out_keys, int32_t* flags, void* stream) {
    return verify_topk_impl(idx, queries, B, ld_q, fuse_mode, b, nullptr, 0.0, stats_enc, w_a, w_b, approx_keys, k_sel, k_out,
                            eps, out_keys, flags, stream);
}

int hs_verify_topk_cand(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t fuse_mode,
                        const double* b_cand, double eps_b, const uint32_t* stats_enc, double w_a, double w_b,
                        const uint64_t* approx_keys, int32_t k_sel, int32_t k_out, double eps, uint64_t* out_keys,
                        int32_t* flags, void* stream) {
    HS_REQUIRE(b_cand != nullptr && eps_b >= 0.0, "hs_verify_topk_cand: b_cand is null");
    return verify_topk_impl(idx, queries, B, ld_q, fuse_mode, nullptr, b_cand, eps_b, stats_enc, w_a, w_b, approx_keys, k_sel,
                            k_out, eps, out_keys, flags, stream);
}

int hs_topk_merge(const uint64_t* keys, int32_t n_lists, int32_t B, int32_t k, uint64_t* out_keys, void* stream) {
    HS_REQUIRE(keys != nullptr && out_keys != nullptr, "hs_topk_merge: null pointer");
    HS_REQUIRE(n_lists > 0 && B > 0 && k > 0 && k <= HS_TOPK_MAX, "hs_topk_merge: bad sizes");
    // layout [n_lists, B, k]: list stride B * k, query stride k
    return merge_dispatch(keys, n_lists, B, k, (int64_t)B * k, (int64_t)k, out_keys, (cudaStream_t)stream);
}

}  // extern "C"
