// K2 -- dense query x doc cosine scan (replaces utils.py:28-54 batch_cosine_sim as called from
// core.py:170-176,252).  HBM-bound streaming kernel for small query batches:
//
//   * persistent CTAs (one per SM).  A producer warp feeds an 8-stage shared-memory ring with TMA
//     bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) of G contiguous rows per stage;
//     stage s belongs to consumer warp s, so each warp waits only on its own mbarrier and up to
//     8 x G x ld x 4 bytes (192 KB at d=384) are in flight per SM.
//   * a consumer warp owns a whole stage (G rows): lane l holds elements 128c + 4l + j of a row
//     (conflict-free LDS.128), BQ query vectors live in registers, so a row is read from HBM once
//     for all BQ queries of the launch.
//   * reduction in the "conformance order" (lane-sequential over (c, j), then lanes combined
//     16, 8, 4, 2, 1).  The G rows of a stage are reduced TOGETHER: at level L the warp exchanges the
//     partial sums of two row groups with one shuffle (lanes with bit (5-L) clear keep the first
//     group, the others the second), so G rows cost G-1 + log2(32/G) shuffles instead of 5G and end
//     with lane l holding the complete dot product of row rho(l) -- which makes the epilogue
//     (division, store, min/max) lane-parallel.  The pairing of partial sums is exactly the
//     butterfly's, so the result is bit-identical to it:
//       EXACT: products/sums in float64 -> bit-identical to oracle/hybrid_oracle.py:cosine_exact
//       FP32 : packed float32 FMAs (two partial sums per lane) -> as precise as the reference's float32 BLAS dot
//   * epilogue fused: cos = f32(dot) / (f32|q| * f32|v|) with the reference's zero-norm rules
//     (utils.py:44-50), per-query min/max folded into the stats slots (utils.py:67-68)
//
// Algorithmic bytes per launch: n_docs * ld * 4 (the corpus pass), independent of BQ.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kStages = kConsumerWarps;
// two consumer warpgroups + one producer warpgroup (only one of its lanes works).  The register file is
// split per SM sub-partition (16 K registers, 3 warps each here), so the kernel starts at <= 168
// registers/thread and then moves registers from the producer group to the consumers with setmaxnreg.
constexpr int kThreads = (kConsumerWarps + 4) * 32;
constexpr int kConsumerRegs = 224;
constexpr int kProducerRegs = 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct DenseParams {
    const float* v;       // [n, ld]
    const float* vnorm;   // [n]
    const float* q;       // [B, ld_q], this launch uses rows b0 .. b0+BQ-1
    float* cos;           // [B, n]
    uint32_t* stats;      // [B, 4] encoded, may be null
    int64_t n, ld, ld_q;
    int32_t dim, b0;
};

template <bool EXACT>
struct Acc {
    using type = float;
};
template <>
struct Acc<true> {
    using type = double;
};

__device__ __forceinline__ float shfl_xor_t(float v, int m) { return __shfl_xor_sync(0xFFFFFFFFu, v, m); }
__device__ __forceinline__ double shfl_xor_t(double v, int m) { return hs_shfl_xor_f64(v, m); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// Partial dot products of 2^L consecutive rows starting at r0, combined over the L highest lane bits.
template <int L, int NCHUNK, int BQ, bool EXACT>
struct GroupSum {
    using acc_t = typename Acc<EXACT>::type;
    static __device__ __forceinline__ void run(const float* __restrict__ tile, int64_t ld, int r0, int lane,
                                               const acc_t (&q)[BQ][NCHUNK][4], acc_t (&out)[BQ]) {
        acc_t a[BQ], b[BQ];
        GroupSum<L - 1, NCHUNK, BQ, EXACT>::run(tile, ld, r0, lane, q, a);
        GroupSum<L - 1, NCHUNK, BQ, EXACT>::run(tile, ld, r0 + (1 << (L - 1)), lane, q, b);
        constexpr int mask = 16 >> (L - 1);
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < BQ; ++i) {
            const acc_t send = upper ? a[i] : b[i];
            const acc_t keep = upper ? b[i] : a[i];
            out[i] = add_rn(keep, shfl_xor_t(send, mask));
        }
    }
};
template <int NCHUNK, int BQ, bool EXACT>
struct GroupSum<0, NCHUNK, BQ, EXACT> {
    using acc_t = typename Acc<EXACT>::type;
    static __device__ __forceinline__ void run(const float* __restrict__ tile, int64_t ld, int r0, int lane,
                                               const acc_t (&q)[BQ][NCHUNK][4], acc_t (&out)[BQ]) {
        const float* row = tile + (size_t)r0 * ld;
        if constexpr (EXACT) {
#pragma unroll
            for (int i = 0; i < BQ; ++i) out[i] = 0.0;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int e = c * 128 + lane * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < ld) v = *reinterpret_cast<const float4*>(row + e);
#pragma unroll
                for (int i = 0; i < BQ; ++i) {
                    out[i] = __fma_rn((double)v.x, q[i][c][0], out[i]);
                    out[i] = __fma_rn((double)v.y, q[i][c][1], out[i]);
                    out[i] = __fma_rn((double)v.z, q[i][c][2], out[i]);
                    out[i] = __fma_rn((double)v.w, q[i][c][3], out[i]);
                }
            }
        } else {
            // float32 mode: packed two-wide FMAs (fma.rn.f32x2, SASS FFMA2) halve the issue slots of the
            // inner loop; each lane keeps an (even j, odd j) pair of partial sums per query
            float2 acc2[BQ];
#pragma unroll
            for (int i = 0; i < BQ; ++i) acc2[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int e = c * 128 + lane * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < ld) v = *reinterpret_cast<const float4*>(row + e);
                const float2 vlo = make_float2(v.x, v.y), vhi = make_float2(v.z, v.w);
#pragma unroll
                for (int i = 0; i < BQ; ++i) {
                    acc2[i] = __ffma2_rn(vlo, make_float2(q[i][c][0], q[i][c][1]), acc2[i]);
                    acc2[i] = __ffma2_rn(vhi, make_float2(q[i][c][2], q[i][c][3]), acc2[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < BQ; ++i) out[i] = __fadd_rn(acc2[i].x, acc2[i].y);
        }
    }
};

// LG = log2(rows per stage); QG = consumer warps sharing a stage, each with its own BQ queries
// (the launch covers BQ * QG queries: warp w works on tile slot w / QG for queries (w % QG) * BQ ...)
template <int NCHUNK, int BQ, bool EXACT, int LG, int QG>
__global__ void __launch_bounds__(kThreads, 1) dense_scan_kernel(const DenseParams p) {
    using acc_t = typename Acc<EXACT>::type;
    constexpr int G = 1 << LG;
    constexpr int kSlots = kConsumerWarps / QG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_floats = (size_t)G * p.ld;
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * stage_floats * sizeof(float));
    uint64_t* empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + G - 1) / G;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], QG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kConsumerWarps) {
        // ---------------- producer warpgroup: gives its registers away; one elected lane issues the bulk
        // copies, stage = tile % 8
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        if (warp == kConsumerWarps && lane == 0) {
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = (int)(it % kStages);
                const uint32_t ph = (uint32_t)((it / kStages) & 1);
                mbar_wait(&empty[s], ph ^ 1u);
                const int64_t row0 = t * G;
                const int64_t rows = (p.n - row0 < G) ? (p.n - row0) : G;
                const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(float));
                mbar_expect_tx(&full[s], bytes);
                tma_bulk_g2s(stages + s * stage_floats, p.v + row0 * p.ld, bytes, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    // query registers: element e = 128c + 4*lane + j (zero beyond dim)
    const int slot = warp / QG;
    const int qbase = p.b0 + (warp % QG) * BQ;
    acc_t q[BQ][NCHUNK][4];
    float qn[BQ];
#pragma unroll
    for (int b = 0; b < BQ; ++b) {
        const float* qb = p.q + (int64_t)(qbase + b) * p.ld_q;
        double qq = 0.0;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = c * 128 + lane * 4 + j;
                const float x = (e < p.dim) ? qb[e] : 0.0f;
                q[b][c][j] = (acc_t)x;
                qq = __fma_rn((double)x, (double)x, qq);
            }
        }
        qq = hs_warp_sum_f64(qq);
        qn[b] = __double2float_rn(__dsqrt_rn(qq));
    }

    // after the LG exchange levels lane l holds row rho(l) of its stage (the remaining low lane bits
    // are reduced by a plain butterfly, so lanes differing only in those bits hold the same row)
    int rho = 0;
#pragma unroll
    for (int L = 1; L <= LG; ++L) rho |= ((lane >> (5 - L)) & 1) << (L - 1);
    constexpr int kRestMask = (32 >> LG) - 1;     // lane bits not consumed by the exchange levels
    const bool writer = (lane & kRestMask) == 0;

    float mn[BQ], mx[BQ];
#pragma unroll
    for (int b = 0; b < BQ; ++b) {
        mn[b] = __int_as_float(0x7f800000);
        mx[b] = __int_as_float(0xff800000);
    }

    // this CTA's tile sequence is t_i = blockIdx.x + i * gridDim.x (stage i % 8); the QG warps of slot s
    // consume i = s, s + kSlots, ...
    int64_t it = slot;
    for (int64_t t = blockIdx.x + (int64_t)slot * gridDim.x; t < n_tiles; t += (int64_t)kSlots * gridDim.x, it += kSlots) {
        const int stage = (int)(it % kStages);
        const uint32_t ph = (uint32_t)((it / kStages) & 1);
        const float* my_stage = stages + (size_t)stage * stage_floats;
        const int64_t row0 = t * G;
        const int rows = (int)((p.n - row0 < G) ? (p.n - row0) : G);
        const bool valid = writer && rho < rows;
        const float vn = valid ? __ldg(p.vnorm + row0 + rho) : 1.0f;   // issued before the wait: latency hidden
        mbar_wait(&full[stage], ph);

        acc_t sum[BQ];
        GroupSum<LG, NCHUNK, BQ, EXACT>::run(my_stage, p.ld, 0, lane, q, sum);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);     // stage can be refilled while we finish the epilogue
#pragma unroll
        for (int m = kRestMask + 1; m > 1;) {           // remaining butterfly steps (masks < 32 >> LG)
            m >>= 1;
#pragma unroll
            for (int b = 0; b < BQ; ++b) sum[b] = add_rn(sum[b], shfl_xor_t(sum[b], m));
        }
        if (valid) {
#pragma unroll
            for (int b = 0; b < BQ; ++b) {
                float dot;
                if constexpr (EXACT)
                    dot = __double2float_rn(sum[b]);
                else
                    dot = sum[b];
                // utils.py:44-52: zero query or zero row -> 0.0, else dot / (|q| * |v|) in float32
                float c = 0.0f;
                if (qn[b] != 0.0f && vn != 0.0f) c = __fdiv_rn(dot, __fmul_rn(qn[b], vn));
                p.cos[(int64_t)(qbase + b) * p.n + row0 + rho] = c;
                mn[b] = fminf(mn[b], c);
                mx[b] = fmaxf(mx[b], c);
            }
        }
    }
    if (p.stats != nullptr) {
#pragma unroll
        for (int b = 0; b < BQ; ++b) {
            float lo = mn[b], hi = mx[b];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, m));
                hi = fmaxf(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, m));
            }
            if (lane == 0 && lo <= hi) {
                atomicMin(&p.stats[(qbase + b) * 4 + HS_STAT_MIN_A], hs_enc_f32(lo));
                atomicMax(&p.stats[(qbase + b) * 4 + HS_STAT_MAX_A], hs_enc_f32(hi));
            }
        }
    }
}

// index-time: vnorm[i] = f32(sqrt(sum64 v[i,:]^2)) in the conformance order; warp per row
__global__ void row_norms_kernel(const float* __restrict__ v, int64_t n, int dim, int64_t ld,
                                 float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nchunk = (dim + 127) / 128;
    for (int64_t r = warp; r < n; r += nwarps) {
        const float* row = v + r * ld;
        double acc = 0.0;
        for (int c = 0; c < nchunk; ++c) {
            const int e = c * 128 + lane * 4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < ld) x = *reinterpret_cast<const float4*>(row + e);
            acc = __fma_rn((double)x.x, (double)x.x, acc);
            acc = __fma_rn((double)x.y, (double)x.y, acc);
            acc = __fma_rn((double)x.z, (double)x.z, acc);
            acc = __fma_rn((double)x.w, (double)x.w, acc);
        }
        acc = hs_warp_sum_f64(acc);
        if (lane == 0) out[r] = __double2float_rn(__dsqrt_rn(acc));
    }
}

template <int NCHUNK, int BQ, bool EXACT, int LG, int QG>
int launch_dense(const DenseParams& p, int num_sms, cudaStream_t st) {
    constexpr int G = 1 << LG;
    const size_t smem = (size_t)kStages * G * p.ld * sizeof(float) + 2 * kStages * sizeof(uint64_t);
    auto kern = dense_scan_kernel<NCHUNK, BQ, EXACT, LG, QG>;
    static size_t smem_set[16] = {0};                 // per instantiation: the attribute is set once, not per launch
    HS_CUDA(hs_smem_limit(kern, smem, smem_set));
    const int64_t n_tiles = (p.n + G - 1) / G;
    const int grid = (int)((n_tiles < num_sms) ? n_tiles : num_sms);
    kern<<<grid, kThreads, smem, st>>>(p);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

// rows per stage by row length: 8 stages must fit in ~200 KB of shared memory
//   NCHUNK <= 3 (ld <= 384): 16 rows   NCHUNK 4..6 (ld <= 768): 8 rows   NCHUNK 8 (ld <= 1024): 4 rows
template <int NCHUNK>
struct StageRows {
    static constexpr int LG = NCHUNK <= 3 ? 4 : (NCHUNK <= 6 ? 3 : 2);
};

// queries per warp: the query registers (NCHUNK * 4 per query, x2 in float64) must not spill
constexpr int max_bq(int nchunk, bool exact) {
    const int budget = 12;     // query registers per lane: NCHUNK * 4 per query (x2 in float64)
    const int cap = 4;         // measured on B200: 4 queries/warp x 2 warps/stage is the sweet spot at d=384
    (void)exact;
    int bq = 1;
    while (bq * 2 <= budget / nchunk && bq * 2 <= cap) bq *= 2;
    return bq;
}

template <int NCHUNK, int BQ, bool EXACT>
int dispatch_qg(int qg, const DenseParams& p, int num_sms, cudaStream_t st) {
    constexpr int LG = StageRows<NCHUNK>::LG;
    if (qg == 4) return launch_dense<NCHUNK, BQ, EXACT, LG, 4>(p, num_sms, st);
    if (qg == 2) return launch_dense<NCHUNK, BQ, EXACT, LG, 2>(p, num_sms, st);
    return launch_dense<NCHUNK, BQ, EXACT, LG, 1>(p, num_sms, st);
}

template <int NCHUNK, bool EXACT>
int dispatch_bq(int bq, int qg, const DenseParams& p, int num_sms, cudaStream_t st) {
    constexpr int kMaxBQ = max_bq(NCHUNK, EXACT);
    if constexpr (kMaxBQ >= 8) {
        if (bq == 8) return dispatch_qg<NCHUNK, 8, EXACT>(qg, p, num_sms, st);
    }
    if constexpr (kMaxBQ >= 4) {
        if (bq == 4) return dispatch_qg<NCHUNK, 4, EXACT>(qg, p, num_sms, st);
    }
    if constexpr (kMaxBQ >= 2) {
        if (bq == 2) return dispatch_qg<NCHUNK, 2, EXACT>(qg, p, num_sms, st);
    }
    if (bq == 1) return dispatch_qg<NCHUNK, 1, EXACT>(qg, p, num_sms, st);
    hs_set_error("dense_scan: internal: no kernel for BQ=%d NCHUNK=%d", bq, NCHUNK);
    return HS_ERR_ARG;
}

template <bool EXACT>
int dispatch_chunks(int nchunk, int bq, int qg, const DenseParams& p, int num_sms, cudaStream_t st) {
    switch (nchunk) {
        case 1: return dispatch_bq<1, EXACT>(bq, qg, p, num_sms, st);
        case 2: return dispatch_bq<2, EXACT>(bq, qg, p, num_sms, st);
        case 3: return dispatch_bq<3, EXACT>(bq, qg, p, num_sms, st);
        case 4: return dispatch_bq<4, EXACT>(bq, qg, p, num_sms, st);
        case 6: return dispatch_bq<6, EXACT>(bq, qg, p, num_sms, st);
        default: return dispatch_bq<8, EXACT>(bq, qg, p, num_sms, st);
    }
}

}  // namespace

extern "C" {

int hs_row_norms(const float* vectors, int64_t n, int32_t dim, int64_t ld, float* vnorm, void* stream) {
    HS_REQUIRE(n >= 0 && dim > 0 && ld >= dim && (ld % 4) == 0, "hs_row_norms: bad shape n=%lld dim=%d ld=%lld",
               (long long)n, dim, (long long)ld);
    if (n == 0) return HS_OK;
    HS_REQUIRE(vectors != nullptr && vnorm != nullptr, "hs_row_norms: null pointer");
    HS_REQUIRE(((uintptr_t)vectors & 15) == 0, "hs_row_norms: vectors must be 16-byte aligned");
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    row_norms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(vectors, n, dim, ld, vnorm);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

int hs_dense_scan(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, float* cos,
                  uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_dense_scan: idx is null");
    if (idx->n_docs == 0 || B == 0) return HS_OK;
    if (idx->vectors == nullptr) {
        hs_set_error("hs_dense_scan: index has no dense matrix (call hs_index_set_dense)");
        return HS_ERR_STATE;
    }
    HS_REQUIRE(queries != nullptr && cos != nullptr && B > 0 && ld_q >= idx->dim, "hs_dense_scan: bad arguments");
    HS_REQUIRE(mode == HS_DENSE_EXACT || mode == HS_DENSE_FP32,
               "hs_dense_scan: mode %d not available in this entry point", mode);
    const int nchunk_raw = (int)((idx->ld + 127) / 128);
    const int nchunk = nchunk_raw <= 4 ? nchunk_raw : (nchunk_raw <= 6 ? 6 : 8);
    DenseParams p;
    p.v = idx->vectors;
    p.vnorm = idx->vnorm;
    p.q = queries;
    p.cos = cos;
    p.stats = stats_enc;
    p.n = idx->n_docs;
    p.ld = idx->ld;
    p.ld_q = ld_q;
    p.dim = idx->dim;
    const bool exact = mode == HS_DENSE_EXACT;
    const int cap = max_bq(nchunk, exact);
    int b0 = 0;
    while (b0 < B) {
        // one corpus pass serves BQ * QG queries: BQ per warp (registers), QG warps per stage
        int bq = cap;
        while (bq > B - b0) bq >>= 1;
        int qg = 4;
        while (qg > 1 && bq * qg > B - b0) qg >>= 1;
        static const char* const force = getenv("HS_DENSE_FORCE");      // tuning aid "bq,qg", read once
        if (const char* f = force) {
            int fb = 0, fq = 0;
            if (sscanf(f, "%d,%d", &fb, &fq) == 2 && fb >= 1 && fb <= cap && fb * fq <= B - b0 &&
                (fq == 1 || fq == 2 || fq == 4) && (fb & (fb - 1)) == 0) {
                bq = fb;
                qg = fq;
            }
        }
        p.b0 = b0;
        int rc = exact ? dispatch_chunks<true>(nchunk, bq, qg, p, idx->num_sms, (cudaStream_t)stream)
                       : dispatch_chunks<false>(nchunk, bq, qg, p, idx->num_sms, (cudaStream_t)stream);
        if (rc != HS_OK) return rc;
        b0 += bq * qg;
    }
    return HS_OK;
}

}  // extern "C"
