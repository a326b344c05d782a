// K2 -- dense query x doc cosine scan (replaces utils.py:28-54 batch_cosine_sim as called from
// core.py:170-176,252).  HBM-bound streaming kernel for small query batches:
//
//   * persistent CTAs (one per SM), a producer warp feeds a 4-stage shared-memory ring with TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx) of R contiguous rows per stage
//   * 8 consumer warps, one doc row per warp at a time: lane l owns elements 128c + 4l + j of the row
//     (conflict-free LDS.128), BQ query vectors live in registers, so a row is read from HBM once for
//     all BQ queries of the launch
//   * reduction in the "conformance order" (lane-sequential over (c, j), butterfly 16..1):
//       EXACT: products/sums in float64 -> bit-identical to oracle/hybrid_oracle.py:cosine_exact
//       FP32 : float32 FMA, same order  -> as precise as the reference's float32 BLAS dot
//   * epilogue fused: cos = f32(dot) / (f32|q| * f32|v|) with the reference's zero-norm rules
//     (utils.py:44-50), per-query min/max folded into the stats slots (utils.py:67-68)
//
// Algorithmic bytes per launch: n_docs * ld * 4 (the corpus pass), independent of BQ.
#include "common.cuh"

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kStages = 4;

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
    int32_t dim, b0, rows_per_stage;
};

template <bool EXACT>
struct Acc {
    using type = float;
};
template <>
struct Acc<true> {
    using type = double;
};

template <int NCHUNK, int BQ, bool EXACT>
__global__ void __launch_bounds__(kThreads, 1) dense_scan_kernel(const DenseParams p) {
    using acc_t = typename Acc<EXACT>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int R = p.rows_per_stage;
    const size_t stage_floats = (size_t)R * p.ld;
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * stage_floats * sizeof(float));
    uint64_t* empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + R - 1) / R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ---------------- producer warp: one elected lane issues the bulk copies
        if (lane == 0) {
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = (int)(it % kStages);
                const uint32_t ph = (uint32_t)((it / kStages) & 1);
                mbar_wait(&empty[s], ph ^ 1u);
                const int64_t row0 = t * R;
                const int64_t rows = (p.n - row0 < R) ? (p.n - row0) : R;
                const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(float));
                mbar_expect_tx(&full[s], bytes);
                tma_bulk_g2s(stages + s * stage_floats, p.v + row0 * p.ld, bytes, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumer warps
    // query registers: element e = 128c + 4*lane + j (zero beyond dim)
    acc_t q[BQ][NCHUNK][4];
    float qn[BQ];
#pragma unroll
    for (int b = 0; b < BQ; ++b) {
        const float* qb = p.q + (int64_t)(p.b0 + b) * p.ld_q;
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

    float my_min = __int_as_float(0x7f800000), my_max = __int_as_float(0xff800000);  // lane b tracks query b
    bool any = false;

    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = (int)(it % kStages);
        const uint32_t ph = (uint32_t)((it / kStages) & 1);
        mbar_wait(&full[s], ph);
        const int64_t row0 = t * R;
        const int rows = (int)((p.n - row0 < R) ? (p.n - row0) : R);
        const float* tile = stages + s * stage_floats;
        for (int r = warp; r < rows; r += kConsumerWarps) {
            const float* row = tile + (size_t)r * p.ld;
            acc_t acc[BQ];
#pragma unroll
            for (int b = 0; b < BQ; ++b) acc[b] = (acc_t)0;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int e = c * 128 + lane * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < p.ld) v = *reinterpret_cast<const float4*>(row + e);
#pragma unroll
                for (int b = 0; b < BQ; ++b) {
                    if (EXACT) {
                        acc[b] = __fma_rn((double)v.x, q[b][c][0], acc[b]);
                        acc[b] = __fma_rn((double)v.y, q[b][c][1], acc[b]);
                        acc[b] = __fma_rn((double)v.z, q[b][c][2], acc[b]);
                        acc[b] = __fma_rn((double)v.w, q[b][c][3], acc[b]);
                    } else {
                        acc[b] = __fmaf_rn(v.x, q[b][c][0], acc[b]);
                        acc[b] = __fmaf_rn(v.y, q[b][c][1], acc[b]);
                        acc[b] = __fmaf_rn(v.z, q[b][c][2], acc[b]);
                        acc[b] = __fmaf_rn(v.w, q[b][c][3], acc[b]);
                    }
                }
            }
            const float vn = p.vnorm[row0 + r];
            float mine = 0.0f;
#pragma unroll
            for (int b = 0; b < BQ; ++b) {
                float dot;
                if (EXACT)
                    dot = __double2float_rn(hs_warp_sum_f64(acc[b]));
                else
                    dot = hs_warp_sum_f32(acc[b]);
                // utils.py:44-52: zero query or zero row -> 0.0, else dot / (|q| * |v|) in float32
                float c = 0.0f;
                if (qn[b] != 0.0f && vn != 0.0f) c = __fdiv_rn(dot, __fmul_rn(qn[b], vn));
                if (lane == b) mine = c;
            }
            if (lane < BQ) {
                p.cos[(int64_t)(p.b0 + lane) * p.n + row0 + r] = mine;
                my_min = fminf(my_min, mine);
                my_max = fmaxf(my_max, mine);
                any = true;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (p.stats != nullptr && lane < BQ && any) {
        atomicMin(&p.stats[(p.b0 + lane) * 4 + HS_STAT_MIN_A], hs_enc_f32(my_min));
        atomicMax(&p.stats[(p.b0 + lane) * 4 + HS_STAT_MAX_A], hs_enc_f32(my_max));
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

template <int NCHUNK, int BQ, bool EXACT>
int launch_dense(const DenseParams& p, int num_sms, cudaStream_t st) {
    const size_t smem = (size_t)kStages * p.rows_per_stage * p.ld * sizeof(float) + 2 * kStages * sizeof(uint64_t);
    auto kern = dense_scan_kernel<NCHUNK, BQ, EXACT>;
    HS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (p.n + p.rows_per_stage - 1) / p.rows_per_stage;
    const int grid = (int)((n_tiles < num_sms) ? n_tiles : num_sms);
    kern<<<grid, kThreads, smem, st>>>(p);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

template <int NCHUNK, bool EXACT>
int dispatch_bq(int bq, const DenseParams& p, int num_sms, cudaStream_t st) {
    constexpr int kMaxBQ = EXACT ? (12 / NCHUNK >= 4 ? 4 : (12 / NCHUNK >= 2 ? 2 : 1))
                                 : (24 / NCHUNK >= 8 ? 8 : (24 / NCHUNK >= 4 ? 4 : 2));
    if constexpr (kMaxBQ >= 8) {
        if (bq == 8) return launch_dense<NCHUNK, 8, EXACT>(p, num_sms, st);
    }
    if constexpr (kMaxBQ >= 4) {
        if (bq == 4) return launch_dense<NCHUNK, 4, EXACT>(p, num_sms, st);
    }
    if constexpr (kMaxBQ >= 2) {
        if (bq == 2) return launch_dense<NCHUNK, 2, EXACT>(p, num_sms, st);
    }
    if (bq == 1) return launch_dense<NCHUNK, 1, EXACT>(p, num_sms, st);
    hs_set_error("dense_scan: internal: no kernel for BQ=%d NCHUNK=%d", bq, NCHUNK);
    return HS_ERR_ARG;
}

template <bool EXACT>
int max_bq(int nchunk) {
    const int budget = EXACT ? 12 : 24;
    int m = budget / nchunk;
    int cap = EXACT ? 4 : 8;
    int bq = 1;
    while (bq * 2 <= m && bq * 2 <= cap) bq *= 2;
    return bq;
}

template <bool EXACT>
int dispatch_chunks(int nchunk, int bq, const DenseParams& p, int num_sms, cudaStream_t st) {
    switch (nchunk) {
        case 1: return dispatch_bq<1, EXACT>(bq, p, num_sms, st);
        case 2: return dispatch_bq<2, EXACT>(bq, p, num_sms, st);
        case 3: return dispatch_bq<3, EXACT>(bq, p, num_sms, st);
        case 4: return dispatch_bq<4, EXACT>(bq, p, num_sms, st);
        case 5:
        case 6: return dispatch_bq<6, EXACT>(bq, p, num_sms, st);
        default: return dispatch_bq<8, EXACT>(bq, p, num_sms, st);
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
    const int nchunk_raw = (idx->dim + 127) / 128;
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
    int rows = (int)(49152 / (idx->ld * sizeof(float)));
    rows = rows / kConsumerWarps * kConsumerWarps;
    if (rows < kConsumerWarps) rows = kConsumerWarps;
    if (rows > 256) rows = 256;
    p.rows_per_stage = rows;
    const bool exact = mode == HS_DENSE_EXACT;
    const int cap = exact ? max_bq<true>(nchunk) : max_bq<false>(nchunk);
    int b0 = 0;
    while (b0 < B) {
        int bq = cap;
        while (bq > B - b0) bq >>= 1;
        p.b0 = b0;
        int rc = exact ? dispatch_chunks<true>(nchunk, bq, p, idx->num_sms, (cudaStream_t)stream)
                       : dispatch_chunks<false>(nchunk, bq, p, idx->num_sms, (cudaStream_t)stream);
        if (rc != HS_OK) return rc;
        b0 += bq;
    }
    return HS_OK;
}

}  // extern "C"
