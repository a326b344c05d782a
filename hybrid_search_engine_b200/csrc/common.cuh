// Shared device/host helpers for the hs_b200 kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hs_b200.h"

// ---------------------------------------------------------------- index handle (host side)
struct hs_index {
    int device = 0;
    int num_sms = 148;
    int64_t n_docs = 0;
    int64_t doc_base = 0;
    // dense
    const float* vectors = nullptr;
    const float* vnorm = nullptr;
    int32_t dim = 0;
    int64_t ld = 0;
    // bf16 copy of the dense matrix for the tcgen05 GEMM path: [n_docs, ld_bf16], ld_bf16 % 64 == 0
    const void* v_bf16 = nullptr;
    int64_t ld_bf16 = 0;
    CUtensorMap tmap_a;          // TMA descriptor over v_bf16: box 64 x 128, 128-byte swizzle
    bool has_tmap_a = false;
    CUtensorMap tmap_f32;        // TMA descriptor over the float32 matrix: box 32 x 128 (tf32x3 GEMM path)
    bool has_tmap_f32 = false;
    // csr
    const int64_t* indptr = nullptr;
    const uint2* postings = nullptr;
    int64_t n_terms = 0;
    int64_t n_postings = 0;
    // doc stats
    const uint32_t* dl = nullptr;
    const double* impact_table = nullptr;   // [(max_dl + 1) * (tf_cap + 1)], see hs_bm25_impact_table
    uint32_t max_dl = 0;
    uint32_t tf_cap = 0;
    double avgdl = 0.0, k1 = 1.5, b = 0.75;
    // hot terms: dense float64 contribution vectors [n_hot, n_docs] + term -> row map (hs_bm25_build_hot)
    const double* hot_c = nullptr;
    const int32_t* hot_of_term = nullptr;
    int32_t n_hot = 0;
};

void hs_gemm_attach_f32(hs_index* idx);      // dense_gemm.cu: builds tmap_f32 (called by hs_index_set_dense)

// ---------------------------------------------------------------- error plumbing
void hs_set_error(const char* fmt, ...);

#define HS_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            hs_set_error(__VA_ARGS__);        \
            return HS_ERR_ARG;                \
        }                                     \
    } while (0)

#define HS_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            hs_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return HS_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define HS_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess) {                                                        \
            hs_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return HS_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------- order-preserving encodings
// float -> uint32 such that a < b  <=>  enc(a) < enc(b) (total order, -0 < +0)
__host__ __device__ __forceinline__ uint32_t hs_enc_f32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float hs_dec_f32(uint32_t e) {
    uint32_t u = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
// ranking key: larger = better under (score desc, doc_id asc). +0.0 and -0.0 rank equal.
__device__ __forceinline__ uint64_t hs_make_key(float score, uint32_t doc_id) {
    if (score == 0.0f) score = 0.0f;  // fold -0.0 into +0.0 so ties fall to doc_id like the host sort
    return ((uint64_t)hs_enc_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - doc_id);
}

// stats slots
#define HS_STAT_MIN_A 0
#define HS_STAT_MAX_A 1
#define HS_STAT_MAX_B 2
#define HS_STAT_MIN_B 3

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double hs_shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xFFFFFFFFu, lo, mask);
    hi = __shfl_xor_sync(0xFFFFFFFFu, hi, mask);
    return __hiloint2double(hi, lo);
}
// butterfly 16,8,4,2,1 -- every lane ends with the same bits (a+b == b+a in IEEE)
__device__ __forceinline__ double hs_warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __dadd_rn(v, hs_shfl_xor_f64(v, m));
    return v;
}
__device__ __forceinline__ float hs_warp_sum_f32(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, m));
    return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a kernel needs more than it was last given on
// this device: `cache` is a function-local static of the launching function (one per kernel instantiation),
// so the steady-state hot path makes no attribute call at all
template <typename Kern>
static inline cudaError_t hs_smem_limit(Kern kern, size_t smem, size_t (&cache)[16]) {
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = cache[dev & 15];
    if (smem <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) have = smem;
    return e;
}

// cosine of query q (float32 [dim], |q| = qn) and row v (|v| = vn) in the CONFORMANCE order, one warp per pair:
// element e = 128 c + 4 lane + j accumulated in float64 over (c, j), lanes combined 16, 8, 4, 2, 1, then the
// reference's float32 steps f32(dot) / (f32|q| * f32|v|) with its zero-norm rules (utils.py:44-52).  Bit-identical to
// dense_scan_kernel<EXACT> and to oracle/hybrid_oracle.py:cosine_exact.  All lanes return the value.
__device__ __forceinline__ float hs_exact_cos_warp(const float* __restrict__ q, float qn, const float* __restrict__ v,
                                                   float vn, int dim, int64_t ld, int lane) {
    double acc = 0.0;
    const int nchunk = (dim + 127) / 128;
    for (int c = 0; c < nchunk; ++c) {
        const int e = c * 128 + lane * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < ld) x = *reinterpret_cast<const float4*>(v + e);
        const float q0 = e + 0 < dim ? q[e + 0] : 0.f, q1 = e + 1 < dim ? q[e + 1] : 0.f;
        const float q2 = e + 2 < dim ? q[e + 2] : 0.f, q3 = e + 3 < dim ? q[e + 3] : 0.f;
        acc = __fma_rn((double)x.x, (double)q0, acc);
        acc = __fma_rn((double)x.y, (double)q1, acc);
        acc = __fma_rn((double)x.z, (double)q2, acc);
        acc = __fma_rn((double)x.w, (double)q3, acc);
    }
    const float dot = __double2float_rn(hs_warp_sum_f64(acc));
    return (qn != 0.0f && vn != 0.0f) ? __fdiv_rn(dot, __fmul_rn(qn, vn)) : 0.0f;
}
// f32(sqrt(sum64 q^2)) in the same order (dense_scan_kernel's query norm); all lanes return the value
__device__ __forceinline__ float hs_exact_norm_warp(const float* __restrict__ q, int dim, int lane) {
    double qq = 0.0;
    const int nchunk = (dim + 127) / 128;
    for (int c = 0; c < nchunk; ++c)
        for (int j = 0; j < 4; ++j) {
            const int e = c * 128 + lane * 4 + j;
            const float x = e < dim ? q[e] : 0.f;
            qq = __fma_rn((double)x, (double)x, qq);
        }
    return __double2float_rn(__dsqrt_rn(hs_warp_sum_f64(qq)));
}

static inline int hs_num_sms(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n > 0 ? n : 148;
}
