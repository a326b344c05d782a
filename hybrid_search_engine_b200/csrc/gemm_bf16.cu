// K2b -- dense scan at LARGE query batch as a bf16 tcgen05 GEMM (north star: "at large query batch it
// becomes a tcgen05 GEMM in bf16"; replaces the same reference site as K2, utils.py:28-54, for
// HS_DENSE_BF16).  cos[b, i] = (Q[b,:] . V[i,:]) / (|q_b| |v_i|), dot in bf16 x bf16 -> fp32 on the
// 5th-generation tensor cores, norms in float32 from the original float32 vectors.
//
// One persistent CTA per SM, warp specialised:
//   warp 0      TMA producer: the query block (B operand, all K) once, then 128-doc x 64-k tiles of the
//               bf16 corpus through a ring of 16 KB stages (cp.async.bulk.tensor, 128-byte swizzle)
//   warp 1      TMEM allocation + single-thread MMA issue: tcgen05.mma.cta_group::1.kind::f16,
//               M = 128 QUERIES (rows = TMEM lanes, zero padded), N = 128 docs (TMEM columns), K = 16 per
//               instruction; two accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1;
//               tcgen05.commit releases smem stages / publishes the accumulator on mbarriers
//   warps 4-11  two epilogue groups of four warps (even tiles / odd tiles = accumulator 0 / 1): with queries on the TMEM lanes every thread owns ONE query and reads its 128 doc
//               scores with tcgen05.ld (32 columns at a time): scaling by 1/|q_b| and the tile's 1/|v_i|
//               (staged in smem), the per-query min/max (two registers, no cross-lane traffic) and the
//               16-byte stores of cos[b, doc tile] are all thread-local
//
// Roofline: 2 * N_q * n * K flop against the measured bf16 tensor peak, n * K * 2 bytes read +
// N_q * n * 4 bytes written against HBM; at N_q <= 128 the kernel is HBM-bound.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 384;       // TMA warp, MMA warp, 2 idle, 2 x 4 epilogue warps
constexpr int kTileM = 128;           // docs per tile = UMMA M
constexpr int kBlockK = 64;           // bf16 elements per 128-byte swizzle row
constexpr int kStageBytes = kTileM * kBlockK * 2;
constexpr int kMaxNQ = 128;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 256;   // two accumulators of up to 128 fp32 columns
constexpr int kEpilogueSmem = 8 * 32 * 33 * 4;   // transpose tiles of the eight epilogue warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// 2-D TMA tile load: coordinates (c0 = element along K, c1 = row)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused with swizzle)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
    float* cos;            // [B, n]
    uint32_t* stats;       // [B, 4] or null
    const float* vnorm;    // [n]
    const float* inv_qn;   // [nq] 1 / |q_b| (0 for a zero query or padding)
    int64_t n;
    int nq_valid;          // real queries of this launch (the A operand is always 128 zero-padded rows)
    int b0;                // first query index of this launch
    int kb;                // K blocks of 64
    int stages;
    int q_resident;        // 1: whole query operand stays in smem; 0: its K blocks stream with the corpus tiles
};

__global__ void __launch_bounds__(kThreads, 1)
dense_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_q,
                       const GemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int q_tile_bytes = kMaxNQ * kBlockK * 2;                // one K block of the query operand (A)
    unsigned char* q_smem = smem;                                 // [kb][128 x 128 B]
    const int q_bytes = p.q_resident ? ((p.kb * q_tile_bytes + 1023) & ~1023) : 0;
    // stage = 128 docs x 64 k (16 KB) [+ the matching 128 queries x 64 k block when the queries stream]
    const int stage_bytes = p.q_resident ? kStageBytes : 2 * kStageBytes;
    unsigned char* a_smem = smem + q_bytes;                       // [stages][stage_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                                        // [stages]
    uint64_t* empty = full + kMaxStages;                          // [stages]
    uint64_t* q_full = empty + kMaxStages;                        // [1]
    uint64_t* tmem_full = q_full + 1;                             // [2]
    uint64_t* tmem_empty = tmem_full + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_inv_vn = reinterpret_cast<float*>(tmem_slot + 2);    // [2 groups][2 tiles][kTileM] 1 / |v_i|
    float* s_tr = s_inv_vn + 4 * kTileM;                          // [4][32 x 33] epilogue transpose tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.n + kTileM - 1) / kTileM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(q_full, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);      // the four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            if (p.q_resident) {
                mbar_expect_tx(q_full, (uint32_t)(p.kb * q_tile_bytes));
                for (int kb = 0; kb < p.kb; ++kb)
                    tma_load_2d(q_smem + (size_t)kb * q_tile_bytes, &tmap_q, kb * kBlockK, 0, q_full);
            }
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = (int)(it % p.stages);
                    const uint32_t ph = (uint32_t)((it / p.stages) & 1);
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    tma_load_2d(a_smem + (size_t)s * stage_bytes, &tmap_a, kb * kBlockK, (int)(t * kTileM), &full[s]);
                    if (!p.q_resident)      // query K block rides along (served from L2 after the first tile)
                        tma_load_2d(a_smem + (size_t)s * stage_bytes + kStageBytes, &tmap_q, kb * kBlockK, 0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            // kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128 queries,
            // N = 128 docs
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileM >> 3) << 17) |
                                   ((uint32_t)(kMaxNQ >> 4) << 24);
            if (p.q_resident) mbar_wait(q_full, 0);
            int64_t it = 0, tile_i = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_i) {
                const int acc = (int)(tile_i & 1);
                mbar_wait(&tmem_empty[acc], (uint32_t)(((tile_i >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kMaxNQ);
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = (int)(it % p.stages);
                    const uint32_t ph = (uint32_t)((it / p.stages) & 1);
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned char* q_blk = p.q_resident ? q_smem + (size_t)kb * q_tile_bytes
                                                              : a_smem + (size_t)s * stage_bytes + kStageBytes;
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(q_blk));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(a_smem + (size_t)s * stage_bytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzled row: +2 in 16-byte units
                        umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);                  // stage reusable once these MMAs have read it
                }
                umma_commit(&tmem_full[acc]);                // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: thread = query 32*(warp%4) + lane;
        // group 0 (warps 4-7) drains accumulator 0 (even tiles), group 1 (warps 8-11) accumulator 1
        const int grp = (warp - 4) >> 2;
        const int e = (warp - 4) & 3;                             // TMEM lane quarter = warp id % 4
        const int etid = e * 32 + lane;                           // 0..127 within the epilogue group
        const int b = e * 32 + lane;
        const bool active = b < p.nq_valid;
        const float inv_qn = active ? __ldg(p.inv_qn + b) : 0.f;
        float* tr = s_tr + (warp - 4) * (32 * 33);                // per-warp 32 x 32 transpose tile (padded)
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        int64_t tile_i = grp;
        for (int64_t t = blockIdx.x + (int64_t)grp * gridDim.x; t < n_tiles; t += 2 * (int64_t)gridDim.x, tile_i += 2) {
            const int acc = grp;
            const int64_t doc0 = t * kTileM;
            float* inv_vn_tile;
            {   // 1 / |v_i| of this tile -> smem (zero row -> 0.0, utils.py:49-50)
                const int64_t d = doc0 + etid;
                float iv = 0.f;
                if (d < p.n) {
                    const float vn = __ldg(p.vnorm + d);
                    iv = vn != 0.f ? 1.0f / vn : 0.f;
                }
                inv_vn_tile = s_inv_vn + (grp * 2 + (int)((tile_i >> 1) & 1)) * kTileM;   // double-buffered:
                inv_vn_tile[etid] = iv;              // a warp may run one tile ahead of its group
            }
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // the four warps of this group only
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            mbar_wait(&tmem_full[acc], (uint32_t)((tile_i >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int ndoc = (int)((p.n - doc0 < kTileM) ? (p.n - doc0) : kTileM);
            const int nrow = (p.nq_valid - e * 32 < 32) ? (p.nq_valid - e * 32) : 32;   // live queries of this warp
#pragma unroll 1
            for (int c0 = 0; c0 < kTileM; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * kTileM + c0), r);
                if (c0 < ndoc && nrow > 0) {                       // warp-uniform
                    // thread = query: scale, fold min/max, then transpose through smem so that the global
                    // stores run along the docs of ONE query (128 contiguous bytes per instruction)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]) * inv_vn_tile[c0 + j] * inv_qn;
                        if (active && c0 + j < ndoc) {
                            mn = fminf(mn, v);
                            mx = fmaxf(mx, v);
                        }
                        tr[lane * 33 + j] = v;
                    }
                    __syncwarp();
                    if (c0 + lane < ndoc) {
                        float* dst = p.cos + (int64_t)(p.b0 + e * 32) * p.n + doc0 + c0 + lane;
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (q < nrow) dst[(int64_t)q * p.n] = tr[q * 33 + lane];
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (p.stats != nullptr && active && mn <= mx) {
            atomicMin(&p.stats[(p.b0 + b) * 4 + HS_STAT_MIN_A], hs_enc_f32(mn));
            atomicMax(&p.stats[(p.b0 + b) * 4 + HS_STAT_MAX_A], hs_enc_f32(mx));
        }
    }
    // ------------------------------------------------ teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// queries float32 [B, ld_q] -> bf16 [nq_pad, kpad] (zero padded) + 1/|q| from the float32 vector
__global__ void gemm_prepare_queries_kernel(const float* __restrict__ q, int64_t ld_q, int dim, int b0, int nq_valid,
                                            int nq_pad, int kpad, __nv_bfloat16* __restrict__ out,
                                            float* __restrict__ inv_qn) {
    const int b = blockIdx.x;       // one warp-sized block per padded query row
    const int lane = threadIdx.x;
    double qq = 0.0;
    for (int e = lane; e < kpad; e += 32) {
        float x = 0.f;
        if (b < nq_valid && e < dim) x = q[(int64_t)(b0 + b) * ld_q + e];
        out[(int64_t)b * kpad + e] = __float2bfloat16_rn(x);
        qq += (double)x * (double)x;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) qq += hs_shfl_xor_f64(qq, m);
    if (lane == 0) {
        const float qn = (float)sqrt(qq);
        inv_qn[b] = (b < nq_valid && qn != 0.f) ? 1.0f / qn : 0.f;   // zero query -> zeros (utils.py:44-45)
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 row-major [rows, kpad] -> TMA map with box (64 x box_rows), 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t kpad, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        hs_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return HS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hs_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return HS_ERR_CUDA;
    }
    return HS_OK;
}

}  // namespace

extern "C" {

int hs_index_set_dense_bf16(hs_index* idx, const void* v_bf16, int64_t ld_bf16) {
    HS_REQUIRE(idx != nullptr, "hs_index_set_dense_bf16: idx is null");
    HS_REQUIRE(idx->vectors != nullptr || idx->n_docs == 0, "hs_index_set_dense_bf16: call hs_index_set_dense first");
    HS_REQUIRE(ld_bf16 >= idx->dim && (ld_bf16 % kBlockK) == 0, "hs_index_set_dense_bf16: ld %lld must be a multiple of 64 >= dim",
               (long long)ld_bf16);
    HS_REQUIRE(v_bf16 != nullptr || idx->n_docs == 0, "hs_index_set_dense_bf16: null matrix");
    HS_REQUIRE(((uintptr_t)v_bf16 & 15) == 0, "hs_index_set_dense_bf16: matrix must be 16-byte aligned");
    idx->v_bf16 = v_bf16;
    idx->ld_bf16 = ld_bf16;
    idx->has_tmap_a = false;
    if (idx->n_docs > 0) {
        int rc = make_tmap(&idx->tmap_a, v_bf16, idx->n_docs, ld_bf16, kTileM);
        if (rc != HS_OK) return rc;
        idx->has_tmap_a = true;
    }
    return HS_OK;
}

size_t hs_dense_scan_bf16_workspace_bytes(const hs_index* idx, int32_t B) {
    if (idx == nullptr || idx->ld_bf16 == 0 || B <= 0) return 0;
    // bf16 query block [128, kpad] + 1/|q| [128], 256-byte aligned pieces
    return (size_t)kMaxNQ * idx->ld_bf16 * 2 + 1024;
}

int hs_dense_scan_bf16(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, void* workspace,
                       size_t workspace_bytes, float* cos, uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_dense_scan_bf16: idx is null");
    if (idx->n_docs == 0 || B == 0) return HS_OK;
    if (!idx->has_tmap_a) {
        hs_set_error("hs_dense_scan_bf16: index has no bf16 matrix (call hs_index_set_dense_bf16)");
        return HS_ERR_STATE;
    }
    HS_REQUIRE(queries != nullptr && cos != nullptr && B > 0 && ld_q >= idx->dim, "hs_dense_scan_bf16: bad arguments");
    HS_REQUIRE(workspace != nullptr && workspace_bytes >= hs_dense_scan_bf16_workspace_bytes(idx, B) &&
                   ((uintptr_t)workspace & 255) == 0,
               "hs_dense_scan_bf16: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int kpad = (int)idx->ld_bf16, kb = kpad / kBlockK;
    const int cap = kMaxNQ;
    __nv_bfloat16* q_bf16 = (__nv_bfloat16*)workspace;
    float* inv_qn = (float*)((unsigned char*)workspace + (size_t)kMaxNQ * kpad * 2);
    for (int b0 = 0; b0 < B; b0 += cap) {
        const int nq_valid = (B - b0 < cap) ? (B - b0) : cap;
        gemm_prepare_queries_kernel<<<kMaxNQ, 32, 0, st>>>(queries, ld_q, idx->dim, b0, nq_valid, kMaxNQ, kpad, q_bf16,
                                                           inv_qn);
        HS_LAUNCH_CHECK();
        CUtensorMap tmap_q;
        int rc = make_tmap(&tmap_q, q_bf16, kMaxNQ, kpad, kMaxNQ);
        if (rc != HS_OK) return rc;
        GemmParams p;
        p.cos = cos;
        p.stats = stats_enc;
        p.vnorm = idx->vnorm;
        p.inv_qn = inv_qn;
        p.n = idx->n_docs;
        p.nq_valid = nq_valid;
        p.b0 = b0;
        p.kb = kb;
        // keep the whole query operand in smem when that still leaves >= 3 corpus stages, else stream it
        p.q_resident = (size_t)kb * kMaxNQ * kBlockK * 2 + 3 * kStageBytes + 4096 + kEpilogueSmem <= 220 * 1024 ? 1 : 0;
        const int q_bytes = p.q_resident ? ((kb * kMaxNQ * kBlockK * 2 + 1023) & ~1023) : 0;
        const int stage_bytes = p.q_resident ? kStageBytes : 2 * kStageBytes;
        int stages = (int)((220 * 1024 - q_bytes - 4096 - kEpilogueSmem) / stage_bytes);
        if (stages > kMaxStages) stages = kMaxStages;
        HS_REQUIRE(stages >= 2, "hs_dense_scan_bf16: not enough shared memory for dim %d", idx->dim);
        p.stages = stages;
        const size_t smem = 1024 + (size_t)q_bytes + (size_t)stages * stage_bytes + 3072 + kEpilogueSmem;
        HS_CUDA(cudaFuncSetAttribute(dense_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t n_tiles = (idx->n_docs + kTileM - 1) / kTileM;
        const int grid = (int)(n_tiles < idx->num_sms ? n_tiles : idx->num_sms);
        dense_gemm_bf16_kernel<<<grid, kThreads, smem, st>>>(idx->tmap_a, tmap_q, p);
        HS_LAUNCH_CHECK();
    }
    return HS_OK;
}

}  // extern "C"
