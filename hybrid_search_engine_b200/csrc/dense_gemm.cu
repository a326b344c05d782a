// K2b -- dense scan at LARGE query batch on the 5th-generation tensor cores (north star: "at large query
// batch it becomes a tcgen05 GEMM"; replaces the same reference site as K2, utils.py:28-54).
//     cos[b, i] = (Q[b,:] . V[i,:]) / (|q_b| |v_i|)          norms in float32 from the float32 vectors
//
// Two operand kinds share one kernel:
//   HS_DENSE_BF16    tcgen05.mma.kind::f16 on a bf16 copy of the matrix (K block = 64 elements); agrees with
//                    the float32 path within 1e-2 (stage-1 retrieval)
//   HS_DENSE_TF32X3  tcgen05.mma.kind::tf32 on the float32 matrix ITSELF (K block = 32 floats), three MMAs per
//                    K step: q_hi*v_hi + q_lo*v_hi + q_hi*v_lo with x_hi = the top 19 bits of x (what the tensor
//                    core reads) and x_lo = x - x_hi (exact).  The corpus tile's lo part is produced in shared
//                    memory by a converter warpgroup while the tile is in flight; float32-grade accuracy
//                    (|cos error| ~ 1e-7), one corpus pass serves 128 queries
// and two epilogues:
//   STORE   cos[b, doc range] float32 + per-query min/max (stats), as K2
//   FILTER  nothing is stored: every (query, doc) score at or above that query's threshold is appended as a
//           ranking key to the query's candidate list (pure-semantic retrieval, multi_stage stage 1)
//
// One persistent CTA per SM, warp specialised:
//   warp 0      TMA producer (cp.async.bulk.tensor, 128-byte swizzle): 128-doc x 128-byte blocks of the corpus
//               through a ring of stages; the query operand is resident in shared memory when it fits, else
//               its K blocks ride in the same stages (served from L2 after the first tile)
//   warp 1      TMEM allocation + single-thread MMA issue: M = 128 QUERIES (TMEM lanes), N = 128 docs (TMEM
//               columns), MT (1 or 2) query tiles per landed corpus block -- one corpus pass serves MT * 128
//               queries; accumulators double-buffered in TMEM (2 * MT * 128 columns), tcgen05.commit releases
//               stages / publishes accumulators on mbarriers
//   warps 4-11  bf16: two epilogue groups (even / odd tiles).  tf32x3: warps 4-7 epilogue, warps 8-11 converter.
//               With queries on the TMEM lanes every epilogue thread owns ONE query per tile: tcgen05.ld 32
//               columns at a time, scale by 1/|q| and the tile's 1/|v_i|, thread-local min/max / threshold test
//
// Rooflines: 2 * B * n * K flop (x3 for tf32x3) against the measured tensor peak; n * K * sizeof(elem) bytes
// per corpus pass (+ B * n * 4 written by STORE) against HBM.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace {

// TMA warp, MMA warp, 2 idle, then bf16: 16 epilogue warps (4 groups); tf32x3: 4 epilogue + 4 converter warps
constexpr int threads_of(int kind) { return kind == 0 ? 640 : 384; }
constexpr int kTileN = 128;           // docs per tile = UMMA N = TMEM columns per accumulator
constexpr int kTileM = 128;           // queries per M tile = UMMA M = TMEM lanes
constexpr int kBlockBytes = 128 * 128;   // one operand K block: 128 rows x 128 bytes (swizzle row) = 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemMax = 227 * 1024;
constexpr int kSmemMisc = 3072;       // barriers, TMEM slot, 1/|v| staging

enum { kKindBf16 = 0, kKindTf32x3 = 1 };
enum { kEpiStore = 0, kEpiFilter = 1 };

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
// the same load MULTICAST to every CTA of the cluster named in `mask`: the tile lands at the same shared-memory offset
// in each of them and each one's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                     uint32_t accumulate) {
    if constexpr (KIND == kKindBf16) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// commit that arrives on the mbarrier at the same offset in EVERY CTA of `mask` (a stage fed by multicast loads may only
// be refilled once all the CTAs that received it have consumed it)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
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
    // STORE epilogue: cos[(b0 + q) * cos_ld + (doc - d0)], float32 -- or binary16 (cos_h, cos == nullptr; cos_ld even)
    float* cos;
    __half* cos_h;
    int64_t cos_ld;
    uint32_t* stats;       // [B, 4] or null
    const float* vnorm;    // [n]
    const float* inv_qn;   // [MT * 128] 1 / |q_b| (0 for a zero query or padding)
    int64_t d0, d1;        // doc range [d0, d1) of this launch (shard-local)
    int nq_valid;          // real queries of this launch (the query operand is MT * 128 zero-padded rows)
    int b0;                // first query index of this launch
    int kb;                // K blocks (64 bf16 / 32 float elements each)
    int stages;
    int q_resident;        // 1: whole query operand stays in smem; 0: its K blocks stream with the corpus blocks
    // FILTER epilogue
    const float* thr;            // [B] keep scores >= thr[b] (null: keep everything)
    // candidates are appended WITHOUT atomics: every (query, CTA, epilogue group) owns one segment of seg_cap keys and
    // its epilogue thread keeps the fill count in a register (an atomic per append stalled the warp ~1 us each)
    unsigned long long* cand;    // [B, n_seg, seg_cap] ranking keys
    unsigned int* cand_cnt;      // [B, n_seg] keys the segment's owner wanted to append (> seg_cap = overflow)
    int seg_cap, n_seg;
    // STORE epilogue, optional (exact verification of an approximate scan, hs_verify_*): every score within eps2 of the
    // thread's running max (min) is appended to the hi (lo) side of the (query, segment) list -- a superset of the docs
    // whose EXACT cosine can be the query's global max (min) when |score - exact| <= eps2 / 2
    unsigned long long* ext;     // [B, n_seg, 2, ext_cap] keys (score, shard-local doc)
    unsigned int* ext_cnt;       // [B, n_seg, 2]
    int ext_cap;
    float eps2;
    uint32_t doc_base;           // global id of shard-local doc 0
};

// CL > 1: the kernel runs in thread-block clusters of CL CTAs and the streamed QUERY blocks -- the same for every CTA --
// are fetched once per cluster: each CTA loads 1/CL of every query block (a slice of 128/CL rows) and TMA-multicasts it
// to all CL CTAs, which divides the L2 -> shared-memory traffic of the query operand by CL (at two query tiles per pass
// that operand is 2/3 of what a stage brings in, and the L2 fabric, not the tensor pipe, bounds the pass).  All CTAs of
// a cluster run the same number of tile iterations (surplus tiles are out-of-range: TMA zero-fills, the epilogue skips).
template <int KIND, int MT, int EPI, int CL>
__global__ void __launch_bounds__(threads_of(KIND), 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_ql, const GemmParams p) {
    constexpr int NQP = (KIND == kKindTf32x3) ? 2 : 1;            // query operand parts (hi, lo)
    constexpr int kElems = (KIND == kKindTf32x3) ? 32 : 64;       // elements per 128-byte swizzle row
    constexpr int V_BYTES = (KIND == kKindTf32x3) ? 2 * kBlockBytes : kBlockBytes;   // raw corpus block (+ its lo part)
    constexpr int Q_BYTES = MT * NQP * kBlockBytes;               // query blocks of one K block
    constexpr int kEpiGroups = (KIND == kKindTf32x3) ? 1 : 4;
    constexpr int kBufs = (KIND == kKindTf32x3 || MT == 2) ? 2 : 4;                    // accumulator buffers in TMEM
    constexpr uint32_t kTmemCols = kBufs * MT * kTileN;           // 256 (tf32x3) or 512 columns

    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char* q_smem = smem;                                 // resident: [kb][mt][part] blocks
    const int q_bytes = p.q_resident ? p.kb * Q_BYTES : 0;
    const int stage_bytes = p.q_resident ? V_BYTES : V_BYTES + Q_BYTES;
    unsigned char* st_smem = smem + q_bytes;                      // [stages][stage_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                                        // [stages] TMA landed
    uint64_t* empty = full + kMaxStages;                          // [stages] MMAs have read the stage
    uint64_t* conv = empty + kMaxStages;                          // [stages] lo part written (tf32x3)
    uint64_t* q_full = conv + kMaxStages;                         // [1]
    uint64_t* tmem_full = q_full + 1;                             // [kBufs <= 4]
    uint64_t* tmem_empty = tmem_full + 4;                         // [kBufs <= 4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
    float* s_inv_vn = reinterpret_cast<float*>(tmem_slot + 2);    // [2][kTileN] 1 / |v_i| (tf32x3)
    float* s_tr = s_inv_vn + 4 * kTileN;                          // STORE: per-warp 32 x 17-word transpose tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (p.d1 - p.d0 + kTileN - 1) / kTileN;
    // tile iterations of this CTA: tile t = blockIdx.x + i * gridDim.x.  In a cluster every CTA runs the same count
    const int64_t n_iter = CL > 1 ? (n_tiles + gridDim.x - 1) / gridDim.x
                                  : (n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    uint32_t crank = 0;
    if constexpr (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    constexpr int kSliceRows = kTileM / CL;                       // rows of a query block this CTA fetches for the cluster

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);          // the MMA commits of every CTA that received the stage's multicast blocks
            mbar_init(&conv[s], 128);          // every converter thread arrives
        }
        mbar_init(q_full, 1);
        for (int a = 0; a < kBufs; ++a) {
            mbar_init(&tmem_full[a], 1);
            // the epilogue warps that drain the buffer: one group, or two (one per query tile) in bf16 with MT = 2
            mbar_init(&tmem_empty[a], (KIND == kKindBf16 && MT == 2) ? 8 : 4);
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
    if constexpr (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anything arrives on them remotely
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            if (p.q_resident) {
                mbar_expect_tx(q_full, (uint32_t)(p.kb * Q_BYTES));
                for (int kb = 0; kb < p.kb; ++kb)
                    for (int mt = 0; mt < MT; ++mt)
                        for (int part = 0; part < NQP; ++part)
                            tma_load_2d(q_smem + (size_t)kb * Q_BYTES + (mt * NQP + part) * kBlockBytes,
                                        part == 0 ? &tmap_q : &tmap_ql, kb * kElems, mt * kTileM, q_full);
            }
            int64_t it = 0;
            for (int64_t i = 0; i < n_iter; ++i) {
                const int64_t t = blockIdx.x + i * gridDim.x;
                const int row0 = (int)(p.d0 + t * kTileN);      // beyond the matrix for a surplus tile: zero fill
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = (int)(it % p.stages);
                    const uint32_t ph = (uint32_t)((it / p.stages) & 1);
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* stg = st_smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], (uint32_t)(kBlockBytes + (p.q_resident ? 0 : Q_BYTES)));
                    tma_load_2d(stg, &tmap_v, kb * kElems, row0, &full[s]);
                    if (!p.q_resident) {    // query K blocks ride along (served from L2 after the first tile)
                        for (int mt = 0; mt < MT; ++mt)
                            for (int part = 0; part < NQP; ++part) {
                                unsigned char* dst = stg + V_BYTES + (mt * NQP + part) * kBlockBytes;
                                if constexpr (CL > 1)       // my row slice of the block, to every CTA of the cluster
                                    tma_load_2d_mc(dst + crank * kSliceRows * 128, part == 0 ? &tmap_q : &tmap_ql, kb * kElems,
                                                   mt * kTileM + (int)crank * kSliceRows, &full[s], kMask);
                                else
                                    tma_load_2d(dst, part == 0 ? &tmap_q : &tmap_ql, kb * kElems, mt * kTileM, &full[s]);
                            }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16 (1) or tf32 (2), both K-major, M = 128 queries,
            // N = 128 docs (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
            constexpr uint32_t fmt = (KIND == kKindTf32x3) ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            if (p.q_resident) mbar_wait(q_full, 0);
            int64_t it = 0;
            for (int64_t tile_i = 0; tile_i < n_iter; ++tile_i) {
                const int buf = (int)(tile_i % kBufs);
                mbar_wait(&tmem_empty[buf], (uint32_t)(((tile_i / kBufs) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = (int)(it % p.stages);
                    const uint32_t ph = (uint32_t)((it / p.stages) & 1);
                    mbar_wait(&full[s], ph);
                    if constexpr (KIND == kKindTf32x3) mbar_wait(&conv[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned char* stg = st_smem + (size_t)s * stage_bytes;
                    const unsigned char* qb = p.q_resident ? q_smem + (size_t)kb * Q_BYTES : stg + V_BYTES;
                    const uint64_t v_desc = umma_desc_sw128(smem_u32(stg));
                    const uint64_t vl_desc = umma_desc_sw128(smem_u32(stg + kBlockBytes));     // tf32x3 only
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((buf * MT + mt) * kTileN);
                        const uint64_t q_desc = umma_desc_sw128(smem_u32(qb + (mt * NQP) * kBlockBytes));
                        const uint64_t ql_desc = umma_desc_sw128(smem_u32(qb + (mt * NQP + NQP - 1) * kBlockBytes));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // one K step = 32 bytes along the swizzled row: +2 in 16-byte descriptor units
                            const uint64_t o = (uint64_t)(k * 2);
                            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                            if constexpr (KIND == kKindTf32x3) {
                                umma<KIND>(d_tmem, ql_desc + o, v_desc + o, idesc, acc);      // q_lo * v_hi
                                umma<KIND>(d_tmem, q_desc + o, vl_desc + o, idesc, 1u);       // q_hi * v_lo
                                umma<KIND>(d_tmem, q_desc + o, v_desc + o, idesc, 1u);        // q_hi * v_hi
                            } else {
                                umma<KIND>(d_tmem, q_desc + o, v_desc + o, idesc, acc);
                            }
                        }
                    }
                    // stage reusable once these MMAs have read it (in a cluster: once EVERY CTA's MMAs have)
                    if constexpr (CL > 1) umma_commit_mc(&empty[s], kMask);
                    else umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);                // accumulators of this tile complete
            }
        }
    } else if (KIND == kKindTf32x3 && warp >= 8) {
        // ------------------------------------------------ converter: lo = x - (top 19 bits of x), element-wise on
        // the landed (swizzled) block into the stage's second buffer -- same addresses, so the swizzle carries over
        const int ctid = threadIdx.x - 256;                 // 0..127
        int64_t it = 0;
        for (int64_t i = 0; i < n_iter; ++i) {
            for (int kb = 0; kb < p.kb; ++kb, ++it) {
                const int s = (int)(it % p.stages);
                const uint32_t ph = (uint32_t)((it / p.stages) & 1);
                mbar_wait(&full[s], ph);
                const float4* raw = reinterpret_cast<const float4*>(st_smem + (size_t)s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(st_smem + (size_t)s * stage_bytes + kBlockBytes);
#pragma unroll
                for (int i = 0; i < kBlockBytes / 16 / 128; ++i) {
                    const float4 x = raw[i * 128 + ctid];
                    float4 y;
                    y.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    y.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    y.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                    y.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                    lo[i * 128 + ctid] = y;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> UMMA reads
                mbar_arrive(&conv[s]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: a group of four warps drains ONE 128-query accumulator
        // tile at a time, thread = query 32 * (warp % 4) + lane of that tile.  bf16 runs FOUR groups (the epilogue, not the
        // tensor pipe or HBM, bounded the pass with two: 2 warps per scheduler cannot hide the TMEM / shared-memory
        // latencies): with one query tile they take every fourth doc tile (four TMEM buffers), with two query tiles groups
        // g and g ^ 1 split the two accumulators of the same doc tile (two TMEM buffers).  tf32x3: one group.
        const int grp = (warp - 4) >> 2;
        const int e = (warp - 4) & 3;                             // TMEM lane quarter = warp id % 4
        const int etid = e * 32 + lane;                           // 0..127 within the group
        const int mtg = (MT == 2) ? (grp & 1) : 0;                // query tile of this group
        const int64_t tile_start = (KIND == kKindTf32x3) ? 0 : (MT == 2 ? (grp >> 1) : grp);
        constexpr int kTileStep = (KIND == kKindTf32x3) ? 1 : (MT == 2 ? 2 : 4);
        const int b = mtg * kTileM + etid;                        // query, relative to b0
        const bool active = b < p.nq_valid;
        const float inv_qn = active ? __ldg(p.inv_qn + b) : 0.f;
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        uint32_t n_app = 0;                                       // FILTER: keys appended to this thread's segment
        uint32_t n_hi = 0, n_lo = 0;                              // STORE + ext: entries of the extreme-candidate lists
        float g_hi = __int_as_float(0xff800000), pub_hi = g_hi;   // ... and the query's max / min over ALL CTAs so far
        float g_lo = __int_as_float(0x7f800000), pub_lo = g_lo;
        const float thr = (EPI == kEpiFilter && active && p.thr != nullptr) ? __ldg(p.thr + p.b0 + b)
                                                                            : __int_as_float(0xff800000);
        const int seg = blockIdx.x * kEpiGroups + grp;            // this group's candidate / extreme-list segment
        const int nrow_raw = p.nq_valid - mtg * kTileM - e * 32;  // live queries of this warp
        const int nrow = nrow_raw < 0 ? 0 : (nrow_raw > 32 ? 32 : nrow_raw);
        float* tr = s_tr + (warp - 4) * (32 * 17);                // per-warp 32 x 16-word transpose tile (padded)
        (void)inv_qn;
        int64_t seq = 0;                                          // tiles this group has drained
        for (int64_t tile_i = tile_start; tile_i < n_iter; tile_i += kTileStep, ++seq) {
            const int64_t t = blockIdx.x + tile_i * gridDim.x;
            const int buf = (int)(tile_i % kBufs);
            const int64_t doc0 = p.d0 + t * kTileN;
            float* inv_vn_tile = s_inv_vn + (int)(seq & 1) * kTileN;   // double-buffered: a warp may run one tile ahead
            if constexpr (KIND != kKindBf16) {
                // 1 / |v_i| of this tile -> smem (zero row -> 0.0, utils.py:49-50).  Not for bf16: its operands are unit
                // vectors, the accumulator is the cosine.
                const int64_t d = doc0 + etid;
                float iv = 0.f;
                if (d < p.d1) {
                    const float vn = __ldg(p.vnorm + d);
                    iv = vn != 0.f ? 1.0f / vn : 0.f;
                }
                inv_vn_tile[etid] = iv;
                asm volatile("bar.sync 1, 128;" ::: "memory");    // the four warps of the (only) group
            }
            mbar_wait(&tmem_full[buf], (uint32_t)((tile_i / kBufs) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int ndoc = (int)((p.d1 - doc0 < kTileN) ? (p.d1 - doc0) : kTileN);
            if constexpr (EPI == kEpiStore) {
                // extreme-candidate lists: the window hangs off the query's max / min over ALL CTAs so far (published through
                // the stats slots once per tile), not off this thread's own running extreme -- a per-thread window triggered
                // in ~4 % of the chunks per thread, i.e. in most chunks per WARP (32 queries), doubling the epilogue
                if (p.ext != nullptr && active) {
                    uint32_t* st = p.stats + (int64_t)(p.b0 + b) * 4;
                    if (mx > pub_hi) {
                        atomicMax(st + HS_STAT_MAX_A, hs_enc_f32(mx));
                        pub_hi = mx;
                    }
                    if (mn < pub_lo) {
                        atomicMin(st + HS_STAT_MIN_A, hs_enc_f32(mn));
                        pub_lo = mn;
                    }
                    // fmaxf / fminf drop the NaN an untouched slot decodes to
                    g_hi = fmaxf(g_hi, hs_dec_f32(__ldcg(st + HS_STAT_MAX_A)));
                    g_lo = fminf(g_lo, hs_dec_f32(__ldcg(st + HS_STAT_MIN_A)));
                }
            }
#pragma unroll 1
            for (int c0 = 0; c0 < kTileN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)((buf * MT + mtg) * kTileN + c0), r);
                if (c0 >= ndoc || nrow == 0) continue;                  // warp-uniform
                // All shared-memory reads of the chunk happen BEFORE any store, and the scores are computed branch-free: a
                // store inside the per-element loop made the compiler keep every later load behind it (possible aliasing
                // through generic pointers), which serialised 256 shared-memory round trips per tile.
                float v[32];
                if constexpr (KIND == kKindBf16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                } else {
                    const float4* ivp = reinterpret_cast<const float4*>(inv_vn_tile + c0);
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float4 iv = ivp[q4];
                        v[4 * q4 + 0] = __uint_as_float(r[4 * q4 + 0]) * iv.x * inv_qn;
                        v[4 * q4 + 1] = __uint_as_float(r[4 * q4 + 1]) * iv.y * inv_qn;
                        v[4 * q4 + 2] = __uint_as_float(r[4 * q4 + 2]) * iv.z * inv_qn;
                        v[4 * q4 + 3] = __uint_as_float(r[4 * q4 + 3]) * iv.w * inv_qn;
                    }
                }
                const int nvalid = ndoc - c0 < 32 ? ndoc - c0 : 32;    // columns of this chunk that are real docs
                // chunk extrema first (a tree of 3-input min / max, no per-element predicates on a full chunk): they
                // feed the query's running min / max, and the FILTER epilogue tests ONE value against the threshold
                float cmn = __int_as_float(0x7f800000), cmx = __int_as_float(0xff800000);
                if (nvalid == 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        cmn = fminf(cmn, v[j]);
                        cmx = fmaxf(cmx, v[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < nvalid) {
                            cmn = fminf(cmn, v[j]);
                            cmx = fmaxf(cmx, v[j]);
                        }
                    }
                }
                if (active) {
                    mn = fminf(mn, cmn);
                    mx = fmaxf(mx, cmx);
                }
                if constexpr (EPI == kEpiStore) {
                    if (p.ext != nullptr && active) {      // warp-uniform on p.ext; both tests are rare after the first tiles
                        unsigned long long* lst = p.ext + ((int64_t)(p.b0 + b) * p.n_seg + seg) * 2 * p.ext_cap;
                        const float ref_hi = fmaxf(mx, g_hi) - p.eps2;      // <= (final global max) - eps2
                        const float ref_lo = fminf(mn, g_lo) + p.eps2;
                        if (cmx >= ref_hi) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (j < nvalid && v[j] >= ref_hi) {
                                    if (n_hi < (uint32_t)p.ext_cap) lst[n_hi] = hs_make_key(v[j], (uint32_t)(doc0 + c0 + j));
                                    ++n_hi;
                                }
                            }
                        }
                        if (cmn <= ref_lo) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (j < nvalid && v[j] <= ref_lo) {
                                    if (n_lo < (uint32_t)p.ext_cap)
                                        lst[p.ext_cap + n_lo] = hs_make_key(v[j], (uint32_t)(doc0 + c0 + j));
                                    ++n_lo;
                                }
                            }
                        }
                    }
                    // transpose through smem so that the global stores run along the docs of ONE query.  Store instruction
                    // i covers rows i and i + 16 of the tile (272 words apart = the other half of the banks; row stride 17
                    // words: conflict-free writes), 64 contiguous bytes of each.
                    // (Storing straight from the registers -- 16 bytes per ROW and instruction, no transpose -- was
                    // measured: 3.7x slower, every store instruction touches 32 different lines.)
                    const int sub = lane & 15, rsel = lane >> 4;
                    if (p.cos_h != nullptr) {
                        // binary16 screen scores, packed BEFORE the transpose: 16 half2 words per query row, one half2 =
                        // two docs per lane
                        uint32_t* trh = reinterpret_cast<uint32_t*>(tr);
#pragma unroll
                        for (int w = 0; w < 16; ++w) {
                            const __half2 h = __floats2half2_rn(v[2 * w], v[2 * w + 1]);
                            trh[lane * 17 + w] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        __syncwarp();
                        const int dcol = c0 + 2 * sub;
                        __half* dst = p.cos_h + (int64_t)(p.b0 + mtg * kTileM + e * 32 + 16 * rsel) * p.cos_ld + (doc0 - p.d0) + dcol;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            if (q + 16 * rsel < nrow) {
                                const uint32_t w = trh[(q + 16 * rsel) * 17 + sub];
                                if (dcol + 1 < ndoc)
                                    *reinterpret_cast<uint32_t*>(dst + (int64_t)q * p.cos_ld) = w;
                                else if (dcol < ndoc)
                                    dst[(int64_t)q * p.cos_ld] = __ushort_as_half((unsigned short)(w & 0xFFFFu));
                            }
                        }
                        __syncwarp();
                    } else {
                        // float32: the 32 columns go through the 16-word tile in two halves
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) tr[lane * 17 + j] = v[16 * h + j];
                            __syncwarp();
                            const int dcol = c0 + 16 * h + sub;
                            if (dcol < ndoc) {
                                float* dst = p.cos + (int64_t)(p.b0 + mtg * kTileM + e * 32 + 16 * rsel) * p.cos_ld + (doc0 - p.d0) + dcol;
#pragma unroll
                                for (int q = 0; q < 16; ++q)
                                    if (q + 16 * rsel < nrow) dst[(int64_t)q * p.cos_ld] = tr[(q + 16 * rsel) * 17 + sub];
                            }
                            __syncwarp();
                        }
                    }
                } else {
                    if (active && cmx >= thr) {        // rare: some column reaches the query's starting bound
                        unsigned long long* sg = p.cand + ((int64_t)(p.b0 + b) * p.n_seg + seg) * p.seg_cap;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nvalid && v[j] >= thr) {
                                if (n_app < (uint32_t)p.seg_cap) sg[n_app] = hs_make_key(v[j], p.doc_base + (uint32_t)(doc0 + c0 + j));
                                ++n_app;
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        if (active) {
            if constexpr (EPI == kEpiFilter) {
                p.cand_cnt[(int64_t)(p.b0 + b) * p.n_seg + seg] = n_app;
            } else if (p.ext != nullptr) {
                unsigned int* c = p.ext_cnt + ((int64_t)(p.b0 + b) * p.n_seg + seg) * 2;
                c[0] = n_hi;
                c[1] = n_lo;
            }
            if (p.stats != nullptr && mn <= mx) {
                atomicMin(&p.stats[(p.b0 + b) * 4 + HS_STAT_MIN_A], hs_enc_f32(mn));
                atomicMax(&p.stats[(p.b0 + b) * 4 + HS_STAT_MAX_A], hs_enc_f32(mx));
            }
        }
    }
    // ------------------------------------------------ teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it / arrive on it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// queries float32 [B, ld_q] -> operand rows [nq_pad, kpad] (zero padded) + 1/|q| from the float32 vector.
//   bf16:   out_hi = bf16(x)
//   tf32x3: out_hi = x (the tensor core reads its top 19 bits), out_lo = x - top19(x)  (float32, exact)
template <int KIND>
__global__ void gemm_prepare_queries_kernel(const float* __restrict__ q, int64_t ld_q, int dim, int b0, int nq_valid,
                                            int kpad, void* __restrict__ out_hi, float* __restrict__ out_lo,
                                            float* __restrict__ inv_qn) {
    const int b = blockIdx.x;       // one warp-sized block per padded query row
    const int lane = threadIdx.x;
    double qq = 0.0;
    for (int e = lane; e < kpad; e += 32) {
        float x = 0.f;
        if (b < nq_valid && e < dim) x = q[(int64_t)(b0 + b) * ld_q + e];
        if constexpr (KIND != kKindBf16) {
            reinterpret_cast<float*>(out_hi)[(int64_t)b * kpad + e] = x;
            out_lo[(int64_t)b * kpad + e] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        }
        qq += (double)x * (double)x;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) qq += hs_shfl_xor_f64(qq, m);
    const float qn = (float)sqrt(qq);
    const float inv = (b < nq_valid && qn != 0.f) ? 1.0f / qn : 0.f;   // zero query -> zeros (utils.py:44-45)
    if constexpr (KIND == kKindBf16) {
        // bf16: BOTH operands are unit vectors (the corpus rows are normalised when the bf16 matrix is built), so the
        // accumulator IS the cosine and the epilogue has no scaling to do
        for (int e = lane; e < kpad; e += 32) {
            float x = 0.f;
            if (b < nq_valid && e < dim) x = q[(int64_t)(b0 + b) * ld_q + e];
            reinterpret_cast<__nv_bfloat16*>(out_hi)[(int64_t)b * kpad + e] = __float2bfloat16_rn(x * inv);
        }
    }
    if (lane == 0) inv_qn[b] = inv;
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

// row-major [rows, ld] of bf16 (elem_bytes 2) or float32 (4) -> TMA map with a 128-byte x 128-row box, 128-byte
// swizzle, zero fill out of bounds (K tail beyond ld, rows beyond the matrix)
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t ld, int elem_bytes, int box_rows = 128) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        hs_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return HS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hs_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return HS_ERR_CUDA;
    }
    return HS_OK;
}

struct Plan {
    int kb, q_resident, stages;
    size_t smem;
};

// shared-memory plan: query operand resident when that still leaves >= 3 stages, else streamed with the corpus blocks
bool make_plan(int kind, int mt, int epi, int64_t ld_elems, Plan& pl) {
    const int nqp = kind == kKindTf32x3 ? 2 : 1;
    const int elems = kind == kKindTf32x3 ? 32 : 64;
    const int v_bytes = kind == kKindTf32x3 ? 2 * kBlockBytes : kBlockBytes;
    const int q_blk = mt * nqp * kBlockBytes;
    const int groups = kind == kKindTf32x3 ? 1 : 4;
    const int tr = epi == kEpiStore ? groups * 4 * 32 * 17 * 4 : 0;
    pl.kb = (int)((ld_elems + elems - 1) / elems);
    const int64_t avail = kSmemMax - 1024 - kSmemMisc - tr;
    const int64_t q_all = (int64_t)pl.kb * q_blk;
    pl.q_resident = (q_all + 3 * (int64_t)v_bytes <= avail) ? 1 : 0;
    const int stage = pl.q_resident ? v_bytes : v_bytes + q_blk;
    int64_t stages = (avail - (pl.q_resident ? q_all : 0)) / stage;
    if (stages > kMaxStages) stages = kMaxStages;
    static const int cap = getenv("HS_GEMM_MAX_STAGES") ? atoi(getenv("HS_GEMM_MAX_STAGES")) : 0;   // experiment switch
    if (cap >= 2 && stages > cap) stages = cap;
    pl.stages = (int)stages;
    pl.smem = 1024 + (size_t)(pl.q_resident ? q_all : 0) + (size_t)pl.stages * stage + kSmemMisc + tr;
    return stages >= 2;
}

template <int KIND, int MT, int EPI, int CL>
int launch_gemm(const CUtensorMap& tv, const CUtensorMap& tq, const CUtensorMap& tql, const GemmParams& p, size_t smem,
                int num_sms, cudaStream_t st) {
    auto kern = dense_gemm_kernel<KIND, MT, EPI, CL>;
    static size_t smem_set[16] = {0};
    HS_CUDA(hs_smem_limit(kern, smem, smem_set));
    const int64_t n_tiles = (p.d1 - p.d0 + kTileN - 1) / kTileN;
    int grid = (int)(n_tiles < num_sms ? n_tiles : num_sms);
    if (CL == 1) {
        kern<<<grid, threads_of(KIND), smem, st>>>(tv, tq, tql, p);
        HS_LAUNCH_CHECK();
        return HS_OK;
    }
    grid = grid / CL * CL;                      // whole clusters only (callers pick CL > 1 for n_tiles >= num_sms)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(threads_of(KIND), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HS_CUDA(cudaLaunchKernelEx(&cfg, kern, tv, tq, tql, p));
    return HS_OK;
}

int dispatch_gemm(int kind, int mt, int epi, int cl, const CUtensorMap& tv, const CUtensorMap& tq, const CUtensorMap& tql,
                  const GemmParams& p, size_t smem, int num_sms, cudaStream_t st) {
#define HS_GEMM_CASE(K, M, E, C) \
    if (kind == K && mt == M && epi == E && cl == C) return launch_gemm<K, M, E, C>(tv, tq, tql, p, smem, num_sms, st)
    HS_GEMM_CASE(kKindBf16, 1, kEpiStore, 1);
    HS_GEMM_CASE(kKindBf16, 2, kEpiStore, 1);
    HS_GEMM_CASE(kKindBf16, 1, kEpiFilter, 1);
    HS_GEMM_CASE(kKindBf16, 2, kEpiFilter, 1);
    HS_GEMM_CASE(kKindTf32x3, 1, kEpiStore, 1);
    HS_GEMM_CASE(kKindTf32x3, 1, kEpiFilter, 1);
    HS_GEMM_CASE(kKindBf16, 1, kEpiStore, 2);
    HS_GEMM_CASE(kKindBf16, 2, kEpiStore, 2);
    HS_GEMM_CASE(kKindBf16, 1, kEpiFilter, 2);
    HS_GEMM_CASE(kKindBf16, 2, kEpiFilter, 2);
    HS_GEMM_CASE(kKindTf32x3, 1, kEpiStore, 2);
    HS_GEMM_CASE(kKindTf32x3, 1, kEpiFilter, 2);
#undef HS_GEMM_CASE
    hs_set_error("dense_gemm: internal: no kernel for kind=%d mt=%d epi=%d cluster=%d", kind, mt, epi, cl);
    return HS_ERR_ARG;
}

// cluster size of the query-block multicast: only when the query operand streams (not resident) and the launch fills the
// GPU; bf16 defaults to 2, tf32x3 (shared-memory bound, not L2 bound) to 1.  HS_GEMM_CLUSTER=1|2 overrides (A/B runs).
int pick_cluster(int kind, int mt, const Plan& pl, int64_t n_tiles, int num_sms) {
    static const int forced = [] {
        const char* e = getenv("HS_GEMM_CLUSTER");
        return e != nullptr ? atoi(e) : 0;
    }();
    if (pl.q_resident || n_tiles < 4 * (int64_t)num_sms || (num_sms % 4) != 0) return 1;
    int cl = kind == kKindBf16 ? 2 : 1;
    if (forced == 1 || forced == 2) cl = forced;      // 4 was measured and dropped: it strands 16 of the 148 SMs (1.8x slower)
    return cl;
}

int gemm_run(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t d0, int64_t d1,
             void* workspace, size_t workspace_bytes, int epi, float* cos, int64_t cos_ld, const float* thr,
             uint64_t* cand, int32_t seg_cap, uint32_t* cand_cnt, uint32_t* stats_enc, void* stream, const char* who,
             uint64_t* ext = nullptr, uint32_t* ext_cnt = nullptr, int32_t ext_cap = 0, float eps2 = 0.f,
             uint16_t* cos_h = nullptr) {
    HS_REQUIRE(idx != nullptr, "%s: idx is null", who);
    HS_REQUIRE(mode == HS_DENSE_BF16 || mode == HS_DENSE_TF32X3, "%s: mode %d is not a tensor-core mode", who, mode);
    HS_REQUIRE(d0 >= 0 && d0 <= d1 && d1 <= idx->n_docs, "%s: doc range [%lld, %lld) outside the shard", who,
               (long long)d0, (long long)d1);
    if (d1 == d0 || B == 0) return HS_OK;
    const int kind = mode == HS_DENSE_BF16 ? kKindBf16 : kKindTf32x3;
    if (kind == kKindBf16 ? !idx->has_tmap_a : !idx->has_tmap_f32) {
        hs_set_error(kind == kKindBf16 ? "%s: index has no bf16 matrix (call hs_index_set_dense_bf16)"
                                       : "%s: index has no float32 TMA map (hs_index_set_dense failed to build one)", who);
        return HS_ERR_STATE;
    }
    HS_REQUIRE(queries != nullptr && B > 0 && ld_q >= idx->dim, "%s: bad arguments", who);
    HS_REQUIRE(workspace != nullptr && workspace_bytes >= hs_dense_gemm_workspace_bytes(idx, B, mode) &&
                   ((uintptr_t)workspace & 255) == 0,
               "%s: workspace too small or not 256-byte aligned", who);
    cudaStream_t st = (cudaStream_t)stream;
    const int elem = kind == kKindBf16 ? 2 : 4;
    const int64_t kpad = kind == kKindBf16 ? idx->ld_bf16 : (idx->ld + 31) / 32 * 32;
    const int max_mt = kind == kKindBf16 ? 2 : 1;
    const size_t op_bytes = (size_t)max_mt * kTileM * kpad * elem;      // one query operand part
    unsigned char* q_hi = (unsigned char*)workspace;
    unsigned char* q_lo = q_hi + op_bytes;                               // tf32x3 only
    float* inv_qn = (float*)(q_hi + (kind == kKindBf16 ? 1 : 2) * op_bytes);
    const CUtensorMap& tv = kind == kKindBf16 ? idx->tmap_a : idx->tmap_f32;
    for (int b0 = 0; b0 < B; b0 += max_mt * kTileM) {
        const int nq_valid = (B - b0 < max_mt * kTileM) ? (B - b0) : max_mt * kTileM;
        const int mt = nq_valid > kTileM ? 2 : 1;
        Plan pl;
        HS_REQUIRE(make_plan(kind, mt, epi, kind == kKindBf16 ? idx->ld_bf16 : idx->ld, pl),
                   "%s: not enough shared memory for dim %d", who, idx->dim);
        if (kind == kKindBf16)
            gemm_prepare_queries_kernel<kKindBf16><<<mt * kTileM, 32, 0, st>>>(queries, ld_q, idx->dim, b0, nq_valid,
                                                                               (int)kpad, q_hi, nullptr, inv_qn);
        else
            gemm_prepare_queries_kernel<kKindTf32x3><<<mt * kTileM, 32, 0, st>>>(queries, ld_q, idx->dim, b0, nq_valid,
                                                                                 (int)kpad, q_hi, (float*)q_lo, inv_qn);
        HS_LAUNCH_CHECK();
        const int cl = pick_cluster(kind, mt, pl, (d1 - d0 + kTileN - 1) / kTileN, idx->num_sms);
        CUtensorMap tq, tql;
        int rc = make_tmap(&tq, q_hi, mt * kTileM, kpad, elem, kTileM / cl);      // box = the row slice one CTA fetches
        if (rc != HS_OK) return rc;
        tql = tq;
        if (kind == kKindTf32x3) {
            rc = make_tmap(&tql, q_lo, mt * kTileM, kpad, elem, kTileM / cl);
            if (rc != HS_OK) return rc;
        }
        GemmParams p;
        p.cos = cos;
        p.cos_h = (__half*)cos_h;
        p.cos_ld = cos_ld;
        p.stats = stats_enc;
        p.vnorm = idx->vnorm;
        p.inv_qn = inv_qn;
        p.d0 = d0;
        p.d1 = d1;
        p.nq_valid = nq_valid;
        p.b0 = b0;
        p.kb = pl.kb;
        p.stages = pl.stages;
        p.q_resident = pl.q_resident;
        p.thr = thr;
        p.cand = (unsigned long long*)cand;
        p.cand_cnt = cand_cnt;
        p.seg_cap = seg_cap;
        p.n_seg = hs_dense_gemm_filter_segments(idx, mode);
        p.ext = (unsigned long long*)ext;
        p.ext_cnt = ext_cnt;
        p.ext_cap = ext_cap;
        p.eps2 = eps2;
        p.doc_base = (uint32_t)idx->doc_base;
        rc = dispatch_gemm(kind, mt, epi, cl, tv, tq, tql, p, pl.smem, idx->num_sms, st);
        if (rc != HS_OK) return rc;
    }
    return HS_OK;
}

}  // namespace

// float32 TMA map over the dense matrix for the tf32x3 path (called by hs_index_set_dense; failure is not an
// error there -- the tensor-core entry points report it when they are used)
void hs_gemm_attach_f32(hs_index* idx) {
    idx->has_tmap_f32 = false;
    if (idx->n_docs > 0 && idx->vectors != nullptr &&
        make_tmap(&idx->tmap_f32, idx->vectors, idx->n_docs, idx->ld, 4) == HS_OK)
        idx->has_tmap_f32 = true;
}

extern "C" {

int hs_index_set_dense_bf16(hs_index* idx, const void* v_bf16, int64_t ld_bf16) {
    HS_REQUIRE(idx != nullptr, "hs_index_set_dense_bf16: idx is null");
    HS_REQUIRE(idx->vectors != nullptr || idx->n_docs == 0, "hs_index_set_dense_bf16: call hs_index_set_dense first");
    HS_REQUIRE(ld_bf16 >= idx->dim && (ld_bf16 % 64) == 0, "hs_index_set_dense_bf16: ld %lld must be a multiple of 64 >= dim",
               (long long)ld_bf16);
    HS_REQUIRE(v_bf16 != nullptr || idx->n_docs == 0, "hs_index_set_dense_bf16: null matrix");
    HS_REQUIRE(((uintptr_t)v_bf16 & 15) == 0, "hs_index_set_dense_bf16: matrix must be 16-byte aligned");
    idx->v_bf16 = v_bf16;
    idx->ld_bf16 = ld_bf16;
    idx->has_tmap_a = false;
    if (idx->n_docs > 0) {
        int rc = make_tmap(&idx->tmap_a, v_bf16, idx->n_docs, ld_bf16, 2);
        if (rc != HS_OK) return rc;
        idx->has_tmap_a = true;
    }
    return HS_OK;
}

size_t hs_dense_gemm_workspace_bytes(const hs_index* idx, int32_t B, int32_t mode) {
    if (idx == nullptr || B <= 0) return 0;
    // query operand rows [256, kpad] bf16, or [128, kpad] float32 hi + lo, then 1/|q|; 256-byte aligned pieces
    if (mode == HS_DENSE_BF16) return idx->ld_bf16 == 0 ? 0 : (size_t)2 * kTileM * idx->ld_bf16 * 2 + 2048;
    if (mode == HS_DENSE_TF32X3) return (size_t)2 * kTileM * ((idx->ld + 31) / 32 * 32) * 4 + 2048;
    return 0;
}

int hs_dense_gemm(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                  int64_t doc_hi, void* workspace, size_t workspace_bytes, float* cos, int64_t cos_ld,
                  uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(cos != nullptr && cos_ld >= doc_hi - doc_lo, "hs_dense_gemm: cos is null or cos_ld < doc range");
    return gemm_run(idx, queries, B, ld_q, mode, doc_lo, doc_hi, workspace, workspace_bytes, kEpiStore, cos, cos_ld,
                    nullptr, nullptr, 0, nullptr, stats_enc, stream, "hs_dense_gemm");
}

int hs_dense_gemm_ext(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                      int64_t doc_hi, void* workspace, size_t workspace_bytes, float* cos, int64_t cos_ld,
                      uint32_t* stats_enc, uint64_t* ext, uint32_t* ext_cnt, int32_t ext_cap, double eps, void* stream) {
    HS_REQUIRE(cos != nullptr && cos_ld >= doc_hi - doc_lo, "hs_dense_gemm_ext: cos is null or cos_ld < doc range");
    HS_REQUIRE(ext != nullptr && ext_cnt != nullptr && ext_cap > 0 && eps >= 0.0 && stats_enc != nullptr,
               "hs_dense_gemm_ext: bad extreme-candidate buffers");
    return gemm_run(idx, queries, B, ld_q, mode, doc_lo, doc_hi, workspace, workspace_bytes, kEpiStore, cos, cos_ld,
                    nullptr, nullptr, 0, nullptr, stats_enc, stream, "hs_dense_gemm_ext", ext, ext_cnt, ext_cap,
                    (float)(2.0 * eps));
}

int hs_dense_gemm_ext_f16(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                          int64_t doc_hi, void* workspace, size_t workspace_bytes, uint16_t* cos_f16, int64_t cos_ld,
                          uint32_t* stats_enc, uint64_t* ext, uint32_t* ext_cnt, int32_t ext_cap, double eps, void* stream) {
    HS_REQUIRE(cos_f16 != nullptr && cos_ld >= doc_hi - doc_lo && (cos_ld & 7) == 0 && ((uintptr_t)cos_f16 & 15) == 0,
               "hs_dense_gemm_ext_f16: cos_f16 is null / misaligned, or cos_ld < doc range or not a multiple of 8");
    HS_REQUIRE(ext != nullptr && ext_cnt != nullptr && ext_cap > 0 && eps >= 0.0 && stats_enc != nullptr,
               "hs_dense_gemm_ext_f16: bad extreme-candidate buffers");
    return gemm_run(idx, queries, B, ld_q, mode, doc_lo, doc_hi, workspace, workspace_bytes, kEpiStore, nullptr, cos_ld,
                    nullptr, nullptr, 0, nullptr, stats_enc, stream, "hs_dense_gemm_ext_f16", ext, ext_cnt, ext_cap,
                    (float)(2.0 * eps), cos_f16);
}

int32_t hs_dense_gemm_filter_segments(const hs_index* idx, int32_t mode) {
    if (idx == nullptr) return 0;
    return idx->num_sms * (mode == HS_DENSE_TF32X3 ? 1 : 4);      // one per CTA and epilogue group
}

int hs_dense_gemm_filter(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode,
                         int64_t doc_lo, int64_t doc_hi, void* workspace, size_t workspace_bytes, const float* thr,
                         uint64_t* cand, int32_t seg_cap, uint32_t* cand_cnt, uint32_t* stats_enc, void* stream) {
    HS_REQUIRE(cand != nullptr && cand_cnt != nullptr && seg_cap > 0, "hs_dense_gemm_filter: bad candidate buffers");
    return gemm_run(idx, queries, B, ld_q, mode, doc_lo, doc_hi, workspace, workspace_bytes, kEpiFilter, nullptr, 0, thr,
                    cand, seg_cap, cand_cnt, stats_enc, stream, "hs_dense_gemm_filter");
}

}  // extern "C"
