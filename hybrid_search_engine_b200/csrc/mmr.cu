// K5 -- greedy Maximal Marginal Relevance re-rank (replaces DiversityPipeline._mmr,
// pipelines.py:531-569, and its 27 M scalar cosine_sim calls per query at C=1000, k=250).
//
// One CTA per query.  Instead of recomputing cosine(candidate, selected) for every pair in every
// round, each round computes the cosine of every live candidate against the ONE newly selected row
// (warp per candidate, conformance-order float64 reduction, float32 cosine as utils.py:5-25) and
// keeps a running float64 max_sim per candidate:  k * C * d multiply-adds in total.
//   mmr_i = lambda * rel_i - (1 - lambda) * max_sim_i      (float64, pipelines.py:561)
// first round max_sim = 0 (pipelines.py:558); the winner is the FIRST maximal candidate in
// candidate order (python max(key=), pipelines.py:565).
//
// Two kernels, same arithmetic and picks:
//   mmr_cluster_kernel  a thread-block CLUSTER of 8 CTAs per query; the candidate rows are staged ONCE in the
//                       cluster's distributed shared memory (C / 8 rows per CTA: 1000 x 384 float32 = 8 x 192 KB), every
//                       round each CTA scores its own rows against the pick read through DSMEM and the argmax is
//                       exchanged through DSMEM slots with one cluster barrier per round -- no global gathers at all
//   mmr_kernel          one CTA per query, rows gathered from global memory every round (any C, any dim)
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChunks = 8;

struct MmrParams {
    const float* v;
    const float* vnorm;
    int64_t ld, n_docs;
    int dim;
    const int64_t* cand;   // [B, C]
    const double* rel;     // [B, C]
    double lam, one_minus_lam;
    int C, k;
    double* max_sim;       // [B, C] workspace
    unsigned char* alive;  // [B, C] workspace
    int32_t* out;          // [B, k]
};

__global__ void __launch_bounds__(kThreads) mmr_kernel(const MmrParams p) {
    __shared__ double red_val[kWarps];
    __shared__ int red_idx[kWarps];
    __shared__ int s_best;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t* cand = p.cand + (int64_t)b * p.C;
    const double* rel = p.rel + (int64_t)b * p.C;
    double* ms = p.max_sim + (int64_t)b * p.C;
    unsigned char* alive = p.alive + (int64_t)b * p.C;
    const int nchunk = (p.dim + 127) / 128;

    for (int i = tid; i < p.C; i += kThreads) {
        const int64_t d = cand[i];
        alive[i] = (d >= 0 && d < p.n_docs) ? 1 : 0;
        ms[i] = 0.0;
    }
    __syncthreads();

    for (int it = 0; it < p.k; ++it) {
        // ---- argmax of mmr over live candidates, first maximal wins
        double bv = 0.0;
        int bi = -1;
        for (int i = tid; i < p.C; i += kThreads) {
            if (!alive[i]) continue;
            const double m = __dsub_rn(__dmul_rn(p.lam, rel[i]), __dmul_rn(p.one_minus_lam, ms[i]));
            if (bi < 0 || m > bv) {
                bv = m;
                bi = i;
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double ov = hs_shfl_xor_f64(bv, m);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, m);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            red_val[warp] = bv;
            red_idx[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            double v = red_val[0];
            int ix = red_idx[0];
            for (int w = 1; w < kWarps; ++w) {
                const int oi = red_idx[w];
                const double ov = red_val[w];
                if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) {
                    v = ov;
                    ix = oi;
                }
            }
            s_best = ix;
            p.out[(int64_t)b * p.k + it] = ix;
            if (ix >= 0) alive[ix] = 0;
        }
        __syncthreads();
        const int best = s_best;
        if (best < 0) {
            for (int j = it + 1 + tid; j < p.k; j += kThreads) p.out[(int64_t)b * p.k + j] = -1;
            return;
        }
        if (it + 1 == p.k) return;
        // ---- cosine of every live candidate against the new pick; running max in float64
        const int64_t drow = cand[best];
        const float* srow = p.v + drow * p.ld;
        const float sn = p.vnorm[drow];
        double q[kMaxChunks][4];
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int e = c * 128 + lane * 4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nchunk && e < p.ld) x = *reinterpret_cast<const float4*>(srow + e);
            q[c][0] = x.x; q[c][1] = x.y; q[c][2] = x.z; q[c][3] = x.w;
        }
        for (int i = warp; i < p.C; i += kWarps) {
            if (!alive[i]) continue;
            const int64_t di = cand[i];
            const float* row = p.v + di * p.ld;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
                if (c < nchunk) {
                    const int e = c * 128 + lane * 4;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < p.ld) x = *reinterpret_cast<const float4*>(row + e);
                    acc = __fma_rn((double)x.x, q[c][0], acc);
                    acc = __fma_rn((double)x.y, q[c][1], acc);
                    acc = __fma_rn((double)x.z, q[c][2], acc);
                    acc = __fma_rn((double)x.w, q[c][3], acc);
                }
            }
            const float dot = __double2float_rn(hs_warp_sum_f64(acc));
            if (lane == 0) {
                const float vn = p.vnorm[di];
                float cs = 0.0f;                                  // utils.py:21-23
                if (sn != 0.0f && vn != 0.0f) cs = __fdiv_rn(dot, __fmul_rn(sn, vn));
                const double sim = (double)cs;
                ms[i] = (it == 0) ? sim : fmax(ms[i], sim);      // python max(similarities)
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- cluster kernel (rows in distributed shared memory)
constexpr int kCluster = 8;
constexpr int kClThreads = 512;                    // 16 warps: the per-round work is latency bound, not throughput bound
constexpr int kClWarps = kClThreads / 32;
constexpr int kRowsAtOnce = 4;                     // candidates a warp scores together (independent float64 chains)

struct ArgSlot {
    double val;
    int idx, pad;
};

// (value, position) argmax step shared by every reduction level: larger value wins, ties go to the lower position
__device__ __forceinline__ void arg_better(double& v, int& ix, double ov, int oi) {
    if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) {
        v = ov;
        ix = oi;
    }
}

__global__ void __launch_bounds__(kClThreads, 1) mmr_cluster_kernel(const MmrParams p, int rpc) {
    extern __shared__ __align__(16) unsigned char mmr_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / kCluster;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // pick [ld] | rows [rpc][ld] | rel [rpc] | ms [rpc] | vn [rpc] | alive [rpc]
    float* pick = reinterpret_cast<float*>(mmr_smem);             // local copy of the round's selected row
    float* rows = pick + p.ld;
    double* rel = reinterpret_cast<double*>(rows + (size_t)rpc * p.ld);
    double* ms = rel + rpc;
    float* vn = reinterpret_cast<float*>(ms + rpc);
    unsigned char* alive = reinterpret_cast<unsigned char*>(vn + rpc);
    __shared__ ArgSlot slot[2];                    // this CTA's best of the round (double-buffered by round parity)
    __shared__ double red_val[kClWarps];
    __shared__ int red_idx[kClWarps];
    __shared__ int s_best;

    const int64_t* cand = p.cand + (int64_t)b * p.C;
    const int c0 = rank * rpc;                     // first candidate (position in the candidate list) of this CTA
    const int nloc = (p.C - c0 < rpc) ? (p.C - c0 > 0 ? p.C - c0 : 0) : rpc;
    const int nchunk = (p.dim + 127) / 128;
    const int ld4 = (int)(p.ld / 4);
    // ---- stage this CTA's candidate rows once
    for (int r = warp; r < rpc; r += kClWarps) {
        const int64_t d = r < nloc ? cand[c0 + r] : -1;
        const bool ok = d >= 0 && d < p.n_docs;
        const float4* src = reinterpret_cast<const float4*>(p.v + (ok ? d : 0) * p.ld);
        float4* dst = reinterpret_cast<float4*>(rows + (size_t)r * p.ld);
        for (int e = lane; e < ld4; e += 32) dst[e] = ok ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) {
            alive[r] = ok ? 1 : 0;
            ms[r] = 0.0;
            rel[r] = ok ? p.rel[(int64_t)b * p.C + c0 + r] : 0.0;
            vn[r] = ok ? p.vnorm[d] : 0.f;
        }
    }
    __syncthreads();

    for (int it = 0; it < p.k; ++it) {
        // ---- local argmax over this CTA's live candidates (first maximal = lowest position)
        double bv = 0.0;
        int bi = -1;
        for (int r = tid; r < nloc; r += kClThreads) {
            if (!alive[r]) continue;
            const double m = __dsub_rn(__dmul_rn(p.lam, rel[r]), __dmul_rn(p.one_minus_lam, ms[r]));
            if (bi < 0 || m > bv) {
                bv = m;
                bi = c0 + r;
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) arg_better(bv, bi, hs_shfl_xor_f64(bv, m), __shfl_xor_sync(0xFFFFFFFFu, bi, m));
        if (lane == 0) {
            red_val[warp] = bv;
            red_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            double v = lane < kClWarps ? red_val[lane] : 0.0;
            int ix = lane < kClWarps ? red_idx[lane] : -1;
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) arg_better(v, ix, hs_shfl_xor_f64(v, m), __shfl_xor_sync(0xFFFFFFFFu, ix, m));
            if (lane == 0) {
                slot[it & 1].val = v;
                slot[it & 1].idx = ix;
            }
        }
        cluster.sync();                            // every CTA's slot of this round is visible cluster-wide
        // ---- global argmax: lanes 0..7 of warp 0 read one peer's slot each through DSMEM (one round trip), butterfly;
        //      every CTA reaches the same decision
        if (warp == 0) {
            double v = 0.0;
            int ix = -1;
            if (lane < kCluster) {
                const ArgSlot* s = cluster.map_shared_rank(&slot[it & 1], lane);
                ix = s->idx;
                v = s->val;
            }
#pragma unroll
            for (int m = 4; m >= 1; m >>= 1) arg_better(v, ix, hs_shfl_xor_f64(v, m), __shfl_xor_sync(0xFFFFFFFFu, ix, m));
            if (lane == 0) {
                s_best = ix;
                if (rank == 0) p.out[(int64_t)b * p.k + it] = ix;
                if (ix >= c0 && ix < c0 + nloc) alive[ix - c0] = 0;
            }
        }
        __syncthreads();
        const int best = s_best;
        if (best < 0) {
            if (rank == 0)
                for (int j = it + 1 + tid; j < p.k; j += kClThreads) p.out[(int64_t)b * p.k + j] = -1;
            break;
        }
        if (it + 1 == p.k) break;
        // ---- the pick's row and norm come from its owner's shared memory (DSMEM); then score the local rows,
        //      kRowsAtOnce per warp so that the float64 chains and butterflies of different rows overlap
        // DSMEM moves ~20 B/clk and the OWNER's shared-memory port pays for every reader, so each CTA fetches the row
        // once (one float4 per thread) into its own shared memory and its 16 warps read the local copy
        const int owner = best / rpc, orow = best - owner * rpc;
        {
            const float4* src = reinterpret_cast<const float4*>(cluster.map_shared_rank(rows, owner) + (size_t)orow * p.ld);
            for (int e = tid; e < ld4; e += kClThreads) reinterpret_cast<float4*>(pick)[e] = src[e];
        }
        const float sn = *(cluster.map_shared_rank(vn, owner) + orow);
        __syncthreads();
        const float* srow = pick;
        double q[kMaxChunks][4];
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int e = c * 128 + lane * 4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nchunk && e < p.ld) x = *reinterpret_cast<const float4*>(srow + e);
            q[c][0] = x.x; q[c][1] = x.y; q[c][2] = x.z; q[c][3] = x.w;
        }
        for (int r0 = warp * kRowsAtOnce; r0 < nloc; r0 += kClWarps * kRowsAtOnce) {
            double acc[kRowsAtOnce];
#pragma unroll
            for (int j = 0; j < kRowsAtOnce; ++j) acc[j] = 0.0;
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
                if (c < nchunk) {
                    const int e = c * 128 + lane * 4;
#pragma unroll
                    for (int j = 0; j < kRowsAtOnce; ++j) {
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e < p.ld && r0 + j < nloc) x = *reinterpret_cast<const float4*>(rows + (size_t)(r0 + j) * p.ld + e);
                        acc[j] = __fma_rn((double)x.x, q[c][0], acc[j]);
                        acc[j] = __fma_rn((double)x.y, q[c][1], acc[j]);
                        acc[j] = __fma_rn((double)x.z, q[c][2], acc[j]);
                        acc[j] = __fma_rn((double)x.w, q[c][3], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {             // hs_warp_sum_f64's butterfly, four rows interleaved
#pragma unroll
                for (int j = 0; j < kRowsAtOnce; ++j) acc[j] = __dadd_rn(acc[j], hs_shfl_xor_f64(acc[j], m));
            }
            if (lane < kRowsAtOnce && r0 + lane < nloc && alive[r0 + lane]) {
                double a = acc[0];
#pragma unroll
                for (int j = 1; j < kRowsAtOnce; ++j) a = (lane == j) ? acc[j] : a;
                const int r = r0 + lane;
                const float dot = __double2float_rn(a);
                const float v2 = vn[r];
                float cs = 0.0f;                                  // utils.py:21-23
                if (sn != 0.0f && v2 != 0.0f) cs = __fdiv_rn(dot, __fmul_rn(sn, v2));
                const double sim = (double)cs;
                ms[r] = (it == 0) ? sim : fmax(ms[r], sim);      // python max(similarities)
            }
        }
        __syncthreads();
    }
    cluster.sync();                                // nobody leaves while a peer may still read its rows / slots
}

size_t mmr_cluster_smem(int rpc, int64_t ld) {
    return (size_t)(rpc + 1) * ld * sizeof(float) + (size_t)rpc * (2 * sizeof(double) + sizeof(float) + 1) + 64;
}

}  // namespace

extern "C" {

size_t hs_mmr_workspace_bytes(int32_t B, int32_t C) {
    if (B <= 0 || C <= 0) return 0;
    const size_t cpad = ((size_t)C + 7) / 8 * 8;
    return (size_t)B * cpad * (sizeof(double) + 1);
}

int hs_mmr(const hs_index* idx, const int64_t* cand_ids, const double* rel, double lambda, int32_t B, int32_t C,
           int32_t k, void* workspace, size_t workspace_bytes, int32_t* out_sel, void* stream) {
    HS_REQUIRE(idx != nullptr, "hs_mmr: idx is null");
    if (B == 0 || k == 0) return HS_OK;
    HS_REQUIRE(B > 0 && C > 0 && k > 0, "hs_mmr: bad sizes B=%d C=%d k=%d", B, C, k);
    HS_REQUIRE(cand_ids != nullptr && rel != nullptr && out_sel != nullptr, "hs_mmr: null pointer");
    if (idx->vectors == nullptr) {
        hs_set_error("hs_mmr: index has no dense matrix");
        return HS_ERR_STATE;
    }
    HS_REQUIRE(workspace != nullptr && workspace_bytes >= hs_mmr_workspace_bytes(B, C), "hs_mmr: workspace too small");
    MmrParams p;
    p.v = idx->vectors;
    p.vnorm = idx->vnorm;
    p.ld = idx->ld;
    p.n_docs = idx->n_docs;
    p.dim = idx->dim;
    p.cand = cand_ids;
    p.rel = rel;
    p.lam = lambda;
    p.one_minus_lam = 1 - lambda;     // python: (1 - self.lambda_param)
    p.C = C;
    p.k = k;
    const size_t cpad = ((size_t)C + 7) / 8 * 8;
    p.max_sim = (double*)workspace;
    p.alive = (unsigned char*)workspace + (size_t)B * cpad * sizeof(double);
    p.out = out_sel;
    // cluster path when a CTA's share of the candidate rows fits its shared memory (C = 1000, d = 384: 192 KB)
    static const bool no_cluster = getenv("HS_MMR_NO_CLUSTER") != nullptr;     // A/B switch, read once
    const int rpc = (C + kCluster - 1) / kCluster;
    const size_t smem = mmr_cluster_smem(rpc, idx->ld);
    if (!no_cluster && C >= 64 && smem <= 225 * 1024 && (int64_t)B * kCluster <= 0x7FFFFFFF) {
        static size_t smem_set[16] = {0};
        HS_CUDA(hs_smem_limit(mmr_cluster_kernel, smem, smem_set));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * kCluster), 1, 1);
        cfg.blockDim = dim3(kClThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        HS_CUDA(cudaLaunchKernelEx(&cfg, mmr_cluster_kernel, p, rpc));
        return HS_OK;
    }
    // note: max_sim / alive are indexed with stride C inside the kernel; cpad only sizes the buffer
    mmr_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(p);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
