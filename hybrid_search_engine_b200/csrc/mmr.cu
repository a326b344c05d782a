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
#include "common.cuh"

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
    // note: max_sim / alive are indexed with stride C inside the kernel; cpad only sizes the buffer
    mmr_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(p);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
