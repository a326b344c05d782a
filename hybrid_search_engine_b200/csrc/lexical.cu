// Fuzzy lexical scorer of the `basic` / `diversity` pipelines (replaces Searcher._lexical_scores,
// core.py:178-197):   0.7 * partial_ratio(query, doc) / 100 + 0.3 * |Q & D| / |Q|   per document.
//
// partial_ratio is third-party code in the reference (rapidfuzz, unpinned, not installable here), so
// this kernel follows the same published definition as the CPU oracle (oracle/hybrid_oracle.py:
// partial_ratio -- PARITY UNPINNED against rapidfuzz itself): the shorter string is the pattern, every
// window of the longer string of at most the pattern's length (partial windows at both ends included)
// is scored (1 - (len_pattern + len_window - 2 * LCS) / (len_pattern + len_window)) * 100, the best window wins.
//
// One CTA per document.  LCS is the bit-parallel Allison-Dix / Hyyro recurrence
//     U = V & PM[c];  V = (V + U) | (V & ~U);   LCS = #zero bits of V in the pattern's m bits
// on W 64-bit words (m <= 64 W), one window per thread at a time; the match masks PM live in shared
// memory (direct table for ASCII, short list for the pattern's other code points).  All floating-point
// steps are float64 in the oracle's order, so the result is bit-identical to the oracle.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxNonAscii = 64;

struct LexParams {
    const uint32_t* doc_chars;   // code points of the lower-cased contents, concatenated
    const int64_t* doc_off;      // [n + 1]
    const uint32_t* q_chars;     // [q_len] lower-cased query
    int q_len;
    const int32_t* doc_tok;      // sorted unique token ids per doc (no stop-word removal), concatenated
    const int64_t* doc_tok_off;  // [n + 1]
    const int32_t* q_tok;        // sorted unique query token ids known to the vocabulary
    int n_q_tok;                 // how many of them
    int q_set_size;              // |Q| = number of distinct query tokens (known or not)
    float* out;                  // [n]
    int* err;                    // set to 1 if a pattern has too many distinct non-ASCII code points
    int64_t n;
};

template <int W>
__device__ __forceinline__ void pm_lookup(uint32_t c, const uint64_t (*pm_ascii)[W], const uint32_t* na_code,
                                          const uint64_t (*na_mask)[W], int n_na, uint64_t (&m)[W]) {
    if (c < 128u) {
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = pm_ascii[c][w];
        return;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) m[w] = 0;
    for (int i = 0; i < n_na; ++i) {
        if (na_code[i] == c) {
#pragma unroll
            for (int w = 0; w < W; ++w) m[w] = na_mask[i][w];
            return;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads) lexical_kernel(const LexParams p) {
    __shared__ uint64_t pm_ascii[128][W];
    __shared__ uint32_t na_code[kMaxNonAscii];
    __shared__ uint64_t na_mask[kMaxNonAscii][W];
    __shared__ int n_na;
    __shared__ double red[kThreads / 32];
    __shared__ int red_cnt[kThreads / 32];

    const int64_t doc = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* dchars = p.doc_chars + p.doc_off[doc];
    const int d_len = (int)(p.doc_off[doc + 1] - p.doc_off[doc]);
    // shorter string is the pattern; on equal lengths the query stays the pattern (no swap, like the oracle)
    const bool q_is_pat = p.q_len <= d_len;
    const uint32_t* pat = q_is_pat ? p.q_chars : dchars;
    const uint32_t* txt = q_is_pat ? dchars : p.q_chars;
    const int m = q_is_pat ? p.q_len : d_len;
    const int n2 = q_is_pat ? d_len : p.q_len;

    double best = 0.0;
    if (m == 0) {
        best = (n2 == 0) ? 100.0 : 0.0;
    } else {
        // ---- match masks of the pattern
        for (int i = tid; i < 128 * W; i += kThreads) (&pm_ascii[0][0])[i] = 0;
        if (tid == 0) n_na = 0;
        __syncthreads();
        for (int i = tid; i < m; i += kThreads) {
            const uint32_t c = pat[i];
            if (c < 128u) atomicOr((unsigned long long*)&pm_ascii[c][i >> 6], 1ull << (i & 63));
        }
        if (tid == 0) {       // other code points are rare: serial insert
            int cnt = 0;
            for (int i = 0; i < m; ++i) {
                const uint32_t c = pat[i];
                if (c < 128u) continue;
                int j = 0;
                while (j < cnt && na_code[j] != c) ++j;
                if (j == cnt) {
                    if (cnt == kMaxNonAscii) {
                        *p.err = 1;
                        continue;
                    }
                    na_code[cnt] = c;
                    for (int w = 0; w < W; ++w) na_mask[cnt][w] = 0;
                    ++cnt;
                }
                na_mask[j][i >> 6] |= 1ull << (i & 63);
            }
            n_na = cnt;
        }
        __syncthreads();
        const int nna = n_na;
        // ---- windows i = -m+1 .. n2-1 of the text, one per thread at a time
        for (int wi = -m + 1 + tid; wi < n2; wi += kThreads) {
            const int lo = wi > 0 ? wi : 0;
            const int hi = (wi + m < n2) ? wi + m : n2;
            uint64_t V[W];
#pragma unroll
            for (int w = 0; w < W; ++w) V[w] = ~0ull;
            for (int t = lo; t < hi; ++t) {
                uint64_t M[W];
                pm_lookup<W>(__ldg(txt + t), pm_ascii, na_code, na_mask, nna, M);
                unsigned carry = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const uint64_t U = V[w] & M[w];
                    const uint64_t s1 = V[w] + U;
                    const uint64_t s2 = s1 + carry;
                    carry = (s1 < U) | (s2 < s1);
                    V[w] = s2 | (V[w] & ~U);
                }
            }
            int zeros = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int bits = m - 64 * w;      // pattern bits living in this word
                if (bits <= 0) continue;
                const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
                zeros += __popcll(~V[w] & mask);
            }
            // rapidfuzz's arithmetic: indel distance = len sum - 2 * LCS, similarity = 1 - dist / len sum, score = sim * 100
            // (reproduces the published partial_ratio_alignment example 83.33333333333334, which 200 * LCS / len sum misses
            // by one ulp)
            const double lensum = (double)(m + (hi - lo));
            const double dist = (double)(m + (hi - lo) - 2 * zeros);
            const double r = __dmul_rn(__dsub_rn(1.0, __ddiv_rn(dist, lensum)), 100.0);
            if (r > best) best = r;
        }
    }
    // ---- token overlap |Q & D|: doc tokens are sorted unique ids, query set is tiny
    int hits = 0;
    const int32_t* dt = p.doc_tok + p.doc_tok_off[doc];
    const int n_dt = (int)(p.doc_tok_off[doc + 1] - p.doc_tok_off[doc]);
    for (int i = tid; i < n_dt; i += kThreads) {
        const int32_t t = dt[i];
        int a = 0, b = p.n_q_tok;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (p.q_tok[mid] < t) a = mid + 1; else b = mid;
        }
        if (a < p.n_q_tok && p.q_tok[a] == t) ++hits;
    }
    // ---- block reduce (max of best, sum of hits)
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        best = fmax(best, hs_shfl_xor_f64(best, s));
        hits += __shfl_xor_sync(0xFFFFFFFFu, hits, s);
    }
    if (lane == 0) {
        red[warp] = best;
        red_cnt[warp] = hits;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            best = fmax(best, red[w]);
            hits += red_cnt[w];
        }
        const double fuzzy = __ddiv_rn(best, 100.0);                    // core.py:185
        double combined = fuzzy;
        if (p.q_set_size > 0 && n_dt > 0) {                            // core.py:189-191
            const double overlap = __ddiv_rn((double)hits, (double)p.q_set_size);
            combined = __dadd_rn(__dmul_rn(fuzzy, 0.7), __dmul_rn(overlap, 0.3));
        }
        p.out[doc] = __double2float_rn(combined);                      // core.py:197 dtype=np.float32
    }
}

}  // namespace

extern "C" {

int hs_lexical_scores(const uint32_t* doc_chars, const int64_t* doc_off, int64_t n_docs, const uint32_t* q_chars,
                      int32_t q_len, const int32_t* doc_tok, const int64_t* doc_tok_off, const int32_t* q_tok,
                      int32_t n_q_tok, int32_t q_set_size, float* out, int32_t* err_flag, void* stream) {
    HS_REQUIRE(n_docs >= 0 && q_len >= 0 && n_q_tok >= 0, "hs_lexical_scores: negative size");
    if (n_docs == 0) return HS_OK;
    HS_REQUIRE(doc_off != nullptr && doc_tok_off != nullptr && out != nullptr && err_flag != nullptr,
               "hs_lexical_scores: null pointer");
    HS_REQUIRE(q_len <= 512, "hs_lexical_scores: queries longer than 512 characters are not supported (%d)", q_len);
    HS_REQUIRE(n_docs <= 0x7FFFFFFF, "hs_lexical_scores: too many docs for one launch");
    LexParams p;
    p.doc_chars = doc_chars;
    p.doc_off = doc_off;
    p.q_chars = q_chars;
    p.q_len = q_len;
    p.doc_tok = doc_tok;
    p.doc_tok_off = doc_tok_off;
    p.q_tok = q_tok;
    p.n_q_tok = n_q_tok;
    p.q_set_size = q_set_size;
    p.out = out;
    p.err = err_flag;
    p.n = n_docs;
    cudaStream_t st = (cudaStream_t)stream;
    // the pattern is never longer than the query, so the query length picks the word count
    if (q_len <= 64)
        lexical_kernel<1><<<(unsigned)n_docs, kThreads, 0, st>>>(p);
    else if (q_len <= 128)
        lexical_kernel<2><<<(unsigned)n_docs, kThreads, 0, st>>>(p);
    else if (q_len <= 256)
        lexical_kernel<4><<<(unsigned)n_docs, kThreads, 0, st>>>(p);
    else
        lexical_kernel<8><<<(unsigned)n_docs, kThreads, 0, st>>>(p);
    HS_LAUNCH_CHECK();
    return HS_OK;
}

}  // extern "C"
