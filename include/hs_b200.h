/*
 * hs_b200.h -- C ABI of the B200-native hybrid scoring path.
 *
 * Drop-in boundary for the hot path behind the reference's
 *   create_pipeline("basic"|"bm25"|"hybrid_bm25"|"multi_stage"|"diversity").index()/search()
 * (reference: search_engine/pipelines.py:617-646).  The reference is pure Python and has no FFI of
 * its own; each entry point below names the reference compute site (file:line, all under
 * /root/reference/search_engine/) it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative hs_status otherwise; hs_last_error() gives the
 *     message for the calling thread
 *   - pointers are DEVICE pointers owned by the caller unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void*; all work is stream ordered, nothing synchronises
 *   - no allocation on the hot path: scratch comes in through `workspace` arguments whose sizes the
 *     *_workspace_bytes() functions report
 *   - doc ids are uint32 positions in the shard plus the shard's doc_base (global id < 2^32)
 *   - ranking order is the total order (score desc, doc_id asc); it is carried in a 64-bit key
 *       key = ordered_u32(score) << 32 | (0xFFFFFFFF - doc_id)          larger key = better rank
 */
#ifndef HS_B200_H
#define HS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_ABI_VERSION 1

typedef enum {
    HS_OK = 0,
    HS_ERR_ARG = -1,       /* bad argument (null pointer, size, alignment, unsupported shape) */
    HS_ERR_CUDA = -2,      /* a CUDA runtime call failed; message holds cudaGetErrorString */
    HS_ERR_STATE = -3      /* index is missing the part this call needs (dense / csr / stats) */
} hs_status;

/* dense arithmetic modes (utils.py:28-54 batch_cosine_sim) */
typedef enum {
    HS_DENSE_EXACT = 0,    /* float64 accumulation in the conformance order: bit-identical to oracle */
    HS_DENSE_FP32 = 1,     /* float32 FMA accumulation, same lane order (as precise as the reference) */
    HS_DENSE_BF16 = 2,     /* bf16 tcgen05 GEMM at large query batch (stage-1 retrieval, 1e-2) */
    HS_DENSE_TF32X3 = 3    /* 3xTF32 tcgen05 GEMM on the float32 matrix: float32-grade accuracy at large batch */
} hs_dense_mode;

/* score fusion formulas */
typedef enum {
    HS_FUSE_RAW = 0,          /* key = a[i]                      bm25.py:141 / utils.py:74-87            */
    HS_FUSE_SEARCHER = 1,     /* f32(norm(a)*f32(w_a)) + f32(norm(b)*f32(w_b))    core.py:264-268        */
    HS_FUSE_HYBRID_BM25 = 2   /* f32(f64(norm(a))*w_a) + f32(b/max_b)*f32(w_b)    pipelines.py:331-340   */
} hs_fuse_mode;

typedef struct hs_index hs_index;

int         hs_abi_version(void);
const char* hs_last_error(void);

/* ---- index handle: one per GPU shard (replaces the pipeline object's numpy/dict state,
 *      pipelines.py:302-313, bm25.py:45-81) ------------------------------------------------------- */
int hs_index_create(int device, int64_t n_docs, int64_t doc_base, hs_index** out);
int hs_index_destroy(hs_index* idx);
/* float32 rows [n_docs, ld] (ld % 4 == 0, base 16-byte aligned, columns >= dim are zero) and their
 * float32 L2 norms from hs_row_norms (indexer.py:285 `vectors`; norms: utils.py:47) */
int hs_index_set_dense(hs_index* idx, const float* vectors, int32_t dim, int64_t ld, const float* vnorm);
/* inverted index over term ids: indptr int64[n_terms+1], postings (doc_id u32, tf u32)[n_postings],
 * doc ids local to the shard and ascending inside each list (bm25.py:62-67 term_freqs, transposed) */
int hs_index_set_csr(hs_index* idx, const int64_t* indptr, const uint32_t* postings, int64_t n_terms,
                     int64_t n_postings);
/* doc lengths u32[n_docs] after stop-word removal (bm25.py:59-60), corpus-global avgdl (bm25.py:71),
 * k1, b (bm25.py:19-33); impact_table double[(max_dl+1) * (tf_cap+1)] from hs_bm25_impact_table, or
 * NULL to compute every posting inline.  With a table the call checks max(dl) <= max_dl on the device
 * (index time: ordered on `stream`, the stream dl and the table were produced on, then synchronised) because the scoring kernels index the table by doc length unchecked; the
 * kernels keep a copy of the table in shared memory when (max_dl+1) * ((tf_cap+1)|1) <= ~10 000 entries */
int hs_index_set_doc_stats(hs_index* idx, const uint32_t* dl, double avgdl, double k1, double b,
                           const double* impact_table, uint32_t max_dl, uint32_t tf_cap, void* stream);

/* ---- index-time kernels ------------------------------------------------------------------------ */
/* vnorm[i] = f32(sqrt(sum64 v[i,:]^2)), conformance order (utils.py:47 np.linalg.norm per row) */
int hs_row_norms(const float* vectors, int64_t n, int32_t dim, int64_t ld, float* vnorm, void* stream);
/* table[dl * (tf_cap+1) + tf] = (tf * (k1+1)) / (tf + k1 * ((1-b) + b * (dl / avgdl))) in float64 with the
 * reference's operation order (bm25.py:107-110); 0 where the reference adds 0 (denominator <= 0) */
int hs_bm25_impact_table(double avgdl, double k1, double b, uint32_t max_dl, uint32_t tf_cap, double* table,
                         void* stream);

/* Hot terms (optional, index time; call after hs_index_set_csr / hs_index_set_doc_stats): for n_hot terms that occur
 * in most docs, hot_c[h, doc] = idf_h * num / den (bm25.py:107-110; 0.0 where the doc lacks the term) is materialised
 * as a dense float64 vector [n_hot, n_docs]; hs_bm25_score then STREAMS that vector for such a term instead of
 * gathering per posting -- same float64 addends in the same order, hence the same bits.  hot_of_term int32 [n_terms]
 * maps a term to its row (-1 = not hot).  n_hot = 0 detaches. */
int hs_bm25_build_hot(hs_index* idx, const int32_t* hot_terms, const double* hot_idf, int32_t n_hot,
                      const int32_t* hot_of_term, double* hot_c, void* stream);

/* ---- hot path ---------------------------------------------------------------------------------- */
/* stats: uint32[B][4] order-preserving encodings {min_a, max_a, max_b, min_b}; reset before a batch */
int hs_stats_reset(uint32_t* stats_enc, int32_t B, void* stream);
int hs_stats_decode(const uint32_t* stats_enc, float* stats, int32_t B, void* stream);
int hs_stats_encode(const float* stats, uint32_t* stats_enc, int32_t B, void* stream);
/* C2 bracket for doc-sharded runs: stats -> float32 [B][4] (-min_a, max_a, max_b, -min_b) so that ONE
 * all-reduce(MAX) yields the corpus-global values (negation is exact; empty shard -> -inf), and back */
int hs_stats_to_maxform(const uint32_t* stats_enc, float* maxform, int32_t B, void* stream);
int hs_stats_from_maxform(const float* maxform, uint32_t* stats_enc, int32_t B, void* stream);
/* fold min/max of x float32 [B, n] into stats slots (utils.py:67-68 for an externally produced vector,
 * e.g. the lexical scores of core.py:261); slot < 0 skips that bound */
int hs_stats_fold_minmax(const float* x, int64_t n, int32_t B, int32_t slot_min, int32_t slot_max,
                         uint32_t* stats_enc, void* stream);

/* K2  batch_cosine_sim (utils.py:28-54) for B queries [B, ld_q] against every row of the shard:
 *     cos[b, i] float32 [B, n_docs]; folds min/max into stats slots 0/1 (utils.py:67-68) */
int hs_dense_scan(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode,
                  float* cos, uint32_t* stats_enc, void* stream);

/* K2b the same scan at large query batch on the tensor cores (tcgen05 GEMM, one corpus pass serves 128-256
 *     queries) over the doc range [doc_lo, doc_hi) of the shard:
 *       HS_DENSE_BF16    needs a bf16 copy of the matrix with UNIT rows, bf16(v_i / |v_i|) (zero rows stay zero),
 *                        [n_docs, ld_bf16] with ld_bf16 a multiple of 64 (zero padded); the queries are normalised the
 *                        same way on the fly, so the accumulator is the cosine itself (no scaling in the epilogue);
 *                        agrees with the float32 path within 2^-8 (stage-1 retrieval)
 *       HS_DENSE_TF32X3  reads the float32 matrix itself, split hi/lo on the fly: float32-grade accuracy
 *     Norms stay float32.  cos[b, i - doc_lo] float32 with row stride cos_ld; min/max folded into stats. */
int hs_index_set_dense_bf16(hs_index* idx, const void* v_bf16, int64_t ld_bf16);
size_t hs_dense_gemm_workspace_bytes(const hs_index* idx, int32_t B, int32_t mode);
int hs_dense_gemm(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                  int64_t doc_hi, void* workspace, size_t workspace_bytes, float* cos, int64_t cos_ld,
                  uint32_t* stats_enc, void* stream);
/* hs_dense_gemm that also lists, per query and (CTA, epilogue group) segment, the docs whose score is within 2 * eps of
 * the segment's running max / min: ext uint64 [B, n_seg, 2, ext_cap] keys (score, shard-local doc; hi side then lo side),
 * ext_cnt uint32 [B, n_seg, 2] (zero first; > ext_cap = overflow).  eps bounds |score - exact cosine| of the mode
 * (bf16: 2^-8 by Cauchy-Schwarz + accumulation slack = 4.2e-3).  Feeds hs_verify_stats. */
int hs_dense_gemm_ext(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                      int64_t doc_hi, void* workspace, size_t workspace_bytes, float* cos, int64_t cos_ld,
                      uint32_t* stats_enc, uint64_t* ext, uint32_t* ext_cnt, int32_t ext_cap, double eps, void* stream);
/* hs_dense_gemm_ext with the SCREEN scores stored as IEEE binary16 [B, cos_ld] (cos_ld a multiple of 8, base 16-byte
 * aligned): half the bytes written here and re-read by hs_fuse_topk_f16.  Statistics and extreme lists come from the
 * float32 accumulators as before; the stored value differs from them by <= 2^-11 (|score| < 2), which the caller adds to
 * the eps it hands hs_verify_topk. */
int hs_dense_gemm_ext_f16(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode, int64_t doc_lo,
                          int64_t doc_hi, void* workspace, size_t workspace_bytes, uint16_t* cos_f16, int64_t cos_ld,
                          uint32_t* stats_enc, uint64_t* ext, uint32_t* ext_cnt, int32_t ext_cap, double eps, void* stream);
/* Exact verification of an approximate scan (screen on the tensor cores, verify in the conformance order; no reference
 * counterpart -- it is what lets the bf16 GEMM return the reference's exact ranking):
 *   hs_verify_stats  replaces stats slots MIN_A / MAX_A by the EXACT min / max cosine of the shard (utils.py:67-68),
 *                    recomputing only the listed candidates; flags[b] |= 1 if a list overflowed
 *   hs_verify_topk   approx_keys [B, k_sel] = the best k_sel docs by the fused score computed from the approximate
 *                    cosine under the exact stats; their cosines are recomputed exactly (utils.py:28-54 in the
 *                    conformance order), the fused score re-evaluated (b = the BM25 vector for HS_FUSE_HYBRID_BM25), the
 *                    list re-sorted: out_keys [B, k_out].  flags[b] |= 2 unless the k_out-th exact score exceeds the
 *                    k_sel-th approximate score by more than |w_a| * eps / (max - min): then no doc outside the list
 *                    can belong to the top k_out and the result is provably the exact mode's; flagged queries must be
 *                    redone in an exact mode by the caller. */
int hs_verify_stats(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, const uint64_t* ext,
                    const uint32_t* ext_cnt, int32_t n_seg, int32_t ext_cap, double eps, uint32_t* stats_enc, int32_t* flags,
                    void* stream);
int hs_verify_topk(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t fuse_mode, const float* b,
                   const uint32_t* stats_enc, double w_a, double w_b, const uint64_t* approx_keys, int32_t k_sel,
                   int32_t k_out, double eps, uint64_t* out_keys, int32_t* flags, void* stream);
/* hs_verify_topk for a select that screened BOTH arrays in binary16 (HS_FUSE_HYBRID_BM25): the b value of candidate i
 * of query q is b_cand[q * k_sel + i], recomputed exactly for the listed docs (hs_keys_local_docs + hs_bm25_score_docs:
 * the float64 BM25 sum, rounded here to the float32 the reference holds, bm25.py:124-126); eps_b bounds the screen's
 * error on b / max_b (2^-11 for binary16) and widens the soundness margin by |w_b| eps_b. */
int hs_verify_topk_cand(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t fuse_mode,
                        const double* b_cand, double eps_b, const uint32_t* stats_enc, double w_a, double w_b,
                        const uint64_t* approx_keys, int32_t k_sel, int32_t k_out, double eps, uint64_t* out_keys,
                        int32_t* flags, void* stream);
/* the same GEMM with the select's pre-filter fused into its epilogue (pure-semantic retrieval: Searcher.search with
 * lexical weight 0, multi_stage stage 1, pipelines.py:474-481): nothing is stored; every (query, doc) cosine >=
 * thr[b] (NULL: all) is appended as a ranking key to the query's candidate lists.  No atomics: each (CTA, epilogue
 * group) of the persistent kernel owns one SEGMENT per query -- cand uint64 [B, n_seg, seg_cap], cand_cnt uint32
 * [B, n_seg] = keys the segment's owner wanted to append (zero it first; > seg_cap afterwards = overflow, the surplus
 * was dropped), n_seg = hs_dense_gemm_filter_segments().  min/max still go to stats. */
int32_t hs_dense_gemm_filter_segments(const hs_index* idx, int32_t mode);
int hs_dense_gemm_filter(const hs_index* idx, const float* queries, int32_t B, int64_t ld_q, int32_t mode,
                         int64_t doc_lo, int64_t doc_hi, void* workspace, size_t workspace_bytes, const float* thr,
                         uint64_t* cand, int32_t seg_cap, uint32_t* cand_cnt, uint32_t* stats_enc, void* stream);

/* K1  BM25.score_batch (bm25.py:83-127) for B queries over the CSR: query b owns tokens
 *     q_off[b]..q_off[b+1]-1 (known terms only, query order, duplicates kept) with their float64 idf
 *     (bm25.py:81).  scores float32 [B, n_docs] (single rounding of the float64 sum, bm25.py:124-126);
 *     folds max into stats slot 2 (pipelines.py:332).  n_tokens = q_off[B] (host copy); workspace
 *     holds the posting offsets of every token at every doc-tile boundary. */
size_t hs_bm25_workspace_bytes(int64_t n_docs, int32_t n_tokens);
int hs_bm25_score(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                  int32_t B, int32_t n_tokens, void* workspace, size_t workspace_bytes, float* scores,
                  uint32_t* stats_enc, void* stream);
/* the same scores stored as IEEE binary16 [B, ld] (ld >= n_docs, a multiple of 8; 16-byte aligned): the BM25 SCREEN of the
 * verified mode -- half the bytes written here and re-read by hs_fuse_topk_f16.  The max folded into stats is that of
 * the float32 scores (exact, pipelines.py:332); exact values of the candidates come from hs_bm25_score_docs. */
int hs_bm25_score_f16(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                      int32_t B, int32_t n_tokens, void* workspace, size_t workspace_bytes, uint16_t* scores_f16,
                      int64_t ld, uint32_t* stats_enc, void* stream);

/* BM25Plus.score over all docs (bm25.py:150-179): idf * (num / den + delta) for EVERY doc and known query
 * token (dense variant, used by no pipeline); same arguments as hs_bm25_score plus delta */
int hs_bm25plus_score(const hs_index* idx, const int32_t* q_terms, const double* q_idf, const int32_t* q_off,
                      int32_t B, int32_t n_tokens, double delta, void* workspace, size_t workspace_bytes,
                      float* scores, uint32_t* stats_enc, void* stream);

/* BM25.score (bm25.py:83-112) for selected docs only: multi_stage stage 2 (pipelines.py:485).
 * doc_ids int64 [B, C] (shard-local, < 0 = padding), out float64 [B, C], unrounded */
int hs_bm25_score_docs(const hs_index* idx, const int32_t* q_terms, const double* q_idf,
                       const int32_t* q_off, int32_t B, const int64_t* doc_ids, int32_t C, double* out,
                       void* stream);

/* BM25Plus.score (bm25.py:161-179) for selected docs: idf * (num / den + delta) per known query token,
 * tf = 0 included; same arguments as hs_bm25_score_docs plus delta */
int hs_bm25plus_score_docs(const hs_index* idx, const int32_t* q_terms, const double* q_idf,
                           const int32_t* q_off, int32_t B, const int64_t* doc_ids, int32_t C, double delta,
                           double* out, void* stream);

/* K3+K4  normalize_scores + weighted fusion (utils.py:57-71, core.py:264-268, pipelines.py:331-340)
 *     fused into the top-k select (core.py:271, bm25.py:141, pipelines.py:342-343): per-CTA partial
 *     top-k lists into `workspace`, then one merge.  a, b: float32 [B, n_docs] (b may be NULL when
 *     w_b == 0).  Only keys < below_key[b] take part (pass NULL for no bound): lets a caller page
 *     past k = HS_TOPK_MAX.  out_keys uint64 [B, k], descending, 0 = no entry. */
#define HS_TOPK_MAX 2048
size_t hs_fuse_topk_workspace_bytes(int64_t n_docs, int32_t B, int32_t k);
int hs_fuse_topk(const hs_index* idx, int32_t fuse_mode, const float* a, const float* b,
                 const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k,
                 const uint64_t* below_key, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                 void* stream);
/* hs_fuse_topk whose a array is the binary16 screen of hs_dense_gemm_ext_f16 (row stride ld elements): the approximate
 * select of the verified mode (fuse_mode SEARCHER or HYBRID_BM25; results go to hs_verify_topk / hs_verify_topk_cand).
 * b: float32 [B, n_docs] (b_is_f16 = 0), or the binary16 BM25 screen of hs_bm25_score_f16, [B, ld] (b_is_f16 = 1,
 * HS_FUSE_HYBRID_BM25 only). */
int hs_fuse_topk_f16(const hs_index* idx, int32_t fuse_mode, const uint16_t* a_f16, const void* b, int32_t b_is_f16,
                     int64_t ld, const uint32_t* stats_enc, double w_a, double w_b, int32_t B, int32_t k, void* workspace,
                     size_t workspace_bytes, uint64_t* out_keys, void* stream);
/* local_ids[i] = shard-local doc id of keys[i], -1 for an empty slot (key 0) or a doc outside [doc_base, doc_base + n_docs):
 * turns a key list into the candidate list of hs_bm25_score_docs */
int hs_keys_local_docs(const uint64_t* keys, int64_t n, int64_t doc_base, int64_t n_docs, int64_t* local_ids, void* stream);
/* top_k_indices (utils.py:74-87) of a float32 [B, n] array with row stride ld that is not a whole shard (e.g. the
 * sample block of the filtered tensor-core scan): keys carry doc_base + position */
int hs_topk_select(const float* x, int64_t n, int64_t ld, int64_t doc_base, int32_t B, int32_t k, void* workspace,
                   size_t workspace_bytes, uint64_t* out_keys, void* stream);
/* thr[b] = score of the kth best key of keys [B, k] (-inf when the list holds fewer than kth keys) */
int hs_keys_kth_score(const uint64_t* keys, int32_t B, int32_t k, int32_t kth, float* thr, void* stream);
/* candidate segments of hs_dense_gemm_filter (+ n_extra keys per query from elsewhere) -> the best k_sel by cosine,
 * re-keyed with the fused score under the FINAL stats (HS_FUSE_SEARCHER, lexical weight 0; HS_FUSE_RAW keeps the
 * cosine) and sorted: out_keys [B, k_out].  *overflow is OR-ed with 1 if any segment overflowed. */
int hs_cand_select(const uint64_t* cand, const uint32_t* cand_cnt, int32_t n_seg, int32_t seg_cap,
                   const uint64_t* extra_keys, int32_t n_extra, int32_t fuse_mode, const uint32_t* stats_enc, double w_a,
                   int32_t B, int32_t k_sel, int32_t k_out, uint64_t* out_keys, int32_t* overflow, void* stream);
/* C1 merge: keys uint64 [n_lists, B, k] (e.g. the all-gathered per-shard lists) -> out_keys [B, k] */
int hs_topk_merge(const uint64_t* keys, int32_t n_lists, int32_t B, int32_t k, uint64_t* out_keys,
                  void* stream);
/* Searcher._semantic_search_faiss bookkeeping (core.py:244-250): scores of the docs named by keys [B, k]
 * scattered into a zeroed float32 [B, n_docs] vector (every other doc keeps 0.0) */
int hs_scatter_keys(const uint64_t* keys, int32_t B, int32_t k, int64_t n_docs, int64_t doc_base, float* out,
                    void* stream);
/* keys -> (float32 score, int64 global doc id; -1 where key == 0) */
int hs_keys_unpack(const uint64_t* keys, int64_t count, float* scores, int64_t* doc_ids, void* stream);

/* K5  DiversityPipeline._mmr (pipelines.py:531-569): greedy MMR over C candidates per query.
 *     cand_ids int64 [B, C] shard-local rows of the dense matrix (< 0 = padding), rel float64 [B, C]
 *     (pipelines.py:589), out_sel int32 [B, k] positions into the candidate list (-1 = none). */
size_t hs_mmr_workspace_bytes(int32_t B, int32_t C);
int hs_mmr(const hs_index* idx, const int64_t* cand_ids, const double* rel, double lambda, int32_t B,
           int32_t C, int32_t k, void* workspace, size_t workspace_bytes, int32_t* out_sel, void* stream);

/* Searcher._lexical_scores (core.py:178-197): 0.7 * partial_ratio(query, doc) / 100 + 0.3 * |Q & D| / |Q|
 * per document (fuzzy part alone when either token set is empty), float64 arithmetic, float32 result.
 * doc_chars / q_chars: code points of the LOWER-CASED strings; doc_tok: sorted unique token ids per doc
 * (no stop-word removal); q_tok: sorted unique query token ids known to the vocabulary; q_set_size = |Q|.
 * partial_ratio follows the published rapidfuzz definition (PARITY UNPINNED). */
int hs_lexical_scores(const uint32_t* doc_chars, const int64_t* doc_off, int64_t n_docs, const uint32_t* q_chars,
                      int32_t q_len, const int32_t* doc_tok, const int64_t* doc_tok_off, const int32_t* q_tok,
                      int32_t n_q_tok, int32_t q_set_size, float* out, int32_t* err_flag, void* stream);

/* ---- index-time tokeniser (BM25.fit, bm25.py:58-67 / extractor.py:15-31): `text` is the lower-cased UTF-8
 *      blob of the documents; flags[i] = 1 where a [a-z0-9_]+ token starts; hashes[t] = 63-bit hash of the
 *      token starting at starts[t] (FNV-1a + splitmix64, host twin in index_build.py) */
int hs_token_flags(const uint8_t* text, int64_t n_bytes, uint8_t* flags, void* stream);
int hs_token_hashes(const uint8_t* text, int64_t n_bytes, const int64_t* starts, int64_t n_tokens, int64_t* hashes,
                    void* stream);

/* ---- index-build primitives (BM25.fit's Counter / dict bookkeeping, bm25.py:56-71, as device passes over 64-bit
 *      (term << 32 | doc) keys): stable LSD radix sort over the key bytes named in byte_mask (bit p = byte p; bytes
 *      that are equal in every key may be left out; tmp: n keys of scratch), binary search in a sorted array, exclusive
 *      prefix sum, run-length encoding of SORTED keys (uniq / start / counts sized n, start n + 1; *total = number
 *      of runs, device scalar), and per-term posting counts df[t] from sorted unique keys (scratch: 2 * n_terms) */
size_t hs_radix_sort_workspace_bytes(int64_t n);
int hs_radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, uint32_t byte_mask, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t hs_scan_workspace_bytes(int64_t n);
int hs_exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* workspace, size_t workspace_bytes,
                          void* stream);
size_t hs_rle_workspace_bytes(int64_t n);
int hs_run_length_encode_u64(const uint64_t* sorted_keys, int64_t n, uint64_t* uniq, int64_t* start, int32_t* counts,
                             int64_t* total, void* workspace, size_t workspace_bytes, void* stream);
int hs_lower_bound_i64(const int64_t* sorted, int64_t n, const int64_t* queries, int64_t m, int64_t* out, void* stream);
int hs_term_doc_freqs(const uint64_t* uniq_keys, int64_t m, int64_t n_terms, int64_t* scratch, int64_t* df, void* stream);

/* ---- synthetic corpus generators (counter-based; hybrid_search_engine_b200/synth.py is the spec) */
int hs_synth_embeddings(float* out, int64_t row0, int64_t n, int32_t dim, int64_t ld, uint64_t seed_key,
                        void* stream);
int hs_synth_doc_lengths(uint32_t* dl, int64_t doc0, int64_t n, uint64_t seed_key, uint32_t min_len,
                         uint32_t span, void* stream);
/* keys[tok_off[i - doc0] + j] = term(i, j) << 32 | (i - doc0)   for doc i in [doc0, doc0 + n) */
int hs_synth_token_keys(uint64_t* keys, const int64_t* tok_off, const uint32_t* dl, int64_t doc0, int64_t n,
                        uint64_t seed_key, const uint64_t* thresholds, int32_t vocab, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HS_B200_H */
