"""Pin the CPU oracle (oracle/hybrid_oracle.py) to outputs of the UNMODIFIED reference.

The golden files were written by ``oracle/make_golden.py`` executing the reference's own
bm25.py / utils.py / core.py / pipelines.py in the dev container.  Integer / float64-ordered work
(BM25, idf, fusion arithmetic given the same inputs, stage-2 scores, MMR picks) must be
bit-exact; the dense cosine is a float32 BLAS reduction of undefined order in the reference, so
it is pinned to 4 ulp(fp32) at |cos|<=1 and ids may differ only inside near-tie groups.
"""
import numpy as np
import pytest

from oracle import hybrid_oracle as orc
from tests.golden_cases import ALL_CASES, load_case

COS_ATOL = 2.4e-7        # 4 ulp(fp32) at 0.5 <= |x| < 1


@pytest.fixture(scope="module", params=ALL_CASES)
def case(request):
    c = load_case(request.param)
    c.ix = orc.build_index(c.docs, c.emb)
    return c


def _nq(c):
    return len(c.queries)


def test_fit_statistics_bit_exact(case):
    st = case.ix.bm25
    assert np.array_equal(st.doc_lengths, case.ref["doc_lengths"])
    assert float(st.avg_doc_len) == float(case.ref["avg_doc_len"])
    terms = case.meta["idf_terms"]
    assert sorted(t for t, i in st.vocab.items() if st.indptr[i + 1] > st.indptr[i]) == terms
    got = np.array([st.idf[st.vocab[t]] for t in terms])
    assert np.array_equal(got, case.ref["idf_vals"])


def test_bm25_score_batch_bit_exact(case):
    for qi, q in enumerate(case.queries):
        got = orc.bm25_score_batch(case.ix.bm25, q)
        assert got.dtype == np.float32
        assert np.array_equal(got, case.ref[f"q{qi}_bm25"]), f"query {qi} {q!r}"


def test_cosine_within_4ulp(case):
    for qi in range(_nq(case)):
        got = orc.cosine_exact(case.q_emb[qi], case.emb, case.ix.vnorm)
        ref = case.ref[f"q{qi}_cos"]
        assert np.max(np.abs(got.astype(np.float64) - ref)) <= COS_ATOL
        assert np.array_equal(got == 0.0, ref == 0.0) or np.allclose(got, ref, atol=COS_ATOL)


def test_normalize_and_fusion_bit_exact_given_reference_cosine(case):
    """Feed the reference's own cosine vector through the oracle's normalise + fuse: every float32
    step (utils.py:57-71, pipelines.py:331-340) must reproduce the reference bit for bit."""
    n = len(case.docs)
    for qi, q in enumerate(case.queries):
        cos = case.ref[f"q{qi}_cos"]
        sem = orc.normalize_scores(cos)
        assert np.array_equal(sem, case.ref[f"q{qi}_sem_norm"])
        fused = orc.hybrid_bm25_fused(sem, case.ref[f"q{qi}_bm25"], 0.6, 0.4)
        ids = case.ref[f"q{qi}_hyb_ids"]
        assert len(ids) == n
        assert np.array_equal(fused[ids], case.ref[f"q{qi}_hyb_scores"])
        # the reference's stable sort == canonical (score desc, id asc)
        assert np.array_equal(orc.canonical_topk(fused, n), ids)


def _near_tie_equal(ids_a, ids_b, score_of, tol):
    """ids equal except inside groups whose reference scores differ by <= tol."""
    if np.array_equal(ids_a, ids_b):
        return True
    for a, b in zip(ids_a, ids_b):
        if a != b and abs(float(score_of[a]) - float(score_of[b])) > tol:
            return False
    return True


def test_hybrid_bm25_end_to_end(case):
    n = len(case.docs)
    for qi, q in enumerate(case.queries):
        ids, sc, fused = orc.search_hybrid_bm25(case.ix, q, case.q_emb[qi], n)
        ref_ids, ref_sc = case.ref[f"q{qi}_hyb_ids"], case.ref[f"q{qi}_hyb_scores"]
        ref_full = np.empty(n, np.float32)
        ref_full[ref_ids] = ref_sc
        np.testing.assert_allclose(fused, ref_full, rtol=1e-5, atol=1e-6)
        assert _near_tie_equal(ids[:100], ref_ids[:100], ref_full, 1e-6)


def test_bm25_pipeline_topk(case):
    for qi, q in enumerate(case.queries):
        ids, sc = orc.search_bm25(case.ix, q, 10)
        ref_ids, ref_sc = case.ref[f"q{qi}_bm25_ids"], case.ref[f"q{qi}_bm25_scores"]
        # scores bit-exact; ids only up to the reference's unstable argsort among exact ties
        assert np.array_equal(sc.astype(np.float64), ref_sc)
        full = case.ref[f"q{qi}_bm25"]
        assert _near_tie_equal(ids, ref_ids, full, 0.0)


def test_multi_stage_stage2(case):
    for qi, q in enumerate(case.queries):
        s1_ref = case.ref[f"q{qi}_ms_stage1_ids"]
        tids = orc.query_term_ids(case.ix.bm25, q)
        bm = orc.bm25_score_docs(case.ix.bm25, tids, s1_ref)
        assert np.array_equal(bm, case.ref[f"q{qi}_ms_stage2_bm25"])      # float64, unrounded
        order = np.lexsort((np.arange(len(s1_ref)), -bm))[:20]
        assert np.array_equal(s1_ref[order], case.ref[f"q{qi}_ms_final_ids"])
        s1, s2, s2sc = orc.search_multi_stage(case.ix, q, case.q_emb[qi], len(s1_ref), 20)
        sem_ref = case.ref[f"q{qi}_sem_norm"]
        assert np.array_equal(sem_ref[s1_ref].astype(np.float64), case.ref[f"q{qi}_ms_stage1_scores"])
        # reference argsort is unstable among exact ties: compare up to (near-)tie groups
        assert _near_tie_equal(s1, s1_ref, sem_ref, 1e-6)


def test_mmr_selection():
    c = load_case("t1_small")
    emb = c.emb[:60].copy()
    emb[31] = emb[30]
    rel = orc.diversity_relevance(np.linspace(1.0, 0.2, 60).tolist())
    assert orc.mmr_select(emb, rel, 0.5, 15) == c.ref["mmr_sel"].tolist()


@pytest.mark.parametrize("name", ["t0_sample_docs", "t1_small"])
def test_basic_and_diversity_with_shared_partial_ratio(name):
    """PARITY UNPINNED for rapidfuzz itself: both sides use the same partial_ratio restatement, so
    this pins everything around it (core.py:178-197,264-271; pipelines.py:571-613)."""
    c = load_case(name)
    ix = orc.build_index(c.docs, c.emb)
    for qi, q in enumerate(c.queries):
        lex = orc.lexical_scores(q, ix.contents)
        assert np.array_equal(lex, c.ref[f"q{qi}_lex"])
        hyb = orc.searcher_hybrid(c.ref[f"q{qi}_cos"], lex, 0.7, 1.0 - 0.7)
        ref_ids = c.ref[f"q{qi}_basic_ids"]
        assert np.array_equal(hyb[ref_ids].astype(np.float64), c.ref[f"q{qi}_basic_scores"])
        ids, sc, full = orc.search_basic(ix, q, c.q_emb[qi], 10)
        np.testing.assert_allclose(sc, c.ref[f"q{qi}_basic_scores"], rtol=1e-5, atol=1e-6)
        assert _near_tie_equal(ids, ref_ids, hyb, 1e-6)
        dids, dsc = orc.search_diversity(ix, q, c.q_emb[qi], 5)
        assert _near_tie_equal(dids, c.ref[f"q{qi}_div_ids"], hyb, 1e-6)


def test_edge_cases():
    # empty corpus: bm25 returns [], Searcher-backed paths raise
    ix = orc.build_index([], np.zeros((0, 8), np.float32))
    ids, sc = orc.search_bm25(ix, "anything", 5)
    assert len(ids) == 0
    with pytest.raises(ValueError):
        orc.search_hybrid_bm25(ix, "anything", np.ones(8, np.float32), 5)
    # all-empty docs: all scores 0.0, no error (avgdl == 0 never divides: idf is empty)
    ix = orc.build_index(["", "the"], np.eye(2, 8, dtype=np.float32))
    assert np.array_equal(orc.bm25_score_batch(ix.bm25, "the fox"), np.zeros(2, np.float32))
    # constant vector normalises to ones (utils.py:69-70)
    assert np.array_equal(orc.normalize_scores(np.full(4, 0.3, np.float32)), np.ones(4, np.float32))
    with pytest.raises(ValueError):
        orc.searcher_hybrid(np.ones(3, np.float32), np.ones(3, np.float32), 0.7, 0.2)
    # top_k > N returns N
    assert len(orc.canonical_topk(np.array([0.1, 0.5], np.float32), 10)) == 2


@pytest.mark.parametrize("t2_name", ["t2_60k", "t2_240k"])
def test_t2_reference_at_60k_docs(t2_name):
    """T2 tier: the oracle against the unmodified reference on a 60 k-doc and a 240 k-doc corpus (BM25 posting lists
    spanning 15 / 59 doc tiles).  The BM25 vector is compared through its sha256 (bit-exact), the cosine at 2048 sampled docs
    (4 ulp), the hybrid top-100 by ids (up to near ties) and scores."""
    import hashlib
    from tests.golden_cases import load_t2
    c = load_t2(t2_name)
    ix = orc.build_index(c.docs, c.emb)
    assert hashlib.sha256(np.asarray(ix.bm25.doc_lengths, np.int64).tobytes()).hexdigest() == str(c.ref["doc_lengths_sha256"])
    assert float(ix.bm25.avg_doc_len) == float(c.ref["avg_doc_len"])
    idx = c.ref["cos_sample_idx"]
    for qi, q in enumerate(c.queries):
        bm = orc.bm25_score_batch(ix.bm25, q)
        assert hashlib.sha256(np.ascontiguousarray(bm).tobytes()).hexdigest() == str(c.ref[f"q{qi}_bm25_sha256"]), q
        top = orc.canonical_topk(bm, 100)
        assert np.array_equal(top, c.ref[f"q{qi}_bm25_top_ids"]) and np.array_equal(bm[top], c.ref[f"q{qi}_bm25_top_scores"])
        cos = orc.cosine_exact(c.q_emb[qi], c.emb, ix.vnorm)
        assert np.max(np.abs(cos[idx].astype(np.float64) - c.ref[f"q{qi}_cos_sample"])) <= COS_ATOL
        ids, sc, fused = orc.search_hybrid_bm25(ix, q, c.q_emb[qi], 100)
        ref_ids, ref_sc = c.ref[f"q{qi}_hyb_ids"], c.ref[f"q{qi}_hyb_scores"]
        np.testing.assert_allclose(fused[ref_ids], ref_sc, rtol=1e-5, atol=1e-6)
        ref_full = fused.copy()
        ref_full[ref_ids] = ref_sc
        assert _near_tie_equal(ids, ref_ids, ref_full, 2e-6), q


def test_partial_ratio_restatement_reproduces_published_rapidfuzz_values():
    """rapidfuzz is not installable here (parity unpinned); the few known-answer values its public
    documentation carries pin the score arithmetic of the shared restatement (distance -> normalised similarity -> x 100;
    the docs' 83.33333333333334 differs from 200 * LCS / len sum by one ulp)."""
    import json
    import os
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rapidfuzz_published.json")))["vectors"]
    assert len(vec) >= 5
    for v in vec:
        assert orc.partial_ratio(v["s1"], v["s2"]) == v["score"], v
        assert orc.partial_ratio(v["s2"], v["s1"]) == v["score"], v          # symmetric in its arguments
