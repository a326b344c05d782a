"""Property-based differential tests: random small corpora through the CUDA pipelines vs the
oracle, bit for bit.  The generators aim at the edge cases the reference's behaviour table lists (section 8b): empty
and stop-word-only docs, duplicate docs (exact score ties), duplicate and unknown query terms, non-ASCII text, zero
vectors / zero queries, all-equal score vectors (min-max -> ones), top_k >= N, one-doc corpora.
"""
import numpy as np
import pytest

from oracle import hybrid_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
hyp = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

WORDS = ["alpha", "beta", "gamma", "delta", "x1", "y_2", "zz", "the", "and", "is", "café", "naïve", "q", "7", "a_b"]
SETTINGS = dict(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@pytest.fixture(scope="module")
def hs():
    import hybrid_search_engine_b200 as hs
    from hybrid_search_engine_b200 import _lib
    _lib.load()
    return hs


doc_st = st.lists(st.sampled_from(WORDS), min_size=0, max_size=12).map(" ".join)
query_st = st.lists(st.sampled_from(WORDS + ["unknownterm", "ALPHA", "Beta!"]), min_size=0, max_size=6).map(" ".join)


@st.composite
def corpus(draw):
    docs = draw(st.lists(doc_st, min_size=1, max_size=40))
    if draw(st.booleans()) and len(docs) > 2:
        docs[-1] = docs[0]                                   # duplicate doc -> exact BM25 ties
    n = len(docs)
    dim = draw(st.sampled_from([4, 7, 48, 130]))
    seed = draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    kind = draw(st.sampled_from(["random", "zero_row", "dup_rows", "all_equal"]))
    if kind == "zero_row":
        emb[rng.integers(n)] = 0.0
    elif kind == "dup_rows" and n > 1:
        emb[-1] = emb[0]
    elif kind == "all_equal":
        emb[:] = emb[0]                                      # every cosine equal -> normalize_scores gives ones
    queries = draw(st.lists(query_st, min_size=1, max_size=4))
    q_emb = rng.standard_normal((len(queries), dim)).astype(np.float32)
    if draw(st.booleans()):
        q_emb[0] = 0.0                                       # zero query -> all-zero cosine
    top_k = draw(st.sampled_from([1, 3, n, n + 3]))
    return docs, emb, queries, q_emb, top_k


@settings(**SETTINGS)
@given(corpus())
def test_bm25_and_hybrid_pipelines_match_the_oracle(hs, c):
    docs, emb, queries, q_emb, top_k = c
    ix = orc.build_index(docs, emb)
    p = hs.create_pipeline("hybrid_bm25")
    p.index(docs, embeddings=emb)
    res = p.search_many(queries, top_k=top_k, query_vectors=q_emb)
    b = hs.create_pipeline("bm25")
    b.index(docs)
    bres = b.search_many(queries, top_k=top_k)
    for qi, q in enumerate(queries):
        ids, sc, _ = orc.search_hybrid_bm25(ix, q, q_emb[qi], top_k)
        assert [r["doc_id"] for r in res[qi].results] == ids.tolist()
        assert np.array_equal(np.array([r["score"] for r in res[qi].results], np.float32), sc)
        assert [r["content"] for r in res[qi].results] == [docs[i] for i in ids]
        ids, sc = orc.search_bm25(ix, q, top_k)
        assert [r["doc_id"] for r in bres[qi].results] == ids.tolist()
        assert [r["score"] for r in bres[qi].results] == [float(s) for s in sc]
        # BM25.score (float64, unrounded) on a few docs
        tids = orc.query_term_ids(ix.bm25, q)
        some = sorted({0, len(docs) - 1, len(docs) // 2})
        want = orc.bm25_score_docs(ix.bm25, tids, some)
        assert [b.bm25.score(q, d) for d in some] == want.tolist()


@settings(**SETTINGS)
@given(corpus())
def test_multi_stage_and_semantic_match_the_oracle(hs, c):
    docs, emb, queries, q_emb, top_k = c
    ix = orc.build_index(docs, emb)
    k1 = min(10, len(docs))
    keep = type("R", (), {"rerank": staticmethod(lambda q, cand, top_k=None: cand[:top_k] if top_k else cand)})()
    p = hs.create_pipeline("multi_stage", stage1_k=k1, stage2_k=5, reranker=keep)
    p.index(docs, embeddings=emb)
    got = p.stages_1_2(queries, query_vectors=q_emb)
    for qi, q in enumerate(queries):
        s1, s2, bm = orc.search_multi_stage(ix, q, q_emb[qi], k1, 5)
        assert [d for _, _, d in got[qi]] == s2.tolist()
        assert [s for s, _, _ in got[qi]] == bm.tolist()


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.integers(1, 3000), st.sampled_from([1, 2, 100, 128, 129, 700]), st.integers(0, 2 ** 16),
       st.sampled_from(["normal", "ties", "constant", "sorted", "signed_zero"]))
def test_topk_select_property(hs, n, k, seed, kind):
    """top_k_indices (utils.py:74-87) under the canonical order for adversarial score vectors, k <= N and k > N."""
    from hybrid_search_engine_b200._lib import HS_FUSE_RAW
    from hybrid_search_engine_b200.engine import SearchEngine
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype(np.float32)
    if kind == "ties":
        x = np.round(x, 1)
    elif kind == "constant":
        x[:] = 0.25
    elif kind == "sorted":
        x = np.sort(x)
    elif kind == "signed_zero":
        x[::2] = 0.0
        x[1::2] = -0.0
    eng = SearchEngine(hs.DeviceIndex("cuda:0", n))
    kk = min(k, n)
    keys = eng._select(HS_FUSE_RAW, torch.from_numpy(x[None, :].copy()).cuda(), None, None, 1.0, 0.0, kk)
    sc, ids = eng.unpack(keys)
    want = orc.canonical_topk(x, kk)
    assert np.array_equal(ids.cpu().numpy()[0], want)
    assert np.array_equal(sc.cpu().numpy()[0], np.where(x[want] == 0, np.float32(0.0), x[want]))
