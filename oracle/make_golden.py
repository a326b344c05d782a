"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the UNMODIFIED reference into tests/golden/.

Run in the dev container (the only place ``/root/reference`` exists):

    python -m oracle.make_golden

Executes the reference's own ``bm25.py`` / ``utils.py`` / ``core.py`` / ``pipelines.py`` through
``oracle/refload.py`` (inert stand-ins for polars / duckdb / sentence_transformers; the shared
``partial_ratio`` restatement where rapidfuzz's value matters) on

* T0: the reference's only fixture -- the 12 ``SAMPLE_DOCUMENTS`` of ``main.py:25-38`` with the 12
  normalised MiniLM rows stored in ``index.faiss`` as their embeddings;
* T1: seeded synthetic Zipf corpora from ``hybrid_search_engine_b200.synth`` with edge cases spliced
  in (empty docs, stop-word-only docs, zero vectors, duplicate docs => exact score ties).

and stores inputs that cannot be regenerated from seeds plus every reference output the parity
tests compare against.  The GPU box never runs this script; it reads the committed ``.npz``/``.json``.
"""
from __future__ import annotations

import ast
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import hybrid_oracle as orc  # noqa: E402
from oracle import refload  # noqa: E402
from hybrid_search_engine_b200 import synth  # noqa: E402
from tests.golden_cases import t1_corpus, t1_extra_queries  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_faiss_flat(path):
    """IndexFlatIP layout: 'IxFI', d, ntotal, ..., u64 count, f32 data."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"IxFI"
    d, = struct.unpack_from("<i", raw, 4)
    ntotal, = struct.unpack_from("<q", raw, 8)
    count, = struct.unpack_from("<Q", raw, len(raw) - 8 - 4 * d * ntotal)
    assert count == d * ntotal
    return np.frombuffer(raw, dtype="<f4", count=d * ntotal, offset=len(raw) - 4 * d * ntotal).reshape(ntotal, d).copy()


def sample_documents():
    src = open(os.path.join(refload.REFERENCE_ROOT, "main.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "SAMPLE_DOCUMENTS":
            return ast.literal_eval(node.value)
    raise RuntimeError("SAMPLE_DOCUMENTS not found")


def run_case(ref, name, docs, emb, queries, q_emb, extra_mmr=None):
    """Run every in-scope pipeline of the reference on one corpus; return dict of arrays."""
    n = len(docs)
    out = {}
    refload.EMBED_DIM[0] = emb.shape[1]
    refload.EMBED_TABLE.clear()
    for t, e in list(zip(map(orc.preprocess_text, docs), emb)) + list(zip(queries, q_emb)):
        # the encoder stand-in is keyed by text: equal texts must carry equal vectors
        assert t not in refload.EMBED_TABLE or np.array_equal(refload.EMBED_TABLE[t], e), t
        refload.EMBED_TABLE[t] = e
    inert = lambda a, b: 50.0                       # weight is 0.0 on these paths (pipelines.py:322-323,479-480)
    refload.PARTIAL_RATIO_FN[0] = inert
    refload.CROSS_ENCODER_FN[0] = lambda pairs: np.arange(len(pairs), 0, -1, dtype=np.float32)

    P = ref.pipelines
    hyb = P.create_pipeline("hybrid_bm25"); hyb.index(docs)
    bm = P.create_pipeline("bm25"); bm.index(docs)
    ms = P.create_pipeline("multi_stage", stage1_k=min(100, n), stage2_k=20); ms.index(docs)
    do_lex = n <= 600                               # pure-python LCS is O(N*|q|*|doc|)
    if do_lex:
        ba = P.create_pipeline("basic"); ba.index(docs)
        dv = P.create_pipeline("diversity", lambda_param=0.5); dv.index(docs)

    out["idf_terms"] = np.array(sorted(hyb.bm25.idf), dtype=object)
    out["idf_vals"] = np.array([hyb.bm25.idf[t] for t in sorted(hyb.bm25.idf)], np.float64)
    out["doc_lengths"] = np.array(hyb.bm25.doc_lengths, np.int64)
    out["avg_doc_len"] = np.float64(hyb.bm25.avg_doc_len)

    for qi, (q, qe) in enumerate(zip(queries, q_emb)):
        k = f"q{qi}_"
        # kernels (bm25.py:114-127, utils.py:28-54, utils.py:57-71)
        out[k + "bm25"] = hyb.bm25.score_batch(q)
        assert np.array_equal(hyb.vectors, emb)
        cos = ref.utils.batch_cosine_sim(qe.astype(np.float32), hyb.vectors)
        out[k + "cos"] = cos
        out[k + "sem_norm"] = ref.utils.normalize_scores(cos)
        # hybrid_bm25 full ranking (pipelines.py:315-357)
        r = hyb.search(q, top_k=n)
        out[k + "hyb_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
        out[k + "hyb_scores"] = np.array([x["score"] for x in r.results], np.float32)
        assert all(type(x["score"]) is np.float32 for x in r.results)
        # bm25 pipeline (pipelines.py:270-280)
        r = bm.search(q, top_k=10)
        out[k + "bm25_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
        out[k + "bm25_scores"] = np.array([x["score"] for x in r.results], np.float64)
        # multi_stage: stage 1 via the Searcher, stage 2 via BM25.score, final via the pipeline with
        # a cross-encoder stand-in that preserves stage-2 order (pipelines.py:470-496)
        s1 = ms.searcher.search(query=q, docs_df=ms.docs_df, vectors=ms.vectors, top_k=ms.stage1_k,
                                semantic_weight=1.0, lexical_weight=0.0)
        out[k + "ms_stage1_ids"] = np.array([d for _, _, d in s1], np.int64)
        out[k + "ms_stage1_scores"] = np.array([s for s, _, _ in s1], np.float64)
        out[k + "ms_stage2_bm25"] = np.array([ms.bm25.score(q, d) for _, _, d in s1], np.float64)
        r = ms.search(q, top_k=20)
        out[k + "ms_final_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
        if do_lex:
            refload.PARTIAL_RATIO_FN[0] = orc.partial_ratio   # shared restatement (parity unpinned)
            r = ba.search(q, top_k=10)
            out[k + "basic_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
            out[k + "basic_scores"] = np.array([x["score"] for x in r.results], np.float64)
            out[k + "lex"] = ba.searcher._lexical_scores(q, [orc.preprocess_text(d) for d in docs])
            r = dv.search(q, top_k=5)
            out[k + "div_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
            out[k + "div_scores"] = np.array([x["score"] for x in r.results], np.float64)
            refload.PARTIAL_RATIO_FN[0] = inert
    # MMR kernel on its own (pipelines.py:531-569)
    if extra_mmr is not None:
        dv = P.create_pipeline("diversity", lambda_param=extra_mmr["lam"])
        e, rel = extra_mmr["emb"], extra_mmr["rel"]
        out["mmr_sel"] = np.array(dv._mmr(None, e, rel, extra_mmr["k"]), np.int64)
    return out


def main():
    ref = refload.load()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- T0: SAMPLE_DOCUMENTS + index.faiss rows
    docs0 = sample_documents()
    emb0 = read_faiss_flat(os.path.join(refload.REFERENCE_ROOT, "index.faiss"))
    assert emb0.shape == (12, 384)
    queries0 = ["wise sayings about starting", "machine learning and AI", "programming languages",
                "the quick brown fox", "data data science", "zzz unknown"]
    rng = np.random.default_rng(7)
    # doc-vs-doc queries: blends of stored rows (no MiniLM here)
    q_emb0 = np.stack([emb0[1] * 0.6 + emb0[7] * 0.4, emb0[8] * 0.7 + emb0[10] * 0.3, emb0[9],
                       emb0[0], emb0[9] * 0.5 + emb0[8] * 0.5,
                       rng.standard_normal(384).astype(np.float32)]).astype(np.float32)
    g0 = run_case(ref, "t0", docs0, emb0, queries0, q_emb0)
    np.savez_compressed(os.path.join(GOLDEN, "t0_sample_docs.npz"), emb=emb0, q_emb=q_emb0,
                        **{k: v for k, v in g0.items() if v.dtype != object})
    json.dump({"docs": docs0, "queries": queries0, "idf_terms": g0["idf_terms"].tolist()},
              open(os.path.join(GOLDEN, "t0_sample_docs.json"), "w"), indent=1)

    # ---- T1: synthetic Zipf corpora (inputs regenerated from seeds by tests/golden_cases.py)
    for name, spec, nq in (
        ("t1_small", synth.SynthSpec(n_docs=400, vocab=300, dim=48, min_len=3, max_len=30), 6),
        ("t1_mid", synth.SynthSpec(n_docs=3000, vocab=2000, dim=384, min_len=20, max_len=60), 8),
    ):
        th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
        docs, emb = t1_corpus(spec, th)
        queries = synth.query_texts(spec, 0, nq, th)
        queries, q_emb = t1_extra_queries(spec, queries)
        mmr = None
        if name == "t1_small":
            mmr = {"emb": emb[:60].copy(), "rel": orc.diversity_relevance(
                np.linspace(1.0, 0.2, 60).tolist()), "k": 15, "lam": 0.5}
            mmr["emb"][31] = mmr["emb"][30]        # duplicate candidate -> sim 1.0
        g = run_case(ref, name, docs, emb, queries, q_emb, mmr)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"),
                            **{k: v for k, v in g.items() if v.dtype != object})
        json.dump({"queries": queries, "n_synth_queries": nq, "idf_terms": g["idf_terms"].tolist(),
                   "spec": spec.__dict__}, open(os.path.join(GOLDEN, f"{name}.json"), "w"), indent=1)
        print(name, "done", {k: v.shape for k, v in list(g.items())[:4]})


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


PLUS_SCORE_DOCS = [0, 1, 7, 31, 32, 199, 398, 399]


def main_plus():
    """BM25Plus (bm25.py:150-179) on the t1_small corpus -> tests/golden/t1_small_bm25plus.npz."""
    ref = refload.load()
    spec = synth.SynthSpec(n_docs=400, vocab=300, dim=48, min_len=3, max_len=30)
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    docs, _ = t1_corpus(spec, th)
    queries, _ = t1_extra_queries(spec, synth.query_texts(spec, 0, 6, th))
    out = {}
    for name, kw in (("d1", dict(delta=1.0)), ("d05_k12", dict(k1=1.2, b=0.5, delta=0.5))):
        bm = ref.bm25.BM25Plus(**kw)
        bm.fit(docs)
        for qi, q in enumerate(queries):
            out[f"{name}_q{qi}"] = bm.score_batch(q)
            # BM25Plus.score (bm25.py:161-179): unrounded python floats for a few docs
            out[f"{name}_q{qi}_score64"] = np.array([bm.score(q, d) for d in PLUS_SCORE_DOCS], np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "t1_small_bm25plus.npz"), **out)
    print("bm25plus golden written", len(out))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "plus":
    main_plus()


T2_SPEC = dict(n_docs=60_000, vocab=100_000, dim=128)
T2_EXTRA_QUERIES = ["t5 t5 t9", "unknownterm t1"]


T2B_SPEC = dict(n_docs=240_000, vocab=200_000, dim=64)


def main_t2(name="t2_60k", spec_kw=None):
    """T2 tier: the unmodified reference at a size where BM25 spans many doc tiles
    (60 k docs x 100-300 tokens, 100 k-term Zipf vocabulary, 128-d).  Full score vectors are too large to commit,
    so the fixture keeps the reference's top-100 of hybrid_bm25, the canonical top-100 of its BM25 vector, a
    sha256 of the whole BM25 vector (the oracle reproduces it bit for bit) and the reference cosine at 2048
    sampled docs -> tests/golden/t2_60k.npz (+ .json).   python -m oracle.make_golden t2"""
    import hashlib
    ref = refload.load()
    spec_kw = spec_kw or T2_SPEC
    spec = synth.SynthSpec(**spec_kw)
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    docs = synth.doc_texts(spec, 0, spec.n_docs, th)
    emb = synth.embeddings(spec, 0, spec.n_docs)
    queries = synth.query_texts(spec, 0, 6, th) + T2_EXTRA_QUERIES
    q_emb = synth.query_embeddings(spec, 0, len(queries))
    refload.EMBED_DIM[0] = emb.shape[1]
    refload.EMBED_TABLE.clear()
    for t, e in list(zip(map(orc.preprocess_text, docs), emb)) + list(zip(queries, q_emb)):
        assert t not in refload.EMBED_TABLE or np.array_equal(refload.EMBED_TABLE[t], e), t[:40]
        refload.EMBED_TABLE[t] = e
    refload.PARTIAL_RATIO_FN[0] = lambda a, b: 50.0        # lexical weight is 0.0 on this path
    hyb = ref.pipelines.create_pipeline("hybrid_bm25")
    hyb.index(docs)
    assert np.array_equal(hyb.vectors, emb)
    out = {"doc_lengths_sha256": np.array(hashlib.sha256(np.asarray(hyb.bm25.doc_lengths, np.int64).tobytes()).hexdigest()),
           "avg_doc_len": np.float64(hyb.bm25.avg_doc_len),
           "cos_sample_idx": np.sort(np.random.default_rng(5).choice(spec.n_docs, 2048, replace=False))}
    for qi, (q, qe) in enumerate(zip(queries, q_emb)):
        k = f"q{qi}_"
        bm = hyb.bm25.score_batch(q)
        out[k + "bm25_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(bm).tobytes()).hexdigest())
        top = orc.canonical_topk(bm, 100)
        out[k + "bm25_top_ids"], out[k + "bm25_top_scores"] = top.astype(np.int64), bm[top]
        cos = ref.utils.batch_cosine_sim(qe.astype(np.float32), hyb.vectors)
        out[k + "cos_sample"] = cos[out["cos_sample_idx"]]
        r = hyb.search(q, top_k=100)
        out[k + "hyb_ids"] = np.array([x["doc_id"] for x in r.results], np.int64)
        out[k + "hyb_scores"] = np.array([x["score"] for x in r.results], np.float32)
        out[k + "cos_at_hyb"] = cos[out[k + "hyb_ids"]]
        out[k + "bm25_at_hyb"] = bm[out[k + "hyb_ids"]]
        out[k + "cos_minmax"] = np.array([cos.min(), cos.max()], np.float32)
        out[k + "bm25_max"] = np.float32(bm.max())
        print("t2 query", qi, repr(q), "done", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
    json.dump({"spec": spec_kw, "queries": queries}, open(os.path.join(GOLDEN, f"{name}.json"), "w"), indent=1)
    print(name, "golden written")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "t2":
    main_t2()
# 240 k docs (59 doc tiles of the BM25 kernel, 200 k-term vocabulary): the largest tier the unmodified reference
# finishes in minutes here (fit 214 us/doc, score_batch 6 us/doc and query)   python -m oracle.make_golden t2b
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "t2b":
    main_t2("t2_240k", T2B_SPEC)
