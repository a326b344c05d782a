"""CPU-only checks: the C-ABI library loads and exports what include/hs_b200.h declares, the host-side
logic (tokeniser, lexical statistics, synthetic generators) matches the oracle, and the product path
fails loudly -- never falls back -- without a GPU."""
import os
import re

import numpy as np
import pytest

from oracle import hybrid_oracle as orc
from tests.golden_cases import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_INDEX_FAISS_SHA256 = "e081a6078b5cac6b2488a55c80a8af4029d0adc35c73edfadd94eb3e1edc39ca"


def test_library_exports_every_declared_symbol():
    from hybrid_search_engine_b200 import _lib, build_native
    build_native.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hs_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hs_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hybrid_search_engine_b200 as hs
    from hybrid_search_engine_b200._lib import HsError
    p = hs.create_pipeline("bm25")
    with pytest.raises(HsError):
        p.index(["alpha beta", "gamma"])
    p = hs.create_pipeline("hybrid_bm25")
    with pytest.raises(HsError):
        p.index(["alpha beta", "gamma"], embeddings=np.ones((2, 8), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid_search_engine_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(\s*['\"]oracle", src,
                                 flags=re.M), fn


def test_create_pipeline_names_and_errors():
    import hybrid_search_engine_b200 as hs
    with pytest.raises(ValueError, match=r"Unknown pipeline: nope\. Choose from \['basic', 'chunked', 'reranked', "
                                         r"'bm25', 'hybrid_bm25', 'rag', 'multi_stage', 'diversity'\]"):
        hs.create_pipeline("nope")
    for name, cls in (("basic", "BasicPipeline"), ("bm25", "BM25Pipeline"), ("hybrid_bm25", "HybridBM25Pipeline"),
                      ("multi_stage", "MultiStagePipeline"), ("diversity", "DiversityPipeline")):
        assert type(hs.create_pipeline(name)).__name__ == cls
    assert hs.create_pipeline().semantic_weight == 0.7          # default "basic", sw = 0.7
    p = hs.create_pipeline("multi_stage")
    assert (p.stage1_k, p.stage2_k, p.final_k) == (100, 20, 5)
    assert hs.create_pipeline("diversity").lambda_param == 0.5
    h = hs.create_pipeline("hybrid_bm25", semantic_weight=0.9, bm25_weight=0.9)   # not validated (reference)
    assert (h.semantic_weight, h.bm25_weight) == (0.9, 0.9)
    with pytest.raises(NotImplementedError):
        hs.create_pipeline("rag")


@pytest.mark.parametrize("name", ["t0_sample_docs", "t1_small"])
def test_lexical_stats_match_oracle_fit(name):
    from hybrid_search_engine_b200.index import LexicalStats, idf_from_df
    c = load_case(name)
    st = LexicalStats().fit(c.docs)
    o = orc.bm25_fit(c.docs)
    assert st.vocab == o.vocab
    assert np.array_equal(st.doc_lengths, c.ref["doc_lengths"])
    assert float(st.avg_doc_len) == float(c.ref["avg_doc_len"])
    assert np.array_equal(st.indptr, o.indptr)
    assert np.array_equal(st.postings[:, 0], o.post_doc) and np.array_equal(st.postings[:, 1], o.post_tf)
    idf = idf_from_df(st.doc_count, st.df)
    terms = c.meta["idf_terms"]
    assert np.array_equal(np.array([idf[st.vocab[t]] for t in terms]), c.ref["idf_vals"])   # reference idf
    for q in c.queries:
        assert st.query_term_ids(q) == [o.vocab[t] for t in orc.extract_tokens(q, True) if t in o.vocab]


def test_tokeniser_matches_oracle():
    from hybrid_search_engine_b200 import extractor as ex
    assert ex.STOPWORDS == orc.STOPWORDS
    for s in ["", "The Quick-brown_fox's 2nd naïve café", "  a\tb\n c  ", "ÀB ſ İx", "the and of"]:
        for rs in (False, True):
            assert ex.extract_tokens(s, rs) == orc.extract_tokens(s, rs)
        assert ex.preprocess_text(s) == orc.preprocess_text(s)
    assert ex.extract_tokens("naïve café") == ["na", "ve", "caf"]


def test_synth_is_deterministic_and_well_formed():
    from hybrid_search_engine_b200 import synth
    spec = synth.SynthSpec(n_docs=1000, vocab=500, dim=64, min_len=5, max_len=20)
    a = synth.embeddings(spec, 10, 20)
    assert np.array_equal(a, synth.embeddings(spec, 0, 30)[10:20])          # counter based: slice-invariant
    assert abs(float(synth.embeddings(spec, 0, 1000).std()) - 1.0) < 0.02
    dl, terms = synth.doc_tokens(spec, 0, 1000)
    assert dl.min() >= 5 and dl.max() <= 20 and terms.max() < 500
    dl2, terms2 = synth.doc_tokens(spec, 500, 1000)
    assert np.array_equal(terms2, terms[dl[:500].sum():])
    th = synth.zipf_thresholds(500)
    assert np.all(th[1:] >= th[:-1]) and th[-1] == np.uint64(0xFFFFFFFFFFFFFFFF)
    # streams of different seeds do not alias (the bug class the seed scrambling prevents)
    assert synth.query_texts(spec, 0, 8) != synth.doc_texts(spec, 0, 8)
    assert not np.array_equal(synth.query_embeddings(spec, 0, 8), synth.embeddings(spec, 0, 8))


def test_faiss_flat_io_reproduces_the_reference_file(tmp_path):
    """The writer must emit the reference's index.faiss byte for byte from its 12 rows (sha256 of the file
    at /root/reference/index.faiss, recorded when the golden fixtures were made)."""
    import hashlib
    from hybrid_search_engine_b200 import faiss_io
    c = load_case("t0_sample_docs")
    path = str(tmp_path / "index.faiss")
    faiss_io.write_index_flat(path, c.emb, metric=0)
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == REFERENCE_INDEX_FAISS_SHA256
    vec, metric = faiss_io.read_index_flat(path)
    assert metric == 0 and vec.shape == (12, 384) and np.array_equal(vec, c.emb)
    faiss_io.write_index_flat(path, c.emb * 3.0, metric=0, normalize=True)      # FAISSIndex.add normalises
    vec, _ = faiss_io.read_index_flat(path)
    assert np.allclose(np.linalg.norm(vec, axis=1), 1.0, atol=1e-6)
    with pytest.raises(ValueError):
        open(path, "wb").write(b"nope" + b"\\0" * 60)
        faiss_io.read_index_flat(path)


def test_token_hash_host_twin():
    """index_build.py: the host twin of the device token hash is a 63-bit value, distinct for the stop words and
    for a sample of synthetic terms (term identity of the device index build)."""
    from hybrid_search_engine_b200.extractor import STOPWORDS
    from hybrid_search_engine_b200.index_build import STOP_HASHES, token_hash
    assert len(set(STOP_HASHES.tolist())) == len(STOPWORDS) == 48
    hs_ = {token_hash(f"t{i}") for i in range(200000)} | {token_hash(w) for w in STOPWORDS}
    assert len(hs_) == 200000 + 48 and all(0 <= h < (1 << 63) for h in hs_)
    assert token_hash("t0") == 1444696336046087048 and token_hash("the") == 3841901135645180336   # pinned values
    assert token_hash("ab") != token_hash("ba") and token_hash("a") != token_hash("a_")


def test_doc_table_reranker_and_weight_validation():
    """Host-only pieces of core.py: the two-column doc table (core.py:240-241 of the reference), the injectable
    stage-3 reranker (reranker.py:50-89: stable sort by the external score, optional cut) and the weight check
    of Searcher.search (core.py:232-233), which fires before any device work."""
    from hybrid_search_engine_b200.core import CrossEncoderReranker, DocTable, Searcher
    t = DocTable(["a b", "c", ""])
    assert len(t) == 3 and t["content"].to_list() == ["a b", "c", ""] and t["doc_id"].to_list() == [0, 1, 2]
    assert t["content"][1] == "c" and len(t["doc_id"]) == 3
    with pytest.raises(KeyError):
        t["nope"]
    rr = CrossEncoderReranker(predict=lambda pairs: [len(c) for _, c in pairs])
    res = [(0.9, "x", 0), (0.8, "yyy", 1), (0.7, "zz", 2), (0.6, "ww", 3)]
    assert rr.rerank("q", res) == [(3.0, "yyy", 1), (2.0, "zz", 2), (2.0, "ww", 3), (1.0, "x", 0)]    # ties keep order
    assert rr.rerank("q", res, top_k=2) == [(3.0, "yyy", 1), (2.0, "zz", 2)]
    assert rr.rerank("q", []) == []
    s = Searcher(encoder=object())
    with pytest.raises(ValueError, match="semantic_weight and lexical_weight must sum to 1.0"):
        s.search("q", DocTable(["a"]), np.ones((1, 4), np.float32), semantic_weight=0.5, lexical_weight=0.6)


def test_bm25_work_partition_model():
    """Python mirror of the work partition of bm25_batch_kernel (csrc/bm25.cu: host choice of queries per item,
    contiguous item runs per group, merging of same-tile items): every (doc tile, query) pair must be scored exactly
    once, in passes of 1..8 queries, for awkward batch sizes and tile counts.  Keep in sync with the kernel."""
    k_max_qpc, k_groups = 8, 3

    def run(n_tiles, B, sms=148):
        qpc = min(B, k_max_qpc)
        while qpc > 1 and n_tiles * ((B + qpc - 1) // qpc) < 16 * k_groups * sms:
            qpc //= 2
        nqg = (B + qpc - 1) // qpc
        n_items = n_tiles * nqg
        grid = min((n_items + k_groups - 1) // k_groups, sms)
        seen = set()
        for me in range(grid * k_groups):
            item, hi = n_items * me // (grid * k_groups), n_items * (me + 1) // (grid * k_groups)
            while item < hi:
                tile, qg = divmod(item, nqg)
                qg_end = qg + (hi - item) if hi - item < nqg - qg else nqg
                if (qg_end - qg) * qpc > k_max_qpc:
                    qg_end = qg + k_max_qpc // qpc
                b0 = qg * qpc
                nb = min(B, qg_end * qpc) - b0
                assert qg_end > qg and 1 <= nb <= k_max_qpc
                for b in range(b0, b0 + nb):
                    assert (tile, b) not in seen
                    seen.add((tile, b))
                item += qg_end - qg
        assert len(seen) == n_tiles * B

    for n_tiles in (1, 2, 7, 100, 305, 977, 2442):
        for B in (1, 2, 3, 5, 7, 8, 9, 13, 19, 32, 33, 100):
            run(n_tiles, B)
    run(50, 19, sms=1)
    run(50, 19, sms=132)
