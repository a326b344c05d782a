"""Shared definition of the golden cases: inputs are rebuilt from seeds here, reference outputs
come from the committed ``tests/golden/*.npz`` (written by ``oracle/make_golden.py`` in the dev
container by executing the unmodified reference)."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from hybrid_search_engine_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def t1_corpus(spec, thresholds):
    """Synthetic docs with the reference's edge cases spliced in at fixed positions."""
    docs = synth.doc_texts(spec, 0, spec.n_docs, thresholds)
    emb = synth.embeddings(spec, 0, spec.n_docs)
    docs[3] = ""                                   # empty doc ...
    emb[3] = 0.0                                   # ... shares the text key "" with the empty query
    docs[7] = "the and of   is"                   # stop words only -> dl 0
    docs[11] = "  T5\tt5\n t5  the T9 "            # case / whitespace / duplicates
    docs[20] = docs[19]                            # duplicate doc -> exact BM25 ties
    emb[20] = emb[19]                              # ... and exact cosine ties
    emb[25] = 0.0                                  # zero vector (utils.py:49-50)
    # the reference embeds by text, so equal texts must carry equal vectors (first occurrence wins)
    first = {}
    for i, d in enumerate(docs):
        key = " ".join(d.split())
        if key in first:
            emb[i] = emb[first[key]]
        else:
            first[key] = i
    return docs, emb


def t1_extra_queries(spec, queries: List[str]):
    nq = len(queries)
    queries = queries + ["t5 t5 t9", "the and of", "", "unknownterm t1", "T3,t4;t3"]
    q_emb = synth.query_embeddings(spec, 0, len(queries))
    q_emb[nq + 2] = 0.0                            # zero query vector (utils.py:44-45)
    return queries, q_emb


@dataclass
class GoldenCase:
    name: str
    docs: List[str]
    emb: np.ndarray
    queries: List[str]
    q_emb: np.ndarray
    ref: Dict[str, np.ndarray]
    meta: dict


def load_case(name: str) -> GoldenCase:
    meta = json.load(open(os.path.join(GOLDEN_DIR, f"{name}.json")))
    ref = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    if name.startswith("t0"):
        return GoldenCase(name, meta["docs"], ref["emb"], meta["queries"], ref["q_emb"], ref, meta)
    spec = synth.SynthSpec(**meta["spec"])
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    docs, emb = t1_corpus(spec, th)
    queries = synth.query_texts(spec, 0, meta["n_synth_queries"], th)
    queries, q_emb = t1_extra_queries(spec, queries)
    assert queries == meta["queries"]
    return GoldenCase(name, docs, emb, queries, q_emb, ref, meta)


ALL_CASES = ("t0_sample_docs", "t1_small", "t1_mid")


T2_CASES = ("t2_60k", "t2_240k")


def load_t2(name: str = "t2_60k"):
    """T2 tier: 60 k- / 240 k-doc corpus regenerated from seeds + the reference outputs frozen by
    ``python -m oracle.make_golden t2`` / ``t2b`` (top-k lists, hashes and samples instead of full vectors)."""
    meta = json.load(open(os.path.join(GOLDEN_DIR, f"{name}.json")))
    ref = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    spec = synth.SynthSpec(**meta["spec"])
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    docs = synth.doc_texts(spec, 0, spec.n_docs, th)
    emb = synth.embeddings(spec, 0, spec.n_docs)
    q_emb = synth.query_embeddings(spec, 0, len(meta["queries"]))
    return GoldenCase(name, docs, emb, meta["queries"], q_emb, ref, meta)
