"""Drop-in pipelines: ``create_pipeline(name).index(documents)`` / ``.search(query, top_k)``.

Mirror of the reference's plugin API for the hot path (pipelines.py:617-646 and the five in-scope
classes): same names, constructor arguments, result dictionaries, metadata and error behaviour.
All scoring runs in the hs_b200 CUDA kernels through the C ABI; this file only tokenises, moves
query bytes and formats results.

Extensions (keyword-only, all optional; the reference has no equivalent because it hard-wires
MiniLM inside Indexer/Searcher, indexer.py:91, core.py:134):

* ``encoder=``      object with ``encode(list[str]) -> float32 [n, d]`` (defaults to
                    sentence-transformers ``all-MiniLM-L6-v2`` if installed, else a clear error)
* ``index_build=``  "host" | "device": where ``BM25.fit`` tokenises and builds the CSR (index_build.py)
* ``device=``       CUDA device of the shard, ``dense_mode=`` "exact" | "fp32" | "tf32x3" | "bf16" | "bf16_exact"
                    (the last three run on the tensor cores; "bf16_exact" = bf16 screen + exact verification,
                    results of "exact" at large query batches)
* ``index(documents, source_paths=None, embeddings=None)``  precomputed document vectors
* ``search(query, top_k, query_vector=None)``               precomputed query vector
* ``search_many(queries, top_k, query_vectors=None)``       one batched launch chain for B queries
* ``group=``        torch.distributed process group: the corpus is sharded by document over its ranks.  Every rank
                    calls ``index`` / ``search`` with the SAME arguments, keeps its contiguous doc range (embedding rows
                    + postings under corpus-global term ids / idf / avgdl) and returns the SAME, unsharded-identical
                    results (C2 stats all-reduce + C1 top-k all-gather inside the engine)
* ``attach_index(shard, documents, query_term_ids=...)``    serve a prebuilt ``DeviceIndex`` (``DeviceIndex.load``
                    checkpoint or a synthetic shard) instead of calling ``index``
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .core import DocTable, Searcher, default_encoder
from .bm25 import BM25
from .engine import QueryBatch
from .extractor import preprocess_text

SearchResult = Tuple[float, str, int]  # (score, content, doc_id)


@dataclass
class PipelineResult:
    """Result from a search pipeline (pipelines.py:24-30)."""
    query: str
    results: List[Dict[str, Any]]
    metadata: Dict[str, Any]
    highlighted: Optional[list] = None


class BasePipeline:
    """pipelines.py:33-59.  Highlighting is outside the hot path."""

    def __init__(self, db_path: str = "index.duckdb", enable_highlighting: bool = False, *,
                 encoder=None, device=None, dense_mode: str = "exact", index_build: str = "host", group=None):
        if enable_highlighting:
            raise NotImplementedError("highlighting is outside the B200 hot path (string post-processing)")
        self.db_path = db_path
        self.docs_df = None
        self.vectors = None
        self.searcher = None
        self.highlighter = None
        self._encoder = encoder
        self._device = device
        self._dense_mode = dense_mode
        self._index_build = index_build
        self._group = group

    def index(self, documents: List[str], **kwargs):
        raise NotImplementedError

    def search(self, query: str, top_k: int = 5, **kwargs) -> PipelineResult:
        raise NotImplementedError

    # ---- shared helpers -----------------------------------------------------------------------
    def _encoder_or_default(self):
        if self._encoder is None:
            self._encoder = default_encoder()
        return self._encoder

    def _index_dense(self, documents: List[str], embeddings=None):
        """Indexer.index_documents (indexer.py:245-285): whitespace-normalised content, doc_id = row,
        float32 [N, d] vectors -- then the Searcher that owns the device copy."""
        contents = [preprocess_text(d) for d in documents]
        self.docs_df = DocTable(contents)
        if embeddings is None:
            vec = self._encoder_or_default().encode(contents)
        else:
            vec = embeddings
        self.vectors = np.array(vec, dtype=np.float32)
        if self.vectors.ndim != 2 or self.vectors.shape[0] != len(documents):
            raise ValueError("embeddings must be [len(documents), dim]")
        self.searcher = Searcher(db_path=self.db_path, encoder=self._encoder, device=self._device,
                                 dense_mode=self._dense_mode, group=self._group)
        self.searcher.attach(self.docs_df, self.vectors)

    def attach_index(self, shard, documents, *, query_term_ids=None, group=None):
        """Serve a prebuilt device shard (extension): ``shard`` is a ``DeviceIndex`` holding the dense rows and -- for
        the BM25 pipelines -- the CSR (``DeviceIndex.load`` checkpoint, ``synth_device.build_synthetic_shard``);
        ``documents`` is any sequence with ``len`` and ``[i]`` covering ALL docs of the corpus (may be lazy);
        ``query_term_ids(query) -> [term id]`` maps a query string to the shard's term ids (defaults to the
        vocabulary of the last ``fit``).  ``group`` as in the constructor."""
        from .engine import SearchEngine
        group = group if group is not None else self._group
        self._group = group
        self.documents = documents
        self.docs_df = DocTable.__new__(DocTable)
        self.docs_df.contents = documents
        self.vectors = None
        self.searcher = Searcher(db_path=self.db_path, encoder=self._encoder, device=shard.device,
                                 dense_mode=self._dense_mode, group=group)
        self.searcher.shard = shard
        self.searcher.engine = SearchEngine(shard, group=group, dense_mode=self._dense_mode)
        self.searcher.doc_range = (shard.doc_base, shard.doc_base + shard.n_docs)
        self.searcher._attached = (self.docs_df, None)
        bm = getattr(self, "bm25", None)
        if bm is not None and shard.has_bm25:
            bm.shard, bm.engine, bm.group, bm.doc_base = shard, self.searcher.engine, group, shard.doc_base
            bm.doc_count = shard.n_docs_global
            bm.avg_doc_len = shard.avgdl
            if query_term_ids is not None:
                bm.stats = type("AttachedVocab", (), {"query_term_ids": staticmethod(query_term_ids)})()

    def _query_vectors(self, queries: Sequence[str], query_vectors=None) -> np.ndarray:
        if query_vectors is not None:
            return np.ascontiguousarray(query_vectors, dtype=np.float32).reshape(len(queries), -1)
        return np.asarray(self._encoder_or_default().encode(list(queries)), dtype=np.float32)

    def _highlight_results(self, results, query):
        return None


# ------------------------------------------------------------------------------------------ basic
class BasicPipeline(BasePipeline):
    """pipelines.py:62-103 -- Searcher.search with (semantic_weight, 1 - semantic_weight)."""

    def __init__(self, db_path: str = "index.duckdb", semantic_weight: float = 0.7,
                 enable_highlighting: bool = False, **ext):
        super().__init__(db_path, enable_highlighting, **ext)
        self.semantic_weight = semantic_weight
        self.lexical_weight = 1.0 - semantic_weight

    def index(self, documents: List[str], source_paths: Optional[List[str]] = None, *, embeddings=None):
        self._index_dense(documents, embeddings)

    def search(self, query: str, top_k: int = 5, *, query_vector=None) -> PipelineResult:
        qv = None if query_vector is None else np.asarray(query_vector, np.float32)[None, :]
        return self.search_many([query], top_k, query_vectors=qv)[0]

    def search_many(self, queries: Sequence[str], top_k: int = 5, *, query_vectors=None) -> List[PipelineResult]:
        """One dense pass + one fuse / select chain for the batch; the lexical kernel runs once per query into the
        rows of a device matrix (Searcher.search_many)."""
        if self.searcher is None:
            raise AttributeError("'NoneType' object has no attribute 'search'")       # pipelines.py:87
        qv = self._query_vectors(queries, query_vectors)
        res = self.searcher.search_many(queries, self.docs_df, self.vectors, top_k=top_k,
                                        semantic_weight=self.semantic_weight, lexical_weight=self.lexical_weight,
                                        query_vectors=qv)
        return [PipelineResult(
            query=q,
            results=[{"score": s, "content": c, "doc_id": d} for s, c, d in r],
            metadata={"pipeline": "basic", "weights": {"semantic": self.semantic_weight}},
            highlighted=None) for q, r in zip(queries, res)]


# ------------------------------------------------------------------------------------------ bm25
class BM25Pipeline(BasePipeline):
    """pipelines.py:253-280 -- pure lexical ranking; no embeddings, no Searcher."""

    def __init__(self, db_path: str = "index.duckdb", k1: float = 1.5, b: float = 0.75, **ext):
        super().__init__(db_path, **ext)
        self.bm25 = BM25(k1=k1, b=b, device=self._device, index_build=self._index_build, group=self._group)
        self.documents: List[str] = []

    def index(self, documents: List[str], source_paths: List[str] = None):
        self.documents = documents
        self.bm25.fit(documents)

    def search(self, query: str, top_k: int = 5) -> PipelineResult:
        return self.search_many([query], top_k)[0]

    def search_many(self, queries: Sequence[str], top_k: int = 5) -> List[PipelineResult]:
        hits = self.bm25.search_many(queries, top_k)
        meta = {"pipeline": "bm25", "k1": self.bm25.k1, "b": self.bm25.b}
        return [PipelineResult(
            query=q,
            results=[{"score": score, "content": self.documents[idx], "doc_id": idx} for idx, score in h],
            metadata=dict(meta)) for q, h in zip(queries, hits)]


# ------------------------------------------------------------------------------------------ hybrid_bm25
class HybridBM25Pipeline(BasePipeline):
    """pipelines.py:283-357 -- min-max cosine * semantic_weight + BM25/max * bm25_weight.

    Weights are not validated (as in the reference); BM25 parameters are the defaults.  Scores come
    back as ``np.float32`` and ``content`` is the RAW document, exactly like the reference.
    """

    def __init__(self, db_path: str = "index.duckdb", semantic_weight: float = 0.6, bm25_weight: float = 0.4,
                 **ext):
        super().__init__(db_path, **ext)
        self.semantic_weight = semantic_weight
        self.bm25_weight = bm25_weight
        self.bm25 = BM25(device=self._device, index_build=self._index_build, group=self._group)
        self.documents: List[str] = []

    def index(self, documents: List[str], source_paths: List[str] = None, *, embeddings=None):
        self.documents = documents
        self._index_dense(documents, embeddings)
        self.bm25.fit(documents, shard=self.searcher.shard)      # same device shard holds both halves

    def search(self, query: str, top_k: int = 5, *, query_vector=None) -> PipelineResult:
        qv = None if query_vector is None else np.asarray(query_vector, np.float32)[None, :]
        return self.search_many([query], top_k, query_vectors=qv)[0]

    def search_many(self, queries: Sequence[str], top_k: int = 5, *, query_vectors=None) -> List[PipelineResult]:
        if self.searcher is None:
            # reference: 'NoneType' object has no attribute 'search' (pipelines.py:317)
            raise AttributeError("'NoneType' object has no attribute 'search'")
        n = len(self.documents)
        if n == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        qv = self._query_vectors(queries, query_vectors)
        qb = QueryBatch(vectors=qv, term_ids=[self.bm25.stats.query_term_ids(q) for q in queries])
        k = min(int(top_k), n) if top_k >= 0 else max(n + int(top_k), 0)     # list slice semantics
        out = []
        if k == 0:
            sc = np.zeros((len(queries), 0), np.float32)
            ids = np.zeros((len(queries), 0), np.int64)
        else:
            s_dev, i_dev = self.searcher.engine.search_hybrid_bm25(qb, k, self.semantic_weight, self.bm25_weight)
            sc, ids = s_dev.cpu().numpy(), i_dev.cpu().numpy()
        meta = {"pipeline": "hybrid_bm25", "semantic_weight": self.semantic_weight,
                "bm25_weight": self.bm25_weight}
        docs = self.documents
        for qi, q in enumerate(queries):
            # iterating a float32 array yields np.float32 scalars (the reference's score type, pipelines.py:347); ids as ints
            out.append(PipelineResult(
                query=q,
                results=[{"score": s, "content": docs[d], "doc_id": d}
                         for s, d in zip(list(sc[qi]), ids[qi].tolist()) if d >= 0],
                metadata=dict(meta), highlighted=None))
        return out


# ------------------------------------------------------------------------------------------ multi_stage
class MultiStagePipeline(BasePipeline):
    """pipelines.py:435-511 -- dense top-stage1_k -> BM25 on those -> external reranker."""

    def __init__(self, db_path: str = "index.duckdb", stage1_k: int = 100, stage2_k: int = 20, final_k: int = 5,
                 *, reranker=None, **ext):
        super().__init__(db_path, **ext)
        self.stage1_k = stage1_k
        self.stage2_k = stage2_k
        self.final_k = final_k
        self.bm25 = BM25(device=self._device, index_build=self._index_build, group=self._group)
        self.documents: List[str] = []
        self._reranker = reranker

    def index(self, documents: List[str], source_paths: List[str] = None, *, embeddings=None):
        self.documents = documents
        self._index_dense(documents, embeddings)
        self.bm25.fit(documents, shard=self.searcher.shard)

    def stages_1_2(self, queries: Sequence[str], *, query_vectors=None):
        """Stages 1-2 for a batch (pipelines.py:474-487).  -> per query list of (bm25 float, content, doc_id)."""
        if self.searcher is None:
            raise AttributeError("'NoneType' object has no attribute 'search'")
        n = len(self.documents)
        if n == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        eng = self.searcher.engine
        qv = self._query_vectors(queries, query_vectors)
        terms = [self.bm25.stats.query_term_ids(q) for q in queries]
        k1 = min(self.stage1_k, n)
        _, ids1 = eng.search_semantic(QueryBatch(vectors=qv), k1, 1.0)          # stage 1: min-max cosine
        ids1 = ids1.clone()
        bm = eng.bm25_score_docs_global(terms, ids1)                             # stage 2: float64 BM25.score
        ids_h, bm_h = ids1.cpu().numpy(), bm.cpu().numpy()
        contents = self.docs_df.contents
        # stable sort, descending: ties keep stage-1 rank (pipelines.py:486) -- one vectorised sort for the batch
        rank = np.broadcast_to(np.arange(k1), bm_h.shape)
        order = np.lexsort((rank, -bm_h), axis=-1)[:, :self.stage2_k]
        return [[(float(bm_h[qi][j]), contents[int(ids_h[qi][j])], int(ids_h[qi][j])) for j in order[qi]]
                for qi in range(len(queries))]

    def search(self, query: str, top_k: int = None, *, query_vector=None) -> PipelineResult:
        qv = None if query_vector is None else np.asarray(query_vector, np.float32)[None, :]
        return self.search_many([query], top_k, query_vectors=qv)[0]

    def search_many(self, queries: Sequence[str], top_k: int = None, *, query_vectors=None) -> List[PipelineResult]:
        """Stages 1-2 for the whole batch in one launch chain (set ``dense_mode="bf16"`` on the pipeline for the
        tcgen05 GEMM at large batch), stage 3 per query through the reranker hook."""
        top_k = top_k or self.final_k
        stage2 = self.stages_1_2(queries, query_vectors=query_vectors)
        if self._reranker is None:
            from .core import CrossEncoderReranker      # stage 3 is transformer inference: external hook
            self._reranker = CrossEncoderReranker()
        out = []
        for q, cand in zip(queries, stage2):
            final = self._reranker.rerank(q, cand, top_k=top_k)
            out.append(PipelineResult(
                query=q,
                results=[{"score": s, "content": c, "doc_id": d, "stage": "final"} for s, c, d in final],
                metadata={"pipeline": "multi_stage", "stage1_k": self.stage1_k, "stage2_k": self.stage2_k,
                          "final_k": top_k},
                highlighted=None))
        return out


# ------------------------------------------------------------------------------------------ diversity
class DiversityPipeline(BasePipeline):
    """pipelines.py:514-613 -- 4*top_k candidates from the default 0.7/0.3 hybrid, then greedy MMR."""

    def __init__(self, db_path: str = "index.duckdb", lambda_param: float = 0.5, **ext):
        super().__init__(db_path, **ext)
        self.lambda_param = lambda_param

    def index(self, documents: List[str], source_paths: Optional[List[str]] = None, *, embeddings=None):
        self._index_dense(documents, embeddings)

    def _mmr(self, query_embedding, doc_embeddings, doc_scores, top_k: int) -> List[int]:
        """pipelines.py:531-569 on the device; ``doc_embeddings`` may be row ids into the index."""
        eng = self.searcher.engine
        rows = np.asarray(doc_embeddings)
        if rows.ndim != 1:
            raise ValueError("_mmr takes candidate row ids of the indexed matrix")
        cand = torch.as_tensor(rows, dtype=torch.int64, device=eng.device)[None, :]
        rel = torch.as_tensor(np.asarray(doc_scores, np.float64), device=eng.device)[None, :]
        k = min(int(top_k), len(rows))
        if k <= 0:
            return []
        sel = eng.mmr(cand, rel, self.lambda_param, k).cpu().numpy()[0]
        return [int(i) for i in sel if i >= 0]

    def search(self, query: str, top_k: int = 5, *, query_vector=None) -> PipelineResult:
        qv = None if query_vector is None else np.asarray(query_vector, np.float32)[None, :]
        return self.search_many([query], top_k, query_vectors=qv)[0]

    def search_many(self, queries: Sequence[str], top_k: int = 5, *, query_vectors=None) -> List[PipelineResult]:
        """Batched diversity search: candidates for every query from ONE batched Searcher pass (default 0.7 / 0.3
        hybrid, pipelines.py:573-578), then one ``hs_mmr`` launch over [B, C] candidates (one CTA per query).
        Doc-sharded: retrieval is sharded (C2 + C1 in the engine), the candidate rows are assembled with one
        all-reduce, the greedy MMR itself is data-parallel over the queries."""
        if self.searcher is None:
            raise AttributeError("'NoneType' object has no attribute 'search'")       # pipelines.py:573
        qv = self._query_vectors(queries, query_vectors)
        cands = self.searcher.search_many(queries, self.docs_df, self.vectors, top_k=top_k * 4, query_vectors=qv)
        B = len(queries)
        Cmax = max((len(c) for c in cands), default=0)
        k = min(int(top_k), Cmax)
        out: List[Optional[PipelineResult]] = [None] * B
        if Cmax == 0 or k <= 0:
            return [PipelineResult(query=q, results=[], metadata={}) if not c else
                    PipelineResult(query=q, results=[], metadata={"pipeline": "diversity", "lambda": self.lambda_param,
                                                                  "method": "mmr"}) for q, c in zip(queries, cands)]
        ids = np.full((B, Cmax), -1, np.int64)
        rel = np.zeros((B, Cmax), np.float64)
        for b, c in enumerate(cands):
            if not c:
                continue
            sc = np.array([s for s, _, _ in c])
            # pipelines.py:589 in float64
            rel[b, :len(c)] = (sc - sc.min()) / (sc.max() - sc.min() + 1e-8)
            ids[b, :len(c)] = [d for _, _, d in c]
        sel = self._mmr_batch(ids, rel, k)
        for b, (q, c) in enumerate(zip(queries, cands)):
            if not c:
                out[b] = PipelineResult(query=q, results=[], metadata={})          # pipelines.py:580-581
                continue
            picked = [int(i) for i in sel[b] if i >= 0][:min(int(top_k), len(c))]
            out[b] = PipelineResult(
                query=q,
                results=[{"score": c[i][0], "content": c[i][1], "doc_id": c[i][2], "diversity_rank": rank}
                         for rank, i in enumerate(picked)],
                metadata={"pipeline": "diversity", "lambda": self.lambda_param, "method": "mmr"})
        return out

    def _mmr_batch(self, doc_ids: np.ndarray, rel: np.ndarray, k: int) -> np.ndarray:
        """int32 [B, k] positions into each query's candidate list.  doc_ids int64 [B, C] GLOBAL ids (-1 = padding)."""
        eng = self.searcher.engine
        dev = eng.device
        B, Cn = doc_ids.shape
        if eng.world == 1:
            local = torch.from_numpy(doc_ids).to(dev) - eng.shard.doc_base
            local = torch.where(local >= 0, local, torch.full_like(local, -1))
            return eng.mmr(local, torch.from_numpy(rel).to(dev), self.lambda_param, k).cpu().numpy()
        return eng.mmr_sharded(doc_ids, rel, self.lambda_param, k)


# ------------------------------------------------------------------------------------------ factory
_PIPELINES = {
    "basic": BasicPipeline,
    "bm25": BM25Pipeline,
    "hybrid_bm25": HybridBM25Pipeline,
    "multi_stage": MultiStagePipeline,
    "diversity": DiversityPipeline,
}
# names the reference also accepts (pipelines.py:633-642) but which are outside the hot path
_OUT_OF_SCOPE = ("chunked", "reranked", "rag")


def create_pipeline(pipeline_type: str = "basic", **kwargs) -> BasePipeline:
    """pipelines.py:617-646."""
    if pipeline_type in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"pipeline {pipeline_type!r} exists in the reference but is outside the B200 hot path "
            "(needs chunker / cross-encoder / LLM); use the reference for it")
    if pipeline_type not in _PIPELINES:
        names = ["basic", "chunked", "reranked", "bm25", "hybrid_bm25", "rag", "multi_stage", "diversity"]
        raise ValueError(f"Unknown pipeline: {pipeline_type}. Choose from {names}")
    return _PIPELINES[pipeline_type](**kwargs)
