"""``Searcher`` -- mirror of the reference's core.py:112-285 for the hot path.

``Searcher.search(query, docs_df, vectors, top_k, semantic_weight, lexical_weight)`` keeps the
reference's signature, defaults (0.7 / 0.3, core.py:229-230), weight check (core.py:232-233) and
return type ``[(float score, content, doc_id)]``; the arithmetic (cosine scan, min-max, weighted
sum, top-k) runs in the hs_b200 kernels.  QueryMemory / DuckDB logging (core.py:20-109,280-281) is
bookkeeping outside the hot path and is not performed.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import QueryBatch, SearchEngine
from .index import DeviceIndex


class _Column:
    def __init__(self, values):
        self._values = values

    def to_list(self):
        return list(self._values)

    def __len__(self):
        return len(self._values)

    def __getitem__(self, i):
        return self._values[i]


class DocTable:
    """The two columns of the reference's polars frame the hot path reads (core.py:240-241):
    ``table['content'].to_list()`` and ``table['doc_id'].to_list()``; doc_id == row (indexer.py:220)."""

    def __init__(self, contents: Sequence[str]):
        self.contents = list(contents)

    def __getitem__(self, col):
        if col == "content":
            return _Column(self.contents)
        if col == "doc_id":
            return _Column(range(len(self.contents)))
        raise KeyError(col)

    def __len__(self):
        return len(self.contents)


def default_encoder():
    """The reference hard-wires SentenceTransformer('all-MiniLM-L6-v2') (core.py:134, indexer.py:91)."""
    try:
        from sentence_transformers import SentenceTransformer
    except ImportError as e:
        raise ImportError(
            "no encoder given and sentence-transformers is not installed; pass encoder= (an object with "
            "encode(list[str]) -> float32 [n, d]) or precomputed embeddings / query vectors") from e
    return SentenceTransformer("all-MiniLM-L6-v2")


class CrossEncoderReranker:
    """Stage-3 hook of multi_stage (reranker.py:50-89): scores (query, content) pairs with an external
    model and re-sorts.  Transformer inference is outside the hot path; ``predict`` is injectable."""

    def __init__(self, model_name: str = "cross-encoder/ms-marco-MiniLM-L-6-v2", device=None, predict=None):
        self.model_name = model_name
        self._predict = predict
        self._device = device

    def rerank(self, query: str, results, top_k: Optional[int] = None):
        if not results:
            return results
        if self._predict is None:
            try:
                from sentence_transformers import CrossEncoder
            except ImportError as e:
                raise ImportError("multi_stage stage 3 needs a cross-encoder: pass reranker= to the pipeline "
                                  "or install sentence-transformers") from e
            model = CrossEncoder(self.model_name, device=self._device)
            self._predict = lambda pairs: model.predict(pairs, show_progress_bar=False)
        scores = self._predict([(query, content) for _, content, _ in results])
        ranked = [(float(s), c, d) for s, (_, c, d) in zip(scores, results)]
        ranked.sort(key=lambda x: x[0], reverse=True)
        return ranked[:top_k] if top_k else ranked


class Searcher:
    """core.py:112-285.  Holds the device shard for the (docs_df, vectors) it was last given."""

    def __init__(self, model_name: str = "all-MiniLM-L6-v2", db_path: str = "index.duckdb", use_faiss: bool = False,
                 faiss_index_path: str = "index.faiss", enable_query_memory: bool = True, *, encoder=None,
                 device=None, dense_mode: str = "exact", lexical_scorer=None, group=None):
        # use_faiss=True (core.py:148-168): the rows of `faiss_index_path` (a FAISS IndexFlatIP file,
        # L2-normalised at add time) replace `vectors`; only the top min(2k, N) inner products keep their
        # score.  Not reachable from the pipelines (they construct Searcher(db_path=...)).  PARITY UNPINNED:
        # faiss is not installable here; cosine of the stored rows stands in for its sgemm inner product.
        self.model = encoder
        self.db_path = db_path
        self.use_faiss = bool(use_faiss)
        self.faiss_index_path = faiss_index_path
        self._faiss_rows = None
        self.query_memory = None
        self._device = device
        self._dense_mode = dense_mode
        self._lexical_scorer = lexical_scorer
        # group (extension): doc-sharded over the ranks of a torch.distributed group; every rank is given the same
        # (docs_df, vectors) and keeps its contiguous doc range; results are identical on every rank
        self.group = group
        self.doc_range = (0, 0)
        self.shard: Optional[DeviceIndex] = None
        self.engine: Optional[SearchEngine] = None
        self._attached = None
        self._doc_lists = None

    # ------------------------------------------------------------------
    def attach(self, docs_df, vectors: np.ndarray, *, force: bool = False):
        """Upload ``vectors`` (float32 [N, d], indexer.py:285) once; later searches reuse the shard."""
        # the cache holds strong references and compares by identity: a recycled id() of a freed array can never
        # alias it.  In-place edits of an attached array are not seen -- call attach(..., force=True) (the
        # reference re-reads `vectors` on every search; here it lives in HBM)
        if (not force and self._attached is not None and self._attached[0] is docs_df
                and self._attached[1] is vectors):
            return
        key = (docs_df, vectors)
        self._doc_lists = None
        if not torch.cuda.is_available():
            raise _lib.HsError("no CUDA device: the hybrid scoring path has no CPU fallback")
        dev = torch.device(self._device) if self._device is not None else torch.device("cuda", torch.cuda.current_device())
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        n = vectors.shape[0]
        lo, hi = 0, n
        if self.group is not None:
            import torch.distributed as dist
            from .parallel import shard_bounds
            lo, hi = shard_bounds(n, dist.get_world_size(self.group), dist.get_rank(self.group))
        self.doc_range = (lo, hi)
        self.shard = DeviceIndex(dev, hi - lo, doc_base=lo)
        if hi > lo:
            self.shard.set_dense(torch.from_numpy(vectors[lo:hi]).to(dev))
        else:
            self.shard.dim = vectors.shape[1] if vectors.ndim == 2 else 0
        self.engine = SearchEngine(self.shard, group=self.group, dense_mode=self._dense_mode)
        self._attached = key

    def _lexical_scorer_obj(self):
        if self._lexical_scorer is None:
            from .lexical import LexicalScorer
            self._lexical_scorer = LexicalScorer(self.shard.device)
        return self._lexical_scorer

    def _lexical_scores(self, query: str, docs: List[str]) -> np.ndarray:
        """core.py:178-197 (partial_ratio + token overlap), float32 [N] on the host."""
        return self._lexical_scorer_obj().scores(query, docs)

    def search(self, query: str, docs_df, vectors: np.ndarray, top_k: int = 5,
               semantic_weight: Optional[float] = None, lexical_weight: Optional[float] = None,
               use_learned_weights: bool = False, *, query_vector=None) -> List[Tuple[float, str, int]]:
        """core.py:199-285 for one query."""
        qv = None if query_vector is None else np.asarray(query_vector, dtype=np.float32)[None, :]
        return self.search_many([query], docs_df, vectors, top_k, semantic_weight, lexical_weight,
                                use_learned_weights, query_vectors=qv)[0]

    def search_many(self, queries: Sequence[str], docs_df, vectors: np.ndarray, top_k: int = 5,
                    semantic_weight: Optional[float] = None, lexical_weight: Optional[float] = None,
                    use_learned_weights: bool = False, *, query_vectors=None) -> List[List[Tuple[float, str, int]]]:
        """``search`` for a batch (extension; the reference's batch endpoint api.py:429-445 is a serial loop): ONE dense
        pass and one fuse / select chain for all queries, the lexical kernel once per query into rows of a device
        matrix.  Same results as one ``search`` per query."""
        if use_learned_weights:
            raise NotImplementedError("learned weights need QueryMemory/DuckDB (core.py:224-226): out of scope")
        semantic_weight = semantic_weight if semantic_weight is not None else 0.7
        lexical_weight = lexical_weight if lexical_weight is not None else 0.3
        if not np.isclose(semantic_weight + lexical_weight, 1.0):
            raise ValueError("semantic_weight and lexical_weight must sum to 1.0")
        if self.use_faiss and self._faiss_rows is None:
            import os
            if os.path.exists(self.faiss_index_path):          # core.py:148-157: silently brute-force otherwise
                from .faiss_io import read_index_flat
                self._faiss_rows, _ = read_index_flat(self.faiss_index_path)
        use_faiss = self.use_faiss and self._faiss_rows is not None
        self.attach(docs_df, self._faiss_rows if use_faiss else vectors)
        # column lists are materialised once per docs_df object (they also key the lexical scorer's device copy)
        if self._doc_lists is None or self._doc_lists[0] is not docs_df:
            docs_all = docs_df["content"].to_list()
            lo, hi = self.doc_range
            local = docs_all if (lo, hi) == (0, len(docs_all)) else docs_all[lo:hi]
            self._doc_lists = (docs_df, docs_all, docs_df["doc_id"].to_list(), local)
        _, docs, doc_ids, local_docs = self._doc_lists
        n = len(docs)
        if n == 0:
            # utils.py:67 on an empty array
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        B = len(queries)
        if query_vectors is None:
            if self.model is None:
                self.model = default_encoder()
            query_vectors = self.model.encode(list(queries))
        q = np.asarray(query_vectors, dtype=np.float32).reshape(B, -1)
        k = min(int(top_k), n) if top_k >= 0 else max(n + int(top_k), 0)
        if k == 0 or B == 0:
            return [[] for _ in range(B)]
        eng = self.engine
        if use_faiss:
            out = []
            for b in range(B):                                  # faiss-style retrieval is per query (not a pipeline path)
                lex = None
                if lexical_weight != 0.0:
                    lex = self._lexical_scorer_obj().scores_device(queries[b], local_docs)[None, :]
                sc, ids = eng.search_faiss_style(QueryBatch(vectors=q[b:b + 1]), lex, k, semantic_weight, lexical_weight)
                out.append((sc.cpu().numpy()[0], ids.cpu().numpy()[0]))
            return [[(float(s), docs[int(i)], doc_ids[int(i)]) for s, i in zip(sc, ids) if i >= 0] for sc, ids in out]
        if lexical_weight == 0.0:
            # lex_norm * 0.0 == +0.0 for every finite lexical vector (core.py:268): skip computing it
            sc, ids = eng.search_semantic(QueryBatch(vectors=q), k, semantic_weight)
            sc, ids = sc.cpu().numpy(), ids.cpu().numpy()
        else:
            # lexical vectors are per query and [B, n] float32: bound the device matrix by chunking the batch
            step = max(1, min(B, (1 << 28) // max(len(local_docs), 1)))
            parts = []
            for s in range(0, B, step):
                lex = self._lexical_scorer_obj().scores_device_many(queries[s:s + step], local_docs)
                a, b_ = eng.search_searcher(QueryBatch(vectors=q[s:s + step]), lex, k, semantic_weight, lexical_weight)
                parts.append((a.cpu().numpy(), b_.cpu().numpy()))
            sc = np.concatenate([p[0] for p in parts])
            ids = np.concatenate([p[1] for p in parts])
        return [[(float(s), docs[int(i)], doc_ids[int(i)]) for s, i in zip(sc[b], ids[b]) if i >= 0] for b in range(B)]
