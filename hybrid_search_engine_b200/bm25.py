"""``BM25`` -- mirror of the reference's bm25.py API with the scoring on the device.

``fit`` tokenises on the host (term identity, bm25.py:58-67) and uploads a CSR inverted index;
``score`` / ``score_batch`` / ``search`` run the hs_b200 BM25 kernels.  ``BM25Okapi`` is the same class
(bm25.py:145-147).  ``BM25Plus`` (bm25.py:150-179) is dense (every doc gets ``idf * delta`` per known
query term) and is used by no pipeline; its ``score_batch`` runs a dedicated tile kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .engine import QueryBatch, SearchEngine
from .index import DeviceIndex, LexicalStats


class BM25:
    def __init__(self, k1: float = 1.5, b: float = 0.75, remove_stopwords: bool = True, *, device=None,
                 index_build: str = "host", group=None):
        """``index_build`` (extension): "host" keeps the vocabulary strings (``doc_freqs`` / ``idf`` dicts as in
        the reference); "device" tokenises and builds the CSR on the GPU (index_build.py) and keeps only term
        hashes -- same scores, ~30x faster ``fit`` on large corpora; the ``doc_freqs`` / ``idf`` dicts are then
        built lazily on first access."""
        if index_build not in ("host", "device"):
            raise ValueError(f"index_build must be 'host' or 'device', got {index_build!r}")
        self.k1 = k1
        self.b = b
        self.remove_stopwords = remove_stopwords
        self._device = device
        self.index_build = index_build
        # group (extension): torch.distributed process group -- the corpus is sharded by document over its ranks
        # (every rank is given the SAME full document list and keeps its contiguous range); scores / rankings are
        # identical to the unsharded index on every rank
        self.group = group
        self.doc_base = 0
        self.stats = LexicalStats(remove_stopwords)
        self.shard: Optional[DeviceIndex] = None
        self.engine: Optional[SearchEngine] = None
        self.doc_count = 0
        self.avg_doc_len = 0.0
        self.doc_lengths: List[int] = []

    # ---- corpus statistics in the reference's shapes (built lazily, host side) -----------------
    def _vocab(self) -> Dict[str, int]:
        return self.stats.vocab          # index_build="device": built lazily on the host (index_build.py)

    @property
    def doc_freqs(self) -> Dict[str, int]:
        return {t: int(self.stats.df[i]) for t, i in self._vocab().items()}

    @property
    def idf(self) -> Dict[str, float]:
        if self.shard is None or self.shard.idf_host is None:
            return {}
        return {t: float(self.shard.idf_host[i]) for t, i in self._vocab().items()}

    def fit(self, documents: Sequence[str], *, shard: Optional[DeviceIndex] = None):
        """bm25.py:45-81.  ``shard``: attach to an existing device shard (the pipeline's dense one)."""
        if self.group is not None:
            return self._fit_sharded(documents, shard)
        if self.index_build == "device":
            return self._fit_device(documents, shard)
        st = self.stats = LexicalStats(self.remove_stopwords).fit(documents)
        self.doc_count = st.doc_count
        self.avg_doc_len = st.avg_doc_len
        self.doc_lengths = st.doc_lengths.tolist()
        if st.doc_count == 0:
            self.shard = self.engine = None
            return
        if not torch.cuda.is_available():
            raise _lib.HsError("no CUDA device: BM25 scoring has no CPU fallback")
        if shard is None:
            dev = torch.device(self._device) if self._device is not None else \
                torch.device("cuda", torch.cuda.current_device())
            shard = DeviceIndex(dev, st.doc_count)
        self.shard = shard
        shard.set_bm25(torch.from_numpy(st.indptr), torch.from_numpy(st.postings.view(np.int32)),
                       torch.from_numpy(st.doc_lengths.astype(np.uint32).view(np.int32)), float(st.avg_doc_len),
                       st.df, st.doc_count, self.k1, self.b)
        self.engine = SearchEngine(shard)

    def _fit_sharded(self, documents: Sequence[str], shard: Optional[DeviceIndex]):
        """Doc-sharded fit: this rank tokenises documents[lo:hi], the vocabulary / df / avgdl are merged over the
        group (LexicalStats.merge_across), the device CSR holds this rank's postings under GLOBAL term ids."""
        import torch.distributed as dist
        from .parallel import shard_bounds
        if self.index_build != "host":
            raise NotImplementedError("group= needs index_build='host' (term identity by string across ranks)")
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(len(documents), world, rank)
        st = self.stats = LexicalStats(self.remove_stopwords).fit(documents[lo:hi]).merge_across(self.group)
        self.doc_count = st.doc_count
        self.avg_doc_len = st.avg_doc_len
        self.doc_lengths = st.doc_lengths.tolist()
        self.doc_base = lo
        if st.doc_count == 0:
            self.shard = self.engine = None
            return
        if not torch.cuda.is_available():
            raise _lib.HsError("no CUDA device: BM25 scoring has no CPU fallback")
        if shard is None:
            dev = torch.device(self._device) if self._device is not None else \
                torch.device("cuda", torch.cuda.current_device())
            shard = DeviceIndex(dev, hi - lo, doc_base=lo)
        if shard.n_docs != hi - lo or shard.doc_base != lo:
            raise ValueError("BM25.fit: the dense shard does not cover this rank's doc range")
        self.shard = shard
        shard.set_bm25(torch.from_numpy(st.indptr), torch.from_numpy(st.postings.view(np.int32)),
                       torch.from_numpy(st.local_doc_lengths.astype(np.uint32).view(np.int32)), float(st.avg_doc_len),
                       st.df, st.doc_count, self.k1, self.b,
                       max_dl=int(st.doc_lengths.max()) if st.doc_count else 0)
        self.engine = SearchEngine(shard, group=self.group)

    def _fit_device(self, documents: Sequence[str], shard: Optional[DeviceIndex]):
        from .index_build import DeviceLexicalStats
        if not torch.cuda.is_available():
            raise _lib.HsError("no CUDA device: the device index build has no CPU fallback")
        if len(documents) == 0:
            self.stats = LexicalStats(self.remove_stopwords).fit(documents)
            self.doc_count, self.avg_doc_len, self.doc_lengths = 0, 0, []
            self.shard = self.engine = None
            return
        dev = shard.device if shard is not None else (
            torch.device(self._device) if self._device is not None else
            torch.device("cuda", torch.cuda.current_device()))
        st = self.stats = DeviceLexicalStats(dev, self.remove_stopwords).fit(documents)
        self.doc_count = st.doc_count
        self.avg_doc_len = st.avg_doc_len
        self.doc_lengths = st.doc_lengths.cpu().tolist()
        if shard is None:
            shard = DeviceIndex(dev, st.doc_count)
        self.shard = shard
        shard.set_bm25(st.indptr, st.postings, st.doc_lengths, float(st.avg_doc_len), st.df, st.doc_count,
                       self.k1, self.b, max_dl=st.max_dl)
        self.engine = SearchEngine(shard)

    # ---- scoring --------------------------------------------------------------------------------
    def score_batch_many(self, queries: Sequence[str]) -> np.ndarray:
        """float32 [B, N] BM25 scores (bm25.py:114-127 for a batch of queries)."""
        if self.doc_count == 0:
            return np.zeros((len(queries), 0), np.float32)
        eng = self.engine
        terms = [self.stats.query_term_ids(q) for q in queries]
        with torch.cuda.device(eng.device):
            qt, qi, qo = eng.upload_terms(terms)
            sc = eng.bm25_score(qt, qi, qo, len(queries), None, plus_delta=self._plus_delta())
            return sc.cpu().numpy()

    def _plus_delta(self):
        return None

    def score_batch(self, query: str) -> np.ndarray:
        return self.score_batch_many([query])[0]

    def score(self, query: str, doc_idx: int) -> float:
        """bm25.py:83-112 -- float64, unrounded."""
        if doc_idx < 0:
            doc_idx += self.doc_count
        if not 0 <= doc_idx < self.doc_count:
            raise IndexError("list index out of range")
        eng = self.engine
        ids = torch.tensor([[doc_idx]], dtype=torch.int64, device=eng.device)
        return float(eng.bm25_score_docs_global([self.stats.query_term_ids(query)], ids,
                                                plus_delta=self._plus_delta()).cpu()[0, 0])

    def search_many(self, queries: Sequence[str], top_k: int = 10) -> List[List[tuple]]:
        """bm25.py:129-142 per query -> [(doc_idx, score)], canonical order (score desc, doc_idx asc)."""
        if self.doc_count == 0:
            return [[] for _ in queries]
        n = self.doc_count
        k = min(int(top_k), n) if top_k >= 0 else max(n + int(top_k), 0)
        if k == 0:
            return [[] for _ in queries]
        qb = QueryBatch(term_ids=[self.stats.query_term_ids(q) for q in queries])
        sc, ids = self.engine.search_bm25(qb, k, plus_delta=self._plus_delta())
        sc, ids = sc.cpu().numpy(), ids.cpu().numpy()
        return [[(int(i), float(s)) for s, i in zip(sc[q], ids[q]) if i >= 0] for q in range(len(queries))]

    def search(self, query: str, top_k: int = 10) -> List[tuple]:
        return self.search_many([query], top_k)[0]


class BM25Okapi(BM25):
    """bm25.py:145-147."""
    pass


class BM25Plus(BM25):
    """bm25.py:150-179 -- adds ``delta`` inside the idf product, for every doc and every known query token
    (tf = 0 included), so scores are dense.  ``score_batch`` / ``search`` run the dense tile kernel (+ the device
    top-k select), ``score`` the float64 per-doc kernel (used by no pipeline in the reference)."""

    def __init__(self, k1: float = 1.5, b: float = 0.75, delta: float = 1.0, **kwargs):
        super().__init__(k1=k1, b=b, **kwargs)
        if k1 < 0 or not 0.0 <= b <= 1.0:
            raise ValueError("BM25Plus on the device needs k1 >= 0 and 0 <= b <= 1")
        self.delta = delta

    def _plus_delta(self):
        return self.delta

    # score (float64, hs_bm25plus_score_docs) and search_many (dense tile kernel + device top-k select) are the
    # base-class methods: _plus_delta() switches every kernel call to the BM25+ formula
