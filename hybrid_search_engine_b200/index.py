"""Device-resident index shard: the HBM layout behind the C-ABI handle ``hs_index``.

HBM layout of one shard (docs ``doc_base .. doc_base + n_docs - 1``):

    vectors   float32 [n_docs, ld]      row-major, ld = dim rounded up to 4, padding columns zero
    vnorm     float32 [n_docs]          f32(sqrt(sum64 v^2)) per row, conformance order
    indptr    int64   [n_terms + 1]     CSR over GLOBAL term ids, restricted to this shard's docs
    postings  uint32  [n_postings, 2]   (doc_id local to the shard, tf), doc ids ascending per term
    dl        uint32  [n_docs]          tokens after stop-word removal (bm25.py:59-60)
    impact    float64 [max_dl+1, tf_cap+1]  tf*(k1+1) / (tf + k1*(1-b+b*dl/avgdl)) per (doc length, tf)
    idf       float64 [n_terms]         host copy too; ln((N - df + .5) / (df + .5) + 1) (bm25.py:81)

Corpus-global statistics (N, df -> idf, avgdl) are replicated on every shard.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def idf_from_df(doc_count: int, df: np.ndarray) -> np.ndarray:
    """bm25.py:76-81 on the host with libm ``math.log`` (one call per distinct df value)."""
    df = np.asarray(df, dtype=np.int64)
    if df.size == 0:
        return np.zeros(0, np.float64)
    uniq, inv = np.unique(df, return_inverse=True)
    vals = np.fromiter((math.log((doc_count - int(d) + 0.5) / (int(d) + 0.5) + 1) for d in uniq),
                       dtype=np.float64, count=len(uniq))
    return vals[inv]


class DeviceIndex:
    """Owns the torch tensors of one shard and the matching ``hs_index`` handle."""

    def __init__(self, device, n_docs: int, doc_base: int = 0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.HsError("DeviceIndex needs a CUDA device: the hot path has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_docs = int(n_docs)
        self.doc_base = int(doc_base)
        h = C.c_void_p()
        check(self.lib.hs_index_create(self.device.index, self.n_docs, self.doc_base, C.byref(h)),
              "hs_index_create")
        self.handle = h
        self.dim = 0
        self.ld = 0
        self.vectors = self.vnorm = None
        self.indptr = self.postings = self.dl = self.impact_table = None
        self.n_terms = 0
        self.avgdl = 0.0
        self.k1, self.b = 1.5, 0.75
        self.idf_host: Optional[np.ndarray] = None      # float64 [n_terms], global
        self.df_host: Optional[np.ndarray] = None       # int64 [n_terms], global

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.hs_index_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ dense part
    def set_dense(self, vectors: torch.Tensor):
        """``vectors`` float32 [n_docs, dim] on this device (indexer.py:285)."""
        if vectors.dim() != 2 or vectors.shape[0] != self.n_docs:
            raise ValueError(f"vectors must be [n_docs={self.n_docs}, dim], got {tuple(vectors.shape)}")
        vectors = vectors.to(self.device, torch.float32)
        dim = vectors.shape[1]
        ld = (dim + 3) // 4 * 4
        if ld != dim:
            padded = torch.zeros((self.n_docs, ld), dtype=torch.float32, device=self.device)
            padded[:, :dim] = vectors
            vectors = padded
        self.vectors = vectors.contiguous()
        self.dim, self.ld = dim, ld
        self.vectors_bf16 = None           # a bf16 copy of an earlier matrix is stale now (ensure_bf16 rebuilds it)
        self.vnorm = torch.empty(self.n_docs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.hs_row_norms(ptr(self.vectors), self.n_docs, dim, ld, ptr(self.vnorm),
                                        stream_ptr(self.device)), "hs_row_norms")
            check(self.lib.hs_index_set_dense(self.handle, ptr(self.vectors), dim, ld, ptr(self.vnorm)),
                  "hs_index_set_dense")

    def ensure_bf16(self):
        """bf16 copy of the matrix with UNIT rows, bf16(v / |v|) (zero rows stay zero), rows padded to a multiple of 64
        elements: the operand of the tcgen05 GEMM, whose accumulator is then the cosine itself."""
        if getattr(self, "vectors_bf16", None) is not None:
            return
        if self.vectors is None:
            raise _lib.HsError("ensure_bf16: no dense matrix")
        ld16 = (self.dim + 63) // 64 * 64
        v16 = torch.zeros((self.n_docs, ld16), dtype=torch.bfloat16, device=self.device)
        step = 1 << 20
        for s in range(0, self.n_docs, step):                   # chunked: no second float32 copy
            vn = self.vnorm[s:s + step]
            inv = torch.where(vn > 0, 1.0 / vn, torch.zeros_like(vn))
            v16[s:s + step, :self.dim] = (self.vectors[s:s + step, :self.dim] * inv[:, None]).to(torch.bfloat16)
        self.vectors_bf16 = v16
        self.ld_bf16 = ld16
        check(self.lib.hs_index_set_dense_bf16(self.handle, ptr(v16), ld16), "hs_index_set_dense_bf16")

    # ------------------------------------------------------------------ lexical part
    def set_bm25(self, indptr: torch.Tensor, postings: torch.Tensor, dl: torch.Tensor, avgdl: float,
                 df_global: np.ndarray, n_docs_global: int, k1: float = 1.5, b: float = 0.75,
                 max_dl: Optional[int] = None):
        """CSR + doc stats.  ``postings`` int32 [P, 2] holding (local doc id, tf) bit patterns."""
        self.indptr = indptr.to(self.device, torch.int64).contiguous()
        self.postings = postings.to(self.device, torch.int32).contiguous().view(-1, 2)
        self.dl = dl.to(self.device, torch.int32).contiguous()
        self.n_terms = self.indptr.numel() - 1
        self.avgdl, self.k1, self.b = avgdl, float(k1), float(b)
        self.df_host = np.asarray(df_global, dtype=np.int64)
        self.n_docs_global = int(n_docs_global)
        self.idf_host = idf_from_df(n_docs_global, self.df_host)
        if max_dl is None:
            max_dl = int(self.dl.max().item()) if self.n_docs else 0
        self.max_dl = int(max_dl)
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            self.impact_table, self.tf_cap = None, 0
            if avgdl > 0 and self.max_dl < (1 << 20):
                # widest table whose shared-memory copy (odd row stride, ~10 000 doubles next to the three tile
                # buffers of bm25_batch_kernel) still fits; else a table of <= 8 MB gathered through L1
                self.tf_cap = 31 if self.max_dl < (1 << 15) else (7 if self.max_dl < (1 << 17) else 0)
                for cap in (31, 15, 7, 3):
                    if (self.max_dl + 1) * ((cap + 1) | 1) <= 10_000:
                        self.tf_cap = cap
                        break
                self.impact_table = torch.empty((self.max_dl + 1) * (self.tf_cap + 1), dtype=torch.float64,
                                                device=self.device)
                check(self.lib.hs_bm25_impact_table(float(avgdl), self.k1, self.b, self.max_dl, self.tf_cap,
                                                    ptr(self.impact_table), st), "hs_bm25_impact_table")
            check(self.lib.hs_index_set_csr(self.handle, ptr(self.indptr), ptr(self.postings), self.n_terms,
                                            self.postings.shape[0]), "hs_index_set_csr")
            check(self.lib.hs_index_set_doc_stats(self.handle, ptr(self.dl), float(avgdl), self.k1, self.b,
                                                  ptr(self.impact_table), self.max_dl, self.tf_cap, st),
                  "hs_index_set_doc_stats")
            self._build_hot_terms(st)

    HOT_MAX_TERMS = 32

    def _build_hot_terms(self, st):
        """Dense contribution vectors for the terms that occur in at least half of this shard's docs (the head of a
        Zipfian vocabulary: ~20 terms carry ~85 % of the postings a query batch touches).  8 bytes per doc and term,
        capped at HOT_MAX_TERMS terms and a quarter of the posting bytes.  HS_BM25_HOT=0 disables (A/B runs)."""
        import os
        self.hot_terms = self.hot_c = self.hot_of_term = None
        n = self.n_docs
        if os.environ.get("HS_BM25_HOT", "1") == "0" or n < 8192 or self.n_terms == 0 or self.avgdl <= 0:
            check(self.lib.hs_bm25_build_hot(self.handle, None, None, 0, None, None, st), "hs_bm25_build_hot")
            return
        df_local = self.indptr[1:] - self.indptr[:-1]
        budget = min(self.HOT_MAX_TERMS, int(self.postings.shape[0] // 4 // max(n, 1)))
        cand = torch.nonzero(df_local * 2 >= n).flatten()
        if budget <= 0 or cand.numel() == 0:
            check(self.lib.hs_bm25_build_hot(self.handle, None, None, 0, None, None, st), "hs_bm25_build_hot")
            return
        if cand.numel() > budget:
            cand = cand[torch.topk(df_local[cand], budget).indices]
        cand = torch.sort(cand).values
        H = int(cand.numel())
        self.hot_terms = cand.to(torch.int32).contiguous()
        idf = torch.from_numpy(self.idf_host[cand.cpu().numpy()]).to(self.device)
        self.hot_of_term = torch.full((self.n_terms,), -1, dtype=torch.int32, device=self.device)
        self.hot_of_term[cand] = torch.arange(H, dtype=torch.int32, device=self.device)
        self.hot_c = torch.empty((H, n), dtype=torch.float64, device=self.device)
        check(self.lib.hs_bm25_build_hot(self.handle, ptr(self.hot_terms), ptr(idf), H, ptr(self.hot_of_term),
                                         ptr(self.hot_c), st), "hs_bm25_build_hot")

    # ------------------------------------------------------------------ persistence (checkpoint / resume)
    def save(self, path: str):
        """Serialise the shard (dense matrix, CSR, doc stats, global df) so that ``index()`` need not be
        repeated -- the reference re-embeds and re-fits on every start (api.py:130-137, cli.py:28-33)."""
        state = {"format": "hs_b200_shard_v1", "n_docs": self.n_docs, "doc_base": self.doc_base, "dim": self.dim}
        if self.vectors is not None:
            state["vectors"] = self.vectors[:, :self.dim].cpu()
        if self.indptr is not None:
            # tensors and plain scalars only: the file loads with weights_only=True (no unpickling of objects)
            state.update(indptr=self.indptr.cpu(), postings=self.postings.cpu(), dl=self.dl.cpu(),
                         avgdl=float(self.avgdl), k1=float(self.k1), b=float(self.b),
                         df=torch.from_numpy(np.ascontiguousarray(self.df_host, dtype=np.int64)),
                         n_docs_global=int(self.n_docs_global), max_dl=int(self.max_dl))
        torch.save(state, path)

    @classmethod
    def load(cls, path: str, device) -> "DeviceIndex":
        state = torch.load(path, map_location="cpu", weights_only=True)
        if not isinstance(state, dict) or state.get("format") != "hs_b200_shard_v1":
            raise ValueError(f"{path}: not an hs_b200 shard file")
        n_docs, doc_base = int(state["n_docs"]), int(state["doc_base"])
        if "vectors" in state:
            v = state["vectors"]
            if v.dim() != 2 or v.shape[0] != n_docs or v.shape[1] != int(state["dim"]):
                raise ValueError(f"{path}: dense matrix {tuple(v.shape)} does not match n_docs={n_docs}, dim={state['dim']}")
        if "indptr" in state:
            indptr, postings, dl, df = state["indptr"], state["postings"], state["dl"], state["df"]
            if (indptr.dim() != 1 or indptr.numel() < 1 or int(indptr[0]) != 0 or postings.dim() != 2
                    or postings.shape[1] != 2 or int(indptr[-1]) != postings.shape[0] or dl.numel() != n_docs
                    or df.numel() != indptr.numel() - 1 or bool((indptr[1:] < indptr[:-1]).any())):
                raise ValueError(f"{path}: inconsistent CSR / doc statistics in shard file")
            if postings.shape[0] and int(postings[:, 0].view(torch.int32).max()) >= n_docs:
                raise ValueError(f"{path}: a posting names a doc outside the shard")
        shard = cls(device, n_docs, doc_base)
        if "vectors" in state:
            shard.set_dense(state["vectors"].to(shard.device))
        if "indptr" in state:
            shard.set_bm25(indptr, postings, dl, float(state["avgdl"]), df.numpy(), int(state["n_docs_global"]),
                           float(state["k1"]), float(state["b"]), max_dl=int(state["max_dl"]))
        return shard

    @property
    def has_dense(self) -> bool:
        return self.vectors is not None

    @property
    def has_bm25(self) -> bool:
        return self.indptr is not None


# ---------------------------------------------------------------------- host-side lexical statistics
class LexicalStats:
    """Host mirror of ``BM25.fit`` (bm25.py:45-81): vocabulary, per-doc term ids, df, dl, avgdl.

    Defines *term identity* (tokeniser + stop words) for the device CSR; the arithmetic itself runs in
    the CUDA kernels.
    """

    def __init__(self, remove_stopwords: bool = True):
        self.remove_stopwords = remove_stopwords
        self.vocab: Dict[str, int] = {}
        self.doc_count = 0
        self.doc_lengths = np.zeros(0, np.int64)
        self.avg_doc_len = 0
        self.indptr = np.zeros(1, np.int64)
        self.postings = np.zeros((0, 2), np.uint32)
        self.df = np.zeros(0, np.int64)

    def fit(self, documents: Sequence[str]):
        from .extractor import extract_tokens
        vocab: Dict[str, int] = {}
        lens: List[int] = []
        flat: List[int] = []
        for doc in documents:
            toks = extract_tokens(doc, remove_stopwords=self.remove_stopwords)
            lens.append(len(toks))
            flat.extend(vocab.setdefault(t, len(vocab)) for t in toks)
        n = len(documents)
        self.vocab = vocab
        self.doc_count = n
        self.doc_lengths = np.asarray(lens, dtype=np.int64)
        # bm25.py:71 -- python int sum / count (0 when the corpus is empty)
        self.avg_doc_len = (int(self.doc_lengths.sum()) / n) if n > 0 else 0
        v = len(vocab)
        if flat:
            terms = np.asarray(flat, dtype=np.int64)
            docs = np.repeat(np.arange(n, dtype=np.int64), self.doc_lengths)
            uk, tf = np.unique(terms * max(n, 1) + docs, return_counts=True)   # sorted by (term, doc)
            pt, pd = uk // max(n, 1), uk % max(n, 1)
        else:
            pt = pd = tf = np.zeros(0, np.int64)
        self.df = np.bincount(pt, minlength=v).astype(np.int64)
        self.indptr = np.concatenate([[0], np.cumsum(self.df)]).astype(np.int64)
        self.postings = np.stack([pd.astype(np.uint32), tf.astype(np.uint32)], axis=1) if len(pd) else \
            np.zeros((0, 2), np.uint32)
        return self

    def merge_across(self, group):
        """Doc-sharded ``fit``: this rank fitted ITS contiguous doc range; make the term ids, df,
        doc count, doc lengths and avgdl corpus-global while the CSR keeps only this rank's docs (local doc ids).

        Term ids come out exactly as a single-process fit over the whole corpus would assign them (first appearance,
        ranks hold consecutive doc ranges), so ``idf`` / ``doc_freqs`` and every score are identical to the unsharded
        index.  Host-side exchange with ``all_gather_object`` (works with gloo on CPU and with NCCL)."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        local_terms = list(self.vocab)                      # dict order == local id order
        gathered = [None] * world
        dist.all_gather_object(gathered, (local_terms, self.df, self.doc_lengths), group=group)
        gvocab: Dict[str, int] = {}
        for terms, _, _ in gathered:
            for t in terms:
                gvocab.setdefault(t, len(gvocab))
        v = len(gvocab)
        df_g = np.zeros(v, np.int64)
        for terms, df, _ in gathered:
            if len(terms):
                df_g[np.fromiter((gvocab[t] for t in terms), np.int64, len(terms))] += df
        all_dl = np.concatenate([np.asarray(dl, np.int64) for _, _, dl in gathered])
        n_global = int(all_dl.size)
        # re-key the local CSR by global term id: posting slices re-ordered, doc ids stay local
        remap = np.fromiter((gvocab[t] for t in local_terms), np.int64, len(local_terms))
        order = np.argsort(remap, kind="stable")
        lens = self.df[order]
        starts = self.indptr[:-1][order]
        total = int(lens.sum())
        if total:
            idx = np.repeat(starts - (np.cumsum(lens) - lens), lens) + np.arange(total, dtype=np.int64)
            self.postings = self.postings[idx]
        counts = np.zeros(v, np.int64)
        counts[remap] = self.df
        self.indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.local_doc_lengths = self.doc_lengths          # this rank's docs (device dl)
        self.n_local = self.doc_count
        self.vocab = gvocab
        self.df = df_g
        self.doc_lengths = all_dl                          # the reference's BM25.doc_lengths covers every doc
        self.doc_count = n_global
        self.avg_doc_len = (int(all_dl.sum()) / n_global) if n_global > 0 else 0      # bm25.py:71
        return self

    def query_term_ids(self, query: str) -> List[int]:
        """bm25.py:94,99-101 -- query tokens in order, duplicates kept, unknown terms dropped."""
        from .extractor import extract_tokens
        out = []
        for t in extract_tokens(query, remove_stopwords=self.remove_stopwords):
            tid = self.vocab.get(t)
            if tid is not None:
                out.append(tid)
        return out
