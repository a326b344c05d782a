"""Batched search over one device shard (and, with a process group, over doc-sharded GPUs).

Data flow of one query batch (B queries), all on the current CUDA stream, nothing synchronises:

    stats_reset
    K2 dense_scan      -> cos  [B, n] float32   + (min, max) per query      (utils.py:28-54,67-68)
    K1 bm25_score      -> bm25 [B, n] float32   + max per query             (bm25.py:83-127, pipelines.py:332)
    [sharded] C2: one all-reduce(MAX) of (-min_cos, max_cos, max_bm, -min_lex) per query
    K3+K4 fuse_topk    -> per-shard top-k ranking keys [B, k]               (core.py:264-271, pipelines.py:331-343)
    [sharded] C1: all-gather of the key lists + merge kernel
    keys_unpack        -> (float32 score, int64 global doc id)

Every rank ends with the same merged result.  Sharded == unsharded bit for bit: min/max are exact
under any reduction order, per-doc scores depend only on the doc and the global statistics, and the
merge is a pure comparison on 64-bit keys.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (DENSE_MODES, HS_FUSE_HYBRID_BM25, HS_FUSE_RAW, HS_FUSE_SEARCHER, HS_TOPK_MAX, check,
                   ptr, stream_ptr)
from .index import DeviceIndex


@dataclass
class QueryBatch:
    """Host-side description of B queries: vectors [B, dim] float32 and/or term-id lists."""
    vectors: Optional[np.ndarray] = None                  # float32 [B, dim]
    term_ids: Optional[Sequence[Sequence[int]]] = None    # per query: known term ids, in order, dups kept

    def __len__(self):
        if self.vectors is not None:
            return len(self.vectors)
        return len(self.term_ids)


class SearchEngine:
    def __init__(self, shard: DeviceIndex, group=None, max_batch: int = 32, dense_mode: str = "exact"):
        self.shard = shard
        self.lib = shard.lib
        self.device = shard.device
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        self.max_batch = int(max_batch)
        self.dense_mode = dense_mode
        self._bufs = {}
        self._pin_slot = 0         # which set of pinned staging buffers the uploads use (search_*_stream)
        self.launches = 0          # kernels launched by this engine (bench.py: gpu_launches)

    # ------------------------------------------------------------------ buffers (never on the hot path twice)
    def _buf(self, name: str, shape, dtype) -> torch.Tensor:
        need = int(np.prod(shape)) if len(shape) else 1
        t = self._bufs.get(name)
        if t is None or t.numel() < need or t.dtype != dtype:
            t = torch.empty(max(need, 1), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t[:need].view(*shape)

    def _pinned(self, name: str, shape, dtype) -> torch.Tensor:
        need = int(np.prod(shape)) if len(shape) else 1
        key = f"pin{self._pin_slot}_{name}"
        t = self._bufs.get(key)
        if t is None or t.numel() < need or t.dtype != dtype:
            # page-locked allocations are slow and synchronise the device: round the capacity up generously so
            # that a batch with a few more tokens than the last one never triggers one in a serving loop
            cap = max(1024, 1 << (max(need, 1) - 1).bit_length())
            t = torch.empty(cap, dtype=dtype, pin_memory=True)
            self._bufs[key] = t
        return t[:need].view(*shape)

    # ------------------------------------------------------------------ query upload
    def upload_vectors(self, q: np.ndarray) -> torch.Tensor:
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.shard.dim:
            raise ValueError(f"query vectors must be [B, {self.shard.dim}], got {q.shape}")
        pin = self._pinned("qv", q.shape, torch.float32)
        pin.copy_(torch.from_numpy(q))
        dev = self._buf("qv", q.shape, torch.float32)
        dev.copy_(pin, non_blocking=True)
        return dev

    def upload_terms(self, term_ids: Sequence[Sequence[int]]):
        """-> (q_terms int32 [T], q_idf float64 [T], q_off int32 [B+1]) on the device."""
        idf = self.shard.idf_host
        df = self.shard.df_host
        off = [0]
        flat: List[int] = []
        for ids in term_ids:
            # bm25.py:100 -- terms that are in no document have no idf entry and are skipped
            flat.extend(int(t) for t in ids if 0 <= int(t) < len(df) and df[int(t)] > 0)
            off.append(len(flat))
        T = len(flat)
        pin_t = self._pinned("qt", (max(T, 1),), torch.int32)
        pin_i = self._pinned("qi", (max(T, 1),), torch.float64)
        pin_o = self._pinned("qo", (len(off),), torch.int32)
        if T:
            arr = np.asarray(flat, dtype=np.int64)
            pin_t[:T].copy_(torch.from_numpy(arr.astype(np.int32)))
            pin_i[:T].copy_(torch.from_numpy(idf[arr]))
        pin_o.copy_(torch.tensor(off, dtype=torch.int32))
        d_t = self._buf("qt", (max(T, 1),), torch.int32)
        d_i = self._buf("qi", (max(T, 1),), torch.float64)
        d_o = self._buf("qo", (len(off),), torch.int32)
        d_t.copy_(pin_t, non_blocking=True)
        d_i.copy_(pin_i, non_blocking=True)
        d_o.copy_(pin_o, non_blocking=True)
        self._n_tokens = T          # host copy of q_off[B] for the kernel's workspace sizing
        return d_t, d_i, d_o

    # ------------------------------------------------------------------ kernels
    def _stats(self, B: int) -> torch.Tensor:
        s = self._buf("stats", (B, 4), torch.int32)
        check(self.lib.hs_stats_reset(ptr(s), B, stream_ptr(self.device)), "hs_stats_reset")
        self.launches += 1
        return s

    def dense_scan(self, q_dev: torch.Tensor, stats: torch.Tensor, mode: Optional[str] = None) -> torch.Tensor:
        B = q_dev.shape[0]
        cos = self._buf("cos", (B, self.shard.n_docs), torch.float32)
        m = DENSE_MODES[mode or self.dense_mode]
        if m == _lib.HS_DENSE_BF16:
            self.shard.ensure_bf16()
            nbytes = self.lib.hs_dense_scan_bf16_workspace_bytes(self.shard.handle, B)
            ws = self._buf("gemm_ws", (max(nbytes // 8, 1) + 32,), torch.int64)
            off = (-ws.data_ptr()) % 256                           # 256-byte aligned view
            check(self.lib.hs_dense_scan_bf16(self.shard.handle, ptr(q_dev), B, q_dev.stride(0),
                                              ws.data_ptr() + off, nbytes, ptr(cos), ptr(stats),
                                              stream_ptr(self.device)), "hs_dense_scan_bf16")
            self.launches += 2 * ((B + 127) // 128)
            return cos
        check(self.lib.hs_dense_scan(self.shard.handle, ptr(q_dev), B, q_dev.stride(0), m, ptr(cos), ptr(stats),
                                     stream_ptr(self.device)), "hs_dense_scan")
        self.launches += self.dense_launches(B, mode)
        return cos

    def dense_launches(self, B: int, mode: Optional[str] = None) -> int:
        nchunk = (self.shard.dim + 127) // 128
        nchunk = nchunk if nchunk <= 4 else (6 if nchunk <= 6 else 8)
        budget, cap = 12, 4
        bq = 1
        while bq * 2 <= budget // nchunk and bq * 2 <= cap:
            bq *= 2
        n, b = 0, B
        while b > 0:               # mirrors hs_dense_scan: BQ queries per warp x QG warps per stage
            step = bq
            while step > b:
                step //= 2
            qg = 4
            while qg > 1 and step * qg > b:
                qg //= 2
            b -= step * qg
            n += 1
        return n

    def bm25_score(self, q_terms, q_idf, q_off, B: int, stats: Optional[torch.Tensor],
                   n_tokens: Optional[int] = None, plus_delta: Optional[float] = None) -> torch.Tensor:
        n_tokens = self._n_tokens if n_tokens is None else int(n_tokens)
        sc = self._buf("bm25", (B, self.shard.n_docs), torch.float32)
        nbytes = self.lib.hs_bm25_workspace_bytes(self.shard.n_docs, n_tokens)
        ws = self._buf("bm25_ws", (max(nbytes // 8, 1),), torch.int64)
        if plus_delta is not None:
            check(self.lib.hs_bm25plus_score(self.shard.handle, ptr(q_terms), ptr(q_idf), ptr(q_off), B, n_tokens,
                                             float(plus_delta), ptr(ws), nbytes, ptr(sc), ptr(stats),
                                             stream_ptr(self.device)), "hs_bm25plus_score")
        else:
            check(self.lib.hs_bm25_score(self.shard.handle, ptr(q_terms), ptr(q_idf), ptr(q_off), B, n_tokens,
                                         ptr(ws), nbytes, ptr(sc), ptr(stats), stream_ptr(self.device)),
                  "hs_bm25_score")
        self.launches += 2 if n_tokens > 0 else 1
        return sc

    def _exchange_stats(self, stats: torch.Tensor, B: int) -> torch.Tensor:
        """C2: global (min, max) per query across shards -- one all-reduce(MAX) of B x 4 floats."""
        if self.group is None or self.world == 1:
            return stats
        import torch.distributed as dist
        f = self._buf("stats_f", (B, 4), torch.float32)
        st = stream_ptr(self.device)
        # (-min, max, max, -min) form: one all-reduce(MAX); same arithmetic as parallel.allreduce_stats
        check(self.lib.hs_stats_to_maxform(ptr(stats), ptr(f), B, st), "hs_stats_to_maxform")
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.group)
        check(self.lib.hs_stats_from_maxform(ptr(f), ptr(stats), B, st), "hs_stats_from_maxform")
        self.launches += 2
        return stats

    def fuse_topk(self, mode: int, a: torch.Tensor, b: Optional[torch.Tensor], stats: Optional[torch.Tensor],
                  wa: float, wb: float, k: int, below: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> ranking keys int64-typed bit patterns [B, k] (merged across shards when sharded)."""
        B = a.shape[0]
        n = self.shard.n_docs
        ws_bytes = self.lib.hs_fuse_topk_workspace_bytes(n, B, k)
        ws = self._buf("topk_ws", (max(ws_bytes // 8, 1),), torch.int64)
        keys = self._buf("keys", (B, k), torch.int64)
        check(self.lib.hs_fuse_topk(self.shard.handle, mode, ptr(a), ptr(b), ptr(stats), float(wa), float(wb),
                                    B, k, ptr(below), ptr(ws), ws_bytes, ptr(keys), stream_ptr(self.device)),
              "hs_fuse_topk")
        self.launches += 2 if n > 0 else 0
        if self.group is not None and self.world > 1:
            from .parallel import allgather_keys
            gathered = allgather_keys(keys, self.group, self._buf("keys_all", (self.world, B, k), torch.int64))  # C1
            merged = self._buf("keys_merged", (B, k), torch.int64)
            check(self.lib.hs_topk_merge(ptr(gathered), self.world, B, k, ptr(merged),
                                         stream_ptr(self.device)), "hs_topk_merge")
            self.launches += 1
            keys = merged
        return keys

    def unpack(self, keys: torch.Tensor):
        B, k = keys.shape
        sc = self._buf("out_sc", (B, k), torch.float32)
        ids = self._buf("out_id", (B, k), torch.int64)
        check(self.lib.hs_keys_unpack(ptr(keys), B * k, ptr(sc), ptr(ids), stream_ptr(self.device)),
              "hs_keys_unpack")
        self.launches += 1
        return sc, ids

    # ------------------------------------------------------------------ paged select for k > HS_TOPK_MAX
    def _select(self, mode, a, b, stats, wa, wb, k):
        k_total = int(k)
        if k_total <= HS_TOPK_MAX:
            return self.fuse_topk(mode, a, b, stats, wa, wb, k_total)
        B = a.shape[0]
        pages, below, got = [], None, 0
        while got < k_total:
            kk = min(HS_TOPK_MAX, k_total - got)
            keys = self.fuse_topk(mode, a, b, stats, wa, wb, kk, below).clone()
            pages.append(keys)
            got += kk
            if got < k_total:
                # next page: strictly below the last key of this page (0 = exhausted -> nothing passes)
                below = keys[:, -1].contiguous().clone()
        return torch.cat(pages, dim=1)

    # ------------------------------------------------------------------ whole searches (device tensors out)
    def _batches(self, B: int):
        for s in range(0, B, self.max_batch):
            yield s, min(B, s + self.max_batch)

    def search_hybrid_bm25(self, qb: QueryBatch, k: int, ws: float, wl: float, dense_mode: Optional[str] = None):
        """HybridBM25Pipeline.search (pipelines.py:315-357) for a batch.  -> (scores [B,k], ids [B,k])."""
        return self._run(qb, k, HS_FUSE_HYBRID_BM25, ws, wl, True, True, dense_mode)

    def search_hybrid_bm25_stream(self, batches, k: int, ws: float, wl: float, dense_mode: Optional[str] = None,
                                  depth: int = 2):
        """Serving loop over an iterable of QueryBatch: yields (scores, ids) as numpy arrays, in order.

        Nothing here waits for the GPU except the hand-over of a finished result: while batch i runs, the host
        flattens and uploads batch i+1 from a second set of pinned staging buffers, and batch i's top-k is
        copied back asynchronously into pinned memory.  Same kernels and bits as ``search_hybrid_bm25``."""
        from collections import deque
        pending = deque()

        def collect(item):
            ev, hs, hi = item
            ev.synchronize()
            return hs.numpy().copy(), hi.numpy().copy()

        try:
            for n, qb in enumerate(batches):
                if len(pending) == depth:          # the slot reused below must have been drained
                    yield collect(pending.popleft())
                self._pin_slot = n % depth
                sc, ids = self.search_hybrid_bm25(qb, k, ws, wl, dense_mode)
                with torch.cuda.device(self.device):
                    hs = self._pinned("out_s", tuple(sc.shape), sc.dtype)
                    hi = self._pinned("out_i", tuple(ids.shape), ids.dtype)
                    hs.copy_(sc, non_blocking=True)
                    hi.copy_(ids, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                pending.append((ev, hs, hi))
            while pending:
                yield collect(pending.popleft())
        finally:
            self._pin_slot = 0

    def search_semantic(self, qb: QueryBatch, k: int, sw: float = 1.0, dense_mode: Optional[str] = None):
        """Searcher.search with lexical weight 0 (pipelines.py:317-324,474-481): min-max cosine * sw."""
        return self._run(qb, k, HS_FUSE_SEARCHER, sw, 0.0, True, False, dense_mode)

    def search_searcher(self, qb: QueryBatch, lex: np.ndarray, k: int, sw: float, lw: float,
                        dense_mode: Optional[str] = None):
        """Searcher.search with a lexical vector (core.py:261-271): lex float32 [B, n_docs], host array or
        device tensor."""
        return self._run(qb, k, HS_FUSE_SEARCHER, sw, lw, True, False, dense_mode, lex=lex)

    def search_faiss_style(self, qb: QueryBatch, lex, k: int, sw: float, lw: float, dense_mode: Optional[str] = None):
        """Searcher.search with use_faiss=True (core.py:244-250,264-271): only the best min(2k, N) cosine scores
        survive, every other doc gets semantic score 0.0 BEFORE the min-max normalisation.  Single shard."""
        if self.world > 1:
            raise NotImplementedError("faiss-style retrieval is not sharded")
        B, n = len(qb), self.shard.n_docs
        k2 = min(2 * int(k), n)
        with torch.cuda.device(self.device):
            stats0 = self._stats(B)
            cos = self.dense_scan(self.upload_vectors(qb.vectors), stats0, dense_mode)
            keys = self._select(HS_FUSE_RAW, cos, None, None, 1.0, 0.0, k2).clone()
            sparse = self._buf("sparse_sem", (B, n), torch.float32)
            check(self.lib.hs_scatter_keys(ptr(keys), B, k2, n, self.shard.doc_base, ptr(sparse),
                                           stream_ptr(self.device)), "hs_scatter_keys")
            stats = self._stats(B)
            check(self.lib.hs_stats_fold_minmax(ptr(sparse), n, B, 0, 1, ptr(stats), stream_ptr(self.device)),
                  "hs_stats_fold_minmax")
            self.launches += 3
            bm = None
            if lex is not None:
                bm = lex.to(self.device, torch.float32).contiguous() if isinstance(lex, torch.Tensor) else \
                    torch.from_numpy(np.ascontiguousarray(lex, dtype=np.float32)).to(self.device)
                check(self.lib.hs_stats_fold_minmax(ptr(bm), n, B, 3, 2, ptr(stats), stream_ptr(self.device)),
                      "hs_stats_fold_minmax")
                self.launches += 1
            out = self._select(HS_FUSE_SEARCHER, sparse, bm, stats, sw, lw if bm is not None else 0.0, k)
            return self.unpack(out)

    def search_bm25(self, qb: QueryBatch, k: int):
        """BM25.search (bm25.py:129-142): raw float32 BM25 score, canonical tie order."""
        return self._run(qb, k, HS_FUSE_RAW, 1.0, 0.0, False, True, None)

    def _run(self, qb, k, mode, wa, wb, use_dense, use_bm25, dense_mode, lex=None):
        B = len(qb)
        out_s, out_i = [], []
        with torch.cuda.device(self.device):
            for s, e in self._batches(B):
                nb = e - s
                stats = self._stats(nb)
                cos = bm = None
                if use_dense:
                    qd = self.upload_vectors(qb.vectors[s:e])
                    cos = self.dense_scan(qd, stats, dense_mode)
                if use_bm25:
                    qt, qi, qo = self.upload_terms(qb.term_ids[s:e])
                    bm = self.bm25_score(qt, qi, qo, nb, stats)
                if lex is not None:
                    if isinstance(lex, torch.Tensor):
                        bm = lex[s:e].to(self.device, torch.float32).contiguous()
                    else:
                        lx = np.ascontiguousarray(lex[s:e], dtype=np.float32)
                        bm = self._buf("lex", lx.shape, torch.float32)
                        bm.copy_(torch.from_numpy(lx), non_blocking=False)
                    check(self.lib.hs_stats_fold_minmax(ptr(bm), self.shard.n_docs, nb, 3, 2, ptr(stats),
                                                        stream_ptr(self.device)), "hs_stats_fold_minmax")
                    self.launches += 1
                if mode != HS_FUSE_RAW:
                    stats = self._exchange_stats(stats, nb)
                a, b = (cos, bm) if use_dense else (bm, None)
                keys = self._select(mode, a, b, stats, wa, wb, k)
                sc, ids = self.unpack(keys)
                out_s.append(sc.clone() if e < B or s > 0 else sc)
                out_i.append(ids.clone() if e < B or s > 0 else ids)
        if len(out_s) == 1:
            return out_s[0], out_i[0]
        return torch.cat(out_s), torch.cat(out_i)

    # ------------------------------------------------------------------ one hybrid step on device tensors
    def hybrid_step_device(self, qd, qt, qi, qo, B: int, n_tokens: int, k: int, ws: float, wl: float,
                           dense_mode: Optional[str] = None):
        """The kernel chain of one hybrid_bm25 batch on device-resident inputs (no host work besides the
        launches): stats reset, K2, K1, [C2], K3+K4, [C1 + merge], unpack.  Capturable in a CUDA graph."""
        stats = self._stats(B)
        cos = self.dense_scan(qd, stats, dense_mode)
        bm = self.bm25_score(qt, qi, qo, B, stats, n_tokens)
        stats = self._exchange_stats(stats, B)
        keys = self.fuse_topk(HS_FUSE_HYBRID_BM25, cos, bm, stats, ws, wl, k)
        return self.unpack(keys)

    def flatten_terms(self, term_ids: Sequence[Sequence[int]]):
        """Host side of upload_terms: (flat known term ids, their idf, offsets)."""
        idf, df = self.shard.idf_host, self.shard.df_host
        off, flat = [0], []
        for ids in term_ids:
            flat.extend(int(t) for t in ids if 0 <= int(t) < len(df) and df[int(t)] > 0)
            off.append(len(flat))
        arr = np.asarray(flat, dtype=np.int64)
        return arr, (idf[arr] if len(arr) else np.zeros(0, np.float64)), np.asarray(off, dtype=np.int32)

    # ------------------------------------------------------------------ small kernels
    def bm25_score_docs(self, term_ids: Sequence[Sequence[int]], doc_ids: torch.Tensor) -> torch.Tensor:
        """BM25.score on candidate docs (pipelines.py:485).  doc_ids int64 [B, C] shard-local."""
        B, Cn = doc_ids.shape
        with torch.cuda.device(self.device):
            qt, qi, qo = self.upload_terms(term_ids)
            out = self._buf("bm25_docs", (B, Cn), torch.float64)
            check(self.lib.hs_bm25_score_docs(self.shard.handle, ptr(qt), ptr(qi), ptr(qo), B,
                                              ptr(doc_ids.contiguous()), Cn, ptr(out), stream_ptr(self.device)),
                  "hs_bm25_score_docs")
            self.launches += 1
        return out

    def mmr(self, cand_ids: torch.Tensor, rel: torch.Tensor, lam: float, k: int) -> torch.Tensor:
        """DiversityPipeline._mmr (pipelines.py:531-569).  cand_ids int64 [B, C], rel float64 [B, C]."""
        B, Cn = cand_ids.shape
        with torch.cuda.device(self.device):
            nbytes = self.lib.hs_mmr_workspace_bytes(B, Cn)
            ws = self._buf("mmr_ws", (max((nbytes + 7) // 8, 1),), torch.int64)
            out = self._buf("mmr_out", (B, k), torch.int32)
            check(self.lib.hs_mmr(self.shard.handle, ptr(cand_ids.contiguous()), ptr(rel.contiguous()), float(lam),
                                  B, Cn, k, ptr(ws), nbytes, ptr(out), stream_ptr(self.device)), "hs_mmr")
            self.launches += 1
        return out
