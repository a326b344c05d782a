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

import os

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
        self._pin_events = {}      # (slot, group) -> event recorded after the last H2D copy out of that staging set
        self.phase_events = None   # bench.py sets a list to collect per-phase CUDA events of the verified chain
        self.verify_wide = False   # bf16_exact: re-score 512 instead of 256 candidates per query (set after fallbacks)
        # bf16_exact: keep the screen scores as binary16 (HS_SCREEN_F32=1 restores the float32 screen, an A/B switch)
        self.screen_f16 = os.environ.get("HS_SCREEN_F32", "0") in ("", "0")
        # bf16_exact: BM25 stored as binary16 too, candidates re-scored exactly (hs_bm25_score_f16 / hs_verify_topk_cand).
        # OFF by default: measured at 10 M docs x 256 queries it saves 0.33 ms in the BM25 kernel but the binary16 x 2 select
        # is 0.73 ms slower than the binary16 + float32 one (16.09 vs 15.56 ms per step); HS_SCREEN_BM25_F16=1 turns it on.
        # Exists in the batched BM25 kernel only (HS_BM25_IMPL=tile selects the A/B kernel).
        self.screen_bm25_f16 = (os.environ.get("HS_SCREEN_BM25_F16", "0") not in ("", "0")
                                and os.environ.get("HS_BM25_IMPL", "batch") != "tile")
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
    # The H2D copies are asynchronous and queue behind whatever the stream is still running, so the host must not
    # rewrite a pinned staging set before the copies out of it have executed: every upload waits for the event of
    # the previous upload from the same set first (a no-op unless the GPU is more than one upload behind).
    def _pin_wait(self, group: str):
        ev = self._pin_events.get((self._pin_slot, group))
        if ev is not None:
            ev.synchronize()

    def _pin_mark(self, group: str):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._pin_events[(self._pin_slot, group)] = ev

    def upload_vectors(self, q: np.ndarray) -> torch.Tensor:
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.shard.dim:
            raise ValueError(f"query vectors must be [B, {self.shard.dim}], got {q.shape}")
        self._pin_wait("qv")
        pin = self._pinned("qv", q.shape, torch.float32)
        pin.copy_(torch.from_numpy(q))
        dev = self._buf("qv", q.shape, torch.float32)
        dev.copy_(pin, non_blocking=True)
        self._pin_mark("qv")
        return dev

    def upload_terms(self, term_ids: Sequence[Sequence[int]]):
        """-> (q_terms int32 [T], q_idf float64 [T], q_off int32 [B+1]) on the device."""
        (d_t, d_i, d_o, T), = self.upload_terms_split(term_ids, max(len(term_ids), 1))
        self._n_tokens = T          # host copy of q_off[B] for the kernel's workspace sizing
        return d_t, d_i, d_o

    def upload_terms_split(self, term_ids: Sequence[Sequence[int]], per: int):
        """One upload for the whole batch, sliced on the device into sub-batches of ``per`` queries:
        -> [(q_terms int32 [T_s], q_idf float64 [T_s], q_off int32 [nb_s + 1] (rebased to 0), T_s), ...]."""
        from itertools import chain
        idf = self.shard.idf_host
        df = self.shard.df_host
        nq = len(term_ids)
        # flatten with numpy (the serving loop runs this once per batch, next to a ~1.5 ms GPU step at 8 ranks)
        lens = np.fromiter(map(len, term_ids), dtype=np.int64, count=nq)
        allt = np.fromiter(chain.from_iterable(term_ids), dtype=np.int64, count=int(lens.sum()))
        # bm25.py:100 -- terms that are in no document have no idf entry and are skipped
        keep = (allt >= 0) & (allt < len(df))
        keep[keep] = df[allt[keep]] > 0
        arr = allt[keep]
        qend = np.cumsum(lens)
        kept_before = np.concatenate([[0], np.cumsum(keep)])                 # kept tokens before position i
        tok_end = kept_before[qend] if nq else np.zeros(0, np.int64)         # kept tokens up to the end of query q
        T = int(arr.size)
        offs: List[int] = []            # concatenated, each sub-batch starting again at 0
        cuts = []                       # (token start, offset start, nb) per sub-batch
        for s in range(0, max(nq, 1), per):
            e_ = min(nq, s + per)
            t0 = int(tok_end[s - 1]) if s > 0 else 0
            cuts.append((t0, len(offs), e_ - s))
            offs.append(0)
            offs.extend((tok_end[s:e_] - t0).tolist())
        self._pin_wait("qt")
        pin_t = self._pinned("qt", (max(T, 1),), torch.int32)
        pin_i = self._pinned("qi", (max(T, 1),), torch.float64)
        pin_o = self._pinned("qo", (len(offs),), torch.int32)
        if T:
            pin_t[:T].copy_(torch.from_numpy(arr.astype(np.int32)))
            pin_i[:T].copy_(torch.from_numpy(idf[arr]))
        pin_o.copy_(torch.from_numpy(np.asarray(offs, dtype=np.int32)))
        d_t = self._buf("qt", (max(T, 1),), torch.int32)
        d_i = self._buf("qi", (max(T, 1),), torch.float64)
        d_o = self._buf("qo", (len(offs),), torch.int32)
        d_t.copy_(pin_t, non_blocking=True)
        d_i.copy_(pin_i, non_blocking=True)
        d_o.copy_(pin_o, non_blocking=True)
        self._pin_mark("qt")
        out = []
        for i, (t0, o0, nb) in enumerate(cuts):
            t1 = cuts[i + 1][0] if i + 1 < len(cuts) else T
            # empty slices keep a valid (never dereferenced) pointer
            out.append((d_t[t0:max(t1, t0 + 1)] if T else d_t, d_i[t0:max(t1, t0 + 1)] if T else d_i,
                        d_o[o0:o0 + nb + 1], t1 - t0))
        return out

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
        if m in (_lib.HS_DENSE_BF16, _lib.HS_DENSE_TF32X3):          # tensor-core GEMM (K2b)
            wsp, nbytes = self._gemm_ws(B, m)
            n = self.shard.n_docs
            check(self.lib.hs_dense_gemm(self.shard.handle, ptr(q_dev), B, q_dev.stride(0), m, 0, n, wsp, nbytes,
                                         ptr(cos), n, ptr(stats), stream_ptr(self.device)), "hs_dense_gemm")
            self.launches += 2 * self._gemm_passes(B, m)
            return cos
        check(self.lib.hs_dense_scan(self.shard.handle, ptr(q_dev), B, q_dev.stride(0), m, ptr(cos), ptr(stats),
                                     stream_ptr(self.device)), "hs_dense_scan")
        self.launches += self.dense_launches(B, mode)
        return cos

    def _gemm_ws(self, B: int, m: int):
        """(256-byte aligned workspace pointer, bytes) for the tensor-core scan; builds the bf16 copy on first use."""
        if m == _lib.HS_DENSE_BF16:
            self.shard.ensure_bf16()
        nbytes = self.lib.hs_dense_gemm_workspace_bytes(self.shard.handle, B, m)
        ws = self._buf("gemm_ws", (max(nbytes // 8, 1) + 32,), torch.int64)
        return ws.data_ptr() + (-ws.data_ptr()) % 256, nbytes

    @staticmethod
    def _gemm_passes(B: int, m: int) -> int:
        """Corpus passes of one GEMM call: 256 queries per pass in bf16 (two query tiles), 128 in tf32x3."""
        per = 256 if m == _lib.HS_DENSE_BF16 else 128
        return (B + per - 1) // per

    def dense_launches(self, B: int, mode: Optional[str] = None) -> int:
        m = DENSE_MODES[mode or self.dense_mode]
        if m in (_lib.HS_DENSE_BF16, _lib.HS_DENSE_TF32X3):
            return self._gemm_passes(B, m)
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
                   n_tokens: Optional[int] = None, plus_delta: Optional[float] = None, half: bool = False) -> torch.Tensor:
        """K1.  ``half``: the scores as binary16 [B, ld] (ld = n_docs rounded up to 8) -- the BM25 screen of the verified
        mode; the max folded into ``stats`` is the float32 one either way."""
        n_tokens = self._n_tokens if n_tokens is None else int(n_tokens)
        nbytes = self.lib.hs_bm25_workspace_bytes(self.shard.n_docs, n_tokens)
        ws = self._buf("bm25_ws", (max(nbytes // 8, 1),), torch.int64)
        if half:
            ld_h = (self.shard.n_docs + 7) // 8 * 8
            sc = self._buf("bm25_h", (B, ld_h), torch.float16)
            check(self.lib.hs_bm25_score_f16(self.shard.handle, ptr(q_terms), ptr(q_idf), ptr(q_off), B, n_tokens,
                                             ptr(ws), nbytes, ptr(sc), ld_h, ptr(stats), stream_ptr(self.device)),
                  "hs_bm25_score_f16")
            self.launches += 2 if n_tokens > 0 else 1
            return sc
        sc = self._buf("bm25", (B, self.shard.n_docs), torch.float32)
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
        from .parallel import all_reduce_
        f = self._buf("stats_f", (B, 4), torch.float32)
        st = stream_ptr(self.device)
        # (-min, max, max, -min) form: one all-reduce(MAX); same arithmetic as parallel.allreduce_stats
        check(self.lib.hs_stats_to_maxform(ptr(stats), ptr(f), B, st), "hs_stats_to_maxform")
        all_reduce_(f, "max", self.group)
        check(self.lib.hs_stats_from_maxform(ptr(f), ptr(stats), B, st), "hs_stats_from_maxform")
        self.launches += 2
        return stats

    def fuse_topk(self, mode: int, a: torch.Tensor, b: Optional[torch.Tensor], stats: Optional[torch.Tensor],
                  wa: float, wb: float, k: int, below: Optional[torch.Tensor] = None, merge: bool = True,
                  out: str = "keys") -> torch.Tensor:
        """-> ranking keys int64-typed bit patterns [B, k] (merged across shards when sharded and ``merge``)."""
        B = a.shape[0]
        n = self.shard.n_docs
        ws_bytes = self.lib.hs_fuse_topk_workspace_bytes(n, B, k)
        ws = self._buf("topk_ws", (max(ws_bytes // 8, 1),), torch.int64)
        keys = self._buf(out, (B, k), torch.int64)
        check(self.lib.hs_fuse_topk(self.shard.handle, mode, ptr(a), ptr(b), ptr(stats), float(wa), float(wb),
                                    B, k, ptr(below), ptr(ws), ws_bytes, ptr(keys), stream_ptr(self.device)),
              "hs_fuse_topk")
        self.launches += self._select_launches(n, k)
        return self._merge_across(keys) if merge else keys

    @staticmethod
    def _select_launches(n: int, k: int) -> int:
        """Kernels one hs_fuse_topk call launches: select + merge, + block maxima and bound when the shard is large enough
        for a starting bound (the rule of fuse_topk_impl in csrc/topk.cu)."""
        if n <= 0:
            return 0
        bound = any(n >= nb * 128 * 2 and k <= nb // 2 for nb in (1024, 512, 256))
        return 4 if bound else 2

    def _merge_across(self, keys: torch.Tensor) -> torch.Tensor:
        """C1: all-gather of the per-shard key lists [B, k] + merge kernel; identity on a single shard."""
        if self.group is None or self.world == 1:
            return keys
        from .parallel import allgather_keys
        B, k = keys.shape
        gathered = allgather_keys(keys, self.group, self._buf("keys_all", (self.world, B, k), torch.int64))
        merged = self._buf("keys_merged", (B, k), torch.int64)
        check(self.lib.hs_topk_merge(ptr(gathered), self.world, B, k, ptr(merged),
                                     stream_ptr(self.device)), "hs_topk_merge")
        self.launches += 1
        return merged

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
            ev, hs, hi, hf, qb = item
            ev.synchronize()
            sc, ids = hs.numpy().copy(), hi.numpy().copy()
            if hf is not None:
                # bf16_exact: the verification flags travel with the result, so the loop never waits for the GPU between
                # batches; a flagged query (rare) is redone in the exact mode here, before its batch is handed over
                bad = np.nonzero(hf.numpy()[:len(qb)])[0].tolist()
                if bad:
                    s2, i2 = self._redo_exact(qb, bad, k, HS_FUSE_HYBRID_BM25, ws, wl, True, True)
                    sc[bad], ids[bad] = s2.cpu().numpy(), i2.cpu().numpy()
            return sc, ids

        try:
            for n, qb in enumerate(batches):
                if len(pending) == depth:          # the slot reused below must have been drained
                    yield collect(pending.popleft())
                self._pin_slot = n % depth
                deferred = []
                sc, ids = self._run(qb, k, HS_FUSE_HYBRID_BM25, ws, wl, True, True, dense_mode, deferred_flags=deferred)
                with torch.cuda.device(self.device):
                    hs = self._pinned("out_s", tuple(sc.shape), sc.dtype)
                    hi = self._pinned("out_i", tuple(ids.shape), ids.dtype)
                    hs.copy_(sc, non_blocking=True)
                    hi.copy_(ids, non_blocking=True)
                    hf = None
                    if deferred:
                        hf = self._pinned("out_f", tuple(deferred[0].shape), deferred[0].dtype)
                        hf.copy_(deferred[0], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                pending.append((ev, hs, hi, hf, qb))
            while pending:
                yield collect(pending.popleft())
        finally:
            self._pin_slot = 0

    def search_semantic(self, qb: QueryBatch, k: int, sw: float = 1.0, dense_mode: Optional[str] = None,
                        filtered: Optional[bool] = None):
        """Searcher.search with lexical weight 0 (pipelines.py:317-324,474-481): min-max cosine * sw.

        In the tensor-core modes on a large shard the select's pre-filter runs inside the GEMM epilogue
        (``_semantic_filtered``): no [B, n] score matrix is written or re-read.  ``filtered`` forces the choice."""
        mode = dense_mode or self.dense_mode
        if filtered is None:
            filtered = (mode in ("bf16", "tf32x3") and self.shard.n_docs >= self.FILTER_MIN_DOCS
                        and 0 < k <= HS_TOPK_MAX - 64)
        if filtered:
            out = self._semantic_filtered(qb, k, sw, mode)
            if out is not None:
                return out
        return self._run(qb, k, HS_FUSE_SEARCHER, sw, 0.0, True, False, dense_mode)

    FILTER_MIN_DOCS = 1 << 18       # below this the stored-matrix path is just as fast
    FILTER_GROUP = 512              # queries per candidate-buffer set (two bf16 corpus passes)

    def _semantic_filtered(self, qb: QueryBatch, k: int, sw: float, mode: str):
        """Pure-semantic top-k with the candidate filter fused into the GEMM epilogue (hs_dense_gemm_filter).

        Per group of queries, all on the stream:  (1) GEMM over a sample block of the shard (docs [0, S), stored);
        (2) exact top-k_sel of the sample -> its k_sel-th score is a valid lower bound ``thr`` on the k_sel-th best
        score of the whole shard (k_sel real docs reach it);  (3) GEMM over docs [S, n) whose epilogue appends the
        (query, doc) pairs with cos >= thr as ranking keys (~k_sel * n / S per query) and folds min/max;
        (4) [sharded] C2;  (5) hs_cand_select: best k_sel by cosine among candidates + sample keys, re-keyed with
        the fused score under the final stats, best k out;  (6) [sharded] C1.  Bit-identical to the stored-matrix
        path (tests/test_gpu_gemm.py); returns None when a candidate list overflowed (caller redoes it unfiltered)."""
        B, n = len(qb), self.shard.n_docs
        m = DENSE_MODES[mode]
        k_sel = 128 if k <= 100 else (512 if k <= 480 else 2048)
        S = min(max(32768, n // 64), n // 2)
        S = (S + 127) // 128 * 128
        # candidates are appended per (CTA, epilogue group) segment: capacity 8x (k_sel = 128) / 4x (larger k_sel) the
        # expected k_sel * (n - S) / S / n_seg appends of a segment, at least 32 keys
        n_seg = self.lib.hs_dense_gemm_filter_segments(self.shard.handle, m)
        expect = k_sel * ((n - S + S - 1) // S)
        cap = max(32, 1 << ((8 if k_sel == 128 else 4) * expect // n_seg).bit_length())
        st = stream_ptr(self.device)
        out_s, out_i = [], []
        with torch.cuda.device(self.device):
            qd_all = self.upload_vectors(qb.vectors)
            overflow = self._buf("cand_ovf", (1,), torch.int32)
            overflow.zero_()
            group = self.FILTER_GROUP if k_sel == 128 else 256          # bounds the candidate buffers (<= 0.5 GB)
            for s in range(0, B, group):
                e = min(B, s + group)
                nb = e - s
                qd = qd_all[s:e]
                stats = self._stats(nb)
                wsp, nbytes = self._gemm_ws(nb, m)
                cos_s = self._buf("cos_sample", (nb, S), torch.float32)
                check(self.lib.hs_dense_gemm(self.shard.handle, ptr(qd), nb, qd.stride(0), m, 0, S, wsp, nbytes,
                                             ptr(cos_s), S, ptr(stats), st), "hs_dense_gemm")
                ws_bytes = self.lib.hs_fuse_topk_workspace_bytes(S, nb, k_sel)
                ws = self._buf("topk_ws", (max(ws_bytes // 8, 1),), torch.int64)
                keys_s = self._buf("keys_sample", (nb, k_sel), torch.int64)
                check(self.lib.hs_topk_select(ptr(cos_s), S, S, self.shard.doc_base, nb, k_sel, ptr(ws), ws_bytes,
                                              ptr(keys_s), st), "hs_topk_select")
                thr = self._buf("cand_thr", (nb,), torch.float32)
                check(self.lib.hs_keys_kth_score(ptr(keys_s), nb, k_sel, k_sel, ptr(thr), st), "hs_keys_kth_score")
                cand = self._buf("cand", (nb, n_seg, cap), torch.int64)
                cnt = self._buf("cand_cnt", (nb, n_seg), torch.int32)
                cnt.zero_()
                check(self.lib.hs_dense_gemm_filter(self.shard.handle, ptr(qd), nb, qd.stride(0), m, S, n, wsp, nbytes,
                                                    ptr(thr), ptr(cand), cap, ptr(cnt), ptr(stats), st),
                      "hs_dense_gemm_filter")
                stats = self._exchange_stats(stats, nb)
                keys = self._buf("keys", (nb, k), torch.int64)
                check(self.lib.hs_cand_select(ptr(cand), ptr(cnt), n_seg, cap, ptr(keys_s), k_sel, HS_FUSE_SEARCHER, ptr(stats),
                                              float(sw), nb, k_sel, k, ptr(keys), ptr(overflow), st), "hs_cand_select")
                self.launches += 4 * self._gemm_passes(nb, m) + 7
                sc, ids = self.unpack(self._merge_across(keys))
                out_s.append(sc.clone() if B > group else sc)
                out_i.append(ids.clone() if B > group else ids)
            if self.group is not None and self.world > 1:      # every rank must take the same branch below
                from .parallel import all_reduce_
                all_reduce_(overflow, "max", self.group)
            if int(overflow.item()) != 0:          # a candidate list overflowed: results may miss docs
                return None
        if len(out_s) == 1:
            return out_s[0], out_i[0]
        return torch.cat(out_s), torch.cat(out_i)

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

    def search_bm25(self, qb: QueryBatch, k: int, plus_delta: Optional[float] = None):
        """BM25.search (bm25.py:129-142): raw float32 BM25 score, canonical tie order (BM25Plus with ``plus_delta``)."""
        return self._run(qb, k, HS_FUSE_RAW, 1.0, 0.0, False, True, None, plus_delta=plus_delta)

    # ------------------------------------------------------------------ screen on the tensor cores, verify exactly
    VERIFY_EXT_CAP = 64

    def _verified_sub_batch(self, qd, nb, stats, bm, mode, wa, wb, k, flags, eps):
        """One sub-batch of ``dense_mode="bf16_exact"``: bf16 GEMM (cos within eps) -> exact min / max from the listed
        extreme candidates -> approximate select of k_sel > k docs -> exact re-scoring + re-sort + soundness check
        (hs_verify_topk).  ``bm`` is produced by ``bm_fn(stats)`` between the scan and the statistics.  -> keys [nb, k]."""
        m = _lib.HS_DENSE_BF16
        n = self.shard.n_docs
        st = stream_ptr(self.device)
        marks = self.phase_events          # bench.py: a list -> one CUDA event per phase boundary of every sub-batch

        def mark(name=None):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                if name is None:
                    marks.append(ev)
                else:                      # finer marks inside the last phase: (name of the part that just ended, event)
                    marks.extra.append((name, ev)) if hasattr(marks, "extra") else None
        mark()
        n_seg = self.lib.hs_dense_gemm_filter_segments(self.shard.handle, m)
        cap = self.VERIFY_EXT_CAP
        wsp, nbytes = self._gemm_ws(nb, m)
        ext = self._buf("vext", (nb, n_seg, 2, cap), torch.int64)
        ext_cnt = self._buf("vext_cnt", (nb, n_seg, 2), torch.int32)
        ext_cnt.zero_()
        f16 = self.screen_f16
        if f16:      # screen scores as binary16 [nb, ld_h]: 2 bytes per (query, doc) written here and re-read by the select
            ld_h = (n + 7) // 8 * 8
            cos = self._buf("cos_h", (nb, ld_h), torch.float16)
            eps = eps + _lib.SCREEN_F16_EPS
            check(self.lib.hs_dense_gemm_ext_f16(self.shard.handle, ptr(qd), nb, qd.stride(0), m, 0, n, wsp, nbytes, ptr(cos),
                                                 ld_h, ptr(stats), ptr(ext), ptr(ext_cnt), cap, float(eps), st),
                  "hs_dense_gemm_ext_f16")
        else:
            cos = self._buf("cos", (nb, n), torch.float32)
            check(self.lib.hs_dense_gemm_ext(self.shard.handle, ptr(qd), nb, qd.stride(0), m, 0, n, wsp, nbytes, ptr(cos), n,
                                             ptr(stats), ptr(ext), ptr(ext_cnt), cap, float(eps), st), "hs_dense_gemm_ext")
        mark()
        # b array: a float32 lexical vector given by the caller, or ("bm25", q_terms, q_idf, q_off, n_tokens) -> BM25 scored
        # here; with the binary16 screen BM25 is stored as binary16 too and re-scored exactly for the candidates below
        terms = bm[1:] if isinstance(bm, tuple) else None
        b_half = terms is not None and f16 and self.screen_bm25_f16
        if terms is not None:
            b = self.bm25_score(terms[0], terms[1], terms[2], nb, stats, terms[3], half=b_half)
        else:
            b = bm
        mark()
        check(self.lib.hs_verify_stats(self.shard.handle, ptr(qd), nb, qd.stride(0), ptr(ext), ptr(ext_cnt), n_seg, cap,
                                       float(eps), ptr(stats), ptr(flags), st), "hs_verify_stats")
        mark("verify_stats")
        stats = self._exchange_stats(stats, nb)
        mark("exchange_stats")
        # candidates re-scored per query: the smaller the list the cheaper the select, but the k-th exact score must clear
        # the bound on the docs outside it (else the query falls back): 2.56 k by default, widened for good once more than
        # 5 % of a call's queries fell back (see _run)
        k_sel = (256 if k <= 100 and not self.verify_wide else 512) if k <= 256 else HS_TOPK_MAX
        if f16:
            ws_bytes = self.lib.hs_fuse_topk_workspace_bytes(n, nb, k_sel)
            ws = self._buf("topk_ws", (max(ws_bytes // 8, 1),), torch.int64)
            approx = self._buf("vkeys", (nb, k_sel), torch.int64)
            check(self.lib.hs_fuse_topk_f16(self.shard.handle, mode, ptr(cos), ptr(b), 1 if b_half else 0, cos.stride(0),
                                            ptr(stats), float(wa), float(wb), nb, k_sel, ptr(ws), ws_bytes, ptr(approx), st),
                  "hs_fuse_topk_f16")
            self.launches += self._select_launches(n, k_sel)
        else:
            approx = self.fuse_topk(mode, cos, b, stats, wa, wb, k_sel, merge=False, out="vkeys")
        mark("select")
        keys = self._buf("keys", (nb, k), torch.int64)
        if b_half:
            # exact BM25 of the candidates only: key list -> shard-local doc ids -> BM25.score (float64, token order)
            cand = self._buf("vcand", (nb, k_sel), torch.int64)
            check(self.lib.hs_keys_local_docs(ptr(approx), nb * k_sel, self.shard.doc_base, n, ptr(cand), st),
                  "hs_keys_local_docs")
            b_cand = self._buf("vcand_bm", (nb, k_sel), torch.float64)
            check(self.lib.hs_bm25_score_docs(self.shard.handle, ptr(terms[0]), ptr(terms[1]), ptr(terms[2]), nb, ptr(cand),
                                              k_sel, ptr(b_cand), st), "hs_bm25_score_docs")
            mark("rescore_bm25")
            check(self.lib.hs_verify_topk_cand(self.shard.handle, ptr(qd), nb, qd.stride(0), mode, ptr(b_cand),
                                               float(_lib.SCREEN_F16_EPS), ptr(stats), float(wa), float(wb), ptr(approx),
                                               k_sel, k, float(eps), ptr(keys), ptr(flags), st), "hs_verify_topk_cand")
            self.launches += 2
        else:
            check(self.lib.hs_verify_topk(self.shard.handle, ptr(qd), nb, qd.stride(0), mode, ptr(b), ptr(stats), float(wa),
                                          float(wb), ptr(approx), k_sel, k, float(eps), ptr(keys), ptr(flags), st),
                  "hs_verify_topk")
        self.launches += 2 * self._gemm_passes(nb, m) + 3
        mark("verify_topk")
        keys = self._merge_across(keys)
        mark()
        return keys

    def _redo_exact(self, qb, bad, k, mode, wa, wb, use_dense, use_bm25, lex=None, plus_delta=None):
        """The queries ``bad`` of ``qb`` again, in the exact mode (every rank calls this with the same list)."""
        self.verify_fallbacks = getattr(self, "verify_fallbacks", 0) + len(bad)
        if len(bad) > max(1, len(qb) // 20):
            self.verify_wide = True
        sub = QueryBatch(vectors=None if qb.vectors is None else qb.vectors[bad],
                         term_ids=None if qb.term_ids is None else [qb.term_ids[i] for i in bad])
        lx = None if lex is None else lex[bad]
        return self._run(sub, k, mode, wa, wb, use_dense, use_bm25, "exact", lex=lx, plus_delta=plus_delta)

    def _run(self, qb, k, mode, wa, wb, use_dense, use_bm25, dense_mode, lex=None, plus_delta=None, deferred_flags=None):
        """``deferred_flags`` (a list): in the verified mode, do not read the verification flags back here -- append the
        device tensor (already reduced over the shards) and leave the exact redo of flagged queries to the caller."""
        B = len(qb)
        out_s, out_i = [], []
        verify = use_dense and (dense_mode or self.dense_mode) == "bf16_exact"
        if verify and (mode == HS_FUSE_RAW or k > 1024 or self.shard.n_docs == 0):
            verify, dense_mode = False, "exact"
        if verify:
            flags = torch.zeros(max(B, 1), dtype=torch.int32, device=self.device)
            eps = _lib.VERIFY_EPS["bf16_exact"]
        with torch.cuda.device(self.device):
            # ONE upload of the whole batch before the sub-batch loop; the loop only slices device tensors (the
            # staging buffers are never rewritten while a sub-batch that reads them is still queued)
            qd_all = self.upload_vectors(qb.vectors) if use_dense else None
            terms_all = self.upload_terms_split(qb.term_ids, self.max_batch) if use_bm25 else None
            for bi, (s, e) in enumerate(self._batches(B)):
                nb = e - s
                stats = self._stats(nb)
                cos = bm = None
                if use_dense and not verify:
                    cos = self.dense_scan(qd_all[s:e], stats, dense_mode)
                if use_bm25 and not verify:
                    qt, qi, qo, n_tok = terms_all[bi]
                    bm = self.bm25_score(qt, qi, qo, nb, stats, n_tok, plus_delta=plus_delta)
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
                if verify:
                    if use_bm25:
                        qt, qi, qo, n_tok = terms_all[bi]
                        bm = ("bm25", qt, qi, qo, n_tok)
                    keys = self._verified_sub_batch(qd_all[s:e], nb, stats, bm, mode, wa, wb, k, flags[s:e], eps)
                else:
                    if mode != HS_FUSE_RAW:
                        stats = self._exchange_stats(stats, nb)
                    a, b = (cos, bm) if use_dense else (bm, None)
                    keys = self._select(mode, a, b, stats, wa, wb, k)
                sc, ids = self.unpack(keys)
                out_s.append(sc.clone() if e < B or s > 0 else sc)
                out_i.append(ids.clone() if e < B or s > 0 else ids)
        sc, ids = (out_s[0], out_i[0]) if len(out_s) == 1 else (torch.cat(out_s), torch.cat(out_i))
        if verify:
            # queries whose verification could not PROVE the result (a candidate list overflowed, or the k-th exact score
            # does not clear the bound on the docs outside the candidate list) are redone in the exact mode
            if self.group is not None and self.world > 1:
                from .parallel import all_reduce_
                all_reduce_(flags, "max", self.group)
            if deferred_flags is not None:
                deferred_flags.append(flags)
                return sc, ids
            bad = torch.nonzero(flags[:B]).flatten().cpu().tolist()
            if bad:
                sc, ids = sc.clone(), ids.clone()
                s2, i2 = self._redo_exact(qb, bad, k, mode, wa, wb, use_dense, use_bm25, lex=lex, plus_delta=plus_delta)
                idx = torch.tensor(bad, dtype=torch.int64, device=self.device)
                sc[idx], ids[idx] = s2, i2
        return sc, ids

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
    def bm25_score_docs(self, term_ids: Sequence[Sequence[int]], doc_ids: torch.Tensor,
                        plus_delta: Optional[float] = None) -> torch.Tensor:
        """BM25.score on candidate docs (pipelines.py:485; BM25Plus.score with ``plus_delta``).
        doc_ids int64 [B, C] shard-local."""
        B, Cn = doc_ids.shape
        with torch.cuda.device(self.device):
            qt, qi, qo = self.upload_terms(term_ids)
            out = self._buf("bm25_docs", (B, Cn), torch.float64)
            if plus_delta is not None:
                check(self.lib.hs_bm25plus_score_docs(self.shard.handle, ptr(qt), ptr(qi), ptr(qo), B,
                                                      ptr(doc_ids.contiguous()), Cn, float(plus_delta), ptr(out),
                                                      stream_ptr(self.device)), "hs_bm25plus_score_docs")
            else:
                check(self.lib.hs_bm25_score_docs(self.shard.handle, ptr(qt), ptr(qi), ptr(qo), B,
                                                  ptr(doc_ids.contiguous()), Cn, ptr(out), stream_ptr(self.device)),
                      "hs_bm25_score_docs")
            self.launches += 1
        return out

    def bm25_score_docs_global(self, term_ids: Sequence[Sequence[int]], doc_ids: torch.Tensor,
                               plus_delta: Optional[float] = None) -> torch.Tensor:
        """``bm25_score_docs`` for GLOBAL doc ids (int64 [B, C], < 0 = padding).  Doc-sharded: every rank scores the
        candidates it owns (the others come out 0.0) and one all-reduce(SUM) of B x C doubles assembles the result --
        exact, because exactly one rank contributes a non-zero term per entry."""
        local = doc_ids - self.shard.doc_base
        local = torch.where((local >= 0) & (local < self.shard.n_docs) & (doc_ids >= 0), local,
                            torch.full_like(local, -1))
        out = self.bm25_score_docs(term_ids, local, plus_delta=plus_delta)
        if self.group is not None and self.world > 1:
            from .parallel import all_reduce_
            all_reduce_(out, "sum", self.group)
        return out

    def gather_rows_global(self, doc_ids: torch.Tensor) -> torch.Tensor:
        """float32 [B, C, ld] rows of the dense matrix for GLOBAL doc ids (rows this rank does not hold arrive through
        one all-reduce(SUM) of zero-filled blocks: exact).  Used to hand MMR its candidate matrix when doc-sharded."""
        B, Cn = doc_ids.shape
        sh = self.shard
        local = doc_ids - sh.doc_base
        own = (local >= 0) & (local < sh.n_docs) & (doc_ids >= 0)
        rows = sh.vectors[torch.where(own, local, torch.zeros_like(local)).reshape(-1)].view(B, Cn, sh.ld)
        rows = rows * own.unsqueeze(-1).to(rows.dtype)
        if self.group is not None and self.world > 1:
            from .parallel import all_reduce_
            all_reduce_(rows, "sum", self.group)
        return rows

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

    def mmr_sharded(self, doc_ids: np.ndarray, rel: np.ndarray, lam: float, k: int, chunk: int = 256) -> np.ndarray:
        """MMR when the corpus is doc-sharded: ``doc_ids`` int64 [B, C] GLOBAL ids (-1 = padding),
        ``rel`` float64 [B, C], the same on every rank.  Per chunk of queries the candidate rows are assembled on every
        rank with one all-reduce (``gather_rows_global``), then the greedy selection -- independent per query -- runs
        data-parallel: rank r takes every world-th query of the chunk; one all-reduce returns all selections to all
        ranks.  -> int32 [B, k] positions into the candidate lists, identical to the unsharded ``mmr``."""
        import torch.distributed as dist
        rank = dist.get_rank(self.group)
        B, Cn = doc_ids.shape
        dev, sh = self.device, self.shard
        out = torch.zeros((B, k), dtype=torch.int32, device=dev)          # selection + 1; 0 = not mine / none
        with torch.cuda.device(dev):
            for s in range(0, B, chunk):
                e = min(B, s + chunk)
                ids = torch.from_numpy(np.ascontiguousarray(doc_ids[s:e])).to(dev)
                rows = self.gather_rows_global(ids)                        # [nb, C, ld] on every rank
                mine = list(range(rank, e - s, self.world))
                if not mine:
                    continue
                mt = torch.tensor(mine, dtype=torch.int64, device=dev)
                sub = rows[mt].reshape(len(mine) * Cn, sh.ld)[:, :sh.dim].contiguous()
                tmp = DeviceIndex(dev, len(mine) * Cn)
                tmp.set_dense(sub)
                cand = torch.arange(len(mine) * Cn, dtype=torch.int64, device=dev).view(len(mine), Cn)
                cand = torch.where(ids[mt] >= 0, cand, torch.full_like(cand, -1))
                r = torch.from_numpy(np.ascontiguousarray(rel[s:e][mine])).to(dev)
                sel = SearchEngine(tmp).mmr(cand, r, lam, k)
                out[s + mt] = sel + 1
                self.launches += 2
            from .parallel import all_reduce_
            all_reduce_(out, "sum", self.group)
        return (out - 1).cpu().numpy()

