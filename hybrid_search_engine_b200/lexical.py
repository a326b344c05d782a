"""Fuzzy lexical scores on the device (mirror of ``Searcher._lexical_scores``, core.py:178-197).

Host side: lower-case the stored contents once, keep their code points and their sorted unique token
ids (tokeniser WITHOUT stop-word removal, core.py:180,188) in HBM; per query upload the lower-cased
query's code points and token-id set.  The arithmetic (bit-parallel LCS over every window, token
overlap, 0.7 / 0.3 blend, float32 rounding) runs in ``csrc/lexical.cu``.

``partial_ratio`` itself is third-party in the reference (rapidfuzz) and PARITY UNPINNED: the kernel is
bit-identical to the shared restatement in ``oracle/hybrid_oracle.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .extractor import extract_tokens


class LexicalScorer:
    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self._key = None
        self.vocab: Dict[str, int] = {}
        self.n = 0

    def prepare(self, docs: Sequence[str]):
        # identity of the caller's document list, held strongly (an id() of a freed list can be recycled)
        if self._key is not None and self._key[0] is docs and self._key[1] == len(docs):
            return
        key = (docs, len(docs))
        lowered = [d.lower() for d in docs]
        # code points (utf-32) of every lower-cased document, concatenated
        lens = np.fromiter((len(d) for d in lowered), dtype=np.int64, count=len(lowered))
        blob = "".join(lowered).encode("utf-32-le")
        chars = np.frombuffer(blob, dtype="<u4") if blob else np.zeros(0, np.uint32)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        # sorted unique token ids per document (core.py:188: set(extract_tokens(doc.lower())))
        vocab: Dict[str, int] = {}
        tok_lists: List[np.ndarray] = []
        for d in lowered:
            ids = {vocab.setdefault(t, len(vocab)) for t in extract_tokens(d)}
            tok_lists.append(np.array(sorted(ids), dtype=np.int32))
        tlen = np.fromiter((len(t) for t in tok_lists), dtype=np.int64, count=len(tok_lists))
        toks = np.concatenate(tok_lists) if tok_lists and tlen.sum() else np.zeros(0, np.int32)
        toff = np.concatenate([[0], np.cumsum(tlen)]).astype(np.int64)
        dev = self.device
        self.d_chars = torch.from_numpy(chars.astype(np.uint32).view(np.int32).copy()).to(dev) if len(chars) else \
            torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_off = torch.from_numpy(off).to(dev)
        self.d_tok = torch.from_numpy(toks).to(dev) if len(toks) else torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_toff = torch.from_numpy(toff).to(dev)
        self.vocab = vocab
        self.n = len(docs)
        self._key = key

    def scores_device_many(self, queries: Sequence[str], docs: Sequence[str]) -> torch.Tensor:
        """float32 [B, N] on the device: one upload of every query's code points / token ids, one kernel launch per
        query (the pattern is a kernel parameter), one error read-back for the batch."""
        self.prepare(docs)
        dev = self.device
        B = len(queries)
        qcs, qts, meta = [], [], []
        for query in queries:
            ql = query.lower()
            if len(ql) > 512:
                raise NotImplementedError("lexical scoring supports queries of at most 512 characters")
            qc = np.frombuffer(ql.encode("utf-32-le"), dtype="<u4") if ql else np.zeros(0, np.uint32)
            q_set = set(extract_tokens(ql))                                 # core.py:180
            known = sorted(self.vocab[t] for t in q_set if t in self.vocab)
            meta.append((len(qc), len(known), len(q_set)))
            qcs.append(qc.astype(np.uint32))
            qts.append(np.asarray(known, np.int32))
        c_off = np.concatenate([[0], np.cumsum([m[0] for m in meta])]).astype(np.int64)
        t_off = np.concatenate([[0], np.cumsum([m[1] for m in meta])]).astype(np.int64)
        all_c = np.concatenate(qcs + [np.zeros(1, np.uint32)]).view(np.int32)
        all_t = np.concatenate(qts + [np.zeros(1, np.int32)])
        d_qc = torch.from_numpy(all_c.copy()).to(dev)
        d_qt = torch.from_numpy(all_t.copy()).to(dev)
        out = torch.empty((B, max(self.n, 1)), dtype=torch.float32, device=dev)
        err = torch.zeros(max(B, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            for b, (n_c, n_t, n_set) in enumerate(meta):
                check(self.lib.hs_lexical_scores(ptr(self.d_chars), ptr(self.d_off), self.n,
                                                 d_qc.data_ptr() + 4 * int(c_off[b]), n_c, ptr(self.d_tok),
                                                 ptr(self.d_toff), d_qt.data_ptr() + 4 * int(t_off[b]), n_t, n_set,
                                                 out.data_ptr() + 4 * b * out.shape[1], err.data_ptr() + 4 * b, st),
                      "hs_lexical_scores")
        if B and int(err.max().item()):
            raise NotImplementedError("lexical scoring supports at most 64 distinct non-ASCII code points per pattern")
        return out[:, :self.n]

    def scores_device(self, query: str, docs: Sequence[str]) -> torch.Tensor:
        """float32 [N] on the device."""
        return self.scores_device_many([query], docs)[0]

    def scores(self, query: str, docs: Sequence[str]) -> np.ndarray:
        return self.scores_device(query, docs).cpu().numpy()
