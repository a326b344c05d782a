"""Mirror of the reference's numeric helpers (utils.py:5-87) on the device.

These are the reference's *stated* primitives; inside the pipelines the same arithmetic is fused into the
scan / select kernels and never materialised.  Every call uploads its arguments, so they are meant for
conformance checks and small inputs, not for the hot path.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib
from ._lib import HS_FUSE_RAW, check, ptr, stream_ptr
from .engine import SearchEngine
from .index import DeviceIndex


def _engine(vectors: np.ndarray, device=None) -> SearchEngine:
    if not torch.cuda.is_available():
        raise _lib.HsError("no CUDA device: the hybrid scoring path has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    shard = DeviceIndex(dev, v.shape[0])
    if v.shape[0]:
        shard.set_dense(torch.from_numpy(v).to(dev))
    return SearchEngine(shard)


def batch_cosine_sim(query: np.ndarray, vectors: np.ndarray, *, mode: str = "exact", device=None) -> np.ndarray:
    """utils.py:28-54 -- float32 [N]; zero query -> zeros, zero row -> 0.0."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.shape[0] == 0:
        return np.zeros(0, np.float32)
    eng = _engine(vectors, device)
    with torch.cuda.device(eng.device):
        stats = eng._stats(1)
        q = eng.upload_vectors(np.asarray(query, np.float32)[None, :])
        return eng.dense_scan(q, stats, mode).cpu().numpy()[0]


def cosine_sim(a: np.ndarray, b: np.ndarray, *, device=None) -> float:
    """utils.py:5-25 -- 0.0 if either norm is zero."""
    return float(batch_cosine_sim(a, np.asarray(b, np.float32)[None, :], device=device)[0])


def normalize_scores(scores: np.ndarray) -> np.ndarray:
    """utils.py:57-71 -- min-max to [0, 1]; a constant vector gives ones; raises on empty input.

    Elementwise host arithmetic in the array's dtype: inside the pipelines this is fused into
    ``hs_fuse_topk`` and has no kernel of its own.
    """
    scores = np.asarray(scores)
    mn, mx = scores.min(), scores.max()
    if mx - mn == 0:
        return np.ones_like(scores)
    return (scores - mn) / (mx - mn)


def top_k_indices(scores: np.ndarray, k: int, *, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:74-87 -- (scores[idx], idx) of the k best, canonical order (score desc, index asc)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    k = min(int(k), len(scores))
    if k <= 0:
        return scores[:0], np.zeros(0, np.int64)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    shard = DeviceIndex(dev, len(scores))
    eng = SearchEngine(shard)
    with torch.cuda.device(dev):
        keys = eng._select(HS_FUSE_RAW, torch.from_numpy(scores).to(dev)[None, :], None, None, 1.0, 0.0, k)
        _, ids = eng.unpack(keys)
        idx = ids.cpu().numpy()[0]
    return scores[idx], idx
