"""Reader / writer for the reference's only persisted vector format: a FAISS ``IndexFlatIP`` /
``IndexFlatL2`` file (``indexer.py:59-66,280-283`` -> ``faiss.write_index``; the repository ships
``index.faiss`` = 12 L2-normalised MiniLM rows).  faiss itself is not needed: the flat index file is

    fourcc 'IxFI' (inner product) or 'IxF2' (L2)      4 bytes
    d            int32        ntotal      int64
    dummy        int64 x 2    (historically 1 << 20)
    is_trained   uint8        metric_type int32  (0 = inner product, 1 = L2)
    vector<float>: count uint64 (= ntotal * d), then count float32 values, row-major

little endian throughout.  ``FAISSIndex.add`` L2-normalises rows before adding (``indexer.py:41-46``), which
``write_index_flat(..., normalize=True)`` reproduces.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

_FOURCC = {b"IxFI": 0, b"IxF2": 1}


def read_index_flat(path: str) -> Tuple[np.ndarray, int]:
    """-> (float32 [ntotal, d] rows, metric_type)."""
    with open(path, "rb") as f:
        raw = f.read()
    four = raw[:4]
    if four not in _FOURCC:
        raise ValueError(f"{path}: not a flat FAISS index (fourcc {four!r})")
    d, = struct.unpack_from("<i", raw, 4)
    ntotal, = struct.unpack_from("<q", raw, 8)
    metric, = struct.unpack_from("<i", raw, 33)
    off = 37
    if metric > 1:                       # metric_arg present for the exotic metrics
        off += 4
    count, = struct.unpack_from("<Q", raw, off)
    off += 8
    if count != ntotal * d or len(raw) < off + 4 * count:
        raise ValueError(f"{path}: corrupt flat index (count {count}, ntotal {ntotal}, d {d})")
    vec = np.frombuffer(raw, dtype="<f4", count=count, offset=off).reshape(ntotal, d).copy()
    return vec, int(metric)


def write_index_flat(path: str, vectors: np.ndarray, metric: int = 0, normalize: bool = False) -> None:
    v = np.array(vectors, dtype="<f4", order="C")
    if v.ndim != 2:
        raise ValueError("vectors must be [n, d]")
    if normalize:                        # faiss.normalize_L2 (indexer.py:44)
        norms = np.sqrt((v.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
        v = np.where(norms > 0, v / np.maximum(norms, 1e-30), v).astype("<f4")
    n, d = v.shape
    four = b"IxFI" if metric == 0 else b"IxF2"
    head = four + struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) + \
        struct.pack("<B", 1) + struct.pack("<i", metric)
    with open(path, "wb") as f:
        f.write(head)
        f.write(struct.pack("<Q", n * d))
        f.write(v.tobytes())
