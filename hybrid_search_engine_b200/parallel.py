"""Doc-sharded execution over the GPUs of one box: host-side plumbing.

The corpus is split into contiguous doc-id ranges, one per rank (keeps the ``doc_id asc`` tie-break
mergeable).  Per query batch the path has exactly two exchange steps, both tiny and latency bound:

* C2  ``allreduce_stats``  global (min cos, max cos, max bm25, min lex) -- needed before fusing because
      min-max / max normalisation is over ALL docs (utils.py:67-71, pipelines.py:331-332).  One
      all-reduce(MAX) of B x 4 floats (minima are negated, which is exact).
* C1  ``allgather_keys``   per-shard top-k ranking keys (uint64 bit patterns carried as int64), then the
      merge kernel ``hs_topk_merge`` on every rank.

Everything here works on CPU tensors with the gloo backend too, which is how the N > 1 host logic is
tested without GPUs (tests/test_sharded_gloo.py).  The reference has no distributed code at all.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def shard_bounds(n_docs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range ``[lo, hi)`` of rank ``rank``; every rank gets ceil(n / world) docs but the last."""
    per = (n_docs + world - 1) // world
    return min(n_docs, rank * per), min(n_docs, (rank + 1) * per)


def _stage_on_host(t: torch.Tensor, group) -> bool:
    """gloo moves CUDA tensors only for some collectives: with a gloo group device tensors are staged through the host
    (the N > 1 logic can then run with several ranks on ONE GPU, which is how the 1-GPU test box covers it)."""
    import torch.distributed as dist
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce_(t: torch.Tensor, op: str = "sum", group=None) -> torch.Tensor:
    """In-place all-reduce ("sum" | "max") over ``group``; NCCL on device tensors, gloo through the host."""
    import torch.distributed as dist
    rop = dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM
    if _stage_on_host(t, group):
        h = t.cpu()
        dist.all_reduce(h, op=rop, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=rop, group=group)
    return t


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In place: columns 0 and 3 become the global minimum, 1 and 2 the global maximum.

    ``stats`` float32 [B, 4] = (min_a, max_a, max_b, min_b) of the local shard.  NaN marks "this shard
    saw nothing" (empty shard) and must lose against any real value, so it is mapped to -inf first.
    """
    import torch.distributed as dist
    stats[:, 0].neg_()
    stats[:, 3].neg_()
    torch.nan_to_num_(stats, nan=float("-inf"), posinf=float("inf"), neginf=float("-inf"))
    dist.all_reduce(stats, op=dist.ReduceOp.MAX, group=group)
    stats[:, 0].neg_()
    stats[:, 3].neg_()
    return stats


def allgather_keys(keys: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, k] int64 per rank -> [world, B, k] on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
    if _stage_on_host(keys, group):
        h = torch.empty(out.shape, dtype=keys.dtype)
        dist.all_gather_into_tensor(h.view(-1), keys.cpu().contiguous().view(-1), group=group)
        out.copy_(h)
    else:
        dist.all_gather_into_tensor(out.view(-1), keys.contiguous().view(-1), group=group)
    return out


# ---------------------------------------------------------------------- key format (host mirror)
def pack_keys(scores: np.ndarray, doc_ids: np.ndarray) -> np.ndarray:
    """numpy twin of ``hs_make_key``: uint64 ``ordered(score) << 32 | (0xFFFFFFFF - doc_id)``."""
    s = np.asarray(scores, dtype=np.float32).copy()
    s[s == 0.0] = 0.0                                   # fold -0.0 into +0.0
    u = s.view(np.uint32).astype(np.uint64)
    enc = np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    return (enc << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(doc_ids, dtype=np.uint64))


def unpack_keys(keys: np.ndarray):
    """-> (float32 scores, int64 doc ids; -1 where key == 0)."""
    k = np.asarray(keys).astype(np.uint64)
    e = (k >> np.uint64(32)).astype(np.uint32)
    u = np.where(e & np.uint32(0x80000000), e & np.uint32(0x7FFFFFFF), ~e)
    ids = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    ids[k == 0] = -1
    sc = u.astype(np.uint32).view(np.float32).copy()
    sc[k == 0] = 0.0
    return sc, ids


def merge_keys_host(gathered: np.ndarray, k: int) -> np.ndarray:
    """numpy twin of ``hs_topk_merge``: [world, B, k] -> [B, k], k largest keys per query, descending."""
    g = np.asarray(gathered).astype(np.uint64)
    world, B, kk = g.shape
    flat = np.transpose(g, (1, 0, 2)).reshape(B, world * kk)
    return np.ascontiguousarray(np.sort(flat, axis=1)[:, ::-1][:, :k])
