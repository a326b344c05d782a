"""Device index-build primitives behind the C ABI (csrc/sort.cu): radix sort, prefix sum, run-length encoding and
per-term posting counts over 64-bit ``term << 32 | doc`` keys.  Hand-written replacements for the library
sort / unique / bincount / cumsum calls the first version of the device index build used.
torch is used for the buffers only."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max((int(nbytes) + 7) // 8, 1), dtype=torch.int64, device=device)


def byte_mask_for(*fields) -> int:
    """Byte mask of a key made of ``(shift, max_value)`` fields: the bytes that can differ between keys."""
    mask = 0
    for shift, vmax in fields:
        bits = max(int(vmax), 1).bit_length()
        for b in range(shift // 8, (shift + bits + 7) // 8):
            mask |= 1 << b
    return mask & 0xFF


def sort_keys_(keys: torch.Tensor, byte_mask: int = 0xFF) -> torch.Tensor:
    """Sort int64 keys (non-negative, as uint64) ascending in place; ``byte_mask`` names the bytes that vary."""
    lib = _lib.load()
    n = keys.numel()
    if n <= 1:
        return keys
    assert keys.dtype == torch.int64 and keys.is_contiguous() and keys.is_cuda
    dev = keys.device
    tmp = torch.empty_like(keys)
    nbytes = lib.hs_radix_sort_workspace_bytes(n)
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev):
        check(lib.hs_radix_sort_u64(ptr(keys), ptr(tmp), n, int(byte_mask), ptr(ws), ws.numel() * 8, stream_ptr(dev)),
              "hs_radix_sort_u64")
    return keys


def exclusive_scan(x: torch.Tensor) -> torch.Tensor:
    """int64 [n] -> int64 [n + 1]: out[i] = sum(x[:i]) (out[n] = total)."""
    lib = _lib.load()
    n = x.numel()
    dev = x.device
    ext = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ext[:n] = x
    ws = _ws(lib.hs_scan_workspace_bytes(n + 1), dev)
    with torch.cuda.device(dev):
        check(lib.hs_exclusive_scan_i64(ptr(ext), ptr(ext), n + 1, ptr(ws), ws.numel() * 8, stream_ptr(dev)),
              "hs_exclusive_scan_i64")
    return ext


def run_length_encode(sorted_keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sorted int64 keys -> (unique keys int64 [m], run lengths int32 [m])."""
    lib = _lib.load()
    n = sorted_keys.numel()
    dev = sorted_keys.device
    if n == 0:
        return sorted_keys, torch.zeros(0, dtype=torch.int32, device=dev)
    uniq = torch.empty(n, dtype=torch.int64, device=dev)
    start = torch.empty(n + 1, dtype=torch.int64, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _ws(lib.hs_rle_workspace_bytes(n), dev)
    with torch.cuda.device(dev):
        check(lib.hs_run_length_encode_u64(ptr(sorted_keys), n, ptr(uniq), ptr(start), ptr(counts), ptr(total), ptr(ws),
                                           ws.numel() * 8, stream_ptr(dev)), "hs_run_length_encode_u64")
    m = int(total.item())
    return uniq[:m], counts[:m]


def term_doc_freqs(uniq_keys: torch.Tensor, n_terms: int) -> torch.Tensor:
    """Sorted unique ``term << 32 | doc`` keys -> int64 [n_terms] postings per term (df, bm25.py:66-67)."""
    lib = _lib.load()
    dev = uniq_keys.device
    df = torch.zeros(max(n_terms, 0), dtype=torch.int64, device=dev)
    if n_terms <= 0:
        return df
    scratch = torch.empty(2 * n_terms, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.hs_term_doc_freqs(ptr(uniq_keys), uniq_keys.numel(), n_terms, ptr(scratch), ptr(df), stream_ptr(dev)),
              "hs_term_doc_freqs")
    return df


def lower_bound(sorted_vals: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    """int64 [m]: first index in ``sorted_vals`` whose value is >= the query (term id of a token hash)."""
    lib = _lib.load()
    dev = queries.device
    out = torch.empty(queries.numel(), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.hs_lower_bound_i64(ptr(sorted_vals), sorted_vals.numel(), ptr(queries.contiguous()), queries.numel(),
                                     ptr(out), stream_ptr(dev)), "hs_lower_bound_i64")
    return out
