"""ctypes binding of the C-ABI library ``libhs_b200.so`` (include/hs_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.  PyTorch is used only
for device memory, streams and ``torch.distributed`` -- tensors go in as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart before our library so both share one runtime)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhs_b200.so")

HS_DENSE_EXACT, HS_DENSE_FP32, HS_DENSE_BF16, HS_DENSE_TF32X3 = 0, 1, 2, 3
HS_FUSE_RAW, HS_FUSE_SEARCHER, HS_FUSE_HYBRID_BM25 = 0, 1, 2
HS_TOPK_MAX = 2048
# "bf16_exact": screen with the bf16 tensor-core GEMM, verify in the conformance order -> the exact mode's results
DENSE_MODES = {"exact": HS_DENSE_EXACT, "fp32": HS_DENSE_FP32, "bf16": HS_DENSE_BF16, "tf32x3": HS_DENSE_TF32X3,
               "bf16_exact": HS_DENSE_BF16}
# bound on |score - exact cosine| of a tensor-core mode: bf16 operands carry a relative rounding error <= 2^-9 each, so
# |sum q~v~ - sum qv| <= (2^-8 + 2^-18) sum|q_i v_i| <= 2^-8 |q||v| (Cauchy-Schwarz); + 2^-13 for the tensor core's
# truncating float32 accumulation (measured 2^-20) and the float32 scaling by the norms
VERIFY_EPS = {"bf16_exact": 2.0 ** -8 * 1.001 + 2.0 ** -13 + 2.0 ** -20}
# the screen scores are STORED as binary16 (half the bytes of the GEMM's write and the select's read): |score| < 2, so the
# stored value is within 2^-11 of the float32 accumulator -- added to the eps of everything that reads the stored screen
SCREEN_F16_EPS = 2.0 ** -11

_vp, _i32, _i64, _u32, _u64, _f64, _sz = (C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64,
                                          C.c_double, C.c_size_t)

# name -> (restype, argtypes); must list every symbol include/hs_b200.h declares
SIGNATURES = {
    "hs_abi_version": (C.c_int, []),
    "hs_last_error": (C.c_char_p, []),
    "hs_index_create": (C.c_int, [C.c_int, _i64, _i64, C.POINTER(_vp)]),
    "hs_index_destroy": (C.c_int, [_vp]),
    "hs_index_set_dense": (C.c_int, [_vp, _vp, _i32, _i64, _vp]),
    "hs_index_set_csr": (C.c_int, [_vp, _vp, _vp, _i64, _i64]),
    "hs_index_set_doc_stats": (C.c_int, [_vp, _vp, _f64, _f64, _f64, _vp, _u32, _u32, _vp]),
    "hs_row_norms": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp]),
    "hs_bm25_impact_table": (C.c_int, [_f64, _f64, _f64, _u32, _u32, _vp, _vp]),
    "hs_bm25_build_hot": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "hs_stats_reset": (C.c_int, [_vp, _i32, _vp]),
    "hs_stats_decode": (C.c_int, [_vp, _vp, _i32, _vp]),
    "hs_stats_encode": (C.c_int, [_vp, _vp, _i32, _vp]),
    "hs_stats_to_maxform": (C.c_int, [_vp, _vp, _i32, _vp]),
    "hs_stats_from_maxform": (C.c_int, [_vp, _vp, _i32, _vp]),
    "hs_stats_fold_minmax": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "hs_dense_scan": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "hs_index_set_dense_bf16": (C.c_int, [_vp, _vp, _i64]),
    "hs_dense_gemm_workspace_bytes": (_sz, [_vp, _i32, _i32]),
    "hs_dense_gemm": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _vp, _sz, _vp, _i64, _vp, _vp]),
    "hs_dense_gemm_filter": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _vp, _sz, _vp, _vp, _i32, _vp, _vp, _vp]),
    "hs_topk_select": (C.c_int, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _sz, _vp, _vp]),
    "hs_keys_kth_score": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "hs_dense_gemm_filter_segments": (_i32, [_vp, _i32]),
    "hs_dense_gemm_ext": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _vp, _sz, _vp, _i64, _vp, _vp, _vp, _i32, _f64, _vp]),
    "hs_dense_gemm_ext_f16": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _vp, _sz, _vp, _i64, _vp, _vp, _vp, _i32, _f64, _vp]),
    "hs_verify_stats": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _f64, _vp, _vp, _vp]),
    "hs_verify_topk": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _f64, _f64, _vp, _i32, _i32, _f64, _vp, _vp, _vp]),
    "hs_cand_select": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _f64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hs_bm25_workspace_bytes": (_sz, [_i64, _i32]),
    "hs_bm25_score": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "hs_bm25_score_f16": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _sz, _vp, _i64, _vp, _vp]),
    "hs_bm25plus_score": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _f64, _vp, _sz, _vp, _vp, _vp]),
    "hs_bm25_score_docs": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp]),
    "hs_bm25plus_score_docs": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _i32, _f64, _vp, _vp]),
    "hs_fuse_topk_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "hs_fuse_topk": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _f64, _f64, _i32, _i32, _vp, _vp, _sz, _vp, _vp]),
    "hs_fuse_topk_f16": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i64, _vp, _f64, _f64, _i32, _i32, _vp, _sz, _vp, _vp]),
    "hs_keys_local_docs": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "hs_verify_topk_cand": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _f64, _vp, _f64, _f64, _vp, _i32, _i32, _f64, _vp, _vp, _vp]),
    "hs_topk_merge": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "hs_scatter_keys": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _vp, _vp]),
    "hs_keys_unpack": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "hs_mmr_workspace_bytes": (_sz, [_i32, _i32]),
    "hs_mmr": (C.c_int, [_vp, _vp, _vp, _f64, _i32, _i32, _i32, _vp, _sz, _vp, _vp]),
    "hs_lexical_scores": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "hs_token_flags": (C.c_int, [_vp, _i64, _vp, _vp]),
    "hs_token_hashes": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "hs_radix_sort_workspace_bytes": (_sz, [_i64]),
    "hs_radix_sort_u64": (C.c_int, [_vp, _vp, _i64, _u32, _vp, _sz, _vp]),
    "hs_lower_bound_i64": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "hs_scan_workspace_bytes": (_sz, [_i64]),
    "hs_exclusive_scan_i64": (C.c_int, [_vp, _vp, _i64, _vp, _sz, _vp]),
    "hs_rle_workspace_bytes": (_sz, [_i64]),
    "hs_run_length_encode_u64": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hs_term_doc_freqs": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "hs_synth_embeddings": (C.c_int, [_vp, _i64, _i64, _i32, _i64, _u64, _vp]),
    "hs_synth_doc_lengths": (C.c_int, [_vp, _i64, _i64, _u64, _u32, _u32, _vp]),
    "hs_synth_token_keys": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _u64, _vp, _i32, _vp]),
}

_lib = None


class HsError(RuntimeError):
    pass


def load():
    """Load the library (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsError(
            f"{LIB_PATH} is missing: build it with `python -m hybrid_search_engine_b200.build_native` "
            "(or __graft_entry__.build()). The hot path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().hs_last_error().decode("utf-8", "replace")
        raise HsError(f"{what or 'hs call'} failed (status {rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream
