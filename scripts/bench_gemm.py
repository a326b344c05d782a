"""Tensor-core dense paths at scale on ONE B200 (CUDA events, device-resident inputs).  One JSON line per point.

    python scripts/bench_gemm.py [--which tf32,bf16,filter,hybrid,cfg4] [--n-docs 10000000]
"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

PEAK_HBM, PEAK_BF16 = 6547.2, 1402.2
ap = argparse.ArgumentParser()
ap.add_argument("--which", default="tf32,bf16,filter,hybrid")
ap.add_argument("--n-docs", type=int, default=10_000_000)
a = ap.parse_args()
which = set(a.which.split(","))
dev = torch.device("cuda:0")


def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


if which & {"tf32", "bf16", "filter", "hybrid"}:
    spec = synth.SynthSpec(n_docs=a.n_docs)
    n, d = spec.n_docs, spec.dim
    shard = synth_device.build_synthetic_shard(spec, 0, n, dev, lexical="hybrid" in which)
    th = synth.zipf_thresholds(spec.vocab)
    eng = SearchEngine(shard, max_batch=256)
    if "tf32" in which:
        for B in (32, 128, 256):
            qd = eng.upload_vectors(synth.query_embeddings(spec, 0, B)).clone()
            stats = eng._stats(B)
            ms = timeit(lambda: eng.dense_scan(qd, stats, "tf32x3"))
            passes = (B + 127) // 128
            flops = 3 * 2.0 * B * n * d
            emit(kernel="dense_gemm tf32x3 STORE", n_docs=n, dim=d, batch=B, ms=round(ms, 3), ms_per_pass=round(ms / passes, 3),
                 dense_qps=round(B / ms * 1e3), tflops_3x=round(flops / ms / 1e9, 1), frac_tf32_peak=round(flops / ms / 1e9 / (PEAK_BF16 / 2), 3),
                 hbm_GBps=round((n * d * 4 * passes + B * n * 4) / ms / 1e6), frac_hbm=round((n * d * 4 * passes + B * n * 4) / ms / 1e6 / PEAK_HBM, 3))
    if "bf16" in which:
        shard.ensure_bf16()
        for B in (128, 256, 512):
            qd = eng.upload_vectors(synth.query_embeddings(spec, 0, B)).clone()
            stats = eng._stats(B)
            eng2 = SearchEngine(shard, max_batch=B)
            ms = timeit(lambda: eng2.dense_scan(qd, stats, "bf16"))
            passes = (B + 255) // 256
            flops = 2.0 * B * n * d
            byts = n * d * 2 * passes + B * n * 4
            emit(kernel="dense_gemm bf16 STORE", n_docs=n, dim=d, batch=B, ms=round(ms, 3), dense_qps=round(B / ms * 1e3),
                 tflops=round(flops / ms / 1e9, 1), frac_bf16_peak=round(flops / ms / 1e9 / PEAK_BF16, 3),
                 hbm_GBps=round(byts / ms / 1e6), frac_hbm=round(byts / ms / 1e6 / PEAK_HBM, 3))
            del eng2
    if "filter" in which:
        for mode, B in (("bf16", 256), ("bf16", 1024), ("tf32x3", 128), ("tf32x3", 512)):
            qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B))
            ms_f = timeit(lambda: eng.search_semantic(qb, 100, 1.0, dense_mode=mode, filtered=True), iters=3, warm=1)
            ms_s = timeit(lambda: eng.search_semantic(qb, 100, 1.0, dense_mode=mode, filtered=False), iters=2, warm=1) if B <= 256 else None
            flops = (3 if mode == "tf32x3" else 1) * 2.0 * B * n * d
            emit(path="semantic top-100, filter epilogue", mode=mode, n_docs=n, dim=d, batch=B, ms_filtered=round(ms_f, 3),
                 ms_stored=None if ms_s is None else round(ms_s, 3), qps=round(B / ms_f * 1e3),
                 frac_tensor_peak=round(flops / ms_f / 1e9 / (PEAK_BF16 / (2 if mode == "tf32x3" else 1)), 3))
    if "hybrid" in which:
        for mode, B, mb in (("fp32", 8, 8), ("fp32", 128, 8), ("tf32x3", 128, 128), ("tf32x3", 256, 128), ("bf16", 256, 256)):
            e3 = SearchEngine(shard, max_batch=mb, dense_mode=mode)
            qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
            ms = timeit(lambda: e3.search_hybrid_bm25(qb, 100, 0.6, 0.4), iters=3, warm=1)
            # parts, device resident
            qd = e3.upload_vectors(qb.vectors[:mb]).clone()
            qt, qi, qo = [t.clone() for t in e3.upload_terms(qb.term_ids[:mb])]
            nt = e3._n_tokens
            stats = e3._stats(mb)
            t_dense = timeit(lambda: e3.dense_scan(qd, stats), iters=3, warm=1)
            t_bm = timeit(lambda: e3.bm25_score(qt, qi, qo, mb, stats, nt), iters=3, warm=1)
            cos = e3.dense_scan(qd, stats); bm = e3.bm25_score(qt, qi, qo, mb, stats, nt)
            t_sel = timeit(lambda: e3.fuse_topk(2, cos, bm, stats, 0.6, 0.4, 100), iters=3, warm=1)
            emit(path="hybrid_bm25 step via SearchEngine.search_hybrid_bm25", mode=mode, n_docs=n, batch=B, sub_batch=mb, ms=round(ms, 3),
                 qps=round(B / ms * 1e3), per_sub_batch_ms=dict(dense=round(t_dense, 3), bm25=round(t_bm, 3), select=round(t_sel, 3)))
            del e3
    del shard, eng
    torch.cuda.empty_cache()

if "cfg4" in which:     # config 4: multi_stage stages 1-2, 10 M x 768 bf16, B = 1024
    spec = synth.SynthSpec(n_docs=a.n_docs, dim=768)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev)
    shard.ensure_bf16()
    th = synth.zipf_thresholds(spec.vocab)
    eng = SearchEngine(shard, max_batch=256, dense_mode="bf16")
    for B in (256, 1024):
        qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())

        def step():
            _, ids1 = eng.search_semantic(qb, 100, 1.0)
            return eng.bm25_score_docs(qb.term_ids, ids1 - shard.doc_base)
        ms = timeit(step, iters=3, warm=1)
        flops = 2.0 * B * spec.n_docs * 768
        emit(config=4, pipeline="multi_stage stages 1-2 (filter epilogue)", n_docs=spec.n_docs, dim=768, batch=B, ms=round(ms, 2),
             qps=round(B / ms * 1e3, 1), tensor_roofline_ms=round(flops / PEAK_BF16 / 1e9, 2),
             frac_tensor_roofline=round(flops / PEAK_BF16 / 1e9 / ms, 3))
