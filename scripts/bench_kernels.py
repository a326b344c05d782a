"""Kernel micro-benchmarks on one GPU (CUDA events): dense scan per (mode, batch), BM25, fuse+topk."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

ap = argparse.ArgumentParser()
ap.add_argument("--n-docs", type=int, default=2_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--vocab", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--lexical", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
spec = synth.SynthSpec(n_docs=a.n_docs, vocab=a.vocab, dim=a.dim)
t0 = time.perf_counter()
shard = synth_device.build_synthetic_shard(spec, 0, a.n_docs, dev, lexical=bool(a.lexical))
torch.cuda.synchronize(); print(f"build {time.perf_counter()-t0:.1f}s  postings={0 if shard.postings is None else shard.postings.shape[0]}")
eng = SearchEngine(shard, max_batch=64)
PEAK = 6547.2

def timeit(fn, iters=a.iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

qv = synth.query_embeddings(spec, 0, 64)
bytes_pass = a.n_docs * shard.ld * 4
for mode in ("exact", "fp32"):
    for B in (1, 2, 4, 8, 16, 32):
        qd = eng.upload_vectors(qv[:B]).clone()
        stats = eng._stats(B)
        ms = timeit(lambda: eng.dense_scan(qd, stats, mode))
        nl = eng.dense_launches(B, mode)
        print(json.dumps({"kernel": "dense", "mode": mode, "B": B, "ms": round(ms, 4), "launches": nl,
                          "GBps_per_launch": round(bytes_pass * nl / ms / 1e6, 1), "frac": round(bytes_pass * nl / ms / 1e6 / PEAK, 3),
                          "qps": round(B / ms * 1e3, 1)}))
if a.lexical:
    th = synth.zipf_thresholds(spec.vocab)
    qt = synth.query_terms(spec, 0, 64, th).tolist()
    indptr = shard.indptr.cpu().numpy()
    for B in (1, 8, 32):
        t, i, o = [x.clone() for x in eng.upload_terms(qt[:B])]
        stats = eng._stats(B)
        P = sum(int(indptr[x + 1] - indptr[x]) for q in qt[:B] for x in q)
        nt = eng._n_tokens
        ms = timeit(lambda: eng.bm25_score(t, i, o, B, stats, nt))
        alg = 8 * P + 8 * a.n_docs * B
        print(json.dumps({"kernel": "bm25", "B": B, "ms": round(ms, 4), "postings": P, "GBps": round(alg / ms / 1e6, 1),
                          "frac": round(alg / ms / 1e6 / PEAK, 3)}))
        qd = eng.upload_vectors(qv[:B]).clone()
        cos = eng.dense_scan(qd, stats, "fp32"); bm = eng.bm25_score(t, i, o, B, stats, nt)
        ms = timeit(lambda: eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, 100))
        alg = 8 * a.n_docs * B
        print(json.dumps({"kernel": "fuse_topk", "B": B, "k": 100, "ms": round(ms, 4), "GBps": round(alg / ms / 1e6, 1),
                          "frac": round(alg / ms / 1e6 / PEAK, 3)}))
