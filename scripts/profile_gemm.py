"""Small fixed workloads for ncu.   python scripts/profile_gemm.py <what> [n_docs] [dim]
   what: filter_bf16 | filter_tf32 | store_bf16_256 | store_tf32 | hybrid_tf32 | hybrid_bf16x | hybrid_bf16x256 | hybrid_fp32"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 384
spec = synth.SynthSpec(n_docs=n, dim=dim)
dev = torch.device("cuda:0")
shard = synth_device.build_synthetic_shard(spec, 0, n, dev, lexical=what.startswith("hybrid"))
th = synth.zipf_thresholds(spec.vocab)
reps = 3
if what.startswith("filter"):
    mode, B = ("bf16", 256) if what == "filter_bf16" else ("tf32x3", 128)
    eng = SearchEngine(shard, max_batch=256, dense_mode=mode)
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B))
    for _ in range(reps):
        eng.search_semantic(qb, 100, 1.0, filtered=True)
elif what.startswith("store"):
    mode, B = ("bf16", 256) if what == "store_bf16_256" else ("tf32x3", 128)
    eng = SearchEngine(shard, max_batch=256, dense_mode=mode)
    qd = eng.upload_vectors(synth.query_embeddings(spec, 0, B)).clone()
    stats = eng._stats(B)
    for _ in range(reps):
        eng.dense_scan(qd, stats)
else:
    mode, B = {"hybrid_tf32": ("tf32x3", 128), "hybrid_bf16x": ("bf16_exact", 128),
               "hybrid_bf16x256": ("bf16_exact", 256)}.get(what, ("fp32", 8))
    eng = SearchEngine(shard, max_batch=B, dense_mode=mode)
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
    for _ in range(reps):
        eng.search_hybrid_bm25(qb, 100, 0.6, 0.4)
torch.cuda.synchronize()
print("done", what, n)
