"""Small fixed workload for ncu: a few hybrid_bm25 steps on a synthetic shard (1 GPU)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

ap = argparse.ArgumentParser()
ap.add_argument("--n-docs", type=int, default=4_000_000)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--mode", default="fp32")
a = ap.parse_args()
spec = synth.SynthSpec(n_docs=a.n_docs)
shard = synth_device.build_synthetic_shard(spec, 0, a.n_docs, "cuda:0")
eng = SearchEngine(shard, max_batch=a.batch, dense_mode=a.mode)
th = synth.zipf_thresholds(spec.vocab)
for s in range(a.steps):
    qb = QueryBatch(vectors=synth.query_embeddings(spec, s * a.batch, (s + 1) * a.batch),
                    term_ids=synth.query_terms(spec, s * a.batch, (s + 1) * a.batch, th).tolist())
    sc, ids = eng.search_hybrid_bm25(qb, 100, 0.6, 0.4)
torch.cuda.synchronize()
print("ok", ids[0, :5].tolist())
