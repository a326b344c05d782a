"""Per-stage CUDA-event timings of one hybrid_bm25 step (device-resident inputs) on one GPU."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import SearchEngine

ap = argparse.ArgumentParser()
ap.add_argument("--n-docs", type=int, default=10_000_000)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
spec = synth.SynthSpec(n_docs=a.n_docs)
shard = synth_device.build_synthetic_shard(spec, 0, a.n_docs, "cuda:0")
eng = SearchEngine(shard, max_batch=a.batch, dense_mode="fp32")
th = synth.zipf_thresholds(spec.vocab)
B = a.batch
qd = eng.upload_vectors(synth.query_embeddings(spec, 0, B)).clone()
qt, qi, qo = [t.clone() for t in eng.upload_terms(synth.query_terms(spec, 0, B, th).tolist())]
nt = eng._n_tokens

def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / a.iters * 1000, 1)

stats = eng._stats(B)
cos = eng.dense_scan(qd, stats); bm = eng.bm25_score(qt, qi, qo, B, stats, nt)
out = {"n_docs": a.n_docs, "B": B,
       "dense_us": timeit(lambda: eng.dense_scan(qd, stats)),
       "bm25_us": timeit(lambda: eng.bm25_score(qt, qi, qo, B, stats, nt)),
       "fuse_topk_us": timeit(lambda: eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, a.k)),
       "step_us": timeit(lambda: eng.hybrid_step_device(qd, qt, qi, qo, B, nt, a.k, 0.6, 0.4))}
print(json.dumps(out))
