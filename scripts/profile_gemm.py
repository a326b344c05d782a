"""Small fixed workload for ncu: bf16 tcgen05 GEMM dense scan, 2 M docs x 384-d, 128 queries."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import SearchEngine
spec = synth.SynthSpec(n_docs=2_000_000)
shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, "cuda:0", lexical=False)
eng = SearchEngine(shard, max_batch=128, dense_mode="bf16")
qd = eng.upload_vectors(synth.query_embeddings(spec, 0, 128)).clone()
stats = eng._stats(128)
for _ in range(4):
    eng.dense_scan(qd, stats, "bf16")
torch.cuda.synchronize()
print("ok")
