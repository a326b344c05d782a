import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
rank, world = int(sys.argv[1]), int(sys.argv[2])
spec = synth.SynthSpec(n_docs=10_000_000)
per = (spec.n_docs + world - 1) // world
lo, hi = rank * per, min(spec.n_docs, (rank + 1) * per)
print("shard", lo, hi, flush=True)
shard = synth_device.build_synthetic_shard(spec, lo, hi, "cuda:0")
torch.cuda.synchronize(); print("built", flush=True)
eng = SearchEngine(shard, max_batch=8, dense_mode="fp32")
th = synth.zipf_thresholds(spec.vocab)
qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, 8), term_ids=synth.query_terms(spec, 0, 8, th).tolist())
B = 8
stats = eng._stats(B); torch.cuda.synchronize(); print("stats ok", flush=True)
cos = eng.dense_scan(eng.upload_vectors(qb.vectors), stats); torch.cuda.synchronize(); print("dense ok", flush=True)
qt, qi, qo = eng.upload_terms(qb.term_ids)
bm = eng.bm25_score(qt, qi, qo, B, stats); torch.cuda.synchronize(); print("bm25 ok", flush=True)
keys = eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, 100); torch.cuda.synchronize(); print("fuse ok", flush=True)
sc, ids = eng.unpack(keys); torch.cuda.synchronize(); print("unpack ok", ids[0, :5].tolist(), flush=True)
