import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
rank, world = int(sys.argv[1]), int(sys.argv[2])
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 25
spec = synth.SynthSpec(n_docs=10_000_000)
per = (spec.n_docs + world - 1) // world
lo, hi = rank * per, min(spec.n_docs, (rank + 1) * per)
# global df from a full-corpus pass is expensive; emulate with shard df (only affects idf values / known terms)
shard = synth_device.build_synthetic_shard(spec, lo, hi, "cuda:0", dense=False)
torch.cuda.synchronize(); print("built", lo, hi, flush=True)
eng = SearchEngine(shard, max_batch=8)
th = synth.zipf_thresholds(spec.vocab)
qt_all = synth.query_terms(spec, 0, 1024, th).tolist()
B = 8
for s in range(nb):
    idx = [(s * B + j) % 1024 for j in range(B)]
    qt, qi, qo = eng.upload_terms([qt_all[i] for i in idx])
    stats = eng._stats(B)
    bm = eng.bm25_score(qt, qi, qo, B, stats)
    torch.cuda.synchronize()
    print("step", s, "ok", float(bm.max()), flush=True)
