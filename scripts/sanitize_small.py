"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every hot-path kernel family once, at sizes
that finish in seconds under the tool.   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_search_engine_b200 as hs
from hybrid_search_engine_b200 import synth, synth_device, devsort
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

which = set((sys.argv[1] if len(sys.argv) > 1 else "bm25,topk,dense,gemm,mmr,sort").split(","))
dev = torch.device("cuda:0")
spec = synth.SynthSpec(n_docs=20_000, vocab=3_000, dim=96, min_len=20, max_len=60)
shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev)            # sort.cu + synth.cu + hot-term build
th = synth.zipf_thresholds(spec.vocab)
B = 12
qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
eng = SearchEngine(shard, max_batch=8)
if "bm25" in which or "topk" in which:
    for mode in ("exact", "fp32"):
        s, i = eng.search_hybrid_bm25(qb, 100, 0.6, 0.4, dense_mode=mode)        # K2 + K1 (batch kernel, hot path) + K3/K4
    s, i = eng.search_bm25(QueryBatch(term_ids=qb.term_ids), 50)
    s, i = eng.search_semantic(QueryBatch(vectors=qb.vectors), 300, 1.0)
    print("hybrid / bm25 / semantic ok", int(i[0, 0]))
if "gemm" in which:
    for mode in ("tf32x3", "bf16"):
        s, i = eng.search_hybrid_bm25(qb, 100, 0.6, 0.4, dense_mode=mode)
    spec2 = synth.SynthSpec(n_docs=270_000, vocab=3_000, dim=64)
    sh2 = synth_device.build_synthetic_shard(spec2, 0, spec2.n_docs, dev, lexical=False)
    e2 = SearchEngine(sh2, max_batch=256)
    q2 = QueryBatch(vectors=synth.query_embeddings(spec2, 0, 140))
    for mode in ("bf16", "tf32x3"):
        s, i = e2.search_semantic(q2, 100, 1.0, dense_mode=mode, filtered=True)   # GEMM filter epilogue (clusters of 2 in bf16)
    print("gemm ok", int(i[0, 0]))
if "mmr" in which:
    cand = torch.stack([torch.randperm(spec.n_docs, device=dev)[:200] for _ in range(3)])
    rel = torch.rand((3, 200), dtype=torch.float64, device=dev)
    a = eng.mmr(cand, rel, 0.5, 40)                                               # cluster / DSMEM kernel
    b = eng.mmr(cand[:, :40].contiguous(), rel[:, :40].contiguous(), 0.5, 10)     # single-CTA kernel
    print("mmr ok", a[0, :3].tolist(), b[0, :3].tolist())
if "sort" in which:
    k = torch.randint(0, 2 ** 40, (100_000,), device=dev, dtype=torch.int64)
    s = devsort.sort_keys_(k.clone())
    u, c = devsort.run_length_encode(s)
    print("sort ok", bool((s[1:] >= s[:-1]).all()), int(c.sum()))
torch.cuda.synchronize()
print("sanitize workload done")
