"""BM25 kernel A/B on one GPU: HS_BM25_IMPL=tile vs the pipelined default; prints time, fraction of the HBM
peak on algorithmic bytes (8 P + 8 N per query) and a checksum of the score bits (must be equal)."""
import argparse, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--n-docs", type=int, default=4_000_000)
ap.add_argument("--child", default="")
a = ap.parse_args()
if not a.child:
    for impl in ("tile", "batch"):
        env = dict(os.environ, HS_BM25_IMPL=impl)
        subprocess.run([sys.executable, __file__, "--n-docs", str(a.n_docs), "--child", impl], env=env, check=True)
    sys.exit(0)

import torch
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import SearchEngine
spec = synth.SynthSpec(n_docs=a.n_docs, dim=16)
shard = synth_device.build_synthetic_shard(spec, 0, a.n_docs, "cuda:0")
eng = SearchEngine(shard, max_batch=32)
th = synth.zipf_thresholds(spec.vocab)
qt = synth.query_terms(spec, 0, 32, th).tolist()
indptr = shard.indptr.cpu().numpy()
for B in (1, 8, 32):
    t, i, o = [x.clone() for x in eng.upload_terms(qt[:B])]
    nt = eng._n_tokens
    stats = eng._stats(B)
    P = sum(int(indptr[x + 1] - indptr[x]) for q in qt[:B] for x in q)
    for _ in range(3):
        out = eng.bm25_score(t, i, o, B, stats, nt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = eng.bm25_score(t, i, o, B, stats, nt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    alg = 8 * P + 8 * a.n_docs * B
    v = out.view(torch.int32).to(torch.int64)
    chk = int((v * (torch.arange(v.numel(), device=v.device).view_as(v) % 1000003 + 1)).sum().item())
    print(json.dumps({"impl": a.child, "B": B, "ms": round(ms, 4), "postings": P, "GBps": round(alg / ms / 1e6, 1),
                      "frac": round(alg / ms / 1e6 / 6547.2, 3), "checksum": chk}))
