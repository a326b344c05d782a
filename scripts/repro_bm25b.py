import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
rank, world = int(sys.argv[1]), int(sys.argv[2])
spec = synth.SynthSpec(n_docs=10_000_000)
per = (spec.n_docs + world - 1) // world
lo, hi = rank * per, min(spec.n_docs, (rank + 1) * per)
shard = synth_device.build_synthetic_shard(spec, lo, hi, "cuda:0", dense=False)
torch.cuda.synchronize(); print("built", lo, hi, flush=True)
ip = shard.indptr; post = shard.postings
docs = post[:, 0].to(torch.int64) & 0xFFFFFFFF
print("P", post.shape[0], "indptr[-1]", int(ip[-1]), "max doc", int(docs.max()), "n", shard.n_docs, "min tf", int(post[:,1].min()), "max tf", int(post[:,1].max()))
# ascending within term
term_of = torch.repeat_interleave(torch.arange(ip.numel()-1, device="cuda"), ip[1:]-ip[:-1])
bad = ((docs[1:] <= docs[:-1]) & (term_of[1:] == term_of[:-1])).nonzero()
print("non-ascending positions:", bad.numel(), bad[:5].flatten().tolist())
print("dl max", int(shard.dl.max()), "max_dl", shard.max_dl, "tf_cap", shard.tf_cap)
eng = SearchEngine(shard, max_batch=8)
th = synth.zipf_thresholds(spec.vocab)
qt_all = synth.query_terms(spec, 0, 1024, th).tolist()
B = 8
n_tiles = (shard.n_docs + 4095) // 4096
for s in range(25):
    idx = [(s * B + j) % 1024 for j in range(B)]
    terms = [qt_all[i] for i in idx]
    qt, qi, qo = eng.upload_terms(terms)
    T = eng._n_tokens
    # run only the ranges kernel through the public call? emulate on host instead:
    flat = qt[:T].cpu().numpy()
    ok = True
    for t in flat:
        a, b = int(ip[t]), int(ip[t + 1])
        d = docs[a:b]
        if d.numel() and not bool((d[1:] > d[:-1]).all()): ok = False
    print("step", s, "terms", terms, "T", T, "df", [int(ip[t+1]-ip[t]) for t in flat][:8], flush=True)
    stats = eng._stats(B)
    bm = eng.bm25_score(qt, qi, qo, B, stats)
    torch.cuda.synchronize()
    ws = eng._bufs["bm25_ws"][: T * (n_tiles + 1)].view(T, n_tiles + 1).cpu()
    # reference ranges
    bounds = torch.arange(0, n_tiles + 1, device="cuda") * 4096
    for ti, t in enumerate(flat):
        a, b = int(ip[t]), int(ip[t + 1])
        ref = a + torch.searchsorted(docs[a:b].contiguous(), bounds)
        ref[0] = a; ref[-1] = b
        if not torch.equal(ref.cpu(), ws[ti]):
            diff = (ref.cpu() != ws[ti]).nonzero().flatten()
            print("  RANGE MISMATCH term", t, "first diffs at boundaries", diff[:5].tolist(), ref.cpu()[diff[:3]].tolist(), ws[ti][diff[:3]].tolist(), flush=True)
    print("   ok max", float(bm.max()), flush=True)
