import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_search_engine_b200 as hs
from oracle import hybrid_oracle as orc
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
docs = []
for i in range(n):
    toks = ["t0"] * int(rng.integers(1, 20)) + [f"t{int(x)}" for x in rng.integers(1, 500, size=int(rng.integers(5, 40)))]
    docs.append(" ".join(toks))
bm = hs.BM25(); bm.fit(docs)
for q in ["t0 t77 t0 t0", "t0 t0 t0 t0 t0 t0", "t5 t0 t0 t1", "t0 t3 t0 t0 t9 t0 t0"]:
    got = bm.score_batch(q); torch.cuda.synchronize()
    st = orc.bm25_fit(docs)
    want = orc.bm25_score_batch(st, q)
    print(q, "equal:", np.array_equal(got, want), float(got.max()), flush=True)
