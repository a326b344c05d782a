"""Dense scan tuning: time every (BQ, QG) variant through HS_DENSE_FORCE (1 GPU, CUDA events)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import SearchEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
spec = synth.SynthSpec(n_docs=n)
shard = synth_device.build_synthetic_shard(spec, 0, n, "cuda:0", lexical=False)
eng = SearchEngine(shard, max_batch=64)
qv = synth.query_embeddings(spec, 0, 64)
bytes_pass = n * shard.ld * 4
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for mode, caps in (("fp32", (1, 2, 4, 8)), ("exact", (1, 2, 4))):
    for bq in caps:
        for qg in (1, 2, 4):
            B = bq * qg
            os.environ["HS_DENSE_FORCE"] = f"{bq},{qg}"
            qd = eng.upload_vectors(qv[:B]).clone()
            stats = eng._stats(B)
            ms = timeit(lambda: eng.dense_scan(qd, stats, mode))
            print(json.dumps({"mode": mode, "bq": bq, "qg": qg, "B": B, "ms": round(ms, 4),
                              "GBps": round(bytes_pass / ms / 1e6, 1), "frac": round(bytes_pass / ms / 1e6 / 6547.2, 3),
                              "qps": round(B / ms * 1e3, 1)}))
