import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import SearchEngine
from hybrid_search_engine_b200.index import DeviceIndex
mode = sys.argv[1]
spec = synth.SynthSpec(n_docs=10_000_000)
lo, hi = 7_500_000, 10_000_000
if mode == "build":
    shard = synth_device.build_synthetic_shard(spec, lo, hi, "cuda:0", dense=False)
    torch.save({"indptr": shard.indptr.cpu(), "postings": shard.postings.cpu(), "dl": shard.dl.cpu(), "avgdl": shard.avgdl,
                "df": shard.df_host}, "/tmp/shard.pt")
    print("saved")
else:
    d = torch.load("/tmp/shard.pt", weights_only=False)
    shard = DeviceIndex("cuda:0", hi - lo, doc_base=lo)
    shard.set_bm25(d["indptr"], d["postings"], d["dl"], d["avgdl"], d["df"], spec.n_docs, max_dl=300)
    eng = SearchEngine(shard, max_batch=8)
    terms = [[0, 486329, 0, 0], [29423, 1, 135, 11615], [203892, 1168, 878653, 52], [19610, 679, 102309, 549], [189446, 1, 30152, 167], [2331, 1, 2127, 1574], [1756, 569369, 0, 4697], [12905, 8, 45122, 6]]
    sel = terms if len(sys.argv) < 3 else [terms[int(sys.argv[2])]]
    qt, qi, qo = eng.upload_terms(sel)
    stats = eng._stats(len(sel))
    bm = eng.bm25_score(qt, qi, qo, len(sel), stats)
    torch.cuda.synchronize()
    print("ok", float(bm.max()))
