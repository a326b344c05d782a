"""Index-build throughput: host mirror of BM25.fit (python regex + dicts) vs the device build
(index_build.py).  Synthetic Zipfian texts (synth.doc_texts), docs/s over the whole fit() call, text
already in host memory.  python scripts/bench_index_build.py [--n-docs 200000]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hybrid_search_engine_b200 import synth  # noqa: E402
from hybrid_search_engine_b200.index import LexicalStats  # noqa: E402
from hybrid_search_engine_b200.index_build import DeviceLexicalStats  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-docs", type=int, default=200000)
ap.add_argument("--host-docs", type=int, default=50000)
a = ap.parse_args()
spec = synth.SynthSpec(n_docs=a.n_docs)
docs = synth.doc_texts(spec, 0, a.n_docs)
nbytes = sum(len(d) for d in docs)
t = time.perf_counter(); h = LexicalStats().fit(docs[:a.host_docs]); th = time.perf_counter() - t
DeviceLexicalStats("cuda:0").fit(docs[:1000])                       # warm-up (module load, allocator)
torch.cuda.synchronize()
t = time.perf_counter(); d = DeviceLexicalStats("cuda:0").fit(docs); torch.cuda.synchronize()
td = time.perf_counter() - t
print(json.dumps({"n_docs": a.n_docs, "text_MB": round(nbytes / 1e6, 1), "host_docs_per_s": round(a.host_docs / th),
                  "device_docs_per_s": round(a.n_docs / td), "device_s": round(td, 3),
                  "postings": int(d.postings.shape[0]), "terms": int(len(d.vocab_hashes))}))
