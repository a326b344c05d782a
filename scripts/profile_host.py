"""Host-side cost of one serving-loop batch: a corpus small enough that the GPU work is negligible, so the loop's
throughput is the host's (flattening, uploads, ~25 launches, result hand-over).  Prints ms per batch and the cProfile top."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from hybrid_search_engine_b200 import synth, synth_device  # noqa: E402
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16_exact"
spec = synth.SynthSpec(n_docs=100_000, vocab=1_000_000, dim=384)
shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"))
if mode in ("bf16", "bf16_exact"):
    shard.ensure_bf16()
th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
eng = SearchEngine(shard, max_batch=256 if mode != "exact" else 8, dense_mode=mode)
batches = [QueryBatch(vectors=synth.query_embeddings(spec, i * B, (i + 1) * B),
                      term_ids=synth.query_terms(spec, i * B, (i + 1) * B, th).tolist()) for i in range(8)]
for _ in eng.search_hybrid_bm25_stream(batches[:4], 100, 0.6, 0.4):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
for _ in eng.search_hybrid_bm25_stream(batches * 5, 100, 0.6, 0.4):
    n += 1
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"B={B} mode={mode}: {dt / n * 1e3:.3f} ms per batch ({n} batches), {n * B / dt:.0f} q/s host-bound")
pr = cProfile.Profile()
pr.enable()
for _ in eng.search_hybrid_bm25_stream(batches * 5, 100, 0.6, 0.4):
    pass
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
