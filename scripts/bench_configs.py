"""Secondary measurements for the other BASELINE.json configs on ONE B200 (CUDA events, device-resident
inputs; bench.py stays the headline).  Prints one JSON line per measurement.

    python scripts/bench_configs.py [--which 2,3,4,5]
"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_search_engine_b200 import synth, synth_device
from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

PEAK_HBM, PEAK_BF16 = 6547.2, 1402.2
ap = argparse.ArgumentParser()
ap.add_argument("--which", default="2,3,4,5")
ap.add_argument("--bm25-docs", type=int, default=20_000_000)
a = ap.parse_args()
which = set(a.which.split(","))
dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


if "2" in which:      # hybrid_bm25, 1 M docs x 384, batch 1 and 256
    spec = synth.SynthSpec(n_docs=1_000_000)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev)
    th = synth.zipf_thresholds(spec.vocab)
    for B, mode in ((1, "exact"), (1, "fp32"), (256, "fp32"), (256, "bf16")):
        eng = SearchEngine(shard, max_batch=32 if mode != "bf16" else 128, dense_mode=mode)
        qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
        ms = timeit(lambda: eng.search_hybrid_bm25(qb, 100, 0.6, 0.4), iters=5 if B > 1 else 20)
        emit(config=2, pipeline="hybrid_bm25", n_docs=spec.n_docs, dim=384, batch=B, dense_mode=mode, ms=round(ms, 3),
             qps=round(B / ms * 1e3, 1), note="through SearchEngine.search_hybrid_bm25 (host query upload included)")
    del shard, eng
    torch.cuda.empty_cache()

if "3" in which:      # pure BM25, Zipfian inverted index, top-100
    spec = synth.SynthSpec(n_docs=a.bm25_docs)
    t0 = time.perf_counter()
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev, dense=False)
    torch.cuda.synchronize()
    build = time.perf_counter() - t0
    th = synth.zipf_thresholds(spec.vocab)
    eng = SearchEngine(shard, max_batch=32)
    indptr = shard.indptr.cpu().numpy()
    for B in (1, 8, 32):
        terms = synth.query_terms(spec, 0, B, th).tolist()
        qt, qi, qo = [t.clone() for t in eng.upload_terms(terms)]
        nt = eng._n_tokens
        P = sum(int(indptr[x + 1] - indptr[x]) for q in terms for x in q)

        def step():
            bm = eng.bm25_score(qt, qi, qo, B, None, nt)
            return eng.unpack(eng.fuse_topk(0, bm, None, None, 1.0, 0.0, 100))
        ms = timeit(step)
        alg = 8 * P + 8 * spec.n_docs * B
        emit(config=3, pipeline="bm25", n_docs=spec.n_docs, postings=int(shard.postings.shape[0]), batch=B, ms=round(ms, 3),
             qps=round(B / ms * 1e3, 1), postings_touched=P, alg_GBps=round(alg / ms / 1e6, 1),
             frac_hbm=round(alg / ms / 1e6 / PEAK_HBM, 3), index_build_s=round(build, 1))
    del shard, eng
    torch.cuda.empty_cache()

if "4" in which:      # multi_stage stages 1-2: dense top-100 (bf16 tcgen05 GEMM) -> BM25 on the 100 -> top-20
    spec = synth.SynthSpec(n_docs=10_000_000, dim=768)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev)
    shard.ensure_bf16()
    th = synth.zipf_thresholds(spec.vocab)
    eng = SearchEngine(shard, max_batch=128, dense_mode="bf16")
    for B in (128, 1024):
        qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())

        def step():
            _, ids1 = eng.search_semantic(qb, 100, 1.0)
            return eng.bm25_score_docs(qb.term_ids, ids1)
        ms = timeit(step, iters=3, warm=1)
        # the GEMM alone, device-resident
        qd = eng.upload_vectors(qb.vectors[:128]).clone()
        stats = eng._stats(128)
        g = timeit(lambda: eng.dense_scan(qd, stats, "bf16"), iters=5)
        flops = 2.0 * 128 * spec.n_docs * 768
        byts = spec.n_docs * 768 * 2 + 128 * spec.n_docs * 4
        emit(config=4, pipeline="multi_stage stages 1-2", n_docs=spec.n_docs, dim=768, batch=B, ms=round(ms, 2),
             qps=round(B / ms * 1e3, 1), gemm_ms_per_128q=round(g, 3), gemm_tflops=round(flops / g / 1e9, 1),
             gemm_frac_bf16_peak=round(flops / g / 1e9 / PEAK_BF16, 3), gemm_GBps=round(byts / g / 1e6, 1),
             gemm_frac_hbm=round(byts / g / 1e6 / PEAK_HBM, 3))
    del shard, eng
    torch.cuda.empty_cache()

if "5" in which:      # diversity: MMR over 1000 candidates x 384-d, top 250
    spec = synth.SynthSpec(n_docs=2_000_000)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev, lexical=False)
    eng = SearchEngine(shard)
    g = torch.Generator(device="cpu").manual_seed(0)
    for B in (64, 1024, 4096):
        cand = torch.stack([torch.randperm(spec.n_docs, generator=g)[:1000] for _ in range(min(B, 256))])
        cand = cand.repeat((B + cand.shape[0] - 1) // cand.shape[0], 1)[:B].to(dev)
        rel = torch.linspace(1.0, 0.0, 1000, dtype=torch.float64).repeat(B, 1).to(dev)
        ms = timeit(lambda: eng.mmr(cand, rel, 0.5, 250), iters=2, warm=1)
        emit(config=5, pipeline="diversity MMR", candidates=1000, k=250, dim=384, batch=B, ms=round(ms, 2),
             us_per_query=round(ms * 1e3 / B, 1), qps=round(B / ms * 1e3, 1))
