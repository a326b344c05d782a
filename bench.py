#!/usr/bin/env python
"""bench.py -- hybrid_bm25 queries/sec on a synthetic 10 M-doc x 384-d corpus (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the UNMODIFIED reference pipeline on the host cores
    python bench.py --workload bm25|multi_stage|diversity ...     # BASELINE configs 3 / 4 / 5

One "step" = one batch of B queries through the whole hot path: dense cosine scan over every doc,
BM25 over the CSR index, min-max / max normalisation + weighted fusion + top-100 select (+ the
stats all-reduce and top-k all-gather/merge when doc-sharded over N GPUs).  The corpus (10 M docs)
is FIXED as N grows (doc-sharded) => "scaling": "strong".

* ``value``  queries/s with the query batch already resident in HBM (device-timed with CUDA events,
             barrier + synchronize on both sides, max over ranks)
* ``e2e``    queries/s through the batched engine API with HOST inputs: pinned host -> device copy of
             the query vectors / term ids and device -> host read of the result inside every step,
             through the serving loop ``SearchEngine.search_hybrid_bm25_stream`` (batch i+1 is uploaded
             while batch i runs; every step's top-k is read back)
* ``e2e_pipeline``  the same through the reference-facing plugin API ``create_pipeline("hybrid_bm25")
             .search_many(query STRINGS, query_vectors=...)``: tokenisation, term lookup and the result
             dictionaries are inside the timed region
* ``roofline``  the dominant kernel of the step + ``kernels[]`` for dense / bm25 / select and the whole
             step against SURVEY.md section 8(d)'s algorithmic bytes (tensor modes: executed MMA flops
             against the measured bf16 peak, halved for TF32).  The default dense mode ``bf16_exact`` screens
             with the bf16 tcgen05 GEMM and verifies in the conformance order: its results are the ``exact``
             mode's bit for bit (``parity.verify_flagged_queries_in_timed_steps`` counts queries whose proof
             failed and that the serving API would redo exactly; the device-timed step does not redo them)
* ``parity``  top-k ids of the last batch vs the float64-accumulated ``exact`` mode, and the sha256 of the
             exact-mode top-100 (ids + scores) of a FIXED 64-query batch compared with the constant committed
             in tests/golden/bench_digests.json (generated at N = 1): sharded == unsharded, bit for bit
* ``cpu_baseline``  the oracle port (numpy/BLAS, all threads) on a bounded sample;
  ``cpu_baseline_reference``  the unmodified reference (oracle/_ref), single thread, on 100 k docs
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "queries/s"
WORKLOADS = {
    # name: (metric, defaults)
    "hybrid": ("hybrid_bm25 queries/sec @10M docs 384-d top-100",
               dict(n_docs=10_000_000, dim=384, batch=256, top_k=100, dense_mode="bf16_exact")),
    "bm25": ("bm25 queries/sec @50M docs top-100 (BASELINE config 3)",
             dict(n_docs=50_000_000, dim=384, batch=32, top_k=100, dense_mode="fp32")),
    "multi_stage": ("multi_stage stages 1-2 queries/sec @10M docs 768-d bf16, dense top-100 -> BM25 top-20 (BASELINE config 4)",
                    dict(n_docs=10_000_000, dim=768, batch=1024, top_k=100, dense_mode="bf16")),
    "diversity": ("diversity MMR queries/sec, top-1000 candidates x 384-d -> 250 picks (BASELINE config 5)",
                  dict(n_docs=10_000_000, dim=384, batch=4096, top_k=250, dense_mode="bf16")),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hybrid", choices=sorted(WORKLOADS))
    ap.add_argument("--n-docs", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="queries per step")
    ap.add_argument("--top-k", type=int, default=None)
    ap.add_argument("--dense-mode", default=None, choices=["exact", "fp32", "tf32x3", "bf16", "bf16_exact"])
    ap.add_argument("--cpu-sample-docs", type=int, default=500_000)
    ap.add_argument("--ref-docs", type=int, default=50_000, help="docs of the unmodified-reference CPU runs")
    ap.add_argument("--ref-workers", type=int, default=0, help="--impl reference: worker processes (0 = all cores, <= 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra measurement points / pipeline e2e")
    a = ap.parse_args()
    for k, v in WORKLOADS[a.workload][1].items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    a.warmup = max(a.warmup, 3)
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j.get("bf16_tflops_sustained", 1402.2), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    In-process NVML polling every 5 ms (nvidia-smi's loop mode cannot sample a 50 ms region reliably);
    only samples taken between mark_start() and mark_end() are reported.
    """
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.t0 = self.t1 = None
        self._stop = False
        self._thread = None
        self._err = None
        self._max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indices follow PCI order; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self._stop:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.rows.append((time.time(), sm, pw, rs))
                    except Exception as e:      # keep sampling; report the error once
                        self._err = repr(e)
                    time.sleep(0.005)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception as e:
            self._err = repr(e)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [f"nvml unavailable: {self._err}"]}
        inside = [r for r in self.rows if self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if not inside:
            inside = [r for r in self.rows if self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
            window = "timed region +-50 ms"
        reasons = set()
        for r in inside:
            for bit, name in self.REASONS.items():
                if r[3] & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": float(self._max), "power_w_max": max([r[2] for r in inside]) if inside else None,
                "samples": len(inside), "window": window, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- CPU arms
def _all_blas_threads():
    """numpy/BLAS on every host core even under torchrun (which exports OMP_NUM_THREADS=1)."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        got = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
        return int(got)
    except Exception:
        return n


def cpu_port_run(args, steps: int, warmup: int):
    """Oracle port of HybridBM25Pipeline.search (pipelines.py:315-357) on the host cores.

    Bounded sample: the first ``cpu_sample_docs`` docs of the SAME synthetic corpus (global vocabulary,
    statistics of the sample), a few queries per step; numpy/BLAS uses every host thread for the
    dense part.  Both hot loops are O(N), so queries/s at the full corpus is extrapolated linearly.
    """
    from hybrid_search_engine_b200 import synth
    from oracle import hybrid_oracle as orc
    threads = _all_blas_threads()
    n_s = min(args.cpu_sample_docs, args.n_docs)
    spec = synth.SynthSpec(n_docs=args.n_docs, vocab=args.vocab, dim=args.dim)
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    t0 = time.perf_counter()
    dl, terms = synth.doc_tokens(spec, 0, n_s, th)
    splits = np.cumsum(dl)[:-1]
    st = orc.bm25_fit_tokens(np.split(terms.astype(np.int64), splits), spec.vocab)
    emb = np.concatenate([synth.embeddings(spec, s, min(n_s, s + 50_000)) for s in range(0, n_s, 50_000)])
    build_s = time.perf_counter() - t0
    nq = max(1, min(args.batch, 4))
    qv = synth.query_embeddings(spec, 0, 1024)
    qt = synth.query_terms(spec, 0, 1024, th)

    def one_query(qi):
        cos = orc.batch_cosine_sim_port(qv[qi], emb)                         # utils.py:28-54 (BLAS, all threads)
        sem = orc.normalize_scores(cos)                                      # utils.py:57-71
        tids = [int(t) for t in qt[qi] if st.indptr[t + 1] > st.indptr[t]]
        bm = orc.bm25_scores_f64(st, tids).astype(np.float32)                # bm25.py:114-127
        fused = orc.hybrid_bm25_fused(sem, bm, 0.6, 0.4)                     # pipelines.py:331-340
        return np.argsort(-fused, kind="stable")[:args.top_k]                # pipelines.py:342-343

    for w in range(warmup):
        one_query(w % 1024)
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        for j in range(nq):
            one_query((s * nq + j) % 1024)
            done += 1
    dt = time.perf_counter() - t0
    qps_sample = done / dt
    return {
        "value": qps_sample * (n_s / args.n_docs), "unit": UNIT, "cores": int(threads), "kind": "port",
        "sample": (f"oracle port (numpy/BLAS) of hybrid_bm25 on the first {n_s} docs of the same synthetic corpus, "
                   f"{done} queries in {dt:.2f}s = {qps_sample:.2f} q/s on the sample, scaled by {n_s}/{args.n_docs} "
                   f"(both hot loops are O(N)); host has {os.cpu_count()} cpus; sample index build {build_s:.1f}s untimed"),
        "ms_per_step": 1e3 * dt / max(steps, 1), "queries_per_step": nq,
    }


class _RefWorld:
    """The unmodified reference's hybrid_bm25 pipeline (oracle/refload.py: /root/reference in the dev container, the
    staged copy under oracle/_ref elsewhere) indexed on the first ``n_ref`` docs of the bench corpus."""

    def __init__(self, args, n_ref):
        from hybrid_search_engine_b200 import synth
        from oracle import hybrid_oracle as orc
        from oracle import refload
        self.ref = refload.load()
        spec = synth.SynthSpec(n_docs=args.n_docs, vocab=args.vocab, dim=args.dim)
        th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
        t0 = time.perf_counter()
        docs = synth.doc_texts(spec, 0, n_ref, th)
        emb = np.concatenate([synth.embeddings(spec, s, min(n_ref, s + 50_000)) for s in range(0, n_ref, 50_000)])
        self.queries = synth.query_texts(spec, 0, 256, th)
        qv = synth.query_embeddings(spec, 0, 256)
        refload.EMBED_DIM[0] = args.dim
        refload.EMBED_TABLE.clear()
        for t, e in zip(map(orc.preprocess_text, docs), emb):
            refload.EMBED_TABLE[t] = e
        for t, e in zip(self.queries, qv):
            refload.EMBED_TABLE[t] = e
        refload.PARTIAL_RATIO_FN[0] = lambda a, b: 50.0        # multiplied by 0.0 on this path (pipelines.py:322-323)
        self.pipe = self.ref.pipelines.create_pipeline("hybrid_bm25")
        self.pipe.index(docs)                                  # Indexer + BM25.fit, unmodified
        self.build_s = time.perf_counter() - t0
        self.top_k = args.top_k
        self.pipe.search(self.queries[0], top_k=self.top_k)    # numba JIT + first-call costs outside the clock

    def run(self, qis):
        for qi in qis:
            self.pipe.search(self.queries[qi % len(self.queries)], top_k=self.top_k)
        return len(qis)


_REF_WORLD = None


def _ref_worker(qis):
    return _REF_WORLD.run(qis)


def cpu_reference_run(args, steps: int, warmup: int, workers: int):
    """Times ``create_pipeline("hybrid_bm25").search(query, top_k)`` of the UNMODIFIED reference (encoder and DuckDB
    replaced by inert stand-ins, which flatters it) on the first ``ref_docs`` docs.  The reference is single-threaded;
    with ``workers`` > 1, forked worker processes answer independent queries concurrently (queries are independent,
    so this is the reference's throughput on all the host cores).  One step = one query per worker."""
    global _REF_WORLD
    from oracle import refload
    if not refload.available():
        return None
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=1)          # workers are processes; BLAS inside each stays serial
    except Exception:
        pass
    n_ref = min(args.ref_docs, args.n_docs)
    _REF_WORLD = world = _RefWorld(args, n_ref)
    if workers <= 1:
        world.run(list(range(warmup)))
        t0 = time.perf_counter()
        done = world.run(list(range(warmup, warmup + steps)))
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            pool.map(_ref_worker, [[w] for w in range(workers)])                       # warm every worker
            chunks = [[(s * workers + w) for s in range(steps)] for w in range(workers)]
            t0 = time.perf_counter()
            done = sum(pool.map(_ref_worker, chunks, chunksize=1))
            dt = time.perf_counter() - t0
    qps_sample = done / dt
    return {
        "value": qps_sample * (n_ref / args.n_docs), "unit": UNIT, "cores": int(max(workers, 1)), "kind": "reference",
        "sample": (f"unmodified reference create_pipeline('hybrid_bm25').search(q, top_k={args.top_k}) on the first {n_ref} docs of "
                   f"the same synthetic corpus, {max(workers, 1)} worker process(es) x 1 thread, {done} queries in {dt:.2f}s = "
                   f"{qps_sample:.3f} q/s on the sample, scaled by {n_ref}/{args.n_docs} (the reference's loops are O(N)); "
                   f"encoder / DuckDB are inert stand-ins; index() {world.build_s:.1f}s untimed; host has {os.cpu_count()} cpus"),
        "ms_per_step": 1e3 * dt / max(steps, 1), "queries_per_step": max(workers, 1),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = args.ref_workers if args.ref_workers > 0 else min(os.cpu_count() or 1, 32)
    cb = None
    if args.workload == "hybrid":
        try:
            cb = cpu_reference_run(args, args.steps, min(args.warmup, 2), workers)
        except Exception as e:      # staged copy missing / unusable: fall back to the port and say so
            print(f"bench.py: unmodified reference unavailable ({e!r}); timing the oracle port", file=sys.stderr)
    if cb is None:
        cb = cpu_port_run(args, args.steps, args.warmup)
    metric = WORKLOADS["hybrid"][0]
    line = {
        "impl": "reference", "metric": metric, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"hybrid_bm25 top-{args.top_k}, {args.n_docs} Zipfian docs x {args.dim}-d fp32, "
                               f"vocab {args.vocab} (CPU: bounded sample, extrapolated linearly in N)",
                   "queries_per_step": cb["queries_per_step"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- ours
class _Marks(list):
    """Phase-boundary CUDA events of one sub-batch (engine.phase_events) + named marks inside the last phase."""

    def __init__(self):
        super().__init__()
        self.extra = []


def digest_of(ids, sc):
    return hashlib.sha256(np.ascontiguousarray(ids, dtype=np.int64).tobytes()
                          + np.ascontiguousarray(sc, dtype=np.float32).tobytes()).hexdigest()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hybrid_search_engine_b200 import _lib, synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    hbm_peak, bf16_peak, peak_src = load_peaks()
    wl = args.workload
    metric = WORKLOADS[wl][0]

    spec = synth.SynthSpec(n_docs=args.n_docs, vocab=args.vocab, dim=args.dim)
    per = (args.n_docs + world - 1) // world
    lo, hi = min(args.n_docs, rank * per), min(args.n_docs, (rank + 1) * per)
    n_shard = hi - lo
    t0 = time.perf_counter()
    shard = synth_device.build_synthetic_shard(spec, lo, hi, device, group=group, dense=wl != "bm25",
                                               lexical=wl != "diversity")
    if args.dense_mode in ("bf16", "bf16_exact") and wl != "bm25":
        shard.ensure_bf16()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    B, k = args.batch, args.top_k
    tensor_mode = args.dense_mode in ("tf32x3", "bf16", "bf16_exact")
    sub = min(B, 128 if args.dense_mode == "tf32x3" else (256 if args.dense_mode in ("bf16", "bf16_exact") else 8))
    vflags = torch.zeros(max(B, 1), dtype=torch.int32, device=device)      # bf16_exact: queries the verification flagged
    if wl == "bm25":
        sub = min(B, 32)
    eng = SearchEngine(shard, group=group, max_batch=sub, dense_mode=args.dense_mode)
    n_q = 4096 if wl == "diversity" else 1024
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    qv_all = synth.query_embeddings(spec, 0, n_q)
    qt_all = synth.query_terms(spec, 0, n_q, th).tolist()
    df_host = shard.df_host if shard.has_bm25 else None

    def batch_of(step, nb=B):
        idx = [(step * nb + j) % n_q for j in range(nb)]
        return QueryBatch(vectors=qv_all[idx], term_ids=[qt_all[i] for i in idx])

    def barrier():
        if world > 1:
            dist.barrier(group=group, device_ids=[local])
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---------------------------------------------------------------- the step of each workload, device-resident inputs
    def stage(qb):
        """Upload one batch and keep private device copies: (vectors, [(terms, idf, offsets, n_tokens) per sub-batch])."""
        qd = eng.upload_vectors(qb.vectors).clone() if wl != "bm25" else None
        terms = [tuple(t.clone() if hasattr(t, "clone") else t for t in x)
                 for x in eng.upload_terms_split(qb.term_ids, sub)] if wl != "diversity" else None
        return qd, terms, qb

    def hybrid_step(staged, timers=None):
        qd, terms, _ = staged
        out = None
        for bi, s in enumerate(range(0, B, sub)):
            nb = min(B, s + sub) - s
            stats = eng._stats(nb)
            if args.dense_mode == "bf16_exact":      # screen (bf16 GEMM) + exact verification: one fused chain
                qt, qi, qo, n_tok = terms[bi]
                eng.phase_events = _Marks() if timers is not None else None
                keys = eng._verified_sub_batch(qd[s:s + nb], nb, stats, ("bm25", qt, qi, qo, n_tok), 2, 0.6, 0.4, k,
                                               vflags[s:s + nb], _lib.VERIFY_EPS["bf16_exact"])
                out = eng.unpack(keys)
                if timers is not None:               # [start, after GEMM, after BM25, after verify + select + merge]
                    timers.append(eng.phase_events)
                    eng.phase_events = None
                continue
            if timers is not None:
                e = [ev() for _ in range(4)]
                e[0].record()
            cos = eng.dense_scan(qd[s:s + nb], stats)
            if timers is not None:
                e[1].record()
            qt, qi, qo, n_tok = terms[bi]
            bm = eng.bm25_score(qt, qi, qo, nb, stats, n_tok)
            if timers is not None:
                e[2].record()
            stats = eng._exchange_stats(stats, nb)
            keys = eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, k)
            out = eng.unpack(keys)
            if timers is not None:
                e[3].record()
                timers.append(e)
        return out

    def bm25_step(staged, timers=None):
        _, terms, _ = staged
        out = None
        for bi, s in enumerate(range(0, B, sub)):
            nb = min(B, s + sub) - s
            if timers is not None:
                e = [ev() for _ in range(4)]
                e[0].record()
                e[1].record()
            qt, qi, qo, n_tok = terms[bi]
            bm = eng.bm25_score(qt, qi, qo, nb, None, n_tok)
            if timers is not None:
                e[2].record()
            out = eng.unpack(eng.fuse_topk(0, bm, None, None, 1.0, 0.0, k))
            if timers is not None:
                e[3].record()
                timers.append(e)
        return out

    def multi_stage_step(staged, timers=None):
        qd, terms, qb = staged
        sc1, ids1 = eng.search_semantic(QueryBatch(vectors=qb.vectors), k, 1.0)       # stage 1 (filter epilogue)
        bm = eng.bm25_score_docs_global(qb.term_ids, ids1)                            # stage 2: float64 BM25.score
        order = torch.sort(-bm, dim=1, stable=True).indices[:, :20]                   # stable: ties keep stage-1 rank
        return torch.gather(bm, 1, order), torch.gather(ids1, 1, order)

    def diversity_step(staged, timers=None):
        qd, _, qb = staged
        sc1, ids1 = eng.search_semantic(QueryBatch(vectors=qb.vectors), 4 * k, 0.7)   # candidates (semantic part)
        sc64 = sc1.to(torch.float64)
        mn, mx = sc64.min(1, keepdim=True).values, sc64.max(1, keepdim=True).values
        rel = (sc64 - mn) / (mx - mn + 1e-8)                                          # pipelines.py:589
        if world == 1:
            return eng.mmr(ids1 - shard.doc_base, rel, 0.5, k), ids1
        sel = eng.mmr_sharded(ids1.cpu().numpy(), rel.cpu().numpy(), 0.5, k)
        return torch.from_numpy(sel), ids1

    step_fn = {"hybrid": hybrid_step, "bm25": bm25_step, "multi_stage": multi_stage_step,
               "diversity": diversity_step}[wl]

    n_staged = args.warmup + args.steps
    distinct = min(n_staged, max(1, n_q // B)) if B <= n_q else 1
    staged_pool = [stage(batch_of(s)) for s in range(distinct)]
    staged = [staged_pool[s % distinct] for s in range(n_staged)]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        step_fn(staged[s])
    barrier()
    sampler.mark_start()
    launches0 = eng.launches
    timers = []
    ev0, ev1 = ev(), ev()
    ev0.record()
    for s in range(args.steps):
        res = step_fn(staged[args.warmup + s], timers=timers)
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    parts = [0.0, 0.0, 0.0]
    chain_ms = {}
    if timers:
        per_step = len(timers) / args.steps
        for e in timers:
            for j in range(3):
                parts[j] += e[j].elapsed_time(e[j + 1])
            prev = e[2]
            for name, evx in list(getattr(e, "extra", [])) + [("exchange_keys", e[3])]:
                chain_ms[name] = chain_ms.get(name, 0.0) + prev.elapsed_time(evx) / args.steps
                prev = evx
        parts = [p / args.steps for p in parts]                 # ms per STEP for dense / bm25 / select chain
    t = torch.tensor([dev_ms] + parts, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dev_ms, dense_ms, bm25_ms, select_ms = [float(x) for x in t]
    step_ms = dev_ms / args.steps
    last = [x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x) for x in res]

    # ---------------------------------------------------------------- end to end with host inputs / host outputs
    host_batches = [batch_of(args.warmup + s) for s in range(args.steps)]
    warm_batches = [batch_of(s) for s in range(3)]
    depth = 2      # two batches in flight: the host prepares / uploads batch i+1 while batch i runs
    e2e_note = None
    if wl == "hybrid":
        for _ in eng.search_hybrid_bm25_stream(warm_batches, k, 0.6, 0.4, depth=depth):
            pass
        barrier()
        t0 = time.perf_counter()
        for res_sc, res_ids in eng.search_hybrid_bm25_stream(host_batches, k, 0.6, 0.4, depth=depth):
            pass                                        # device -> host read of every step's result
        barrier()
        e2e_s = time.perf_counter() - t0
        same = bool(np.array_equal(res_ids, last[1]))
        h2d = B * args.dim * 4 + sum(len(x) for x in batch_of(0).term_ids) * 12 + (B // sub + B + 1) * 4
        d2h = B * k * (4 + 8)
    else:
        def host_step(qb):
            st_ = stage(qb)
            out = step_fn(st_)
            return [x.cpu() if hasattr(x, "cpu") else x for x in out]
        for qb in warm_batches[:2]:
            host_step(qb)
        barrier()
        t0 = time.perf_counter()
        for qb in host_batches:
            host_out = host_step(qb)
        barrier()
        e2e_s = time.perf_counter() - t0
        same = bool(np.array_equal(np.asarray(host_out[1]), last[1]))
        depth = 1
        h2d = (B * args.dim * 4 if wl != "bm25" else 0) + (sum(len(x) for x in batch_of(0).term_ids) * 12 if wl != "diversity" else 0)
        d2h = int(sum(np.asarray(x).nbytes for x in host_out))
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX, group=group)
    e2e_s = float(te[0])

    # ---------------------------------------------------------------- plugin-API end to end (hybrid, query strings)
    e2e_pipe = None
    if wl == "hybrid" and not args.no_extras:
        import hybrid_search_engine_b200 as hs
        pipe = hs.create_pipeline("hybrid_bm25", device=str(device), dense_mode=args.dense_mode)
        pipe.attach_index(shard, synth.LazyDocs(spec), query_term_ids=synth.SynthVocab(spec).query_term_ids, group=group)
        pipe.searcher.engine.max_batch = sub
        q_txt = synth.query_texts(spec, 0, n_q, th)
        txt_batches = [([q_txt[(s * B + j) % n_q] for j in range(B)], batch_of(s).vectors)
                       for s in range(args.warmup, args.warmup + args.steps)]
        for qs, qv in txt_batches[:2]:
            pipe.search_many(qs, top_k=k, query_vectors=qv)
        barrier()
        t0 = time.perf_counter()
        for qs, qv in txt_batches:
            pres = pipe.search_many(qs, top_k=k, query_vectors=qv)
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX, group=group)
        p_ids = np.array([[r["doc_id"] for r in x.results] for x in pres])
        e2e_pipe = {"value": args.steps * B / float(tp[0]), "unit": UNIT,
                    "api": "create_pipeline('hybrid_bm25').search_many(query strings, top_k, query_vectors=...)",
                    "includes": "tokenisation + term lookup + H2D + kernels + D2H + result dictionaries (contents are "
                                "placeholders: the synthetic corpus exists only as token ids in HBM)",
                    "same_ids_as_device_run": bool(np.array_equal(p_ids, last[1]))}

    # ---------------------------------------------------------------- parity
    parity = {}
    if wl == "hybrid":
        qb_last = batch_of(args.warmup + args.steps - 1)
        _, ex_ids = eng.search_hybrid_bm25(qb_last, k, 0.6, 0.4, dense_mode="exact")
        ex_ids = ex_ids.cpu().numpy()
        parity["topk_ids_equal_to_exact_mode"] = float(np.mean(ex_ids == last[1]))
        parity["queries_with_identical_topk"] = float(np.mean(np.all(ex_ids == last[1], axis=1)))
        if args.dense_mode == "bf16":
            parity["recall_at_k_vs_exact"] = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(last[1], ex_ids)]))
        # fixed 64-query batch, exact mode: the digest is a constant of the corpus -> equal at every rank count
        qb64 = QueryBatch(vectors=qv_all[:64], term_ids=qt_all[:64])
        old_mb = eng.max_batch
        eng.max_batch = 8
        d_sc, d_ids = eng.search_hybrid_bm25(qb64, k, 0.6, 0.4, dense_mode="exact")
        eng.max_batch = old_mb
        dg = digest_of(d_ids.cpu().numpy(), d_sc.cpu().numpy())
        key = f"hybrid_n{args.n_docs}_v{args.vocab}_d{args.dim}_k{k}_q64"
        try:
            want = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_digests.json"))).get(key)
        except Exception:
            want = None
        parity.update({"sharded_digest": dg, "sharded_digest_key": key, "sharded_digest_committed": want,
                       "sharded_digest_equal": (dg == want) if want else None,
                       "note": "exact mode == CPU oracle bit for bit (tests/test_gpu_parity.py); digest = sha256(ids, scores) "
                               "of the exact-mode top-100 of queries 0..63, committed constant generated at N=1"})
    elif wl == "multi_stage" and args.dense_mode == "bf16":
        qb_last = batch_of(args.warmup + args.steps - 1, min(B, 64))
        _, a = eng.search_semantic(QueryBatch(vectors=qb_last.vectors), k, 1.0, dense_mode="bf16")
        a = a.cpu().numpy().copy()
        _, b_ = eng.search_semantic(QueryBatch(vectors=qb_last.vectors), k, 1.0, dense_mode="exact")
        b_ = b_.cpu().numpy()
        parity["stage1_recall_at_k_vs_exact"] = float(np.mean([len(set(x) & set(y)) / k for x, y in zip(a, b_)]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- rooflines (rank 0; per-shard figures)
    kernels = []
    passes = eng.dense_launches(sub) * ((B + sub - 1) // sub) if wl in ("hybrid",) else 0
    split = True
    if wl == "hybrid":
        elem = 4
        if args.dense_mode in ("bf16", "bf16_exact"):
            elem = 2
        # + the score matrix the tensor-core modes write: float32, or binary16 for the screen of bf16_exact
        f16_screen = args.dense_mode == "bf16_exact" and eng.screen_f16
        dense_bytes = n_shard * shard.ld * elem * passes + (B * n_shard * (2 if f16_screen else 4) if tensor_mode else 0)
        gname = {"bf16_exact": "bf16 STORE + extreme lists"}.get(args.dense_mode, args.dense_mode)
        ent = {"name": "dense_gemm_kernel (tcgen05 %s)" % gname if tensor_mode else "dense_scan_kernel",
               "ms_per_step": dense_ms, "launches_per_step": passes, "alg_bytes_per_step": dense_bytes,
               "hbm_GBps": dense_bytes / dense_ms / 1e6, "frac_hbm": dense_bytes / dense_ms / 1e6 / hbm_peak}
        if tensor_mode:
            mul = 3 if args.dense_mode == "tf32x3" else 1
            peak_t = bf16_peak / (2 if args.dense_mode == "tf32x3" else 1)
            ent["note"] = "bf16 MMA against the measured sustained bf16 peak; the kernel is HBM-bound at <= 256 queries per pass"
            fl = 2.0 * B * n_shard * shard.ld * mul
            ent.update({"mma_tflops": fl / dense_ms / 1e9, "tensor_peak_tflops": peak_t, "frac_tensor": fl / dense_ms / 1e9 / peak_t,
                        "algorithmic_tflops": 2.0 * B * n_shard * args.dim / dense_ms / 1e9,
                        "note": "tf32x3 issues 3 TF32 MMAs per K step (hi*hi + lo*hi + hi*lo); peak = measured bf16 sustained / 2"
                        if mul == 3 else "bf16 MMA against the measured sustained bf16 peak"})
        kernels.append(ent)
    if (wl == "hybrid" and split) or wl == "bm25":
        # bytes_bm25(query) = 8 * P(q) + 8 * N_shard, P from the GLOBAL df scaled to the shard
        P = 0.0
        for s in range(args.warmup, args.warmup + args.steps):
            for q in staged[s][2].term_ids:
                P += sum(int(df_host[t_]) for t_ in q if 0 <= t_ < len(df_host))
        P = P / args.steps * (n_shard / args.n_docs)
        bm_bytes = 8 * P + 8 * n_shard * B
        kernels.append({"name": "bm25_ranges_kernel + bm25_batch_kernel", "ms_per_step": bm25_ms, "alg_bytes_per_step": bm_bytes,
                        "postings_per_step": P, "hbm_GBps": bm_bytes / bm25_ms / 1e6, "frac_hbm": bm_bytes / bm25_ms / 1e6 / hbm_peak})
        sel_bytes = (8 if wl == "hybrid" else 4) * n_shard * B
        if wl == "hybrid" and args.dense_mode == "bf16_exact" and eng.screen_f16:
            sel_bytes = ((4 if eng.screen_bm25_f16 else 6)) * n_shard * B      # binary16 screen + float32 (or binary16) BM25
        sname = "fuse_blockmax + fuse_bound + fuse_topk + topk_merge + keys_unpack"
        if args.dense_mode == "bf16_exact" and wl == "hybrid":
            sname = "verify_stats + fuse_blockmax + fuse_bound + fuse_topk (k' = 256) + topk_merge + verify_topk + keys_unpack"
        kernels.append({"name": sname + (" (+ C2/C1 exchange)" if world > 1 else ""),
                        "ms_per_step": select_ms, "alg_bytes_per_step": sel_bytes, "hbm_GBps": sel_bytes / select_ms / 1e6,
                        "frac_hbm": sel_bytes / select_ms / 1e6 / hbm_peak})
        if chain_ms:        # this rank's split of the chain (CUDA events between its parts)
            kernels[-1]["parts_ms_rank0"] = {k_: round(v_, 4) for k_, v_ in chain_ms.items()}
    if wl == "hybrid" and split:
        # bytes_hybrid(B) = N d 4 + B (8 P + 8 N) + B 8 N   (one corpus pass whatever B)
        step_bytes = n_shard * shard.ld * 4 + kernels[1]["alg_bytes_per_step"] + 8 * n_shard * B
        step_ent = {"name": "step (bytes_hybrid of SURVEY 8(d))", "ms_per_step": step_ms, "alg_bytes_per_step": step_bytes,
                    "hbm_GBps": step_bytes / step_ms / 1e6, "frac_hbm": step_bytes / step_ms / 1e6 / hbm_peak}
        dom = max(kernels, key=lambda e: e["ms_per_step"])
        if "frac_tensor" in dom and dom["frac_tensor"] > dom["frac_hbm"]:
            roof = {"bound": "tensor", "kernel": dom["name"], "achieved": dom["mma_tflops"], "peak": dom["tensor_peak_tflops"],
                    "unit": "TFLOP/s", "frac": dom["frac_tensor"]}
        else:
            roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom["frac_hbm"]}
        roof.update({"launch_ms": dom["ms_per_step"] / max(dom.get("launches_per_step", 1), 1),
                     "share_of_step": dom["ms_per_step"] / step_ms})
        kernels.append(step_ent)
    elif wl == "hybrid":
        pass
    elif wl == "bm25":
        dom = kernels[0]
        roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dom["frac_hbm"], "share_of_step": dom["ms_per_step"] / step_ms}
    elif wl == "multi_stage":
        fl = 2.0 * B * n_shard * shard.ld_bf16 if args.dense_mode == "bf16" else 2.0 * B * n_shard * shard.ld
        roof = {"bound": "tensor", "kernel": "dense_gemm_kernel (filter epilogue) + candidate select + bm25_docs (whole step)",
                "achieved": fl / step_ms / 1e9, "peak": bf16_peak, "unit": "TFLOP/s", "frac": fl / step_ms / 1e9 / bf16_peak,
                "roofline_ms": fl / bf16_peak / 1e9,
                "hbm_frac_of_step": (n_shard * (shard.ld_bf16 if args.dense_mode == "bf16" else shard.ld) * 2 * ((B + 255) // 256)) / step_ms / 1e6 / hbm_peak}
    else:
        roof = {"bound": "latency", "kernel": "dense_gemm_kernel (filter) + mmr_kernel", "achieved": None, "peak": None, "unit": "us/query",
                "frac": None, "us_per_query": step_ms * 1e3 / B}
    roof.update({"peak_source": peak_src, "traffic": None, "kernels": kernels})
    try:    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, when one matches
        ent = json.load(open(os.path.join(ROOT, "profiles", "dense_traffic.json")))["entries"]
        if "bm25_batch" in str(roof.get("kernel", "")):      # the dominant kernel is BM25: its own capture
            roof["traffic"] = ent.get(f"bm25_batch_n{n_shard}_q{B}")
        else:
            roof["traffic"] = ent.get(f"{args.dense_mode}_n{n_shard}_ld{shard.ld}_q{sub}") if tensor_mode else \
                ent.get(f"n{n_shard}_ld{shard.ld}_q{min(sub, 8)}")
    except Exception:
        pass

    dtype = {"fp32": "f32", "exact": "f32 (f64-accumulated dot, f64 BM25)", "tf32x3": "f32 (3xTF32 tensor-core dot, f64 BM25)",
             "bf16": "bf16 (dense) + f64 BM25",
             "bf16_exact": "bf16 screen + f64-accumulated exact verification (results of the exact mode), f64 BM25"}[args.dense_mode]
    if args.dense_mode == "bf16_exact":
        parity["verify_flagged_queries_in_timed_steps"] = int((vflags != 0).sum().item())
    workload = {
        "hybrid": f"hybrid_bm25 (0.6/0.4, k1=1.5, b=0.75) top-{k} over {args.n_docs} Zipfian docs x {args.dim}-d fp32, vocab {args.vocab}, avg 200 tokens/doc",
        "bm25": f"bm25 top-{k} over a {args.n_docs}-doc Zipfian inverted index (avg 200 terms/doc, vocab {args.vocab})",
        "multi_stage": f"multi_stage stages 1-2: dense top-{k} -> BM25.score on the {k} -> top-20, {args.n_docs} docs x {args.dim}-d {args.dense_mode}",
        "diversity": f"diversity: semantic top-{4 * k} candidates x {args.dim}-d ({args.dense_mode}) -> MMR lambda=0.5 -> {k} picks, {args.n_docs} docs "
                     "(fuzzy-lexical term of the candidate search excluded at this scale: rapidfuzz parity unpinned)",
    }[wl]
    line = {
        "metric": metric, "value": args.steps * B / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload, "queries_per_step": B, "queries_per_corpus_pass": sub, "dense_mode": args.dense_mode,
                   "parallelism": f"doc-shard x{world}",
                   "l2": "inputs larger than L2 (every step streams the whole shard: >= 1.9 GB of embeddings + postings per GPU)",
                   "index_build_s": round(build_s, 1)},
        "e2e": {"value": args.steps * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "batches_in_flight": depth,
                "d2h_bytes_per_step": int(d2h), "same_ids_as_device_run": same},
        "e2e_pipeline": e2e_pipe,
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
    }

    # ---------------------------------------------------------------- extra points of the north star (N = 1, hybrid)
    if wl == "hybrid" and world == 1 and not args.no_extras:
        pts = []
        # SURVEY 8(d) north-star points B in {1, 32, 256} (+ 8, the largest batch of ONE pass of the CUDA-core fp32 scan)
        for pb, mode in ((1, "fp32"), (8, "fp32"), (32, "bf16_exact"), (128, "bf16_exact"), (128, "tf32x3"), (256, "bf16_exact")):
            if pb == B and mode == args.dense_mode:
                continue
            e2 = SearchEngine(shard, max_batch=min(pb, 128 if mode == "tf32x3" else 256), dense_mode=mode)
            st_ = []
            for s in range(8):
                qb = batch_of(s, pb)
                qd = e2.upload_vectors(qb.vectors).clone()
                qt, qi, qo = [x.clone() for x in e2.upload_terms(qb.term_ids)]
                st_.append((qd, qt, qi, qo, e2._n_tokens, qb))

            def small_step(x, tm=None):
                if mode == "bf16_exact":        # the verified chain through the engine call (host inputs)
                    return e2.search_hybrid_bm25(x[5], k, 0.6, 0.4)
                stats = e2._stats(pb)
                if tm is not None:
                    a, b_ = ev(), ev()
                    a.record()
                cos = e2.dense_scan(x[0], stats)
                if tm is not None:
                    b_.record()
                    tm.append((a, b_))
                bm = e2.bm25_score(x[1], x[2], x[3], pb, stats, x[4])
                return e2.unpack(e2.fuse_topk(2, cos, bm, stats, 0.6, 0.4, k))
            for s in range(4):
                small_step(st_[s % 8])
            torch.cuda.synchronize()
            tm = []
            a, b_ = ev(), ev()
            a.record()
            for s in range(16):
                small_step(st_[s % 8], tm)
            b_.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b_) / 16
            pt = {"queries_per_step": pb, "dense_mode": mode, "value": pb / ms * 1e3, "unit": UNIT, "ms_per_step": ms,
                  "note": "device-resident inputs" if mode != "bf16_exact" else "through SearchEngine.search_hybrid_bm25 (host inputs)"}
            if tm:
                dms = float(np.mean([x.elapsed_time(y) for x, y in tm])) / e2.dense_launches(pb)
                pt.update({"dense_launch_ms": dms, "dense_frac_hbm": n_shard * shard.ld * 4 / dms / 1e6 / hbm_peak})
                if mode == "tf32x3":
                    pt["dense_frac_tensor"] = 6.0 * pb * n_shard * shard.ld / dms / 1e9 / (bf16_peak / 2)
            pts.append(pt)
            del e2
        line["points"] = pts
    if world == 1 and not args.no_cpu_baseline and wl == "hybrid":
        cb = cpu_port_run(args, steps=8, warmup=1)
        line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_extras:
            try:
                cr = cpu_reference_run(args, steps=5, warmup=1, workers=1)
                line["cpu_baseline_reference"] = None if cr is None else {kk: cr[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["cpu_baseline_reference"] = {"unavailable": repr(e)}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
