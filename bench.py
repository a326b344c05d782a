#!/usr/bin/env python
"""bench.py -- hybrid_bm25 queries/sec on a synthetic 10 M-doc x 384-d corpus (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # CPU port of the reference path on the host cores

One "step" = one batch of B queries through the whole hot path: dense cosine scan over every doc,
BM25 over the CSR index, min-max / max normalisation + weighted fusion + top-100 select (+ the
stats all-reduce and top-k all-gather/merge when doc-sharded over N GPUs).  The corpus (10 M docs)
is FIXED as N grows (doc-sharded) => "scaling": "strong".

* ``value``  queries/s with the query batch already resident in HBM (device-timed with CUDA events,
             barrier + synchronize on both sides, max over ranks)
* ``e2e``    queries/s through the public batched API with HOST inputs: pinned host -> device copy of
             the query vectors / term ids and device -> host read of the result inside every step,
             through the serving loop ``SearchEngine.search_hybrid_bm25_stream`` (batch i+1 is uploaded
             while batch i runs; every step's top-k is read back)
* ``roofline``  the dense scan kernel: algorithmic bytes per launch (n_shard * ld * 4) over
             its mean launch duration measured with CUDA events inside the timed region
* ``cpu_baseline``  the oracle port of the reference path timed on the host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hybrid_bm25 queries/sec @10M docs 384-d top-100"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--batch", type=int, default=8, help="queries per step")
    ap.add_argument("--top-k", type=int, default=100)
    ap.add_argument("--dense-mode", default="fp32", choices=["exact", "fp32"])
    ap.add_argument("--cpu-sample-docs", type=int, default=500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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


# --------------------------------------------------------------------------------------------- CPU port
def cpu_reference_run(args, steps: int, warmup: int):
    """Oracle port of HybridBM25Pipeline.search (pipelines.py:315-357) on the host cores.

    Bounded sample: the first ``cpu_sample_docs`` docs of the SAME synthetic corpus (global vocabulary,
    statistics of the sample), ``batch`` queries per step; numpy/BLAS uses every host thread for the
    dense part.  Both hot loops are O(N), so queries/s at the full corpus is extrapolated linearly.
    """
    from hybrid_search_engine_b200 import synth
    from oracle import hybrid_oracle as orc
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
        order = np.argsort(-fused, kind="stable")[:args.top_k]               # pipelines.py:342-343
        return order

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
    qps_full = qps_sample * (n_s / args.n_docs)
    try:
        import threadpoolctl
        threads = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    return {
        "value": qps_full, "unit": UNIT, "cores": int(threads), "kind": "port",
        "sample": (f"oracle port (numpy/BLAS) of hybrid_bm25 on the first {n_s} docs of the same synthetic corpus, "
                   f"{done} queries in {dt:.2f}s = {qps_sample:.2f} q/s on the sample, scaled by {n_s}/{args.n_docs} "
                   f"(both hot loops are O(N)); host has {os.cpu_count()} cpus; sample index build {build_s:.1f}s untimed"),
        "ms_per_step": 1e3 * dt / max(steps, 1), "queries_per_step": nq,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"hybrid_bm25 top-{args.top_k}, {args.n_docs} Zipfian docs x {args.dim}-d fp32, "
                               f"vocab {args.vocab} (CPU: bounded sample, extrapolated)",
                   "queries_per_step": cb["queries_per_step"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hybrid_search_engine_b200 import synth, synth_device
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

    spec = synth.SynthSpec(n_docs=args.n_docs, vocab=args.vocab, dim=args.dim)
    per = (args.n_docs + world - 1) // world
    lo, hi = min(args.n_docs, rank * per), min(args.n_docs, (rank + 1) * per)
    t0 = time.perf_counter()
    shard = synth_device.build_synthetic_shard(spec, lo, hi, device, group=group)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    eng = SearchEngine(shard, group=group, max_batch=args.batch, dense_mode=args.dense_mode)

    B, k = args.batch, args.top_k
    n_q = 1024
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    qv_all = synth.query_embeddings(spec, 0, n_q)
    qt_all = synth.query_terms(spec, 0, n_q, th).tolist()

    def batch_of(step):
        idx = [(step * B + j) % n_q for j in range(B)]
        return QueryBatch(vectors=qv_all[idx], term_ids=[qt_all[i] for i in idx])

    def barrier():
        if world > 1:
            dist.barrier(group=group, device_ids=[local])
        torch.cuda.synchronize()

    # ---- device-resident timing: upload once per step OUTSIDE the timed events
    def device_step(qd, qt, qi, qo, n_tok, timers=None):
        stats = eng._stats(B)
        if timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        cos = eng.dense_scan(qd, stats)
        if timers is not None:
            e1.record()
            timers.append((e0, e1))
        bm = eng.bm25_score(qt, qi, qo, B, stats, n_tok)
        stats = eng._exchange_stats(stats, B)
        keys = eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, k)
        return eng.unpack(keys)

    # host-side query batches for the end-to-end loop are prepared up front: generating them between the two timed
    # loops would leave the GPU idle for ~0.5 s and the second loop would start on ramping clocks
    host_batches = [batch_of(args.warmup + s) for s in range(args.steps)]
    warm_batches = [batch_of(s) for s in range(max(args.warmup, 3))]
    staged = []
    for s in range(args.warmup + args.steps):
        qb = batch_of(s)
        qd = eng.upload_vectors(qb.vectors).clone()
        qt, qi, qo = [t.clone() for t in eng.upload_terms(qb.term_ids)]
        staged.append((qd, qt, qi, qo, eng._n_tokens))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        device_step(*staged[s])
    barrier()
    sampler.mark_start()
    launches0 = eng.launches
    timers = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        sc, ids = device_step(*staged[args.warmup + s], timers=timers)
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    dense_ms = float(np.mean([a.elapsed_time(b) for a, b in timers]))
    dense_launches_per_step = eng.dense_launches(B)
    t = torch.tensor([dev_ms, dense_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dev_ms, dense_ms = float(t[0]), float(t[1])
    last_ids = ids.cpu().numpy()

    # ---- end to end through the public batched API with host inputs / host outputs: the serving loop
    # (SearchEngine.search_hybrid_bm25_stream) uploads batch i+1 from pinned memory while batch i runs and reads
    # every step's result back to the host; the query batches themselves are prepared before the clock starts
    # batches in flight: 2 overlaps the upload / launch work of batch i+1 with batch i on the GPU (+5..12 % end to
    # end at 1-4 ranks).  At 8 ranks (0.55 ms steps, two NCCL collectives per step) the only measurements so far are
    # 12.3 k q/s with 1 in flight and 8.8 k q/s with 2 -- the latter taken before the warm-up below existed, i.e.
    # with the pinned allocations of the second buffer set inside the timed region -- so 8 ranks keep 1 for now.
    depth = 2 if world <= 4 else 1
    # warm-up through the same loop: allocates both sets of pinned staging buffers before the clock starts
    for _ in eng.search_hybrid_bm25_stream(warm_batches, k, 0.6, 0.4, depth=depth):
        pass
    barrier()
    t0 = time.perf_counter()
    for res_sc, res_ids in eng.search_hybrid_bm25_stream(host_batches, k, 0.6, 0.4, depth=depth):
        pass                                        # device -> host read of every step's result
    res_ids = torch.from_numpy(res_ids)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX, group=group)
    e2e_s = float(te[0])
    same = bool(np.array_equal(res_ids.numpy(), last_ids))
    # conformance check outside the timed regions: the same batch in the float64-accumulated `exact` dense
    # mode (bit-identical to the CPU oracle in the parity tests) must give the same top-k ids
    a, b_ = eng.search_hybrid_bm25(batch_of(args.warmup + args.steps - 1), k, 0.6, 0.4, dense_mode="exact")
    exact_ids = b_.cpu().numpy()
    ids_equal_exact = float(np.mean(exact_ids == res_ids.numpy()))
    h2d = B * args.dim * 4 + sum(len(x) for x in batch_of(0).term_ids) * 12 + (B + 1) * 4
    d2h = B * k * (4 + 8)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        n_shard = hi - lo
        alg_bytes = n_shard * shard.ld * 4
        traffic = None      # DRAM bytes per launch from the committed ncu --set full capture, if one matches
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", "dense_traffic.json")))["entries"]
            traffic = ent.get(f"n{n_shard}_ld{shard.ld}_q{min(B, 8)}")
        except Exception:
            pass
        achieved = alg_bytes / (dense_ms / dense_launches_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": args.steps * B / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.dense_mode == "fp32" else "f32 (f64-accumulated dot, f64 BM25)",
            "data": "synthetic",
            "config": {"workload": f"hybrid_bm25 (0.6/0.4, k1=1.5, b=0.75) top-{k} over {args.n_docs} Zipfian docs x "
                                   f"{args.dim}-d fp32, vocab {args.vocab}, avg 200 tokens/doc",
                       "queries_per_step": B, "dense_mode": args.dense_mode, "parallelism": f"doc-shard x{world}",
                       "l2": "inputs larger than L2 (corpus pass 15.4 GB/step per shard set)",
                       "index_build_s": round(build_s, 1)},
            "e2e": {"value": args.steps * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "batches_in_flight": depth,
                    "d2h_bytes_per_step": d2h, "same_ids_as_device_run": same},
            "parity": {"topk_ids_equal_to_exact_mode": ids_equal_exact,
                       "note": "exact mode == CPU oracle bit for bit (tests/test_gpu_parity.py)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "dense_scan_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": alg_bytes, "launch_ms": dense_ms / dense_launches_per_step,
                         "launches_per_step": dense_launches_per_step,
                         "dense_share_of_step": dense_ms / (dev_ms / args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(args, steps=8, warmup=1)
            line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
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
