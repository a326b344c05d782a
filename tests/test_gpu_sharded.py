"""Doc-sharded execution on the GPU: G shards must give the 1-shard result bit for bit.

* ``test_two_shards_on_one_gpu``: both shards live on cuda:0, the exchange steps are done by hand
  (stats max/min, key lists concatenated) and merged with the CUDA merge kernel -- runs on a 1-GPU box.
* ``test_nccl_two_ranks``: real 2-process NCCL run through ``SearchEngine(group=...)`` -- needs 2 GPUs.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _spec():
    from hybrid_search_engine_b200 import synth
    return synth.SynthSpec(n_docs=50_001, vocab=3000, dim=384, min_len=20, max_len=60)


def test_two_shards_on_one_gpu():
    from hybrid_search_engine_b200 import _lib, parallel, synth, synth_device
    from hybrid_search_engine_b200._lib import check, ptr, stream_ptr
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = _spec()
    dev = torch.device("cuda:0")
    full = SearchEngine(synth_device.build_synthetic_shard(spec, 0, spec.n_docs, dev), max_batch=8)
    B, k = 6, 100
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B).tolist())
    ref_sc, ref_ids = [t.cpu().numpy().copy() for t in full.search_hybrid_bm25(qb, k, 0.6, 0.4)]

    # global statistics: build each shard with the df / avgdl of the whole corpus (what the all-reduce
    # inside build_synthetic_shard(group=...) produces on a real multi-GPU run)
    lib = _lib.load()
    engines, stats_f, parts = [], [], []
    for r in range(2):
        lo, hi = parallel.shard_bounds(spec.n_docs, 2, r)
        sh = synth_device.build_synthetic_shard(spec, lo, hi, dev)
        sh.set_bm25(sh.indptr, sh.postings, sh.dl, full.shard.avgdl, full.shard.df_host, spec.n_docs,
                    max_dl=spec.max_len)
        engines.append(SearchEngine(sh, max_batch=8))
    for eng in engines:                                   # local phase: scores + local stats
        stats = eng._stats(B)
        cos = eng.dense_scan(eng.upload_vectors(qb.vectors), stats)
        qt, qi, qo = eng.upload_terms(qb.term_ids)
        bm = eng.bm25_score(qt, qi, qo, B, stats)
        f = torch.empty((B, 4), dtype=torch.float32, device=dev)
        check(lib.hs_stats_decode(ptr(stats), ptr(f), B, stream_ptr(dev)))
        stats_f.append(f)
        parts.append((eng, cos, bm, stats))
    g = torch.stack(stats_f)                              # C2 by hand: min of col 0, max of cols 1, 2
    glob = torch.stack([g[:, :, 0].min(0).values, g[:, :, 1].max(0).values, g[:, :, 2].max(0).values,
                        g[0, :, 3]], dim=1).contiguous()
    lists = []
    for eng, cos, bm, stats in parts:
        check(lib.hs_stats_encode(ptr(glob), ptr(stats), B, stream_ptr(dev)))
        lists.append(eng.fuse_topk(2, cos, bm, stats, 0.6, 0.4, k).clone())
    gathered = torch.stack(lists).contiguous()            # C1 by hand: [2, B, k]
    merged = torch.empty((B, k), dtype=torch.int64, device=dev)
    check(lib.hs_topk_merge(ptr(gathered), 2, B, k, ptr(merged), stream_ptr(dev)))
    sc, ids = [t.cpu().numpy() for t in engines[0].unpack(merged)]
    assert np.array_equal(ids, ref_ids)
    assert np.array_equal(sc, ref_sc)
    # and the device merge agrees with the numpy twin used by the gloo test
    host = parallel.merge_keys_host(gathered.cpu().numpy().view(np.uint64), k)
    assert np.array_equal(host, merged.cpu().numpy().view(np.uint64))


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from hybrid_search_engine_b200 import parallel, synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec = _spec()
    lo, hi = parallel.shard_bounds(spec.n_docs, world, rank)
    eng = SearchEngine(synth_device.build_synthetic_shard(spec, lo, hi, dev, group=dist.group.WORLD),
                       group=dist.group.WORLD, max_batch=8)
    B, k = 6, 100
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B).tolist())
    sc, ids = eng.search_hybrid_bm25(qb, k, 0.6, 0.4)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), sc=sc.cpu().numpy(), ids=ids.cpu().numpy())
    sc2, ids2 = eng.search_bm25(QueryBatch(term_ids=qb.term_ids), k)
    np.savez(os.path.join(out_dir, f"bm25_rank{rank}.npz"), sc=sc2.cpu().numpy(), ids=ids2.cpu().numpy())
    dist.destroy_process_group()


def test_nccl_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    spec = _spec()
    full = SearchEngine(synth_device.build_synthetic_shard(spec, 0, spec.n_docs, "cuda:0"), max_batch=8)
    B, k = 6, 100
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B).tolist())
    sc, ids = [t.cpu().numpy().copy() for t in full.search_hybrid_bm25(qb, k, 0.6, 0.4)]
    sc2, ids2 = [t.cpu().numpy().copy() for t in full.search_bm25(QueryBatch(term_ids=qb.term_ids), k)]
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["ids"], ids) and np.array_equal(got["sc"], sc)
        got = np.load(tmp_path / f"bm25_rank{r}.npz")
        assert np.array_equal(got["ids"], ids2) and np.array_equal(got["sc"], sc2)


# ---------------------------------------------------------------------- doc-sharded PLUGIN API on real text
def _pipe_worker(rank, world, port, out_dir, backend):
    """create_pipeline(..., group=WORLD) on the T1 corpus: every rank indexes the same documents, keeps its doc range,
    and must return the 1-rank results.  backend 'nccl' = one GPU per rank; 'gloo' = all ranks on cuda:0 (the
    exchange steps are staged through the host), which is how a 1-GPU box covers the N > 1 path."""
    import pickle
    import torch.distributed as dist
    import hybrid_search_engine_b200 as hs
    from tests.golden_cases import load_case
    from tests.test_gpu_parity import TableEncoder
    from oracle import hybrid_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev_i = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev_i)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev_i))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    out = _run_all_pipelines(hs, load_case("t1_small"), TableEncoder, orc, dist.group.WORLD, f"cuda:{dev_i}")
    pickle.dump(out, open(os.path.join(out_dir, f"pipe{rank}.pkl"), "wb"))
    dist.destroy_process_group()


def _run_all_pipelines(hs, c, TableEncoder, orc, group, device):
    table = {orc.preprocess_text(d): e for d, e in zip(c.docs, c.emb)}
    table.update({q: e for q, e in zip(c.queries, c.q_emb)})
    enc = TableEncoder(table, c.emb.shape[1])
    rer = type("R", (), {"rerank": staticmethod(lambda q, cand, top_k=None: cand[:top_k] if top_k else cand)})()
    out = {}
    for name, kw, k in (("hybrid_bm25", {}, 50), ("bm25", {}, 30), ("multi_stage", dict(reranker=rer), 20),
                        ("basic", {}, 25), ("diversity", {}, 7)):
        p = hs.create_pipeline(name, encoder=enc, device=device, group=group, **kw)
        p.index(c.docs)
        res = p.search_many(c.queries, top_k=k)
        out[name] = [[(r["doc_id"], float(r["score"]), r["content"]) for r in x.results] for x in res]
        if name == "bm25":
            out["bm25_score"] = [p.bm25.score(c.queries[0], d) for d in (0, 1, 200, 399)]
            out["bm25_idf"] = sorted(p.bm25.idf.items())[:20]
    return out


def test_pipelines_doc_sharded_equal_unsharded(tmp_path):
    """All five pipelines with group= over 2 ranks == the same pipelines on one rank, bit for bit (ids, scores,
    contents), on real text with edge-case docs (T1 corpus)."""
    import pickle
    import torch.multiprocessing as mp
    import hybrid_search_engine_b200 as hs
    from oracle import hybrid_oracle as orc
    from tests.golden_cases import load_case
    from tests.test_gpu_parity import TableEncoder
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_pipe_worker, args=(2, port, str(tmp_path), backend), nprocs=2, join=True)
    want = _run_all_pipelines(hs, load_case("t1_small"), TableEncoder, orc, None, "cuda:0")
    for r in range(2):
        got = pickle.load(open(tmp_path / f"pipe{r}.pkl", "rb"))
        for name in want:
            assert got[name] == want[name], (backend, r, name)
