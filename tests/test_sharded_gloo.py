"""N > 1 host logic on CPU: world_size-2 and -3 gloo runs of the doc-sharded exchange (C2 stats all-reduce,
C1 key all-gather + merge) around per-shard scores computed by the oracle.  The sharded result must
equal the unsharded oracle bit for bit (SURVEY.md section 8e "exactness under sharding")."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybrid_search_engine_b200 import parallel, synth
from oracle import hybrid_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = synth.SynthSpec(n_docs=1501, vocab=400, dim=48, min_len=3, max_len=30)
    docs = synth.doc_texts(spec, 0, spec.n_docs)
    emb = synth.embeddings(spec, 0, spec.n_docs)
    ix = orc.build_index(docs, emb)                       # global statistics (idf, avgdl) on every rank
    lo, hi = parallel.shard_bounds(spec.n_docs, world, rank)
    queries = synth.query_texts(spec, 0, 5)
    qv = synth.query_embeddings(spec, 0, 5)
    B, k = len(queries), 37
    cos = np.stack([orc.cosine_exact(qv[b], emb[lo:hi], ix.vnorm[lo:hi]) for b in range(B)])
    bm = np.stack([orc.bm25_score_batch(ix.bm25, q)[lo:hi] for q in queries])
    nan = np.float32("nan")
    st = np.stack([cos.min(1), cos.max(1), bm.max(1), np.full(B, nan, np.float32)], axis=1).astype(np.float32)
    g = parallel.allreduce_stats(torch.from_numpy(st.copy())).numpy()          # C2
    keys = np.zeros((B, k), np.uint64)
    for b in range(B):
        rng_a = g[b, 1] - g[b, 0]
        sem = np.ones_like(cos[b]) if rng_a == 0 else (cos[b] - g[b, 0]) / rng_a
        mb = g[b, 2] if g[b, 2] > 0 else np.float32(1.0)
        t1 = (sem.astype(np.float64) * 0.6).astype(np.float32)
        fused = (t1 + (bm[b] / mb) * np.float32(0.4)).astype(np.float32)
        kk = parallel.pack_keys(fused, np.arange(lo, hi))
        top = np.sort(kk)[::-1][:k]
        keys[b, :len(top)] = top
    gathered = parallel.allgather_keys(torch.from_numpy(keys.view(np.int64)))   # C1
    merged = parallel.merge_keys_host(gathered.numpy().view(np.uint64), k)
    sc, ids = parallel.unpack_keys(merged)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), sc=sc, ids=ids, lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spec = synth.SynthSpec(n_docs=1501, vocab=400, dim=48, min_len=3, max_len=30)
    docs = synth.doc_texts(spec, 0, spec.n_docs)
    ix = orc.build_index(docs, synth.embeddings(spec, 0, spec.n_docs))
    queries = synth.query_texts(spec, 0, 5)
    qv = synth.query_embeddings(spec, 0, 5)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    r0 = res[0]
    assert [(int(r["lo"]), int(r["hi"])) for r in res] == [parallel.shard_bounds(1501, world, r) for r in range(world)]
    if world == 2:
        assert (int(res[1]["lo"]), int(res[1]["hi"])) == (751, 1501)
    for r in res[1:]:                                                             # every rank: same result
        assert np.array_equal(r0["ids"], r["ids"]) and np.array_equal(r0["sc"], r["sc"])
    for b, q in enumerate(queries):
        ids, sc, _ = orc.search_hybrid_bm25(ix, q, qv[b], 37)
        assert np.array_equal(r0["ids"][b], ids)
        assert np.array_equal(r0["sc"][b], sc)


def test_shard_bounds_and_key_roundtrip():
    for n, w in ((10, 1), (10, 3), (7, 8), (0, 4), (10_000_000, 8)):
        b = [parallel.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[:4] = [0.0, -0.0, np.float32(1e-38), -np.float32(1e-38)]
    ids = rng.permutation(1000)
    k = parallel.pack_keys(s, ids)
    sc, di = parallel.unpack_keys(k)
    assert np.array_equal(di, ids) and np.array_equal(sc, np.where(s == 0, 0, s).astype(np.float32))
    order = np.argsort(k)[::-1]
    assert np.array_equal(ids[order], ids[orc.canonical_topk(s, 1000)])         # key order == canonical order


# ---------------------------------------------------------------------- doc-sharded BM25.fit (LexicalStats.merge_across)
def _fit_worker(rank, world, port, out_dir):
    from hybrid_search_engine_b200.index import LexicalStats
    from tests.golden_cases import load_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    docs = load_case("t1_small").docs
    lo, hi = parallel.shard_bounds(len(docs), world, rank)
    st = LexicalStats().fit(docs[lo:hi]).merge_across(dist.group.WORLD)
    np.savez(os.path.join(out_dir, f"fit{rank}.npz"), terms=np.array(list(st.vocab), dtype=object), df=st.df,
             indptr=st.indptr, postings=st.postings, dl=st.doc_lengths, local_dl=st.local_doc_lengths,
             n=st.doc_count, avg=np.float64(st.avg_doc_len), lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fit_equals_single_process_fit(tmp_path, world):
    """Every rank fits its doc range; after the merge the vocabulary (term ids), df, doc lengths and avgdl are those of a
    single-process fit (== the reference's, tests/test_host_cpu.py) and the local CSR is the global CSR restricted to
    the rank's docs with rebased doc ids."""
    from hybrid_search_engine_b200.index import LexicalStats
    from tests.golden_cases import load_case
    mp.spawn(_fit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    docs = load_case("t1_small").docs
    full = LexicalStats().fit(docs)
    for r in range(world):
        g = np.load(tmp_path / f"fit{r}.npz", allow_pickle=True)
        lo, hi = int(g["lo"]), int(g["hi"])
        assert g["terms"].tolist() == list(full.vocab)
        assert np.array_equal(g["df"], full.df) and np.array_equal(g["dl"], full.doc_lengths)
        assert int(g["n"]) == full.doc_count and float(g["avg"]) == float(full.avg_doc_len)
        assert np.array_equal(g["local_dl"], full.doc_lengths[lo:hi])
        want_ptr, want_post = [0], []
        for t in range(len(full.df)):
            sl = full.postings[full.indptr[t]:full.indptr[t + 1]]
            sl = sl[(sl[:, 0] >= lo) & (sl[:, 0] < hi)].copy()
            sl[:, 0] -= lo
            want_post.append(sl)
            want_ptr.append(want_ptr[-1] + len(sl))
        assert np.array_equal(g["indptr"], np.asarray(want_ptr))
        assert np.array_equal(g["postings"], np.concatenate(want_post))
