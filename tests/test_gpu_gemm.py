"""GPU tests of the tensor-core dense modes (csrc/dense_gemm.cu): 3xTF32 and bf16 tcgen05 GEMM, the
two-query-tile pass, doc ranges, and the filter epilogue of the pure-semantic search.

Tolerances: ``tf32x3`` is a float32-grade mode -- |cos - exact| <= 2e-6 absolute, measured 9e-7 at worst (the
tensor core accumulates its float32 partial sums with truncation, which dominates the 3xTF32 split error of
~1e-7; the ``fp32`` CUDA-core mode is within 2.4e-7, the reference's own BLAS/numba cosine is pinned to 4 ulp).
On fused scores that is ~2e-6 relative, inside the north star's 1e-5.  ``bf16`` within 1e-2 (north star).  The filtered search must equal the stored-matrix search of the SAME mode bit for bit.
"""
import numpy as np
import pytest

from oracle import hybrid_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def hs():
    import hybrid_search_engine_b200 as hs
    from hybrid_search_engine_b200 import _lib
    _lib.load()
    return hs


def _dec(e):
    e = int(e)
    u = (e & 0x7FFFFFFF) if (e & 0x80000000) else ((~e) & 0xFFFFFFFF)
    return np.array([u], np.uint32).view(np.float32)[0]


def _engine(hs, v, **kw):
    from hybrid_search_engine_b200.engine import SearchEngine
    shard = hs.DeviceIndex("cuda:0", v.shape[0])
    shard.set_dense(torch.from_numpy(v).cuda())
    return SearchEngine(shard, **kw)


@pytest.mark.parametrize("n,d,B", [(1000, 384, 5), (20000, 384, 130), (3000, 768, 40), (777, 100, 33), (4097, 130, 128),
                                   (129, 64, 1), (50000, 384, 128)])
def test_tf32x3_within_fp32_tolerance_of_exact(hs, n, d, B):
    """HS_DENSE_TF32X3: one corpus pass per 128 queries at float32-grade accuracy; stats match the produced
    vectors; zero query / zero row rules; two runs give the same bits."""
    rng = np.random.default_rng(n + d + B)
    v = rng.standard_normal((n, d)).astype(np.float32)
    v[3] = 0.0
    v[5] *= 1e-3                                           # small-norm row: relative accuracy must hold
    q = rng.standard_normal((B, d)).astype(np.float32)
    if B > 1:
        q[1] = 0.0
    eng = _engine(hs, v, max_batch=256)
    stats = eng._stats(B)
    qd = eng.upload_vectors(q)
    cos = eng.dense_scan(qd, stats, "tf32x3").cpu().numpy().copy()
    st = stats.cpu().numpy().view(np.uint32).copy()
    stats2 = eng._stats(B)
    cos2 = eng.dense_scan(qd, stats2, "tf32x3").cpu().numpy()
    assert np.array_equal(cos, cos2), "tf32x3 is not bit-for-bit reproducible"
    worst = 0.0
    for b in sorted({0, min(1, B - 1), B // 2, B - 1}):
        want = orc.cosine_exact(q[b], v)
        worst = max(worst, float(np.max(np.abs(cos[b] - want))))
        assert _dec(st[b, 0]) == cos[b].min() and _dec(st[b, 1]) == cos[b].max()
    print(f"tf32x3 max|cos - exact| = {worst:.3e} (n={n}, d={d}, B={B})")
    assert worst <= 2e-6
    if B > 1:
        assert np.all(cos[1] == 0.0)
    assert np.all(cos[:, 3] == 0.0)


@pytest.mark.parametrize("n,d,B", [(20000, 384, 200), (6000, 768, 256), (5000, 100, 300), (100000, 384, 200),
                                   (80000, 768, 130), (76000, 768, 300)])
def test_bf16_two_query_tiles_per_pass(hs, n, d, B):
    """B > 128 in bf16: two 128-query tiles share every landed corpus block (one pass per 256 queries).  From ~76 k docs
    on (>= 4 tiles per SM) the streamed query blocks are TMA-multicast across clusters of 2 CTAs; the ragged tile counts
    here make some CTAs of a cluster run surplus (out-of-range) tiles."""
    rng = np.random.default_rng(n + d + B)
    v = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((B, d)).astype(np.float32)
    eng = _engine(hs, v, max_batch=512)
    stats = eng._stats(B)
    cos = eng.dense_scan(eng.upload_vectors(q), stats, "bf16").cpu().numpy()
    st = stats.cpu().numpy().view(np.uint32)
    for b in (0, 127, 128, 129, B - 1):
        want = orc.cosine_exact(q[b], v)
        assert np.max(np.abs(cos[b] - want)) <= 1e-2
        assert _dec(st[b, 0]) == cos[b].min() and _dec(st[b, 1]) == cos[b].max()
    # the single-tile launch (first 100 queries alone) gives the same bits for those queries
    stats1 = eng._stats(100)
    cos1 = eng.dense_scan(eng.upload_vectors(q[:100]), stats1, "bf16").cpu().numpy()
    assert np.array_equal(cos1, cos[:100])


@pytest.mark.parametrize("mode", ["tf32x3", "bf16"])
def test_gemm_doc_range_matches_full_pass(hs, mode):
    """hs_dense_gemm over [doc_lo, doc_hi) writes exactly the columns of the full pass (ragged ends)."""
    from hybrid_search_engine_b200 import _lib
    from hybrid_search_engine_b200._lib import check, ptr, stream_ptr
    rng = np.random.default_rng(9)
    n, d, B = 9000, 96, 37
    v = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((B, d)).astype(np.float32)
    eng = _engine(hs, v)
    m = _lib.DENSE_MODES[mode]
    stats = eng._stats(B)
    qd = eng.upload_vectors(q)
    full = eng.dense_scan(qd, stats, mode).cpu().numpy().copy()
    lo, hi = 1234, 7777
    out = torch.full((B, hi - lo + 5), -7.0, dtype=torch.float32, device="cuda")
    wsp, nbytes = eng._gemm_ws(B, m)
    st2 = eng._stats(B)
    check(eng.lib.hs_dense_gemm(eng.shard.handle, ptr(qd), B, qd.stride(0), m, lo, hi, wsp, nbytes, ptr(out),
                                out.stride(0), ptr(st2), stream_ptr(eng.device)), "hs_dense_gemm")
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :hi - lo], full[:, lo:hi])
    assert np.all(got[:, hi - lo:] == -7.0)
    s2 = st2.cpu().numpy().view(np.uint32)
    for b in (0, B - 1):
        assert _dec(s2[b, 0]) == full[b, lo:hi].min() and _dec(s2[b, 1]) == full[b, lo:hi].max()


@pytest.mark.parametrize("mode,B,k", [("bf16", 300, 100), ("tf32x3", 130, 100), ("bf16", 64, 7), ("tf32x3", 5, 300)])
def test_filtered_semantic_search_equals_stored_matrix_search(hs, mode, B, k):
    """The GEMM-epilogue candidate filter (no [B, n] matrix) returns the same keys as GEMM -> store -> select."""
    from hybrid_search_engine_b200.engine import QueryBatch
    rng = np.random.default_rng(B + k)
    n, d = 300_000, 64
    v = rng.standard_normal((n, d)).astype(np.float32)
    v[1000:1100] = v[2000:2100]                           # duplicate rows: exact score ties across the corpus
    v[17] = 0.0
    q = rng.standard_normal((B, d)).astype(np.float32)
    q[2] = v[1003] * 0.5                                  # best hits are a tied pair
    eng = _engine(hs, v, max_batch=128)
    qb = QueryBatch(vectors=q)
    s_f, i_f = eng.search_semantic(qb, k, 0.7, dense_mode=mode, filtered=True)
    s_f, i_f = s_f.cpu().numpy().copy(), i_f.cpu().numpy().copy()
    s_s, i_s = eng.search_semantic(qb, k, 0.7, dense_mode=mode, filtered=False)
    s_s, i_s = s_s.cpu().numpy(), i_s.cpu().numpy()
    assert np.array_equal(i_f, i_s)
    assert np.array_equal(s_f, s_s)
    assert set(i_f[2][:2].tolist()) == {1003, 2003}


def test_tf32x3_hybrid_ids_match_exact_up_to_near_ties(hs):
    """hybrid_bm25 through the engine with dense_mode='tf32x3' at B=128: same top-100 as the exact mode except
    inside groups whose fused scores differ by <= 4 ulp."""
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = synth.SynthSpec(n_docs=200_000, vocab=50_000, dim=384)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"))
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    B = 128
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
    eng = SearchEngine(shard, max_batch=128)
    s_t, i_t = eng.search_hybrid_bm25(qb, 100, 0.6, 0.4, dense_mode="tf32x3")
    s_t, i_t = s_t.cpu().numpy().copy(), i_t.cpu().numpy().copy()
    s_e, i_e = eng.search_hybrid_bm25(qb, 100, 0.6, 0.4, dense_mode="exact")
    s_e, i_e = s_e.cpu().numpy(), i_e.cpu().numpy()
    assert np.max(np.abs(s_t - s_e)) <= 1e-5 * np.max(np.abs(s_e))
    same = 0
    for b in range(B):
        if np.array_equal(i_t[b], i_e[b]):
            same += 1
            continue
        sc = dict(zip(i_e[b].tolist(), s_e[b].tolist()))
        sc.update({i: s for i, s in zip(i_t[b].tolist(), s_t[b].tolist()) if i not in sc})
        for a, c in zip(i_t[b], i_e[b]):
            assert a == c or abs(sc[int(a)] - sc[int(c)]) <= 4 * np.spacing(np.float32(abs(sc[int(c)]))), (b, a, c)
    print(f"tf32x3 hybrid: {same}/{B} queries with identical top-100 ids, the rest differ only inside 4-ulp ties")


@pytest.mark.parametrize("B,k", [(300, 100), (17, 10), (130, 400)])
def test_bf16_exact_mode_is_bit_identical_to_exact(hs, B, k):
    """dense_mode="bf16_exact": bf16 tensor-core screen + verification in the conformance order returns the ids AND float32
    scores of the exact mode (== oracle) for hybrid_bm25, semantic-only and searcher-with-lexical-vector fusion; the
    count of queries that needed the exact fallback is printed (0 expected on this corpus)."""
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = synth.SynthSpec(n_docs=300_000, vocab=50_000, dim=96)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"))
    v = shard.vectors
    v[1000:1050] = v[2000:2050]                          # duplicate rows: exact ties near the top
    v[17] = 0.0
    shard.set_dense(v[:, :spec.dim].clone())
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    qv = synth.query_embeddings(spec, 0, B)
    qv[2] = v[1003, :spec.dim].cpu().numpy() * 0.5
    if B > 5:
        qv[5] = 0.0                                      # zero query: every cosine 0 -> constant -> ones
    qb = QueryBatch(vectors=qv, term_ids=synth.query_terms(spec, 0, B, th).tolist())
    eng = SearchEngine(shard, max_batch=256)
    got = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="bf16_exact")]
    eng.max_batch = 8
    want = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="exact")]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    eng.max_batch = 256
    got = [t.cpu().numpy().copy() for t in eng.search_semantic(QueryBatch(vectors=qv), k, 0.7, dense_mode="bf16_exact")]
    eng.max_batch = 8
    want = [t.cpu().numpy().copy() for t in eng.search_semantic(QueryBatch(vectors=qv), k, 0.7, dense_mode="exact")]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    lex = torch.rand((min(B, 40), spec.n_docs), device="cuda", dtype=torch.float32)
    qs = QueryBatch(vectors=qv[:lex.shape[0]])
    eng.max_batch = 256
    got = [t.cpu().numpy().copy() for t in eng.search_searcher(qs, lex, k, 0.7, 0.3, dense_mode="bf16_exact")]
    eng.max_batch = 8
    want = [t.cpu().numpy().copy() for t in eng.search_searcher(qs, lex, k, 0.7, 0.3, dense_mode="exact")]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    print(f"bf16_exact: {getattr(eng, 'verify_fallbacks', 0)} of {2 * B + lex.shape[0]} queries fell back to the exact mode")
    # BM25 screened in binary16 as well, candidates re-scored exactly (HS_SCREEN_BM25_F16=1; off by default)
    eng.max_batch, eng.screen_bm25_f16 = 256, True
    got = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="bf16_exact")]
    eng.max_batch, eng.screen_bm25_f16 = 8, False
    want = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="exact")]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    # the float32 screen (HS_SCREEN_F32=1; the default keeps the screen scores as binary16) proves the same result
    eng.max_batch, eng.screen_f16 = 256, False
    got = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="bf16_exact")]
    eng.max_batch = 8
    want = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="exact")]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])


@pytest.mark.parametrize("n_docs", [50_003, 4_099, 130])
def test_bf16_exact_ragged_shard_sizes(hs, n_docs):
    """Shard sizes that are not a multiple of 2 / 8 / 128: the binary16 screen rows are padded to 8 elements, the last
    tile stores a lone half, the select's vector path falls back on the ragged tail -- still the exact mode's bits."""
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = synth.SynthSpec(n_docs=n_docs, vocab=5_000, dim=72)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"))
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    B, k = 37, 20
    qb = QueryBatch(vectors=synth.query_embeddings(spec, 0, B), term_ids=synth.query_terms(spec, 0, B, th).tolist())
    eng = SearchEngine(shard, max_batch=64)
    want = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="exact")]
    for bm25_f16 in (False, True):
        eng.screen_bm25_f16 = bm25_f16
        got = [t.cpu().numpy().copy() for t in eng.search_hybrid_bm25(qb, k, 0.6, 0.4, dense_mode="bf16_exact")]
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0]), bm25_f16


def test_bf16_exact_pipeline_matches_oracle_on_real_text(hs):
    """create_pipeline("hybrid_bm25", dense_mode="bf16_exact") on the T1 corpus (ties, empty docs, zero vectors) == oracle."""
    from tests.golden_cases import load_case
    c = load_case("t1_mid")
    ix = orc.build_index(c.docs, c.emb)
    p = hs.create_pipeline("hybrid_bm25", dense_mode="bf16_exact")
    p.index(c.docs, embeddings=c.emb)
    res = p.search_many(c.queries, top_k=50, query_vectors=c.q_emb)
    for qi, q in enumerate(c.queries):
        ids, sc, _ = orc.search_hybrid_bm25(ix, q, c.q_emb[qi], 50)
        assert [r["doc_id"] for r in res[qi].results] == ids.tolist(), q
        assert np.array_equal(np.array([r["score"] for r in res[qi].results], np.float32), sc), q


def test_bf16_exact_falls_back_when_the_bound_cannot_be_proven(hs):
    """A corpus of near-duplicate vectors: every cosine lies inside the bf16 error band, so no top-k can be PROVEN from
    the screen -- every query must be flagged and redone in the exact mode, in the one-shot call and in the serving
    loop (where the flags are read back with the result, not between batches).  Result == exact mode, bit for bit."""
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = synth.SynthSpec(n_docs=60_000, vocab=20_000, dim=96)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"))
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(spec.dim, device="cuda", generator=g)
    v = base[None, :] + 1e-4 * torch.randn((spec.n_docs, spec.dim), device="cuda", generator=g)
    shard.set_dense(v.contiguous())
    th = synth.zipf_thresholds(spec.vocab, spec.zipf_s)
    sizes = [20, 7, 20]
    batches, o = [], 0
    for B in sizes:
        batches.append(QueryBatch(vectors=synth.query_embeddings(spec, o, o + B),
                                  term_ids=synth.query_terms(spec, o, o + B, th).tolist()))
        o += B
    eng = SearchEngine(shard, max_batch=32)
    want = []
    for qb in batches:
        sc, ids = eng.search_hybrid_bm25(qb, 50, 0.6, 0.4, dense_mode="exact")
        want.append((sc.cpu().numpy().copy(), ids.cpu().numpy().copy()))
    eng.verify_fallbacks = 0
    for qb, (ws_, wi) in zip(batches, want):
        sc, ids = eng.search_hybrid_bm25(qb, 50, 0.6, 0.4, dense_mode="bf16_exact")
        assert np.array_equal(ids.cpu().numpy(), wi) and np.array_equal(sc.cpu().numpy(), ws_)
    assert eng.verify_fallbacks >= sum(sizes) // 2, eng.verify_fallbacks
    eng.verify_fallbacks = 0
    got = list(eng.search_hybrid_bm25_stream(batches, 50, 0.6, 0.4, dense_mode="bf16_exact"))
    assert eng.verify_fallbacks >= sum(sizes) // 2, eng.verify_fallbacks
    for (gs, gi), (ws_, wi) in zip(got, want):
        assert np.array_equal(gi, wi) and np.array_equal(gs, ws_)


def test_large_k_select_with_many_lists_is_stable(hs):
    """k > 512 runs the select without a pre-computed bound: the first bound is published by the query's own CTAs while later
    CTAs start.  Every CTA must read it ONCE (a per-thread read split a CTA between two code paths with different barriers:
    an intermittent cudaErrorIllegalInstruction once a query had more CTAs than fit on the GPU at a time).  40 repetitions of
    the call that used to fail 1 time in ~15, each equal to the first."""
    from hybrid_search_engine_b200 import synth, synth_device
    from hybrid_search_engine_b200.engine import QueryBatch, SearchEngine
    spec = synth.SynthSpec(n_docs=300_000, vocab=50_000, dim=96)
    shard = synth_device.build_synthetic_shard(spec, 0, spec.n_docs, torch.device("cuda:0"), lexical=False)
    B, k = 130, 400
    qv = synth.query_embeddings(spec, 0, B)
    qv[5] = 0.0
    eng = SearchEngine(shard, max_batch=256)
    first = None
    for _ in range(40):
        s, i = eng.search_semantic(QueryBatch(vectors=qv), k, 0.7, dense_mode="bf16_exact")
        got = (s.cpu().numpy().copy(), i.cpu().numpy().copy())
        if first is None:
            first = got
        assert np.array_equal(got[1], first[1]) and np.array_equal(got[0], first[0])
    eng.max_batch = 8
    s, i = eng.search_semantic(QueryBatch(vectors=qv), k, 0.7, dense_mode="exact")
    assert np.array_equal(i.cpu().numpy(), first[1]) and np.array_equal(s.cpu().numpy(), first[0])
