"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's hybrid scoring path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline.  The product package
``hybrid_search_engine_b200`` never imports it and has no CPU fallback.

Parity status: every function below except ``partial_ratio``/``lexical_scores`` is pinned against
outputs of the unmodified reference executed in the dev container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).  ``partial_ratio`` restates the
*published* algorithm of the third-party ``rapidfuzz`` (``requirements.txt:6``, no version pin, not
vendored, not installable here) and is therefore **parity unpinned** (SURVEY.md section 8c).

All ``file:line`` citations are into ``/root/reference/search_engine/``.

Dense "conformance order".  The reference's cosine (``utils.py:28-54``) is a numba ``fastmath``
float32 BLAS reduction whose summation order is implementation defined, so it cannot be restated
bit-for-bit.  The oracle (and the CUDA ``exact`` mode, which matches it bit-for-bit) define the
reduction instead as: products and sums in float64 (a product of two float32 is exact in float64),
element ``e = 128*c + 4*l + j`` accumulated sequentially over ``(c, j)`` into lane ``l`` of 32,
lanes combined by the butterfly ``16, 8, 4, 2, 1``; then the reference's own float32 steps
``f32(dot) / (f32(|q|) * f32(|v|))``.  This is within ~1 ulp(fp32) of the reference on every doc
(checked in the golden tests) and is order-deterministic.
"""
from __future__ import annotations

import math
import re
from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------- tokeniser
# extractor.py:6-12 (48 words)
STOPWORDS = frozenset(
    "a an the and or but in on at to for of with by from is are was were be been being have has "
    "had do does did will would could should may might must shall can this that these those i "
    "you he she it we they".split()
)
_TOKEN_RE = re.compile(r"[A-Za-z0-9_]+")
_WS_RE = re.compile(r"\s+")


def extract_tokens(text: str, remove_stopwords: bool = False) -> List[str]:
    """extractor.py:15-31 -- lower-case, ASCII ``[A-Za-z0-9_]+`` runs, optional stop-word removal."""
    if not text:
        return []
    toks = _TOKEN_RE.findall(text.lower())
    if remove_stopwords:
        toks = [t for t in toks if t not in STOPWORDS]
    return toks


def preprocess_text(text: str) -> str:
    """extractor.py:34-52 (remove_stopwords=False branch) -- collapse whitespace, strip."""
    if not text:
        return ""
    return _WS_RE.sub(" ", text.strip())


# --------------------------------------------------------------------------- BM25
@dataclass
class BM25State:
    """bm25.py:19-81 restated over an inverted CSR instead of per-doc dicts."""
    k1: float = 1.5
    b: float = 0.75
    doc_count: int = 0
    avg_doc_len: float = 0.0
    doc_lengths: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    vocab: Dict[str, int] = field(default_factory=dict)
    indptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))
    post_doc: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # ascending per term
    post_tf: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    idf: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    remove_stopwords: bool = True


def idf_from_df(doc_count: int, df: np.ndarray) -> np.ndarray:
    """bm25.py:81 -- ``math.log((N - df + 0.5) / (df + 0.5) + 1)`` (libm log, one call per df value)."""
    df = np.asarray(df, dtype=np.int64)
    uniq, inv = np.unique(df, return_inverse=True)
    vals = np.array([math.log((doc_count - int(d) + 0.5) / (int(d) + 0.5) + 1) for d in uniq],
                    dtype=np.float64)
    return vals[inv] if len(uniq) else np.zeros(0, np.float64)


def bm25_fit_tokens(doc_terms: Sequence[np.ndarray], n_vocab: int, k1=1.5, b=0.75,
                    vocab: Optional[Dict[str, int]] = None) -> BM25State:
    """bm25.py:45-81 on pre-tokenised docs (arrays of term ids, duplicates = tf)."""
    n = len(doc_terms)
    dl = np.array([len(t) for t in doc_terms], dtype=np.int64)
    if n and dl.sum():
        doc_of = np.repeat(np.arange(n, dtype=np.int64), dl)
        terms = np.concatenate([np.asarray(t, dtype=np.int64) for t in doc_terms])
        key = terms * n + doc_of
        uk, tf = np.unique(key, return_counts=True)          # sorted by (term, doc)
        pt, pd = uk // n, uk % n
    else:
        pt = pd = tf = np.zeros(0, np.int64)
    df = np.bincount(pt, minlength=n_vocab).astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(df)]).astype(np.int64)
    # bm25.py:71 -- sum(doc_lengths) / doc_count, 0 if no docs
    avg = (int(dl.sum()) / n) if n > 0 else 0
    idf = idf_from_df(n, df)
    return BM25State(k1=k1, b=b, doc_count=n, avg_doc_len=avg, doc_lengths=dl,
                     vocab=vocab or {}, indptr=indptr, post_doc=pd, post_tf=tf.astype(np.int64),
                     idf=idf)


def bm25_fit(documents: Sequence[str], k1=1.5, b=0.75, remove_stopwords=True) -> BM25State:
    """bm25.py:45-74 -- tokenise (stop words removed), dl = len(tokens), tf, df, avgdl, idf."""
    vocab: Dict[str, int] = {}
    doc_terms = []
    for doc in documents:
        ids = [vocab.setdefault(t, len(vocab)) for t in extract_tokens(doc, remove_stopwords)]
        doc_terms.append(np.array(ids, dtype=np.int64))
    st = bm25_fit_tokens(doc_terms, len(vocab), k1, b, vocab)
    st.remove_stopwords = remove_stopwords
    return st


def query_term_ids(st: BM25State, query: str) -> List[int]:
    """bm25.py:94,99-101 -- query tokens in order, duplicates kept, unknown terms skipped."""
    out = []
    for t in extract_tokens(query, st.remove_stopwords):
        tid = st.vocab.get(t)
        if tid is not None and st.indptr[tid + 1] > st.indptr[tid]:
            out.append(tid)
    return out


def _bm25_contrib(st: BM25State, tid: int, tf: np.ndarray, dl: np.ndarray) -> np.ndarray:
    """bm25.py:103-110 in float64 with the reference's operation order."""
    tf = tf.astype(np.float64)
    num = tf * (st.k1 + 1)
    den = tf + st.k1 * (1 - st.b + st.b * (dl.astype(np.float64) / st.avg_doc_len))
    with np.errstate(divide="ignore", invalid="ignore"):
        c = st.idf[tid] * (num / den)
    return np.where(den > 0, c, 0.0)


def bm25_scores_f64(st: BM25State, term_ids: Sequence[int], delta: Optional[float] = None) -> np.ndarray:
    """bm25.py:83-112 for all docs, float64, accumulating per query token *in order*.

    ``delta`` not None gives BM25Plus (bm25.py:150-179): every doc gets ``idf * (num/den + delta)``
    for every known query term, tf = 0 included.
    """
    s = np.zeros(st.doc_count, dtype=np.float64)
    for tid in term_ids:
        lo, hi = st.indptr[tid], st.indptr[tid + 1]
        d = st.post_doc[lo:hi]
        if delta is None:
            s[d] += _bm25_contrib(st, tid, st.post_tf[lo:hi], st.doc_lengths[d])
        else:
            tf = np.zeros(st.doc_count, np.int64)
            tf[d] = st.post_tf[lo:hi]
            tff = tf.astype(np.float64)
            num = tff * (st.k1 + 1)
            den = tff + st.k1 * (1 - st.b + st.b * (st.doc_lengths.astype(np.float64) / st.avg_doc_len))
            s += np.where(den > 0, st.idf[tid] * ((num / den) + delta), 0.0)
    return s


def bm25_score_batch(st: BM25State, query: str) -> np.ndarray:
    """bm25.py:114-127 -- float64 scores rounded once to float32."""
    return bm25_scores_f64(st, query_term_ids(st, query)).astype(np.float32)


def bm25_score_docs(st: BM25State, term_ids: Sequence[int], doc_ids: Sequence[int]) -> np.ndarray:
    """bm25.py:83-112 for selected docs only (multi_stage stage 2, pipelines.py:485): float64, unrounded."""
    doc_ids = np.asarray(doc_ids, dtype=np.int64)
    s = np.zeros(len(doc_ids), dtype=np.float64)
    for tid in term_ids:
        lo, hi = st.indptr[tid], st.indptr[tid + 1]
        d = st.post_doc[lo:hi]
        pos = np.searchsorted(d, doc_ids)
        pos_c = np.minimum(pos, max(len(d) - 1, 0))
        hit = (pos < len(d)) & (d[pos_c] == doc_ids) if len(d) else np.zeros(len(doc_ids), bool)
        if hit.any():
            c = _bm25_contrib(st, tid, st.post_tf[lo:hi][pos_c[hit]], st.doc_lengths[doc_ids[hit]])
            s[hit] += c
    return s


# --------------------------------------------------------------------------- dense cosine
def _lane_sum64(p: np.ndarray) -> np.ndarray:
    """Sum float64 ``p[n, d]`` per row in the conformance order (see module docstring)."""
    n, d = p.shape
    nchunk = (d + 127) // 128
    if d != nchunk * 128:
        p = np.concatenate([p, np.zeros((n, nchunk * 128 - d), np.float64)], axis=1)
    p = p.reshape(n, nchunk, 32, 4)
    acc = np.zeros((n, 32), np.float64)
    for c in range(nchunk):
        for j in range(4):
            acc = acc + p[:, c, :, j]
    for half in (16, 8, 4, 2, 1):
        acc = acc[:, :half] + acc[:, half:2 * half]
    return acc[:, 0]


def row_norms(v: np.ndarray, block: int = 65536) -> np.ndarray:
    """float32 ``f32(sqrt(sum64 v*v))`` per row, conformance order."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty(v.shape[0], np.float32)
    for s in range(0, v.shape[0], block):
        v64 = v[s:s + block].astype(np.float64)
        out[s:s + block] = np.sqrt(_lane_sum64(v64 * v64)).astype(np.float32)
    return out


def cosine_exact(q: np.ndarray, v: np.ndarray, vnorm: Optional[np.ndarray] = None,
                 block: int = 65536) -> np.ndarray:
    """utils.py:28-54 with the reduction in the conformance order; float32[N]."""
    q = np.asarray(q, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    n = v.shape[0]
    q64 = q.astype(np.float64)
    qn = np.float32(np.sqrt(_lane_sum64((q64 * q64)[None, :])[0]))
    if qn == 0.0:                                   # utils.py:44-45
        return np.zeros(n, np.float32)
    if vnorm is None:
        vnorm = row_norms(v, block)
    out = np.empty(n, np.float32)
    for s in range(0, n, block):
        dot = _lane_sum64(v[s:s + block].astype(np.float64) * q64[None, :]).astype(np.float32)
        vn = vnorm[s:s + block]
        den = (qn * vn).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            c = (dot / den).astype(np.float32)
        out[s:s + block] = np.where(vn == 0.0, np.float32(0.0), c)   # utils.py:49-50
    return out


def batch_cosine_sim_port(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """utils.py:28-54 as a CPU user would run it: float32 BLAS, norms recomputed per query.

    Used only as the timed CPU baseline (all host threads via BLAS); not bit-matched to anything.
    """
    q = np.asarray(q, np.float32)
    qn = np.linalg.norm(q)
    if qn == 0.0:
        return np.zeros(v.shape[0], np.float32)
    vn = np.sqrt(np.einsum("ij,ij->i", v, v, dtype=np.float32))
    with np.errstate(divide="ignore", invalid="ignore"):
        c = (v @ q) / (qn * vn)
    return np.where(vn == 0.0, np.float32(0.0), c).astype(np.float32)


def cosine_sim_f32(a: np.ndarray, b: np.ndarray) -> float:
    """utils.py:5-25 -- scalar cosine, conformance-order reduction, returns python float."""
    c = cosine_exact(np.asarray(a, np.float32), np.asarray(b, np.float32)[None, :])
    na = row_norms(np.asarray(a, np.float32)[None, :])[0]
    return 0.0 if na == 0.0 else float(c[0])


# --------------------------------------------------------------------------- normalise / fuse / select
def normalize_scores(x: np.ndarray) -> np.ndarray:
    """utils.py:57-71 -- min-max in the array's dtype; constant vector -> ones."""
    mn, mx = x.min(), x.max()
    if mx - mn == 0:
        return np.ones_like(x)
    return (x - mn) / (mx - mn)


def searcher_hybrid(sem_raw: np.ndarray, lex_raw: np.ndarray, sw: float, lw: float) -> np.ndarray:
    """core.py:232-233,264-268 -- weights must sum to 1; float32 ``norm(sem)*sw + norm(lex)*lw``."""
    if not np.isclose(sw + lw, 1.0):
        raise ValueError("semantic_weight and lexical_weight must sum to 1.0")
    sem = normalize_scores(sem_raw.astype(np.float32))
    lex = normalize_scores(lex_raw.astype(np.float32))
    return (sem * sw) + (lex * lw)


def hybrid_bm25_fused(sem_norm: np.ndarray, bm25: np.ndarray, ws: float, wl: float) -> np.ndarray:
    """pipelines.py:331-340 under NumPy >= 2 scalar promotion (python floats are weak):

    ``f32( f32(f64(sem_i)/max_sem * ws) + f32(f32(bm_i/max_bm) * f32(wl)) )``; ``max_sem`` is the
    python-float maximum of the (already min-max normalised) semantic scores, ``max_bm`` is
    ``bm.max()`` if positive else 1.
    """
    sem_norm = sem_norm.astype(np.float32)
    bm25 = bm25.astype(np.float32)
    max_sem = float(sem_norm.max()) if len(sem_norm) else 1
    max_bm = bm25.max() if bm25.max() > 0 else np.float32(1.0)
    t1 = ((sem_norm.astype(np.float64) / max_sem) * ws).astype(np.float32)
    t2 = (bm25 / max_bm) * np.float32(wl)
    return (t1 + t2).astype(np.float32)


def canonical_topk(scores: np.ndarray, k: int) -> np.ndarray:
    """Total order ``(score desc, doc_id asc)`` -- equals the reference wherever it is deterministic
    (stable ``list.sort(reverse=True)`` over ascending i, pipelines.py:342); the reference's
    ``np.argsort(x)[::-1]`` (core.py:271, bm25.py:141) is unstable among exact ties."""
    scores = np.asarray(scores)
    k = min(k, len(scores))
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    return order[:k].astype(np.int64)


# --------------------------------------------------------------------------- fuzzy lexical (PARITY UNPINNED)
def _lcs_len(a: str, b: str) -> int:
    if not a or not b:
        return 0
    prev = [0] * (len(b) + 1)
    for ca in a:
        cur = [0]
        for j, cb in enumerate(b):
            cur.append(prev[j] + 1 if ca == cb else max(prev[j + 1], cur[j]))
        prev = cur
    return prev[-1]


def partial_ratio(s1: str, s2: str) -> float:
    """rapidfuzz >= 3 ``fuzz.partial_ratio`` restated from its published description: best normalised
    indel similarity ``(1 - (len1 + len_window - 2*LCS) / (len1 + len_window)) * 100`` of the shorter string
    against every window of the longer one of at most its length, partial windows at both ends included.
    The arithmetic (distance -> normalised distance -> similarity -> x 100) is rapidfuzz's own and is
    pinned by the few known-answer vectors its public documentation carries (tests/golden/
    rapidfuzz_published.json); the window set is NOT pinned by them -- rapidfuzz itself is not installable
    here (PARITY UNPINNED beyond those vectors)."""
    if len(s1) > len(s2):
        s1, s2 = s2, s1
    n1, n2 = len(s1), len(s2)
    if n1 == 0:
        return 100.0 if n2 == 0 else 0.0
    best = 0.0
    for i in range(-n1 + 1, n2):
        w = s2[max(0, i):min(n2, i + n1)]
        lensum = n1 + len(w)
        r = (1.0 - (lensum - 2 * _lcs_len(s1, w)) / lensum) * 100.0
        if r > best:
            best = r
    return best


def lexical_scores(query: str, docs: Sequence[str], ratio_fn=partial_ratio) -> np.ndarray:
    """core.py:178-197 -- ``0.7*partial_ratio/100 + 0.3*|Q&D|/|Q|`` (token sets, stop words kept)."""
    qt = set(extract_tokens(query.lower()))
    out = []
    for doc in docs:
        fz = ratio_fn(query.lower(), doc.lower()) / 100.0
        dt = set(extract_tokens(doc.lower()))
        if qt and dt:
            out.append((fz * 0.7) + ((len(qt & dt) / len(qt)) * 0.3))
        else:
            out.append(fz)
    return np.array(out, dtype=np.float32)


# --------------------------------------------------------------------------- MMR
def mmr_select(emb: np.ndarray, rel: np.ndarray, lam: float, top_k: int,
               vnorm: Optional[np.ndarray] = None) -> List[int]:
    """pipelines.py:531-569 -- greedy MMR; ``max(key=)`` keeps the FIRST maximal candidate in
    remaining (= candidate rank) order; first pick has ``max_sim = 0``; cosine in float32
    (utils.py:5-25, here in the conformance order) combined in float64."""
    emb = np.ascontiguousarray(emb, np.float32)
    c = emb.shape[0]
    rel = np.asarray(rel, np.float64)
    if vnorm is None:
        vnorm = row_norms(emb)
    selected: List[int] = []
    alive = np.ones(c, bool)
    max_sim = np.zeros(c, np.float64)
    have = False
    while len(selected) < top_k and alive.any():
        ms = max_sim if have else np.zeros(c, np.float64)
        mmr = lam * rel - (1 - lam) * ms
        mmr = np.where(alive, mmr, -np.inf)
        best = int(np.argmax(mmr))                      # first maximal
        selected.append(best)
        alive[best] = False
        # cosine of every candidate against the new pick (utils.py:21-25: 0 if either norm is 0)
        sim = cosine_exact(emb[best], emb, vnorm).astype(np.float64)
        if vnorm[best] == 0.0:
            sim[:] = 0.0
        max_sim = sim if not have else np.maximum(max_sim, sim)
        have = True
    return selected


def diversity_relevance(scores: Sequence[float]) -> np.ndarray:
    """pipelines.py:586-589 -- float64 ``(s - min) / (max - min + 1e-8)``."""
    s = np.array(list(scores))
    return (s - s.min()) / (s.max() - s.min() + 1e-8)


# --------------------------------------------------------------------------- whole pipelines
@dataclass
class OracleIndex:
    documents: List[str]
    contents: List[str]            # whitespace-normalised (indexer.py:264)
    vectors: np.ndarray            # float32 [N, d]
    vnorm: np.ndarray
    bm25: BM25State


def build_index(documents: Sequence[str], vectors: np.ndarray, k1=1.5, b=0.75) -> OracleIndex:
    vectors = np.ascontiguousarray(vectors, np.float32)
    return OracleIndex(list(documents), [preprocess_text(d) for d in documents], vectors,
                       row_norms(vectors) if len(vectors) else np.zeros(0, np.float32),
                       bm25_fit(documents, k1, b))


def search_bm25(ix: OracleIndex, query: str, top_k: int):
    """pipelines.py:270-280 -> bm25.py:129-142.  Returns (ids, float scores)."""
    s = bm25_score_batch(ix.bm25, query)
    ids = canonical_topk(s, top_k)
    return ids, s[ids]


def semantic_only(ix: OracleIndex, q_vec: np.ndarray) -> np.ndarray:
    """Searcher.search with weights (1.0, 0.0): ``norm(cos)*1.0 + norm(lex)*0.0`` == norm(cos)."""
    if len(ix.vectors) == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    return normalize_scores(cosine_exact(q_vec, ix.vectors, ix.vnorm))


def search_hybrid_bm25(ix: OracleIndex, query: str, q_vec: np.ndarray, top_k: int,
                       ws: float = 0.6, wl: float = 0.4):
    """pipelines.py:315-357.  Returns (ids, float32 fused scores, full fused vector)."""
    sem = semantic_only(ix, q_vec)
    bm = bm25_score_batch(ix.bm25, query)
    fused = hybrid_bm25_fused(sem, bm, ws, wl)
    ids = canonical_topk(fused, top_k)
    return ids, fused[ids], fused


def search_multi_stage(ix: OracleIndex, query: str, q_vec: np.ndarray,
                       stage1_k: int = 100, stage2_k: int = 20):
    """pipelines.py:470-487 stages 1-2.  Returns (stage1 ids, stage2 ids, stage2 float64 scores)."""
    sem = semantic_only(ix, q_vec)
    s1 = canonical_topk(sem, stage1_k)
    bm = bm25_score_docs(ix.bm25, query_term_ids(ix.bm25, query), s1)
    order = np.lexsort((np.arange(len(s1)), -bm))        # stable: ties keep stage-1 rank
    keep = order[:stage2_k]
    return s1, s1[keep], bm[keep]


def search_basic(ix: OracleIndex, query: str, q_vec: np.ndarray, top_k: int, sw: float = 0.7,
                 ratio_fn=partial_ratio):
    """pipelines.py:85-103 -> core.py:199-285.  Returns (ids, float32 scores, full vector)."""
    if len(ix.vectors) == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    cos = cosine_exact(q_vec, ix.vectors, ix.vnorm)
    lex = lexical_scores(query, ix.contents, ratio_fn)
    hyb = searcher_hybrid(cos, lex, sw, 1.0 - sw)
    ids = canonical_topk(hyb, top_k)
    return ids, hyb[ids], hyb


def search_diversity(ix: OracleIndex, query: str, q_vec: np.ndarray, top_k: int,
                     lam: float = 0.5, ratio_fn=partial_ratio):
    """pipelines.py:571-613.  Returns (doc ids in selection order, their hybrid scores)."""
    cand, sc, _ = search_basic(ix, query, q_vec, top_k * 4, 0.7, ratio_fn)
    if len(cand) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32)
    rel = diversity_relevance([float(x) for x in sc])
    sel = mmr_select(ix.vectors[cand], rel, lam, top_k, ix.vnorm[cand])
    return cand[sel], sc[sel]


def search_faiss_style(ix: OracleIndex, query: str, q_vec: np.ndarray, top_k: int, sw: float = 0.7,
                       ratio_fn=partial_ratio):
    """core.py:244-250 + 264-276 with use_faiss=True (PARITY UNPINNED: faiss is not installable; the inner
    product of the stored L2-normalised rows with the normalised query is restated as the conformance cosine).
    Only the best min(2k, N) semantic scores survive, all other docs keep 0.0 before min-max."""
    n = len(ix.vectors)
    cos = cosine_exact(q_vec, ix.vectors, ix.vnorm)
    keep = canonical_topk(cos, min(2 * top_k, n))
    sem = np.zeros(n, np.float32)
    sem[keep] = cos[keep]
    lw = 1.0 - sw
    lex = lexical_scores(query, ix.contents, ratio_fn) if lw != 0.0 else np.zeros(n, np.float32)
    hyb = searcher_hybrid(sem, lex, sw, lw)
    ids = canonical_topk(hyb, top_k)
    return ids, hyb[ids]
