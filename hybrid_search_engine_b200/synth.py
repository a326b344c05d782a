"""Counter-based synthetic corpus -- host (numpy) definition.

Every value is a pure function of ``(seed, index...)`` through splitmix64, built from integer and
single-rounding IEEE operations only, so the CUDA generators in ``csrc/synth.cu`` reproduce it
bit-for-bit on the device without the corpus ever existing on the host.

* vocabulary     ``t0 .. t{V-1}`` -- ASCII alnum, survives the reference tokeniser
                  (``extractor.py:28``), none is a stop word (``extractor.py:6-12``)
* doc length     ``L_i = min_len + mix(seed_len, i, 0) mod (max_len - min_len + 1)``
* token (i, j)   inverse-CDF Zipf(s) over V of the 64-bit hash ``mix(seed_tok, i, j)`` against an
                  integer threshold table (``zipf_thresholds``), no floating point at lookup time
* embedding      Irwin-Hall(4) of the four 16-bit fields of ``mix(seed_emb, i, j)``, centred and
                  scaled to unit variance: exact small integer times one fp32 constant
* queries        ``n_q_terms`` tokens from the same Zipf with ``seed_q`` (duplicates kept -- the
                  reference double counts them, ``bm25.py:99``), embedding with ``seed_qemb``
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)

# 1 / sqrt(4 * (65536**2 - 1) / 12): Irwin-Hall(4) over 16-bit uniforms -> unit variance
EMB_SCALE = np.float32(1.0 / 37837.22723)
EMB_CENTRE = 131070  # 2 * 65535


@dataclass(frozen=True)
class SynthSpec:
    n_docs: int
    vocab: int = 1_000_000
    dim: int = 384
    min_len: int = 100
    max_len: int = 300
    zipf_s: float = 1.0
    n_q_terms: int = 4
    seed_tok: int = 1234
    seed_len: int = 1235
    seed_emb: int = 1236
    seed_q: int = 1237
    seed_qemb: int = 1238


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + _GOLDEN) & _M64
        z = ((z ^ (z >> np.uint64(30))) * _C1) & _M64
        z = ((z ^ (z >> np.uint64(27))) * _C2) & _M64
        return z ^ (z >> np.uint64(31))


def seed_key(seed: int) -> np.uint64:
    """Scramble a small user seed into a 64-bit stream key (computed once on the host)."""
    return splitmix64(np.array([seed], dtype=np.uint64))[0]


def mix(seed: int, a, b) -> np.ndarray:
    """``splitmix64(splitmix64(seed_key(seed) + a) + b)`` -- the one hash every generator uses.

    The seed is scrambled first so that streams of nearby seeds never alias for nearby ``a``.
    """
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return splitmix64((splitmix64((seed_key(seed) + a) & _M64) + b) & _M64)


def zipf_thresholds(vocab: int, s: float = 1.0) -> np.ndarray:
    """uint64[V] table T with ``term(h) = #{r : T[r] <= h}`` clipped to V-1."""
    w = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    cdf = np.minimum(cdf, 1.0 - 2.0 ** -53)
    t = (cdf * 2.0 ** 63).astype(np.uint64) << np.uint64(1)
    t[-1] = _M64
    return t


def terms_from_hash(h: np.ndarray, thresholds: np.ndarray) -> np.ndarray:
    r = np.searchsorted(thresholds, h, side="right")
    return np.minimum(r, len(thresholds) - 1).astype(np.uint32)


def doc_lengths(spec: SynthSpec, lo: int, hi: int) -> np.ndarray:
    i = np.arange(lo, hi, dtype=np.uint64)
    span = np.uint64(spec.max_len - spec.min_len + 1)
    return (np.uint64(spec.min_len) + mix(spec.seed_len, i, 0) % span).astype(np.uint32)


def doc_tokens(spec: SynthSpec, lo: int, hi: int, thresholds: np.ndarray | None = None):
    """Return ``(lengths u32[n], terms u32[sum lengths])`` for docs ``lo..hi-1`` (doc-major)."""
    if thresholds is None:
        thresholds = zipf_thresholds(spec.vocab, spec.zipf_s)
    dl = doc_lengths(spec, lo, hi)
    doc = np.repeat(np.arange(lo, hi, dtype=np.uint64), dl)
    start = np.concatenate([[0], np.cumsum(dl, dtype=np.int64)])
    pos = (np.arange(start[-1], dtype=np.int64) - np.repeat(start[:-1], dl)).astype(np.uint64)
    return dl, terms_from_hash(mix(spec.seed_tok, doc, pos), thresholds)


def doc_texts(spec: SynthSpec, lo: int, hi: int, thresholds: np.ndarray | None = None):
    dl, terms = doc_tokens(spec, lo, hi, thresholds)
    out, p = [], 0
    for n in dl.tolist():
        out.append(" ".join(f"t{t}" for t in terms[p:p + n].tolist()))
        p += n
    return out


def _irwin_hall(h: np.ndarray) -> np.ndarray:
    m = np.uint64(0xFFFF)
    s = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
    return (s.astype(np.int64) - EMB_CENTRE).astype(np.float32) * EMB_SCALE


def embeddings(spec: SynthSpec, lo: int, hi: int, seed: int | None = None) -> np.ndarray:
    seed = spec.seed_emb if seed is None else seed
    i = np.arange(lo, hi, dtype=np.uint64)[:, None]
    j = np.arange(spec.dim, dtype=np.uint64)[None, :]
    return _irwin_hall(mix(seed, i, j))


def query_terms(spec: SynthSpec, lo: int, hi: int, thresholds: np.ndarray | None = None):
    """uint32[hi-lo, n_q_terms] term ids of queries lo..hi-1."""
    if thresholds is None:
        thresholds = zipf_thresholds(spec.vocab, spec.zipf_s)
    i = np.arange(lo, hi, dtype=np.uint64)[:, None]
    j = np.arange(spec.n_q_terms, dtype=np.uint64)[None, :]
    return terms_from_hash(mix(spec.seed_q, i, j), thresholds)


def query_texts(spec: SynthSpec, lo: int, hi: int, thresholds: np.ndarray | None = None):
    return [" ".join(f"t{t}" for t in row) for row in query_terms(spec, lo, hi, thresholds).tolist()]


def query_embeddings(spec: SynthSpec, lo: int, hi: int) -> np.ndarray:
    return embeddings(spec, lo, hi, seed=spec.seed_qemb)


class SynthVocab:
    """Term identity of the synthetic corpus for the plugin API: the vocabulary is ``t0 .. t{V-1}`` and term ``t<i>``
    has id ``i``, so a query STRING maps to term ids exactly as ``BM25.fit`` on the materialised texts would map it
    (tokeniser + stop-word removal of extractor.py:15-31; ids of terms outside the vocabulary are dropped)."""

    def __init__(self, spec: SynthSpec):
        self.vocab_size = spec.vocab

    def query_term_ids(self, query: str):
        from .extractor import extract_tokens
        out = []
        for t in extract_tokens(query, remove_stopwords=True):
            if len(t) > 1 and t[0] == "t" and t[1:].isdigit() and int(t[1:]) < self.vocab_size:
                out.append(int(t[1:]))
        return out


class LazyDocs:
    """``documents`` stand-in for a corpus that exists only as token ids in HBM: ``len`` and ``[i]`` (the text of doc i is
    regenerated on demand; benchmark result dictionaries then carry real contents without 10 M strings on the host)."""

    def __init__(self, spec: SynthSpec, materialise: bool = False):
        self.spec, self.materialise = spec, materialise
        self._th = None

    def __len__(self):
        return self.spec.n_docs

    def __getitem__(self, i):
        i = int(i)
        if not 0 <= i < self.spec.n_docs:
            raise IndexError(i)
        if not self.materialise:
            return f"<synthetic doc {i}>"
        if self._th is None:
            self._th = zipf_thresholds(self.spec.vocab, self.spec.zipf_s)
        return doc_texts(self.spec, i, i + 1, self._th)[0]

