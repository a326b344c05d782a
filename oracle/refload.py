"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Imports ``search_engine.{bm25,utils,extractor,core,indexer,pipelines}`` straight from
``/root/reference`` (read-only, exists only in the dev container, never on the GPU box) so that
``oracle/make_golden.py`` can execute the reference's own code and freeze its outputs into
``tests/golden/``.  Nothing in the product package, in ``-m gpu`` tests or ``smoke()`` may import this
file; ``bench.py`` uses it for its CPU baseline legs only (``--impl reference`` / ``cpu_baseline_reference``),
where it loads the same unmodified modules from the staged archive ``oracle/_ref/search_engine_ref.zip``
(``python -m oracle.build_ref``) when ``/root/reference`` is absent, i.e. on the GPU box.

The reference package ``__init__`` imports every hard dependency (``search_engine/__init__.py:7``),
several of which are absent here (polars, duckdb, rapidfuzz, sentence_transformers).  We therefore
register an empty parent package whose ``__path__`` points at the reference directory and seed
``sys.modules`` with inert stand-ins that carry *no hot-path arithmetic*:

* ``polars.DataFrame``          -- column dict with ``__getitem__(col).to_list()`` and ``__len__``
                                   (all that ``core.py:240-241`` / ``indexer.py:219-227,277`` use)
* ``duckdb.connect``            -- no-op connection (``core.py:33-39``, ``indexer.py:100-201``)
* ``sentence_transformers``     -- ``SentenceTransformer.encode`` = lookup in an injected
                                   text -> vector table; ``CrossEncoder.predict`` = injected callable
* ``rapidfuzz.fuzz.partial_ratio`` -- swappable; inert wherever its weight is 0.0
                                   (``pipelines.py:322-323,479-480``).  For ``basic``/``diversity``
                                   it is the shared restatement in ``oracle/hybrid_oracle.py``
                                   ("parity unpinned").
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("HS_REFERENCE_ROOT", "/root/reference")
# staged copy of the same unmodified modules for boxes without /root/reference (oracle/build_ref.py)
STAGED_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "search_engine_ref.zip")

# text -> float32 vector table used by the SentenceTransformer stand-in
EMBED_TABLE: dict = {}
EMBED_DIM = [384]
# callable(pairs) -> scores for the CrossEncoder stand-in
CROSS_ENCODER_FN = [None]
# callable(a, b) -> float in [0, 100]
PARTIAL_RATIO_FN = [lambda a, b: 0.0]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "search_engine")) or os.path.exists(STAGED_ZIP)


def _package_dir() -> str:
    """Directory holding the reference's ``search_engine`` modules: the read-only tree in the dev container, else
    the staged archive unpacked into a fresh temporary directory (GPU box; used by bench.py's reference arm only)."""
    live = os.path.join(REFERENCE_ROOT, "search_engine")
    if os.path.isdir(live):
        return live
    import zipfile
    tmp = tempfile.mkdtemp(prefix="hs_ref_")
    with zipfile.ZipFile(STAGED_ZIP) as z:
        z.extractall(tmp)
    return os.path.join(tmp, "search_engine")


class _Series:
    def __init__(self, values):
        self._v = list(values)

    def to_list(self):
        return list(self._v)

    def __len__(self):
        return len(self._v)


class _DataFrame:
    def __init__(self, data):
        self._cols = {k: list(v) for k, v in data.items()}

    def __getitem__(self, col):
        return _Series(self._cols[col])

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    @property
    def columns(self):
        return list(self._cols)


class _Cursor:
    def fetchone(self):
        return (0,)

    def fetchall(self):
        return []


class _Con:
    def execute(self, *a, **k):
        return _Cursor()

    def register(self, *a, **k):
        pass

    def unregister(self, *a, **k):
        pass

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _SentenceTransformer:
    def __init__(self, name="stub", *a, **k):
        self.name = name

    def encode(self, texts, show_progress_bar=False, convert_to_numpy=True, **k):
        out = np.zeros((len(texts), EMBED_DIM[0]), dtype=np.float32)
        for i, t in enumerate(texts):
            if t not in EMBED_TABLE:
                raise KeyError(f"no synthetic embedding registered for text {t[:40]!r}")
            out[i] = EMBED_TABLE[t]
        return out


class _CrossEncoder:
    def __init__(self, name="stub", device=None, *a, **k):
        pass

    def predict(self, pairs, show_progress_bar=False, **k):
        if CROSS_ENCODER_FN[0] is None:
            raise RuntimeError("no cross-encoder function injected")
        return CROSS_ENCODER_FN[0](pairs)


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path modules (cached)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present (dev container) and no staged copy under oracle/_ref "
                           "(python -m oracle.build_ref)")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_ref_cache"))

    polars = types.ModuleType("polars")
    polars.DataFrame = _DataFrame
    duckdb = types.ModuleType("duckdb")
    duckdb.connect = lambda *a, **k: _Con()
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = _SentenceTransformer
    st.CrossEncoder = _CrossEncoder
    rapidfuzz = types.ModuleType("rapidfuzz")
    fuzz = types.ModuleType("rapidfuzz.fuzz")
    fuzz.partial_ratio = lambda a, b, **k: PARTIAL_RATIO_FN[0](a, b)
    fuzz.ratio = lambda a, b, **k: 0.0
    rapidfuzz.fuzz = fuzz
    process = types.ModuleType("rapidfuzz.process")
    rapidfuzz.process = process
    for name, mod in (("polars", polars), ("duckdb", duckdb), ("sentence_transformers", st),
                      ("rapidfuzz", rapidfuzz), ("rapidfuzz.fuzz", fuzz),
                      ("rapidfuzz.process", process)):
        sys.modules.setdefault(name, mod)

    pkg = types.ModuleType("search_engine")
    pkg.__path__ = [_package_dir()]
    sys.modules["search_engine"] = pkg
    try:
        from loguru import logger
        logger.remove()  # the reference logs INFO twice per query (core.py:235,284)
    except Exception:
        pass
    for sub in ("extractor", "utils", "bm25", "core", "indexer", "pipelines", "reranker"):
        _loaded[sub] = importlib.import_module(f"search_engine.{sub}")
    return types.SimpleNamespace(**_loaded)
