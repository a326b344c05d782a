"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference modules under ``oracle/_ref/``.

    python -m oracle.build_ref

``/root/reference`` exists only in the dev container.  ``oracle/_ref/`` is git-ignored (the reference's
sources never enter this repository's history) but NOT gpurun-ignored, so the staged copy travels to the
GPU box with the snapshot -- exactly like the built ``.so`` -- and lets ``bench.py --impl reference`` /
``cpu_baseline_reference`` time the reference's own ``create_pipeline("hybrid_bm25").search()`` there,
single-threaded as the reference is, through ``oracle/refload.py`` (inert stand-ins for polars / duckdb /
sentence_transformers / rapidfuzz, none of which carries hot-path arithmetic on this path).

The modules are stored byte for byte in ONE archive, ``oracle/_ref/search_engine_ref.zip`` (``refload`` unpacks
it into a temporary directory at run time); a manifest with their sha256 is written next to it so that a reader
can check the copy against the upstream tree.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/search_engine"
DST = os.path.join(HERE, "_ref", "search_engine_ref.zip")
# the modules oracle/refload.py imports, plus what pipelines.py imports at module level (pipelines.py:13-17)
FILES = ["extractor.py", "utils.py", "bm25.py", "core.py", "indexer.py", "pipelines.py", "reranker.py",
         "chunker.py", "highlighter.py"]


def build(verbose: bool = True) -> bool:
    """Copy the files; returns False (and does nothing) when the reference tree is not present."""
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle.build_ref: {SRC} not present -- keeping whatever is staged under oracle/_ref", file=sys.stderr)
        return False
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    manifest = {}
    with zipfile.ZipFile(DST, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            raw = open(os.path.join(SRC, f), "rb").read()
            z.writestr(zipfile.ZipInfo(f"search_engine/{f}", date_time=(2020, 1, 1, 0, 0, 0)), raw)
            manifest[f] = hashlib.sha256(raw).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle.build_ref: staged {len(FILES)} unmodified reference modules in {DST}")
    return True


if __name__ == "__main__":
    build()
