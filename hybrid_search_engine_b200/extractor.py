"""Host-side tokeniser: defines term identity for the device CSR.

Same observable behaviour as the reference's ``extract_tokens`` / ``STOPWORDS`` / ``preprocess_text``
(extractor.py:6-52): lower-case first, ASCII ``[A-Za-z0-9_]+`` runs (so ``"naive cafe"`` with
accents splits at the non-ASCII letters), optional removal of the 48-word stop list, whitespace
collapse for stored ``content``.
"""
from __future__ import annotations

import re
from typing import List

_STOP_TEXT = """
a an the and or but in on at to for of with by from
is are was were be been being have has had do does did
will would could should may might must shall can
this that these those i you he she it we they
"""
STOPWORDS = frozenset(_STOP_TEXT.split())
assert len(STOPWORDS) == 48

_TOKEN = re.compile(r"[a-z0-9_]+")      # applied after lower(): same runs as [A-Za-z0-9_]+
_SPACE = re.compile(r"\s+")


def extract_tokens(text: str, remove_stopwords: bool = False) -> List[str]:
    if not text:
        return []
    found = _TOKEN.findall(text.lower())
    if not remove_stopwords:
        return found
    stop = STOPWORDS
    return [tok for tok in found if tok not in stop]


def preprocess_text(text: str, remove_stopwords: bool = False) -> str:
    if not text:
        return ""
    flat = _SPACE.sub(" ", text.strip())
    if remove_stopwords:
        return " ".join(extract_tokens(flat, remove_stopwords=True))
    return flat
