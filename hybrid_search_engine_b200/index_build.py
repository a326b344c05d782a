"""Inverted-index construction on the device for real text (the step before the hot path).

``BM25.fit`` in the reference tokenises every document with a regex and counts terms in Python dicts
(bm25.py:58-67, ~214 us/doc).  Here the host only lower-cases and concatenates the documents; tokenising,
term hashing, stop-word removal, doc lengths, document frequencies and the (term, doc, tf) CSR are built on
the GPU (``csrc/tokenize.cu`` + device-wide sort / unique / prefix sums -- index construction uses torch for
those primitives as plumbing).  Term identity is a 63-bit hash of the token, so no vocabulary strings exist:
queries are hashed the same way and looked up in the sorted hash table.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib, devsort
from ._lib import check, ptr, stream_ptr
from .extractor import STOPWORDS, extract_tokens

_M64 = (1 << 64) - 1
_SEP = "\x1e"


def token_hash(tok: str) -> int:
    """Host twin of ``token_hash_kernel``: FNV-1a over the (ASCII, lower-case) token, splitmix64 finaliser."""
    h = 0xCBF29CE484222325
    data = tok.encode("ascii")
    for c in data:
        h = ((h ^ c) * 0x100000001B3) & _M64
    z = ((h ^ (len(data) << 48)) + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return (z ^ (z >> 31)) >> 1


STOP_HASHES = np.array(sorted(token_hash(w) for w in STOPWORDS), dtype=np.int64)


class DeviceLexicalStats:
    """Same statistics as ``index.LexicalStats`` (bm25.py:45-81), produced on the device.

    Attributes mirror what ``DeviceIndex.set_bm25`` needs: ``indptr``, ``postings`` (device tensors),
    ``doc_lengths`` (device int32), ``df`` (host int64), ``avg_doc_len``, ``doc_count`` and the sorted
    ``vocab_hashes`` (host int64) that define term ids.
    """

    def __init__(self, device, remove_stopwords: bool = True):
        self.device = torch.device(device)
        self.remove_stopwords = remove_stopwords
        self.lib = _lib.load()
        self._documents = None
        self._vocab = None

    def fit(self, documents: Sequence[str], chunk_bytes: int = 1 << 28):
        dev, lib = self.device, self.lib
        n = len(documents)
        self.doc_count = n
        self._documents, self._vocab = documents, None      # only for the lazy `vocab` property
        hashes: List[torch.Tensor] = []
        docs_of: List[torch.Tensor] = []
        start = 0
        with torch.cuda.device(dev):
            while start < n:
                # ---- host: concatenate one chunk.  ASCII documents go up as they are (the kernel folds A-Z);
                #      only documents with non-ASCII characters pay for str.lower().  Documents are joined by
                #      U+001E, a byte that is never part of a token; the device finds the boundaries.
                end, total = start, 0
                while end < n and (total < chunk_bytes or end == start):
                    total += len(documents[end]) + 1
                    end += 1
                chunk = [x if x.isascii() else x.lower() for x in documents[start:end]]
                blob = (_SEP.join(chunk) + _SEP).encode("utf-8")
                text = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
                doc_end = (text == 0x1E).nonzero().flatten()
                if doc_end.numel() != end - start:          # a document contains U+001E itself: blank it out
                    chunk = [x.replace(_SEP, " ") for x in chunk]
                    blob = (_SEP.join(chunk) + _SEP).encode("utf-8")
                    text = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
                    doc_end = (text == 0x1E).nonzero().flatten()
                # ---- device: token starts, hashes, owning doc
                flags = torch.empty(len(blob), dtype=torch.uint8, device=dev)
                st = stream_ptr(dev)
                check(lib.hs_token_flags(ptr(text), len(blob), ptr(flags), st), "hs_token_flags")
                pos = flags.nonzero().flatten()
                h = torch.empty(pos.numel(), dtype=torch.int64, device=dev)
                check(lib.hs_token_hashes(ptr(text), len(blob), ptr(pos), pos.numel(), ptr(h), st), "hs_token_hashes")
                d = torch.bucketize(pos, doc_end, right=True) + start
                hashes.append(h)
                docs_of.append(d)
                start = end
            h = torch.cat(hashes) if hashes else torch.zeros(0, dtype=torch.int64, device=dev)
            d = torch.cat(docs_of) if docs_of else torch.zeros(0, dtype=torch.int64, device=dev)
            # token-level passes: the hand-written primitives of csrc/sort.cu (radix sort, binary search, run-length
            # encoding, boundary counts, prefix sum)
            if self.remove_stopwords and h.numel():
                stop = torch.from_numpy(np.sort(STOP_HASHES)).to(dev)
                at = devsort.lower_bound(stop, h).clamp_(max=stop.numel() - 1)
                keep = stop[at] != h
                h, d = h[keep], d[keep]
            # ---- statistics (bm25.py:59-71): dl counts tokens after stop-word removal, duplicates included.  Tokens
            #      are in document order, so d << 32 is a sorted key column and dl is its per-"term" count
            dl = devsort.term_doc_freqs(d << 32, n) if n else torch.zeros(0, dtype=torch.int64, device=dev)
            self.doc_lengths = dl.to(torch.int32)
            self.avg_doc_len = (int(dl.sum().item()) / n) if n > 0 else 0
            if h.numel():
                hs_sorted = devsort.sort_keys_(h.clone())                     # vocabulary = distinct hashes, ascending
                uniq, _ = devsort.run_length_encode(hs_sorted)
                del hs_sorted
                inv = devsort.lower_bound(uniq, h)                            # term id of every token
            else:
                uniq, inv = h, h
            V = int(uniq.numel())
            keys = ((inv << 32) | d).contiguous()
            if keys.numel():
                devsort.sort_keys_(keys, devsort.byte_mask_for((0, max(n - 1, 1)), (32, max(V - 1, 1))))
                uk, tf = devsort.run_length_encode(keys)
            else:
                uk, tf = keys, keys
            df = devsort.term_doc_freqs(uk, V) if V else torch.zeros(0, dtype=torch.int64, device=dev)
            self.indptr = devsort.exclusive_scan(df) if V else torch.zeros(1, dtype=torch.int64, device=dev)
            if uk.numel():
                self.postings = torch.stack([(uk & 0xFFFFFFFF).to(torch.int32), tf.to(torch.int32)], dim=1)
            else:
                self.postings = torch.zeros((0, 2), dtype=torch.int32, device=dev)
            self.df = df.cpu().numpy()
            self.vocab_hashes = uniq.cpu().numpy()
            self.max_dl = int(dl.max().item()) if n else 0
        return self

    @property
    def vocab(self):
        """term string -> term id, as ``LexicalStats.vocab`` -- built on first use by tokenising the documents on
        the host (slow; the device build itself never needs the strings).  Also the place where a collision of the
        63-bit term hash would surface: two different strings with one id raise."""
        if self._vocab is None:
            vh, out, owner = self.vocab_hashes, {}, {}
            for doc in self._documents or ():
                for t in extract_tokens(doc, remove_stopwords=self.remove_stopwords):
                    if t in out:
                        continue
                    i = int(np.searchsorted(vh, token_hash(t)))
                    if owner.setdefault(i, t) != t:
                        raise RuntimeError(f"term hash collision between {owner[i]!r} and {t!r}")
                    out[t] = i
            self._vocab = out
        return self._vocab

    def query_term_ids(self, query: str) -> List[int]:
        """bm25.py:94,99-101 -- query tokens in order, duplicates kept, unknown terms dropped."""
        out = []
        vh = self.vocab_hashes
        for t in extract_tokens(query, remove_stopwords=self.remove_stopwords):
            hv = token_hash(t)
            i = int(np.searchsorted(vh, hv))
            if i < len(vh) and vh[i] == hv:
                out.append(i)
        return out
