"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the gensim 4.3.3 pieces the hot path calls.

gensim==4.3.3 (reference requirements.txt:25, UTF-16 file) is NOT vendored under
/root/reference and is NOT installable here (no network), so the arithmetic at this
boundary is restated from gensim's published behaviour and
anchored on the reference's own call sites: webui.py:352 and webui.py:205 (``index[vec]``),
webui.py:670 (load) and genmodel.py:168-175 (build).  **Parity at the gensim boundary is
UNPINNED**: the reference ships no tests, fixtures or golden vectors for it.

What is restated:
  * ``matutils.unitvec`` on a gensim-sparse list: ``length = sqrt(sum(val**2))`` in Python
    floats over ALL entries; asserts length > 0; divides every value.
  * ``matutils.sparse2full(doc, length)``: ``zeros(length, float32)``; ``dict(doc)`` so the LAST
    duplicate id wins; values cast to float32.
  * ``Similarity.__getitem__`` / ``MatrixSimilarity.get_similarities``: per shard of
    ``shardsize = 32768`` rows ``numpy.dot(index_fp32, query_fp32)``, results ``hstack``-ed -> fp32[N].
    Stored rows are used as given (ndarray input is NOT normalised, genmodel.py:168-173).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

SHARD_ROWS = 32768


def unitvec_sparse(vec: Sequence[Tuple[int, float]]) -> List[Tuple[int, float]]:
    length = 1.0 * math.sqrt(sum(val ** 2 for _, val in vec))
    assert length > 0.0, "sparse documents must not contain any explicit zero entries"
    if length != 1.0:
        return [(termid, val / length) for termid, val in vec]
    return list(vec)


def sparse2full(doc: Sequence[Tuple[int, float]], length: int) -> np.ndarray:
    result = np.zeros(length, dtype=np.float32)
    doc = ((int(id_), float(val_)) for (id_, val_) in doc)
    doc = dict(doc)
    result[list(doc)] = list(doc.values())
    return result


def dense_query(vec: Sequence[Tuple[int, float]], dim: int) -> np.ndarray:
    """What the engine must be handed for ``index[vec]``: fp32[dim] = sparse2full(unitvec(vec))."""
    return sparse2full(unitvec_sparse(vec), dim)


class SimilarityStub:
    """``index`` object of webui.py:27 as far as the hot path uses it: ``index[vec] -> fp32[N]``."""

    def __init__(self, rows: np.ndarray, shard_rows: int = SHARD_ROWS):
        assert rows.dtype == np.float32 and rows.ndim == 2
        self.rows = rows
        self.shard_rows = shard_rows
        self.num_features = rows.shape[1]
        self.queries_seen: List[np.ndarray] = []   # instrumentation for tests

    def __len__(self) -> int:
        return self.rows.shape[0]

    def __getitem__(self, vec: Sequence[Tuple[int, float]]) -> np.ndarray:
        q = dense_query(vec, self.num_features)
        self.queries_seen.append(q)
        parts = []
        for lo in range(0, self.rows.shape[0], self.shard_rows):
            parts.append(np.dot(self.rows[lo: lo + self.shard_rows], q.T).T)
        if not parts:
            return np.zeros(0, dtype=np.float32)
        return np.hstack(parts)
