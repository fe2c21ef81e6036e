"""Host-side query preparation: the parts of the path that stay Python (north_star: gensim query
inference and the query grammar are untouched).  Output of this module is what crosses to the GPU:
a dense fp32 unit query vector and the {term id: weight} table.

Reference lines: query grammar webui.py:354-371, query vector webui.py:82-117, PRF centroid
webui.py:195-203, gensim ``unitvec`` / ``sparse2full`` as invoked by ``index[vec]`` (webui.py:352,205;
gensim 4.3.3 behaviour per SURVEY.md Appendix B).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from .engine import DIM, Query

SparseVec = List[Tuple[int, float]]


def _has_weight(last: str) -> bool:
    return last[:1] in ("+", "-") or last.isdigit()


def split_token(token: str) -> Tuple[str, str]:
    """'tag:with:colons:+2' -> ('tag:with:colons', '+2'); 'tag' -> ('tag', '')."""
    fields = token.split(":")
    if len(fields) >= 2 and _has_weight(fields[-1]):
        return ":".join(fields[:-1]), fields[-1]
    return token, ""


def parse_weights(new_doc: str, token2id: Dict[str, int], magic: int = 1000) -> Dict[int, int]:
    """BM25 side of the grammar.  '+W' -> magic + W (required), '-W' -> negative (exclude), 'W' -> W,
    bare tag -> 1; a repeated tag keeps its LAST weight; unknown tag / empty token -> KeyError,
    'tag:+' -> ValueError - exactly what the reference raises."""
    table: Dict[int, int] = {}
    for token in new_doc.split(" "):
        tag, w = split_token(token)
        if w == "":
            table[token2id[tag]] = 1
        elif w[0] == "+":
            table[token2id[tag]] = magic + int(w)
        else:
            table[token2id[tag]] = int(w)
    return table


def _canon_parens(tag: str) -> str:
    plain = tag.replace("\\(", "(").replace("\\)", ")")
    return plain.replace("(", "\\(").replace(")", "\\)")


def query_vector(new_doc: str, infer_vector: Callable[[List[str]], np.ndarray], dim: int = DIM) -> SparseVec:
    """Doc2Vec side: weighted sum of the unit vectors of every token's tag (every occurrence counts,
    '-W' subtracts), divided by the weight sum (0 -> 1; a negative sum flips the direction) and
    re-normalised (norm 0 / inf -> 1)."""
    acc = np.zeros(dim)
    total = 0
    for token in new_doc.split(" "):
        tag, w = split_token(token)
        weight = int(w) if w else 1
        total += weight
        v = infer_vector([_canon_parens(tag)])
        v = v / np.linalg.norm(v)
        acc += weight * v
    acc = acc / (total if total != 0 else 1)
    norm = np.linalg.norm(acc)
    if math.isinf(norm) or norm == 0:
        norm = 1.0
    acc = acc / norm
    return [(i, val) for i, val in enumerate(acc)]


def dense_query(vec: Sequence[Tuple[int, float]], dim: int = DIM) -> np.ndarray:
    """What gensim hands to numpy.dot for ``index[vec]``: unitvec over all list entries (python floats,
    asserts a positive length), then sparse2full where the LAST duplicate id wins, cast to fp32."""
    length = 1.0 * math.sqrt(sum(val ** 2 for _, val in vec))
    assert length > 0.0, "sparse documents must not contain any explicit zero entries"
    if length != 1.0:
        vec = [(i, val / length) for i, val in vec]
    slots = dict((int(i), float(val)) for i, val in vec)
    out = np.zeros(dim, dtype=np.float32)
    out[list(slots)] = list(slots.values())
    return out


def prf_query(top_vectors: Sequence[SparseVec], weights: Sequence[float]) -> SparseVec:
    """The re-query 'vector' exactly as webui.py:200-203 builds it, index column and all (the ids
    collapse to 0 after the Frobenius normalisation + round())."""
    mean = np.average(top_vectors, axis=0, weights=list(weights))
    mean = mean / np.linalg.norm(mean)
    return [(round(i), val) for i, val in mean.tolist()]


def make_query(new_doc: str, token2id: Dict[str, int], infer_vector, dim: int = DIM, magic: int = 1000) -> Query:
    """Everything find_similar_documents does on the host before the O(N) work (webui.py:349-371)."""
    vec = dense_query(query_vector(new_doc, infer_vector, dim), dim)
    table = parse_weights(new_doc, token2id, magic)
    return Query(vec, np.fromiter(table.keys(), dtype=np.int32, count=len(table)),
                 np.fromiter(table.values(), dtype=np.float64, count=len(table)))
