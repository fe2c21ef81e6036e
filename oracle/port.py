"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the reference's query-scoring path.

Every function cites the reference lines it follows (paths under /root/reference).  The
restatement is PINNED against the reference itself: oracle/verbatim.py executes the
reference's own functions (AST-extracted, unchanged) in this container, oracle/make_golden.py
records their outputs under tests/golden/, and tests/test_oracle_golden.py checks this port
against those fixtures bit-for-bit.  The gensim piece (``index[vec]``) follows
oracle/gensim_stub.py and is UNPINNED (gensim 4.3.3 is absent; see that file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg may
import this module - as the checker or the timed CPU baseline, never as the shipped path.

Two flavours of the O(N) loops:
  * ``faithful=True``  - the reference's own data structures and loop shapes (a Python list of
    per-doc dicts, ``dict.get`` per doc per term, Python ``sorted`` over N tuples).  This is what
    bench.py times as the CPU baseline (kind "port").
  * ``faithful=False`` - vectorised numpy over CSR/posting arrays with the SAME per-element
    arithmetic (same operations, same order, same dtypes), so results are bit-identical; used
    to check the GPU at sizes where the faithful loops take minutes.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .gensim_stub import SimilarityStub, dense_query

# webui.py:51-60
BM25_WEIGHT = 0.5
DOC2VEC_WEIGHT = 0.5
ORIGINAL_SCORE_WEIGHT = 0.7
RERANKED_SCORE_WEIGHT = 0.3
DIFF_FILTER_THRESH = 1e-6
REQUIRE_TAG_MAGIC_NUMBER = 1000
K1 = 1.5    # webui.py:126
B = 0.75    # webui.py:127
PRF_DEPTH = 10  # webui.py:193-195


def _weight_suffix(last: str) -> bool:
    # webui.py:89 / webui.py:360 - the token's last ':' field is a weight
    return last.startswith("+") or last.startswith("-") or last.isdigit()


def parse_query_weights(new_doc: str, token2id: Dict[str, int], magic: int = REQUIRE_TAG_MAGIC_NUMBER):
    """webui.py:354-371.  Returns ({term_id: weight} in insertion order, required_tags, exclude_tags).
    Unknown tag / empty token -> KeyError; 'tag:+' -> ValueError (int('+'))."""
    weights: Dict[int, float] = {}
    required: List[str] = []
    exclude: List[str] = []
    for term in new_doc.split(" "):
        parts = term.split(":")
        if len(parts) >= 2 and _weight_suffix(parts[-1]):
            tag = ":".join(parts[:-1])
            if parts[-1].startswith("+"):
                weights[token2id[tag]] = magic + int(parts[-1])
                required.append(tag)
            else:
                weights[token2id[tag]] = int(parts[-1])
                exclude.append(tag)
        else:
            weights[token2id[":".join(parts)]] = 1
    return weights, required, exclude


def parse_vector_terms(new_doc: str) -> Tuple[List[Tuple[str, int]], int]:
    """webui.py:83-102.  [(paren-escaped tag text, int weight)], sum of weights (0 -> 1)."""
    out: List[Tuple[str, int]] = []
    total = 0
    for tag in new_doc.split(" "):
        parts = tag.split(":")
        if len(parts) >= 2 and _weight_suffix(parts[-1]):
            text = ":".join(parts[:-1]).replace("\\(", "(").replace("\\)", ")")
            w = int(parts[-1])
        else:
            text = ":".join(parts).replace("\\(", "(").replace("\\)", ")")
            w = 1
        out.append((text.replace("(", "\\(").replace(")", "\\)"), w))
        total += w
    if total == 0:
        total = 1
    return out, total


def filter_searched_result(sorted_scores: Sequence[Tuple[int, float]], thresh: float = DIFF_FILTER_THRESH):
    """webui.py:63-80."""
    s = np.array([p[1] for p in sorted_scores])
    diff = s[:-1] - s[1:]
    diff = np.where(diff == 0, np.inf, diff)
    t = len(sorted_scores)
    found = np.where(diff < thresh)[0]
    if len(found) == 1:
        t = found[0]
    elif len(found) >= 2:
        t = found[1]
    max_val = s.max()
    return [(sorted_scores[i][0], sorted_scores[i][1] / float(max_val)) for i in range(int(t))
            if sorted_scores[i][1] > 0]


class OraclePort:
    def __init__(self, idx, faithful: bool = False, infer: Optional[Callable[[List[str]], np.ndarray]] = None):
        """idx: a synth.SynthIndex-like object (row_ptr/term_ids/tfs/doc_len/avgdl/idf/rows/tag_names)."""
        self.idx = idx
        self.faithful = faithful
        self.n = idx.n_docs
        self.token2id = idx.token2id
        self.index = SimilarityStub(idx.rows)
        self.dim = idx.rows.shape[1]
        self.doc_len = idx.doc_len
        self.avgdl = idx.avgdl
        self.bm25_D = idx.n_docs
        self.idf_dict = idx.bm25_idf_dict()
        self.consts = dict(BM25_WEIGHT=BM25_WEIGHT, DOC2VEC_WEIGHT=DOC2VEC_WEIGHT,
                           ORIGINAL_SCORE_WEIGHT=ORIGINAL_SCORE_WEIGHT, RERANKED_SCORE_WEIGHT=RERANKED_SCORE_WEIGHT,
                           DIFF_FILTER_THRESH=DIFF_FILTER_THRESH, REQUIRE_TAG_MAGIC_NUMBER=REQUIRE_TAG_MAGIC_NUMBER)
        if infer is None:
            t2i = self.token2id
            infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
        self.infer_vector = infer
        if faithful:
            self.corpus = idx.bm25_corpus()
            self.csv = idx.csv_lines()
        else:
            self.post_ptr, self.post_doc, self.post_tf = idx.postings()

    # ---- webui.py:82-117 ----------------------------------------------------
    def query_vector(self, new_doc: str) -> List[Tuple[int, float]]:
        terms, total = parse_vector_terms(new_doc)
        got = np.zeros(self.dim)
        for tag, w in terms:
            v = self.infer_vector([tag])
            v = v / np.linalg.norm(v)
            got += w * v
        got = got / total
        norm = np.linalg.norm(got)
        if math.isinf(norm) or norm == 0:
            norm = 1.0
        got = got / norm
        return [(i, val) for i, val in enumerate(got)]

    # ---- webui.py:119-172 ---------------------------------------------------
    def bm25_scores(self, query_weights: Dict[int, float]) -> np.ndarray:
        magic = self.consts["REQUIRE_TAG_MAGIC_NUMBER"]
        scores = np.zeros(self.bm25_D)
        dl = self.doc_len
        for term_id in list(query_weights.keys()):
            idf = self.idf_dict.get(term_id, 0)
            weight = query_weights.get(term_id, 1.0)
            if self.faithful:
                tfs = np.array([doc.get(term_id, 0) for doc in self.corpus])
                denom = tfs + K1 * (1 - B + B * (dl / self.avgdl))
                score = idf * ((tfs * (K1 + 1)) / denom)
                if weight < 0:
                    hit = [i for i, doc in enumerate(self.corpus) if term_id in doc]
                    scores[hit] = -np.inf
                elif weight > magic:
                    miss = [i for i, doc in enumerate(self.corpus) if term_id not in doc]
                    scores += (weight - magic) * score
                    scores[miss] = -np.inf
                else:
                    scores += weight * score
                continue
            # vectorised: only docs in the posting list have tf != 0; for every other doc the
            # reference adds weight * (idf * (0 / denom)) = 0, which leaves the sum unchanged.
            if 0 <= term_id < len(self.post_ptr) - 1:
                a, b = int(self.post_ptr[term_id]), int(self.post_ptr[term_id + 1])
            else:
                a = b = 0
            docs = self.post_doc[a:b]
            tfs = self.post_tf[a:b].astype(np.int64)
            denom = tfs + K1 * (1 - B + B * (dl[docs] / self.avgdl))
            score = idf * ((tfs * (K1 + 1)) / denom)
            if weight < 0:
                scores[docs] = -np.inf
            elif weight > magic:
                scores[docs] += (weight - magic) * score
                miss = np.ones(self.bm25_D, dtype=bool)
                miss[docs] = False
                scores[miss] = -np.inf
            else:
                scores[docs] += weight * score
        return scores

    # ---- webui.py:376-383 ---------------------------------------------------
    def combine(self, sims: np.ndarray, bm25: np.ndarray) -> np.ndarray:
        if sims.max() > 0:
            sims = sims / sims.max()
        if bm25.max() > 0:
            bm25 = bm25 / bm25.max()
        return self.consts["BM25_WEIGHT"] * bm25 + self.consts["DOC2VEC_WEIGHT"] * sims

    # ---- webui.py:182-187 ---------------------------------------------------
    def doc_vector_pairs(self, doc_id_1based: int) -> List[Tuple[int, float]]:
        d = doc_id_1based - 1
        if self.faithful:
            tags = self.csv[d].split(",")[1:]
        else:
            tags = [self.idx.tag_names[t] for t in self.idx.doc_tags(d)]
        v = self.infer_vector(tags)
        return [(i, val) for i, val in enumerate(v)]

    # ---- webui.py:195-203 ---------------------------------------------------
    @staticmethod
    def prf_query(top_vectors: List[List[Tuple[int, float]]], weights: List[float]) -> List[Tuple[int, float]]:
        mean = np.average(top_vectors, axis=0, weights=weights)
        mean = mean / np.linalg.norm(mean)
        return [(round(docid), val) for docid, val in mean.tolist()]

    # ---- webui.py:189-253 ---------------------------------------------------
    def rerank_sorted(self, final_scores: np.ndarray) -> List[Tuple[int, float]]:
        """The sorted (doc id, score) list of webui.py:189-237 right BEFORE filter_searched_result."""
        n = len(final_scores)
        if self.faithful:
            sims = sorted(list(enumerate(final_scores)), key=lambda it: -it[1])
        else:
            order = np.argsort(-final_scores, kind="stable")
            sims = None
        if n > PRF_DEPTH:
            if self.faithful:
                top = sims[:PRF_DEPTH]
            else:
                top = [(int(d), final_scores[d]) for d in order[:PRF_DEPTH]]
            top_ids = [d for d, _ in top]
            vecs = [self.doc_vector_pairs(d + 1) for d in top_ids]
            q2 = self.prf_query(vecs, [s for _, s in top])
            rer = self.index[q2]
            R = self.consts["ORIGINAL_SCORE_WEIGHT"] * final_scores + self.consts["RERANKED_SCORE_WEIGHT"] * rer
            if R.max() > 0:
                R = R / R.max()
            head = [(d, 1.0) for d in top_ids]
            if self.faithful:
                top_set = set(top_ids)
                rest = [it for it in enumerate(R) if it[0] not in top_set]
                rest = sorted(rest, key=lambda it: -it[1])
            else:
                ro = np.argsort(-R, kind="stable")
                ro = ro[~np.isin(ro, np.asarray(top_ids))]
                rest = list(zip(ro.tolist(), R[ro]))
            return head + rest
        if not self.faithful:
            sims = list(zip(order.tolist(), final_scores[order]))
        return sims

    def rerank(self, final_scores: np.ndarray, topn: int) -> List[Tuple[int, float]]:
        out = filter_searched_result(self.rerank_sorted(final_scores), self.consts["DIFF_FILTER_THRESH"])
        return out[: min(topn, len(out))]

    def find_sorted(self, new_doc: str) -> List[Tuple[int, float]]:
        """find_similar_documents up to (not including) the filter: for tolerance analysis in tests."""
        vec = self.query_vector(new_doc)
        sims = self.index[vec]
        weights, _, _ = parse_query_weights(new_doc, self.token2id, self.consts["REQUIRE_TAG_MAGIC_NUMBER"])
        return self.rerank_sorted(self.combine(sims, self.bm25_scores(weights)))

    # ---- webui.py:345-390 ---------------------------------------------------
    def find_similar_documents(self, new_doc: str, topn: int = 50) -> List[Tuple[int, float]]:
        vec = self.query_vector(new_doc)
        sims = self.index[vec]
        weights, _, _ = parse_query_weights(new_doc, self.token2id, self.consts["REQUIRE_TAG_MAGIC_NUMBER"])
        bm25 = self.bm25_scores(weights)
        final = self.combine(sims, bm25)
        return self.rerank(final, topn)

    # ---- the same path without N-long Python lists (for >= 1 M docs) -------------
    def rerank_arrays(self, final_scores: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """rerank_sorted (webui.py:189-237) as (doc ids, scores) ARRAYS: same arithmetic, same stable sorts, no tuples."""
        n = len(final_scores)
        order = np.argsort(-final_scores, kind="stable")
        if n <= PRF_DEPTH:
            return order, final_scores[order]
        top_ids = order[:PRF_DEPTH]
        vecs = [self.doc_vector_pairs(int(d) + 1) for d in top_ids]
        q2 = self.prf_query(vecs, [final_scores[d] for d in top_ids])
        rer = self.index[q2]
        R = self.consts["ORIGINAL_SCORE_WEIGHT"] * final_scores + self.consts["RERANKED_SCORE_WEIGHT"] * rer
        if R.max() > 0:
            R = R / R.max()
        ro = np.argsort(-R, kind="stable")
        ro = ro[~np.isin(ro, top_ids)]
        return np.concatenate([top_ids, ro]), np.concatenate([np.ones(PRF_DEPTH), R[ro]])

    @staticmethod
    def filter_arrays(ids: np.ndarray, s: np.ndarray, thresh: float, topn: int) -> List[Tuple[int, float]]:
        """filter_searched_result (webui.py:63-80) + [:topn] (webui.py:243-246) on arrays."""
        diff = s[:-1] - s[1:]
        diff = np.where(diff == 0, np.inf, diff)
        t = len(s)
        found = np.where(diff < thresh)[0]
        if len(found) == 1:
            t = int(found[0])
        elif len(found) >= 2:
            t = int(found[1])
        max_val = float(s.max())
        keep = np.nonzero(s[:t] > 0)[0][:topn]           # `if score > 0` over range(t), then [:topn]: order kept
        return [(int(ids[i]), s[i] / max_val) for i in keep]

    def find_fast(self, new_doc: str, topn: int = 50) -> List[Tuple[int, float]]:
        """find_similar_documents(new_doc, topn), bit-identical to the list-based path above (tests/test_oracle_golden.py)."""
        vec = self.query_vector(new_doc)
        sims = self.index[vec]
        weights, _, _ = parse_query_weights(new_doc, self.token2id, self.consts["REQUIRE_TAG_MAGIC_NUMBER"])
        ids, s = self.rerank_arrays(self.combine(sims, self.bm25_scores(weights)))
        return self.filter_arrays(ids, s, self.consts["DIFF_FILTER_THRESH"], topn)

    def find_sorted_arrays(self, new_doc: str) -> Tuple[np.ndarray, np.ndarray]:
        vec = self.query_vector(new_doc)
        sims = self.index[vec]
        weights, _, _ = parse_query_weights(new_doc, self.token2id, self.consts["REQUIRE_TAG_MAGIC_NUMBER"])
        return self.rerank_arrays(self.combine(sims, self.bm25_scores(weights)))

    # ---- intermediate products, for kernel-level parity tests -----------------
    def stages(self, new_doc: str) -> Dict[str, np.ndarray]:
        vec = self.query_vector(new_doc)
        q = dense_query(vec, self.dim)
        sims = self.index[vec]
        weights, _, _ = parse_query_weights(new_doc, self.token2id, self.consts["REQUIRE_TAG_MAGIC_NUMBER"])
        bm25 = self.bm25_scores(weights)
        final = self.combine(sims, bm25)
        return {"q": q, "sims": sims, "bm25": bm25, "final": final, "weights": weights}
