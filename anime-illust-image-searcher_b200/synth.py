"""Synthetic tagged-image index + query generator.

Produces, for N docs over a V-tag vocabulary, exactly the pieces the reference's
index files hold (reference formats: genmodel.py:51-99 for the BM25 side,
genmodel.py:168-175 for the doc-vector side, genmodel.py:21-43 for the csv):

* doc x tag incidence as CSR (``row_ptr int64[N+1]``, ``term_ids int32[nnz]``,
  ``tfs int32[nnz]``), doc order = csv line order;
* ``doc_len int64[N]`` (= sum of tf), ``avgdl`` (np.float64 mean), ``idf``
  (float64[V], ``ln(1 + (D - df + 0.5)/(df + 0.5))``, 0 where df == 0);
* ``rows fp32[N x 300]`` = the stand-in for ``Doc2Vec.infer_vector(tags)`` applied
  to every doc, stored RAW (not normalised) like gensim stores ndarray input;
* tag names / token2id, csv lines ``path,tag,tag,...``.

The ``infer_vector`` stand-in is a pure deterministic function of the tag-id
sequence so that the reference's PRF step (which RE-infers the top-10 docs,
webui.py:182-187) can be reproduced by the oracle, by the host callback and by
the device "stored rows" mode alike.

Everything here is numpy; `synth_torch.py` holds the on-device generator used
for the >= 10 M-doc bench configurations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

DIM = 300
DEFAULT_VOCAB = 10861

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, vectorised over uint64 arrays (wraps silently)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        x = x ^ (x >> np.uint64(31))
    return x


def _u01(h: np.ndarray) -> np.ndarray:
    """uint64 hash -> float64 in (0, 1)."""
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _hash_gauss(h_doc: np.ndarray, n_cols: int) -> np.ndarray:
    """Deterministic N(0,1) matrix [len(h_doc), n_cols] from per-doc hashes (Box-Muller)."""
    cols = np.arange(n_cols, dtype=np.uint64)[None, :]
    base = h_doc[:, None]
    with np.errstate(over="ignore"):
        h1 = _mix64(base ^ (cols * np.uint64(2) + np.uint64(1)) * np.uint64(0xD6E8FEB86659FD93))
        h2 = _mix64(base ^ (cols * np.uint64(2) + np.uint64(2)) * np.uint64(0xA0761D6478BD642F))
    u1 = _u01(h1)
    u2 = _u01(h2)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def tag_sequence_hash(padded: np.ndarray, lengths: np.ndarray, seed: int) -> np.ndarray:
    """Order-dependent fold of each row's first `lengths[i]` tag ids into a uint64."""
    n, width = padded.shape
    h = np.full(n, np.uint64(seed) ^ np.uint64(0x5851F42D4C957F2D), dtype=np.uint64)
    for k in range(width):
        live = lengths > k
        if not live.any():
            break
        hk = _mix64(h ^ (padded[:, k].astype(np.uint64) + np.uint64(1)))
        h = np.where(live, hk, h)
    return h


class InferVectorStub:
    """Deterministic stand-in for gensim ``Doc2Vec.infer_vector`` (webui.py:106,185).

    f(tags) = s * (mean_k E[tag_k] + 0.3 * xi(hash(tags))),  s = exp(0.25 * g(hash)),
    evaluated in float32 exactly the same way for one doc or a million.
    """

    def __init__(self, vocab_size: int, seed: int, dim: int = DIM):
        self.vocab_size = vocab_size
        self.seed = int(seed)
        self.dim = dim
        rng = np.random.default_rng(self.seed ^ 0xE3B)
        self.E = rng.standard_normal((vocab_size, dim), dtype=np.float32)

    def batch(self, padded: np.ndarray, lengths: np.ndarray) -> np.ndarray:
        """padded int32 [n, width] (entries past lengths[i] ignored) -> fp32 [n, dim]."""
        n, width = padded.shape
        acc = np.zeros((n, self.dim), dtype=np.float32)
        for k in range(width):
            live = lengths > k
            if not live.any():
                break
            idx = np.where(live, padded[:, k], 0)
            acc += self.E[idx] * live[:, None].astype(np.float32)
        denom = np.maximum(lengths, 1).astype(np.float32)[:, None]
        mean = acc / denom
        h = tag_sequence_hash(padded, lengths, self.seed)
        g = _hash_gauss(h, self.dim + 1)
        xi = g[:, : self.dim].astype(np.float32)
        s = np.exp(0.25 * g[:, self.dim]).astype(np.float32)[:, None]
        return (s * (mean + np.float32(0.3) * xi)).astype(np.float32)

    def one(self, tag_ids: Sequence[int]) -> np.ndarray:
        ids = np.asarray(list(tag_ids), dtype=np.int32)[None, :]
        if ids.shape[1] == 0:
            ids = np.zeros((1, 1), dtype=np.int32)
            return self.batch(ids, np.array([0]))[0]
        return self.batch(ids, np.array([ids.shape[1]]))[0]


@dataclass
class SynthIndex:
    n_docs: int
    vocab_size: int
    seed: int
    row_ptr: np.ndarray          # int64 [N+1]
    term_ids: np.ndarray         # int32 [nnz]   (doc-major, in csv tag order)
    tfs: np.ndarray              # int32 [nnz]
    doc_len: np.ndarray          # int64 [N]
    avgdl: np.float64
    idf: np.ndarray              # float64 [V], 0 where df == 0
    df: np.ndarray               # int64 [V]
    rows: np.ndarray             # fp32 [N, 300] raw stored vectors
    tag_names: List[str]
    infer: InferVectorStub
    doc_tag_seq: Optional[List[np.ndarray]] = None   # per doc: tag-id sequence incl. repeats (csv order)
    popularity: np.ndarray = field(default=None)     # float64 [V] sampling weights

    @property
    def token2id(self) -> Dict[str, int]:
        return {t: i for i, t in enumerate(self.tag_names)}

    def csv_lines(self) -> List[str]:
        out = []
        for d in range(self.n_docs):
            seq = self.doc_tags(d)
            out.append(",".join(["img/%07d.png" % d] + [self.tag_names[t] for t in seq]))
        return out

    def doc_tags(self, d: int) -> np.ndarray:
        if self.doc_tag_seq is not None:
            return self.doc_tag_seq[d]
        a, b = int(self.row_ptr[d]), int(self.row_ptr[d + 1])
        return np.repeat(self.term_ids[a:b], self.tfs[a:b])

    def bm25_corpus(self) -> List[Dict[int, int]]:
        """The reference's in-memory form (genmodel.py:64-68): one {term_id: tf} per doc."""
        out: List[Dict[int, int]] = []
        tid = self.term_ids.tolist()
        tf = self.tfs.tolist()
        rp = self.row_ptr.tolist()
        for d in range(self.n_docs):
            out.append({tid[p]: tf[p] for p in range(rp[d], rp[d + 1])})
        return out

    def bm25_idf_dict(self) -> Dict[int, np.float64]:
        return {int(t): np.float64(self.idf[t]) for t in np.nonzero(self.df)[0]}

    def postings(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Tag-major posting lists: (post_ptr int64[V+1], post_doc int32[nnz], post_tf int32[nnz]),
        doc ids ascending inside each list."""
        return csr_to_postings(self.row_ptr, self.term_ids, self.tfs, self.vocab_size)


def csr_to_postings(row_ptr: np.ndarray, term_ids: np.ndarray, tfs: np.ndarray, vocab_size: int):
    n = len(row_ptr) - 1
    counts = np.diff(row_ptr)
    docs = np.repeat(np.arange(n, dtype=np.int32), counts)
    order = np.argsort(term_ids, kind="stable")
    post_doc = docs[order]
    post_tf = tfs[order].astype(np.int32)
    df = np.bincount(term_ids, minlength=vocab_size).astype(np.int64)
    post_ptr = np.zeros(vocab_size + 1, dtype=np.int64)
    np.cumsum(df, out=post_ptr[1:])
    return post_ptr, post_doc, post_tf


def default_tag_names(vocab_size: int) -> List[str]:
    names = ["t%d" % i for i in range(vocab_size)]
    # a few names that exercise the reference's ':' and paren handling (webui.py:88-99,358-371)
    specials = {5: "re:zero", 6: "fate_(series)", 7: "tag:with:colons", 8: "1girl", 11: "3:4"}
    for k, v in specials.items():
        if k < vocab_size:
            names[k] = v
    return names


def zipf_popularity(vocab_size: int, s: float = 1.0) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, vocab_size + 1, dtype=np.float64), s)
    return w / w.sum()


def generate_index(
    n_docs: int,
    vocab_size: int = DEFAULT_VOCAB,
    seed: int = 20260101,
    mean_tags: float = 28.0,
    sigma: float = 0.45,
    min_tags: int = 3,
    max_tags: int = 120,
    tf_gt1_fraction: float = 0.0,
    chunk: int = 65536,
    keep_sequences: bool = True,
    with_rows: bool = True,
    rows: str = "infer",
) -> SynthIndex:
    """Build a synthetic index (see module docstring).  rows="random": plain N(0,1) fp32 rows instead of the (slow, hash-based)
    infer_vector stand-in - for CPU timing runs at >= 1 M docs, where the row VALUES do not matter."""
    rng = np.random.default_rng(seed)
    pop = zipf_popularity(vocab_size)
    cdf = np.cumsum(pop)
    cdf[-1] = 1.0
    max_tags = min(max_tags, max(min_tags, vocab_size // 2))
    infer = InferVectorStub(vocab_size, seed)

    row_counts = np.zeros(n_docs, dtype=np.int64)
    tid_chunks: List[np.ndarray] = []
    tf_chunks: List[np.ndarray] = []
    rows_mode = rows
    rows_arr = np.zeros((n_docs, DIM), dtype=np.float32) if with_rows else np.zeros((0, DIM), dtype=np.float32)
    seqs: Optional[List[np.ndarray]] = [] if keep_sequences else None
    doc_len = np.zeros(n_docs, dtype=np.int64)

    for lo in range(0, n_docs, chunk):
        hi = min(n_docs, lo + chunk)
        m = hi - lo
        want = np.clip(np.rint(rng.lognormal(np.log(mean_tags), sigma, size=m)), min_tags, max_tags).astype(np.int64)
        # oversample with replacement, keep the first `want` DISTINCT tags per doc
        width = int(min(vocab_size, max_tags * 3 + 8))
        draws = np.searchsorted(cdf, rng.random((m, width)), side="right").astype(np.int32)
        np.clip(draws, 0, vocab_size - 1, out=draws)
        # first-occurrence mask per row
        order = np.argsort(draws, axis=1, kind="stable")
        sorted_draws = np.take_along_axis(draws, order, axis=1)
        dup_sorted = np.zeros_like(sorted_draws, dtype=bool)
        dup_sorted[:, 1:] = sorted_draws[:, 1:] == sorted_draws[:, :-1]
        dup = np.zeros_like(dup_sorted)
        np.put_along_axis(dup, order, dup_sorted, axis=1)
        first = ~dup
        rank = np.cumsum(first, axis=1)            # 1-based rank among distinct tags
        keep = first & (rank <= want[:, None])
        got = keep.sum(axis=1)
        want = np.minimum(want, got)               # (rarely) fewer distinct tags than asked
        # padded [m, max_tags] distinct-tag sequences in draw order
        width_out = int(want.max())
        padded = np.zeros((m, width_out), dtype=np.int32)
        r_idx, c_idx = np.nonzero(keep)
        padded[r_idx, rank[r_idx, c_idx] - 1] = draws[r_idx, c_idx]
        lengths = want.copy()
        tf_pad = np.ones((m, width_out), dtype=np.int32)
        seq_pad, seq_len = padded, lengths
        if tf_gt1_fraction > 0:
            # a slice of docs repeats its first tag (tf = 2) and a few also their second (tf = 3 total 2+... )
            rep = rng.random(m) < tf_gt1_fraction
            tf_pad[rep, 0] = 2
            rep3 = rep & (rng.random(m) < 0.3)
            tf_pad[rep3, 1] = 3
            # the csv sequence repeats the tag at the end of the line
            extra = (tf_pad - 1) * (np.arange(width_out)[None, :] < lengths[:, None])
            n_extra = extra.sum(axis=1)
            seq_len = lengths + n_extra
            seq_pad = np.zeros((m, int(seq_len.max())), dtype=np.int32)
            seq_pad[:, :width_out] = padded
            for i in np.nonzero(rep)[0]:
                reps = np.repeat(padded[i, : lengths[i]], extra[i, : lengths[i]])
                seq_pad[i, lengths[i] : lengths[i] + len(reps)] = reps
        live = np.arange(width_out)[None, :] < lengths[:, None]
        tid_chunks.append(padded[live])
        tf_chunks.append(tf_pad[live])
        row_counts[lo:hi] = lengths
        doc_len[lo:hi] = (tf_pad * live).sum(axis=1)
        if with_rows:
            rows_arr[lo:hi] = infer.batch(seq_pad, seq_len) if rows_mode == "infer" else rng.standard_normal((m, DIM), dtype=np.float32)
        if seqs is not None:
            for i in range(m):
                seqs.append(seq_pad[i, : seq_len[i]].copy())

    row_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(row_counts, out=row_ptr[1:])
    term_ids = np.concatenate(tid_chunks) if tid_chunks else np.zeros(0, np.int32)
    tfs = np.concatenate(tf_chunks) if tf_chunks else np.zeros(0, np.int32)
    df = np.bincount(term_ids, minlength=vocab_size).astype(np.int64)
    avgdl = np.mean(doc_len)
    D = n_docs
    with np.errstate(divide="ignore"):
        idf = np.log(1 + (D - df + 0.5) / (df + 0.5))
    idf = np.where(df > 0, idf, 0.0).astype(np.float64)
    return SynthIndex(
        n_docs=n_docs, vocab_size=vocab_size, seed=seed, row_ptr=row_ptr, term_ids=term_ids.astype(np.int32),
        tfs=tfs.astype(np.int32), doc_len=doc_len, avgdl=np.float64(avgdl), idf=idf, df=df, rows=rows_arr,
        tag_names=default_tag_names(vocab_size), infer=infer, doc_tag_seq=seqs, popularity=pop,
    )


def generate_queries(idx: SynthIndex, n_queries: int, seed: int = 7, max_terms: int = 6,
                     p_required: float = 0.3, p_exclude: float = 0.3) -> List[str]:
    """Weighted tag queries in the reference's grammar."""
    rng = np.random.default_rng(seed)
    present = np.nonzero(idx.df)[0]
    w = np.sqrt(idx.popularity[present])
    w = w / w.sum()
    top_pop = present[np.argsort(-idx.df[present], kind="stable")[: max(1, min(200, len(present) // 4 + 1))]]
    out: List[str] = []
    for _ in range(n_queries):
        t = int(rng.integers(1, max_terms + 1))
        tags = rng.choice(present, size=min(t, len(present)), replace=False, p=w)
        toks = []
        for tg in tags:
            wt = int(rng.integers(1, 6))
            name = idx.tag_names[int(tg)]
            toks.append(name if wt == 1 and rng.random() < 0.5 else "%s:%d" % (name, wt))
        used = set(int(x) for x in tags)
        if rng.random() < p_required:
            cand = [int(c) for c in top_pop[:50] if int(c) not in used]
            if cand:
                tg = cand[int(rng.integers(0, len(cand)))]
                used.add(tg)
                toks.append("%s:+%d" % (idx.tag_names[tg], int(rng.integers(1, 4))))
        if rng.random() < p_exclude:
            cand = [int(c) for c in top_pop if int(c) not in used]
            if cand:
                tg = cand[int(rng.integers(0, len(cand)))]
                toks.append("%s:-%d" % (idx.tag_names[tg], int(rng.integers(1, 4))))
        order = rng.permutation(len(toks))
        out.append(" ".join(toks[i] for i in order))
    return out
