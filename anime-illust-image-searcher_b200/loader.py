"""Native loader for the reference's index files (replaces the O(N)-Python part of load_model(),
webui.py:649-689, and needs NO gensim for the index itself).

Files, all in the current working directory like the reference (written by genmodel.py:84-97,155-156,175):

* ``bm25_corpus``      pickle of a list of N dicts {term id: tf}.  Read by the C opcode walker in the shared library
                       (csrc/pickle_csr.cpp) straight into CSR arrays - no N Python dicts; falls back to ``pickle.load``
                       if the file is outside the walker's opcode subset.
* ``bm25_idf`` / ``bm25_avgdl`` / ``bm25_D`` / ``bm25_doc_lengths``   small pickles (a dict of V floats, two scalars, one
                       ndarray): plain ``pickle.load``.
* ``doc2vec_index``    gensim ``Similarity`` pickled by ``SaveLoad.save``; its ``shards`` list names the shard files
  ``doc2vec_index.N``  each a pickled ``MatrixSimilarity`` whose fp32 matrix ``index`` [<= 32768 x 300] is either inline or,
                       above gensim's 10 MB ``sep_limit``, split out as ``doc2vec_index.N.index.npy``.  The pickles are read
                       with an unpickler that maps every ``gensim.*`` class onto a plain attribute holder (only attribute
                       values are needed, never gensim code); the .npy matrices are memory-mapped and streamed to the GPU
                       shard by shard (rows are stored RAW).
* ``doc2vec_dictionary``  pickled gensim ``Dictionary``: only ``token2id`` is used (webui.py:133,364-371).

PARITY STATUS: gensim 4.3.3 is not installable in the build container, so the gensim-side layout above follows its
documented ``SaveLoad`` behaviour and is exercised with files written in that layout by
tests/make_gensim_layout.py - "unpinned" until checked against files written by a real gensim.  The BM25 side is pinned:
the test files are written by the reference builder's own ``pickle.dump`` calls.
"""
from __future__ import annotations

import ctypes as C
import io
import os
import pickle
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from . import binding as B

DIM = B.DIM


# ---- BM25 side ----------------------------------------------------------------------------------------------------
def read_bm25_corpus_csr(path: str = "bm25_corpus") -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (row_ptr int64[N+1], term_ids int32[nnz], tfs int32[nnz]) in doc order / dict insertion order."""
    n_docs, nnz = C.c_int64(0), C.c_int64(0)
    st = B.lib.ais_pickle_csr_scan(os.fsencode(path), C.byref(n_docs), C.byref(nnz))
    if st == B.AIS_OK:
        row_ptr = np.zeros(n_docs.value + 1, dtype=np.int64)
        term_ids = np.zeros(nnz.value, dtype=np.int32)
        tfs = np.zeros(nnz.value, dtype=np.int32)
        st = B.lib.ais_pickle_csr_fill(os.fsencode(path), n_docs.value, nnz.value, row_ptr.ctypes.data, term_ids.ctypes.data,
                                       tfs.ctypes.data)
        if st == B.AIS_OK:
            return row_ptr, term_ids, tfs
    if st != B.AIS_ERR_UNSUPPORTED:
        raise B.AisError(st, B.lib.ais_pickle_last_error().decode("utf-8", "replace"))
    # outside the opcode subset (e.g. numpy scalars as keys): the reference's own way, one dict per doc
    with open(path, "rb") as f:
        corpus = pickle.load(f)
    return corpus_to_csr(corpus)


def corpus_to_csr(corpus) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    counts = np.fromiter((len(d) for d in corpus), dtype=np.int64, count=len(corpus))
    row_ptr = np.zeros(len(corpus) + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    term_ids = np.fromiter((k for d in corpus for k in d.keys()), dtype=np.int32, count=nnz)
    tfs = np.fromiter((v for d in corpus for v in d.values()), dtype=np.int32, count=nnz)
    return row_ptr, term_ids, tfs


def read_bm25_index(dirpath: str = ".") -> dict:
    """The five BM25 files of genmodel.py:84-97 -> arrays the engine stages (CSR + dense idf table)."""
    p = lambda name: os.path.join(dirpath, name)
    row_ptr, term_ids, tfs = read_bm25_corpus_csr(p("bm25_corpus"))
    with open(p("bm25_idf"), "rb") as f:
        idf_dict = pickle.load(f)
    with open(p("bm25_avgdl"), "rb") as f:
        avgdl = pickle.load(f)
    with open(p("bm25_D"), "rb") as f:
        D = pickle.load(f)
    with open(p("bm25_doc_lengths"), "rb") as f:
        doc_len = np.asarray(pickle.load(f), dtype=np.int64)
    n_terms = int(max(max(idf_dict.keys(), default=-1), int(term_ids.max()) if len(term_ids) else -1)) + 1
    idf = np.zeros(n_terms, dtype=np.float64)                      # 0 where absent: bm25_idf.get(term_id, 0), webui.py:140
    for t, v in idf_dict.items():
        idf[int(t)] = float(v)
    if int(D) != len(row_ptr) - 1 or len(doc_len) != len(row_ptr) - 1:
        raise ValueError("inconsistent BM25 files: bm25_D=%d, %d corpus docs, %d doc lengths" % (int(D), len(row_ptr) - 1, len(doc_len)))
    return {"row_ptr": row_ptr, "term_ids": term_ids, "tfs": tfs, "idf": idf, "idf_dict": idf_dict, "avgdl": float(avgdl), "D": int(D),
            "doc_len": doc_len, "n_terms": n_terms}


# ---- gensim side (no gensim import) ------------------------------------------------------------------------------------
class _Holder:
    """Stands in for any ``gensim.*`` class while unpickling: keeps the attribute dict, runs no gensim code."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2:
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)


class _GensimFreeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "gensim" or module.startswith("gensim."):
            return type(name, (_Holder,), {"__module__": module})
        return super().find_class(module, name)


def load_holder(path: str):
    with open(path, "rb") as f:
        return _GensimFreeUnpickler(f).load()


def iter_similarity_shards(prefix: str = "doc2vec_index") -> Iterator[np.ndarray]:
    """fp32 [rows, 300] matrices of a saved gensim ``Similarity`` in doc order, memory-mapped where gensim split them out."""
    sim = load_holder(prefix)
    shards = getattr(sim, "shards", None)
    if shards is None:                                             # a bare MatrixSimilarity saved under the prefix
        yield _shard_matrix(prefix, sim)
        return
    here = os.path.dirname(os.path.abspath(prefix))
    for sh in shards:
        fname = os.path.join(here, os.path.basename(getattr(sh, "fname")))      # Shard.fullname(): dirname is re-based on load
        yield _shard_matrix(fname, load_holder(fname), expect_rows=getattr(sh, "length", None))


def _shard_matrix(fname: str, obj, expect_rows: Optional[int] = None) -> np.ndarray:
    m = getattr(obj, "index", None)
    if m is None:                                                  # SaveLoad stored it separately: <fname>.index.npy
        m = np.load(fname + ".index.npy", mmap_mode="r")
    m = np.asarray(m) if not isinstance(m, np.memmap) else m
    if m.ndim != 2 or m.shape[1] != DIM:
        raise ValueError("%s: shard matrix has shape %r, expected [rows, %d]" % (fname, m.shape, DIM))
    if m.dtype != np.float32:
        m = m.astype(np.float32)
    if expect_rows is not None and m.shape[0] != expect_rows:
        raise ValueError("%s: %d rows, the index says %d" % (fname, m.shape[0], expect_rows))
    return m


def read_token2id(path: str = "doc2vec_dictionary") -> Dict[str, int]:
    d = load_holder(path)
    t2i = d if isinstance(d, dict) else getattr(d, "token2id")
    return dict(t2i)


# ---- everything into an engine ---------------------------------------------------------------------------------------------
def stage_index(engine, dirpath: str = ".", index_prefix: str = "doc2vec_index") -> dict:
    """Stage the reference's index files into ``engine`` (a SearchEngine): doc vectors shard by shard, then the BM25
    index as tag-major posting lists.  Returns the BM25 dict (host arrays) for callers that need token-level access."""
    from .synth import csr_to_postings
    n = 0
    shards = list(iter_similarity_shards(os.path.join(dirpath, index_prefix)))
    total = sum(s.shape[0] for s in shards)
    engine.reserve_docs(total)
    for m in shards:
        engine.load_vectors(np.ascontiguousarray(m), first_row=n)          # one cudaMemcpy per (memory-mapped) shard
        n += m.shape[0]
    bm = read_bm25_index(dirpath)
    if bm["D"] != n:
        raise ValueError("doc2vec_index holds %d docs but the BM25 index %d" % (n, bm["D"]))
    post_ptr, post_doc, post_tf = csr_to_postings(bm["row_ptr"], bm["term_ids"], bm["tfs"], bm["n_terms"])
    tf = post_tf if (len(post_tf) and int(post_tf.max()) > 1) else None
    engine.load_bm25(post_ptr, post_doc, tf, bm["idf"], bm["doc_len"], bm["avgdl"])
    engine.set_shard(0, n)
    return bm
