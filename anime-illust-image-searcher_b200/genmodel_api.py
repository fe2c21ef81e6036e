"""Drop-in for genmodel.py's BM25 index builder (genmodel.py:51-99), computed on the GPU.

    from ais_b200.genmodel_api import gen_and_save_bm25_index      # same signature, same five files

The term-frequency counting, doc lengths, document frequencies and the posting lists the scoring engine
reads are built by CUDA kernels (csrc/build.cuh) through ``ais_build_bm25``; the IDF table uses
``numpy.log`` and avgdl ``numpy.mean`` exactly like the reference so the pickled values are bit-identical.
``genmodel.py --update`` rebuilds this index from scratch over all docs (genmodel.py:134,177), so an update
is simply another call.
"""
from __future__ import annotations

import pickle
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine as _eng


def tokens_to_csr(corpus: Sequence[Sequence[str]], token2id: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray]:
    """genmodel.py:59-61: tags -> term ids, unknown tags dropped.  -> (seq_ptr int64[N+1], seq_ids int32[total])"""
    ptr = np.zeros(len(corpus) + 1, dtype=np.int64)
    ids: List[int] = []
    for i, tags in enumerate(corpus):
        ids.extend(token2id[t] for t in tags if t in token2id)
        ptr[i + 1] = len(ids)
    return ptr, np.asarray(ids, dtype=np.int32)


def idf_table(df: np.ndarray, n_docs: int) -> Dict[int, np.float64]:
    """genmodel.py:79-82, term by term with numpy scalars like the reference."""
    out: Dict[int, np.float64] = {}
    for t in np.nonzero(df)[0]:
        d = int(df[t])
        out[int(t)] = np.log(1 + (n_docs - d + 0.5) / (d + 0.5))
    return out


def build_index(engine: _eng.SearchEngine, seq_ptr: np.ndarray, seq_ids: np.ndarray, n_terms: int):
    """GPU build + staging: after this call the engine can score BM25 for these docs.
    -> (doc_lengths int64[N], avgdl np.float64, idf {term: np.float64}, df int64[V])"""
    df, doc_len = engine.build_bm25(seq_ptr, seq_ids, n_terms)
    n_docs = len(doc_len)
    avgdl = np.mean(doc_len) if n_docs else np.float64(0.0)          # genmodel.py:76
    idf = idf_table(df, n_docs)
    dense = np.zeros(n_terms, dtype=np.float64)
    for t, v in idf.items():
        dense[t] = v
    engine.finish_bm25(dense, float(avgdl))
    return doc_len, avgdl, idf, df


def gen_and_save_bm25_index(corpus: List[List[str]], dictionary, engine: Optional[_eng.SearchEngine] = None,
                            device: int = 0) -> _eng.SearchEngine:
    """genmodel.py:51-99: writes bm25_corpus / bm25_idf / bm25_avgdl / bm25_D / bm25_doc_lengths into the CWD and
    returns the engine that now holds the index on the device."""
    token2id = dictionary.token2id
    eng = engine or _eng.SearchEngine(device=device)
    seq_ptr, seq_ids = tokens_to_csr(corpus, token2id)
    n_terms = (max(token2id.values()) + 1) if token2id else 1
    doc_len, avgdl, idf, _df = build_index(eng, seq_ptr, seq_ids, n_terms)
    # the pickle holds python dicts (genmodel.py:64-68); their content comes back from the device postings
    ptr, doc, tf = eng.export_postings()
    bm25_corpus: List[Dict[int, int]] = [dict() for _ in range(len(corpus))]
    order = {}
    for i in range(len(corpus)):                      # first-occurrence key order, like the reference's dict
        seen = {}
        for t in seq_ids[seq_ptr[i]: seq_ptr[i + 1]].tolist():
            if t not in seen:
                seen[t] = 0
        bm25_corpus[i] = seen
    terms = np.repeat(np.arange(n_terms), np.diff(ptr))
    for t, d, f in zip(terms.tolist(), doc.tolist(), tf.tolist()):
        bm25_corpus[d][t] = f
    for name, obj in (("bm25_corpus", bm25_corpus), ("bm25_idf", idf), ("bm25_avgdl", avgdl), ("bm25_D", len(corpus)),
                      ("bm25_doc_lengths", doc_len)):
        with open(name, "wb") as fh:
            pickle.dump(obj, fh)
    print("BM25 index generated")
    return eng
