"""Drop-in for the callables webui.py's search path uses - same names, arguments,
return values and exceptions, computed on the B200 through the C ABI.

    from ais_b200.webui_api import load_model, find_similar_documents     # instead of webui.py's own

Module globals mirror webui.py:24-60.  The constants are read at call time, as in the reference
("modifiable", webui.py:51-52).  ``PRF_MODE`` chooses how the pseudo-relevance-feedback re-query
vector is obtained: "callback" (default, reference behaviour: the top docs are RE-INFERRED with
``model.infer_vector`` on the host, webui.py:182-187) or "stored_rows" (device-only: the stored
rows of the top docs are used instead - identical whenever inference is deterministic).
"""
from __future__ import annotations

import os
import pickle
import threading
from itertools import chain
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine as _eng
from . import query as _q

# ---- webui.py:24-60 ------------------------------------------------------------------------------
model: Any = None
index: Any = None
dictionary: Any = None
image_files_name_tags_arr: List[str] = []
file_tag_index_dict: Dict[str, Dict[str, bool]] = {}
filepath_docid_dict: Dict[str, int] = {}
bm25_D: int = 0

BM25_WEIGHT = 0.5
DOC2VEC_WEIGHT = 0.5
ORIGINAL_SCORE_WEIGHT = 0.7
RERANKED_SCORE_WEIGHT = 0.3
DIFF_FILTER_THRESH = 1e-6
REQUIRE_TAG_MAGIC_NUMBER = 1000

PRF_MODE = "callback"
DEVICE = int(os.environ.get("AIS_DEVICE", "0"))

_engine: Optional[_eng.SearchEngine] = None
_loaded_from: Optional[str] = None
_lock = threading.Lock()

INDEX_CSV = "tags-wd-tagger_doc2vec_idx.csv"


class GpuSimilarity:
    """The ``index`` object: ``index[vec] -> float32[N]`` like gensim's Similarity (webui.py:352,205)."""

    def __init__(self, engine: _eng.SearchEngine):
        self._engine = engine
        self.num_features = _eng.DIM

    def __len__(self) -> int:
        return self._engine.n_docs

    def __getitem__(self, vec: Sequence[Tuple[int, float]]) -> np.ndarray:
        return self._engine.dot_scores(_q.dense_query(vec, self.num_features))


def _sync_constants() -> _eng.SearchEngine:
    if _engine is None:
        raise RuntimeError("index not loaded: call load_model() (or install()) first")
    p = _engine.params
    want = dict(bm25_weight=BM25_WEIGHT, doc2vec_weight=DOC2VEC_WEIGHT, original_score_weight=ORIGINAL_SCORE_WEIGHT,
                reranked_score_weight=RERANKED_SCORE_WEIGHT, diff_filter_thresh=DIFF_FILTER_THRESH,
                require_magic=float(REQUIRE_TAG_MAGIC_NUMBER))
    if any(getattr(p, k) != v for k, v in want.items()):
        _engine.set_params(**want)
    return _engine


def _prf_mode() -> int:
    return {"callback": _eng.PRF_CALLBACK, "stored_rows": _eng.PRF_STORED_ROWS,
            "stored_rows_full": _eng.PRF_STORED_ROWS_FULL, "off": _eng.PRF_OFF}[PRF_MODE]


# ---- index staging -----------------------------------------------------------------------------------
def corpus_to_postings(corpus: Sequence[Dict[int, int]], n_terms: int):
    """bm25_corpus (genmodel.py:64-68: one {term_id: tf} per doc) -> tag-major posting lists."""
    lens = np.fromiter((len(d) for d in corpus), dtype=np.int64, count=len(corpus))
    nnz = int(lens.sum())
    tids = np.fromiter(chain.from_iterable(d.keys() for d in corpus), dtype=np.int64, count=nnz)
    tfs = np.fromiter(chain.from_iterable(d.values() for d in corpus), dtype=np.int64, count=nnz)
    if nnz and (tids.min() < 0 or tids.max() >= n_terms):
        raise ValueError("term id outside [0, %d)" % n_terms)
    docs = np.repeat(np.arange(len(corpus), dtype=np.int32), lens)
    order = np.argsort(tids, kind="stable")
    df = np.bincount(tids, minlength=n_terms).astype(np.int64)
    post_ptr = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(df, out=post_ptr[1:])
    post_tf = tfs[order].astype(np.int32)
    return post_ptr, docs[order], (post_tf if nnz and post_tf.max() > 1 else None)


def stage_bm25(engine: _eng.SearchEngine, corpus, doc_lengths, avgdl, idf: Dict[int, float], n_terms: Optional[int] = None):
    """The five pickles of genmodel.py:84-97 -> engine."""
    top = max(chain((max(d) for d in corpus if d), idf.keys()), default=-1) + 1
    n_terms = max(n_terms or 0, top)
    post_ptr, post_doc, post_tf = corpus_to_postings(corpus, n_terms)
    idf_arr = np.zeros(n_terms, dtype=np.float64)
    for t, v in idf.items():
        idf_arr[t] = v
    engine.load_bm25(post_ptr, post_doc, post_tf, idf_arr, np.asarray(doc_lengths, dtype=np.int64), float(avgdl))


def install(engine: _eng.SearchEngine, model_: Any, dictionary_: Any, csv_lines: List[str]) -> None:
    """Adopt an already staged engine plus the host-side gensim objects (tests, benchmarks, custom loaders)."""
    global _engine, model, index, dictionary, image_files_name_tags_arr, filepath_docid_dict, bm25_D, _loaded_from
    _engine = engine
    model = model_
    dictionary = dictionary_
    index = GpuSimilarity(engine)
    image_files_name_tags_arr = csv_lines
    filepath_docid_dict = {line.split(",")[0]: i for i, line in enumerate(csv_lines)}
    bm25_D = engine.n_docs
    _loaded_from = "<installed>"


MAX_BATCH = int(os.environ.get("AIS_MAX_BATCH", "64"))     # queries that share each pass in find_similar_documents_batch
MODEL_LOADER = None      # callable(path) -> object with infer_vector / dv; default: gensim's Doc2Vec.load (webui.py:668)


class _Dictionary:
    """``dictionary`` as far as the path uses it: ``token2id`` (webui.py:133,364-371)."""

    def __init__(self, token2id: Dict[str, int]):
        self.token2id = token2id


def _load_doc2vec(path: str):
    if MODEL_LOADER is not None:
        return MODEL_LOADER(path)
    try:
        from gensim.models.doc2vec import Doc2Vec
    except ImportError as exc:      # query inference stays gensim's (north_star); everything else loads without it
        raise ImportError("load_model() needs gensim for doc2vec_model (query inference, webui.py:106,185) - the index, "
                          "dictionary and BM25 files are read natively; set webui_api.MODEL_LOADER to supply another "
                          "infer_vector implementation") from exc
    return Doc2Vec.load(path)


def load_model() -> None:
    """webui.py:649-689: reads the index files from the CWD - ONCE; repeat calls (webui calls it on every search,
    webui.py:585) are no-ops while the CWD is unchanged.  The doc-vector shards, the dictionary and the five BM25 files
    are read by ais_b200.loader (memory-mapped .npy shards, C walker over the bm25_corpus pickle: no gensim import, no N
    Python dicts); only ``doc2vec_model`` goes through gensim, because query inference stays gensim's."""
    global _engine, model, index, dictionary, image_files_name_tags_arr, file_tag_index_dict, filepath_docid_dict
    global bm25_D, _loaded_from
    from . import loader
    with _lock:
        cwd = os.getcwd()
        if _engine is not None and _loaded_from == cwd:
            return
        with open(INDEX_CSV, "r", encoding="utf-8") as f:
            lines = [line.strip() for line in f]
        tag_index: Dict[str, Dict[str, bool]] = {}
        for line in lines:
            parts = line.split(",")
            tag_index[parts[0]] = {t: True for t in parts[1:]}
        mdl = _load_doc2vec("doc2vec_model")
        dct = _Dictionary(loader.read_token2id("doc2vec_dictionary"))
        eng = _eng.SearchEngine(device=DEVICE, max_batch=MAX_BATCH)
        bm = loader.stage_index(eng, ".")
        if len(lines) != eng.n_docs:
            raise ValueError("%s has %d lines but the index holds %d docs" % (INDEX_CSV, len(lines), eng.n_docs))
        install(eng, mdl, dct, lines)
        file_tag_index_dict = tag_index
        bm25_D = int(bm["D"])
        _loaded_from = cwd


# ---- webui.py:82-117 ------------------------------------------------------------------------------------
def normalize_and_apply_weight_doc2vec(new_doc: str) -> List[Tuple[int, float]]:
    return _q.query_vector(new_doc, model.infer_vector, len(model.dv[0]))


# ---- webui.py:119-172 -----------------------------------------------------------------------------------
def compute_bm25_scores(query_terms: List[str] = [], query_weights: Optional[Dict[int, float]] = None) -> np.ndarray:
    eng = _sync_constants()
    if query_weights is not None:
        ids = list(query_weights.keys())
        weights = [query_weights[t] for t in ids]
    else:
        ids = [dictionary.token2id[t] for t in query_terms if t in dictionary.token2id]
        weights = [1.0] * len(ids)
    return eng.bm25_scores(ids, weights)


# ---- webui.py:182-187 -----------------------------------------------------------------------------------
def get_embedded_vector_by_doc_id(doc_id: int) -> List[Tuple[int, float]]:
    tags = image_files_name_tags_arr[doc_id - 1].split(",")[1:]
    return [(i, val) for i, val in enumerate(model.infer_vector(tags))]


def _prf_callback(_qi: int, doc_ids: np.ndarray, scores: np.ndarray) -> np.ndarray:
    vectors = [get_embedded_vector_by_doc_id(int(d) + 1) for d in doc_ids]
    return _q.dense_query(_q.prf_query(vectors, scores.tolist()), _eng.DIM)


# ---- webui.py:63-80 ---------------------------------------------------------------------------------------
def filter_searched_result(sorted_scores: List[Tuple[int, float]]) -> List[Tuple[int, float]]:
    eng = _sync_constants()
    ids, scores = eng.filter_sorted([d for d, _ in sorted_scores], [s for _, s in sorted_scores])
    return list(zip(ids.tolist(), scores.tolist()))


# ---- webui.py:189-253 -------------------------------------------------------------------------------------
def get_doc2vec_based_reranked_scores(final_scores, topn: int) -> List[Tuple[int, float]]:
    eng = _sync_constants()
    mode = _prf_mode()
    return eng.rerank(np.asarray(final_scores, dtype=np.float64), topn, mode,
                      _prf_callback if mode == _eng.PRF_CALLBACK else None)


# ---- webui.py:345-390 -------------------------------------------------------------------------------------
def find_similar_documents(new_doc: str, topn: int = 50) -> List[Tuple[int, float]]:
    eng = _sync_constants()
    q = _q.make_query(new_doc, dictionary.token2id, model.infer_vector, len(model.dv[0]), REQUIRE_TAG_MAGIC_NUMBER)
    mode = _prf_mode()
    return eng.search([q], topn, mode, _prf_callback if mode == _eng.PRF_CALLBACK else None)[0]


def find_similar_documents_batch(new_docs: Sequence[str], topn: int = 50) -> List[List[Tuple[int, float]]]:
    """Extension: several queries share each pass over the doc vectors (engine max_batch)."""
    eng = _sync_constants()
    qs = [_q.make_query(d, dictionary.token2id, model.infer_vector, len(model.dv[0]), REQUIRE_TAG_MAGIC_NUMBER)
          for d in new_docs]
    mode = _prf_mode()
    return eng.search(qs, topn, mode, _prf_callback if mode == _eng.PRF_CALLBACK else None)
