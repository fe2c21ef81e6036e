"""SearchEngine: one GPU, one contiguous shard of the documents, behind the C ABI (binding.py).

Host-side responsibilities only: keep numpy/torch buffers alive across the ctypes calls, turn
status codes into the exceptions the reference raises, and expose the staged
calls with torch tensors so shard.py can put its collectives between them.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import binding as B
from .binding import lib, check

PRF_CALLBACK = B.AIS_PRF_CALLBACK
PRF_STORED_ROWS = B.AIS_PRF_STORED_ROWS
PRF_STORED_ROWS_FULL = B.AIS_PRF_STORED_ROWS_FULL
PRF_OFF = B.AIS_PRF_OFF
DIM = B.DIM


@dataclass
class Query:
    """A parsed weighted tag query (the parsing itself stays Python: webui.py:354-371).  Immutable after construction:
    the C-ABI record of the query (pointers into the three arrays) is formed once."""
    vec: np.ndarray        # fp32[300]: sparse2full(unitvec(normalize_and_apply_weight_doc2vec(q)))
    term_ids: np.ndarray   # int32[T]  dict keys in insertion order
    weights: np.ndarray    # float64[T] 1000+W = required, negative = exclude

    def __post_init__(self):
        self.vec = np.ascontiguousarray(self.vec, dtype=np.float32)
        self.term_ids = np.ascontiguousarray(self.term_ids, dtype=np.int32)
        self.weights = np.ascontiguousarray(self.weights, dtype=np.float64)
        if self.vec.shape != (DIM,):
            raise ValueError("query vector must have %d components" % DIM)
        if self.term_ids.shape != self.weights.shape or self.term_ids.ndim != 1:
            raise ValueError("term_ids / weights must be 1-D and of equal length")
        if len(self.term_ids) > B.AIS_MAX_TERMS:
            raise ValueError("at most %d query terms" % B.AIS_MAX_TERMS)
        # the ais_query record of this query (three pointers + n_terms, 32 bytes), formed once: the arrays above are
        # owned by the object and never reallocated, so the addresses stay valid for its lifetime
        self._rec = (self.vec.__array_interface__["data"][0], self.term_ids.__array_interface__["data"][0],
                     self.weights.__array_interface__["data"][0], len(self.term_ids))


def raise_for_status(status: int) -> None:
    """The exception the reference raises in the corresponding situation."""
    if status == B.AIS_Q_OK:
        return
    if status == B.AIS_Q_NAN_WEIGHTS:
        raise ValueError("cannot convert float NaN to integer")          # round(nan), webui.py:202-203
    if status == B.AIS_Q_ZERO_WEIGHT_SUM:
        raise ZeroDivisionError("Weights sum to zero, can't be normalized")  # np.average, webui.py:200
    if status == B.AIS_Q_ZERO_VECTOR:
        raise AssertionError("sparse documents must not contain any explicit zero entries")  # gensim unitvec
    raise RuntimeError("PRF callback failed (status %d)" % status)


def _ptr(a) -> C.c_void_p:
    """Address of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


class _Packed:
    """An array of ais_query records in one numpy buffer (kept alive with the object): building ctypes structures
    field by field costs ~15 us per query, which at 256 queries per batch is milliseconds of idle GPU per step."""
    __slots__ = ("buf", "_as_parameter_")

    def __init__(self, queries: Sequence[Query]):
        assert C.sizeof(B.AisQuery) == 32
        self.buf = np.array([q._rec for q in queries], dtype=np.uint64).reshape(-1, 4)   # little-endian: n_terms in the low half
        self._as_parameter_ = C.cast(self.buf.ctypes.data, C.POINTER(B.AisQuery))


def _pack_queries(queries: Sequence[Query]):
    return _Packed(queries)


_NULL_CB = C.cast(None, B.INFER_CB)


class SearchEngine:
    def __init__(self, device: int = 0, max_batch: int = 1, **params):
        self._h = C.c_void_p(0)
        self._lock = threading.Lock()
        p = B.AisParams()
        lib.ais_default_params(C.byref(p))
        p.max_batch = max_batch
        for k, v in params.items():
            if not hasattr(p, k):
                raise TypeError("unknown engine parameter %r" % k)
            setattr(p, k, v)
        self.params = p
        check(lib.ais_create(C.byref(self._h), device, C.byref(p)))
        self.device = device
        self.n_docs = 0
        self.first_doc = 0
        self.n_total = 0

    # ---- lifecycle --------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            lib.ais_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **params) -> None:
        for k, v in params.items():
            if not hasattr(self.params, k):
                raise TypeError("unknown engine parameter %r" % k)
            setattr(self.params, k, v)
        check(lib.ais_set_params(self._h, C.byref(self.params)))

    @property
    def torch_device(self):
        import torch
        return torch.device("cuda", self.device)

    def use_torch_stream(self) -> None:
        """Run on torch's current CUDA stream so torch events / NCCL collectives order with the kernels."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        # handle 0 is torch's (legacy) default stream; the C ABI reserves NULL for "the engine's own stream",
        # so name the default stream by its explicit handle cudaStreamLegacy (0x1)
        check(lib.ais_set_stream(self._h, C.c_void_p(s if s else 1)))

    # ---- index staging (load_model, webui.py:649-689) ------------------------------------------
    def load_vectors(self, rows, first_row: int = 0) -> None:
        """rows: fp32 [n, 300] numpy array or torch tensor (host or device), stored RAW."""
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            n, dim = rows.shape
        else:
            rows = rows.contiguous()
            assert str(rows.dtype) == "torch.float32"
            n, dim = rows.shape
        check(lib.ais_load_vectors(self._h, _ptr(rows), n, dim, first_row))
        self.n_docs = max(self.n_docs, first_row + n)

    def reserve_docs(self, n: int) -> None:
        check(lib.ais_reserve_docs(self._h, n))

    def rows_tensor(self, n_docs: int):
        """torch view [n_docs, 300] of the engine-owned row store (write the rows in place)."""
        import torch
        p = C.c_void_p(0)
        check(lib.ais_vectors_device_ptr(self._h, n_docs, C.byref(p)))
        self.n_docs = n_docs

        class _Mem:     # __cuda_array_interface__ wrapper
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (n_docs, DIM), "typestr": "<f4", "data": (p.value, False), "version": 2}
        with torch.cuda.device(self.device):
            return torch.as_tensor(m, device="cuda:%d" % self.device)

    def load_bm25(self, post_ptr, post_doc, post_tf, idf, doc_len, avgdl: float) -> None:
        """Tag-major posting lists with LOCAL ascending doc ids (numpy arrays or torch tensors)."""
        def prep(a, dt):
            if a is None:
                return None
            if isinstance(a, np.ndarray):
                return np.ascontiguousarray(a, dtype=dt)
            return a.contiguous()
        post_ptr = prep(post_ptr, np.int64)
        post_doc = prep(post_doc, np.int32)
        post_tf = prep(post_tf, np.int32)
        idf = prep(idf, np.float64)
        doc_len = prep(doc_len, np.int64)
        n_terms = len(post_ptr) - 1
        n_docs = len(doc_len)
        check(lib.ais_load_bm25(self._h, _ptr(post_ptr), _ptr(post_doc), _ptr(post_tf), n_terms, n_docs, _ptr(idf),
                                _ptr(doc_len), float(avgdl)))
        self.n_bm25 = n_docs

    def build_bm25(self, seq_ptr, seq_ids, n_terms: int):
        """gen_and_save_bm25_index (genmodel.py:51-99) on the GPU from the docs' term-id sequences (CSR).
        -> (df int64[n_terms], doc_len int64[n_docs]); call finish_bm25(idf, avgdl) afterwards."""
        seq_ptr = np.ascontiguousarray(seq_ptr, dtype=np.int64) if isinstance(seq_ptr, np.ndarray) else seq_ptr.contiguous()
        seq_ids = np.ascontiguousarray(seq_ids, dtype=np.int32) if isinstance(seq_ids, np.ndarray) else seq_ids.contiguous()
        n_docs = len(seq_ptr) - 1
        df = np.zeros(n_terms, dtype=np.int64)
        doc_len = np.zeros(n_docs, dtype=np.int64)
        with self._lock:
            check(lib.ais_build_bm25(self._h, _ptr(seq_ptr), _ptr(seq_ids), n_docs, n_terms, _ptr(df), _ptr(doc_len)))
        self._n_built, self._v_built = n_docs, n_terms
        return df, doc_len

    def finish_bm25(self, idf: np.ndarray, avgdl: float) -> None:
        idf = np.ascontiguousarray(idf, dtype=np.float64)
        with self._lock:
            check(lib.ais_finish_bm25(self._h, _ptr(idf), float(avgdl)))
        self.n_bm25 = self._n_built

    def export_postings(self, with_tf: bool = True):
        ptr = np.zeros(self._v_built + 1, dtype=np.int64)
        check(lib.ais_export_postings(self._h, _ptr(ptr), None, None))
        n = int(ptr[-1])
        doc = np.zeros(n, dtype=np.int32)
        tf = np.zeros(n, dtype=np.int32) if with_tf else None
        check(lib.ais_export_postings(self._h, _ptr(ptr), _ptr(doc), _ptr(tf)))
        return ptr, doc, tf

    def set_shard(self, first_doc: int, n_total: int) -> None:
        check(lib.ais_set_shard(self._h, first_doc, n_total))
        self.first_doc, self.n_total = first_doc, n_total

    @classmethod
    def from_index(cls, idx, device: int = 0, max_batch: int = 1, lo: int = 0, hi: Optional[int] = None, **params):
        """Stage docs [lo, hi) of a SynthIndex-like object (rows/row_ptr/term_ids/tfs/idf/doc_len/avgdl).
        IDF / avgdl stay the GLOBAL values of the index files."""
        from .synth import csr_to_postings
        hi = idx.n_docs if hi is None else hi
        eng = cls(device=device, max_batch=max_batch, **params)
        a, b = int(idx.row_ptr[lo]), int(idx.row_ptr[hi])
        row_ptr = idx.row_ptr[lo:hi + 1] - idx.row_ptr[lo]
        post_ptr, post_doc, post_tf = csr_to_postings(row_ptr, idx.term_ids[a:b], idx.tfs[a:b], idx.vocab_size)
        tf = post_tf if (len(post_tf) and post_tf.max() > 1) else None
        eng.reserve_docs(hi - lo)
        eng.load_vectors(idx.rows[lo:hi])
        eng.load_bm25(post_ptr, post_doc, tf, idx.idf, idx.doc_len[lo:hi], float(idx.avgdl))
        eng.set_shard(lo, idx.n_docs)
        return eng

    # ---- seams --------------------------------------------------------------------------------
    def dot_scores(self, q: np.ndarray) -> np.ndarray:
        """index[vec] (webui.py:352,205) for an already dense unit query."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.empty(self.n_docs, dtype=np.float32)
        with self._lock:
            check(lib.ais_dot_scores(self._h, _ptr(q), _ptr(out)))
        return out

    def bm25_scores(self, term_ids, weights) -> np.ndarray:
        """compute_bm25_scores(query_weights=...) webui.py:119-172."""
        t = np.ascontiguousarray(term_ids, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if len(t) > B.AIS_MAX_TERMS:
            raise ValueError("at most %d query terms" % B.AIS_MAX_TERMS)
        out = np.empty(self.n_bm25, dtype=np.float64)
        with self._lock:
            check(lib.ais_bm25_scores(self._h, _ptr(t), _ptr(w), len(t), _ptr(out)))
        return out

    def final_scores(self, q: Query) -> np.ndarray:
        out = np.empty(self.n_docs, dtype=np.float64)
        arr = _pack_queries([q])
        with self._lock:
            check(lib.ais_final_scores(self._h, arr, _ptr(out)))
        return out

    def filter_sorted(self, ids, scores) -> Tuple[np.ndarray, np.ndarray]:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        n = len(ids)
        oi = np.empty(n, dtype=np.int64)
        os_ = np.empty(n, dtype=np.float64)
        cnt = C.c_int64(0)
        with self._lock:
            check(lib.ais_filter_sorted(self._h, _ptr(ids), _ptr(scores), n, _ptr(oi), _ptr(os_), C.byref(cnt)))
        return oi[:cnt.value], os_[:cnt.value]

    # ---- the fused path ------------------------------------------------------------------------
    def _make_cb(self, infer_cb, errors: list):
        if infer_cb is None:
            return _NULL_CB

        def trampoline(_ctx, qi, ids_p, scores_p, depth, out_p):
            try:
                ids = np.ctypeslib.as_array(ids_p, shape=(depth,)).copy()
                scores = np.ctypeslib.as_array(scores_p, shape=(depth,)).copy()
                vec = np.ascontiguousarray(infer_cb(int(qi), ids, scores), dtype=np.float32)
                if vec.shape != (DIM,):
                    raise ValueError("PRF callback must return %d floats" % DIM)
                C.memmove(out_p, vec.ctypes.data, DIM * 4)
                return 0
            except BaseException as exc:      # noqa: BLE001 - re-raised by the caller after the C call returns
                errors.append((int(qi), exc))
                return 1
        return B.INFER_CB(trampoline)

    def search_raw(self, queries: Sequence[Query], topn: int, prf_mode: int = PRF_STORED_ROWS,
                   infer_cb: Optional[Callable] = None):
        """-> (ids int64[nq, topn], scores float64[nq, topn], counts int32[nq], status int32[nq], cb_errors)."""
        nq = len(queries)
        arr = _pack_queries(queries)
        ids = np.zeros((nq, topn), dtype=np.int64)
        scores = np.zeros((nq, topn), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.int32)
        status = np.zeros(nq, dtype=np.int32)
        errors: list = []
        cb = self._make_cb(infer_cb, errors)
        with self._lock:
            check(lib.ais_search(self._h, arr, nq, topn, prf_mode, cb, None, _ptr(ids), _ptr(scores), _ptr(counts),
                                 _ptr(status)))
        return ids, scores, counts, status, errors

    def search(self, queries: Sequence[Query], topn: int, prf_mode: int = PRF_STORED_ROWS,
               infer_cb: Optional[Callable] = None) -> List[List[Tuple[int, float]]]:
        """find_similar_documents for a batch of parsed queries; raises what the reference raises."""
        ids, scores, counts, status, errors = self.search_raw(queries, topn, prf_mode, infer_cb)
        if errors:
            raise errors[0][1]
        out = []
        for q in range(len(queries)):
            raise_for_status(int(status[q]))
            c = int(counts[q])
            out.append(list(zip(ids[q, :c].tolist(), scores[q, :c].tolist())))
        return out

    def rerank(self, final_scores: np.ndarray, topn: int, prf_mode: int = PRF_STORED_ROWS,
               infer_cb: Optional[Callable] = None) -> List[Tuple[int, float]]:
        """get_doc2vec_based_reranked_scores(final_scores, topn) webui.py:189-253."""
        f = np.ascontiguousarray(final_scores, dtype=np.float64)
        if len(f) != self.n_docs:
            raise ValueError("final_scores must have one entry per doc")
        ids = np.zeros(topn, dtype=np.int64)
        scores = np.zeros(topn, dtype=np.float64)
        cnt = C.c_int32(0)
        st = C.c_int32(0)
        errors: list = []
        cb = self._make_cb(infer_cb, errors)
        with self._lock:
            check(lib.ais_rerank(self._h, _ptr(f), topn, prf_mode, cb, None, _ptr(ids), _ptr(scores), C.byref(cnt),
                                 C.byref(st)))
        if errors:
            raise errors[0][1]
        raise_for_status(st.value)
        return list(zip(ids[:cnt.value].tolist(), scores[:cnt.value].tolist()))

    # ---- staged calls (torch tensors on this engine's device; see include/ais_b200.h) ----------
    def max_select_k(self) -> int:
        return int(lib.ais_max_select_k())

    def stage_score(self, queries: Sequence[Query], maxes) -> None:
        check(lib.ais_stage_score(self._h, _pack_queries(queries), len(queries), _ptr(maxes)))

    def stage_combine(self, nq: int, maxes, k: int, keys, ids) -> None:
        check(lib.ais_stage_combine(self._h, nq, _ptr(maxes), k, _ptr(keys), _ptr(ids)))

    def stage_top(self, nq: int, n_lists: int, k: int, keys, ids, want_host: bool, rows=None):
        depth = self.params.prf_depth
        top_ids = np.zeros((nq, depth), dtype=np.int64) if want_host else None
        top_scores = np.zeros((nq, depth), dtype=np.float64) if want_host else None
        check(lib.ais_stage_top(self._h, nq, n_lists, k, _ptr(keys), _ptr(ids), _ptr(top_ids), _ptr(top_scores), _ptr(rows)))
        return top_ids, top_scores

    def stage_set_status(self, status: np.ndarray) -> None:
        s = np.ascontiguousarray(status, dtype=np.int32)
        check(lib.ais_stage_set_status(self._h, len(s), _ptr(s)))

    def stage_requery(self, nq: int, q2: Optional[np.ndarray], rows, prf_mode: int, k: int, max_r, keys, ids) -> None:
        if q2 is not None:
            q2 = np.ascontiguousarray(q2, dtype=np.float32)
            assert q2.shape == (nq, DIM)
        check(lib.ais_stage_requery(self._h, nq, _ptr(q2), _ptr(rows), prf_mode, k, _ptr(max_r), _ptr(keys), _ptr(ids)))

    def stage_requery_select(self, nq: int, k: int, keys, ids) -> None:
        check(lib.ais_stage_requery_select(self._h, nq, k, _ptr(keys), _ptr(ids)))

    def stage_finish(self, nq: int, n_lists: int, k: int, keys, ids, max_r, topn: int, witness=None):
        """-> (ids, scores, counts, status, ambiguous, last_keys)"""
        out_ids = np.zeros((nq, topn), dtype=np.int64)
        out_scores = np.zeros((nq, topn), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.int32)
        status = np.zeros(nq, dtype=np.int32)
        amb = np.zeros(nq, dtype=np.int32)
        last = np.zeros(nq, dtype=np.uint64)
        check(lib.ais_stage_finish(self._h, nq, n_lists, k, _ptr(keys), _ptr(ids), _ptr(max_r), _ptr(witness), topn,
                                   _ptr(out_ids), _ptr(out_scores), _ptr(counts), _ptr(status), _ptr(amb), _ptr(last)))
        return out_ids, out_scores, counts, status, amb, last

    def stage_witness(self, amb: np.ndarray, last_keys: np.ndarray, second_pass: bool, max_r, witness) -> None:
        amb = np.ascontiguousarray(amb, dtype=np.int32)
        last_keys = np.ascontiguousarray(last_keys, dtype=np.uint64)
        check(lib.ais_stage_witness(self._h, len(amb), _ptr(amb), _ptr(last_keys), 1 if second_pass else 0, _ptr(max_r),
                                    _ptr(witness)))

    def stage_export_keys(self, query: int, second_pass: bool, keys, ids) -> None:
        check(lib.ais_stage_export_keys(self._h, query, 1 if second_pass else 0, _ptr(keys), _ptr(ids)))

    def sort_capacity(self, n: int) -> int:
        return int(lib.ais_sort_capacity(n))

    def stage_sort_finish(self, query: int, keys, ids, n_entries: int, max_r, topn: int):
        out_ids = np.zeros(topn, dtype=np.int64)
        out_scores = np.zeros(topn, dtype=np.float64)
        cnt = C.c_int32(0)
        st = C.c_int32(0)
        check(lib.ais_stage_sort_finish(self._h, query, _ptr(keys), _ptr(ids), n_entries, _ptr(max_r), topn, _ptr(out_ids),
                                        _ptr(out_scores), C.byref(cnt), C.byref(st)))
        return out_ids, out_scores, cnt.value, st.value

    def debug_read(self, which: str, query: int) -> np.ndarray:
        """Per-doc work array of the current batch: 'sim' | 'bm25' | 'fin' | 'rer' (test seam)."""
        code = {"sim": 0, "bm25": 1, "fin": 2, "rer": 3}[which]
        out = np.empty(self.n_docs, dtype=np.float32 if code in (0, 3) else np.float64)
        check(lib.ais_debug_read(self._h, code, query, _ptr(out)))
        return out

    # ---- introspection -----------------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        check(lib.ais_set_profiling(self._h, 1 if on else 0))

    def stats(self) -> dict:
        s = B.AisStats()
        check(lib.ais_get_stats(self._h, C.byref(s)))
        out = {}
        for name, _ in B.AisStats._fields_:
            v = getattr(s, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        out["kernels"] = {k: {"ms": out["kind_ms"][i], "brackets": out["kind_launches"][i]} for i, k in enumerate(B.KIND_NAMES)}
        return out

    def reset_stats(self) -> None:
        check(lib.ais_reset_stats(self._h))

    def synchronize(self) -> None:
        check(lib.ais_synchronize(self._h))
