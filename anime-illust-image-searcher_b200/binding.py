"""ctypes view of the C ABI declared in include/ais_b200.h (libais_b200.so, built for sm_100a).

There is NO CPU path: if the shared library has not been built the import of this module
fails loudly, and ``ais_create`` fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AIS_B200_LIB") or os.path.join(_HERE, "libais_b200.so")   # override: instrumented debug builds

AIS_OK = 0
AIS_ERR_INVALID, AIS_ERR_CUDA, AIS_ERR_NOT_LOADED, AIS_ERR_UNSUPPORTED, AIS_ERR_CALLBACK = 1, 2, 3, 4, 5
AIS_Q_OK, AIS_Q_NAN_WEIGHTS, AIS_Q_ZERO_WEIGHT_SUM, AIS_Q_ZERO_VECTOR, AIS_Q_CALLBACK_FAILED = 0, 1, 2, 3, 4
AIS_PRF_CALLBACK, AIS_PRF_STORED_ROWS, AIS_PRF_STORED_ROWS_FULL, AIS_PRF_OFF = 0, 1, 2, 3
AIS_MAX_TERMS = 64
DIM = 300


class AisParams(C.Structure):
    _fields_ = [
        ("k1", C.c_double), ("b", C.c_double), ("bm25_weight", C.c_double), ("doc2vec_weight", C.c_double),
        ("original_score_weight", C.c_double), ("reranked_score_weight", C.c_double),
        ("diff_filter_thresh", C.c_double), ("require_magic", C.c_double),
        ("prf_depth", C.c_int32), ("max_batch", C.c_int32),
    ]


class AisQuery(C.Structure):
    _fields_ = [
        ("vec", C.POINTER(C.c_float)), ("term_ids", C.POINTER(C.c_int32)), ("weights", C.POINTER(C.c_double)),
        ("n_terms", C.c_int32),
    ]


class AisStats(C.Structure):
    _fields_ = [
        ("n_docs", C.c_int64), ("n_postings", C.c_int64), ("dim", C.c_int32), ("n_terms", C.c_int32),
        ("scan_launches", C.c_int64), ("scan_ms_total", C.c_double), ("kernel_launches", C.c_int64),
        ("fullsort_fallbacks", C.c_int64), ("bytes_device", C.c_int64), ("column_scan_launches", C.c_int64),
        ("tiles_per_seg", C.c_int64),
        ("kind_ms", C.c_double * 8), ("kind_launches", C.c_int64 * 8), ("bound_passes", C.c_int64), ("bitmap_batches", C.c_int64),
        ("pair_scan_launches", C.c_int64),
    ]


KIND_NAMES = ("scan", "bm25_slices", "bm25_score", "combine", "select", "requery", "tail", "witness")


INFER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int32,
                       C.POINTER(C.c_float))

_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/ais_b200.h declares
SIGNATURES = {
    "ais_last_error": (C.c_char_p, []),
    "ais_abi_version": (C.c_int, []),
    "ais_default_params": (None, [C.POINTER(AisParams)]),
    "ais_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(AisParams)]),
    "ais_destroy": (C.c_int, [_vp]),
    "ais_set_params": (C.c_int, [_vp, C.POINTER(AisParams)]),
    "ais_set_stream": (C.c_int, [_vp, _vp]),
    "ais_set_shard": (C.c_int, [_vp, _i64, _i64]),
    "ais_load_vectors": (C.c_int, [_vp, _vp, _i64, _i32, _i64]),
    "ais_reserve_docs": (C.c_int, [_vp, _i64]),
    "ais_vectors_device_ptr": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "ais_load_bm25": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _dbl]),
    "ais_build_bm25": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "ais_finish_bm25": (C.c_int, [_vp, _vp, _dbl]),
    "ais_export_postings": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ais_pickle_csr_scan": (C.c_int, [C.c_char_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "ais_pickle_csr_fill": (C.c_int, [C.c_char_p, _i64, _i64, _vp, _vp, _vp]),
    "ais_pickle_last_error": (C.c_char_p, []),
    "ais_dot_scores": (C.c_int, [_vp, _vp, _vp]),
    "ais_bm25_scores": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "ais_final_scores": (C.c_int, [_vp, C.POINTER(AisQuery), _vp]),
    "ais_search": (C.c_int, [_vp, C.POINTER(AisQuery), _i32, _i32, _i32, INFER_CB, _vp, _vp, _vp, _vp, _vp]),
    "ais_rerank": (C.c_int, [_vp, _vp, _i32, _i32, INFER_CB, _vp, _vp, _vp, _vp, _vp]),
    "ais_filter_sorted": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "ais_stage_score": (C.c_int, [_vp, C.POINTER(AisQuery), _i32, _vp]),
    "ais_stage_combine": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "ais_stage_top": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ais_stage_set_status": (C.c_int, [_vp, _i32, _vp]),
    "ais_stage_requery": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "ais_stage_requery_select": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "ais_stage_finish": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ais_stage_witness": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "ais_stage_export_keys": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "ais_stage_sort_finish": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ais_max_select_k": (C.c_int, []),
    "ais_sort_capacity": (_i64, [_i64]),
    "ais_debug_read": (C.c_int, [_vp, _i32, _i32, _vp]),
    "ais_set_profiling": (C.c_int, [_vp, C.c_int]),
    "ais_get_stats": (C.c_int, [_vp, C.POINTER(AisStats)]),
    "ais_reset_stats": (C.c_int, [_vp]),
    "ais_synchronize": (C.c_int, [_vp]),
}


class AisError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("ais_b200 error %d: %s" % (code, message))
        self.code = code


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  ais_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status: int) -> None:
    if status != AIS_OK:
        raise AisError(status, lib.ais_last_error().decode("utf-8", "replace"))
