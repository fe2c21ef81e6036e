"""The C-ABI library loads and exports every symbol include/ais_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

import ais_b200  # noqa: F401
from ais_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "ais_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    names = set(re.findall(r"\b(ais_[a-z0-9_]+)\s*\(", text))
    return names - {"ais_infer_cb"}


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(binding.LIB_PATH)
    for n in sorted(names):
        assert hasattr(raw, n), "libais_b200.so does not export %s" % n
        assert n in binding.SIGNATURES, "binding.py does not bind %s" % n
    assert set(binding.SIGNATURES) <= names, set(binding.SIGNATURES) - names


def test_abi_version_and_defaults():
    assert binding.lib.ais_abi_version() == 1
    p = binding.AisParams()
    binding.lib.ais_default_params(ctypes.byref(p))
    # webui.py:51-60,126-127,193-195
    assert (p.k1, p.b, p.bm25_weight, p.doc2vec_weight) == (1.5, 0.75, 0.5, 0.5)
    assert (p.original_score_weight, p.reranked_score_weight, p.diff_filter_thresh, p.require_magic) == (0.7, 0.3, 1e-6, 1000.0)
    assert (p.prf_depth, p.max_batch) == (10, 1)
    assert binding.lib.ais_max_select_k() == 1024
    assert binding.lib.ais_sort_capacity(5) == 2048 and binding.lib.ais_sort_capacity(2049) == 4096


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ais_b200 import engine
    with pytest.raises(binding.AisError) as ei:
        engine.SearchEngine()
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "anime-illust-image-searcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
