"""TEST INFRASTRUCTURE ONLY (oracle).  Runs the reference's OWN hot-path functions, verbatim.

``webui.py`` cannot be imported here (it imports gensim / streamlit / icecream / gen_cfeatures at
top level and calls ``main()`` on import, webui.py:5-20,788), but its hot-path functions are
pure numpy/Python.  This harness parses ``/root/reference/webui.py`` and ``genmodel.py`` with
``ast``, keeps ONLY the function definitions and constants of the path and ``exec``s them,
unchanged, in a namespace whose ``index`` / ``model`` / ``dictionary`` / ``ss`` globals are the
small stubs below.  Nothing from the reference is copied into this repository: the source is
read where it lies, at run time, and only in the build container (``/root/reference`` does not
exist on the GPU box - callers must check ``available()``).

It is used (a) to generate the committed golden fixtures (oracle/make_golden.py) and (b) to
validate oracle/port.py.  It is never imported by the product package.
"""
from __future__ import annotations

import ast
import math
import os
import pickle
import typing
import warnings
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .gensim_stub import SimilarityStub

REFERENCE_DIR = os.environ.get("AIS_REFERENCE_DIR", "/root/reference")

WEBUI_FUNCS = (
    "filter_searched_result",                 # webui.py:63-80
    "normalize_and_apply_weight_doc2vec",     # webui.py:82-117
    "compute_bm25_scores",                    # webui.py:119-172
    "get_embedded_vector_by_doc_id",          # webui.py:182-187
    "get_doc2vec_based_reranked_scores",      # webui.py:189-253
    "find_similar_documents",                 # webui.py:345-390
)
WEBUI_CONSTS = (
    "BM25_WEIGHT", "DOC2VEC_WEIGHT", "ORIGINAL_SCORE_WEIGHT", "RERANKED_SCORE_WEIGHT",
    "DIFF_FILTER_THRESH", "REQUIRE_TAG_MAGIC_NUMBER",          # webui.py:51-60
)
GENMODEL_FUNCS = ("gen_and_save_bm25_index",)  # genmodel.py:51-99


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "webui.py"))


def _extract(path: str, func_names: Sequence[str], const_names: Sequence[str]) -> ast.Module:
    with open(path, "r", encoding="utf-8") as f:
        src = f.read()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)   # '\(' escapes at webui.py:92-98
        tree = ast.parse(src, filename=path)
    keep: List[ast.stmt] = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in func_names:
            keep.append(node)
        elif isinstance(node, ast.Assign):
            if any(isinstance(t, ast.Name) and t.id in const_names for t in node.targets):
                keep.append(node)
        elif isinstance(node, ast.AnnAssign):
            if isinstance(node.target, ast.Name) and node.target.id in const_names:
                keep.append(node)
    found = {n.name for n in keep if isinstance(n, ast.FunctionDef)}
    missing = set(func_names) - found
    if missing:
        raise RuntimeError("reference functions not found in %s: %s" % (path, sorted(missing)))
    return ast.Module(body=keep, type_ignores=[])


class _ModelStub:
    """``model`` of webui.py:26: ``infer_vector(list_of_tag_strings)`` and ``dv[0]`` (webui.py:104,106,185)."""

    def __init__(self, infer, token2id: Dict[str, int], dim: int):
        self._infer = infer
        self._token2id = token2id
        self.dv = [np.zeros(dim, dtype=np.float32)]
        self.calls: List[List[str]] = []

    def infer_vector(self, words: List[str]) -> np.ndarray:
        self.calls.append(list(words))
        ids = [self._token2id[w] for w in words if w in self._token2id]   # gensim skips unknown words
        return self._infer.one(ids)


class _DictionaryStub:
    def __init__(self, token2id: Dict[str, int]):
        self.token2id = token2id


class ReferenceWorld:
    """The reference's module globals + functions, populated from a synthetic index."""

    def __init__(self, idx, use_reference_bm25_builder: bool = True, workdir: Optional[str] = None):
        if not available():
            raise RuntimeError("reference sources not present at %s" % REFERENCE_DIR)
        token2id = idx.token2id
        ns: Dict[str, Any] = {
            "np": np, "math": math, "ndarray": np.ndarray, "pickle": pickle,
            "List": typing.List, "Tuple": typing.Tuple, "Dict": typing.Dict, "Any": typing.Any,
            "Optional": typing.Optional,
        }
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            code = compile(_extract(os.path.join(REFERENCE_DIR, "webui.py"), WEBUI_FUNCS, WEBUI_CONSTS),
                           "webui.py<extracted>", "exec")
            exec(code, ns)
        self.ns = ns
        self.index = SimilarityStub(idx.rows)
        self.model = _ModelStub(idx.infer, token2id, idx.rows.shape[1])
        ns["index"] = self.index
        ns["model"] = self.model
        ns["dictionary"] = _DictionaryStub(token2id)
        ns["ss"] = {"search_mode": "normal"}
        ns["image_files_name_tags_arr"] = idx.csv_lines()
        if use_reference_bm25_builder:
            self._build_bm25_with_reference(idx, workdir)
        else:
            ns["bm25_corpus"] = idx.bm25_corpus()
            ns["bm25_doc_lengths"] = idx.doc_len
            ns["bm25_avgdl"] = idx.avgdl
            ns["bm25_idf"] = idx.bm25_idf_dict()
            ns["bm25_D"] = idx.n_docs

    def _build_bm25_with_reference(self, idx, workdir: Optional[str]) -> None:
        """Run genmodel.gen_and_save_bm25_index (genmodel.py:51-99) on the docs' tag strings and
        load the five pickles the way load_model does (webui.py:680-684)."""
        import tempfile
        gns: Dict[str, Any] = {"np": np, "pickle": pickle, "List": typing.List,
                               "corpora": type("corpora", (), {"Dictionary": object})}
        exec(compile(_extract(os.path.join(REFERENCE_DIR, "genmodel.py"), GENMODEL_FUNCS, ()),
                     "genmodel.py<extracted>", "exec"), gns)
        gns["print"] = lambda *a, **k: None
        corpus = [[idx.tag_names[t] for t in idx.doc_tags(d)] for d in range(idx.n_docs)]
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory(dir=workdir) as tmp:
            os.chdir(tmp)
            try:
                gns["gen_and_save_bm25_index"](corpus, self.ns["dictionary"])
                for name, key in (("bm25_corpus", "bm25_corpus"), ("bm25_doc_lengths", "bm25_doc_lengths"),
                                  ("bm25_avgdl", "bm25_avgdl"), ("bm25_idf", "bm25_idf"), ("bm25_D", "bm25_D")):
                    with open(name, "rb") as f:
                        self.ns[key] = pickle.load(f)
            finally:
                os.chdir(cwd)

    # --- reference callables -------------------------------------------------
    def set_const(self, name: str, value: float) -> None:
        assert name in WEBUI_CONSTS
        self.ns[name] = value

    def __getattr__(self, name: str):
        ns = object.__getattribute__(self, "ns")
        if name in WEBUI_FUNCS or name in WEBUI_CONSTS or name.startswith("bm25_"):
            return ns[name]
        raise AttributeError(name)
