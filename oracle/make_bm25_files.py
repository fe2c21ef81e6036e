"""TEST INFRASTRUCTURE ONLY (oracle).  Writes tests/golden/bm25_files_main/{bm25_corpus,bm25_idf,bm25_avgdl,bm25_D,
bm25_doc_lengths} by RUNNING the reference's own builder (genmodel.py:51-99, AST-extracted by oracle/verbatim.py and
executed unchanged) on the docs of the committed "main" golden index - the bytes on disk are what the reference's
`pickle.dump` calls produce (genmodel.py:84-97).  tests/test_loader.py reads them back with the native loader.

Run in the build container (needs /root/reference):   python -m oracle.make_bm25_files
"""
from __future__ import annotations

import os
import pickle
import shutil
import sys
import typing

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import verbatim  # noqa: E402


def main():
    from golden_util import load_index
    idx = load_index("main")
    out = os.path.join(ROOT, "tests", "golden", "bm25_files_main")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    gns = {"np": np, "pickle": pickle, "List": typing.List, "corpora": type("corpora", (), {"Dictionary": object})}
    exec(compile(verbatim._extract(os.path.join(verbatim.REFERENCE_DIR, "genmodel.py"), verbatim.GENMODEL_FUNCS, ()),
                 "genmodel.py<extracted>", "exec"), gns)
    gns["print"] = lambda *a, **k: None
    corpus = [[idx.tag_names[t] for t in idx.doc_tags(d)] for d in range(idx.n_docs)]
    cwd = os.getcwd()
    os.chdir(out)
    try:
        gns["gen_and_save_bm25_index"](corpus, verbatim._DictionaryStub(idx.token2id))
    finally:
        os.chdir(cwd)
    print({f: os.path.getsize(os.path.join(out, f)) for f in sorted(os.listdir(out))})


if __name__ == "__main__":
    main()
