"""TEST INFRASTRUCTURE.  Writes a doc-vector index and a dictionary in the ON-DISK LAYOUT of gensim 4.3.3 without gensim:

    doc2vec_index            pickle of a ``gensim.similarities.docsim.Similarity`` (attributes per its __init__ / save)
    doc2vec_index.N          pickle of a ``gensim.similarities.docsim.MatrixSimilarity`` per shard; a matrix above
                             ``SaveLoad``'s 10 MB sep_limit is NOT in the pickle but in ``doc2vec_index.N.index.npy``
                             (the pickle then lists the attribute under ``__numpys``), smaller ones are inline
    doc2vec_dictionary       pickle of a ``gensim.corpora.dictionary.Dictionary`` (token2id, id2token, dfs, cfs, ...)

The class PATHS in the pickles are gensim's (temporary stand-in modules are registered under those names while pickling
and removed afterwards), the attribute sets follow gensim's documented ``SaveLoad`` behaviour.
gensim itself is not installable in the build container, so this layout is UNPINNED until compared with files written by
a real gensim 4.3.3 (genmodel.py:155-156,175).
"""
from __future__ import annotations

import os
import pickle
import sys
import types

import numpy as np

SEP_LIMIT = 10 * 1024 ** 2          # gensim.utils.SaveLoad.save(sep_limit=10 MiB)


def _fake_gensim():
    mods = {}
    for name in ("gensim", "gensim.similarities", "gensim.similarities.docsim", "gensim.corpora", "gensim.corpora.dictionary",
                 "gensim.utils"):
        mods[name] = types.ModuleType(name)
    docsim, dictionary = mods["gensim.similarities.docsim"], mods["gensim.corpora.dictionary"]
    for cls_name, mod in (("Similarity", docsim), ("Shard", docsim), ("MatrixSimilarity", docsim), ("Dictionary", dictionary)):
        cls = type(cls_name, (object,), {"__module__": mod.__name__})
        setattr(mod, cls_name, cls)
    return mods


def write_index(dirpath: str, rows: np.ndarray, token2id: dict, shardsize: int = 32768, prefix: str = "doc2vec_index"):
    """rows fp32 [n, 300] -> the files above in `dirpath`; returns the list of files written."""
    mods = _fake_gensim()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    written = []
    try:
        docsim, dictionary = mods["gensim.similarities.docsim"], mods["gensim.corpora.dictionary"]
        shards = []
        for i, lo in enumerate(range(0, len(rows), shardsize)):
            m = np.ascontiguousarray(rows[lo: lo + shardsize], dtype=np.float32)
            fname = os.path.join(dirpath, "%s.%d" % (prefix, i))
            ms = docsim.MatrixSimilarity()
            ms.__dict__.update(num_features=m.shape[1], num_best=None, normalize=True, chunksize=256, corpus_len=len(m))
            if m.nbytes >= SEP_LIMIT:                       # SaveLoad._smart_save: large arrays go to <fname>.<attr>.npy
                np.save(fname + ".index.npy", m)
                written.append(fname + ".index.npy")
                ms.__dict__.update({"index": None, "__numpys": ["index"], "__scipys": [], "__ignoreds": [], "__recursive_saveloads": []})
                del ms.__dict__["index"]
            else:
                ms.__dict__.update({"index": m, "__numpys": [], "__scipys": [], "__ignoreds": [], "__recursive_saveloads": []})
            with open(fname, "wb") as f:
                pickle.dump(ms, f, protocol=4)
            written.append(fname)
            sh = docsim.Shard()
            # a path from ANOTHER machine: gensim re-bases it on load (Shard.fullname uses the current dirname)
            sh.__dict__.update(dirname="/somewhere/else", fname="%s.%d" % (prefix, i), length=len(m), cls=docsim.MatrixSimilarity)
            shards.append(sh)
        sim = docsim.Similarity()
        sim.__dict__.update(output_prefix=os.path.join("/somewhere/else", prefix), shardsize=shardsize, shards=shards, fresh_docs=[],
                            fresh_nnz=0, num_features=rows.shape[1], num_best=None, norm=False, chunksize=256, maintain_sparsity=False)
        with open(os.path.join(dirpath, prefix), "wb") as f:
            pickle.dump(sim, f, protocol=4)
        written.append(os.path.join(dirpath, prefix))
        d = dictionary.Dictionary()
        d.__dict__.update(token2id=dict(token2id), id2token={}, cfs={}, dfs={}, num_docs=len(rows), num_pos=0, num_nnz=0)
        with open(os.path.join(dirpath, "doc2vec_dictionary"), "wb") as f:
            pickle.dump(d, f)
        written.append(os.path.join(dirpath, "doc2vec_dictionary"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return written
