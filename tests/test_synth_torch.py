"""The on-device generator used by bench.py, exercised on CPU tensors: shards of any world size
add up to the same corpus, posting lists are well formed."""
import numpy as np
import torch

import ais_b200  # noqa: F401
from ais_b200 import shard, synth_torch


def _gen(lo, hi, vocab=500):
    rows = torch.zeros((hi - lo, 300), dtype=torch.float32)
    sh = synth_torch.generate_shard(lo, hi, rows, vocab=vocab, seed=11, chunk=8192)
    return rows, sh


def test_shards_add_up_to_the_same_corpus():
    n = 2 * 8192 + 3000
    rows1, one = _gen(0, n)
    parts = [_gen(*shard.shard_bounds(n, 3, r)) for r in range(3)]
    assert torch.equal(torch.cat([p[0] for p in parts]), rows1)
    assert torch.equal(sum(p[1].df for p in parts), one.df)
    assert torch.equal(torch.cat([p[1].doc_len for p in parts]), one.doc_len)
    # posting lists: ascending local doc ids inside every term, each (term, doc) once
    for rows, sh in parts + [(rows1, one)]:
        ptr = sh.post_ptr.numpy()
        doc = sh.post_doc.numpy()
        assert ptr[0] == 0 and ptr[-1] == len(doc) == int(sh.doc_len.sum())
        seg_start = np.zeros(len(doc), dtype=bool)
        seg_start[ptr[:-1][ptr[:-1] < len(doc)]] = True
        inc = np.diff(doc) > 0
        assert np.all(inc | seg_start[1:])
        assert doc.min() >= 0 and doc.max() < sh.n_docs
    dl = one.doc_len.numpy()
    assert dl.min() >= 3 and dl.max() <= 120 and 24 < dl.mean() < 34


def test_queries_are_well_formed():
    rows, sh = _gen(0, 20000)
    E = synth_torch.embedding_table(500, 11, torch.device("cpu")).numpy()
    texts, parsed = synth_torch.make_queries(sh.df.numpy(), E, 50, seed=3)
    assert len(texts) == len(parsed) == 50
    for vec, terms, weights in parsed:
        assert vec.dtype == np.float32 and abs(np.linalg.norm(vec) - 1) < 1e-5
        assert len(terms) == len(weights) == len(set(terms.tolist())) >= 1
