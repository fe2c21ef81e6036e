"""Host-side logic of the product (query grammar, query vector, PRF vector, postings conversion,
shard bounds) against the golden fixtures made by running the reference."""
import warnings

import numpy as np
import pytest

from golden_util import load_index, load_results, load_seams
import ais_b200  # noqa: F401
from ais_b200 import query as Q, shard, synth, webui_api
from oracle import port

warnings.filterwarnings("ignore", category=RuntimeWarning)


@pytest.fixture(scope="module")
def main_index():
    return load_index("main")


def _infer(ix):
    t2i = ix.token2id
    return lambda words: ix.infer.one([t2i[w] for w in words if w in t2i])


def test_make_query_matches_reference_seams(main_index):
    z = load_seams()
    for k in range(4):
        q = Q.make_query(str(z["q%d_text" % k]), main_index.token2id, _infer(main_index))
        assert np.array_equal(q.vec, z["q%d_dense" % k])
        assert q.term_ids.tolist() == z["q%d_terms" % k].tolist()
        assert q.weights.tolist() == z["q%d_weights" % k].tolist()


def test_query_errors_match_reference(main_index):
    for rec in load_results("main")["results"]:
        if rec.get("error") in ("KeyError", ):
            with pytest.raises(KeyError) as ei:
                Q.make_query(rec["query"], main_index.token2id, _infer(main_index))
            assert str(ei.value) == rec["message"]
        elif rec.get("error") == "ValueError" and "invalid literal" in rec["message"]:
            with pytest.raises(ValueError) as ei:
                Q.make_query(rec["query"], main_index.token2id, _infer(main_index))
            assert str(ei.value) == rec["message"]


def test_prf_query_collapses_like_the_reference(main_index):
    P = port.OraclePort(main_index)
    vecs = [P.doc_vector_pairs(d + 1) for d in range(10)]
    w = list(np.linspace(1.0, 0.5, 10))
    a = Q.prf_query(vecs, w)
    b = P.prf_query(vecs, w)
    assert a == b
    assert {i for i, _ in a} == {0}
    dense = Q.dense_query(a)
    assert np.count_nonzero(dense) == 1 and dense[0] != 0


def test_corpus_to_postings_roundtrip(main_index):
    ix = main_index
    corpus = ix.bm25_corpus()
    pp, pd, pt = webui_api.corpus_to_postings(corpus, ix.vocab_size)
    rp, rd, rt = ix.postings()
    assert np.array_equal(pp, rp) and np.array_equal(pd, rd)
    assert pt is not None and np.array_equal(pt, rt)         # the main fixture holds tf > 1 docs


def test_shard_bounds_cover_everything():
    for n in (0, 1, 9, 10, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_builder_host_side_matches_reference_outputs(main_index):
    """tokens_to_csr / idf_table (the host half of the GPU BM25 builder) against the reference builder's outputs."""
    import os
    from golden_util import GOLDEN
    from ais_b200 import genmodel_api as G
    z = np.load(os.path.join(GOLDEN, "bm25_build_main.npz"))
    ix = main_index
    corpus = [[ix.tag_names[t] for t in seq] + ["unknown_tag"] for seq in ix.doc_tag_seq]
    ptr, ids = G.tokens_to_csr(corpus, ix.token2id)
    assert np.array_equal(np.diff(ptr), z["doc_lengths"])
    assert np.array_equal(ids, np.concatenate(ix.doc_tag_seq))
    idf = G.idf_table(ix.df, ix.n_docs)
    assert sorted(idf) == z["idf_terms"].tolist() and all(idf[t] == z["idf"][t] for t in idf)
