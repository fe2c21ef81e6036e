"""-m gpu: the GPU BM25 index builder (ais_build_bm25, genmodel.py:51-99) against what the reference's own
gen_and_save_bm25_index produced for the golden corpus, and against the numpy transposition on larger corpora."""
import os
import pickle

import numpy as np
import pytest

from golden_util import GOLDEN, load_index
import ais_b200  # noqa: F401
from ais_b200 import engine as E, genmodel_api as G, synth

pytestmark = pytest.mark.gpu


def _seq_csr(idx):
    ptr = np.cumsum([0] + [len(s) for s in idx.doc_tag_seq]).astype(np.int64)
    ids = np.concatenate(idx.doc_tag_seq).astype(np.int32) if idx.n_docs else np.zeros(0, np.int32)
    return ptr, ids


def test_build_matches_reference_builder_outputs():
    idx = load_index("main")
    z = np.load(os.path.join(GOLDEN, "bm25_build_main.npz"))
    eng = E.SearchEngine()
    ptr, ids = _seq_csr(idx)
    doc_len, avgdl, idf, df = G.build_index(eng, ptr, ids, idx.vocab_size)
    assert np.array_equal(doc_len, z["doc_lengths"])                       # genmodel.py:69,75
    assert avgdl == z["avgdl"] and type(avgdl) is np.float64                # genmodel.py:76, bit-equal
    assert sorted(idf.keys()) == z["idf_terms"].tolist()
    assert all(idf[t] == z["idf"][t] for t in idf)                          # genmodel.py:79-82, bit-equal
    # bm25_corpus (genmodel.py:64-68) back from the device posting lists
    pp, pd, pt = eng.export_postings()
    got = [dict() for _ in range(idx.n_docs)]
    terms = np.repeat(np.arange(idx.vocab_size), np.diff(pp))
    for t, d, f in zip(terms.tolist(), pd.tolist(), pt.tolist()):
        got[d][t] = f
    cptr, cterms, ctfs = z["corpus_ptr"], z["corpus_terms"], z["corpus_tfs"]
    want = [dict(zip(cterms[cptr[i]: cptr[i + 1]].tolist(), ctfs[cptr[i]: cptr[i + 1]].tolist())) for i in range(idx.n_docs)]
    assert got == want
    assert int(ctfs.max()) > 1                                              # the fixture exercises tf > 1
    for t in range(idx.vocab_size):                                         # ascending doc ids inside every list
        seg = pd[pp[t]: pp[t + 1]]
        assert np.all(np.diff(seg) > 0)


def test_built_index_scores_like_the_loaded_one():
    idx = synth.generate_index(30000, vocab_size=1500, seed=404, tf_gt1_fraction=0.02)
    loaded = E.SearchEngine.from_index(idx)
    built = E.SearchEngine()
    ptr, ids = _seq_csr(idx)
    doc_len, avgdl, idf, df = G.build_index(built, ptr, ids, idx.vocab_size)
    assert np.array_equal(df, idx.df) and np.array_equal(doc_len, idx.doc_len) and avgdl == idx.avgdl
    rp, rd, rt = idx.postings()
    pp, pd, pt = built.export_postings()
    assert np.array_equal(pp, rp) and np.array_equal(pd, rd) and np.array_equal(pt, rt)
    rng = np.random.default_rng(1)
    for _ in range(8):
        terms = rng.choice(np.nonzero(idx.df)[0], size=4, replace=False)
        w = np.array([2.0, 1003.0, -1.0, 1.0])
        a = loaded.bm25_scores(terms, w)
        b = built.bm25_scores(terms, w)
        assert np.array_equal(a, b)


def test_genmodel_drop_in_writes_the_reference_pickles(tmp_path):
    idx = load_index("main")
    z = np.load(os.path.join(GOLDEN, "bm25_build_main.npz"))
    corpus = [[idx.tag_names[t] for t in seq] + (["not_in_dictionary"] if i % 7 == 0 else []) for i, seq in enumerate(idx.doc_tag_seq)]

    class Dict_:
        token2id = idx.token2id
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        G.gen_and_save_bm25_index(corpus, Dict_())
        got = {n: pickle.load(open(n, "rb")) for n in ("bm25_corpus", "bm25_idf", "bm25_avgdl", "bm25_D", "bm25_doc_lengths")}
    finally:
        os.chdir(cwd)
    assert got["bm25_D"] == int(z["D"]) and got["bm25_avgdl"] == z["avgdl"]
    assert np.array_equal(got["bm25_doc_lengths"], z["doc_lengths"])
    assert {t: float(v) for t, v in got["bm25_idf"].items()} == {int(t): float(z["idf"][t]) for t in z["idf_terms"]}
    cptr, cterms, ctfs = z["corpus_ptr"], z["corpus_terms"], z["corpus_tfs"]
    for i in range(idx.n_docs):
        want = dict(zip(cterms[cptr[i]: cptr[i + 1]].tolist(), ctfs[cptr[i]: cptr[i + 1]].tolist()))
        assert got["bm25_corpus"][i] == want
        assert list(got["bm25_corpus"][i].keys()) == list(want.keys())      # same insertion order as the reference's dict


def test_build_rejects_bad_input():
    from ais_b200.binding import AisError
    eng = E.SearchEngine()
    with pytest.raises(AisError):
        eng.build_bm25(np.array([0, 2], np.int64), np.array([0, 99], np.int32), 10)      # term id out of range
    with pytest.raises(AisError):
        eng.build_bm25(np.array([0, 2049], np.int64), np.arange(2049, dtype=np.int32) % 400, 400)  # longer than a sort chunk


def test_build_docs_with_long_tag_lists():
    """no 256-tags cap.  Docs of 300 / 1500 / 2048 tokens (with repeats) next to ordinary ones: doc lengths,
    document frequencies and the (term, doc, tf) postings equal the dict-based transposition of genmodel.py:64-73."""
    rng = np.random.default_rng(77)
    V = 900
    lens = [30, 300, 12, 1500, 2048, 0, 45, 257, 31]
    seqs = [rng.integers(0, V, size=n).astype(np.int32) for n in lens]
    ptr = np.cumsum([0] + lens).astype(np.int64)
    ids = np.concatenate(seqs)
    eng = E.SearchEngine()
    doc_len, avgdl, idf, df = G.build_index(eng, ptr, ids, V)
    assert doc_len.tolist() == lens
    want = {}
    for d, s in enumerate(seqs):
        terms, tfs = np.unique(s, return_counts=True)
        for t, f in zip(terms.tolist(), tfs.tolist()):
            want[(t, d)] = f
    pp, pd, pt = eng.export_postings()
    terms = np.repeat(np.arange(V), np.diff(pp))
    got = {(t, d): f for t, d, f in zip(terms.tolist(), pd.tolist(), pt.tolist())}
    assert got == want
    assert np.array_equal(df, np.bincount([t for t, _ in want], minlength=V))
    for t in range(V):
        assert np.all(np.diff(pd[pp[t]: pp[t + 1]]) > 0)
