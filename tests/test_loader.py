"""Native loader (ais_b200/loader.py + csrc/pickle_csr.cpp) - CPU tests: the C opcode walker over
`bm25_corpus` pickles written by the REFERENCE's own builder (tests/golden/bm25_files_main, oracle/make_bm25_files.py) and
by Python's pickle at every protocol, its fallbacks, and the gensim-free reader of the doc-vector index layout."""
import os
import pickle

import numpy as np
import pytest

from golden_util import GOLDEN, load_index
import ais_b200  # noqa: F401
from ais_b200 import binding as B, loader, synth


def test_reference_written_bm25_files_round_trip():
    """The five files genmodel.py:84-97 wrote for the "main" golden index -> CSR == the index the fixtures hold."""
    idx = load_index("main")
    bm = loader.read_bm25_index(os.path.join(GOLDEN, "bm25_files_main"))
    assert bm["D"] == idx.n_docs
    assert np.array_equal(bm["row_ptr"], idx.row_ptr)
    assert np.array_equal(bm["term_ids"], idx.term_ids) and np.array_equal(bm["tfs"], idx.tfs)
    assert np.array_equal(bm["doc_len"], idx.doc_len)
    assert bm["avgdl"] == float(idx.avgdl)
    assert np.array_equal(bm["idf"][: idx.vocab_size], idx.idf[: len(bm["idf"])]) and not bm["idf"][idx.df[: len(bm["idf"])] == 0].any()
    # and it is the walker, not the pickle.load fallback, that read the corpus
    n, nnz = B.C.c_int64(0), B.C.c_int64(0)
    assert B.lib.ais_pickle_csr_scan(os.fsencode(os.path.join(GOLDEN, "bm25_files_main", "bm25_corpus")), B.C.byref(n), B.C.byref(nnz)) == 0
    assert (n.value, nnz.value) == (idx.n_docs, len(idx.term_ids))


@pytest.mark.parametrize("protocol", [2, 3, 4, 5])
def test_walker_equals_pickle_load_for_every_protocol(tmp_path, protocol):
    idx = synth.generate_index(7000, vocab_size=70000 if protocol == 4 else 500, seed=3 + protocol, tf_gt1_fraction=0.05)
    corpus = idx.bm25_corpus()
    corpus[5] = {}                                   # a doc whose tags were all unknown to the dictionary (genmodel.py:59)
    corpus[6] = {2 ** 31 - 1: 300, 70000: 65536}     # BININT / BININT2 / LONG1-sized values
    path = str(tmp_path / "bm25_corpus")
    with open(path, "wb") as f:
        pickle.dump(corpus, f, protocol=protocol)
    row_ptr, term_ids, tfs = loader.read_bm25_corpus_csr(path)
    want = loader.corpus_to_csr(corpus)
    for a, b in zip((row_ptr, term_ids, tfs), want):
        assert np.array_equal(a, b)
    n, nnz = B.C.c_int64(0), B.C.c_int64(0)
    assert B.lib.ais_pickle_csr_scan(os.fsencode(path), B.C.byref(n), B.C.byref(nnz)) == 0 and n.value == len(corpus)


def test_walker_falls_back_and_reports(tmp_path):
    """numpy scalars as keys pickle through REDUCE: outside the subset -> AIS_ERR_UNSUPPORTED -> pickle.load fallback;
    a truncated file is an error, not a silent short read."""
    corpus = [{np.int64(3): 1, np.int64(9): 2}, {np.int64(1): 1}]
    path = str(tmp_path / "np_keys")
    with open(path, "wb") as f:
        pickle.dump(corpus, f)
    n, nnz = B.C.c_int64(0), B.C.c_int64(0)
    assert B.lib.ais_pickle_csr_scan(os.fsencode(path), B.C.byref(n), B.C.byref(nnz)) == B.AIS_ERR_UNSUPPORTED
    assert b"opcode" in B.lib.ais_pickle_last_error()
    row_ptr, term_ids, tfs = loader.read_bm25_corpus_csr(path)
    assert row_ptr.tolist() == [0, 2, 3] and term_ids.tolist() == [3, 9, 1] and tfs.tolist() == [1, 2, 1]
    good = str(tmp_path / "good")
    with open(good, "wb") as f:
        pickle.dump([{1: 2, 3: 4}] * 50, f)
    data = open(good, "rb").read()
    cut = str(tmp_path / "cut")
    with open(cut, "wb") as f:
        f.write(data[: len(data) // 2])
    with pytest.raises((B.AisError, pickle.UnpicklingError, EOFError)):       # the walker, or the fallback it hands over to
        loader.read_bm25_corpus_csr(cut)
    with pytest.raises(B.AisError):
        loader.read_bm25_corpus_csr(str(tmp_path / "missing"))


def test_gensim_layout_reader_without_gensim(tmp_path):
    """doc2vec_index (+ shards, one with its matrix split out as .index.npy, one inline) and doc2vec_dictionary written
    in gensim's layout (tests/make_gensim_layout.py; UNPINNED: no real gensim here) -> rows in doc order, token2id."""
    import sys
    from make_gensim_layout import write_index
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((9000 + 9000 + 700, 300)).astype(np.float32)       # 9000 x 1200 B = 10.8 MB > sep_limit
    t2i = {"t%d" % i: i for i in range(50)}
    files = write_index(str(tmp_path), rows, t2i, shardsize=9000)
    assert any(f.endswith(".index.npy") for f in files) and "gensim" not in sys.modules
    shards = list(loader.iter_similarity_shards(str(tmp_path / "doc2vec_index")))
    assert [s.shape[0] for s in shards] == [9000, 9000, 700]
    assert isinstance(shards[0], np.memmap) and not isinstance(shards[2], np.memmap)
    assert np.array_equal(np.concatenate(shards), rows)
    assert loader.read_token2id(str(tmp_path / "doc2vec_dictionary")) == t2i


@pytest.mark.gpu
def test_staged_from_files_equals_staged_from_arrays(tmp_path):
    """load_model()'s job end to end: the reference-format files -> engine -> same search results as the arrays."""
    from make_gensim_layout import write_index
    from ais_b200 import engine as E, query as Q
    idx = synth.generate_index(20000, vocab_size=800, seed=44, tf_gt1_fraction=0.01)
    write_index(str(tmp_path), idx.rows, idx.token2id, shardsize=9000)
    for name, obj in (("bm25_corpus", idx.bm25_corpus()), ("bm25_idf", idx.bm25_idf_dict()), ("bm25_avgdl", idx.avgdl),
                      ("bm25_D", idx.n_docs), ("bm25_doc_lengths", idx.doc_len)):
        with open(str(tmp_path / name), "wb") as f:
            pickle.dump(obj, f)
    a = E.SearchEngine(device=0, max_batch=8)
    loader.stage_index(a, str(tmp_path))
    b = E.SearchEngine.from_index(idx, max_batch=8)
    t2i = loader.read_token2id(str(tmp_path / "doc2vec_dictionary"))
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(t, t2i, infer) for t in synth.generate_queries(idx, 8, seed=2)]
    ra, rb = a.search_raw(qs, 100, E.PRF_STORED_ROWS), b.search_raw(qs, 100, E.PRF_STORED_ROWS)
    for x, y in zip(ra[:4], rb[:4]):
        assert np.array_equal(x, y)
    a.close(); b.close()


@pytest.mark.gpu
def test_load_model_end_to_end(tmp_path, monkeypatch):
    """webui_api.load_model() (webui.py:649-689) on a directory of reference-format files, gensim replaced only for
    doc2vec_model (MODEL_LOADER): find_similar_documents then equals the oracle, load_model is idempotent per CWD, and the
    engine was created for batched use (find_similar_documents_batch shares passes)."""
    from make_gensim_layout import write_index
    from gpu_util import ModelStub, assert_same_or_filter_unstable, capture
    from ais_b200 import webui_api
    from oracle import port
    idx = synth.generate_index(12000, vocab_size=600, seed=71)
    write_index(str(tmp_path), idx.rows, idx.token2id, shardsize=9000)
    for name, obj in (("bm25_corpus", idx.bm25_corpus()), ("bm25_idf", idx.bm25_idf_dict()), ("bm25_avgdl", idx.avgdl),
                      ("bm25_D", idx.n_docs), ("bm25_doc_lengths", idx.doc_len)):
        with open(str(tmp_path / name), "wb") as f:
            pickle.dump(obj, f)
    with open(str(tmp_path / webui_api.INDEX_CSV), "w", encoding="utf-8") as f:
        f.write("\n".join(idx.csv_lines()) + "\n")
    with open(str(tmp_path / "doc2vec_model"), "wb") as f:
        f.write(b"stub")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(webui_api, "MODEL_LOADER", lambda path: ModelStub(idx))
    monkeypatch.setattr(webui_api, "_engine", None)
    monkeypatch.setattr(webui_api, "_loaded_from", None)
    webui_api.load_model()
    eng = webui_api._engine
    webui_api.load_model()                                   # every search calls it (webui.py:585): must be a no-op now
    assert webui_api._engine is eng and eng.n_docs == idx.n_docs and webui_api.bm25_D == idx.n_docs
    assert eng.params.max_batch == webui_api.MAX_BATCH
    P = port.OraclePort(idx)
    texts = synth.generate_queries(idx, 10, seed=5)
    webui_api.PRF_MODE = "callback"
    for t in texts:
        assert_same_or_filter_unstable(capture(webui_api.find_similar_documents, t, 100), capture(P.find_similar_documents, t, 100),
                                       lambda: P.find_sorted(t), 1e-6, 100, t)
    ok = [t for t in texts if capture(P.find_similar_documents, t, 50)[0] == "ok"]
    batch = webui_api.find_similar_documents_batch(ok, 50)
    for t, got in zip(ok, batch):
        assert_same_or_filter_unstable(("ok", [d for d, _ in got], [s for _, s in got]), capture(P.find_similar_documents, t, 50),
                                       lambda: P.find_sorted(t), 1e-6, 50, t)
    eng.close()
    webui_api._engine = None
