"""-m gpu: the CUDA path, called through the C ABI / the webui.py-compatible module, against
(a) the golden fixtures recorded from the reference itself and (b) the oracle port on fresh indexes."""
import numpy as np
import pytest

from golden_util import load_filter_cases, load_index, load_results, load_seams
from gpu_util import SCORE_RTOL, assert_same, assert_same_or_filter_unstable, capture, install
import ais_b200  # noqa: F401
from ais_b200 import engine as E, query as Q, synth, webui_api
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def main_index():
    return load_index("main")


@pytest.mark.parametrize("prf_mode", ["callback", "stored_rows"])
def test_golden_results_main(main_index, prf_mode):
    """Every recorded find_similar_documents outcome of the reference: ids, order, scores, exceptions."""
    install(main_index, prf_mode=prf_mode)
    for rec in load_results("main")["results"]:
        got = capture(webui_api.find_similar_documents, rec["query"], rec["topn"])
        want = ("err", rec["error"], rec["message"]) if "error" in rec else ("ok", rec["ids"], rec["scores"])
        assert_same(got, want, rec["query"])


def test_golden_results_tiny_no_prf_branch():
    """N <= 10: webui.py:247-253."""
    install(load_index("tiny"))
    for rec in load_results("tiny")["results"]:
        got = capture(webui_api.find_similar_documents, rec["query"], rec["topn"])
        assert_same(got, ("ok", rec["ids"], rec["scores"]), rec["query"])


def test_golden_seams(main_index):
    """index[vec], compute_bm25_scores and the combined scores against the reference's own vectors."""
    eng = install(main_index)
    z = load_seams()
    for k in range(4):
        text = str(z["q%d_text" % k])
        vec = webui_api.normalize_and_apply_weight_doc2vec(text)
        sims = webui_api.index[vec]
        ref = z["q%d_sims" % k]
        assert sims.dtype == np.float32 and sims.shape == ref.shape
        assert np.abs(sims - ref).max() <= 1e-5 * np.abs(ref).max()
        wts = dict(zip(z["q%d_terms" % k].tolist(), z["q%d_weights" % k].tolist()))
        bm25 = webui_api.compute_bm25_scores(query_weights=wts)
        assert bm25.dtype == np.float64
        assert np.array_equal(bm25, z["q%d_bm25" % k])                    # bit-exact, -inf masks included
        q = Q.make_query(text, main_index.token2id, webui_api.model.infer_vector)
        fin = eng.final_scores(q)
        rf = z["q%d_final" % k]
        assert np.array_equal(np.isneginf(fin), np.isneginf(rf))          # masks bit-exact
        live = ~np.isneginf(rf)
        assert np.abs(fin[live] - rf[live]).max() <= 1e-5 * np.abs(rf[live]).max()
    names = z["terms_form_tags"].tolist()
    assert np.array_equal(webui_api.compute_bm25_scores(query_terms=names), z["terms_form_bm25"])


def test_filter_known_answers(main_index):
    install(main_index)
    for c in load_filter_cases():
        res = webui_api.filter_searched_result(c["input"])
        assert [d for d, _ in res] == c["ids"]
        assert [float(s) for _, s in res] == c["scores"]


def test_rerank_seam_matches_oracle(main_index):
    install(main_index)
    P = port.OraclePort(main_index)
    z = load_seams()
    for k in range(4):
        fin = z["q%d_final" % k]
        for topn in (5, 100, 800):
            got = capture(webui_api.get_doc2vec_based_reranked_scores, fin, topn)
            want = capture(P.rerank, fin, topn)
            assert_same(got, want, "rerank q%d topn %d" % (k, topn))


@pytest.mark.parametrize("n_docs,vocab,tf_frac", [(50000, 3000, 0.0), (7000, 400, 0.03), (33, 20, 0.0), (11, 12, 0.0)])
def test_fresh_index_vs_oracle(n_docs, vocab, tf_frac):
    idx = synth.generate_index(n_docs, vocab_size=vocab, seed=1234 + n_docs, tf_gt1_fraction=tf_frac)
    P = port.OraclePort(idx)
    queries = synth.generate_queries(idx, 24, seed=n_docs)
    for prf_mode in ("callback", "stored_rows"):
        install(idx, prf_mode=prf_mode)
        for q in queries:
            for topn in ((100, 800) if n_docs > 1000 else (100,)):
                assert_same_or_filter_unstable(capture(webui_api.find_similar_documents, q, topn),
                                               capture(P.find_similar_documents, q, topn),
                                               lambda: P.find_sorted(q), 1e-6, topn, q)


def test_config1_10k_docs_1024_queries_batched():
    """BASELINE.json configs[0] / SURVEY.md 8(d) config 1: 10^4 docs, V = 10 861, 1 024 weighted queries with +required /
    -exclude tags, top-100, PRF on - every result of the batched engine path (64 queries per pass: tcgen05 scan, BM25
    records, tile-filtered select) against the oracle's find_similar_documents: ids, order, scores, exceptions."""
    from ais_b200.engine import raise_for_status
    idx = synth.generate_index(10000, vocab_size=10861, seed=20260101)
    P = port.OraclePort(idx)
    texts = synth.generate_queries(idx, 1024, seed=1)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(t, t2i, infer) for t in texts]
    eng = E.SearchEngine.from_index(idx, max_batch=64)
    n_err = 0
    for b0 in range(0, len(qs), 64):
        ids, scores, counts, status, _ = eng.search_raw(qs[b0:b0 + 64], 100, E.PRF_STORED_ROWS)
        for j in range(len(ids)):
            text = texts[b0 + j]
            want = capture(P.find_similar_documents, text, 100)
            try:
                raise_for_status(int(status[j]))
                c = int(counts[j])
                got = ("ok", ids[j, :c].tolist(), scores[j, :c].tolist())
            except Exception as e:   # noqa: BLE001
                got = ("err", type(e).__name__, str(e))
                n_err += 1
            assert_same_or_filter_unstable(got, want, lambda: P.find_sorted(text), 1e-6, 100, text)
    eng.close()
    assert n_err < 1024            # the generator rejects most queries that leave fewer than 10 docs


def test_batched_queries_equal_single_queries():
    """<= 4 queries per pass run the same fp32 SIMT scan (bit-equal results for any batching); >= 5 queries per
    pass run a tensor-core scan (3xTF32, fp32-level accuracy; mma.sync up to 16, tcgen05 beyond): same ranking within
    the score tolerance."""
    from gpu_util import same_ranking
    idx = synth.generate_index(40000, vocab_size=2000, seed=77)
    queries = synth.generate_queries(idx, 37, seed=3)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(q, t2i, infer) for q in queries]
    single = E.SearchEngine.from_index(idx, max_batch=1)
    ref = single.search_raw(qs, 100, E.PRF_STORED_ROWS)
    for mb in (2, 3, 4, 8, 13, 16, 29, 37):
        eng = E.SearchEngine.from_index(idx, max_batch=mb)
        got = eng.search_raw(qs, 100, E.PRF_STORED_ROWS)
        assert np.array_equal(got[3], ref[3])
        for q in range(len(qs)):
            c = ref[2][q]
            if mb <= 4:
                assert got[2][q] == c
                assert np.array_equal(got[0][q, :c], ref[0][q, :c]), (mb, q)
                assert np.array_equal(got[1][q, :c], ref[1][q, :c]), (mb, q)      # same kernels, same order: bit-equal
            elif got[2][q] == c:
                msg = same_ranking(got[0][q, :c].tolist(), got[1][q, :c].tolist(), ref[0][q, :c].tolist(), ref[1][q, :c].tolist())
                assert msg is None, (mb, q, msg)
            else:       # a near-threshold gap of filter_searched_result fell on the other side (see gpu_util)
                m = min(c, got[2][q])
                assert same_ranking(got[0][q, :m].tolist(), got[1][q, :m].tolist(), ref[0][q, :m].tolist(), ref[1][q, :m].tolist()) is None
        eng.close()


@pytest.mark.parametrize("mb", [8, 16, 24, 32, 47, 64, 100, 200])
def test_tensor_core_scan_accuracy(mb):
    """index[vec] through the batched scans against the fp64 dot product, every doc of every query: fp32-level error.
    8/16: mma.sync 3xTF32 (scan.cuh); 17..32: tcgen05 32 queries per pass; > 32: tcgen05 64 per pass (scan_tc.cuh).
    The bound is what numpy's own fp32 dot product (the reference's arithmetic) stays within, too."""
    import torch
    n_docs = 128 * 148 + 77                                   # ragged last tile, more tiles than CTAs
    idx = synth.generate_index(n_docs, vocab_size=500, seed=12)
    rng = np.random.default_rng(4 + mb)
    eng = E.SearchEngine.from_index(idx, max_batch=mb)
    vecs = rng.standard_normal((mb, 300)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    qs = [E.Query(v, np.array([0], np.int32), np.array([1.0])) for v in vecs]
    maxes = torch.empty((mb, 2), dtype=torch.float64, device="cuda")
    eng.stage_score(qs, maxes)
    torch.cuda.synchronize()
    eng.synchronize()
    want = idx.rows.astype(np.float64) @ vecs.T.astype(np.float64)          # [n][mb]
    scale = np.abs(want).max(axis=0)
    got_max = maxes.cpu().numpy()[:, 1]
    assert np.abs(got_max - want.max(axis=0)).max() <= 2e-6 * scale.max()
    for q in range(mb):
        got = eng.debug_read("sim", q)
        assert np.abs(got - want[:, q]).max() <= 2e-6 * scale[q], (mb, q)
    eng.close()


@pytest.mark.parametrize("n_batch", [12, 40])
@pytest.mark.parametrize("n_docs", [11, 33, 127, 129, 300])
def test_batched_path_on_tiny_indexes(n_docs, n_batch):
    """Ragged / single-tile shards through the batched kernels (tcgen05 scan with one partial 128-row tile - 40 queries: the
    CTA-pair kernel, whose second CTA then works on a tile that lies wholly beyond the shard -, BM25 records and tile
    maxima of a partial 256-doc tile): a batch returns what single-query searches return."""
    from gpu_util import same_ranking
    idx = synth.generate_index(n_docs, vocab_size=40, seed=500 + n_docs)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = []
    for text in synth.generate_queries(idx, 40 + 3 * n_batch, seed=n_docs):
        try:
            qs.append(Q.make_query(text, t2i, infer))
        except KeyError:             # the special tag "3:4" parses as tag "3", weight 4 (webui.py:358-371): host-side error
            continue
    qs = qs[:n_batch]
    assert len(qs) == n_batch
    single = E.SearchEngine.from_index(idx, max_batch=1)
    many = E.SearchEngine.from_index(idx, max_batch=n_batch)
    for mode in (E.PRF_STORED_ROWS, E.PRF_STORED_ROWS_FULL, E.PRF_OFF):
        ref = [single.search_raw([q], 100, mode) for q in qs]
        got = many.search_raw(qs, 100, mode)
        for j in range(len(qs)):
            assert got[3][j] == ref[j][3][0], (mode, j)
            if got[3][j] != 0:
                continue
            c = int(ref[j][2][0])
            if int(got[2][j]) != c:          # a near-threshold gap of filter_searched_result fell on the other side
                c = min(c, int(got[2][j]))
            msg = same_ranking(got[0][j, :c].tolist(), got[1][j, :c].tolist(), ref[j][0][0, :c].tolist(), ref[j][1][0, :c].tolist())
            assert msg is None, (mode, j, msg)
    single.close()
    many.close()


def test_dense_requery_modes_match_the_oracle_full_centroid():
    """AIS_PRF_STORED_ROWS_FULL (the un-collapsed centroid: a dense second pass) at batch sizes that take the fp32 SIMT
    scan and the tcgen05 scan: same ranking."""
    from gpu_util import same_ranking
    idx = synth.generate_index(30000, vocab_size=1500, seed=17)
    queries = synth.generate_queries(idx, 40, seed=2)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(q, t2i, infer) for q in queries]
    single = E.SearchEngine.from_index(idx, max_batch=1)
    ref = [single.search_raw([q], 100, E.PRF_STORED_ROWS_FULL) for q in qs]
    eng = E.SearchEngine.from_index(idx, max_batch=40)
    got = eng.search_raw(qs, 100, E.PRF_STORED_ROWS_FULL)
    assert eng.stats()["column_scan_launches"] == 0
    n_same = 0
    for j in range(len(qs)):
        assert got[3][j] == ref[j][3][0]
        c = int(ref[j][2][0])
        if got[3][j] != 0 or int(got[2][j]) != c:
            continue
        msg = same_ranking(got[0][j, :c].tolist(), got[1][j, :c].tolist(), ref[j][0][0, :c].tolist(), ref[j][1][0, :c].tolist())
        assert msg is None, (j, msg)
        n_same += 1
    assert n_same >= 30
    single.close()
    eng.close()


def test_requery_column_scan_equals_dense_scan(monkeypatch):
    """The reference's PRF re-query vector is [c, 0, ..., 0]; the engine serves it with a column scan
    (one sector per doc).  Same bits as the dense fp32 scan of that vector, same search results, and it is counted."""
    idx = synth.generate_index(30000, vocab_size=1500, seed=91)
    queries = synth.generate_queries(idx, 8, seed=5)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(q, t2i, infer) for q in queries]
    out = {}
    for dense in ("1", "0"):
        monkeypatch.setenv("AIS_REQUERY_DENSE", dense)
        eng = E.SearchEngine.from_index(idx, max_batch=1)
        eng.reset_stats()
        res, rers = [], []
        for q in qs:
            res.append(eng.search_raw([q], 100, E.PRF_STORED_ROWS))
            rers.append(eng.debug_read("rer", 0))
        out[dense] = (res, rers, eng.stats()["column_scan_launches"])
        eng.close()
    assert out["1"][2] == 0 and out["0"][2] == len(qs)
    for (a, ra), (b, rb) in zip(zip(out["1"][0], out["1"][1]), zip(out["0"][0], out["0"][1])):
        assert np.array_equal(ra, rb)                                  # fp32 SIMT dense scan == column scan, bit for bit
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(x, y)


def test_column_cache_follows_row_updates():
    """The engine caches column 0 of the row store for the collapsed PRF re-query; reloading rows must refresh it."""
    idx = synth.generate_index(8000, vocab_size=600, seed=3)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    q = Q.make_query(synth.generate_queries(idx, 1, seed=8)[0], t2i, infer)
    eng = E.SearchEngine.from_index(idx, max_batch=1)
    eng.search_raw([q], 50, E.PRF_STORED_ROWS)
    r1 = eng.debug_read("rer", 0)
    rows2 = idx.rows.copy()
    rows2[:, 0] = -3.0 * rows2[:, 0] + 0.25
    eng.load_vectors(rows2)
    eng.search_raw([q], 50, E.PRF_STORED_ROWS)
    r2 = eng.debug_read("rer", 0)
    assert not np.array_equal(r1, r2)
    # r2 = RN(rows2[:, 0] * c) for the new re-query scalar c: recover c from one doc and check all the others
    d0 = int(np.argmax(np.abs(rows2[:, 0])))
    c = np.float32(r2[d0] / rows2[d0, 0])
    assert np.allclose(r2, rows2[:, 0] * c, rtol=2e-6, atol=0)
    eng.close()


def test_constants_are_honoured_at_call_time():
    idx = synth.generate_index(6000, vocab_size=500, seed=5)
    P = port.OraclePort(idx)
    install(idx)
    q = synth.generate_queries(idx, 3, seed=9)
    try:
        webui_api.BM25_WEIGHT, webui_api.DOC2VEC_WEIGHT = 0.3, 0.7
        webui_api.ORIGINAL_SCORE_WEIGHT, webui_api.RERANKED_SCORE_WEIGHT = 0.6, 0.4
        webui_api.DIFF_FILTER_THRESH = 1e-4              # near-ties now common: exercises the exact filter fallback
        P.consts.update(BM25_WEIGHT=0.3, DOC2VEC_WEIGHT=0.7, ORIGINAL_SCORE_WEIGHT=0.6, RERANKED_SCORE_WEIGHT=0.4,
                        DIFF_FILTER_THRESH=1e-4)
        for text in q:
            for topn in (20, 100, 800):
                assert_same_or_filter_unstable(capture(webui_api.find_similar_documents, text, topn),
                                               capture(P.find_similar_documents, text, topn),
                                               lambda: P.find_sorted(text), 1e-4, topn, text)
    finally:
        webui_api.BM25_WEIGHT = webui_api.DOC2VEC_WEIGHT = 0.5
        webui_api.ORIGINAL_SCORE_WEIGHT, webui_api.RERANKED_SCORE_WEIGHT = 0.7, 0.3
        webui_api.DIFF_FILTER_THRESH = 1e-6


def test_filter_fallback_is_exact_and_counted():
    """with exactly one near-tie inside the first k entries the outcome depends on whether
    ANY other near-tie exists in the whole list - the engine must sort everything to decide."""
    n = 5000
    idx = synth.generate_index(n, vocab_size=300, seed=21)
    rng = np.random.default_rng(0)
    base = 1.0 - np.arange(n) * 1e-4                 # every gap 1e-4 ...
    base[17:] += 1e-4 - 3e-7                         # ... except ONE near-tie between ranks 16 and 17
    final = np.empty(n)
    perm = rng.permutation(n)
    final[perm] = base
    for tweak in ("one", "two_far", "none"):
        f = final.copy()
        if tweak == "two_far":
            f[perm[4000:]] += 1e-4 - 2e-7            # a second near-tie far beyond any top-k prefix
        if tweak == "none":
            f[perm[17:]] -= 1e-4 - 3e-7              # no near-tie at all
        # PRF off: sorted(final) -> filter -> [:topn]  (the webui.py:247-253 branch semantics)
        eng = install(idx, prf_mode="off")
        eng.reset_stats()
        order = np.argsort(-f, kind="stable")
        want = port.filter_searched_result(list(zip(order.tolist(), f[order])))[:20]
        got = webui_api.get_doc2vec_based_reranked_scores(f, 20)
        assert [d for d, _ in got] == [d for d, _ in want], tweak
        assert [s for _, s in got] == [s for _, s in want], tweak
        assert len(got) == {"one": 16, "two_far": 20, "none": 20}[tweak]
        # "two_far" is settled by the near-tie witness pass, only "one" needs the full sort
        assert eng.stats()["fullsort_fallbacks"] == (1 if tweak == "one" else 0), tweak
        # PRF on, with the re-query weight at 0 so that R = 0.7 * final keeps the crafted gaps
        eng = install(idx, prf_mode="stored_rows", reranked_score_weight=0.0)
        webui_api.RERANKED_SCORE_WEIGHT = 0.0
        try:
            P = port.OraclePort(idx)
            P.consts["RERANKED_SCORE_WEIGHT"] = 0.0
            eng.reset_stats()
            assert_same(capture(webui_api.get_doc2vec_based_reranked_scores, f, 40), capture(P.rerank, f, 40), tweak)
            assert eng.stats()["fullsort_fallbacks"] == (1 if tweak == "one" else 0), tweak
        finally:
            webui_api.RERANKED_SCORE_WEIGHT = 0.3


@pytest.mark.parametrize("prf_mode", ["stored_rows", "off"])
def test_topn_beyond_the_selector_is_not_truncated(prf_mode):
    """find_similar_documents(new_doc, topn) accepts any topn (webui.py:345); the select stages return at most
    ais_max_select_k() = 1024 candidates, so a longer result goes through the exact full sort of every doc's key -
    single engine and doc-sharded - instead of being cut silently."""
    from ais_b200 import shard
    idx = synth.generate_index(30000, vocab_size=1500, seed=61)
    P = port.OraclePort(idx)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    texts = synth.generate_queries(idx, 5, seed=4)
    TH = 1e-13                       # (almost) no near-tie cut: the results really are longer than the selector's lists
    P.consts["DIFF_FILTER_THRESH"] = TH
    eng = install(idx, prf_mode=prf_mode, diff_filter_thresh=TH)
    webui_api.DIFF_FILTER_THRESH = TH
    engines = [E.SearchEngine.from_index(idx, lo=lo, hi=hi, diff_filter_thresh=TH)
               for lo, hi in (shard.shard_bounds(idx.n_docs, 2, r) for r in range(2))]
    S = shard.ShardedSearch(engines, idx.n_docs)
    mode = E.PRF_STORED_ROWS if prf_mode == "stored_rows" else E.PRF_OFF
    n_long = 0
    for text in texts:
        for topn in (1100, 2500):
            if prf_mode == "off":
                st = P.stages(text)
                order = np.argsort(-st["final"], kind="stable")
                srt = list(zip(order.tolist(), st["final"][order]))
                want = capture(lambda: port.filter_searched_result(srt, TH)[:topn])
                got = capture(lambda: eng.search([Q.make_query(text, t2i, infer)], topn, mode)[0])
                sorted_fn = lambda s=srt: s
            else:
                want = capture(P.find_similar_documents, text, topn)
                got = capture(webui_api.find_similar_documents, text, topn)
                sorted_fn = lambda t=text: P.find_sorted(t)
            assert_same_or_filter_unstable(got, want, sorted_fn, TH, topn, (text, topn))
            got2 = capture(lambda: S.search([Q.make_query(text, t2i, infer)], topn, mode)[0])
            assert_same_or_filter_unstable(got2, want, sorted_fn, TH, topn, ("sharded", text, topn))
            n_long += want[0] == "ok" and len(want[1]) > 1034
    webui_api.DIFF_FILTER_THRESH = 1e-6
    assert n_long > 0, "no query returned more than the selector's 1024 + 10 docs: the test did not reach the long path"

    for e in engines:
        e.close()


def test_bitmap_path_equals_record_path(monkeypatch):
    """The BM25 side has two exact implementations: per-tile records from a posting walk (bm25_score_kernel; the default)
    and term bitmaps (tf == 1 indexes, AIS_BM25_BITMAP=1; bm25_bitmap_kernel + bm25_tile_values: every pass re-derives the
    values, IEEE division instead of the FMA-corrected one).  On a tf == 1 index both must return the same bits:
    compute_bm25_scores (-inf masks included), combined scores, search results."""
    idx = synth.generate_index(45000, vocab_size=1200, seed=23)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    texts = synth.generate_queries(idx, 40, seed=12) + ["t3:+2 t12:+1 t9", "t0:-1 t1:-2 t2", "t13:0 t14:+0 t4:-0 t20:2"]
    qs = [Q.make_query(t, t2i, infer) for t in texts]
    many_terms = E.Query(qs[0].vec, np.arange(40, dtype=np.int32), np.r_[np.arange(1, 39, dtype=np.float64), 1003.0, -2.0])
    qs.append(many_terms)                             # 40 terms: more than one group of four in flight
    P = port.OraclePort(idx)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("AIS_BM25_BITMAP", flag)
        eng = E.SearchEngine.from_index(idx, max_batch=len(qs))
        eng.reset_stats()
        res = eng.search_raw(qs, 100, E.PRF_STORED_ROWS)
        fins = [eng.debug_read("fin", j) for j in (0, 7, len(qs) - 1)]
        bm = [eng.bm25_scores(q.term_ids, q.weights) for q in qs[-6:]]
        assert (eng.stats()["bitmap_batches"] > 0) == (flag == "1")
        out[flag] = (res, fins, bm)
        eng.close()
    for x, y in zip(out["0"][0][:4], out["1"][0][:4]):
        assert np.array_equal(x, y)
    for x, y in zip(out["0"][1] + out["0"][2], out["1"][1] + out["1"][2]):
        assert np.array_equal(x, y, equal_nan=True)
    for q, got in zip(qs[-6:], out["0"][2]):         # and both equal the oracle's compute_bm25_scores, bit for bit
        want = P.bm25_scores(dict(zip(q.term_ids.tolist(), q.weights.tolist())))
        assert np.array_equal(got, want)


def test_fma_corrected_division_equals_ieee_division(monkeypatch):
    """bm25 / max(bm25) (webui.py:379-380, fp64) runs as RN(1/max) + two FMAs (finals.cuh ddiv_by_max); AIS_IEEE_DIV=1
    switches to __ddiv_rn.  Every combined score of every doc and the search results must be the same bits."""
    idx = synth.generate_index(60000, vocab_size=900, seed=29)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(t, t2i, infer) for t in synth.generate_queries(idx, 32, seed=21)]
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("AIS_IEEE_DIV", flag)
        eng = E.SearchEngine.from_index(idx, max_batch=len(qs))
        res = eng.search_raw(qs, 100, E.PRF_STORED_ROWS)
        out[flag] = (res, [eng.debug_read("fin", j) for j in range(len(qs))])
        eng.close()
    for x, y in zip(out["0"][0][:4], out["1"][0][:4]):
        assert np.array_equal(x, y)
    for x, y in zip(out["0"][1], out["1"][1]):
        assert np.array_equal(x, y, equal_nan=True)


def test_pass2_tile_skip_equals_full_pass(monkeypatch):
    """Pass 2 of the reference's collapsed re-query: the maxima kernel skips tiles whose upper bound on the blend R stays
    below a threshold taken from the pass-1 candidates (select2.cuh, rerank_max_kernel<1>; default), visits every tile
    (AIS_NO_TILE_SKIP=1), or the collect itself applies the bound (AIS_TILE_BOUND=1, collect_kernel<2, 1>); and the generic
    streaming kernel (AIS_NO_RERANK_MAX=1, segmax_kernel<2>).  Same results bit for bit, max R included."""
    idx = synth.generate_index(300000, vocab_size=2000, seed=88, rows="random")
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    qs = [Q.make_query(q, t2i, infer) for q in synth.generate_queries(idx, 24, seed=6)]
    out = {}
    for name, env in (("skip", {}), ("full", {"AIS_NO_TILE_SKIP": "1"}), ("collect_bound", {"AIS_TILE_BOUND": "1"}),
                      ("generic", {"AIS_NO_RERANK_MAX": "1"})):
        for k in ("AIS_NO_TILE_SKIP", "AIS_TILE_BOUND", "AIS_NO_RERANK_MAX"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = E.SearchEngine.from_index(idx, max_batch=24)
        out[name] = [eng.search_raw(qs, topn, E.PRF_STORED_ROWS) for topn in (100, 800)]
        assert (eng.stats()["bound_passes"] > 0) == (name in ("skip", "collect_bound")), name
        eng.close()
    for name in ("full", "collect_bound", "generic"):
        for a, b in zip(out["skip"], out[name]):
            for x, y in zip(a[:4], b[:4]):
                assert np.array_equal(x, y), name


def test_clustered_top_docs_overflow_the_streaming_select():
    """All the best docs sit in one contiguous id range: the segment-maximum threshold lets thousands of
    survivors through, the streaming select overflows and the gated buffer select must take over."""
    n = 100000
    idx = synth.generate_index(n, vocab_size=300, seed=33, with_rows=True)
    rng = np.random.default_rng(1)
    f = rng.random(n) * 0.5
    f[20000:26000] = 0.6 + np.linspace(0.3, 0.0, 6000)          # 6000 clustered top docs, descending with the id
    f[77] = -np.inf
    P = port.OraclePort(idx)
    for mode in ("off", "stored_rows"):
        install(idx, prf_mode=mode)
        for topn in (100, 800):
            if mode == "off":
                order = np.argsort(-f, kind="stable")
                want = port.filter_searched_result(list(zip(order.tolist(), f[order])))[:topn]
                want = ("ok", [d for d, _ in want], [s for _, s in want])
            else:
                want = capture(P.rerank, f, topn)
            assert_same(capture(webui_api.get_doc2vec_based_reranked_scores, f, topn), want, (mode, topn))


def test_load_rejects_corrupt_posting_lists():
    """ais_load_bm25 validates what the kernels index with - doc ids inside the shard and strictly ascending per
    term (one device pass at load time); a corrupt index is an error, not an out-of-bounds write."""
    from ais_b200 import binding as B
    eng = E.SearchEngine(device=0, max_batch=1)
    n_docs, n_terms = 1000, 3
    idf = np.ones(n_terms)
    doc_len = np.full(n_docs, 2, dtype=np.int64)
    ptr = np.array([0, 3, 5, 6], dtype=np.int64)
    good = np.array([1, 5, 999, 0, 7, 42], dtype=np.int32)
    eng.load_bm25(ptr, good, None, idf, doc_len, 2.0)
    for bad, term in ((np.array([1, 5, 1000, 0, 7, 42], np.int32), 0),        # beyond the shard
                      (np.array([1, 5, 999, 7, 7, 42], np.int32), 1),         # not strictly ascending
                      (np.array([1, 5, 999, 0, 7, -1], np.int32), 2)):        # negative
        with pytest.raises(B.AisError) as ei:
            eng.load_bm25(ptr, bad, None, idf, doc_len, 2.0)
        assert "term %d" % term in str(ei.value)
    eng.close()
