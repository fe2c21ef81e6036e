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


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's CPU algorithm on a bounded sample) needs no GPU: one JSON line with
    the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample-docs", "3000", "--quick-reference"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    # kind: the reference's verbatim functions where /root/reference exists (this container), the oracle port elsewhere
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["batch"] == 256 and line["config"]["docs"] == 10_000_000          # the b200 arm's config
    assert line["ms_per_step"] > 0 and line["measured"][0]["docs"] == 3000                  # what was timed
    assert line["extrapolated"]["to_docs"] == 10_000_000
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_names_the_scan_kernel_of_a_batch():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.scan_kernel_for(1)[1] == "scan_simt" and bench.scan_kernel_for(8)[1] == "scan_mma8"
    assert bench.scan_kernel_for(9)[1] == "scan_tc32" and bench.scan_kernel_for(32)[1] == "scan_tc32"
    os.environ.pop("AIS_SCAN_PAIR", None)
    assert bench.scan_kernel_for(33)[1] == "scan_pair64" and bench.scan_kernel_for(256)[1] == "scan_pair64"     # CTA pairs (default)
    os.environ["AIS_SCAN_PAIR"] = "0"
    try:
        assert bench.scan_kernel_for(64)[1] == "scan_tc64"
    finally:
        del os.environ["AIS_SCAN_PAIR"]
    t2, src2 = bench.load_traffic("scan_pair64", 10_000_000)
    assert src2 and 14.4e9 < t2 < 14.7e9                      # 12.0 GB of rows read once + 2.55 GB of scores written
    traffic, src = bench.load_traffic("scan_tc64", 5_000_000)
    assert src and abs(traffic - 14556012000 / 2) < 1e6       # scaled to the shard's rows


def test_packed_query_records_match_the_c_struct():
    """engine._pack_queries builds the ais_query array in one numpy buffer: same bytes as filling the ctypes structures."""
    import ctypes as C
    import numpy as np
    import ais_b200  # noqa: F401
    from ais_b200 import binding as B, engine as E
    rng = np.random.default_rng(0)
    qs = [E.Query(rng.standard_normal(300).astype(np.float32), np.arange(1 + i % 5, dtype=np.int32) + i,
                  rng.standard_normal(1 + i % 5)) for i in range(37)]
    packed = E._pack_queries(qs)
    rec = packed._as_parameter_
    assert C.sizeof(B.AisQuery) == 32 and packed.buf.shape == (37, 4)
    for i, q in enumerate(qs):
        assert rec[i].n_terms == len(q.term_ids)
        assert C.addressof(rec[i].vec.contents) == q.vec.ctypes.data
        assert C.addressof(rec[i].term_ids.contents) == q.term_ids.ctypes.data
        assert C.addressof(rec[i].weights.contents) == q.weights.ctypes.data
        assert rec[i].vec[299] == q.vec[299] and rec[i].weights[0] == q.weights[0]


def test_merged_prefix_of_short_shard_lists_is_exact():
    """The argument behind shard.py's short per-shard lists (select.cuh prefix_bound_kernel, mirrored in tests/fake_engine.py):
    every doc a shard did not return scores at most that shard's last returned key, so the merged list is the true global
    order strictly above the largest such key (and for its first k entries in any case).  Random scores with ties, random
    splits: the prefix the rule certifies equals the true global ranking, and it is never shorter than k."""
    import numpy as np
    rng = np.random.default_rng(5)
    for case in range(300):
        n = int(rng.integers(20, 400))
        shards = int(rng.integers(2, 9))
        k = int(rng.integers(1, 40))
        scores = rng.integers(0, 60, size=n).astype(np.float64)         # many ties
        ids = np.arange(n)
        cuts = np.sort(rng.choice(np.arange(1, n), size=shards - 1, replace=False))
        order_all = np.lexsort((ids, -scores))                            # score desc, id asc: the engine's order
        lists, bound = [], -np.inf
        for lo, hi in zip(np.r_[0, cuts], np.r_[cuts, n]):
            o = lo + np.lexsort((ids[lo:hi], -scores[lo:hi]))[:k]
            lists.append(o)
            if len(o) == k and hi - lo >= k:                              # a full list: the shard may hold more at <= its last key
                bound = max(bound, scores[o[-1]])
        cand = np.concatenate(lists)
        merged = cand[np.lexsort((ids[cand], -scores[cand]))][:1024]
        exact = max(min(k, len(merged)), int((scores[merged] > bound).sum()))
        assert exact >= min(k, len(merged))
        assert np.array_equal(merged[:exact], order_all[:exact]), (case, n, shards, k)


def test_shard_list_depth_policy(monkeypatch):
    """ShardedSearch asks every shard for kmax / shards + 128 candidates (at least what the result needs, at most half the
    shard's segment count); AIS_SHARD_CUT=0 restores the single-engine depth."""
    from types import SimpleNamespace
    import ais_b200  # noqa: F401
    from ais_b200 import shard
    eng = SimpleNamespace(params=SimpleNamespace(prf_depth=10, max_batch=4), max_select_k=lambda: 1024)
    monkeypatch.delenv("AIS_SELECT_DEPTH", raising=False)
    monkeypatch.delenv("AIS_SHARD_CUT", raising=False)
    S = shard.ShardedSearch([eng], 10_000_000)
    S.n_shards = 8
    assert S._select_depth(91, 1014) == 256                    # 1024 / 8 + 128
    assert S._select_depth(500, 1014) == 500                   # never less than the result needs
    S.n_shards = 1
    assert S._select_depth(91, 1014) == 576                    # one engine: the engine's own default cap (engine.cu sel_deep)
    S.n_shards = 8
    monkeypatch.setenv("AIS_SHARD_CUT", "0")
    assert S._select_depth(91, 1014) == 814                    # 1.25 M docs per shard: 1628 segments / 2
