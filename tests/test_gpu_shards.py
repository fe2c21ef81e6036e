"""-m gpu: doc-sharded search.  Several engines, each holding a contiguous doc range, are driven by
shard.ShardedSearch in one process (records reduced locally instead of over NCCL): the merged result
must equal the single-engine result bit for bit, and the oracle within tolerance."""
import numpy as np
import pytest

from gpu_util import assert_same_or_filter_unstable, capture
import ais_b200  # noqa: F401
from ais_b200 import engine as E, query as Q, shard, synth
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_shards", [2, 3, 8])
@pytest.mark.parametrize("mode", [E.PRF_STORED_ROWS, E.PRF_CALLBACK, E.PRF_OFF])
def test_sharded_equals_single(n_shards, mode):
    idx = synth.generate_index(30000, vocab_size=1500, seed=404)
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    queries = synth.generate_queries(idx, 16, seed=8)
    qs = [Q.make_query(q, t2i, infer) for q in queries]

    def cb(qi, ids, scores):
        vecs = [[(i, v) for i, v in enumerate(idx.infer.one(idx.doc_tags(int(d))))] for d in ids]
        return Q.dense_query(Q.prf_query(vecs, scores.tolist()))

    single = E.SearchEngine.from_index(idx, max_batch=4)
    ref = single.search_raw(qs, 100, mode, cb if mode == E.PRF_CALLBACK else None)
    engines = []
    for r in range(n_shards):
        lo, hi = shard.shard_bounds(idx.n_docs, n_shards, r)
        engines.append(E.SearchEngine.from_index(idx, max_batch=4, lo=lo, hi=hi))
    S = shard.ShardedSearch(engines, idx.n_docs)
    for lo in range(0, len(qs), 4):
        got = S.search_raw(qs[lo:lo + 4], 100, mode, cb if mode == E.PRF_CALLBACK else None)
        for j in range(len(qs[lo:lo + 4])):
            q = lo + j
            assert got[3][j] == ref[3][q] and got[2][j] == ref[2][q], (q, got[2][j], ref[2][q])
            c = ref[2][q]
            assert np.array_equal(got[0][j, :c], ref[0][q, :c]), q
            assert np.array_equal(got[1][j, :c], ref[1][q, :c]), q
    if mode != E.PRF_OFF:
        P = port.OraclePort(idx)
        S2 = shard.ShardedSearch(engines, idx.n_docs)
        for text, q in list(zip(queries, qs))[:6]:
            got = capture(lambda t, n: S2.search([q], n, mode, cb if mode == E.PRF_CALLBACK else None)[0], text, 100)
            assert_same_or_filter_unstable(got, capture(P.find_similar_documents, text, 100), lambda: P.find_sorted(text),
                                           1e-6, 100, text)
    for e in engines:
        e.close()
    single.close()


def test_sharded_exact_filter_fallback():
    idx = synth.generate_index(3000, vocab_size=300, seed=21)
    P = port.OraclePort(idx)
    P.consts["DIFF_FILTER_THRESH"] = 3e-5
    t2i = idx.token2id
    infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
    engines = []
    for r in range(3):
        lo, hi = shard.shard_bounds(idx.n_docs, 3, r)
        engines.append(E.SearchEngine.from_index(idx, lo=lo, hi=hi, diff_filter_thresh=3e-5))
    S = shard.ShardedSearch(engines, idx.n_docs)
    for text in synth.generate_queries(idx, 30, seed=2):
        q = Q.make_query(text, t2i, infer)
        got = capture(lambda t, n: S.search([q], n, E.PRF_STORED_ROWS)[0], text, 12)
        assert_same_or_filter_unstable(got, capture(P.find_similar_documents, text, 12), lambda: P.find_sorted(text),
                                       3e-5, 12, text)
