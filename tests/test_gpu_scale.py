"""-m gpu: parity at the BENCHMARKED configuration - the sizes where the engine leaves its small-index branches.

Every parity test in test_gpu_parity.py holds <= 100 000 docs = <= 391 tiles, so a select segment is always ONE tile.
From 1 M docs (BASELINE configs[1]: 3 907 tiles of 256 docs) a segment spans several tiles and the segment table of the
combine kernel wraps; max_batch = 256 (configs[2]) runs four 64-query tensor-core passes.  Here the corpus comes from the
same on-device generator bench.py times (synth_torch), and every result is compared with the oracle port
(webui.py:345-390 restated, oracle/port.py) - ids in order, scores within 1e-5, statuses.
"""
import numpy as np
import pytest

from gpu_util import assert_same_or_filter_unstable
import ais_b200  # noqa: F401
from ais_b200 import engine as E

pytestmark = pytest.mark.gpu

N_DOCS = 1_000_000
TOPN = 100


@pytest.fixture(scope="module")
def corpus():
    import scale_util as SU
    engines, view = SU.build_corpus(N_DOCS, max_batch=256)
    P = SU.StoredRowOracle(view)
    yield engines[0], view, P
    engines[0].close()


def _check(eng, P, texts, qs, res, topn=TOPN, what=""):
    import scale_util as SU
    want = SU.oracle_results(P, texts, topn)
    ids, scores, counts, status = res[:4]
    n_ok = 0
    for j, text in enumerate(texts):
        got = SU.engine_outcome(ids, scores, counts, status, j)
        assert_same_or_filter_unstable(got, want[j], lambda t=text: P.find_sorted_arrays(t), 1e-6, topn, (what, j, text))
        n_ok += want[j][0] == "ok"
    return n_ok


def test_1m_docs_single_queries(corpus):
    """configs[1]: 1 M docs, V = 10 861, single weighted queries with +required / -exclude, top-100 (fp32 SIMT scan)."""
    import scale_util as SU
    eng, view, P = corpus
    texts, qs = SU.make_queries(view, 12, seed=101)
    eng.reset_stats()
    outs = [eng.search_raw([q], TOPN, E.PRF_STORED_ROWS) for q in qs]
    res = [np.concatenate([o[i] for o in outs]) for i in range(4)]
    assert _check(eng, P, texts, qs, res, what="1M single") >= 8
    st = eng.stats()
    assert st["tiles_per_seg"] > 1, "a 1 M-doc shard must take the multi-tile segment branch"
    assert st["scan_launches"] >= len(qs)


def test_1m_docs_batch64(corpus):
    """One 64-query tcgen05 pass on the 1 M-doc index."""
    import scale_util as SU
    eng, view, P = corpus
    texts, qs = SU.make_queries(view, 64, seed=202)
    res = eng.search_raw(qs, TOPN, E.PRF_STORED_ROWS)
    assert _check(eng, P, texts, qs, res, what="1M batch 64") >= 50
    assert eng.stats()["tiles_per_seg"] > 1


def test_1m_docs_batch256_four_passes(corpus):
    """configs[2]'s batch shape: max_batch = 256, one full batch (four 64-query passes), every result vs the oracle."""
    import scale_util as SU
    eng, view, P = corpus
    texts, qs = SU.make_queries(view, 256, seed=303)
    eng.reset_stats()
    res = eng.search_raw(qs, TOPN, E.PRF_STORED_ROWS)
    assert _check(eng, P, texts, qs, res, what="1M batch 256") >= 200
    st = eng.stats()
    assert st["column_scan_launches"] >= 1 and st["tiles_per_seg"] > 1


def test_1m_docs_topn_800_and_prf_modes(corpus):
    """webui.py's own topn (800, webui.py:586) and the other PRF modes at 1 M docs (dense second pass / no PRF)."""
    import scale_util as SU
    from gpu_util import capture
    eng, view, P = corpus
    texts, qs = SU.make_queries(view, 24, seed=404)
    res = eng.search_raw(qs, 800, E.PRF_STORED_ROWS)
    assert _check(eng, P, texts, qs, res, topn=800, what="1M topn 800") >= 16
    # PRF off == webui.py:247-253 applied to the combined scores
    res = eng.search_raw(qs[:8], TOPN, E.PRF_OFF)
    for j, text in enumerate(texts[:8]):
        st = P.stages(text)
        order = np.argsort(-st["final"], kind="stable")
        srt_scores = st["final"][order]
        want = capture(lambda: P.filter_arrays(order, srt_scores, 1e-6, TOPN))
        got = SU.engine_outcome(res[0], res[1], res[2], res[3], j)
        assert_same_or_filter_unstable(got, want, lambda o=order, s=srt_scores: (o, s), 1e-6, TOPN, ("1M prf off", j, text))


def test_sharded_1m_docs_three_engines_vs_oracle():
    """The doc-sharded driver (shard.py) at 1 M docs, three shards on one GPU, a 64-query batch vs the oracle."""
    import scale_util as SU
    from ais_b200 import shard
    engines, view = SU.build_corpus(N_DOCS, max_batch=64, n_shards=3)
    P = SU.StoredRowOracle(view)
    texts, qs = SU.make_queries(view, 64, seed=505)
    S = shard.ShardedSearch(engines, N_DOCS)
    res = S.search_raw(qs, TOPN, E.PRF_STORED_ROWS)
    assert _check(engines[0], P, texts, qs, res, what="1M 3 shards") >= 50
    for e in engines:
        e.close()


def test_nccl_two_ranks_vs_oracle(tmp_path):
    """The NCCL data path itself: two processes, one GPU each (torch.distributed.run), 1 M docs sharded by document,
    a 64-query batch through shard.ShardedSearch; rank 0 compares every result with the oracle (tests/nccl_worker.py).
    Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "nccl_parity.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(here, "nccl_worker.py"), "--docs", str(N_DOCS), "--batch", "64", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_PARITY" in r.stdout
