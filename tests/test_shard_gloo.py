"""world_size-2 (and 3) gloo runs of shard.ShardedSearch on CPU with numpy stage engines (tests/fake_engine.py):
the collective sequence, gather layouts and the ambiguity protocol of the N>1 path, checked against the
oracle's find_similar_documents on the whole index (same arithmetic -> bit-exact)."""
import os
import socket
import sys
import traceback
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, engines_per_rank, thresh, out_q):
    try:
        for p in (ROOT, HERE):
            if p not in sys.path:
                sys.path.insert(0, p)
        warnings.filterwarnings("ignore", category=RuntimeWarning)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port_no)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import ais_b200  # noqa: F401
        from ais_b200 import engine as E, query as Q, shard, synth
        from fake_engine import FakeStageEngine
        from gpu_util import filter_alternatives, same_ranking
        from oracle import port
        idx = synth.generate_index(3000, vocab_size=300, seed=21)
        P = port.OraclePort(idx)
        P.consts["DIFF_FILTER_THRESH"] = thresh
        n_sh = world * engines_per_rank
        engines = []
        for j in range(engines_per_rank):
            lo, hi = shard.shard_bounds(idx.n_docs, n_sh, rank * engines_per_rank + j)
            engines.append(FakeStageEngine(idx, lo, hi, max_batch=4, thresh=thresh))
        S = shard.ShardedSearch(engines, idx.n_docs)
        t2i = idx.token2id
        infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])

        def cb(qi, ids, scores):
            vecs = [[(i, v) for i, v in enumerate(idx.infer.one(idx.doc_tags(int(d))))] for d in ids]
            return Q.dense_query(Q.prf_query(vecs, scores.tolist()))

        texts = synth.generate_queries(idx, 12, seed=2)
        n_checked = 0
        for mode in (E.PRF_STORED_ROWS, E.PRF_CALLBACK):
            for topn in (12, 100):
                for text in texts:
                    try:
                        want = ("ok", P.find_similar_documents(text, topn))
                    except Exception as e:   # noqa: BLE001
                        want = ("err", type(e).__name__)
                    try:
                        got = ("ok", S.search([Q.make_query(text, t2i, infer)], topn, mode, cb if mode == E.PRF_CALLBACK else None)[0])
                    except Exception as e:   # noqa: BLE001
                        got = ("err", type(e).__name__, traceback.format_exc()[-600:])
                    if want[0] == "err":
                        assert got[:2] == want or (rank != 0 and got[0] == "err"), (text, got, want)
                    else:
                        assert got[0] == "ok", (text, got)
                        # per-shard sgemv may block its fp32 sums differently from the whole-matrix sgemv
                        msg = same_ranking([d for d, _ in got[1]], [s for _, s in got[1]],
                                           [d for d, _ in want[1]], [s for _, s in want[1]])
                        if msg is not None and len(got[1]) != len(want[1]):
                            alts = filter_alternatives(P.find_sorted(text), thresh, topn)
                            msg = None if any(same_ranking([d for d, _ in got[1]], [s for _, s in got[1]], [d for d, _ in a],
                                                           [s for _, s in a]) is None for a in alts) else msg
                        assert msg is None, (text, mode, topn, msg)
                    n_checked += 1
        # PRF off == sorted(final) -> filter -> [:topn]
        for text in texts[:4]:
            st = P.stages(text)
            order = np.argsort(-st["final"], kind="stable")
            want = port.filter_searched_result(list(zip(order.tolist(), st["final"][order])), thresh)[:50]
            got = S.search([Q.make_query(text, t2i, infer)], 50, E.PRF_OFF)[0]
            assert same_ranking([d for d, _ in got], [s for _, s in got], [d for d, _ in want], [s for _, s in want]) is None, text
        out_q.put((rank, "ok", n_checked, S.fullsort_fallbacks))
    except Exception:   # noqa: BLE001
        out_q.put((rank, "fail", traceback.format_exc(), 0))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,per_rank,thresh", [(2, 1, 1e-6), (2, 2, 3e-5), (3, 1, 3e-5)])
def test_sharded_search_over_gloo(world, per_rank, thresh):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, per_rank, thresh, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info, _ in sorted(results):
        assert status == "ok", "rank %d failed:\n%s" % (rank, info)
    assert all(r[2] > 0 for r in results)
