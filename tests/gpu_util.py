"""Shared helpers of the -m gpu parity tests: the CUDA path (through the C ABI) next to the oracle."""
import warnings

import numpy as np

import ais_b200  # noqa: F401
from ais_b200 import engine as E, query as Q, webui_api

warnings.filterwarnings("ignore", category=RuntimeWarning)

SCORE_RTOL = 1e-5        # north_star: returned scores within 1e-5 relative (fp32 dot summation order differs from BLAS)


class ModelStub:
    """webui.py's ``model`` as far as the path uses it (infer_vector, dv[0]) over the synthetic generator."""

    def __init__(self, idx):
        self._t2i = idx.token2id
        self._infer = idx.infer
        self.dv = [np.zeros(idx.rows.shape[1], dtype=np.float32)]

    def infer_vector(self, words):
        return self._infer.one([self._t2i[w] for w in words if w in self._t2i])


class DictStub:
    def __init__(self, idx):
        self.token2id = idx.token2id


def install(idx, max_batch=1, prf_mode="callback", **params):
    eng = E.SearchEngine.from_index(idx, device=0, max_batch=max_batch, **params)
    webui_api.install(eng, ModelStub(idx), DictStub(idx), idx.csv_lines())
    webui_api.PRF_MODE = prf_mode
    return eng


def capture(fn, *args):
    try:
        res = fn(*args)
    except Exception as e:   # noqa: BLE001 - exceptions are part of the reference's behaviour
        return ("err", type(e).__name__, str(e))
    return ("ok", [int(d) for d, _ in res], [float(s) for _, s in res])


def same_ranking(got_ids, got_scores, want_ids, want_scores, rtol=SCORE_RTOL):
    """ids identical in order except swaps inside groups of scores tied within tolerance; scores within rtol."""
    if len(got_ids) != len(want_ids):
        return "length %d != %d" % (len(got_ids), len(want_ids))
    gs, ws = np.asarray(got_scores), np.asarray(want_scores)
    if len(gs) and not np.allclose(gs, ws, rtol=rtol, atol=0):
        bad = int(np.argmax(np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-300)))
        return "score[%d] %r != %r" % (bad, gs[bad], ws[bad])
    if list(got_ids) == list(want_ids):
        return None
    # permutations are only tolerated inside runs of near-tied reference scores
    i = 0
    n = len(want_ids)
    while i < n:
        j = i + 1
        while j < n and abs(ws[j] - ws[j - 1]) <= rtol * max(abs(ws[j - 1]), 1e-300):
            j += 1
        if sorted(got_ids[i:j]) != sorted(want_ids[i:j]):
            return "ids differ in positions %d..%d: %r vs %r" % (i, j, got_ids[i:j], want_ids[i:j])
        i = j
    return None


def assert_same(got, want, what=""):
    if want[0] == "err":
        assert got[0] == "err" and got[1] == want[1] and got[2] == want[2], (what, got, want)
        return
    assert got[0] == "ok", (what, got)
    msg = same_ranking(got[1], got[2], want[1], want[2])
    assert msg is None, (what, msg)
