"""Shared helpers of the -m gpu parity tests: the CUDA path (through the C ABI) next to the oracle."""
import warnings

import numpy as np

import ais_b200  # noqa: F401
from ais_b200 import engine as E, query as Q, webui_api

warnings.filterwarnings("ignore", category=RuntimeWarning)

# north_star: returned scores within 1e-5 relative.  The only arithmetic that differs from the reference is
# the summation order of the fp32 dot products (BLAS sgemv vs. the scan kernel); its error is ~1 fp32 ulp
# of the max-normalised scale (scores are normalised to max 1.0), so scores that cancel toward 0
# (0.7*final + 0.3*rer with rer < 0) get an absolute floor of 2 fp32 ulps of 1.0.
SCORE_RTOL = 1e-5
SCORE_ATOL = 2.4e-7


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
    if len(gs) and not np.allclose(gs, ws, rtol=rtol, atol=SCORE_ATOL):
        bad = int(np.argmax(np.abs(gs - ws) - rtol * np.abs(ws)))
        return "score[%d] %r != %r" % (bad, gs[bad], ws[bad])
    if list(got_ids) == list(want_ids):
        return None
    # permutations are only tolerated inside runs of near-tied reference scores
    i = 0
    n = len(want_ids)
    while i < n:
        j = i + 1
        while j < n and abs(ws[j] - ws[j - 1]) <= rtol * abs(ws[j - 1]) + SCORE_ATOL:
            j += 1
        if sorted(got_ids[i:j]) != sorted(want_ids[i:j]):
            return "ids differ in positions %d..%d: %r vs %r" % (i, j, got_ids[i:j], want_ids[i:j])
        i = j
    return None


FILTER_GAP_NOISE = 4e-7   # fp32 rounding of two adjacent max-normalised scores (each ~1.2e-7 at the 1.0 scale)


def filter_alternatives(sorted_list, thresh, topn):
    """filter_searched_result (webui.py:63-80) cuts where an adjacent gap is < 1e-6.  A gap that lies
    within fp32 rounding of that threshold may legitimately fall on either side when the dot products
    are summed in another order; return the outcomes for every threshold in thresh +- FILTER_GAP_NOISE.
    `sorted_list`: list of (id, score) or an (ids, scores) pair of arrays (the >= 1 M-doc tests).  Only the first two
    "found" gaps decide the outcome, so near-threshold gaps beyond the second CERTAIN near-tie (gap < thresh - noise)
    cannot matter and are not enumerated - at 1 M docs nearly every gap deep in the list is near 1e-6."""
    from oracle import port
    if isinstance(sorted_list, tuple):
        ids, s = np.asarray(sorted_list[0]), np.asarray(sorted_list[1], dtype=np.float64)
    else:
        ids = np.array([p[0] for p in sorted_list], dtype=np.int64)
        s = np.array([p[1] for p in sorted_list], dtype=np.float64)
    with np.errstate(invalid="ignore"):
        gaps = s[:-1] - s[1:]
    gaps = np.where(gaps == 0, np.inf, gaps)
    sure = np.nonzero(gaps < thresh - FILTER_GAP_NOISE)[0]
    limit = int(sure[1]) + 1 if len(sure) >= 2 else len(gaps)
    region = gaps[:limit]
    near = np.unique(region[np.isfinite(region) & (np.abs(region - thresh) <= FILTER_GAP_NOISE)])
    if len(near) > 48:                                    # keep the enumeration bounded: the gaps closest to the threshold
        near = near[np.argsort(np.abs(near - thresh))[:48]]
    cuts = sorted(set([thresh - FILTER_GAP_NOISE, thresh + FILTER_GAP_NOISE] + [float(g) for g in near] +
                      [float(np.nextafter(g, np.inf)) for g in near]))
    outs = []
    for t in cuts:
        res = port.OraclePort.filter_arrays(ids, s, t, topn)
        key = [d for d, _ in res]
        if all(key != [d for d, _ in o] for o in outs):
            outs.append(res)
    return outs


def assert_same_or_filter_unstable(got, want, sorted_list_fn, thresh, topn, what=""):
    """assert_same, except that a result whose LENGTH differs is accepted when it equals the reference's
    outcome for a threshold within fp32 noise of DIFF_FILTER_THRESH (see filter_alternatives)."""
    if want[0] == "err" or got[0] == "err" or len(got[1]) == len(want[1]):
        return assert_same(got, want, what)
    for alt in filter_alternatives(sorted_list_fn(), thresh, topn):
        if len(alt) == len(got[1]) and same_ranking(got[1], got[2], [d for d, _ in alt], [s for _, s in alt]) is None:
            return
    raise AssertionError((what, "length %d != %d and no threshold within noise explains it" % (len(got[1]), len(want[1]))))


def assert_same(got, want, what=""):
    if want[0] == "err":
        assert got[0] == "err" and got[1] == want[1] and got[2] == want[2], (what, got, want)
        return
    assert got[0] == "ok", (what, got)
    msg = same_ranking(got[1], got[2], want[1], want[2])
    assert msg is None, (what, msg)
