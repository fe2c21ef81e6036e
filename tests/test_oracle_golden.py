"""The oracle port (oracle/port.py) against what the REFERENCE itself returned (tests/golden)."""
import warnings

import numpy as np
import pytest

from golden_util import load_filter_cases, load_index, load_results, load_seams, run_and_capture
from oracle import port, verbatim

warnings.filterwarnings("ignore", category=RuntimeWarning)


@pytest.fixture(scope="module")
def main_index():
    return load_index("main")


@pytest.mark.parametrize("faithful", [False, True])
def test_port_matches_reference_results_bit_exact(main_index, faithful):
    P = port.OraclePort(main_index, faithful=faithful)
    for rec in load_results("main")["results"]:
        got = run_and_capture(P.find_similar_documents, rec["query"], rec["topn"])
        if "error" in rec:
            assert got[0] == "err" and got[1] == rec["error"], (rec["query"], got)
            assert got[2] == rec["message"], (rec["query"], got)
        else:
            assert got[0] == "ok", (rec["query"], got)
            assert got[1] == rec["ids"], rec["query"]
            assert got[2] == rec["scores"], rec["query"]     # bit-exact floats


@pytest.mark.parametrize("faithful", [False, True])
def test_port_tiny_index_no_prf_branch(faithful):
    P = port.OraclePort(load_index("tiny"), faithful=faithful)
    for rec in load_results("tiny")["results"]:
        got = run_and_capture(P.find_similar_documents, rec["query"], rec["topn"])
        assert got == ("ok", rec["ids"], rec["scores"]), rec["query"]


def test_port_seam_vectors_bit_exact(main_index):
    z = load_seams()
    P = port.OraclePort(main_index)
    for k in range(4):
        st = P.stages(str(z["q%d_text" % k]))
        assert np.array_equal(st["q"], z["q%d_dense" % k])
        assert np.array_equal(st["sims"], z["q%d_sims" % k])
        assert np.array_equal(st["bm25"], z["q%d_bm25" % k])           # -inf masks included
        assert np.array_equal(st["final"], z["q%d_final" % k])
        assert list(st["weights"].keys()) == z["q%d_terms" % k].tolist()
        assert list(st["weights"].values()) == z["q%d_weights" % k].tolist()


def test_filter_known_answers():
    for c in load_filter_cases():
        res = port.filter_searched_result(c["input"])
        assert [d for d, _ in res] == c["ids"]
        assert [float(s) for _, s in res] == c["scores"]


@pytest.mark.skipif(not verbatim.available(), reason="reference sources only exist in the build container")
def test_port_matches_live_reference_on_fresh_index():
    from ais_b200 import synth
    ix = synth.generate_index(2500, vocab_size=500, seed=99, tf_gt1_fraction=0.01)
    W = verbatim.ReferenceWorld(ix)
    P = port.OraclePort(ix)
    for q in synth.generate_queries(ix, 25, seed=3):
        a = run_and_capture(W.find_similar_documents, q, 100)
        b = run_and_capture(P.find_similar_documents, q, 100)
        assert a == b, q


def test_array_flavour_of_the_port_is_bit_identical(main_index):
    """find_fast (no N-long Python lists; what the >= 1 M-doc GPU tests and bench.py --verify compare with) returns exactly
    what the reference returned - golden results incl. the exception types/messages - and what the list flavour returns on
    fresh indexes (tf > 1, the N <= 10 branch, topn 5/100/800)."""
    from ais_b200 import synth
    P = port.OraclePort(main_index)
    for rec in load_results("main")["results"]:
        got = run_and_capture(P.find_fast, rec["query"], rec["topn"])
        if "error" in rec:
            assert got[0] == "err" and got[1] == rec["error"] and got[2] == rec["message"], (rec["query"], got)
        else:
            assert got == ("ok", rec["ids"], rec["scores"]), rec["query"]
    Pt = port.OraclePort(load_index("tiny"))
    for rec in load_results("tiny")["results"]:
        assert run_and_capture(Pt.find_fast, rec["query"], rec["topn"]) == ("ok", rec["ids"], rec["scores"]), rec["query"]
    ix = synth.generate_index(4000, vocab_size=300, seed=5, tf_gt1_fraction=0.02)
    P2 = port.OraclePort(ix)
    for thresh in (1e-6, 1e-4):                      # 1e-4: near-ties are common, the cut branches run
        P2.consts["DIFF_FILTER_THRESH"] = thresh
        for q in synth.generate_queries(ix, 30, seed=11):
            for topn in (5, 100, 800):
                assert run_and_capture(P2.find_fast, q, topn) == run_and_capture(P2.find_similar_documents, q, topn), (q, topn)
