"""TEST INFRASTRUCTURE ONLY (oracle).  Generates tests/golden/*.npz|json by RUNNING THE REFERENCE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Each fixture holds a small synthetic index (arrays stored explicitly so nothing depends on RNG
stream stability), a list of queries, and what the reference's own functions (oracle/verbatim.py:
AST-extracted webui.py / genmodel.py, executed unchanged) returned for them: the final
``find_similar_documents`` lists, the ``compute_bm25_scores`` / ``index[vec]`` / combined-score
vectors, or the exception type and message.  gensim is stubbed per oracle/gensim_stub.py.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ais_b200  # noqa: E402
from ais_b200 import synth  # noqa: E402
from oracle import verbatim  # noqa: E402
from oracle.gensim_stub import dense_query  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def index_arrays(ix):
    return dict(
        n_docs=np.int64(ix.n_docs), vocab_size=np.int64(ix.vocab_size), seed=np.int64(ix.seed),
        row_ptr=ix.row_ptr, term_ids=ix.term_ids, tfs=ix.tfs, doc_len=ix.doc_len,
        avgdl=np.float64(ix.avgdl), idf=ix.idf, df=ix.df, rows=ix.rows, E=ix.infer.E,
        seq_ptr=np.cumsum([0] + [len(s) for s in ix.doc_tag_seq]).astype(np.int64),
        seq_ids=np.concatenate(ix.doc_tag_seq).astype(np.int32),
    )


def run_query(world, q, topn):
    try:
        res = world.find_similar_documents(q, topn=topn)
        return {"query": q, "topn": topn, "ids": [int(d) for d, _ in res],
                "scores": [float(s) for _, s in res]}
    except Exception as e:  # the reference has no error handling on this path: exceptions ARE the behaviour
        return {"query": q, "topn": topn, "error": type(e).__name__, "message": str(e)}


def fixture_main():
    ix = synth.generate_index(900, vocab_size=300, seed=20260101, tf_gt1_fraction=0.02)
    world = verbatim.ReferenceWorld(ix)
    names = ix.tag_names
    by_df = np.argsort(-ix.df, kind="stable")
    pop = [names[int(t)] for t in by_df[:12]]
    rare = [names[int(t)] for t in by_df if ix.df[t] > 0][-6:]
    absent = [names[int(t)] for t in np.nonzero(ix.df == 0)[0][:2]]

    queries = synth.generate_queries(ix, 48, seed=11)
    # quirks of SURVEY.md A.7 + grammar corners of A.1
    queries += [
        "%s:+0 %s" % (pop[0], pop[3]),                      # weight exactly 1000: NOT required
        "%s:-0 %s" % (pop[1], pop[4]),                      # weight 0: no exclusion
        "%s %s:2" % (pop[2], pop[2]),                       # duplicate tag: last weight wins for BM25
        "%s:3 %s:+2 %s:-1 %s" % (pop[0], pop[3], pop[9], pop[7]),
        "re:zero:2 %s" % pop[0],                             # tag containing ':'
        "fate_(series) %s:2" % pop[1],                       # parens (escaped only on the vector side)
        "3:4 %s" % pop[2],                                   # '3:4' parses as tag '3' weight 4 -> KeyError
        "tag:with:colons:+1",
        "%s  %s" % (pop[0], pop[1]),                         # double space -> empty token -> KeyError
        "no_such_tag",                                       # KeyError
        "%s:+" % pop[0],                                     # ValueError from int('+')
        "%s:1.5" % pop[0],                                   # literal tag 'x:1.5' -> KeyError
        " ".join("%s:+1" % r for r in rare[:3]),             # (almost) nothing survives -> NaN -> ValueError
        "%s:-1 %s:-1 %s:-1" % (pop[0], pop[1], pop[2]),      # only exclusions: bm25 all 0/-inf, negative weight sum
        "%s:-3" % pop[5],                                    # negative weight sum flips the query vector
        "%s:+2 %s:+1" % (pop[0], pop[1]),
        "%s:5" % rare[0],
        "%s:0" % pop[2],                                     # weight 0 -> vector weight sum 0 -> 1
    ]
    if absent:
        queries.append("%s %s" % (absent[0], pop[0]))        # term without idf entry (df == 0) -> idf 0
    results = [run_query(world, q, 100) for q in queries]
    results += [run_query(world, q, 800) for q in queries[:6]]     # webui.py:586 uses topn=800
    results += [run_query(world, q, 5) for q in queries[:3]]       # topn < PRF depth

    # vector-valued seams for a handful of queries
    seam = {}
    for k, q in enumerate([queries[0], queries[3], queries[48], queries[51]]):
        vec = world.normalize_and_apply_weight_doc2vec(q)
        sims = world.index[vec]
        wts = {}
        # query weights exactly as find_similar_documents builds them (webui.py:354-371)
        from oracle.port import parse_query_weights
        wts, _, _ = parse_query_weights(q, ix.token2id)
        bm25 = world.compute_bm25_scores(query_weights=wts)
        s2 = sims / sims.max() if sims.max() > 0 else sims
        b2 = bm25 / bm25.max() if bm25.max() > 0 else bm25
        final = world.BM25_WEIGHT * b2 + world.DOC2VEC_WEIGHT * s2
        seam["q%d_text" % k] = np.array(q)
        seam["q%d_dense" % k] = dense_query(vec, 300)
        seam["q%d_sims" % k] = sims
        seam["q%d_bm25" % k] = bm25
        seam["q%d_final" % k] = final
        seam["q%d_terms" % k] = np.array(list(wts.keys()), dtype=np.int64)
        seam["q%d_weights" % k] = np.array(list(wts.values()), dtype=np.float64)
    # the string-list form of compute_bm25_scores (webui.py:130-134)
    seam["terms_form_bm25"] = world.compute_bm25_scores(query_terms=[pop[0], pop[2], "no_such_tag"])
    seam["terms_form_tags"] = np.array([pop[0], pop[2], "no_such_tag"])

    # what the reference's own gen_and_save_bm25_index (genmodel.py:51-99) produced for this corpus
    ns = world.ns
    corpus = ns["bm25_corpus"]
    cptr = np.cumsum([0] + [len(d) for d in corpus]).astype(np.int64)
    idf_dense = np.zeros(ix.vocab_size)
    for t, v in ns["bm25_idf"].items():
        idf_dense[t] = v
    np.savez_compressed(os.path.join(OUT, "bm25_build_main.npz"),
                        doc_lengths=np.asarray(ns["bm25_doc_lengths"]), avgdl=np.float64(ns["bm25_avgdl"]), D=np.int64(ns["bm25_D"]),
                        idf=idf_dense, idf_terms=np.array(sorted(ns["bm25_idf"].keys()), dtype=np.int64),
                        corpus_ptr=cptr, corpus_terms=np.array([t for d in corpus for t in d.keys()], dtype=np.int64),
                        corpus_tfs=np.array([f for d in corpus for f in d.values()], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, "index_main.npz"), **index_arrays(ix))
    np.savez_compressed(os.path.join(OUT, "seams_main.npz"), **seam)
    with open(os.path.join(OUT, "results_main.json"), "w") as f:
        json.dump({"tag_names": names, "results": results}, f, indent=0)
    print("main: %d results, %d errors" % (len(results), sum("error" in r for r in results)))


def fixture_tiny():
    """N <= 10 takes the no-PRF branch (webui.py:247-253)."""
    ix = synth.generate_index(9, vocab_size=40, seed=77)
    world = verbatim.ReferenceWorld(ix)
    names = ix.tag_names
    present = [names[int(t)] for t in np.argsort(-ix.df, kind="stable")[:6]]
    queries = ["%s" % present[0], "%s:2 %s" % (present[1], present[2]), "%s:+1" % present[0],
               "%s:-1 %s" % (present[0], present[3]),
               " ".join("%s:-1" % p for p in present)]
    results = [run_query(world, q, 100) for q in queries] + [run_query(world, queries[0], 3)]
    np.savez_compressed(os.path.join(OUT, "index_tiny.npz"), **index_arrays(ix))
    with open(os.path.join(OUT, "results_tiny.json"), "w") as f:
        json.dump({"tag_names": names, "results": results}, f, indent=0)
    print("tiny: %d results, %d errors" % (len(results), sum("error" in r for r in results)))


def fixture_filter():
    """Known-answer cases for filter_searched_result (webui.py:63-80), incl. SURVEY A.6."""
    world_ns = verbatim.ReferenceWorld(synth.generate_index(12, vocab_size=40, seed=5)).ns
    f = world_ns["filter_searched_result"]
    cases = [
        [1, .9, .9 - 1e-7, .8, .7, .6, .5, .4, .4 - 1e-7, .3],
        [1, .9, .9 - 1e-7, .8],                     # prefix of the above: only one near-tie
        [1.0, 1.0, 1.0, .5, .25],
        [1.0, .5, .5, .5 - 1e-9, .2, -1.0],
        [.9, .8, .7, 0.0, -.1, float("-inf"), float("-inf")],
        [1.0] * 10 + [.99, .98, .98 - 5e-7, .97 - 1e-8, .97 - 2e-8, .5],
        [.5],
        [2.0, 1.0, 1.0 - 1e-7, 1.0 - 3e-7, .1],
    ]
    out = []
    for c in cases:
        lst = [(i * 3 + 1, float(s)) for i, s in enumerate(c)]
        res = f(lst)
        out.append({"input": [[d, (s if np.isfinite(s) else "-inf")] for d, s in lst],
                    "ids": [int(d) for d, _ in res], "scores": [float(s) for _, s in res]})
    with open(os.path.join(OUT, "filter_cases.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("filter: %d cases" % len(out))


if __name__ == "__main__":
    if not verbatim.available():
        raise SystemExit("needs the reference at %s" % verbatim.REFERENCE_DIR)
    os.makedirs(OUT, exist_ok=True)
    fixture_main()
    fixture_tiny()
    fixture_filter()
