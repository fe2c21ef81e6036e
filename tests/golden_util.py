"""Loads the committed golden fixtures (made by oracle/make_golden.py from the reference itself)."""
import json
import os

import numpy as np

import ais_b200  # noqa: F401
from ais_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_index(name: str) -> synth.SynthIndex:
    z = np.load(os.path.join(GOLDEN, "index_%s.npz" % name))
    v = int(z["vocab_size"])
    infer = synth.InferVectorStub(v, int(z["seed"]))
    infer.E = z["E"]
    seq_ptr, seq_ids = z["seq_ptr"], z["seq_ids"]
    seqs = [seq_ids[seq_ptr[i]: seq_ptr[i + 1]] for i in range(int(z["n_docs"]))]
    names = load_results(name)["tag_names"]
    return synth.SynthIndex(
        n_docs=int(z["n_docs"]), vocab_size=v, seed=int(z["seed"]), row_ptr=z["row_ptr"], term_ids=z["term_ids"],
        tfs=z["tfs"], doc_len=z["doc_len"], avgdl=np.float64(z["avgdl"]), idf=z["idf"], df=z["df"], rows=z["rows"],
        tag_names=names, infer=infer, doc_tag_seq=seqs, popularity=synth.zipf_popularity(v))


def load_results(name: str):
    with open(os.path.join(GOLDEN, "results_%s.json" % name)) as f:
        return json.load(f)


def load_seams():
    return np.load(os.path.join(GOLDEN, "seams_main.npz"))


def load_filter_cases():
    with open(os.path.join(GOLDEN, "filter_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        c["input"] = [(d, float("-inf") if s == "-inf" else float(s)) for d, s in c["input"]]
    return cases


def run_and_capture(fn, query, topn):
    """-> ('ok', ids, scores) or ('err', type_name, message)"""
    try:
        res = fn(query, topn)
    except Exception as e:
        return ("err", type(e).__name__, str(e))
    return ("ok", [int(d) for d, _ in res], [float(s) for _, s in res])
