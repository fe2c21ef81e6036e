"""Helpers of the parity tests at the BENCHMARKED sizes (>= 1 M docs): the on-device generator bench.py uses
(ais_b200.synth_torch) staged into an engine, plus a host view of the same shard for the oracle port.

The stand-in for gensim's ``infer_vector`` on these indexes is, by definition, "whatever produced the stored row":
``infer_vector(tags of doc d) := rows[d]`` and ``infer_vector([tag]) := E[tag]`` - exactly what bench.py's device PRF
mode (stored rows) and its query generator assume, so the oracle below checks the very path that is timed.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

import ais_b200  # noqa: F401
from ais_b200 import engine as E, query as Q, shard as SH, synth_torch
from oracle import port

VOCAB = 10861
SEED = 20260101


class _Infer:
    def __init__(self, emb: np.ndarray):
        self.E = emb

    def one(self, tag_ids: Sequence[int]) -> np.ndarray:
        ids = list(tag_ids)
        assert len(ids) == 1, "single-tag inference only (query side, webui.py:106)"
        return self.E[ids[0]]


class TorchIndexView:
    """What oracle.port.OraclePort(faithful=False) needs, over host copies of a synth_torch corpus."""

    def __init__(self, n_docs, rows, post_ptr, post_doc, doc_len, idf, df, avgdl, emb):
        self.n_docs = int(n_docs)
        self.vocab_size = len(idf)
        self.rows = rows
        self._post = (post_ptr, post_doc, np.ones(len(post_doc), dtype=np.int32))
        self.doc_len = doc_len
        self.avgdl = np.float64(avgdl)
        self.idf = idf
        self.df = df
        self.tag_names = ["t%d" % i for i in range(self.vocab_size)]
        self.token2id: Dict[str, int] = {t: i for i, t in enumerate(self.tag_names)}
        self.infer = _Infer(emb)

    def postings(self):
        return self._post

    def bm25_idf_dict(self):
        return {int(t): np.float64(self.idf[t]) for t in np.nonzero(self.df)[0]}


class StoredRowOracle(port.OraclePort):
    """OraclePort whose doc re-inference (webui.py:182-187) returns the stored row (see the module docstring)."""

    def doc_vector_pairs(self, doc_id_1based: int):
        v = self.idx.rows[doc_id_1based - 1]
        return [(i, val) for i, val in enumerate(v)]


def build_corpus(n_docs: int, device: int = 0, max_batch: int = 1, n_shards: int = 1, **params):
    """-> (engines [one per shard, all on `device`], TorchIndexView of the WHOLE corpus).  The rows / postings of every
    shard are generated on the device straight into the engine (as bench.py does) and copied back for the oracle."""
    import torch
    dev = torch.device("cuda", device)
    engines, rows_h, ptrs, docs, lens = [], [], [], [], []
    df_total = torch.zeros((VOCAB,), dtype=torch.int64, device=dev)
    shards = []
    tot_len = 0
    for r in range(n_shards):
        lo, hi = SH.shard_bounds(n_docs, n_shards, r)
        eng = E.SearchEngine(device=device, max_batch=max_batch, **params)
        rows = eng.rows_tensor(hi - lo)
        sh = synth_torch.generate_shard(lo, hi, rows, vocab=VOCAB, seed=SEED)
        df_total += sh.df
        tot_len += sh.total_len
        shards.append((eng, sh, lo, hi))
        rows_h.append(rows.cpu().numpy())
    dfd = df_total.to(torch.float64)
    idf = torch.where(df_total > 0, torch.log(1 + (n_docs - dfd + 0.5) / (dfd + 0.5)), torch.zeros_like(dfd))
    avgdl = float(tot_len) / float(n_docs)
    for eng, sh, lo, hi in shards:
        eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
        eng.set_shard(lo, n_docs)
        engines.append(eng)
        ptrs.append(sh.post_ptr.cpu().numpy())
        docs.append(sh.post_doc.cpu().numpy().astype(np.int64) + lo)
        lens.append(sh.doc_len.cpu().numpy())
    # whole-corpus posting lists (global doc ids ascending per term): concatenate the shards' slices term by term
    if n_shards == 1:
        post_ptr, post_doc = ptrs[0], docs[0].astype(np.int32)
    else:
        df_h = df_total.cpu().numpy()
        post_ptr = np.zeros(VOCAB + 1, dtype=np.int64)
        np.cumsum(df_h, out=post_ptr[1:])
        post_doc = np.empty(int(post_ptr[-1]), dtype=np.int32)
        fill = post_ptr[:-1].copy()
        for p, d in zip(ptrs, docs):
            cnt = np.diff(p)
            term_of = np.repeat(np.arange(VOCAB), cnt)
            pos = fill[term_of] + (np.arange(len(d)) - p[term_of])
            post_doc[pos] = d
            fill += cnt
    emb = synth_torch.embedding_table(VOCAB, SEED, dev).cpu().numpy()
    view = TorchIndexView(n_docs, np.concatenate(rows_h) if n_shards > 1 else rows_h[0], post_ptr, post_doc,
                          np.concatenate(lens), idf.cpu().numpy(), df_total.cpu().numpy(), avgdl, emb)
    del shards
    torch.cuda.empty_cache()
    return engines, view


def make_queries(view: TorchIndexView, n: int, seed: int):
    """(texts, engine queries): the benchmark's query generator; the query vectors go through the product's own
    query.py (webui.py:82-117) with E[tag] as the per-tag inference, exactly like the oracle."""
    texts, _ = synth_torch.make_queries(view.df, view.infer.E, n, seed=seed)
    t2i = view.token2id
    infer = lambda words: view.infer.E[t2i[words[0]]]
    return texts, [Q.make_query(t, t2i, infer) for t in texts]


def oracle_results(P: port.OraclePort, texts: Sequence[str], topn: int, workers: int = 0):
    """[capture(find_similar_documents)] for every text; numpy releases the GIL in its O(N) kernels, so threads help."""
    from gpu_util import capture
    workers = workers or min(16, os.cpu_count() or 1)
    with cf.ThreadPoolExecutor(workers) as ex:
        return list(ex.map(lambda t: capture(P.find_fast, t, topn), texts))


def engine_outcome(ids, scores, counts, status, j):
    from ais_b200.engine import raise_for_status
    try:
        raise_for_status(int(status[j]))
        c = int(counts[j])
        return ("ok", ids[j, :c].tolist(), scores[j, :c].tolist())
    except Exception as e:   # noqa: BLE001
        return ("err", type(e).__name__, str(e))
