import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ais_b200
from ais_b200 import engine as E, query as Q, synth
idx = synth.generate_index(900, vocab_size=300, seed=1)
t2i = idx.token2id
infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
eng = E.SearchEngine.from_index(idx)
q = Q.make_query("t80:1 t24:+1", t2i, infer)
step = sys.argv[1] if len(sys.argv) > 1 else "all"
if step in ("dot", "all"):
    s = eng.dot_scores(q.vec); print("dot ok", np.abs(s - idx.rows @ q.vec).max())
if step in ("bm25", "all"):
    b = eng.bm25_scores(q.term_ids, q.weights); print("bm25 ok", np.isneginf(b).sum())
if step in ("final", "all"):
    f = eng.final_scores(q); print("final ok", np.nanmax(f))
if step in ("search", "all"):
    r = eng.search([q], 100, E.PRF_STORED_ROWS); print("search ok", r[0][:3])
