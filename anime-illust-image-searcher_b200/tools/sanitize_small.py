"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the batched and the single-query
path on a 6 000-doc index, PRF modes stored rows / full centroid / off, topn 100 and 1100 (full-sort route)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ais_b200  # noqa: F401,E402
from ais_b200 import engine as E, query as Q, synth  # noqa: E402

idx = synth.generate_index(6000, vocab_size=400, seed=9, tf_gt1_fraction=0.02)
t2i = idx.token2id
infer = lambda words: idx.infer.one([t2i[w] for w in words if w in t2i])
qs = [Q.make_query(t, t2i, infer) for t in synth.generate_queries(idx, 40, seed=1)]
for mb in (1, 40):
    eng = E.SearchEngine.from_index(idx, max_batch=mb)
    for mode in (E.PRF_STORED_ROWS, E.PRF_STORED_ROWS_FULL, E.PRF_OFF):
        for topn in (100, 1100):
            r = eng.search_raw(qs[: (3 if mb == 1 else 40)], topn, mode)
            print(mb, mode, topn, int(r[2].sum()), r[3].tolist()[:6])
    print(eng.stats())
    eng.close()
print("sanitize run done")
