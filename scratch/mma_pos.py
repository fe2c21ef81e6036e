import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ais_b200
from ais_b200 import engine as E
rng = np.random.default_rng(0)
X = rng.standard_normal((64, 300)).astype(np.float32)
rows = np.zeros((400, 300), np.float32)
offs = [0, 80, 168, 201, 300]
for o in offs: rows[o:o+64] = X
eng = E.SearchEngine(device=0, max_batch=16)
eng.load_vectors(rows)
eng.load_bm25(np.zeros(2, np.int64), np.zeros(0, np.int32), None, np.zeros(1), np.full(400, 5, np.int64), 5.0)
eng.set_shard(0, 400)
vecs = rng.standard_normal((16, 300)).astype(np.float32); vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
qs = [E.Query(v, np.zeros(0, np.int32), np.zeros(0)) for v in vecs]
maxes = torch.empty((16, 2), dtype=torch.float64, device="cuda")
eng.stage_score(qs, maxes); eng.synchronize()
ref = (X.astype(np.float64) @ vecs.T.astype(np.float64))
for q in (0, 7, 15):
    s = eng.debug_read("sim", q)
    base = s[0:64]
    print("q", q, "err vs f64", np.abs(base - ref[:, q]).max(), [bool(np.array_equal(s[o:o+64], base)) for o in offs], [float(np.abs(s[o:o+64]-base).max()) for o in offs])
