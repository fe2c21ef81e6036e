"""First-light check of scan_tc_kernel: full sim matrix of a batch against the fp64 dot product."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ais_b200  # noqa
from ais_b200 import engine as E, synth

for n_docs in (20011, 128 * 148 * 3 + 5):
    idx = synth.generate_index(n_docs, vocab_size=500, seed=12)
    rng = np.random.default_rng(4)
    for mb in (32, 24, 64, 47, 100):
        eng = E.SearchEngine.from_index(idx, max_batch=mb)
        vecs = rng.standard_normal((mb, 300)).astype(np.float32)
        vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
        qs = [E.Query(v, np.array([0], np.int32), np.array([1.0])) for v in vecs]
        maxes = torch.empty((mb, 2), dtype=torch.float64, device="cuda")
        eng.stage_score(qs, maxes)
        torch.cuda.synchronize(); eng.synchronize()
        want = idx.rows.astype(np.float64) @ vecs.T.astype(np.float64)      # [n][mb]
        w32 = idx.rows @ vecs.T
        worst = 0.0
        for q in range(mb):
            got = eng.debug_read("sim", q)
            err = np.abs(got - want[:, q]).max() / np.abs(want[:, q]).max()
            worst = max(worst, err)
        e32 = (np.abs(w32 - want).max(axis=0) / np.abs(want).max(axis=0)).max()
        gm = maxes.cpu().numpy()[:, 1]
        print("n=%d mb=%d: worst rel-to-max err %.3e (numpy fp32: %.3e); max err %.3e" %
              (n_docs, mb, worst, e32, np.abs(gm - want.max(axis=0)).max() / np.abs(want).max()), flush=True)
        eng.close()
