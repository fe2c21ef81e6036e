import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ais_b200
from ais_b200 import engine as E, shard, synth_torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
V = 10861; SEED = 20260101
dev = torch.device("cuda", 0)
def build(lo, hi, idf=None, avgdl=None):
    eng = E.SearchEngine(device=0, max_batch=B)
    rows = eng.rows_tensor(hi - lo)
    sh = synth_torch.generate_shard(lo, hi, rows, vocab=V, seed=SEED)
    return eng, sh
eng1, sh1 = build(0, N)
idf, avgdl, df = synth_torch.global_stats(sh1, N)
eng1.load_bm25(sh1.post_ptr, sh1.post_doc, None, idf, sh1.doc_len, avgdl); eng1.set_shard(0, N)
Eh = synth_torch.embedding_table(V, SEED, dev).cpu().numpy()
texts, parsed = synth_torch.make_queries(df.cpu().numpy(), Eh, 256, seed=7)
pool = [E.Query(*p) for p in parsed]
engines = []
for r in range(4):
    lo, hi = shard.shard_bounds(N, 4, r)
    e, sh = build(lo, hi)
    e.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl); e.set_shard(lo, N)
    engines.append(e)
S = shard.ShardedSearch(engines, N)
eng1.use_torch_stream()
nbad = 0
for lo in range(0, 256, B):
    qs = pool[lo:lo + B]
    a = eng1.search_raw(qs, 100, E.PRF_STORED_ROWS)
    b = S.search_raw(qs, 100, E.PRF_STORED_ROWS)
    for j in range(len(qs)):
        ca, cb = int(a[2][j]), int(b[2][j])
        same = ca == cb and np.array_equal(a[0][j, :ca], b[0][j, :cb]) and np.array_equal(a[1][j, :ca], b[1][j, :cb])
        if not same or a[3][j] != b[3][j]:
            nbad += 1
            print("DIFF q", lo + j, texts[lo + j], "counts", ca, cb, "status", a[3][j], b[3][j])
            m = min(ca, cb)
            d = np.nonzero(a[0][j, :m] != b[0][j, :m])[0]
            print("   first id diff at", d[:5], a[0][j, :m][d[:3]], b[0][j, :m][d[:3]], "score diff max", np.abs(a[1][j, :m] - b[1][j, :m]).max() if m else None)
print("bad", nbad, "fallbacks", eng1.stats()["fullsort_fallbacks"], S.fullsort_fallbacks)
# stage-level comparison for the first 16 queries
qs = pool[:16]
r1 = eng1.search_raw(qs, 100, E.PRF_STORED_ROWS)
r4 = S.search_raw(qs, 100, E.PRF_STORED_ROWS)
for which in ("sim", "bm25", "fin", "rer"):
    for q in (0, 5):
        a = eng1.debug_read(which, q)
        b = np.concatenate([e.debug_read(which, q) for e in engines])
        fin = np.isfinite(a)
        print(which, q, "equal", np.array_equal(a, b), "maxdiff", np.abs(a[fin] - b[fin]).max() if fin.any() else None)
