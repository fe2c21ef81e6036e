"""One engine over the whole synthetic index vs the same index cut into E doc shards (E engines of THIS process driven by
shard.ShardedSearch - the code path of `bench.py --gpus E`, collectives replaced by local reductions): the result of
every query must be identical (ids, counts; scores to 1e-12).  Prints the queries that differ.

    python tools/shard_consistency.py --docs 10000000 --shards 8 --batch 256 --steps 20
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--topn", type=int, default=100)
    args = ap.parse_args()
    import numpy as np
    import torch
    import ais_b200  # noqa: F401
    from ais_b200 import engine as E, shard, synth_torch
    VOCAB, SEED = 10861, 1234
    dev = torch.device("cuda", 0)

    def stage(lo, hi, idf=None, avgdl=None):
        eng = E.SearchEngine(device=0, max_batch=args.batch)
        rows = eng.rows_tensor(hi - lo)
        sh = synth_torch.generate_shard(lo, hi, rows, vocab=VOCAB, seed=SEED)
        if idf is None:
            idf, avgdl, df = synth_torch.global_stats(sh, args.docs)
        else:
            df = None
        eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
        eng.set_shard(lo, args.docs)
        return eng, idf, avgdl, df

    whole, idf, avgdl, df = stage(0, args.docs)
    parts = []
    for r in range(args.shards):
        lo, hi = shard.shard_bounds(args.docs, args.shards, r)
        parts.append(stage(lo, hi, idf, avgdl)[0])
    S = shard.ShardedSearch(parts, args.docs)
    emb = synth_torch.embedding_table(VOCAB, SEED, dev).cpu().numpy()
    n_pool = max(64, min(args.batch, 256) * 4)
    texts, parsed = synth_torch.make_queries(df.cpu().numpy(), emb, n_pool, seed=7)
    pool = [E.Query(*p) for p in parsed]
    bad = 0
    for s in range(args.steps):
        idx = [(s * args.batch + j) % n_pool for j in range(args.batch)]
        qs = [pool[i] for i in idx]
        a = whole.search_raw(qs, args.topn, E.PRF_STORED_ROWS)
        b = S.search_raw(qs, args.topn, E.PRF_STORED_ROWS)
        for j in range(args.batch):
            ca, cb = int(a[2][j]), int(b[2][j])
            same = ca == cb and int(a[3][j]) == int(b[3][j]) and np.array_equal(a[0][j, :ca], b[0][j, :cb]) and \
                np.allclose(a[1][j, :ca], b[1][j, :cb], rtol=1e-12, atol=0)
            if not same:
                bad += 1
                if bad <= 10:
                    print("DIFF step %d query %d (pool %d) %r" % (s, j, idx[j], texts[idx[j]]))
                    print("  whole : count %d status %d" % (ca, int(a[3][j])))
                    print("  shards: count %d status %d" % (cb, int(b[3][j])))
                    n = min(ca, cb)
                    d = np.nonzero(a[0][j, :n] != b[0][j, :n])[0]
                    if len(d):
                        p = int(d[0])
                        print("  first differing rank %d: ids %r vs %r" % (p, a[0][j, max(0, p - 2):p + 3].tolist(), b[0][j, max(0, p - 2):p + 3].tolist()))
                        print("  scores %r vs %r" % (a[1][j, max(0, p - 2):p + 3].tolist(), b[1][j, max(0, p - 2):p + 3].tolist()))
    print("SHARD_CONSISTENCY shards=%d steps=%d batch=%d differing=%d fullsort_fallbacks=%d" % (args.shards, args.steps, args.batch, bad, S.fullsort_fallbacks))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
