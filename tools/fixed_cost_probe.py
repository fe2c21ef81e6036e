"""What a batch costs when the index is tiny: stages N docs (default 8192), runs batches of 256 queries.  Run it under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel launch list of the part of a step that does not shrink
with the shard (profiles/r02_*_fixed_launches.csv).

    python tools/fixed_cost_probe.py [--docs 8192] [--batch 256] [--steps 3]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import ais_b200  # noqa: F401
    from ais_b200 import engine as E, shard, synth_torch
    dev = torch.device("cuda", 0)
    eng = E.SearchEngine(device=0, max_batch=args.batch)
    rows = eng.rows_tensor(args.docs)
    sh = synth_torch.generate_shard(0, args.docs, rows, vocab=10861, seed=1234)
    idf, avgdl, df = synth_torch.global_stats(sh, args.docs)
    eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
    eng.set_shard(0, args.docs)
    emb = synth_torch.embedding_table(10861, 1234, dev).cpu().numpy()
    texts, parsed = synth_torch.make_queries(df.cpu().numpy(), emb, args.batch, seed=7)
    qs = [E.Query(*p) for p in parsed]
    for _ in range(2):
        eng.search_raw(qs, 100, E.PRF_STORED_ROWS)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.search_raw(qs, 100, E.PRF_STORED_ROWS)
    torch.cuda.synchronize()
    print("ms per step: %.3f" % ((time.perf_counter() - t0) * 1e3 / args.steps))


if __name__ == "__main__":
    main()
