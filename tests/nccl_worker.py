"""Worker of tests/test_gpu_scale.py::test_nccl_two_ranks_vs_oracle (one process per GPU, torch.distributed.run).

Every rank stages its contiguous shard of a synth_torch corpus (the generator bench.py times), the batch goes through
shard.ShardedSearch with the records exchanged over NCCL, and rank 0 compares every result with the ORACLE
(oracle/port.py over a host copy of the whole corpus) - not with a single-engine run.  Exit code 0 = parity.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/nccl_worker.py --docs 1000000 --batch 64
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--topn", type=int, default=100)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import ais_b200  # noqa: F401
    from ais_b200 import engine as E, shard, synth_torch
    import scale_util as SU
    from gpu_util import assert_same_or_filter_unstable

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    lo, hi = shard.shard_bounds(args.docs, world, rank)
    eng = E.SearchEngine(device=local, max_batch=args.batch)
    rows = eng.rows_tensor(hi - lo)
    sh = synth_torch.generate_shard(lo, hi, rows, vocab=SU.VOCAB, seed=SU.SEED)
    idf, avgdl, df = synth_torch.global_stats(sh, args.docs)          # NCCL all-reduce of df / total length
    eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
    eng.set_shard(lo, args.docs)
    del sh, rows
    S = shard.ShardedSearch([eng], args.docs)

    emb = synth_torch.embedding_table(SU.VOCAB, SU.SEED, dev).cpu().numpy()
    df_h = df.cpu().numpy()
    texts, _ = synth_torch.make_queries(df_h, emb, args.batch, seed=909)
    t2i = {"t%d" % i: i for i in range(SU.VOCAB)}
    from ais_b200 import query as Q
    qs = [Q.make_query(t, t2i, lambda words: emb[t2i[words[0]]]) for t in texts]
    res = S.search_raw(qs, args.topn, E.PRF_STORED_ROWS)
    torch.cuda.synchronize()
    dist.barrier()

    ok = 1
    report = {}
    if rank == 0:
        try:
            # the whole corpus once more on this GPU (same chunk seeds -> same docs), copied to the host for the oracle
            engines, view = SU.build_corpus(args.docs, device=local, max_batch=1)
            for e in engines:
                e.close()
            assert np.array_equal(view.df, df_h) and abs(float(view.avgdl) - avgdl) == 0.0
            P = SU.StoredRowOracle(view)
            want = SU.oracle_results(P, texts, args.topn)
            n_ok = 0
            for j, text in enumerate(texts):
                got = SU.engine_outcome(res[0], res[1], res[2], res[3], j)
                assert_same_or_filter_unstable(got, want[j], lambda t=text: P.find_sorted_arrays(t), 1e-6, args.topn, ("nccl", j, text))
                n_ok += want[j][0] == "ok"
            report = {"ranks": world, "docs": args.docs, "batch": args.batch, "checked_vs_oracle": len(texts),
                      "ok_results": n_ok, "backend": dist.get_backend(), "fullsort_fallbacks": S.fullsort_fallbacks}
            print("NCCL_PARITY " + json.dumps(report), flush=True)
            if args.out:
                with open(args.out, "w") as f:
                    json.dump(report, f)
        except BaseException as exc:   # noqa: BLE001 - reported through the exit code
            import traceback
            traceback.print_exc()
            ok = 0
    flag = torch.tensor([ok], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
