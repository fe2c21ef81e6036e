#!/usr/bin/env python
"""bench.py - queries/sec of the query-time scoring path (BM25 + Doc2Vec dot + PRF re-rank + top-100)
over a synthetic 10 M-doc index, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--docs 10000000] [--batch 256] [--impl reference]

A "step" is ONE batch of `--batch` queries through the whole path (both passes over the doc vectors).
N > 1 (launched with torch.distributed.run): the 10 M docs are sharded by document across the ranks
(strong scaling: total work per query is fixed), NCCL carries the per-query records between stages.
Prints ONE JSON line on rank 0 (contract in the task statement): value / e2e / roofline / cpu_baseline /
clocks / gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec @10M docs top-100 (BM25+Doc2Vec+rerank)"
UNIT = "queries/s"
TOPN = 100
VOCAB = 10861
SEED = 20260101


def scan_kernel_for(batch: int):
    """(kernel name, key into profiles/traffic.json) of the scan launch that dominates a batch of this size"""
    left = min(batch, 256)
    if left >= 33:
        return "scan_tc_kernel<64> (tcgen05 kind::tf32 3xTF32, 64 queries per pass)", "scan_tc64"
    if left >= 9:
        return "scan_tc_kernel<32> (tcgen05 kind::tf32 3xTF32, 32 queries per pass)", "scan_tc32"
    if left >= 5:
        return "scan_mma_kernel<8> (mma.sync 3xTF32)", "scan_mma8"
    return "scan_kernel<%d> (fp32 SIMT, lane per row)" % (1 if left <= 1 else 2 if left <= 2 else 4), "scan_simt"


def load_traffic(key: str, n_docs_local: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at 10 M docs), scaled to this shard's rows; None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[key]
        return float(t["dram_bytes_per_launch"]) * n_docs_local / float(t["docs"]), t["source"]
    except Exception:
        return None, None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, text)
        self.t_from = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """only samples arriving after this call count (the sampler is started before the warm-up)"""
        self.t_from = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t_arr, ln in self.lines:
            if t_arr < self.t_from:
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_docs: int, n_queries: int, n_docs_metric: int, warmup: int = 1):
    """The reference's CPU algorithm (oracle/port.py, faithful=True: its own data structures and Python loops)
    on a bounded sample of the workload; the O(N) cost is extrapolated linearly to `n_docs_metric` docs."""
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    import ais_b200  # noqa: F401
    from ais_b200 import synth
    from oracle import port
    idx = synth.generate_index(sample_docs, vocab_size=VOCAB, seed=SEED, keep_sequences=False)
    P = port.OraclePort(idx, faithful=True)
    queries = synth.generate_queries(idx, n_queries + warmup, seed=7)
    times = []
    for i, q in enumerate(queries):
        t0 = time.perf_counter()
        try:
            P.find_similar_documents(q, TOPN)
        except (ValueError, ZeroDivisionError):
            pass                                    # fewer than 10 survivors: the reference raises too
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_query = float(np.mean(times))
    qps_sample = 1.0 / per_query
    return qps_sample * sample_docs / n_docs_metric, qps_sample, per_query


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample_docs
    t_all = time.perf_counter()
    # every step = one query over the bounded sample
    n = args.steps
    value, qps_sample, per_query = cpu_reference_rate(sample, n, args.docs, warmup=max(1, args.warmup))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_query * 1e3 * args.docs / sample, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": "%d docs, V=%d, ~30 tags/doc, 300-d fp32 rows, single weighted queries with +required/-exclude, "
                               "top-%d, PRF re-rank" % (args.docs, VOCAB, TOPN)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "host_cores_available": cores,
                         "sample": "oracle/port.py faithful=True (the reference's list-of-dicts BM25 loops, Python sorts, "
                                   "numpy sgemv) timed on %d docs: %.3f s/query = %.3f q/s, scaled linearly (O(N)) to %d docs; "
                                   "the reference's loops are single-threaded, BLAS may use all %d cores"
                                   % (sample, per_query, qps_sample, args.docs, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=256,
                    help="queries per engine batch (1..256; BASELINE configs[2]: 10 M docs, batches of 256 queries with PRF); "
                         "up to 64 share one pass over the doc vectors")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-docs", type=int, default=100_000)
    ap.add_argument("--cpu-queries", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra batch-1 / batch-4 operating points")
    ap.add_argument("--sweep", action="store_true", help="also time batch sizes 1..128 (extra key batch_sweep): SIMT -> mma.sync -> tcgen05 crossover")
    ap.add_argument("--prf", default="stored_rows", choices=["stored_rows", "full"],
                    help="stored_rows: the reference's re-query (collapsed centroid [c,0,...,0]: served by the one-sector-per-doc "
                         "column scan); full: the un-collapsed centroid (a second dense pass over the rows)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ais_b200  # noqa: F401
    from ais_b200 import engine as E, shard, synth_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ais_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- stage the shard [lo, hi) of the synthetic index into HBM -----------------------------------
    t_build = time.perf_counter()
    lo, hi = shard.shard_bounds(args.docs, world, rank)
    eng = E.SearchEngine(device=local_rank, max_batch=max(args.batch, 256) if args.sweep else args.batch)
    rows = eng.rows_tensor(hi - lo)
    sh = synth_torch.generate_shard(lo, hi, rows, vocab=VOCAB, seed=SEED)
    idf, avgdl, df = synth_torch.global_stats(sh, args.docs)
    eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
    eng.set_shard(lo, args.docs)
    nnz_local = int(sh.post_doc.numel())
    E_host = synth_torch.embedding_table(VOCAB, SEED, dev).cpu().numpy()
    df_host = df.cpu().numpy()
    del sh, rows
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    n_pool = max(64, args.batch * 4)
    texts, parsed = synth_torch.make_queries(df_host, E_host, n_pool, seed=7)
    pool = [E.Query(*p) for p in parsed]
    eng.use_torch_stream()                               # CUDA events below see the engine's kernels
    S = shard.ShardedSearch([eng], args.docs) if world > 1 else None
    mode = E.PRF_STORED_ROWS if args.prf == "stored_rows" else E.PRF_STORED_ROWS_FULL

    def search(qs):
        # one GPU: the C-ABI call ais_search; several: the staged calls with NCCL between them
        return S.search_raw(qs, TOPN, mode) if S is not None else eng.search_raw(qs, TOPN, mode)

    def batch_at(step, b):
        return [pool[(step * b + j) % n_pool] for j in range(b)]

    def run_steps(n_steps, b, first_step=0):
        """device time (CUDA events on the engine's stream = torch's current stream), max over ranks"""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        done = []
        for s in range(n_steps):
            qs = batch_at(first_step + s, b)
            ids, scores, counts, status, _ = search(qs)          # host query buffers in, host result arrays out
            done.append((qs, ids, scores, counts, status))
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        # bookkeeping outside the timed region: bytes moved per step, result checksum
        h2d = d2h = 0
        n_results = 0
        checksum = 0
        for qs, ids, scores, counts, status in done:
            h2d += sum(q.vec.nbytes + q.term_ids.nbytes + q.weights.nbytes for q in qs)
            d2h += ids.nbytes + scores.nbytes + counts.nbytes + status.nbytes
            n_results += int(counts.sum())
            for j in range(len(qs)):
                checksum = (checksum * 1000003 + int(ids[j, :counts[j]].sum()) + 7 * int(counts[j])) % (1 << 61)
        dev_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), h2d // max(n_steps, 1), d2h // max(n_steps, 1), n_results, checksum

    # ---- warm-up, then the timed region ----------------------------------------------------------------
    b = args.batch
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    run_steps(args.warmup, b, 0)
    eng.set_profiling(True)
    eng.reset_stats()
    sampler.mark()
    dev_ms, wall_ms, h2d, d2h, n_results, checksum = run_steps(args.steps, b, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()
    eng.set_profiling(False)

    n_queries = args.steps * b
    value = n_queries / (dev_ms * 1e-3)
    e2e = n_queries / (wall_ms * 1e-3)
    peak, peak_src = load_peaks()
    scan_ms = st["scan_ms_total"] / max(1, st["scan_launches"])
    scan_bytes = (hi - lo) * 1200                       # every stored fp32 row read exactly once per launch
    achieved = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    # whole-step figure: algorithmic bytes of a step (2 passes over the rows + the posting ids of the batch)
    post_bytes = 0
    for s in range(args.steps):
        for q in batch_at(args.warmup + s, b):
            post_bytes += int(sum(4 * df_host[t] for t in q.term_ids)) // world
    # pass 2: the reference's collapsed re-query needs column 0 only (4 B per doc); the dense variant a second full pass
    requery_bytes = (hi - lo) * 4 if st["column_scan_launches"] > 0 else scan_bytes
    step_bytes = scan_bytes + requery_bytes + post_bytes / args.steps
    step_gbs = step_bytes / (dev_ms / args.steps * 1e-3) / 1e9

    kernel_name, traffic_key = scan_kernel_for(b)
    traffic, traffic_src = load_traffic(traffic_key, hi - lo)

    sweep = None
    if args.sweep or not args.no_modes:
        # other operating points of the same engine: batch 1 = single-query latency mode (both scans at the HBM
        # roofline), batch 4 = largest batch of the fp32 SIMT scan; --sweep adds the rest
        sweep = {}
        for bb in ((1, 2, 4, 8, 16, 32, 64, 128, 256) if args.sweep else (1, 64)):
            if bb > eng.params.max_batch:
                break
            run_steps(2, bb, 0)
            eng.set_profiling(True); eng.reset_stats()
            ms, wms, _, _, _, _ = run_steps(max(4, args.steps // 2), bb, 2)
            s2 = eng.stats(); eng.set_profiling(False)
            k = max(4, args.steps // 2)
            sm = s2["scan_ms_total"] / max(1, s2["scan_launches"])
            sweep[str(bb)] = {"qps": k * bb / (ms * 1e-3), "e2e_qps": k * bb / (wms * 1e-3), "scan_kernel": scan_kernel_for(bb)[0],
                              "scan_ms": sm,
                              "scan_gbs": scan_bytes / (sm * 1e-3) / 1e9, "scan_frac_of_peak": scan_bytes / (sm * 1e-3) / 1e9 / peak,
                              "scan_share": s2["scan_ms_total"] / ms,
                              "whole_step_frac_of_peak": ((scan_bytes + ((hi - lo) * 4 if s2["column_scan_launches"] > 0 else scan_bytes)) * bb)
                                                         / (ms / k * 1e-3) / 1e9 / peak / bb}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, qps_s, per_q = cpu_reference_rate(args.cpu_sample_docs, args.cpu_queries, args.docs)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": "oracle/port.py faithful=True on a %d-doc index of the same generator family, %d queries: "
                         "%.3f s/query = %.3f q/s, scaled linearly (O(N)) to %d docs; the reference's Python loops "
                         "are single-threaded" % (args.cpu_sample_docs, args.cpu_queries, per_q, qps_s, args.docs)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 dot / f64 BM25+combine", "data": "synthetic",
            "config": {"workload": "%d docs sharded over %d GPU(s), V=%d, ~30 distinct tags/doc (%d postings on rank 0), 300-d fp32 rows; "
                                   "weighted queries with +required/-exclude, top-%d, PRF re-rank (device stored-rows mode); "
                                   "%d queries per engine batch" % (args.docs, world, VOCAB, nnz_local, TOPN, b),
                       "docs": args.docs, "batch": b, "topn": TOPN, "prf": args.prf,
                       "requery": "column scan (single non-zero component, SURVEY.md A.5)" if st["column_scan_launches"] > 0 else "dense scan", "parallelism": "doc-shard x%d" % world,
                       "l2": "inputs larger than L2 (%.1f GB of rows per GPU re-read every pass)" % (scan_bytes / 1e9)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "host wall clock around the public API call (ctypes -> C ABI) with host query buffers and host result arrays"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "%s, %d launches" % (kernel_name, st["scan_launches"]),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "traffic_frac": (traffic / (scan_ms * 1e-3) / 1e9 / peak) if traffic and scan_ms > 0 else None,
                         "peak_source": peak_src, "bytes_per_launch": scan_bytes, "ms_per_launch": scan_ms,
                         "scan_share_of_step": st["scan_ms_total"] / dev_ms,
                         "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / peak}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "results_per_step": n_results / args.steps,
            "results_checksum": checksum,      # same queries -> same value for every --gpus N (doc ids of all results)
            "fullsort_fallbacks": int(st["fullsort_fallbacks"]) + (S.fullsort_fallbacks if S is not None else 0),
            "index_build_s": t_build,
        }
        if sweep:
            line["batch_sweep" if args.sweep else "other_batch_sizes"] = sweep
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
