#!/usr/bin/env python
"""bench.py - queries/sec of the query-time scoring path (BM25 + Doc2Vec dot + PRF re-rank + top-100)
over a synthetic 10 M-doc index, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--docs 10000000] [--batch 256] [--impl reference]

A "step" is ONE batch of `--batch` queries through the whole path (both passes over the doc vectors).
N > 1 (launched with torch.distributed.run): the docs are sharded by document across the ranks
(strong scaling: total work per query is fixed), NCCL carries the per-query records between stages.
Prints ONE JSON line on rank 0 (contract in the task statement): value / e2e / roofline (+ kernels[]) /
cpu_baseline / clocks / gpu_launches / parity_checked.

`--impl reference` times the reference's own CPU implementation of the path (webui.py:345-390): the verbatim
reference functions when /root/reference is present (build container), else the oracle port with the reference's own
loop shapes (oracle/port.py faithful=True) - at 10^4, 10^5 and 10^6 docs, median + p95, BLAS threads 1 and all.
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


def workload_text(docs, world, batch):
    return ("%d docs sharded over %d GPU(s), V=%d, ~30 distinct tags/doc, 300-d fp32 rows; weighted queries with "
            "+required/-exclude, top-%d, PRF re-rank (device stored-rows mode = the reference's collapsed centroid); "
            "%d queries per engine batch" % (docs, world, VOCAB, TOPN, batch))


def scan_kernel_for(batch: int):
    """(kernel name, key into profiles/traffic.json, queries per pass) of the scan launch that dominates a batch of this size"""
    left = min(batch, 256)
    if left >= 33:
        if os.environ.get("AIS_SCAN_PAIR", "9") != "0":
            return "scan_pair_kernel<9> (tcgen05 cta_group::2 kind::tf32 3xTF32, CTA pairs, 64 queries per pass)", "scan_pair64", 64
        return "scan_tc_kernel<64> (tcgen05 kind::tf32 3xTF32, 64 queries per pass)", "scan_tc64", 64
    if left >= 9:
        return "scan_tc_kernel<32> (tcgen05 kind::tf32 3xTF32, 32 queries per pass)", "scan_tc32", 32
    if left >= 5:
        return "scan_mma_kernel<8> (mma.sync 3xTF32)", "scan_mma8", 8
    return "scan_kernel<%d> (fp32 SIMT, lane per row)" % (1 if left <= 1 else 2 if left <= 2 else 4), "scan_simt", max(1, left)


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
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))), \
            "measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops_sustained)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.4 PFLOP/s sustained bf16)"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


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
        self.t_to = float("inf")

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

    def mark_end(self):
        self.t_to = time.perf_counter() + 0.15

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
            if t_arr < self.t_from or t_arr > self.t_to:
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
# The reference's CPU path
def _reference_world(n_docs: int):
    """(callable find_similar_documents(text, topn), query texts generator, kind) over a synth index of n_docs docs.
    kind "reference": the reference's own functions executed verbatim (oracle/verbatim.py; only where /root/reference
    exists); kind "port": oracle/port.py faithful=True (its list-of-dicts BM25 loops, Python sorts, numpy sgemv)."""
    import ais_b200  # noqa: F401
    from ais_b200 import synth
    from oracle import port, verbatim
    idx = synth.generate_index(n_docs, vocab_size=VOCAB, seed=SEED, keep_sequences=n_docs <= 200_000,
                               rows="random" if n_docs > 200_000 else "infer")
    if verbatim.available() and n_docs <= 200_000:
        W = verbatim.ReferenceWorld(idx, use_reference_bm25_builder=False)
        return W.find_similar_documents, idx, "reference"
    P = port.OraclePort(idx, faithful=True)
    return P.find_similar_documents, idx, "port"


def _time_queries(fn, texts, warmup: int):
    times = []
    for i, q in enumerate(texts):
        t0 = time.perf_counter()
        try:
            fn(q, TOPN)
        except (ValueError, ZeroDivisionError):
            pass                                    # fewer than 10 survivors: the reference raises too
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def cpu_reference_measure(n_docs: int, n_queries: int, warmup: int, blas_threads):
    """median / p95 seconds per query of the reference's path on an n_docs index, per BLAS thread setting"""
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    from ais_b200 import synth
    fn, idx, kind = _reference_world(n_docs)
    texts = synth.generate_queries(idx, n_queries + warmup, seed=7)
    out = {"docs": n_docs, "kind": kind, "queries": n_queries, "per_threads": {}}
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    for nt in blas_threads:
        if threadpool_limits is not None:
            with threadpool_limits(limits=nt):
                ts = _time_queries(fn, texts, warmup)
        else:
            ts = _time_queries(fn, texts, warmup)
        out["per_threads"][str(nt)] = {"median_s": float(np.median(ts)), "p95_s": float(np.percentile(ts, 95)),
                                       "mean_s": float(np.mean(ts)), "qps": 1.0 / float(np.median(ts))}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  The timed steps are
    `--steps` single queries on a 10^5-doc index (what fits a driver run); 10^4 and 10^6 are measured beside it with
    fewer queries, and the 10 M figure (`value`: the metric is defined at 10 M docs) is the O(N) extrapolation of the
    largest measured size - stated as such in `extrapolated`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    cores = os.cpu_count() or 1
    threads = [1] if cores == 1 else [1, cores]
    main = cpu_reference_measure(args.cpu_sample_docs, args.steps, max(1, args.warmup), threads)
    sizes = [main]
    if not args.quick_reference:
        sizes.insert(0, cpu_reference_measure(10_000, max(20, args.steps), 3, threads))
        sizes.append(cpu_reference_measure(1_000_000, 3, 1, threads))
    best_t = min(main["per_threads"], key=lambda k: main["per_threads"][k]["median_s"])
    per_query = main["per_threads"][best_t]["median_s"]
    largest = sizes[-1]
    lt = min(largest["per_threads"], key=lambda k: largest["per_threads"][k]["median_s"])
    per_doc = largest["per_threads"][lt]["median_s"] / largest["docs"]
    value = 1.0 / (per_doc * args.docs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_query * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": workload_text(args.docs, max(1, args.gpus), args.batch), "docs": args.docs, "batch": args.batch, "topn": TOPN,
                   "prf": args.prf,
                   "sample": "the reference scores one query at a time (no batching, webui.py:345-390): a timed step = one query on a %d-doc index of the same generator family (ms_per_step); "
                             "value = O(N) extrapolation of the %d-doc measurement to %d docs"
                             % (main["docs"], largest["docs"], args.docs)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(lt), "kind": largest["kind"],
                         "host_cores_available": cores, "cpu_model": cpu_model(),
                         "sample": "%s path (%s), single queries, top-%d, PRF on: median s/query at %s docs = %s (BLAS threads %s); "
                                   "the reference's O(N) Python loops are single-threaded, only the two sgemv calls use BLAS threads"
                                   % ("verbatim reference functions" if largest["kind"] == "reference" else "oracle/port.py faithful=True",
                                      "webui.py:345-390", TOPN, "/".join(str(s["docs"]) for s in sizes),
                                      "/".join("%.3f" % s["per_threads"][min(s["per_threads"], key=lambda k: s["per_threads"][k]["median_s"])]["median_s"]
                                               for s in sizes), lt)},
        "measured": sizes,
        "extrapolated": {"to_docs": args.docs, "from_docs": largest["docs"], "s_per_query": per_doc * args.docs,
                         "assumption": "linear in N (measured: s/query/doc at each size in `measured`)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def verify_against_oracle(eng, rows, sh_host, idf_h, df_h, avgdl, n_docs, emb, texts, results, k_check):
    """--verify: K queries of a timed batch re-checked against the oracle port (oracle/port.py: webui.py:345-390 restated)
    on a HOST COPY OF THE BENCHMARKED INDEX itself - ids in order (swaps only inside score ties within tolerance),
    scores within 1e-5.  Test infrastructure used as the checker, outside every timed region."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scale_util as SU
    from gpu_util import assert_same_or_filter_unstable
    post_ptr, post_doc, doc_len = sh_host
    view = SU.TorchIndexView(n_docs, rows, post_ptr, post_doc, doc_len, idf_h, df_h, avgdl, emb)
    P = SU.StoredRowOracle(view)
    pick = list(range(0, len(texts), max(1, len(texts) // k_check)))[:k_check]
    want = SU.oracle_results(P, [texts[j] for j in pick], TOPN, workers=min(len(pick), 8))
    ids, scores, counts, status = results[:4]
    n_ok = 0
    for w, j in zip(want, pick):
        got = SU.engine_outcome(ids, scores, counts, status, j)
        assert_same_or_filter_unstable(got, w, lambda t=texts[j]: P.find_sorted_arrays(t), 1e-6, TOPN, ("verify", j, texts[j]))
        n_ok += 1
    return n_ok


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
    ap.add_argument("--quick-reference", action="store_true", help="--impl reference: only the --cpu-sample-docs size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra operating points (batch 1 / 64, configs[1])")
    ap.add_argument("--sweep", action="store_true", help="also time batch sizes 1..4096 (extra key batch_sweep): SIMT -> mma.sync -> tcgen05 crossover")
    ap.add_argument("--verify", type=int, default=4, help="queries of a timed batch re-checked against the oracle on a host copy of "
                                                          "the benchmarked index (0 = off; 1 GPU only)")
    ap.add_argument("--dump-results", default="", help="write the ids / scores / counts of every timed step to this .npz "
                                                       "(rank 0): lets runs at different --gpus be diffed query by query")
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
    def stage(n_docs, max_batch):
        lo, hi = shard.shard_bounds(n_docs, world, rank)
        eng = E.SearchEngine(device=local_rank, max_batch=max_batch)
        rows = eng.rows_tensor(hi - lo)
        sh = synth_torch.generate_shard(lo, hi, rows, vocab=VOCAB, seed=SEED)
        idf, avgdl, df = synth_torch.global_stats(sh, n_docs)
        eng.load_bm25(sh.post_ptr, sh.post_doc, None, idf, sh.doc_len, avgdl)
        eng.set_shard(lo, n_docs)
        return eng, rows, sh, idf, avgdl, df, lo, hi

    t_build = time.perf_counter()
    eng, rows, sh, idf, avgdl, df, lo, hi = stage(args.docs, min(256, max(args.batch, 256) if args.sweep else args.batch))
    nnz_local = int(sh.post_doc.numel())
    E_host = synth_torch.embedding_table(VOCAB, SEED, dev).cpu().numpy()
    df_host = df.cpu().numpy()
    want_verify = args.verify > 0 and world == 1
    host_copy = None
    if want_verify:
        try:
            with open("/proc/meminfo") as f:
                avail_kb = [int(ln.split()[1]) for ln in f if ln.startswith("MemAvailable")][0]
            need = (hi - lo) * 1200 * 2.2 + nnz_local * 16
            if avail_kb * 1024 > need:
                host_copy = (rows.cpu().numpy(), (sh.post_ptr.cpu().numpy(), sh.post_doc.cpu().numpy(), sh.doc_len.cpu().numpy()),
                             idf.cpu().numpy())
            else:
                host_copy = "skipped: %.0f GB of host memory needed, %.0f GB available" % (need / 1e9, avail_kb * 1024 / 1e9)
        except Exception as exc:   # noqa: BLE001
            host_copy = "skipped: %r" % (exc,)
    del sh, rows
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    n_pool = max(64, min(args.batch, 256) * 4)
    texts, parsed = synth_torch.make_queries(df_host, E_host, n_pool, seed=7)
    pool = [E.Query(*p) for p in parsed]
    eng.use_torch_stream()                               # CUDA events below see the engine's kernels
    S = shard.ShardedSearch([eng], args.docs) if world > 1 else None
    mode = E.PRF_STORED_ROWS if args.prf == "stored_rows" else E.PRF_STORED_ROWS_FULL

    def make_search(engine, sharded):
        # one GPU: the C-ABI call ais_search; several: the staged calls with NCCL between them
        return (lambda qs: sharded.search_raw(qs, TOPN, mode)) if sharded is not None else (lambda qs: engine.search_raw(qs, TOPN, mode))

    search = make_search(eng, S)

    def batch_at(step, b):
        return [pool[(step * b + j) % n_pool] for j in range(b)]

    def run_steps(n_steps, b, first_step=0, fn=None, keep=False):
        """device time (CUDA events on the engine's stream = torch's current stream), max over ranks"""
        fn = fn or search
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        done = []
        for s in range(n_steps):
            qs = batch_at(first_step + s, b)
            ids, scores, counts, status, _ = fn(qs)          # host query buffers in, host result arrays out
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
        out = (float(t[0]), float(t[1]), h2d // max(n_steps, 1), d2h // max(n_steps, 1), n_results, checksum)
        return out + ((done,) if keep else ())

    # ---- warm-up, then the timed region ----------------------------------------------------------------
    b = args.batch
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    run_steps(args.warmup, b, 0)
    eng.set_profiling(True)
    eng.reset_stats()
    sampler.mark()
    dev_ms, wall_ms, h2d, d2h, n_results, checksum, done = run_steps(args.steps, b, args.warmup, keep=True)
    sampler.mark_end()
    st = eng.stats()
    eng.set_profiling(False)
    if args.dump_results and rank == 0:
        import numpy as np
        np.savez_compressed(args.dump_results, ids=np.stack([d[1] for d in done]), scores=np.stack([d[2] for d in done]),
                            counts=np.stack([d[3] for d in done]), status=np.stack([d[4] for d in done]))

    n_queries = args.steps * b
    value = n_queries / (dev_ms * 1e-3)
    e2e = n_queries / (wall_ms * 1e-3)
    peak, peak_tf, peak_src = load_peaks()
    n_loc = hi - lo
    scan_ms = st["scan_ms_total"] / max(1, st["scan_launches"])
    scan_bytes = n_loc * 1200                           # every stored fp32 row read exactly once per launch
    achieved = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    # whole-step figure: algorithmic bytes of a step (2 passes over the rows + the posting ids of the batch)
    post_bytes = 0
    for s in range(args.steps):
        for q in batch_at(args.warmup + s, b):
            post_bytes += int(sum(4 * df_host[t] for t in q.term_ids)) // world
    post_bytes_step = post_bytes / args.steps
    # pass 2: the reference's collapsed re-query needs column 0 only (4 B per doc); the dense variant a second full pass
    requery_bytes = n_loc * 4 if st["column_scan_launches"] > 0 else scan_bytes
    step_bytes = scan_bytes + requery_bytes + post_bytes_step
    step_gbs = step_bytes / (dev_ms / args.steps * 1e-3) / 1e9

    kernel_name, traffic_key, per_pass = scan_kernel_for(b)
    if traffic_key == "scan_pair64" and st.get("pair_scan_launches", 0) == 0:
        # the engine fell back to the single-CTA kernel (clusters of two CTAs cannot be resident on this device)
        kernel_name, traffic_key = "scan_tc_kernel<64> (tcgen05 kind::tf32 3xTF32, 64 queries per pass)", "scan_tc64"
    traffic, traffic_src = load_traffic(traffic_key, n_loc)
    passes_per_step = st["scan_launches"] / args.steps
    # per kernel class (CUDA-event brackets inside the engine): ms per step, share, algorithmic bytes, fraction of peak
    step_ms = dev_ms / args.steps
    alg = {"scan": scan_bytes * passes_per_step,                # rows read once per pass (64 queries share a pass)
           "bm25_score": post_bytes_step,                       # 4 B per posting id of the batch's terms
           "bm25_slices": 0.0, "combine": 0.0, "select": 0.0, "requery": requery_bytes, "tail": 0.0, "witness": 0.0}
    model = {"combine": "reads 4 B dot score per doc-query + ~9 B per BM25 record; writes 8 B per 256 docs (tile maxima)",
             "select": "reads 8 B per 256 docs per query (tile maxima) + recomputes the combined scores of the tiles that reach the threshold",
             "bm25_slices": "one binary search per (query term, 256-doc tile)",
             "requery": "PRF seeds, centroid; the collapsed re-query reads column 0 of the rows (4 B per doc, once per index)",
             "tail": "merge + filter_searched_result on <= 1034 candidates per query, result copy to the host",
             "witness": "near-tie witness pass for ambiguous filter outcomes (12 B per doc of such a query)"}
    kernels = []
    for name, rec in st["kernels"].items():
        ms = rec["ms"] / args.steps
        if rec["brackets"] == 0:
            continue
        k = {"name": name, "ms_per_step": ms, "share_of_step": ms / step_ms, "algorithmic_bytes_per_step": alg.get(name, 0.0)}
        if alg.get(name, 0.0) > 0 and ms > 0:
            k["achieved_gbs"] = alg[name] / (ms * 1e-3) / 1e9
            k["frac_of_hbm_peak"] = k["achieved_gbs"] / peak
        if name in model:
            k["traffic_model"] = model[name]
        kernels.append(k)
    kernels.sort(key=lambda k: -k["ms_per_step"])
    # the dense scan on the tensor pipe: 3 TF32 products per element (hi*hi + hi*lo + lo*hi), K padded to 304
    tensor = None
    if per_pass >= 5 and scan_ms > 0:
        flops = 3 * 2.0 * n_loc * 304 * per_pass
        tf = flops / (scan_ms * 1e-3) / 1e12
        tensor = {"achieved_tflops_tf32": tf, "peak_tflops_tf32": peak_tf / 2.0, "frac": tf / (peak_tf / 2.0),
                  "note": "kind::tf32 runs at half the measured bf16 rate; 3xTF32 = three tensor products per fp32-accurate product; "
                          "at 64 queries per pass the HBM stream (12 GB per pass) still bounds the kernel"}

    # ---- parity re-check of the timed batch against the oracle (outside the timed region) -------------------
    parity = None
    if want_verify and rank == 0:
        if isinstance(host_copy, str) or host_copy is None:
            parity = {"checked": 0, "note": host_copy}
        else:
            t_v = time.perf_counter()
            try:
                qs0, ids0, scores0, counts0, status0 = done[0]
                texts0 = [texts[(args.warmup * b + j) % n_pool] for j in range(len(qs0))]
                n_ok = verify_against_oracle(eng, host_copy[0], host_copy[1], host_copy[2], df_host, avgdl, args.docs, E_host,
                                             texts0, (ids0, scores0, counts0, status0), min(args.verify, len(qs0)))
                parity = {"checked": n_ok, "of_batch": len(qs0), "docs": args.docs, "against": "oracle/port.py on a host copy of the benchmarked index",
                          "criterion": "ids in order (swaps only inside score ties), scores within 1e-5 relative", "ok": True,
                          "seconds": time.perf_counter() - t_v}
            except AssertionError as exc:
                parity = {"checked": 0, "ok": False, "error": str(exc)[:500]}
            except Exception as exc:   # noqa: BLE001
                parity = {"checked": 0, "ok": None, "error": "verification could not run: %r" % (exc,)}
        host_copy = None

    # ---- where a sharded step's device time goes (events at the stage boundaries, a separate untimed run) and what part
    # of a step does not shrink with the shard (the same batch over an index of 8 192 docs per GPU)
    stage_ms = None
    if S is not None:
        S.trace = True
        run_steps(max(2, args.steps // 2), b, args.warmup)
        S.trace = False
        stage_ms = {k: v / max(1, S.traced_steps) for k, v in S.stage_ms.items()}
    fixed = None
    if not args.no_modes and world == 1:             # N > 1: stage_ms_per_step above shows the same thing per stage
        try:
            n_small = 8192 * world
            e0 = stage(n_small, min(256, b))[0]
            e0.use_torch_stream()
            S0 = shard.ShardedSearch([e0], n_small) if world > 1 else None
            f0 = make_search(e0, S0)
            run_steps(3, b, 0, fn=f0)
            ms0 = run_steps(args.steps, b, 3, fn=f0)[0]
            e0.set_profiling(True); e0.reset_stats()
            run_steps(args.steps, b, 3, fn=f0)
            s0 = e0.stats(); e0.set_profiling(False)
            fixed = {"ms_per_step": ms0 / args.steps, "docs": n_small,
                     "kernels_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in s0["kernels"].items() if v["brackets"]},
                     "launches_per_step": s0["kernel_launches"] / args.steps,
                     "what": "the same batches over an index of 8192 docs per GPU: launches, collectives, merges, host work - "
                             "everything that does not shrink with the shard"}
            e0.close()
        except Exception as exc:   # noqa: BLE001
            fixed = {"error": repr(exc)}

    sweep = None
    other_configs = None
    if args.sweep or not args.no_modes:
        # other operating points of the same engine: batch 1 = single-query latency mode (both scans at the HBM
        # roofline), batch 64 = one tensor-core pass; --sweep adds the rest up to 4096 queries per call
        sweep = {}
        for bb in ((1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096) if args.sweep else (1, 64)):
            if bb > eng.params.max_batch and not args.sweep:
                break
            run_steps(2, bb, 0)
            eng.set_profiling(True); eng.reset_stats()
            k = max(2, (args.steps // 2) if bb <= 256 else 2)
            ms, wms, _, _, _, _ = run_steps(k, bb, 2)
            s2 = eng.stats(); eng.set_profiling(False)
            sm = s2["scan_ms_total"] / max(1, s2["scan_launches"])
            sweep[str(bb)] = {"qps": k * bb / (ms * 1e-3), "e2e_qps": k * bb / (wms * 1e-3), "ms_per_step": ms / k,
                              "scan_kernel": scan_kernel_for(bb)[0], "scan_ms": sm,
                              "scan_gbs": scan_bytes / (sm * 1e-3) / 1e9 if sm > 0 else None,
                              "scan_frac_of_peak": scan_bytes / (sm * 1e-3) / 1e9 / peak if sm > 0 else None,
                              "scan_share": s2["scan_ms_total"] / ms,
                              "whole_step_frac_of_peak": (scan_bytes * (-(-min(bb, 256) // scan_kernel_for(bb)[2])) * (-(-bb // 256)) +
                                                          (n_loc * 4 if s2["column_scan_launches"] > 0 else scan_bytes))
                                                         / (ms / k * 1e-3) / 1e9 / peak}
    if not args.no_modes and world == 1 and args.docs >= 2_000_000:
        # BASELINE configs[1]: 1 M docs, single weighted queries with +required / -exclude, top-100, one B200
        try:
            e1, r1, s1, i1, a1, d1, lo1, hi1 = stage(1_000_000, 1)
            del r1, s1
            e1.use_torch_stream()
            t1, p1 = synth_torch.make_queries(d1.cpu().numpy(), E_host, 64, seed=11)
            pool1 = [E.Query(*p) for p in p1]
            f1 = lambda qs: e1.search_raw(qs, TOPN, mode)
            saved = pool, n_pool
            pool, n_pool = pool1, 64
            run_steps(10, 1, 0, fn=f1)
            e1.set_profiling(True); e1.reset_stats()
            ms, wms, _, _, _, _ = run_steps(100, 1, 10, fn=f1)
            s1s = e1.stats(); e1.set_profiling(False)
            pool, n_pool = saved
            sm = s1s["scan_ms_total"] / max(1, s1s["scan_launches"])
            other_configs = {"configs[1]: 1M docs, single weighted query (+required/-exclude), top-100, 1xB200": {
                "ms_per_query": ms / 100, "qps": 100 / (ms * 1e-3), "e2e_qps": 100 / (wms * 1e-3), "scan_ms": sm,
                "scan_frac_of_peak": 1_000_000 * 1200 / (sm * 1e-3) / 1e9 / peak if sm > 0 else None,
                "whole_query_frac_of_peak": (1_000_000 * 1204) / (ms / 100 * 1e-3) / 1e9 / peak,
                "kernel_launches_per_query": s1s["kernel_launches"] / 100}}
            e1.close()
        except Exception as exc:   # noqa: BLE001
            other_configs = {"configs[1]": "failed: %r" % (exc,)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = cpu_reference_measure(args.cpu_sample_docs, args.cpu_queries, 1, [1])
        per_q = m["per_threads"]["1"]["median_s"]
        cpu = {"value": 1.0 / (per_q * args.docs / m["docs"]), "unit": UNIT, "cores": 1, "kind": m["kind"],
               "host_cores_available": os.cpu_count(), "cpu_model": cpu_model(),
               "sample": "%s on a %d-doc index of the same generator family, %d single queries: median %.3f s/query, scaled "
                         "linearly (O(N)) to %d docs; the reference's Python loops are single-threaded; `bench.py --impl reference` "
                         "measures 10^4 / 10^5 / 10^6 docs with 1 and all BLAS threads"
                         % ("verbatim reference functions (webui.py:345-390)" if m["kind"] == "reference" else "oracle/port.py faithful=True",
                            m["docs"], args.cpu_queries, per_q, args.docs)}

    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 dot (3xTF32 on tcgen05 from 9 queries per pass) / f64 BM25+combine", "data": "synthetic",
            "config": {"workload": workload_text(args.docs, world, b),
                       "docs": args.docs, "batch": b, "topn": TOPN, "prf": args.prf, "postings_rank0": nnz_local,
                       "requery": "column scan (single non-zero component, SURVEY.md A.5)" if st["column_scan_launches"] > 0 else "dense scan",
                       "parallelism": "doc-shard x%d" % world,
                       "l2": "inputs larger than L2 (%.1f GB of rows per GPU re-read every pass)" % (scan_bytes / 1e9)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "host wall clock around the public API call (ctypes -> C ABI) with host query buffers and host result arrays"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "%s, %d launches" % (kernel_name, st["scan_launches"]),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "traffic_frac": (traffic / (scan_ms * 1e-3) / 1e9 / peak) if traffic and scan_ms > 0 else None,
                         "peak_source": peak_src, "bytes_per_launch": scan_bytes, "ms_per_launch": scan_ms,
                         "scan_share_of_step": st["scan_ms_total"] / dev_ms,
                         "tensor": tensor,
                         "kernels": kernels,
                         "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / peak,
                                        "note": "rows counted once per batch (SURVEY 8d); a 256-query batch is read in %d passes of "
                                                "%d queries, so the dense arithmetic (3xTF32), not this figure, bounds the step"
                                                % (int(round(passes_per_step)), per_pass)}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity_checked": parity,
            "results_per_step": n_results / args.steps,
            "results_checksum": checksum,      # same queries -> same value for every --gpus N (doc ids of all results)
            "fullsort_fallbacks": int(st["fullsort_fallbacks"]) + (S.fullsort_fallbacks if S is not None else 0),
            "bound_passes": int(st["bound_passes"]), "tiles_per_seg": int(st["tiles_per_seg"]),
            "device_bytes": int(st["bytes_device"]),
            "index_build_s": t_build,
            "fixed_ms_per_step": fixed,
        }
        if stage_ms:
            line["stage_ms_per_step"] = stage_ms
        if sweep:
            line["batch_sweep" if args.sweep else "other_batch_sizes"] = sweep
        if other_configs:
            line["other_configs"] = other_configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
