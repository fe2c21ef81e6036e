"""Doc-sharded search: one process per GPU, each holding a contiguous range of the documents
(SURVEY.md 8e).  Every doc is scored independently, so the data path needs no collective; the only
exchanges are the small per-query records between the engine's stages:

    stage_score    -> all-reduce(MAX)  [nq, 2] float64   {max bm25, max dot}          (webui.py:377-380)
    stage_combine  -> all-gather       [nq, depth] (key, id) local PRF seeds          (webui.py:191-195)
    stage_top      -> all-reduce(SUM)  [nq, depth, 300] float32 stored seed rows      (device PRF mode)
    stage_requery  -> all-reduce(MAX)  [nq] float64 max R ; all-gather [nq, k] (key, id)
    stage_finish   on every rank (identical inputs -> identical results)

A process may hold several engines (``engines=[...]``): their records are reduced locally first.
That is how the N>1 logic is exercised on one GPU and, with numpy stage engines, on CPU over gloo.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import PRF_CALLBACK, PRF_OFF, Query

AIS_Q_NAN_WEIGHTS, AIS_Q_ZERO_WEIGHT_SUM, AIS_Q_CALLBACK_FAILED = 1, 2, 4


def shard_bounds(n_docs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous doc-id ranges of ceil(N / world) docs (global id = base + local id)."""
    per = (n_docs + world - 1) // world
    lo = min(n_docs, rank * per)
    return lo, min(n_docs, lo + per)


class _Dist:
    """The three collectives the search needs, over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def all_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.on and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def all_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.on and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """[...] -> [world, ...]"""
        if not (self.on and self.world > 1):
            return t.unsqueeze(0)
        flat = t.contiguous().view(-1)
        out = torch.empty((self.world * flat.numel(),), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)       # flat: the same call shape for NCCL and gloo
        return out.view((self.world,) + tuple(t.shape))

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.on and self.world > 1:
            self.dist.broadcast(t, src=src, group=self.group)
        return t


def _np_pairwise_sum(w: np.ndarray) -> float:
    return float(np.add.reduce(np.asarray(w, dtype=np.float64)))


class ShardedSearch:
    def __init__(self, engines: Sequence, n_total: int, group=None):
        """engines: this process' stage engines (SearchEngine or anything with the same stage_* methods,
        a ``params.prf_depth``, ``max_select_k()`` and a ``torch_device``)."""
        self.engines = list(engines)
        for e in self.engines:
            if hasattr(e, "use_torch_stream"):
                e.use_torch_stream()          # engine kernels and NCCL collectives order on torch's stream
        self.n_total = int(n_total)
        self.comm = _Dist(group)
        self.depth = int(self.engines[0].params.prf_depth)
        self.kmax = int(self.engines[0].max_select_k())
        self.fullsort_fallbacks = 0
        self.n_shards = self.comm.world * len(self.engines)
        # stage trace (bench.py `stage_ms_per_step`): CUDA events on torch's stream at the stage boundaries, read after
        # the step's own final synchronisation - shows where a sharded step's device time goes, gaps included
        self.trace = False
        self.stage_ms: dict = {}
        self.traced_steps = 0
        self._marks: list = []

    def _mark(self, name: str) -> None:
        if self.trace and torch.cuda.is_available():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def _trace_flush(self) -> None:
        if not self._marks:
            return
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
            self.stage_ms[name] = self.stage_ms.get(name, 0.0) + a.elapsed_time(b)
        self._marks = []
        self.traced_steps += 1

    def _select_depth(self, need: int, cap: int) -> int:
        """Candidates a select stage asks for - the engine's own policy (engine.cu select_depth), from GLOBAL quantities so
        that every rank gathers lists of the same length: a longer exact prefix is nearly free and lets the filter see its
        second near-tie (webui.py:66-77) without the witness pass; kept below half a shard's segment count."""
        import os
        per = -(-self.n_total // max(1, self.n_shards))
        n_tiles = -(-per // 256)
        tps = max(1, -(-n_tiles // 2048))
        deep = (-(-n_tiles // tps)) // 2
        # the merged list is exact down to the shards' largest last key (select.cuh prefix_bound_kernel): n_shards
        # lists of kmax / n_shards (+ slack for an uneven split) give the same ~kmax-deep exact prefix as one engine
        if self.n_shards > 1 and os.environ.get("AIS_SHARD_CUT", "1") != "0":
            deep = min(deep, -(-self.kmax // self.n_shards) + 128)
        elif self.n_shards == 1:
            deep = min(deep, 576)                            # the engine's own default (engine.cu sel_deep)
        env = os.environ.get("AIS_SELECT_DEPTH")
        if env is not None and int(env) >= 0:
            deep = min(deep, int(env))
        return max(1, min(cap, max(need, deep)))

    # ---- helpers ----------------------------------------------------------------------------------
    def _dev(self, eng):
        return getattr(eng, "torch_device", torch.device("cpu"))

    def _gather_lists(self, keys: List[torch.Tensor], ids: List[torch.Tensor]):
        """per-engine [nq, k] -> [L, nq, k] with L = world * engines-per-process (on engine 0's device).
        keys[j] / ids[j] are the two halves of ONE buffer (see _cand_buffers): a single all-gather moves both."""
        dev = self._dev(self.engines[0])
        packs = [k._base if k._base is not None else torch.stack([k, i]) for k, i in zip(keys, ids)]   # [2, nq, k] each
        local = packs[0].unsqueeze(0) if len(packs) == 1 else torch.stack([p.to(dev) for p in packs])   # [E, 2, nq, k]
        g = self.comm.all_gather(local)                                                                  # [W, E, 2, nq, k]
        g = g.reshape((-1,) + tuple(local.shape[1:]))                                                    # [L, 2, nq, k]
        if g.shape[0] == 1:
            return g[0, 0].unsqueeze(0), g[0, 1].unsqueeze(0)
        return g[:, 0].contiguous(), g[:, 1].contiguous()

    def _reduce_local(self, tensors: List[torch.Tensor], op) -> torch.Tensor:
        if len(tensors) == 1:
            return tensors[0]
        dev = self._dev(self.engines[0])
        acc = tensors[0].to(dev).clone()
        for t in tensors[1:]:
            acc = op(acc, t.to(dev))
        return acc

    def _cand_buffers(self, nq: int, k: int):
        packs = [torch.empty((2, nq, k), dtype=torch.int64, device=self._dev(e)) for e in self.engines]
        return [p[0] for p in packs], [p[1] for p in packs]

    # ---- one batch ---------------------------------------------------------------------------------
    def search_raw(self, queries: Sequence[Query], topn: int, prf_mode: int, infer_cb: Optional[Callable] = None):
        """-> (ids [nq, topn], scores [nq, topn], counts [nq], status [nq], callback errors); same on every rank."""
        E = self.engines
        nq = len(queries)
        depth = self.depth
        errors: list = []
        # --- pass 1: score, global maxima
        self._mark("start")
        maxes_l = [torch.empty((nq, 2), dtype=torch.float64, device=self._dev(e)) for e in E]
        for e, m in zip(E, maxes_l):
            e.stage_score(queries, m)
        self._mark("score")
        maxes = self.comm.all_max(self._reduce_local(maxes_l, torch.maximum))
        maxes_e = [maxes.to(self._dev(e)) for e in E]
        self._mark("all_max(maxes)")
        prf = prf_mode != PRF_OFF and self.n_total > depth

        # a result longer than the selector returns: every query goes through the exact full sort (_resolve_ambiguous)
        too_long = topn + 1 > self.kmax
        if not prf:
            k = 1 if too_long else self._select_depth(topn + 1, self.kmax)
            keys, ids = self._cand_buffers(nq, k)
            for e, m, kk, ii in zip(E, maxes_e, keys, ids):
                e.stage_combine(nq, m, k, kk, ii)
            self._mark("combine+select")
            gk, gi = self._gather_lists(keys, ids)
            self._mark("all_gather(lists)")
            res = self._finish(nq, k, gk, gi, None, topn, second_pass=False, all_ambiguous=too_long)
            self._mark("finish")
            self._trace_flush()
            return self._resolve_ambiguous(res, nq, topn, None, second_pass=False) + (errors,)

        # --- pass 1: every shard's best k1 = depth + k2 docs (the engine keeps the list: its pass-2 threshold starts from
        # it); the global PRF seeds are the best `depth` of the shards' first `depth` entries
        k2 = (self.kmax - depth) if too_long else self._select_depth(max(1, topn + 1 - depth), self.kmax - depth)
        k1 = depth + k2
        keys, ids = self._cand_buffers(nq, k1)
        for e, m, kk, ii in zip(E, maxes_e, keys, ids):
            e.stage_combine(nq, m, k1, kk, ii)
        self._mark("combine+select")
        gk, gi = self._gather_lists([kk[:, :depth].clone() for kk in keys], [ii[:, :depth].clone() for ii in ids])
        self._mark("all_gather(seeds)")
        host = prf_mode == PRF_CALLBACK
        rows_l = None if host else [torch.empty((nq, depth, 300), dtype=torch.float32, device=self._dev(e)) for e in E]
        top = None
        for j, e in enumerate(E):
            r = e.stage_top(nq, gk.shape[0], depth, gk.to(self._dev(e)), gi.to(self._dev(e)), host and j == 0,
                            None if host else rows_l[j])
            if j == 0:
                top = r
        self._mark("top rows")
        q2 = None
        rows_e = [None] * len(E)
        if host:
            # rank 0 runs the (possibly non-deterministic) host inference and broadcasts the outcome
            pack = torch.zeros((nq, 301), dtype=torch.float32)
            if self.comm.rank == 0:
                top_ids, top_scores = top
                for q in range(nq):
                    w = top_scores[q]
                    st = 0
                    if not np.isfinite(w).all():
                        st = AIS_Q_NAN_WEIGHTS
                    elif _np_pairwise_sum(w) == 0.0:
                        st = AIS_Q_ZERO_WEIGHT_SUM
                    else:
                        try:
                            pack[q, :300] = torch.from_numpy(np.ascontiguousarray(infer_cb(q, top_ids[q].copy(), w.copy()),
                                                                                  dtype=np.float32))
                        except BaseException as exc:  # noqa: BLE001 - surfaced by the caller
                            errors.append((q, exc))
                            st = AIS_Q_CALLBACK_FAILED
                    pack[q, 300] = float(st)
            dev0 = self._dev(E[0])
            pack = self.comm.broadcast(pack.to(dev0), 0).cpu()
            q2 = pack[:, :300].contiguous().numpy()
            status = pack[:, 300].to(torch.int32).numpy()
            if status.any():
                for e in E:
                    e.stage_set_status(status)
        else:
            rows = self.comm.all_sum(self._reduce_local(rows_l, torch.add))
            rows_e = [rows.to(self._dev(e)) for e in E]
        self._mark("all_sum(rows)")

        # --- pass 2
        k = k2
        keys, ids = self._cand_buffers(nq, k)
        maxr_l = [torch.empty((nq,), dtype=torch.float64, device=self._dev(e)) for e in E]
        for e, r, mr, kk, ii in zip(E, rows_e, maxr_l, keys, ids):
            e.stage_requery(nq, q2, r, prf_mode, k, mr, kk, ii)
        self._mark("requery+select")
        maxr = self.comm.all_max(self._reduce_local(maxr_l, torch.maximum))
        gk, gi = self._gather_lists(keys, ids)
        self._mark("all_max(maxr)+all_gather(lists)")
        res = self._finish(nq, k, gk, gi, maxr, topn, second_pass=True, all_ambiguous=too_long)
        self._mark("finish")
        self._trace_flush()
        return self._resolve_ambiguous(res, nq, topn, maxr, second_pass=True) + (errors,)

    def _finish(self, nq: int, k: int, gk, gi, maxr, topn: int, second_pass: bool, all_ambiguous: bool = False):
        """stage_finish on engine 0 (every rank holds identical inputs); an ambiguous filter outcome first
        goes through the near-tie witness pass on every shard (flags all-reduced with MAX)."""
        E = self.engines
        res = E[0].stage_finish(nq, gk.shape[0], k, gk, gi, maxr, topn)
        amb, last = res[4], res[5]
        if all_ambiguous:
            return res[:4] + (np.ones(nq, dtype=np.int32),)
        if amb.any():
            flags = []
            for e in E:
                w = torch.zeros((nq,), dtype=torch.int32, device=self._dev(e))
                e.stage_witness(amb, last, second_pass, None if maxr is None else maxr.to(self._dev(e)), w)
                flags.append(w)
            wit = self.comm.all_max(self._reduce_local(flags, torch.maximum))
            res = E[0].stage_finish(nq, gk.shape[0], k, gk, gi, maxr, topn, witness=wit)
        return res[:5]

    def _resolve_ambiguous(self, res, nq: int, topn: int, maxr, second_pass: bool):
        """Exact fallback: all-gather every shard's keys for the query and sort them (rare)."""
        out_ids, out_scores, counts, status, amb = res
        E = self.engines
        for q in range(nq):
            if not amb[q]:
                continue
            self.fullsort_fallbacks += 1
            dev0 = self._dev(E[0])
            n_loc = torch.tensor([sum(e.n_docs for e in E)], dtype=torch.int64, device=dev0)
            n_max = int(self.comm.all_max(n_loc.clone()).item())
            lk = torch.zeros((n_max,), dtype=torch.int64, device=dev0)                 # KEY_EMPTY
            li = torch.full((n_max,), 0x7FFFFFFFFFFFFFFF, dtype=torch.int64, device=dev0)   # ID_EMPTY
            at = 0
            for e in E:
                if e.n_docs == 0:
                    continue
                kk = torch.empty((e.n_docs,), dtype=torch.int64, device=self._dev(e))
                ii = torch.empty((e.n_docs,), dtype=torch.int64, device=self._dev(e))
                e.stage_export_keys(q, second_pass, kk, ii)
                lk[at:at + e.n_docs] = kk.to(dev0)
                li[at:at + e.n_docs] = ii.to(dev0)
                at += e.n_docs
            gk = self.comm.all_gather(lk).reshape(-1)
            gi = self.comm.all_gather(li).reshape(-1)
            n_entries = gk.numel()
            cap = E[0].sort_capacity(n_entries)
            sk = torch.zeros((cap,), dtype=torch.int64, device=dev0)
            si = torch.zeros((cap,), dtype=torch.int64, device=dev0)
            sk[:n_entries] = gk
            si[:n_entries] = gi
            ids_q, scores_q, cnt, st = E[0].stage_sort_finish(q, sk, si, n_entries, maxr, topn)
            out_ids[q, :cnt] = ids_q[:cnt]
            out_scores[q, :cnt] = scores_q[:cnt]
            counts[q] = cnt
            status[q] = st
        return out_ids, out_scores, counts, status

    def search(self, queries: Sequence[Query], topn: int, prf_mode: int, infer_cb: Optional[Callable] = None):
        from .engine import raise_for_status
        max_batch = int(self.engines[0].params.max_batch)
        out: List[List[Tuple[int, float]]] = []
        for lo in range(0, len(queries), max_batch):
            chunk = queries[lo: lo + max_batch]
            cb = (lambda qi, ids, sc, _lo=lo: infer_cb(_lo + qi, ids, sc)) if infer_cb else None
            ids, scores, counts, status, errors = self.search_raw(chunk, topn, prf_mode, cb)
            if errors:
                raise errors[0][1]
            for q in range(len(chunk)):
                raise_for_status(int(status[q]))
                c = int(counts[q])
                out.append(list(zip(ids[q, :c].tolist(), scores[q, :c].tolist())))
        return out
