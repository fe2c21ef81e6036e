"""TEST INFRASTRUCTURE: a numpy stage engine with the interface shard.ShardedSearch drives (the staged
calls of include/ais_b200.h on CPU tensors), built from the oracle's arithmetic.  It lets the N>1 host
logic of shard.py (collective sequence, gather layouts, ambiguity protocol) run under gloo without a GPU.
It is NOT a product path: the product's engines are ais_b200.engine.SearchEngine (CUDA only)."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import port

KEY_EMPTY = np.uint64(0)
ID_EMPTY = np.int64(0x7FFFFFFFFFFFFFFF)
SIGN = np.uint64(0x8000000000000000)


def dkey(x):
    b = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).view(np.uint64).copy()
    b[b == SIGN] = 0
    neg = (b & SIGN) != 0
    return np.where(neg, ~b, b | SIGN)


def dkey_inv(k):
    k = np.asarray(k, dtype=np.uint64)
    b = np.where((k & SIGN) != 0, k & ~SIGN, ~k)
    return b.view(np.float64)


def top_sorted(keys, ids, k):
    """best-first (key desc, id asc) top-k of candidate arrays, KEY_EMPTY skipped, padded to k."""
    live = keys != KEY_EMPTY
    keys, ids = keys[live], ids[live]
    order = np.lexsort((ids, np.iinfo(np.uint64).max - keys))[:k]
    ok = np.full(k, KEY_EMPTY, dtype=np.uint64)
    oi = np.full(k, ID_EMPTY, dtype=np.int64)
    ok[:len(order)] = keys[order]
    oi[:len(order)] = ids[order]
    return ok, oi, len(order)


class FakeStageEngine:
    def __init__(self, idx, lo, hi, max_batch=4, thresh=1e-6):
        self.idx, self.lo, self.hi = idx, lo, hi
        self.n_docs = hi - lo
        self.n_total = idx.n_docs
        self.params = SimpleNamespace(prf_depth=10, max_batch=max_batch)
        self.thresh = thresh
        self.torch_device = torch.device("cpu")
        self.P = port.OraclePort(idx)              # bm25 / consts of the whole index; sliced to the shard below
        self.rows = idx.rows[lo:hi]

    def max_select_k(self):
        return 1024

    # -- helpers
    @staticmethod
    def _np(t):
        return t.numpy() if isinstance(t, torch.Tensor) else t

    def _keys_view(self, t):
        return self._np(t).view(np.uint64)

    def stage_score(self, queries, maxes):
        self.nq = len(queries)
        self.sim = [np.dot(self.rows, q.vec) for q in queries]
        self.bm25 = [self.P.bm25_scores(dict(zip(q.term_ids.tolist(), q.weights.tolist())))[self.lo:self.hi] for q in queries]
        m = self._np(maxes)
        for i in range(self.nq):
            m[i, 0] = self.bm25[i].max() if self.n_docs else -np.inf
            m[i, 1] = float(self.sim[i].max()) if self.n_docs else -np.inf
        self.status = np.zeros(self.nq, dtype=np.int32)

    def stage_combine(self, nq, maxes, k, keys, ids):
        m = self._np(maxes)
        kk, ii = self._keys_view(keys), self._np(ids)
        self.fin = []
        c = self.P.consts
        for i in range(nq):
            s, b = self.sim[i], self.bm25[i]
            maxs, maxb = np.float32(m[i, 1]), m[i, 0]
            if maxs > 0:
                s = s / maxs
            if maxb > 0:
                b = b / maxb
            f = c["BM25_WEIGHT"] * b + c["DOC2VEC_WEIGHT"] * s
            self.fin.append(f)
            kk[i], ii[i], _ = top_sorted(dkey(f), np.arange(self.lo, self.hi, dtype=np.int64), k)

    def stage_top(self, nq, n_lists, k, keys, ids, want_host, rows=None):
        kk, ii = self._keys_view(keys).reshape(n_lists, nq, k), self._np(ids).reshape(n_lists, nq, k)
        d = self.params.prf_depth
        self.top_ids = np.zeros((nq, d), dtype=np.int64)
        self.top_scores = np.zeros((nq, d))
        for i in range(nq):
            tk, ti, _ = top_sorted(kk[:, i].ravel(), ii[:, i].ravel(), d)
            self.top_ids[i], self.top_scores[i] = ti, dkey_inv(tk)
            if rows is not None:
                r = self._np(rows)
                for t in range(d):
                    loc = ti[t] - self.lo
                    r[i, t] = self.rows[loc] if 0 <= loc < self.n_docs else 0.0
        return (self.top_ids.copy(), self.top_scores.copy()) if want_host else (None, None)

    def stage_set_status(self, status):
        self.status = np.asarray(status, dtype=np.int32).copy()

    def stage_requery(self, nq, q2, rows, prf_mode, k, max_r, keys, ids):
        c = self.P.consts
        self.R = []
        mr = self._np(max_r)
        for i in range(nq):
            if q2 is not None:
                qv = q2[i]
            else:
                w = self.top_scores[i]
                if not np.isfinite(w).all():
                    self.status[i] = 1
                    qv = np.zeros(300, np.float32)
                elif np.add.reduce(w) == 0:
                    self.status[i] = 2
                    qv = np.zeros(300, np.float32)
                else:
                    vecs = [[(j, v) for j, v in enumerate(self._np(rows)[i, t])] for t in range(len(w))]
                    from oracle.gensim_stub import dense_query
                    qv = dense_query(port.OraclePort.prf_query(vecs, w.tolist()), 300)
            rer = np.dot(self.rows, qv.astype(np.float32))
            R = c["ORIGINAL_SCORE_WEIGHT"] * self.fin[i] + c["RERANKED_SCORE_WEIGHT"] * rer
            self.R.append(R)
            mr[i] = R.max() if self.n_docs else -np.inf
        self.stage_requery_select(nq, k, keys, ids)

    def stage_requery_select(self, nq, k, keys, ids):
        kk, ii = self._keys_view(keys), self._np(ids)
        gid = np.arange(self.lo, self.hi, dtype=np.int64)
        for i in range(nq):
            keep = ~np.isin(gid, self.top_ids[i])
            kk[i], ii[i], _ = top_sorted(dkey(self.R[i])[keep], gid[keep], k)

    def _tail(self, vals_rest, ids_rest, top_ids, depth, topn, complete, witness):
        s = np.concatenate([np.ones(depth), vals_rest])
        ids = np.concatenate([top_ids[:depth], ids_rest])
        with np.errstate(invalid="ignore"):
            diff = s[:-1] - s[1:]
        diff = np.where(diff == 0, np.inf, diff)
        found = np.where(diff < self.thresh)[0]
        lim = min(len(s), topn)
        pos = np.where(~(s[:lim] > 0))[0]
        npos = pos[0] if len(pos) else lim
        amb = False
        if len(found) >= 2:
            t = found[1]
        elif len(found) == 1:
            if complete:
                t = found[0]
            else:
                t = lim
                amb = found[0] < npos and not witness
        else:
            t = len(s) if complete else lim
        cnt = int(min(t, npos, lim))
        mx = s[0] if len(s) else 1.0
        return ids[:cnt], s[:cnt] / mx, cnt, amb

    def stage_finish(self, nq, n_lists, k, keys, ids, max_r, topn, witness=None):
        kk, ii = self._keys_view(keys).reshape(n_lists, nq, k), self._np(ids).reshape(n_lists, nq, k)
        depth = self.params.prf_depth if max_r is not None else 0
        out_ids = np.zeros((nq, topn), dtype=np.int64)
        out_scores = np.zeros((nq, topn))
        counts = np.zeros(nq, dtype=np.int32)
        amb = np.zeros(nq, dtype=np.int32)
        last = np.zeros(nq, dtype=np.uint64)
        wit = None if witness is None else self._np(witness)
        for i in range(nq):
            # engine.cu do_finish: merged list up to SEL_KMAX deep, cut to its certainly-exact prefix (prefix_bound_kernel)
            k_out = k if n_lists == 1 else min(1024, n_lists * k)
            rk, ri, m = top_sorted(kk[:, i].ravel(), ii[:, i].ravel(), k_out)
            if k_out > k:
                bound = kk[:, i, k - 1].max()
                m = max(min(k, m), int((rk[:m] > bound).sum()))
            vals = dkey_inv(rk[:m])
            if max_r is not None and self._np(max_r)[i] > 0:
                vals = vals / self._np(max_r)[i]
            last[i] = rk[m - 1] if m else np.uint64(0xFFFFFFFFFFFFFFFF)
            if self.status[i]:
                continue
            oi, os_, c, a = self._tail(vals, ri[:m], getattr(self, "top_ids", np.zeros((nq, 0), np.int64))[i], depth, topn,
                                       depth + m >= self.n_total, bool(wit[i]) if wit is not None else False)
            out_ids[i, :c], out_scores[i, :c], counts[i], amb[i] = oi, os_, c, a
        return out_ids, out_scores, counts, self.status.copy(), amb, last

    def _values(self, q, second_pass, max_r):
        v = self.R[q] if second_pass else self.fin[q]
        if max_r is not None and self._np(max_r)[q] > 0:
            v = v / self._np(max_r)[q]
        return v

    def stage_witness(self, amb, last_keys, second_pass, max_r, witness):
        w = self._np(witness)
        w[:] = 0
        gid = np.arange(self.lo, self.hi, dtype=np.int64)
        for q in np.nonzero(amb)[0]:
            raw = self.R[q] if second_pass else self.fin[q]
            ok = (dkey(raw) <= last_keys[q]) & np.isfinite(raw)
            if second_pass:
                ok &= ~np.isin(gid, self.top_ids[q])
            v = np.unique(self._values(q, second_pass, max_r)[ok])
            d = np.diff(v)
            w[q] = int(np.any((d != 0) & (d < self.thresh)))

    def stage_export_keys(self, query, second_pass, keys, ids):
        kk, ii = self._keys_view(keys), self._np(ids)
        gid = np.arange(self.lo, self.hi, dtype=np.int64)
        k = dkey(self.R[query] if second_pass else self.fin[query])
        if second_pass:
            seed = np.isin(gid, self.top_ids[query])
            k[seed] = KEY_EMPTY
            gid = np.where(seed, ID_EMPTY, gid)
        kk[:], ii[:] = k, gid

    def sort_capacity(self, n):
        p = 2048
        while p < n:
            p <<= 1
        return p

    def stage_sort_finish(self, query, keys, ids, n_entries, max_r, topn):
        kk, ii = self._keys_view(keys)[:n_entries], self._np(ids)[:n_entries]
        rk, ri, m = top_sorted(kk, ii, n_entries)
        vals = dkey_inv(rk[:m])
        if max_r is not None and self._np(max_r)[query] > 0:
            vals = vals / self._np(max_r)[query]
        depth = self.params.prf_depth if max_r is not None else 0
        oi, os_, c, _ = self._tail(vals, ri[:m], getattr(self, "top_ids", np.zeros((1, 0), np.int64))[query] if depth else np.zeros(0, np.int64),
                                   depth, topn, True, False)
        out_ids = np.zeros(topn, dtype=np.int64)
        out_scores = np.zeros(topn)
        out_ids[:c], out_scores[:c] = oi, os_
        return out_ids, out_scores, c, int(self.status[query])
