"""On-device synthetic index generator for the >= 1 M-doc benchmark configurations.

Same distributional recipe as synth.py (V tags with Zipf(1) popularity, tags/doc ~ clip(round(lognormal(
ln 28, 0.45)), 3, 120) distinct tags, stored rows = s * (mean E[tags] + 0.3 * noise), NOT normalised, tf = 1),
but drawn with torch's CUDA generator chunk by chunk so that a 10 M-doc shard (12 GB of rows, 3e8 postings)
is built in seconds without touching host memory.  Chunks are seeded by their GLOBAL chunk index, so the
union of the shards is the same corpus for every world size.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch

from .synth import DEFAULT_VOCAB, DIM, zipf_popularity

CHUNK = 1 << 16


class TorchShard:
    """Everything an engine needs for docs [lo, hi): rows are written straight into the engine's store."""

    def __init__(self):
        self.n_docs = 0
        self.post_ptr = None      # int64 [V+1] device
        self.post_doc = None      # int32 [nnz] device, local doc ids ascending per term
        self.doc_len = None       # int64 [n] device
        self.df = None            # int64 [V] device (this shard)
        self.total_len = 0


def embedding_table(vocab: int, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed ^ 0xE3B)
    return torch.randn((vocab, DIM), generator=g, device=device, dtype=torch.float32)


def generate_shard(lo: int, hi: int, rows_out: torch.Tensor, vocab: int = DEFAULT_VOCAB, seed: int = 20260101,
                   mean_tags: float = 28.0, sigma: float = 0.45, min_tags: int = 3, max_tags: int = 120,
                   width: int = 192, chunk: int = CHUNK) -> TorchShard:
    """Fill rows_out [hi-lo, 300] (device) and return the shard's posting lists."""
    device = rows_out.device
    n = hi - lo
    sh = TorchShard()
    sh.n_docs = n
    cdf = torch.from_numpy(np.cumsum(zipf_popularity(vocab))).to(device=device, dtype=torch.float32)
    cdf[-1] = 1.0
    E = embedding_table(vocab, seed, device)
    keys: List[torch.Tensor] = []
    lens: List[torch.Tensor] = []
    c0 = lo // chunk
    c1 = (hi + chunk - 1) // chunk if n > 0 else c0
    for c in range(c0, c1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + c)
        m = chunk
        want = torch.exp(math.log(mean_tags) + sigma * torch.randn((m,), generator=g, device=device))
        want = want.round().clamp_(min_tags, max_tags).to(torch.int64)
        draws = torch.searchsorted(cdf, torch.rand((m, width), generator=g, device=device)).clamp_(max=vocab - 1)
        noise = torch.randn((m, DIM), generator=g, device=device)
        scale = torch.exp(0.25 * torch.randn((m, 1), generator=g, device=device))
        # first occurrence of every tag inside its row, in draw order
        srt, order = torch.sort(draws, dim=1, stable=True)
        dup_sorted = torch.zeros_like(srt, dtype=torch.bool)
        dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
        dup = torch.zeros_like(dup_sorted)
        dup.scatter_(1, order, dup_sorted)
        first = ~dup
        rank = torch.cumsum(first, dim=1)
        keep = first & (rank <= want[:, None])
        # restrict the chunk to the docs of this shard
        d0 = c * chunk
        a, b = max(lo, d0) - d0, min(hi, d0 + chunk) - d0
        keep, draws, noise, scale = keep[a:b], draws[a:b], noise[a:b], scale[a:b]
        cnt = keep.sum(dim=1)
        r_idx, c_idx = torch.nonzero(keep, as_tuple=True)                  # row-major: draw order inside a doc
        tags = draws[r_idx, c_idx]
        # stored row = s * (mean E[tags] + 0.3 * noise); the sum runs over the tag positions in draw order so that
        # it is bit-reproducible (index_add_ on CUDA uses atomics, i.e. an arbitrary summation order)
        width_out = int(cnt.max().item()) if cnt.numel() else 0
        padded = torch.zeros((b - a, max(width_out, 1)), dtype=torch.int64, device=device)
        padded[r_idx, rank[a:b][r_idx, c_idx] - 1] = tags
        acc = torch.zeros((b - a, DIM), device=device, dtype=torch.float32)
        for kpos in range(width_out):
            live = (cnt > kpos).to(torch.float32)[:, None]
            acc += E[padded[:, kpos]] * live
        acc /= cnt.clamp(min=1).to(torch.float32)[:, None]
        local0 = d0 + a - lo
        rows_out[local0: local0 + (b - a)] = scale * (acc + 0.3 * noise)
        keys.append((tags.to(torch.int64) << 32) | (r_idx + local0))
        lens.append(cnt)
    if keys:
        key = torch.sort(torch.cat(keys))[0]                              # (term, local doc) ascending
        sh.post_doc = (key & 0xFFFFFFFF).to(torch.int32)
        terms = key >> 32
        sh.df = torch.bincount(terms, minlength=vocab)
        sh.doc_len = torch.cat(lens)
    else:
        sh.post_doc = torch.zeros((0,), dtype=torch.int32, device=device)
        sh.df = torch.zeros((vocab,), dtype=torch.int64, device=device)
        sh.doc_len = torch.zeros((0,), dtype=torch.int64, device=device)
    sh.post_ptr = torch.zeros((vocab + 1,), dtype=torch.int64, device=device)
    sh.post_ptr[1:] = torch.cumsum(sh.df, 0)
    sh.total_len = int(sh.doc_len.sum().item())
    return sh


def global_stats(sh: TorchShard, n_total: int) -> Tuple[torch.Tensor, float, torch.Tensor]:
    """IDF / avgdl over ALL shards (genmodel.py:76-82): all-reduce df and the length sum when distributed."""
    import torch.distributed as dist
    df = sh.df.clone()
    tot = torch.tensor([sh.total_len], dtype=torch.int64, device=df.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(df)
        dist.all_reduce(tot)
    dfd = df.to(torch.float64)
    idf = torch.log(1 + (n_total - dfd + 0.5) / (dfd + 0.5))
    idf = torch.where(df > 0, idf, torch.zeros_like(idf))
    avgdl = float(tot.item()) / float(n_total)
    return idf, avgdl, df


def make_queries(df: np.ndarray, E: np.ndarray, n_queries: int, seed: int = 7, max_terms: int = 6,
                 p_required: float = 0.3, p_exclude: float = 0.3):
    """Weighted tag queries: (texts, [(vec fp32[300], term ids, weights)]).  The query vector
    follows webui.py:104-115 with E[tag] standing in for infer_vector([tag])."""
    rng = np.random.default_rng(seed)
    vocab = len(df)
    pop = zipf_popularity(vocab)
    present = np.nonzero(df)[0]
    w = np.sqrt(pop[present])
    w /= w.sum()
    top_pop = present[np.argsort(-df[present], kind="stable")[:200]]
    texts, parsed = [], []
    for _ in range(n_queries):
        t = int(rng.integers(1, max_terms + 1))
        tags = [int(x) for x in rng.choice(present, size=t, replace=False, p=w)]
        vec_w = [int(rng.integers(1, 6)) for _ in tags]
        bm_w = [float(x) for x in vec_w]
        used = set(tags)
        if rng.random() < p_required:
            cand = [int(c) for c in top_pop[:50] if int(c) not in used]
            tg = cand[int(rng.integers(0, len(cand)))]
            k = int(rng.integers(1, 4))
            tags.append(tg); vec_w.append(k); bm_w.append(1000.0 + k); used.add(tg)
        if rng.random() < p_exclude:
            cand = [int(c) for c in top_pop if int(c) not in used]
            tg = cand[int(rng.integers(0, len(cand)))]
            k = int(rng.integers(1, 4))
            tags.append(tg); vec_w.append(-k); bm_w.append(-float(k))
        acc = np.zeros(DIM)
        for tg, wt in zip(tags, vec_w):
            v = E[tg].astype(np.float64)
            acc += wt * (v / np.linalg.norm(v))
        tot = sum(vec_w) or 1
        acc /= tot
        nrm = np.linalg.norm(acc)
        acc /= nrm if (nrm > 0 and np.isfinite(nrm)) else 1.0
        texts.append(" ".join("t%d:%s%d" % (tg, "+" if bw > 1000 else "", wt) for tg, wt, bw in zip(tags, vec_w, bm_w)))
        parsed.append((acc.astype(np.float32), np.asarray(tags, np.int32), np.asarray(bm_w, np.float64)))
    return texts, parsed
