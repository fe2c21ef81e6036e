// BM25 over tag-major posting lists, term-at-a-time, bit-exact in fp64.
//
// Replaces compute_bm25_scores (webui.py:119-172).  The reference evaluates, per query term in
// dict order,  score = idf * (tf*(k1+1) / (tf + k1*(1 - b + b*(dl/avgdl))))  over ALL docs and then
// scores += weight*score  /  -inf masks.  Docs without the term get tf = 0 -> score = +0, which
// leaves the running sum unchanged, so walking only the posting list is exact.  fp64 addition is not
// associative, so the per-doc accumulation ORDER must be the query's term order: each CTA owns a
// tile of BM25_TILE consecutive docs (accumulators in shared memory), reads every term's slice of its
// posting list from a table built by bm25_slices_kernel (doc ids are ascending), and processes the terms
// one after another with a barrier in between.  All arithmetic uses the _rn intrinsics (no FMA contraction).
//
// K_d = k1*(1 - b + b*(dl/avgdl)) is precomputed per doc at load time with the same operation order.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int BM25_TILE = 8192;                  // docs per block: fewer, fatter blocks amortise the per-term barriers
constexpr int BM25_THREADS = 512;
constexpr int BM25_SMEM = BM25_TILE * (8 + 1 + 1);   // fp64 accumulators + exclude flags + required-term counters

struct QueryTerms {  // device-resident, one per query of the pass
    int32_t n_terms;
    int32_t term[MAX_TERMS];
    double weight[MAX_TERMS];
};

__global__ void kd_kernel(const int64_t* __restrict__ doc_len, int64_t n, double avgdl, double k1, double b,
                          double one_minus_b, double* __restrict__ kd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // k1 * (1 - b + b * (dl / bm25_avgdl))      webui.py:145
    const double r = __ddiv_rn((double)doc_len[i], avgdl);
    kd[i] = __dmul_rn(k1, __dadd_rn(one_minus_b, __dmul_rn(b, r)));
}

// slices[(q * t_cap + j) * (n_tiles + 1) + tile] = first posting of query q's j-th term whose doc id is
// >= tile * BM25_TILE.  One thread per entry: the binary searches are independent, so their latency is
// hidden by parallelism instead of being paid serially at the head of every bm25 block.
__global__ void bm25_slices_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc, int32_t n_vocab,
                                   const QueryTerms* __restrict__ queries, int t_cap, int64_t n_tiles,
                                   int64_t* __restrict__ slices) {
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int qi = blockIdx.y / t_cap, j = blockIdx.y - qi * t_cap;
    if (tile > n_tiles || j >= queries[qi].n_terms) return;
    const int t = queries[qi].term[j];
    int64_t a = 0, b = 0;
    if (t >= 0 && t < n_vocab) { a = post_ptr[t]; b = post_ptr[t + 1]; }
    const int64_t target = tile * BM25_TILE;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if ((int64_t)post_doc[m] < target) a = m + 1; else b = m;
    }
    slices[((int64_t)qi * t_cap + j) * (n_tiles + 1) + tile] = a;
}

__global__ void __launch_bounds__(BM25_THREADS)
bm25_kernel(const int64_t* __restrict__ slices, int t_cap, int64_t n_tiles, const int32_t* __restrict__ post_doc,
            const int32_t* __restrict__ post_tf,  // may be null: tf == 1
            const double* __restrict__ idf, const double* __restrict__ kd, int64_t n, int32_t n_vocab,
            const QueryTerms* __restrict__ queries, double magic, double k1p1,
            double* __restrict__ out, int64_t ld,       // [nq][ld]
            uint64_t* __restrict__ max_keys) {          // [nq] dkey images
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    double* acc = reinterpret_cast<double*>(bm25_smem);
    uint8_t* excl = bm25_smem + (size_t)BM25_TILE * 8;
    uint8_t* reqc = excl + BM25_TILE;
    __shared__ uint64_t wmax[BM25_THREADS / 32];

    const int qi = blockIdx.y;
    const QueryTerms& Q = queries[qi];
    const int T = Q.n_terms;
    const int tid = threadIdx.x;
    const int64_t lo = (int64_t)blockIdx.x * BM25_TILE;
    const int64_t hi = (lo + BM25_TILE < n) ? lo + BM25_TILE : n;
    const int64_t* sl = slices + (int64_t)qi * t_cap * (n_tiles + 1) + blockIdx.x;

    // does any term of the query touch this tile at all?  (most tiles of most queries: no)
    bool touched = false;
    for (int j = 0; j < T; ++j) touched = touched || (sl[(int64_t)j * (n_tiles + 1)] != sl[(int64_t)j * (n_tiles + 1) + 1]);
    int n_required = 0;
    if (touched) {
        for (int i = tid; i < BM25_TILE; i += BM25_THREADS) {
            acc[i] = 0.0;
            excl[i] = 0;
            reqc[i] = 0;
        }
        __syncthreads();
    }
    for (int j = 0; j < T; ++j) {
        const double w = Q.weight[j];
        const bool required = w > magic;                // webui.py:161  (1000 itself is NOT required)
        if (required) ++n_required;
        if (!touched) continue;
        const int t = Q.term[j];
        const int64_t a = sl[(int64_t)j * (n_tiles + 1)], b = sl[(int64_t)j * (n_tiles + 1) + 1];
        if (w < 0.0) {
            // webui.py:154-160: docs CONTAINING the term -> -inf, nothing added
            for (int64_t p = a + tid; p < b; p += BM25_THREADS) excl[post_doc[p] - lo] = 1;
        } else {
            const double mult = required ? (w - magic) : w;
            const double idfv = (t >= 0 && t < n_vocab) ? idf[t] : 0.0;
            for (int64_t p = a + tid; p < b; p += BM25_THREADS) {
                const int d = post_doc[p];
                const double tf = post_tf ? (double)post_tf[p] : 1.0;
                const double denom = __dadd_rn(tf, kd[d]);                       // webui.py:145
                const double numer = __dmul_rn(tf, k1p1);                        // webui.py:146
                const double score = __dmul_rn(idfv, __ddiv_rn(numer, denom));   // webui.py:147
                const int l = (int)(d - lo);
                acc[l] = __dadd_rn(acc[l], __dmul_rn(mult, score));              // webui.py:167,170
                if (required) reqc[l] = (uint8_t)(reqc[l] + 1);
            }
        }
        __syncthreads();
    }

    uint64_t best = dkey(-INFINITY);
    if (touched) {
        for (int64_t d = lo + tid; d < hi; d += BM25_THREADS) {
            const int l = (int)(d - lo);
            // webui.py:160,168: excluded hit, or a required term missing -> -inf (absorbing under +=)
            const double v = (excl[l] || reqc[l] != n_required) ? -INFINITY : acc[l];
            out[(int64_t)qi * ld + d] = v;
            const uint64_t k = dkey(v);
            best = k > best ? k : best;
        }
    } else {
        // untouched tile: every doc scores +0.0, or -inf when the query has a required term (all of them lack it)
        const double v = n_required > 0 ? -INFINITY : 0.0;
        for (int64_t d = lo + tid; d < hi; d += BM25_THREADS) out[(int64_t)qi * ld + d] = v;
        if (lo < hi) best = dkey(v);
    }
    best = warp_max_u64(best);
    if ((tid & 31) == 0) wmax[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < BM25_THREADS / 32; ++w) best = wmax[w] > best ? wmax[w] : best;
        atomicMax(reinterpret_cast<unsigned long long*>(&max_keys[qi]), (unsigned long long)best);
    }
}

}  // namespace ais
