// BM25 over tag-major posting lists, term-at-a-time, bit-exact in fp64.
//
// Replaces compute_bm25_scores (webui.py:119-172).  The reference evaluates, per query term in
// dict order,  score = idf * (tf*(k1+1) / (tf + k1*(1 - b + b*(dl/avgdl))))  over ALL docs and then
// scores += weight*score  /  -inf masks.  Docs without the term get tf = 0 -> score = +0, which
// leaves the running sum unchanged, so walking only the posting list is exact.  fp64 addition is not
// associative, so the per-doc accumulation ORDER must be the query's term order: each CTA owns a
// tile of BM25_TILE consecutive docs (accumulators in shared memory), finds every term's slice of
// its posting list by binary search (doc ids are ascending), and processes the terms one after
// another with a barrier in between.  All arithmetic uses the _rn intrinsics (no FMA contraction).
//
// K_d = k1*(1 - b + b*(dl/avgdl)) is precomputed per doc at load time with the same operation order.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int BM25_TILE = 2048;
constexpr int BM25_THREADS = 256;

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

__global__ void __launch_bounds__(BM25_THREADS)
bm25_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc,
            const int32_t* __restrict__ post_tf,  // may be null: tf == 1
            const double* __restrict__ idf, const double* __restrict__ kd, int64_t n, int32_t n_vocab,
            const QueryTerms* __restrict__ queries, double magic, double k1p1,
            double* __restrict__ out, int64_t ld,       // [nq][ld]
            uint64_t* __restrict__ max_keys) {          // [nq] dkey images
    __shared__ double acc[BM25_TILE];
    __shared__ uint8_t excl[BM25_TILE];
    __shared__ uint8_t reqc[BM25_TILE];
    __shared__ int64_t slice[2 * MAX_TERMS];
    __shared__ uint64_t wmax[BM25_THREADS / 32];

    const int qi = blockIdx.y;
    const QueryTerms& Q = queries[qi];
    const int T = Q.n_terms;
    const int tid = threadIdx.x;
    const int64_t lo = (int64_t)blockIdx.x * BM25_TILE;
    const int64_t hi = (lo + BM25_TILE < n) ? lo + BM25_TILE : n;

    for (int i = tid; i < BM25_TILE; i += BM25_THREADS) {
        acc[i] = 0.0;
        excl[i] = 0;
        reqc[i] = 0;
    }
    // lower_bound of lo / hi inside every term's posting list
    if (tid < 2 * T) {
        const int j = tid >> 1;
        const int t = Q.term[j];
        int64_t a = 0, b = 0;
        if (t >= 0 && t < n_vocab) {
            a = post_ptr[t];
            b = post_ptr[t + 1];
        }
        const int64_t target = (tid & 1) ? hi : lo;
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if ((int64_t)post_doc[m] < target) a = m + 1; else b = m;
        }
        slice[tid] = a;
    }
    __syncthreads();

    int n_required = 0;
    for (int j = 0; j < T; ++j) {
        const double w = Q.weight[j];
        const int t = Q.term[j];
        const int64_t a = slice[2 * j], b = slice[2 * j + 1];
        if (w < 0.0) {
            // webui.py:154-160: docs CONTAINING the term -> -inf, nothing added
            for (int64_t p = a + tid; p < b; p += BM25_THREADS) excl[post_doc[p] - lo] = 1;
        } else {
            const bool required = w > magic;            // webui.py:161  (1000 itself is NOT required)
            const double mult = required ? (w - magic) : w;
            const double idfv = (t >= 0 && t < n_vocab) ? idf[t] : 0.0;
            if (required) ++n_required;
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
    for (int64_t d = lo + tid; d < hi; d += BM25_THREADS) {
        const int l = (int)(d - lo);
        // webui.py:160,168: excluded hit, or a required term missing -> -inf (absorbing under +=)
        const double v = (excl[l] || reqc[l] != n_required) ? -INFINITY : acc[l];
        out[(int64_t)qi * ld + d] = v;
        const uint64_t k = dkey(v);
        best = k > best ? k : best;
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
