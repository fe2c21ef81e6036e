// BM25 over tag-major posting lists, term-at-a-time, bit-exact in fp64.
//
// Replaces compute_bm25_scores (webui.py:119-172).  The reference evaluates, per query term in
// dict order,  score = idf * (tf*(k1+1) / (tf + k1*(1 - b + b*(dl/avgdl))))  over ALL docs and then
// scores += weight*score  /  -inf masks.  Docs without the term get tf = 0 -> score = +0, which
// leaves the running sum unchanged, so walking only the posting list is exact.  fp64 addition is not
// associative, so the per-doc accumulation ORDER must be the query's term order: each WARP owns a
// sub-tile of BM25_SUB consecutive docs for one query (accumulators in shared memory), reads every term's
// slice of its posting list from a table built by bm25_slices_kernel (doc ids are ascending), and adds the
// terms one after another.  All arithmetic uses the _rn intrinsics (no FMA contraction).
//
// K_d = k1*(1 - b + b*(dl/avgdl)) and, for tf == 1, the whole quotient (k1+1)/(1 + K_d) are precomputed per doc
// at load time with the same operation order.  Two kernels around the global maxima (webui.py:379 divides by the
// maximum over ALL docs): bm25_score_kernel leaves a compact per-sub-tile record, bm25_combine_kernel reads it.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int BM25_SUB = 256;                    // docs per warp-private sub-tile
constexpr int BM25_WARPS = 8;                    // warps (= consecutive sub-tiles of one query) per block
constexpr int BM25_THREADS = 32 * BM25_WARPS;
constexpr int BM25_STAGE = 256;                  // staged postings per (sub-tile, query); denser slices take the direct path
constexpr int BM25_WARP_SMEM = BM25_SUB * 8 + BM25_STAGE * 8 + 264 + MAX_TERMS * 8 + 2 * BM25_SUB;
constexpr int BM25_SMEM = BM25_WARPS * BM25_WARP_SMEM;

struct QueryTerms {  // device-resident, one per query of the pass
    int32_t n_terms;
    int32_t term[MAX_TERMS];
    double weight[MAX_TERMS];
};

__global__ void kd_kernel(const int64_t* __restrict__ doc_len, int64_t n, double avgdl, double k1, double b,
                          double one_minus_b, double k1p1, double* __restrict__ kd, double* __restrict__ g1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // k1 * (1 - b + b * (dl / bm25_avgdl))      webui.py:145
    const double r = __ddiv_rn((double)doc_len[i], avgdl);
    const double k = __dmul_rn(k1, __dadd_rn(one_minus_b, __dmul_rn(b, r)));
    kd[i] = k;
    // tf == 1 (every real tagger output: tags of an image are unique): tf*(k1+1) / (tf + K_d) with the same operations,
    // once per doc instead of once per posting and query            webui.py:145-147
    g1[i] = __ddiv_rn(__dmul_rn(1.0, k1p1), __dadd_rn(1.0, k));
}

// slices[(q * t_cap + j) * (n_sub + 1) + sub] = first posting of query q's j-th term whose doc id is
// >= sub * BM25_SUB (absolute index into post_doc).  One thread per entry: the binary searches are
// independent, so their latency is hidden by parallelism instead of being paid serially in the scoring kernel.
__global__ void bm25_slices_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc, int32_t n_vocab,
                                   const QueryTerms* __restrict__ queries, int t_cap, int64_t n_sub,
                                   int64_t* __restrict__ slices, double magic, int32_t* __restrict__ n_required) {
    const int64_t sub = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int qi = blockIdx.y / t_cap, j = blockIdx.y - qi * t_cap;
    if (sub == 0 && j == 0) {                            // per-query constant of the combine phase (webui.py:161)
        int r = 0;
        for (int t = 0; t < queries[qi].n_terms; ++t) r += queries[qi].weight[t] > magic;
        n_required[qi] = r;
    }
    if (sub > n_sub || j >= queries[qi].n_terms) return;
    const int t = queries[qi].term[j];
    int64_t a = 0, b = 0;
    if (t >= 0 && t < n_vocab) { a = post_ptr[t]; b = post_ptr[t + 1]; }
    const int64_t target = sub * BM25_SUB;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if ((int64_t)post_doc[m] < target) a = m + 1; else b = m;
    }
    slices[((int64_t)qi * t_cap + j) * (n_sub + 1) + sub] = a;
}

struct Bm25Args {
    const int64_t* slices; int t_cap; int64_t n_sub;
    const int32_t* post_doc; const int32_t* post_tf;          // post_tf may be null: tf == 1
    const double* idf; const double* kd; const double* g1; int64_t n; int32_t n_vocab;   // g1: the tf == 1 quotient per doc
    const QueryTerms* queries; double magic, k1p1;
    // phase 0: per-query maximum + the sub-tile's record (+ optional dense scores for the compute_bm25_scores seam)
    uint64_t* max_keys; double* dense_out; int64_t ld;
    uint32_t* tile_hdr;                                        // [q][tile_ld][8]: bitmap of the docs whose BM25 value is not the default
    const int32_t* n_required;                                 // [q] number of required terms (bm25_slices_kernel)
    // phase 1: normalise + combine with the dot scores (webui.py:376-383), store the combined scores, segment maxima
    const float* sim; double* fin; const double* maxes; double wb; float wd;
    uint64_t* seg_max; int seg_mod;                            // seg_max[q][sub % seg_mod]
    uint64_t* tile_max; int64_t tile_ld;                       // tile_max[q][sub]: best key of the sub-tile (the collect pass skips by it)
};

// Phase 0.  One WARP per (sub-tile of BM25_SUB docs, query), no block barriers: every warp runs its own dependency chain
// (slice bounds -> postings -> K_d -> fp64 contribution -> ordered accumulation -> record), so an SM overlaps ~40 of
// them.  The postings of ALL the query's terms that fall into the sub-tile are first staged in the warp's shared memory
// with their contributions (independent global loads), then summed term by term, in the query's term order (fp64
// addition is not associative; webui.py:139-170 adds term by term), out of shared memory only.
//
// Result: the per-query maximum (webui.py:379 needs it before anything can be combined) and the sub-tile's RECORD -
// a 256-bit map of the docs whose BM25 value differs from the default of the query (0, or -inf when the query has a
// required term) in tile_hdr, and those values, compacted in doc order, in the first slots of the sub-tile's own range
// of the combined-score array `fin` (which phase 1 of the same warp position overwrites with the final scores).  Phase 1
// therefore never touches posting lists, K_d or an fp64 division per posting again.
__global__ void __launch_bounds__(BM25_THREADS, 5)
bm25_score_kernel(Bm25Args A) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t sub = (int64_t)blockIdx.x * BM25_WARPS + warp;
    if (sub >= A.n_sub) return;
    unsigned char* base = bm25_smem + (size_t)warp * BM25_WARP_SMEM;
    double* acc = reinterpret_cast<double*>(base);
    double* st_val = reinterpret_cast<double*>(base + BM25_SUB * 8);
    int* off = reinterpret_cast<int*>(base + BM25_SUB * 8 + BM25_STAGE * 8);                 // [MAX_TERMS + 1]
    int64_t* sl_a = reinterpret_cast<int64_t*>(base + BM25_SUB * 8 + BM25_STAGE * 8 + 264);  // [MAX_TERMS]
    uint8_t* reqc = base + BM25_SUB * 8 + BM25_STAGE * 8 + 264 + MAX_TERMS * 8;
    uint8_t* st_doc = reqc + BM25_SUB;

    const int qi = blockIdx.y;
    const QueryTerms& Q = A.queries[qi];
    const int T = Q.n_terms;
    const int64_t lo = sub * BM25_SUB;
    const int64_t hi = (lo + BM25_SUB < A.n) ? lo + BM25_SUB : A.n;
    const int64_t* sl = A.slices + (int64_t)qi * A.t_cap * (A.n_sub + 1) + sub;
    uint32_t* hdr = A.tile_hdr + ((int64_t)qi * A.tile_ld + sub) * 8;

    // slice bounds of every term, exclusive prefix sum of their lengths, number of required terms
    int n_required = 0, carry = 0;
    for (int j0 = 0; j0 < T; j0 += 32) {
        const int j = j0 + lane;
        int len = 0;
        bool req = false;
        if (j < T) {
            const int64_t a = sl[(int64_t)j * (A.n_sub + 1)], b = sl[(int64_t)j * (A.n_sub + 1) + 1];
            sl_a[j] = a;
            len = (int)(b - a);
            req = Q.weight[j] > A.magic;                 // webui.py:161 (1000 itself is NOT required)
        }
        n_required += __popc(__ballot_sync(0xffffffffu, req));
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (j < T) off[j + 1] = carry + inc;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) off[0] = 0;
    const int E = carry;
    __syncwarp();
    // a doc no term touches scores +0.0, or -inf when the query has a required term (webui.py:160,168)
    const double untouched = n_required > 0 ? -INFINITY : 0.0;

    if (E == 0) {
        if (lane < 8) hdr[lane] = 0u;
        if (A.dense_out)
            for (int64_t d = lo + lane; d < hi; d += 32) A.dense_out[(int64_t)qi * A.ld + d] = untouched;
        const uint64_t k = dkey(untouched);              // one relaxed check instead of 256 identical keys
        if (lane == 0 && k > *(volatile uint64_t*)&A.max_keys[qi])
            atomicMax(reinterpret_cast<unsigned long long*>(&A.max_keys[qi]), (unsigned long long)k);
        return;
    }

    // An excluded doc simply becomes -inf in its accumulator (absorbing under the later additions, webui.py:160); the
    // per-doc count of required terms exists only for queries that have one (30 % of the benchmark's queries).
    const bool has_req = n_required > 0;
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) acc[u * 32 + lane] = 0.0;
    if (has_req)
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u) reqc[u * 32 + lane] = 0;
    __syncwarp();
    if (E <= BM25_STAGE) {
        // ---- stage every posting of the sub-tile with its contribution (all global loads independent) ----
        for (int e = lane; e < E; e += 32) {
            int j = 0;
            while (e >= off[j + 1]) ++j;
            const int64_t p = sl_a[j] + (e - off[j]);
            const int d = A.post_doc[p];
            const double w = Q.weight[j];
            double c = 0.0;
            if (!(w < 0.0)) {
                const int t = Q.term[j];
                const double mult = (w > A.magic) ? (w - A.magic) : w;
                const double idfv = (t >= 0 && t < A.n_vocab) ? A.idf[t] : 0.0;
                double quot;
                if (A.post_tf) {
                    const double tf = (double)A.post_tf[p];
                    const double denom = __dadd_rn(tf, A.kd[d]);                   // webui.py:145
                    const double numer = __dmul_rn(tf, A.k1p1);                    // webui.py:146
                    quot = __ddiv_rn(numer, denom);
                } else {
                    quot = A.g1[d];                                                // the same quotient for tf == 1, precomputed
                }
                const double score = __dmul_rn(idfv, quot);                        // webui.py:147
                c = __dmul_rn(mult, score);                                        // webui.py:167,170
            }
            st_doc[e] = (uint8_t)(d - lo);
            st_val[e] = c;
        }
        __syncwarp();
        // ---- ordered accumulation, shared memory only (doc ids are unique inside a term) ----
        for (int j = 0; j < T; ++j) {
            const double w = Q.weight[j];
            const bool required = w > A.magic;
            for (int e = off[j] + lane; e < off[j + 1]; e += 32) {
                const int l = st_doc[e];
                if (w < 0.0) acc[l] = -INFINITY;                                   // webui.py:154-160
                else {
                    acc[l] = __dadd_rn(acc[l], st_val[e]);
                    if (required) reqc[l] = (uint8_t)(reqc[l] + 1);
                }
            }
            __syncwarp();
        }
    } else {
        // ---- direct path for very dense sub-tiles ----
        for (int j = 0; j < T; ++j) {
            const double w = Q.weight[j];
            const int t = Q.term[j];
            const int64_t a = sl_a[j], b = a + (off[j + 1] - off[j]);
            if (w < 0.0) {
                for (int64_t p = a + lane; p < b; p += 32) acc[A.post_doc[p] - lo] = -INFINITY;
            } else {
                const bool required = w > A.magic;
                const double mult = required ? (w - A.magic) : w;
                const double idfv = (t >= 0 && t < A.n_vocab) ? A.idf[t] : 0.0;
                for (int64_t p = a + lane; p < b; p += 32) {
                    const int d = A.post_doc[p];
                    double quot;
                    if (A.post_tf) {
                        const double tf = (double)A.post_tf[p];
                        quot = __ddiv_rn(__dmul_rn(tf, A.k1p1), __dadd_rn(tf, A.kd[d]));
                    } else {
                        quot = A.g1[d];
                    }
                    const double score = __dmul_rn(idfv, quot);
                    const int l = (int)(d - lo);
                    acc[l] = __dadd_rn(acc[l], __dmul_rn(mult, score));
                    if (required) reqc[l] = (uint8_t)(reqc[l] + 1);
                }
            }
            __syncwarp();
        }
    }

    // ---- record: bitmap + compacted values (doc order), and the maximum over the sub-tile ----
    // webui.py:160,168: excluded hit, or a required term missing -> -inf (absorbing under +=)
    double* rec = A.fin + (int64_t)qi * A.ld + lo;
    double bd = -INFINITY;                                        // maximum in the double domain, one key at the end
    bool any = false;
    int n_rec = 0;
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        const int l = u * 32 + lane;
        const bool in = lo + l < hi;
        double v = acc[l];
        if (has_req && reqc[l] != n_required) v = -INFINITY;      // webui.py:168: a required term is missing
        if (A.dense_out && in) A.dense_out[(int64_t)qi * A.ld + lo + l] = v;
        const bool rec_it = in && v != untouched;                 // NaN never appears: idf, K_d and the weights are finite
        const unsigned m = __ballot_sync(0xffffffffu, rec_it);
        if (lane == 0) hdr[u] = m;
        if (rec_it) {
            rec[n_rec + __popc(m & ((1u << lane) - 1u))] = v;
            bd = fmax(bd, v);
            any = true;
        }
        n_rec += __popc(m);
    }
    if (n_rec < (int)(hi - lo)) {                                 // some doc keeps the default
        bd = fmax(bd, untouched);
        any = true;
    }
    uint64_t best = any ? dkey(bd) : KEY_EMPTY;
    {   // 64-bit warp maximum with two redux.sync
        const uint32_t bh = (uint32_t)(best >> 32), bl = (uint32_t)best;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, bh);
        const uint32_t ml = __reduce_max_sync(0xffffffffu, bh == mh ? bl : 0u);
        best = ((uint64_t)mh << 32) | ml;
    }
    if (lane == 0 && best > *(volatile uint64_t*)&A.max_keys[qi])
        atomicMax(reinterpret_cast<unsigned long long*>(&A.max_keys[qi]), (unsigned long long)best);
}

// x / m rounded to nearest with three instructions (Markstein: y = RN(1/m), q = RN(x*y), r = x - m*q exactly by FMA,
// q' = RN(q + r*y) is the correctly rounded quotient when nothing over- or underflows); operands outside a safe
// exponent window take the full IEEE division.  webui.py:377-378 divides fp32 by fp32.
__device__ __forceinline__ float div_by_max(float x, float m, float y, bool m_safe) {
    const uint32_t ex = (__float_as_uint(x) >> 23) & 0xffu;
    if (m_safe && ex - 64u < 128u) {                   // 2^-63 <= |x| < 2^65
        const float q = __fmul_rn(x, y);
        const float r = __fmaf_rn(-m, q, x);
        return __fmaf_rn(r, y, q);
    }
    return __fdiv_rn(x, m);
}

constexpr int BM25C_WARPS = 8;
constexpr int BM25C_THREADS = 32 * BM25C_WARPS;

// Phase 1 (after the global maxima are known).  One warp per (sub-tile, query) again: read the record phase 0 left
// (32-byte bitmap + compacted values), bm25 / max for the recorded docs (full warps on the fp64 division), then per
// doc  final = wb * bm25n + wd * (sim / max sim)  (webui.py:376-383; fp32 / fp32, the products and the sum in the
// reference's precisions), store it, and keep the sub-tile's best key for the select.
__global__ void __launch_bounds__(BM25C_THREADS, 4)
bm25_combine_kernel(Bm25Args A) {
    __shared__ double vals_all[BM25C_WARPS][BM25_SUB + 1];        // + 1: the branch-free lookup below may touch slot n_rec
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t sub = (int64_t)blockIdx.x * BM25C_WARPS + warp;
    if (sub >= A.n_sub) return;
    double* vals = vals_all[warp];
    const int qi = blockIdx.y;
    const int64_t lo = sub * BM25_SUB;
    const int64_t hi = (lo + BM25_SUB < A.n) ? lo + BM25_SUB : A.n;
    const float* simq = A.sim + (int64_t)qi * A.ld;
    double* finq = A.fin + (int64_t)qi * A.ld;

    // all the independent global loads first: the dot scores of the sub-tile, the bitmap (every lane holds all of it:
    // two broadcast 16-byte loads), the per-query constants
    float sv[BM25_SUB / 32];
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        const int64_t d = lo + u * 32 + lane;
        sv[u] = d < hi ? __ldcs(simq + d) : 0.0f;
    }
    const uint4* hp = reinterpret_cast<const uint4*>(A.tile_hdr + ((int64_t)qi * A.tile_ld + sub) * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    // the first 64 record slots are fetched before the bitmap says how many there are (average: ~50): one dependent
    // round trip less for most sub-tiles
    double pre[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) pre[r] = lo + 32 * r + lane < hi ? finq[lo + 32 * r + lane] : 0.0;
    const double maxb = A.maxes[2 * qi];
    const float maxs = (float)A.maxes[2 * qi + 1];
    const int n_required = A.n_required[qi];
    const uint32_t mex = (__float_as_uint(maxs) >> 23) & 0xffu;
    const bool m_safe = maxs > 0.0f && mex - 64u < 128u;          // 2^-63 <= max < 2^65
    const float rmax = m_safe ? __frcp_rn(maxs) : 0.0f;
    const uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    int n_rec = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) n_rec += __popc(w[u]);

    double dflt = n_required > 0 ? -INFINITY : 0.0;               // bm25n of a doc without a record: 0 / max = 0, -inf / max = -inf
    if (n_rec > 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = 32 * r + lane;
            if (i < n_rec) vals[i] = maxb > 0.0 ? __ddiv_rn(pre[r], maxb) : pre[r];     // webui.py:379-380
        }
        for (int i = 64 + lane; i < n_rec; i += 32) {
            double v = finq[lo + i];
            if (maxb > 0.0) v = __ddiv_rn(v, maxb);
            vals[i] = v;
        }
        __syncwarp();
    }
    const double wb_dflt = __dmul_rn(A.wb, dflt);

    double fbest = -INFINITY;
    bool any = false, nan_seen = false;
    const unsigned lt = (1u << lane) - 1u;
    // Fast path (all but the last sub-tile of a shard, ordinary magnitudes): no branch per doc.  The exact division
    // runs as three instructions; the record lookup is a select (the slot index is always in bounds); a NaN score is
    // detected by a running sum (the scores are finite or -inf, so the sum is NaN only if a score is).
    bool ok = m_safe && hi - lo == BM25_SUB;
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        const uint32_t ex = (__float_as_uint(sv[u]) >> 23) & 0xffu;
        ok = ok && (ex - 64u < 128u || sv[u] == 0.0f);
    }
    if (__all_sync(0xffffffffu, ok)) {
        double* fp = finq + lo + lane;
        double fsum = 0.0;
        int prefix = 0;
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u) {
            const float x = sv[u];
            const float q = __fmul_rn(x, rmax);                                    // webui.py:377-378 (fp32 / fp32), see div_by_max
            const float sn = __fmaf_rn(__fmaf_rn(-maxs, q, x), rmax, q);
            const double v = vals[prefix + __popc(w[u] & lt)];
            const double wbb = ((w[u] >> lane) & 1u) ? __dmul_rn(A.wb, v) : wb_dflt;
            prefix += __popc(w[u]);
            const double f = __dadd_rn(wbb, (double)__fmul_rn(A.wd, sn));          // webui.py:383
            __stcs(fp + u * 32, f);
            fsum += f;
            fbest = fmax(fbest, f);
        }
        any = true;
        nan_seen = fsum != fsum;
    } else {
        int prefix = 0;
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u) {
            const int64_t d = lo + u * 32 + lane;
            double wbb = wb_dflt;
            if ((w[u] >> lane) & 1u) wbb = __dmul_rn(A.wb, vals[prefix + __popc(w[u] & lt)]);
            prefix += __popc(w[u]);
            if (d < hi) {
                float sn = sv[u];
                if (maxs > 0.0f) sn = div_by_max(sn, maxs, rmax, m_safe);          // webui.py:377-378 (fp32 / fp32)
                const double f = __dadd_rn(wbb, (double)__fmul_rn(A.wd, sn));      // webui.py:383
                __stcs(finq + d, f);
                nan_seen = nan_seen || (f != f);
                fbest = any ? fmax(fbest, f) : f;
                any = true;
            }
        }
    }
    uint64_t best = any ? dkey(fbest) : KEY_EMPTY;
    if (nan_seen) best = 0xFFF8000000000000ull;                    // a NaN score (e.g. weight 0 x -inf) sorts first, as dkey(NaN) does
    best = warp_max_u64(best);
    if (lane == 0) A.tile_max[(int64_t)qi * A.tile_ld + sub] = best;
    if (lane == 0 && best != KEY_EMPTY)
        atomicMax(reinterpret_cast<unsigned long long*>(&A.seg_max[(size_t)qi * 2048 + (int)(sub % A.seg_mod)]),
                  (unsigned long long)best);
}

}  // namespace ais
