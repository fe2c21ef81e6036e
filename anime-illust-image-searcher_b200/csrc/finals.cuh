// Combined scores on demand.
//
// webui.py:376-383 forms  final = BM25_WEIGHT * bm25 / max(bm25) + DOC2VEC_WEIGHT * sims / max(sims)  for every doc.
// The engine never stores that fp64 array (8 B written + 8 B re-read per doc and query were the largest traffic of a
// batched step): what stays resident per query is the fp32 dot score of every doc (`sim`, 4 B) and the BM25 RECORD of
// every 256-doc tile - a 256-bit map of the docs whose BM25 value differs from the query's default (0, or -inf when the
// query has a required term), their fp64 values and their positions inside the tile, compacted in doc order in a
// per-query pool (bm25.cuh).  Every kernel that needs combined scores (tile maxima, the collect passes, the near-tie
// witness, the exact fallbacks, the test seams) recomputes them for the tiles it visits with tile_finals(), with the
// reference's operations and precisions - so all of them see the same bits.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int FIN_TILE = 256;                    // docs per tile (= BM25_SUB = SEL_TILE)
constexpr int FIN_U = FIN_TILE / 32;             // docs per lane

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

// the per-query constants of the combine (webui.py:376-383)
struct QNorm {
    double maxb; float maxs; float rmax; bool m_safe; double dflt;      // dflt: bm25 / max of a doc without a record
    double wb_dflt;
};

struct FinSrc {
    const double* fin_ext;      // non-null: the combined scores were SUPPLIED (ais_rerank), [n] for query 0
    const float* sim; int64_t ld;                 // [q][ld] dot scores
    const uint32_t* tile_hdr;                     // [q][tile_ld][8] bitmap of the docs with a record
    const uint32_t* tile_off;                     // [q][tile_ld] first record slot of the tile, relative to rec_base[q]
    int64_t tile_ld;
    const double* rec_val; const uint8_t* rec_pos; const int64_t* rec_base;   // pools; rec_base [q]
    const double* maxes;        // [q][2] {max bm25, max dot} over ALL docs of ALL shards
    const int32_t* n_required;  // [q]
    double wb; float wd;        // BM25_WEIGHT (fp64 multiply), DOC2VEC_WEIGHT (fp32 multiply)
    int64_t n;

    __device__ __forceinline__ QNorm qnorm(int qi) const {
        QNorm c;
        c.maxb = maxes[2 * qi];
        c.maxs = (float)maxes[2 * qi + 1];
        const uint32_t mex = (__float_as_uint(c.maxs) >> 23) & 0xffu;
        c.m_safe = c.maxs > 0.0f && mex - 64u < 128u;             // 2^-63 <= max < 2^65
        c.rmax = c.m_safe ? __frcp_rn(c.maxs) : 0.0f;
        c.dflt = n_required[qi] > 0 ? -INFINITY : 0.0;            // 0 / max = 0, -inf / max = -inf (webui.py:379-380)
        c.wb_dflt = __dmul_rn(wb, c.dflt);
        return c;
    }
    // sims / max(sims) if max > 0 (webui.py:377-378), fp32
    __device__ __forceinline__ float sim_norm(const QNorm& c, float x) const {
        return c.maxs > 0.0f ? div_by_max(x, c.maxs, c.rmax, c.m_safe) : x;
    }
    // bm25 / max(bm25) if max > 0 (webui.py:379-380), fp64
    __device__ __forceinline__ double bm25_norm(const QNorm& c, double v) const {
        return c.maxb > 0.0 ? __ddiv_rn(v, c.maxb) : v;
    }
    // webui.py:383
    __device__ __forceinline__ double blend(double bm25n_or_wb_product, float simn) const {
        return __dadd_rn(bm25n_or_wb_product, (double)__fmul_rn(wd, simn));
    }
};

// Combined scores of the docs  tile * 256 + 32 * u + lane  (u = 0..7) of query qi, one warp per call.
// Docs beyond n get -inf and a cleared bit in `valid` (bit u).
__device__ __forceinline__ void tile_finals(const FinSrc& S, int qi, int64_t tile, int lane, double (&f)[FIN_U], unsigned& valid) {
    const int64_t lo = tile * FIN_TILE;
    valid = 0u;
    if (S.fin_ext) {
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const int64_t d = lo + 32 * u + lane;
            const bool in = d < S.n;
            f[u] = in ? S.fin_ext[d] : -INFINITY;
            valid |= in ? (1u << u) : 0u;
        }
        return;
    }
    const float* simq = S.sim + (int64_t)qi * S.ld;
    float sv[FIN_U];
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int64_t d = lo + 32 * u + lane;
        sv[u] = d < S.n ? simq[d] : 0.0f;
    }
    const uint4* hp = reinterpret_cast<const uint4*>(S.tile_hdr + ((int64_t)qi * S.tile_ld + tile) * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    const uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    const QNorm c = S.qnorm(qi);
    const double* rec = S.rec_val + S.rec_base[qi] + S.tile_off[(int64_t)qi * S.tile_ld + tile];
    const unsigned lt = (1u << lane) - 1u;
    int prefix = 0;
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int64_t d = lo + 32 * u + lane;
        double wbb = c.wb_dflt;
        if ((w[u] >> lane) & 1u) wbb = __dmul_rn(S.wb, S.bm25_norm(c, rec[prefix + __popc(w[u] & lt)]));
        prefix += __popc(w[u]);
        const bool in = d < S.n;
        f[u] = in ? S.blend(wbb, S.sim_norm(c, sv[u])) : -INFINITY;
        valid |= in ? (1u << u) : 0u;
    }
}

// one thread, one doc (the scattered readers: PRF threshold, debug seams)
__device__ __forceinline__ double doc_final(const FinSrc& S, int qi, int64_t d) {
    if (S.fin_ext) return S.fin_ext[d];
    const int64_t tile = d / FIN_TILE;
    const int l = (int)(d - tile * FIN_TILE), u = l >> 5, ln = l & 31;
    const uint32_t* w = S.tile_hdr + ((int64_t)qi * S.tile_ld + tile) * 8;
    const QNorm c = S.qnorm(qi);
    double wbb = c.wb_dflt;
    if ((w[u] >> ln) & 1u) {
        int idx = __popc(w[u] & ((1u << ln) - 1u));
        for (int k = 0; k < u; ++k) idx += __popc(w[k]);
        const double* rec = S.rec_val + S.rec_base[qi] + S.tile_off[(int64_t)qi * S.tile_ld + tile];
        wbb = __dmul_rn(S.wb, S.bm25_norm(c, rec[idx]));
    }
    return S.blend(wbb, S.sim_norm(c, S.sim[(int64_t)qi * S.ld + d]));
}

}  // namespace ais
