// Combined scores on demand.
//
// webui.py:376-383 forms  final = BM25_WEIGHT * bm25 / max(bm25) + DOC2VEC_WEIGHT * sims / max(sims)  for every doc.
// The engine never stores that fp64 array (8 B written + 8 B re-read per doc and query were the largest traffic of a
// batched step).  What stays resident per query is the fp32 dot score of every doc (`sim`, 4 B); the BM25 side of a
// 256-doc tile is re-derived wherever it is needed from one of two sources:
//   * BITMAPS (tf == 1 indexes - every real tagger output): one 256-bit presence map per (distinct query term of the
//     batch, tile), built once per batch by streaming each term's posting list (bm25.cuh, bm25_bitmap_kernel).  A
//     tile's BM25 values are then a pure function of <= T bytes per lane, the per-doc quotient g1 and the query's
//     weights - ONE round of independent loads, no posting walk, no per-query intermediate at all.
//   * RECORDS (indexes with tf > 1, or when the bitmaps would not fit): a 256-bit map of the docs whose BM25 value
//     differs from the query's default (0, or -inf when the query has a required term) + those fp64 values and their
//     positions, compacted in doc order in a per-query pool (bm25.cuh, bm25_score_kernel).
// Every kernel that needs combined scores (tile maxima, the collect passes, the near-tie witness, the exact fallbacks,
// the test seams) calls tile_finals(), with the reference's operations and precisions - so all of them see the same bits.
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

// v / m for fp64 the same way (webui.py:379-380 divides the BM25 scores by their maximum): y = RN(1/m), q = RN(v*y),
// r = v - m*q exactly (FMA), q' = RN(q + r*y).  Checked against exact rational arithmetic on 4e5 operand pairs incl. the
// all-ones-mantissa divisors (tools/check_fma_division.py: 0 mismatches); operands outside a wide exponent window, zero,
// infinities and an all-ones divisor take the IEEE division.  ~3 fp64 instructions instead of the ~40 of __ddiv_rn.
__device__ __forceinline__ double ddiv_by_max(double v, double m, double y, bool m_safe) {
    const uint32_t ex = ((uint32_t)(__double2hiint(v)) >> 20) & 0x7ffu;
    if (m_safe && ex - 523u < 1000u) {                 // 2^-500 <= |v| < 2^500
        const double q = __dmul_rn(v, y);
        const double r = __fma_rn(-m, q, v);
        return __fma_rn(r, y, q);
    }
    if (m_safe && (v == -INFINITY || v == 0.0)) return v;     // masked docs (webui.py:160,168) / zeros: m is finite and positive
    return __ddiv_rn(v, m);
}

// ---- BM25 values of a tile from the batch's term bitmaps (tf == 1) ------------------------------------------------
// Layout: bits[(slot * n_tiles + tile) * 32 + l] is a BYTE whose bit u says "doc tile*256 + 32*u + l carries the term":
// lane l of a warp needs exactly that byte for its 8 docs, so a term costs one coalesced 32-byte load per tile.
struct BitSrc {
    const uint8_t* bits; int64_t n_tiles;
    const QueryTerms* queries; const double* q_idf;     // q_idf [q][MAX_TERMS]: idf of the query's terms (0 if absent, webui.py:140)
    const double* g1;                                   // per doc: (k1 + 1) / (1 + K_d), the tf == 1 quotient (webui.py:145-147)
    double magic;                                       // REQUIRE_TAG_MAGIC_NUMBER
};

// v[u] = compute_bm25_scores(...)[tile*256 + 32*u + lane] (webui.py:139-170): the terms are added in the query's order
// (fp64 addition is not associative), `scores += weight * (idf * quotient)` exactly as the reference rounds it; an
// excluded term's docs become -inf (absorbing under the later additions), a doc that misses a required term -inf.
// Docs beyond n get the query's default.
__device__ __forceinline__ void bm25_tile_values(const BitSrc& B, int qi, int64_t tile, int lane, int64_t n, double (&v)[FIN_U]) {
    const QueryTerms& Q = B.queries[qi];
    const int T = Q.n_terms;
    const int n_required = Q.n_required;
    const int64_t lo = tile * FIN_TILE;
    double g[FIN_U];
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int64_t d = lo + 32 * u + lane;
        g[u] = d < n ? B.g1[d] : 0.0;
        v[u] = 0.0;
    }
    unsigned cnt_lo = 0u, cnt_hi = 0u;                   // required-term counters, one byte per u
    for (int j0 = 0; j0 < T; j0 += 4) {
        // up to four terms in flight: their presence bytes are independent loads
        unsigned by[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + r;
            const int slot = j < T ? Q.slot[j] : -1;
            by[r] = slot >= 0 ? (unsigned)B.bits[((int64_t)slot * B.n_tiles + tile) * 32 + lane] : 0u;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + r;
            if (j >= T) break;
            if (!__any_sync(0xffffffffu, by[r] != 0u)) continue;       // no doc of the tile carries the term
            const double w = Q.weight[j];
            const double idfv = B.q_idf[(size_t)qi * MAX_TERMS + j];
            if (w < 0.0) {
#pragma unroll
                for (int u = 0; u < FIN_U; ++u)
                    if ((by[r] >> u) & 1u) v[u] = -INFINITY;                                   // webui.py:154-160
            } else {
                const bool required = w > B.magic;
                const double mult = required ? (w - B.magic) : w;
#pragma unroll
                for (int u = 0; u < FIN_U; ++u)
                    if ((by[r] >> u) & 1u) v[u] = __dadd_rn(v[u], __dmul_rn(mult, __dmul_rn(idfv, g[u])));   // webui.py:147,167,170
                if (required) {                                      // per-doc count of the required terms present
                    const unsigned b = by[r];
                    cnt_lo += (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21);
                    cnt_hi += ((b >> 4) & 1u) | (((b >> 4) & 2u) << 7) | (((b >> 4) & 4u) << 14) | (((b >> 4) & 8u) << 21);
                }
            }
        }
    }
    if (n_required > 0) {
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const unsigned c = ((u < 4 ? cnt_lo : cnt_hi) >> (8 * (u & 3))) & 0xffu;
            if ((int)c != n_required) v[u] = -INFINITY;               // webui.py:168: a required term is missing
        }
    }
}

// the per-query constants of the combine (webui.py:376-383)
struct QNorm {
    double maxb; float maxs; float rmax; bool m_safe; double dflt;      // dflt: bm25 / max of a doc without a record
    double wb_dflt;
    double rmaxb; bool b_safe;                                           // RN(1 / max bm25) for ddiv_by_max
};

struct FinSrc {
    const double* fin_ext;      // non-null: the combined scores were SUPPLIED (ais_rerank), [n] for query 0
    const float* sim; int64_t ld;                 // [q][ld] dot scores
    int use_bits;               // 1: BM25 from the term bitmaps (B), 0: from the per-tile records below
    int ieee_div;               // 1: bm25 / max with __ddiv_rn instead of the FMA-corrected quotient (AIS_IEEE_DIV=1; cross-check)
    BitSrc B;
    const uint32_t* tile_hdr;                     // [q][tile_ld][8] bitmap of the docs with a record
    const uint32_t* tile_off;                     // [q][tile_ld] first record slot of the tile, relative to rec_base[q]
    int64_t tile_ld;
    const double* rec_val; const uint8_t* rec_pos; const int64_t* rec_base;   // pools; rec_base [q]
    const double* maxes;        // [q][2] {max bm25, max dot} over ALL docs of ALL shards
    const int32_t* n_required;  // [q]
    const QNorm* qtab;          // [q] the constants below, formed once per batch by qnorm_kernel (null: formed per call)
    double wb; float wd;        // BM25_WEIGHT (fp64 multiply), DOC2VEC_WEIGHT (fp32 multiply)
    int64_t n;

    __device__ __forceinline__ QNorm qnorm(int qi) const {
        if (qtab) return qtab[qi];
        return qnorm_compute(qi);
    }
    __device__ __forceinline__ QNorm qnorm_compute(int qi) const {
        QNorm c;
        c.maxb = maxes[2 * qi];
        c.maxs = (float)maxes[2 * qi + 1];
        const uint32_t mex = (__float_as_uint(c.maxs) >> 23) & 0xffu;
        c.m_safe = c.maxs > 0.0f && mex - 64u < 128u;             // 2^-63 <= max < 2^65
        c.rmax = c.m_safe ? __frcp_rn(c.maxs) : 0.0f;
        c.dflt = n_required[qi] > 0 ? -INFINITY : 0.0;            // 0 / max = 0, -inf / max = -inf (webui.py:379-380)
        c.wb_dflt = __dmul_rn(wb, c.dflt);
        const uint32_t bhi = (uint32_t)__double2hiint(c.maxb), bex = (bhi >> 20) & 0x7ffu;
        const bool ones = (bhi & 0xfffffu) == 0xfffffu && (uint32_t)__double2loint(c.maxb) == 0xffffffffu;
        c.b_safe = c.maxb > 0.0 && bex - 523u < 1000u && !ones && !ieee_div;
        c.rmaxb = c.b_safe ? __drcp_rn(c.maxb) : 0.0;
        return c;
    }
    // sims / max(sims) if max > 0 (webui.py:377-378), fp32
    __device__ __forceinline__ float sim_norm(const QNorm& c, float x) const {
        return c.maxs > 0.0f ? div_by_max(x, c.maxs, c.rmax, c.m_safe) : x;
    }
    // bm25 / max(bm25) if max > 0 (webui.py:379-380), fp64
    __device__ __forceinline__ double bm25_norm(const QNorm& c, double v) const {
        return c.maxb > 0.0 ? ddiv_by_max(v, c.maxb, c.rmaxb, c.b_safe) : v;
    }
    // webui.py:383
    __device__ __forceinline__ double blend(double bm25n_or_wb_product, float simn) const {
        return __dadd_rn(bm25n_or_wb_product, (double)__fmul_rn(wd, simn));
    }
};

// the per-query constants once per batch (every tile of every pass would otherwise re-derive them: two reciprocals, the
// exponent-window tests)
__global__ void qnorm_kernel(FinSrc S, int nq, QNorm* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < nq) out[qi] = S.qnorm_compute(qi);
}

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
    const QNorm c = S.qnorm(qi);
    if (S.use_bits) {
        double v[FIN_U];
        bm25_tile_values(S.B, qi, tile, lane, S.n, v);
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const int64_t d = lo + 32 * u + lane;
            const bool in = d < S.n;
            // a doc at the query's default skips the division: 0 / max = 0, -inf / max = -inf
            const double wbb = v[u] == c.dflt ? c.wb_dflt : __dmul_rn(S.wb, S.bm25_norm(c, v[u]));
            f[u] = in ? S.blend(wbb, S.sim_norm(c, sv[u])) : -INFINITY;
            valid |= in ? (1u << u) : 0u;
        }
        return;
    }
    const uint4* hp = reinterpret_cast<const uint4*>(S.tile_hdr + ((int64_t)qi * S.tile_ld + tile) * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    const uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
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

}  // namespace ais
