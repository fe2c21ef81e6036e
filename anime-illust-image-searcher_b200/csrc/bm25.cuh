// BM25 over tag-major posting lists, term-at-a-time, bit-exact in fp64.
//
// Replaces compute_bm25_scores (webui.py:119-172).  The reference evaluates, per query term in
// dict order,  score = idf * (tf*(k1+1) / (tf + k1*(1 - b + b*(dl/avgdl))))  over ALL docs and then
// scores += weight*score  /  -inf masks.  Docs without the term get tf = 0 -> score = +0, which
// leaves the running sum unchanged, so walking only the posting list is exact.  fp64 addition is not
// associative, so the per-doc accumulation ORDER must be the query's term order: each WARP owns a
// tile of BM25_SUB consecutive docs for one query (accumulators in shared memory), reads every term's
// slice of its posting list from a table built by bm25_slices_kernel (doc ids are ascending), and adds the
// terms one after another.  All arithmetic uses the _rn intrinsics (no FMA contraction).
//
// K_d = k1*(1 - b + b*(dl/avgdl)) and, for tf == 1, the whole quotient (k1+1)/(1 + K_d) are precomputed per doc
// at load time with the same operation order.  Two kernels around the global maxima (webui.py:379 divides by the
// maximum over ALL docs): bm25_score_kernel leaves a compact per-tile RECORD in a per-query pool, bm25_combine_kernel
// reads it and emits only the best combined score of every tile - the combined scores themselves are never stored
// (finals.cuh recomputes them for the few tiles a later stage visits).
#pragma once
#include "common.cuh"
#include "finals.cuh"

namespace ais {

constexpr int BM25_SUB = FIN_TILE;               // docs per warp-private tile
constexpr int BM25_WARPS = 8;                    // warps (= consecutive tiles of one query) per block
constexpr int BM25_THREADS = 32 * BM25_WARPS;
constexpr int BM25_WARP_SMEM = BM25_SUB * 8 + BM25_SUB;           // fp64 accumulators + required-term counters
constexpr int BM25_SMEM = BM25_WARPS * BM25_WARP_SMEM;


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

// K_d and the tf == 1 quotient depend on the doc only through its length: tables indexed by the length, and the length
// of every posting's doc next to the posting (uint16), turn the score kernel's per-posting gather into the 80 MB per-doc
// arrays into one coalesced 2-byte load plus an L1 hit.  Same operations as kd_kernel -> same bits.
__global__ void max_len_kernel(const int64_t* __restrict__ doc_len, int64_t n, unsigned long long* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0ull;
    if (i < n) v = doc_len[i] < 0 ? ~0ull : (unsigned long long)doc_len[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    if ((threadIdx.x & 31) == 0 && v > 0ull) atomicMax(out, v);
}
__global__ void kd_table_kernel(int64_t max_len, double avgdl, double k1, double b, double one_minus_b, double k1p1,
                                double* __restrict__ kd_tab, double* __restrict__ g1_tab) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > max_len) return;
    const double r = __ddiv_rn((double)i, avgdl);
    const double k = __dmul_rn(k1, __dadd_rn(one_minus_b, __dmul_rn(b, r)));
    kd_tab[i] = k;
    g1_tab[i] = __ddiv_rn(__dmul_rn(1.0, k1p1), __dadd_rn(1.0, k));
}
__global__ void post_len_kernel(const int32_t* __restrict__ post_doc, int64_t n_post, const int64_t* __restrict__ doc_len,
                                uint16_t* __restrict__ post_len) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_post) post_len[p] = (uint16_t)doc_len[post_doc[p]];
}

// slices[(q * t_cap + j) * (n_sub + 1) + sub] = first posting of query q's j-th term whose doc id is
// >= sub * BM25_SUB (absolute index into post_doc).  One thread per entry: the binary searches are
// independent, so their latency is hidden by parallelism instead of being paid serially in the scoring kernel.
// Also fills q_idf[q][j] (idf of the query's j-th term, 0 if absent: webui.py:140).
__global__ void bm25_slices_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc, int32_t n_vocab,
                                   const QueryTerms* __restrict__ queries, const double* __restrict__ idf, int t_cap, int64_t n_sub,
                                   int64_t* __restrict__ slices, int32_t* __restrict__ n_required,
                                   double* __restrict__ q_idf) {
    const int64_t sub = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int qi = blockIdx.y / t_cap, j = blockIdx.y - qi * t_cap;
    if (sub == 0 && j == 0) n_required[qi] = queries[qi].n_required;   // per-query constant of the combine phase (webui.py:161)
    if (j >= queries[qi].n_terms) return;
    const int t = queries[qi].term[j];
    const bool known = t >= 0 && t < n_vocab;
    if (sub == 0) q_idf[(size_t)qi * MAX_TERMS + j] = known ? idf[t] : 0.0;
    if (sub > n_sub) return;
    int64_t a = 0, b = 0;
    if (known) { a = post_ptr[t]; b = post_ptr[t + 1]; }
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
    const double* kd; const double* g1; int64_t n;            // g1: the tf == 1 quotient per doc
    const uint16_t* post_len; const double* kd_tab; const double* g1_tab;   // per-posting doc length + per-length tables (null: per-doc arrays)
    const QueryTerms* queries; const double* q_idf; double magic, k1p1;      // q_idf [q][MAX_TERMS]
    // score kernel: per-query maximum + the tile's record (+ optional dense scores for the compute_bm25_scores seam)
    uint64_t* max_keys; double* dense_out; int64_t ld;
    uint32_t* tile_hdr;                                        // [q][tile_ld][8]: bitmap of the docs whose BM25 value is not the default
    uint32_t* tile_off;                                        // [q][tile_ld]: first record slot of the tile (relative to rec_base[q])
    double* rec_val; uint8_t* rec_pos; const int64_t* rec_base;   // record pools, [q] pool offsets
    int64_t tile_ld;
};

// One WARP per (tile of BM25_SUB docs, query), no block barriers: every warp runs its own dependency chain (slice bounds
// -> postings -> per-doc quotient -> fp64 contribution -> ordered accumulation -> record), so an SM overlaps ~50 of them.
// The postings of ALL the query's terms that fall into the tile are enumerated term after term (e = 0 .. E-1); a lane
// takes two of them per round (64 postings = the usual tile in one round), issues their global loads back to back,
// and the warp then adds the contributions term by term, in the query's term order (fp64 addition is not associative;
// webui.py:139-170 adds term by term), into the shared-memory accumulators.
//
// Result: the per-query maximum (webui.py:379 needs it before anything can be combined) and the tile's RECORD - a
// 256-bit map of the docs whose BM25 value differs from the default of the query (0, or -inf when the query has a
// required term) in tile_hdr, and those values with their positions inside the tile, compacted in doc order, at
// rec_base[q] + tile_off[q][tile] of the record pools.  tile_off = the number of postings of the query's terms that lie
// BEFORE the tile (read off the slice table): a tile has at most as many records as postings, so the regions never
// overlap and no cursor is shared (a per-query atomic cursor serialised ~10 M atomics per batch on 256 addresses).
// One-off check at load time: every posting names a doc of the shard and the doc ids of a term ascend strictly
// - what the tile slices (binary search) and the per-tile accumulators (acc[doc - lo]) rely on.  One warp per term.
__global__ void validate_postings_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc, int n_vocab,
                                         int64_t n_docs, int* __restrict__ bad_term) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_vocab; t += warps) {
        const int64_t a = post_ptr[t], b = post_ptr[t + 1];
        bool bad = false;
        for (int64_t p = a + lane; p < b; p += 32) {
            const int32_t d = post_doc[p];
            bad = bad || d < 0 || d >= n_docs || (p > a && post_doc[p - 1] >= d);
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicMax(bad_term, t + 1);
    }
}

template <int MINB>
__global__ void __launch_bounds__(BM25_THREADS, MINB)
bm25_score_kernel(Bm25Args A) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t sub = (int64_t)blockIdx.x * BM25_WARPS + warp;
    if (sub >= A.n_sub) return;
    unsigned char* base = bm25_smem + (size_t)warp * BM25_WARP_SMEM;
    double* acc = reinterpret_cast<double*>(base);
    uint8_t* reqc = base + BM25_SUB * 8;

    const int qi = blockIdx.y;
    const QueryTerms& Q = A.queries[qi];
    const int T = Q.n_terms;
    const int n_required = Q.n_required;
    const int64_t lo = sub * BM25_SUB;
    const int64_t hi = (lo + BM25_SUB < A.n) ? lo + BM25_SUB : A.n;
    const int64_t* sl = A.slices + (int64_t)qi * A.t_cap * (A.n_sub + 1) + sub;
    uint32_t* hdr = A.tile_hdr + ((int64_t)qi * A.tile_ld + sub) * 8;
    // a doc no term touches scores +0.0, or -inf when the query has a required term (webui.py:160,168)
    const bool has_req = n_required > 0;
    const double untouched = has_req ? -INFINITY : 0.0;

    bool zeroed = false;
    unsigned int rec_off = 0u;
    // terms in groups of 32 (lane = term of the group); groups, and the terms inside one, are processed in query order
    // The loop runs to t_cap (a launch constant >= every query's term count) and a lane's loads do not wait for the
    // query record: slice bounds, weight and idf of term j are fetched together with T itself (one round trip instead of
    // two); table rows beyond the query's terms hold stale values and are masked by `j < T`.
    for (int j0 = 0; j0 < A.t_cap; j0 += 32) {
        const int j = j0 + lane;
        int64_t a = 0, b1 = 0, a0 = 0;
        double w = 0.0, idfv = 0.0;
        if (j < A.t_cap) {
            a = sl[(int64_t)j * (A.n_sub + 1)];
            b1 = sl[(int64_t)j * (A.n_sub + 1) + 1];
            a0 = sl[(int64_t)j * (A.n_sub + 1) - sub];   // slices[..][0] = start of the list
            w = Q.weight[j];
            idfv = A.q_idf[(size_t)qi * MAX_TERMS + j];
        }
        if (j0 >= T) break;                              // warp-uniform (T arrives with the loads above)
        const int len = j < T ? (int)(b1 - a) : 0;
        const unsigned int before = j < T ? (unsigned int)(a - a0) : 0u;      // postings of term j in the tiles before this one
        rec_off += __reduce_add_sync(0xffffffffu, before);
        int inc = len;                                   // inclusive prefix sum of the slice lengths
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int E = __shfl_sync(0xffffffffu, inc, 31);
        if (E == 0) continue;
        if (!zeroed) {
#pragma unroll
            for (int u = 0; u < BM25_SUB / 64; ++u) reinterpret_cast<double2*>(acc)[u * 32 + lane] = make_double2(0.0, 0.0);
            if (has_req) reinterpret_cast<uint2*>(reqc)[lane] = make_uint2(0u, 0u);
            zeroed = true;
            __syncwarp();
        }
        const int tg = (T - j0 < 32) ? T - j0 : 32;      // terms of this group
        const int64_t rel = a - (int64_t)(inc - len);    // posting index = rel(term) + e
        for (int e0 = 0; e0 < E; e0 += 64) {
            // two postings per lane: term index by counting the slice ends that are <= e
            int jj[2], l[2];
            double c[2];
            bool act[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int e = e0 + 32 * r + lane;
                act[r] = e < E;
                int t = 0;
                for (int k = 0; k < tg; ++k) t += (e >= __shfl_sync(0xffffffffu, inc, k)) ? 1 : 0;
                t = act[r] ? t : 0;
                jj[r] = t;
                const int64_t p = __shfl_sync(0xffffffffu, rel, t) + e;
                const double wt = __shfl_sync(0xffffffffu, w, t);
                const double iv = __shfl_sync(0xffffffffu, idfv, t);
                l[r] = 0;
                c[r] = 0.0;
                if (act[r]) {
                    const int d = A.post_doc[p];
                    l[r] = (int)(d - lo);
                    if (!(wt < 0.0)) {
                        double quot;
                        if (A.post_tf) {
                            const double tf = (double)A.post_tf[p];
                            const double kdv = A.post_len ? A.kd_tab[A.post_len[p]] : A.kd[d];
                            quot = __ddiv_rn(__dmul_rn(tf, A.k1p1), __dadd_rn(tf, kdv));       // webui.py:145-146
                        } else {
                            quot = A.post_len ? A.g1_tab[A.post_len[p]] : A.g1[d];             // the same quotient for tf == 1, precomputed
                        }
                        const double mult = (wt > A.magic) ? (wt - A.magic) : wt;
                        c[r] = __dmul_rn(mult, __dmul_rn(iv, quot));                           // webui.py:147,167,170
                    }
                }
            }
            // ordered accumulation: the round's postings span terms t_first..t_last (ascending with e)
            const int t_first = __shfl_sync(0xffffffffu, jj[0], 0);
            int t_last = act[1] ? jj[1] : (act[0] ? jj[0] : 0);
            t_last = __reduce_max_sync(0xffffffffu, t_last);
            for (int t = t_first; t <= t_last; ++t) {
                const double wt = __shfl_sync(0xffffffffu, w, t);
                const bool excl = wt < 0.0, required = wt > A.magic;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (act[r] && jj[r] == t) {                 // doc ids are unique inside a term: no two lanes share a slot
                        if (excl) acc[l[r]] = -INFINITY;                                       // webui.py:154-160
                        else {
                            acc[l[r]] = __dadd_rn(acc[l[r]], c[r]);
                            if (required) reqc[l[r]] = (uint8_t)(reqc[l[r]] + 1);
                        }
                    }
                }
                __syncwarp();
            }
        }
    }

    if (!zeroed) {                                       // no posting of the query falls into this tile
        if (lane < 8) hdr[lane] = 0u;
        if (lane == 0) A.tile_off[(int64_t)qi * A.tile_ld + sub] = 0u;
        if (A.dense_out)
            for (int64_t d = lo + lane; d < hi; d += 32) A.dense_out[(int64_t)qi * A.ld + d] = untouched;
        const uint64_t k = dkey(untouched);              // one relaxed check instead of 256 identical keys
        if (lane == 0 && k > *(volatile uint64_t*)&A.max_keys[qi])
            atomicMax(reinterpret_cast<unsigned long long*>(&A.max_keys[qi]), (unsigned long long)k);
        return;
    }

    // ---- record: bitmap, then values + positions compacted in doc order, and the maximum over the tile ----
    // webui.py:160,168: excluded hit, or a required term missing -> -inf (absorbing under +=).  This part runs for every
    // (tile, query): 32-bit indices, no fmax (the values are finite or -inf, never NaN: idf, K_d and the weights are finite).
    const int n_valid = (int)(hi - lo);
    double v[BM25_SUB / 32];
    uint32_t m[BM25_SUB / 32];
    int n_rec = 0;
    if (has_req) {
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u) {
            const int l = u * 32 + lane;
            v[u] = reqc[l] != n_required ? -INFINITY : acc[l];         // webui.py:168: a required term is missing
        }
    } else {
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u) v[u] = acc[u * 32 + lane];
    }
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        m[u] = __ballot_sync(0xffffffffu, u * 32 + lane < n_valid && v[u] != untouched);
        n_rec += __popc(m[u]);
    }
    if (A.dense_out) {
        double* dq = A.dense_out + (int64_t)qi * A.ld + lo;
#pragma unroll
        for (int u = 0; u < BM25_SUB / 32; ++u)
            if (u * 32 + lane < n_valid) dq[u * 32 + lane] = v[u];
    }
    const unsigned int off = rec_off;
    const int64_t rb = A.rec_base[qi] + off;
    double* rv = A.rec_val + rb;
    uint8_t* rp = A.rec_pos + rb;
    const unsigned lt = (1u << lane) - 1u;
    double bd = -INFINITY;                                        // maximum in the double domain, one key at the end
    int prefix = 0;
    uint32_t mine = 0u;                                           // lane u keeps word u of the bitmap
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        if ((m[u] >> lane) & 1u) {
            const int slot = prefix + __popc(m[u] & lt);
            rv[slot] = v[u];
            rp[slot] = (uint8_t)(u * 32 + lane);
            bd = v[u] > bd ? v[u] : bd;
        }
        prefix += __popc(m[u]);
        mine = lane == u ? m[u] : mine;
    }
    if (lane < 8) hdr[lane] = mine;
    if (lane == 0) A.tile_off[(int64_t)qi * A.tile_ld + sub] = off;
    uint64_t best = dkey(bd);                                     // dkey(-inf) is a valid (smallest real) key
    if (n_rec < n_valid) {                                        // some doc keeps the default
        const uint64_t ku = dkey(untouched);
        best = ku > best ? ku : best;
    }
    best = warp_max_u64_redux(best);
    if (lane == 0 && best > *(volatile uint64_t*)&A.max_keys[qi])
        atomicMax(reinterpret_cast<unsigned long long*>(&A.max_keys[qi]), (unsigned long long)best);
}

// ---- the bitmap path (tf == 1 indexes) ---------------------------------------------------------------------------
// One presence bitmap per DISTINCT term of the batch, in the lane-major byte layout of finals.cuh (BitSrc): built by
// streaming every term's posting list once - coalesced, no per-(tile, query) dependency chain - and shared by all the
// queries of the batch that use the term.  grid (chunks, n_slots); the bitmaps are zeroed by the caller.
constexpr int BITMAP_THREADS = 256;
__global__ void __launch_bounds__(BITMAP_THREADS)
bm25_bitmap_kernel(const int64_t* __restrict__ post_ptr, const int32_t* __restrict__ post_doc, const int32_t* __restrict__ slot_terms,
                   int64_t n_tiles, uint32_t* __restrict__ bits /* [slot][n_tiles][8 words] */) {
    const int s = blockIdx.y, lane = threadIdx.x & 31;
    const int t = slot_terms[s];
    const int64_t a = post_ptr[t], b = post_ptr[t + 1];
    uint32_t* mine = bits + (int64_t)s * n_tiles * 8;
    const int64_t stride = (int64_t)gridDim.x * BITMAP_THREADS;
    for (int64_t p0 = a + (int64_t)blockIdx.x * BITMAP_THREADS + (threadIdx.x - lane); p0 < b; p0 += stride) {   // warp-uniform
        const int64_t p = p0 + lane;
        const bool act = p < b;
        int64_t word = -1 - lane;                                   // inactive lanes: distinct, matching nobody
        uint32_t bit = 0u;
        if (act) {
            const int d = post_doc[p];
            const int l = d & 31, u = (d >> 5) & 7;                 // doc = tile*256 + 32*u + l -> byte l, bit u of the tile
            word = ((int64_t)(d >> 8) << 3) + (l >> 2);
            bit = 1u << (8 * (l & 3) + u);
        }
        // doc ids ascend, so lanes that hit the same 32-bit word are neighbours: one atomic per distinct word
        const unsigned peers = __match_any_sync(0xffffffffu, word);
        const uint32_t all = __reduce_or_sync(peers, bit);
        if (act && lane == __ffs(peers) - 1) atomicOr(&mine[word], all);
    }
}

// q_idf[q][j] = idf of query q's j-th term (0 if absent, webui.py:140), n_required[q]: the per-query constants of the
// bitmap path (the general path fills them in bm25_slices_kernel)
__global__ void query_idf_kernel(const QueryTerms* __restrict__ queries, int nq, const double* __restrict__ idf, int32_t n_vocab,
                                 double* __restrict__ q_idf, int32_t* __restrict__ n_required) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * MAX_TERMS) return;
    const int qi = i / MAX_TERMS, j = i - qi * MAX_TERMS;
    if (j == 0) n_required[qi] = queries[qi].n_required;
    if (j >= queries[qi].n_terms) return;
    const int t = queries[qi].term[j];
    q_idf[i] = (t >= 0 && t < n_vocab) ? idf[t] : 0.0;
}

// per-query maximum of the BM25 scores (webui.py:379 needs it before anything can be combined); one block = one tile

// x 8 queries, so the tile's g1 values and the bytes of shared terms come out of L1 for seven of the eight warps
constexpr int BITQ_WARPS = 8;
__global__ void __launch_bounds__(32 * BITQ_WARPS)
bm25_max_bits_kernel(BitSrc B, int64_t n, int nq, uint64_t* __restrict__ max_keys, double* __restrict__ dense_out, int64_t ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const int qi = blockIdx.y * BITQ_WARPS + warp;
    if (qi >= nq) return;
    double v[FIN_U];
    bm25_tile_values(B, qi, tile, lane, n, v);
    double bd = -INFINITY;
    bool any = false;
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int64_t d = tile * FIN_TILE + 32 * u + lane;
        if (d < n) {
            if (dense_out) dense_out[(int64_t)qi * ld + d] = v[u];
            bd = fmax(bd, v[u]);
            any = true;
        }
    }
    uint64_t best = any ? dkey(bd) : KEY_EMPTY;
    {   // 64-bit warp maximum with two redux.sync
        const uint32_t bh = (uint32_t)(best >> 32), bl = (uint32_t)best;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, bh);
        const uint32_t ml = __reduce_max_sync(0xffffffffu, bh == mh ? bl : 0u);
        best = ((uint64_t)mh << 32) | ml;
    }
    if (lane == 0 && best != KEY_EMPTY && best > *(volatile uint64_t*)&max_keys[qi])
        atomicMax(reinterpret_cast<unsigned long long*>(&max_keys[qi]), (unsigned long long)best);
}

constexpr int BM25C_WARPS = 8;
constexpr int BM25C_THREADS = 32 * BM25C_WARPS;
constexpr uint64_t KEY_NAN = 0xFFF8000000000000ull;     // a NaN score sorts first, as dkey(NaN) does

struct CombineArgs {
    FinSrc S;
    int64_t n_sub;
    uint64_t* seg_max; int seg_stride; int tiles_per_seg;     // seg_max[q * seg_stride + tile / tiles_per_seg]
    uint64_t* tile_max;                                        // tile_max[q][tile]: best combined key of the tile
};

// After the global maxima are known.  One warp per (tile, query): the best combined score of the tile
//   final = wb * bm25 / max bm25 + wd * (sim / max sim)         (webui.py:376-383)
// and nothing else - no per-doc store.  Docs WITHOUT a BM25 record share the query's default BM25 value, and
// RN(x / max), RN(wd * .), the widening and RN(default + .) are all monotone in x, so their best combined score is that
// of the largest (wd >= 0) dot score among them: one FMNMX per doc.  Docs WITH a record (~45 of 256) take the exact
// formula, one doc per lane, reading the compacted record (value + position) - the dot score comes from the tile's 1 KB
// the warp has just pulled through L1.
template <int MINB>
__global__ void __launch_bounds__(BM25C_THREADS, MINB)
bm25_combine_kernel(CombineArgs A) {
    const FinSrc& S = A.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = blockIdx.x * BM25C_WARPS + warp;               // 32-bit tile index: n_sub < 2^23 (shards hold < 2^31 docs)
    if (sub >= (int)A.n_sub) return;
    const int qi = blockIdx.y;
    const int64_t lo = (int64_t)sub * BM25_SUB;
    const int n_valid = (int)((lo + BM25_SUB < S.n ? lo + BM25_SUB : S.n) - lo);
    const float* simt = S.sim + (int64_t)qi * S.ld + lo;           // the tile's dot scores
    const int64_t tq = (int64_t)qi * S.tile_ld + sub;

    float sv[BM25_SUB / 32];
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) sv[u] = u * 32 + lane < n_valid ? simt[u * 32 + lane] : 0.0f;
    const uint4* hp = reinterpret_cast<const uint4*>(S.tile_hdr + tq * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    const int64_t rbase = S.rec_base[qi] + S.tile_off[tq];
    const QNorm c = S.qnorm(qi);
    const uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    int n_rec = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) n_rec += __popc(w[u]);
    // the first 64 records are fetched before anything depends on them
    const double* rvp = S.rec_val + rbase;
    const uint8_t* rpp = S.rec_pos + rbase;
    double pv[2];
    int pp[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = 32 * r + lane;
        pv[r] = i < n_rec ? rvp[i] : 0.0;
        pp[r] = i < n_rec ? (int)rpp[i] : 0;
    }

    // docs without a record: the extreme dot score among them (NaN dot scores - NaN rows - are caught by x != x)
    const bool up = !(S.wd < 0.0f);
    float ext = up ? -INFINITY : INFINITY;
    bool nan_seen = false;
#pragma unroll
    for (int u = 0; u < BM25_SUB / 32; ++u) {
        const bool plain = !((w[u] >> lane) & 1u) && u * 32 + lane < n_valid;
        const float x = plain ? sv[u] : ext;
        nan_seen = nan_seen || x != x;
        ext = up ? fmaxf(ext, x) : fminf(ext, x);
    }
    double fbest = -INFINITY;
    bool any = false;
    if (n_rec < n_valid) {                                         // warp-uniform
        // order-preserving integer image: one redux per direction instead of five shuffle rounds
        const uint32_t k = fkey(ext);
        const uint32_t kk = up ? __reduce_max_sync(0xffffffffu, k) : __reduce_min_sync(0xffffffffu, k);
        fbest = S.blend(c.wb_dflt, S.sim_norm(c, fkey_inv(kk)));                      // webui.py:383
        nan_seen = nan_seen || fbest != fbest;
        any = true;
    }
    // docs with a record: exact, one per lane
    auto exact = [&](double val, int pos) {
        const double f = S.blend(__dmul_rn(S.wb, S.bm25_norm(c, val)), S.sim_norm(c, simt[pos]));
        nan_seen = nan_seen || f != f;
        fbest = f > fbest ? f : fbest;
        any = true;
    };
    if (lane < n_rec) exact(pv[0], pp[0]);
    if (32 + lane < n_rec) exact(pv[1], pp[1]);
    for (int i = 64 + lane; i < n_rec; i += 32) exact(rvp[i], (int)rpp[i]);
    uint64_t best = any ? dkey(fbest) : KEY_EMPTY;
    if (nan_seen) best = KEY_NAN;                                  // e.g. weight 0 x -inf
    best = warp_max_u64_redux(best);
    if (lane == 0) {
        A.tile_max[tq] = best;
        if (best != KEY_EMPTY)
            atomicMax(reinterpret_cast<unsigned long long*>(&A.seg_max[(size_t)qi * A.seg_stride + sub / A.tiles_per_seg]),
                      (unsigned long long)best);
    }
}

// The same stage on the bitmap path: the tile's BM25 values are re-derived from the term bitmaps (no records exist),
// every doc takes the exact formula; block = one tile x 8 queries (shared g1 / bitmap bytes through L1).
__global__ void __launch_bounds__(32 * BITQ_WARPS)
bm25_combine_bits_kernel(CombineArgs A, int nq) {
    const FinSrc& S = A.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t sub = blockIdx.x;
    const int qi = blockIdx.y * BITQ_WARPS + warp;
    if (qi >= nq) return;
    double f[FIN_U];
    unsigned valid;
    tile_finals(S, qi, sub, lane, f, valid);
    double fbest = -INFINITY;
    bool nan_seen = false;
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        if ((valid >> u) & 1u) {
            nan_seen = nan_seen || f[u] != f[u];
            fbest = fmax(fbest, f[u]);
        }
    }
    uint64_t best = valid ? dkey(fbest) : KEY_EMPTY;
    if (nan_seen) best = KEY_NAN;
    best = warp_max_u64(best);
    if (lane == 0) {
        A.tile_max[(int64_t)qi * S.tile_ld + sub] = best;
        if (best != KEY_EMPTY)
            atomicMax(reinterpret_cast<unsigned long long*>(&A.seg_max[(size_t)qi * A.seg_stride + (int)(sub / A.tiles_per_seg)]),
                      (unsigned long long)best);
    }
}

}  // namespace ais
