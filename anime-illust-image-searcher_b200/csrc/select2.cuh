// Streaming threshold select (the fast path of the top-k stages) and the near-tie witness pass.
//
// Exact top-k of N keys without a sort and without a stored score array:
//   1. tile / segment maxima: the best combined key of every 256-doc tile (pass 1: written by bm25_combine_kernel;
//      supplied scores and the dense re-query: segmax_kernel) and of every segment of whole tiles (<= 2048 per query).
//   2. threshold: pass 1: T = k-th largest segment maximum - at least k docs (one per such segment) have key >= T, so
//      the top-k is contained in {key >= T}; with docs spread over the segments |{key >= T}| ~ -S ln(1 - k/S).
//      Pass 2 of the reference's (collapsed) PRF re-query: T = k-th largest blend R among this shard's pass-1
//      candidates (rerank_threshold_kernel) - again at least k docs reach it.
//   3. collect: one lane per tile decides from 8-16 bytes whether the tile can hold a key >= T (pass 2, column mode: an
//      UPPER BOUND of R over the tile from the tile's best combined score and the tile's extreme column-0 values - the
//      blend is monotone in both); the warp recomputes the combined scores of the (few hundred of 39 063) tiles that
//      can (finals.cuh) and appends every doc with key >= T to a survivor list (warp-aggregated atomics); then ONE block
//      sorts the survivors (key desc, doc id asc) and writes the top-k.
// If the survivors exceed SURV_CAP (top docs clustered in few segments) the query's gate flag is raised
// and the streaming buffer select (stream_select_kernel), launched right behind and gated on that flag, redoes it.
//
// The witness pass settles filter_searched_result's "is there a second near-tie anywhere?" question
//  without sorting: two DISTINCT scores closer than DIFF_FILTER_THRESH anywhere below the
// returned prefix imply an adjacent near-tie there, and two such scores falling into one bucket of width
// thresh are found with a single atomicMax per doc.
#pragma once
#include "finals.cuh"
#include "select.cuh"

namespace ais {

constexpr int SEG_MAX = 2048;            // segments per query
constexpr int SEG_WARPS = 8;             // warps (= segments) per block
constexpr int SURV_CAP = 4096;
constexpr int COLLECT_THREADS = 256;
constexpr int SEL_TILE = FIN_TILE;       // docs per tile of the tile-maximum table

// pass 2: webui.py:208 blend of the combined score with the re-query score
struct Rerank {
    const float* rer;           // this query's re-query scores [n], or the shared column buffer
    float cq;                   // 1, or the query's scalar (column mode): rer[i] * cq is RN-exact either way
    CombineParams cp;
    __device__ __forceinline__ double blend(double fin, float rr) const {
        return __dadd_rn(__dmul_rn(cp.wo, fin), (double)__fmul_rn(cp.wr, __fmul_rn(rr, cq)));
    }
};

struct SelectArgs {        // everything the kernels share
    FinSrc S;
    const float* rer; int64_t rer_qstride;   // pass 2: re-query scores; qstride ld, or 0 when `rer` is the shared column buffer
    const float* rer_scale;                  // null (scale 1) or the queries' scalars [q * DIM]
    int64_t n, id_base;
    CombineParams cp;
    const int64_t* seeds_all;   // [nq][MAX_DEPTH] (mode 2)
    int depth;
    int n_seg; int tiles_per_seg; int64_t n_tiles;
    uint64_t* seg_max;          // [nq][SEG_MAX]
    uint64_t* tile_max;         // [nq][tile_ld]: best key of every tile, seeds included (an upper bound: skip filter only)
    const float* col_lo; const float* col_hi;    // [n_tiles] extreme column values per tile (pass 2, column mode) or null
    uint64_t* max_all;          // [nq] atomicMax over ALL docs (mode 2: max R) or null
    uint64_t* thr;              // [nq]
    int* surv_count;            // [nq]
    uint64_t* surv_keys;        // [nq][SURV_CAP]
    int64_t* surv_ids;
    int* gate;                  // [nq] raised when the survivors overflow

    __device__ __forceinline__ Rerank rerank(int qi) const {
        return Rerank{rer + (size_t)qi * rer_qstride, rer_scale ? rer_scale[(size_t)qi * DIM] : 1.0f, cp};
    }
};

// scores of one tile for the select kernels: MODE 1 the combined scores, MODE 2 the blend R
template <int MODE>
__device__ __forceinline__ void tile_scores(const SelectArgs& a, const Rerank& rk, int qi, int64_t tile, int lane,
                                            double (&f)[FIN_U], unsigned& valid) {
    tile_finals(a.S, qi, tile, lane, f, valid);
    if (MODE == 2) {
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const int64_t d = tile * SEL_TILE + 32 * u + lane;
            if ((valid >> u) & 1u) f[u] = rk.blend(f[u], rk.rer[d]);
        }
    }
}

// ---- 1. segment maxima (+ tile maxima) by streaming --------------------------------------------------------
// Used where no tile table exists yet: supplied scores (ais_rerank, MODE 1) and the dense re-query (MODE 2).  The stream
// runs in the double domain (one fmax per doc); keys are formed once per tile / segment.  A NaN score (only possible
// with a zero weight times -inf, or NaN rows) sorts first like dkey(NaN) does: it is tracked by a flag.
template <int MODE>
__global__ void __launch_bounds__(32 * SEG_WARPS)
segmax_kernel(SelectArgs a, int nq) {
    // block = one segment x SEG_WARPS queries (warp = query): the per-doc data the queries share (g1, bitmap bytes of
    // common terms, the re-query column) is pulled through L1 once for the eight of them
    __shared__ int64_t seeds_all[SEG_WARPS][MAX_DEPTH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.y * SEG_WARPS + warp;
    const int seg = blockIdx.x;
    if (qi >= nq) return;
    int64_t* seeds = seeds_all[warp];
    if (MODE == 2) {
        if (lane < MAX_DEPTH) seeds[lane] = lane < a.depth ? a.seeds_all[qi * MAX_DEPTH + lane] : -1;
        __syncwarp();
    }
    uint64_t all_best = KEY_EMPTY;
    {
        const Rerank rk = a.rerank(qi);
        const int64_t t0 = (int64_t)seg * a.tiles_per_seg;
        const int64_t t1 = t0 + a.tiles_per_seg < a.n_tiles ? t0 + a.tiles_per_seg : a.n_tiles;
        uint64_t* tmax = a.tile_max + (int64_t)qi * a.S.tile_ld;
        double bestd = -INFINITY;              // best non-seed score of the segment so far (valid once `has`)
        bool has = false, best_nan = false;
#pragma unroll 1
        for (int64_t tile = t0; tile < t1; ++tile) {
            const int64_t lo = tile * SEL_TILE;
            const int64_t hi = lo + SEL_TILE < a.n ? lo + SEL_TILE : a.n;
            double f[FIN_U];
            unsigned valid;
            tile_scores<MODE>(a, rk, qi, tile, lane, f, valid);
            double tb = -INFINITY;
            bool tnan = false;
            // the PRF seeds are not candidates (webui.py:217); at most `depth` of the shard's tiles hold one
            const bool seeded = MODE == 2 && __any_sync(0xffffffffu, lane < a.depth && seeds[lane] >= a.id_base + lo &&
                                                                         seeds[lane] < a.id_base + hi);
#pragma unroll
            for (int u = 0; u < FIN_U; ++u) {
                if (!((valid >> u) & 1u)) continue;
                const double r = f[u];
                const bool isn = r != r;
                tnan = tnan || isn;
                tb = fmax(tb, r);
                bool cand = true;
                if (seeded) {
                    const int64_t id = a.id_base + lo + 32 * u + lane;
                    for (int t = 0; t < a.depth; ++t) cand = cand && seeds[t] != id;
                }
                if (cand) {
                    best_nan = best_nan || isn;
                    bestd = fmax(bestd, r);
                    has = true;
                }
            }
            uint64_t tbest = tnan ? KEY_NAN : (valid ? dkey(tb) : KEY_EMPTY);
            tbest = warp_max_u64_redux(tbest);
            if (lane == 0) tmax[tile] = tbest;
            all_best = tbest > all_best ? tbest : all_best;
        }
        uint64_t best = best_nan ? KEY_NAN : (has ? dkey(bestd) : KEY_EMPTY);
        best = warp_max_u64(best);
        if (lane == 0) a.seg_max[(size_t)qi * SEG_MAX + seg] = best;
    }
    if (a.max_all && lane == 0 && all_best != KEY_EMPTY && all_best > *(volatile uint64_t*)&a.max_all[qi])
        atomicMax(reinterpret_cast<unsigned long long*>(&a.max_all[qi]), (unsigned long long)all_best);
}

// ---- records path: the scores of one tile in two classes ----------------------------------------------------------
// Docs WITHOUT a BM25 record share the query's default BM25 value: their combined score needs the dot score only (~12
// instructions per doc, straight from registers).  The ~45 docs WITH a record are handled compactly, one per lane,
// through the stored (value, position) pairs - no per-doc bitmap-prefix lookup, no dynamic register indexing.
// fn(valid, score, l) is called warp-uniformly (every lane, every round): l = doc index inside the tile.
// MODE 1: combined score (webui.py:376-383); MODE 2: the blend R with the re-query score (webui.py:208).
template <int MODE, typename F>
__device__ __forceinline__ void tile_two_class(const SelectArgs& a, const Rerank& rk, int qi, int64_t tile, int lane, F&& fn) {
    const FinSrc& S = a.S;
    const int64_t lo = tile * SEL_TILE;
    const int64_t hi = lo + SEL_TILE < a.n ? lo + SEL_TILE : a.n;
    const float* simq = S.sim + (int64_t)qi * S.ld;
    float sv[FIN_U], rv[FIN_U];
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int64_t d = lo + u * 32 + lane;
        sv[u] = d < hi ? simq[d] : 0.0f;
        rv[u] = (MODE == 2 && d < hi) ? rk.rer[d] : 0.0f;
    }
    const uint4* hp = reinterpret_cast<const uint4*>(S.tile_hdr + ((int64_t)qi * S.tile_ld + tile) * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    const int64_t rbase = S.rec_base[qi] + S.tile_off[(int64_t)qi * S.tile_ld + tile];
    const QNorm c = S.qnorm(qi);
    const uint32_t w[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    int n_rec = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) n_rec += __popc(w[u]);
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int l = u * 32 + lane;
        const bool plain = !((w[u] >> lane) & 1u) && lo + l < hi;
        double f = S.blend(c.wb_dflt, S.sim_norm(c, sv[u]));
        if (MODE == 2) f = rk.blend(f, rv[u]);
        fn(plain, f, l);
    }
    for (int i0 = 0; i0 < n_rec; i0 += 32) {
        const int i = i0 + lane;
        const bool live = i < n_rec;
        const double val = live ? S.rec_val[rbase + i] : 0.0;
        const int pos = live ? (int)S.rec_pos[rbase + i] : 0;
        double f = S.blend(__dmul_rn(S.wb, S.bm25_norm(c, val)), S.sim_norm(c, simq[lo + pos]));
        if (MODE == 2) f = rk.blend(f, rk.rer[lo + pos]);
        fn(live, f, pos);
    }
}

// ---- 1b. pass 2 on the records path: tile / segment maxima of the blend R in one light pass ---------------------------
// R = wo * final + wr * rer (webui.py:208) for every doc of the tiles that matter, from the dot score (4 B), the re-query
// score (the shared column or the query's dense array, 4 B) and the tile's BM25 records.  A warp owns 32 consecutive
// tiles of one query (block = 8 queries over the same tiles: the column comes through L1 once for the eight).  The PRF
// seeds are NOT excluded here (webui.py:217 drops them from the candidates): the caller asks the segment-maximum
// threshold for `depth` more segments instead, which keeps it a valid lower bound - at most `depth` maxima are seeds.
// BOUND = 1 (the reference's collapsed re-query with non-negative blend weights): a.thr holds a lower bound T of the
// k-th best R of this shard (rerank_threshold_kernel).  ONE LANE per tile forms the tile's upper bound - the blend of its
// best COMBINED score (pass 1's tile table) with its extreme column value, every rounding monotone; a tile below T holds
// no candidate and not the maximum of R either and costs 16 bytes.  Measured on the benchmark: 12 % of the tiles remain.
template <int BOUND>
__global__ void __launch_bounds__(32 * SEG_WARPS)
rerank_max_kernel(SelectArgs a, int nq, const uint64_t* __restrict__ tile_max_fin) {
    const FinSrc& S = a.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tb = (int64_t)blockIdx.x * 32;
    const int qi = blockIdx.y * SEG_WARPS + warp;
    if (qi >= nq) return;
    const Rerank rk = a.rerank(qi);
    const int64_t my_tile = tb + lane;
    bool visit = my_tile < a.n_tiles;
    if (BOUND == 1 && visit) {
        uint64_t bound = tile_max_fin[(int64_t)qi * S.tile_ld + my_tile];
        if (bound != KEY_EMPTY && bound < KEY_NAN) {
            const float cx = rk.cq < 0.0f ? a.col_lo[my_tile] : a.col_hi[my_tile];
            bound = dkey(rk.blend(dkey_inv(bound), cx));
        }
        visit = bound >= a.thr[qi];
        if (!visit) a.tile_max[(int64_t)qi * S.tile_ld + my_tile] = KEY_EMPTY;
    }
    unsigned todo = __ballot_sync(0xffffffffu, visit);
    uint64_t all_best = KEY_EMPTY;
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t tile = tb + b;
        double best = -INFINITY;
        bool any = false, nan_seen = false;
        tile_two_class<2>(a, rk, qi, tile, lane, [&](bool valid, double r, int) {
            if (valid) {
                nan_seen = nan_seen || r != r;
                best = fmax(best, r);
                any = true;
            }
        });
        uint64_t key = nan_seen ? KEY_NAN : (any ? dkey(best) : KEY_EMPTY);
        key = warp_max_u64_redux(key);
        all_best = key > all_best ? key : all_best;
        if (lane == 0) {
            a.tile_max[(int64_t)qi * S.tile_ld + tile] = key;
            if (key != KEY_EMPTY)
                atomicMax(reinterpret_cast<unsigned long long*>(&a.seg_max[(size_t)qi * SEG_MAX + (int)(tile / a.tiles_per_seg)]),
                          (unsigned long long)key);
        }
    }
    if (a.max_all && lane == 0 && all_best != KEY_EMPTY && all_best > *(volatile uint64_t*)&a.max_all[qi])
        atomicMax(reinterpret_cast<unsigned long long*>(&a.max_all[qi]), (unsigned long long)all_best);
}

// ---- 2a. threshold = k-th largest segment maximum ------------------------------------------------------
__global__ void __launch_bounds__(256)
threshold_kernel(const uint64_t* __restrict__ seg_max, int n_seg, int k, uint64_t* __restrict__ thr,
                 int* __restrict__ surv_count, int* __restrict__ gate, int use_floor) {
    __shared__ uint64_t v[SEG_MAX];
    const int qi = blockIdx.x, tid = threadIdx.x;
    int P = 32;
    while (P < n_seg) P <<= 1;
    for (int i = tid; i < P; i += 256) v[i] = i < n_seg ? seg_max[(size_t)qi * SEG_MAX + i] : KEY_EMPTY;
    __syncthreads();
    for (unsigned size = 2; size <= (unsigned)P; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = tid; t < (unsigned)P / 2; t += 256) {
                const unsigned i = 2 * t - (t & (stride - 1));
                const unsigned l = i + stride;
                const bool up = ((i & size) == 0);
                const uint64_t x = v[i], y = v[l];
                if ((y > x) == up) { v[i] = y; v[l] = x; }       // descending
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        uint64_t t = (k <= P) ? v[k - 1] : KEY_EMPTY;
        t = t != KEY_EMPTY ? t : 1ull;                // fewer than k non-empty segments: every doc qualifies
        // use_floor: thr[qi] already holds ANOTHER valid lower bound of the k-th best key (rerank_threshold_kernel): keep the tighter
        if (use_floor && thr[qi] > t && thr[qi] < KEY_NAN) t = thr[qi];
        thr[qi] = t;
        surv_count[qi] = 0;
        gate[qi] = 0;
    }
}

// ---- 2b. pass 2, column mode: threshold = k-th largest blend R among this shard's pass-1 candidates -------------
// cand_*: the shard's best `k1` docs by combined score (sorted, KEY_EMPTY padded; the key IS the exact score).  The
// global PRF seeds are not candidates (webui.py:217).  At least k of the shard's docs reach the threshold, so the
// shard's top-k by R is contained in {R >= T}; with fewer than k candidates every doc qualifies (T = 1).
constexpr int RTHR_CAP = 1024;           // = SEL_KMAX
__global__ void __launch_bounds__(256)
rerank_threshold_kernel(const uint64_t* __restrict__ cand_keys, const int64_t* __restrict__ cand_ids, int k1, SelectArgs a, int k) {
    __shared__ uint64_t v[RTHR_CAP];
    __shared__ int64_t seeds[MAX_DEPTH];
    const int qi = blockIdx.x, tid = threadIdx.x;
    if (tid < MAX_DEPTH) seeds[tid] = tid < a.depth ? a.seeds_all[qi * MAX_DEPTH + tid] : -1;
    __syncthreads();
    const Rerank rk = a.rerank(qi);
    int P = 32;
    while (P < k1) P <<= 1;
    for (int i = tid; i < P; i += 256) {
        uint64_t key = KEY_EMPTY;
        if (i < k1) {
            const uint64_t ck = cand_keys[(size_t)qi * k1 + i];
            const int64_t id = cand_ids[(size_t)qi * k1 + i];
            bool ok = ck != KEY_EMPTY;
            for (int t = 0; t < MAX_DEPTH; ++t) ok = ok && seeds[t] != id;
            if (ok) key = dkey(rk.blend(dkey_inv(ck), rk.rer[id - a.id_base]));
        }
        v[i] = key;
    }
    __syncthreads();
    for (unsigned size = 2; size <= (unsigned)P; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = tid; t < (unsigned)P / 2; t += 256) {
                const unsigned i = 2 * t - (t & (stride - 1));
                const unsigned l = i + stride;
                const bool up = ((i & size) == 0);
                const uint64_t x = v[i], y = v[l];
                if ((y > x) == up) { v[i] = y; v[l] = x; }       // descending
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        uint64_t t = (k <= P) ? v[k - 1] : KEY_EMPTY;
        // a NaN key at the cut would be an unusable bound (NaN compares above everything): take every doc then
        a.thr[qi] = (t != KEY_EMPTY && t < KEY_NAN) ? t : 1ull;
        a.surv_count[qi] = 0;
        a.gate[qi] = 0;
    }
}

// ---- 3a. collect the survivors -------------------------------------------------------------------------
// One lane per tile decides whether the tile can reach the threshold; the warp then walks the (rare) tiles that do.
// BOUND = 0: tile_max holds the tile's best key of THIS pass.  BOUND = 1 (pass 2, column mode): tile_max holds the best
// COMBINED key (pass 1) and the bound on R = wo * fin + wr * (col * c) comes from the tile's extreme column values:
// with wo, wr >= 0 (checked by the host) every rounding in the blend is monotone, so evaluating it at
// (max fin, extreme col * c) bounds every doc of the tile from above.  Visited tiles also feed the running max of R
// (webui.py:209 takes it over ALL docs): the doc that attains it reaches the threshold, so its tile is visited.
template <int MODE, int BOUND>
__global__ void __launch_bounds__(COLLECT_THREADS)
collect_kernel(SelectArgs a) {
    __shared__ int64_t seeds[MAX_DEPTH];
    const int qi = blockIdx.y, lane = threadIdx.x & 31;
    if (MODE == 2) {
        if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = threadIdx.x < a.depth ? a.seeds_all[qi * MAX_DEPTH + threadIdx.x] : -1;
        __syncthreads();
    }
    const uint64_t T = a.thr[qi];
    int* cnt = a.surv_count + qi;
    uint64_t* sk = a.surv_keys + (size_t)qi * SURV_CAP;
    int64_t* si = a.surv_ids + (size_t)qi * SURV_CAP;
    const uint64_t* tmax = a.tile_max + (int64_t)qi * a.S.tile_ld;
    const Rerank rk = a.rerank(qi);
    const int64_t warp_id = ((int64_t)blockIdx.x * COLLECT_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * COLLECT_THREADS) >> 5;
    uint64_t all_best = KEY_EMPTY;
    for (int64_t tb = warp_id * 32; tb < a.n_tiles; tb += n_warps * 32) {
        const int64_t my_tile = tb + lane;
        bool visit = false;
        if (my_tile < a.n_tiles) {
            uint64_t bound = tmax[my_tile];
            if (BOUND == 1 && bound != KEY_EMPTY && bound < KEY_NAN) {
                const float cx = rk.cq < 0.0f ? a.col_lo[my_tile] : a.col_hi[my_tile];
                bound = dkey(rk.blend(dkey_inv(bound), cx));
            }
            visit = bound >= T;
        }
        unsigned todo = __ballot_sync(0xffffffffu, visit);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t tile = tb + b, lo = tile * SEL_TILE;
            double f[FIN_U];
            unsigned valid;
            tile_scores<MODE>(a, rk, qi, tile, lane, f, valid);
#pragma unroll 1
            for (int u = 0; u < FIN_U; ++u) {
                const int64_t i = lo + u * 32 + lane;
                const bool in = (valid >> u) & 1u;
                const uint64_t key = in ? dkey(f[u]) : KEY_EMPTY;
                if (BOUND == 1) all_best = key > all_best ? key : all_best;
                bool pass = in && key >= T;
                if (MODE == 2 && pass)
                    for (int t = 0; t < MAX_DEPTH; ++t) pass = pass && seeds[t] != a.id_base + i;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(cnt, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (pass) {
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < SURV_CAP) { sk[pos] = key; si[pos] = a.id_base + i; }
                    }
                }
            }
        }
    }
    if (BOUND == 1 && a.max_all) {
        all_best = warp_max_u64(all_best);
        if (lane == 0 && all_best != KEY_EMPTY)
            atomicMax(reinterpret_cast<unsigned long long*>(&a.max_all[qi]), (unsigned long long)all_best);
    }
}

// the same collect on the records path (no supplied scores, no bitmaps): tile_two_class instead of the generic per-doc
// evaluation; tile_max holds the tile's best key of THIS pass
template <int MODE>
__global__ void __launch_bounds__(COLLECT_THREADS)
collect_fast_kernel(SelectArgs a) {
    __shared__ int64_t seeds[MAX_DEPTH];
    const int qi = blockIdx.y, lane = threadIdx.x & 31;
    if (MODE == 2) {
        if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = threadIdx.x < a.depth ? a.seeds_all[qi * MAX_DEPTH + threadIdx.x] : -1;
        __syncthreads();
    }
    const uint64_t T = a.thr[qi];
    int* cnt = a.surv_count + qi;
    uint64_t* sk = a.surv_keys + (size_t)qi * SURV_CAP;
    int64_t* si = a.surv_ids + (size_t)qi * SURV_CAP;
    const uint64_t* tmax = a.tile_max + (int64_t)qi * a.S.tile_ld;
    const Rerank rk = a.rerank(qi);
    const int64_t warp_id = ((int64_t)blockIdx.x * COLLECT_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * COLLECT_THREADS) >> 5;
    for (int64_t tb = warp_id * 32; tb < a.n_tiles; tb += n_warps * 32) {
        const int64_t my_tile = tb + lane;
        unsigned todo = __ballot_sync(0xffffffffu, my_tile < a.n_tiles && tmax[my_tile] >= T);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t tile = tb + b, lo = tile * SEL_TILE;
            tile_two_class<MODE>(a, rk, qi, tile, lane, [&](bool valid, double f, int l) {
                const uint64_t key = valid ? dkey(f) : KEY_EMPTY;
                const int64_t id = a.id_base + lo + l;
                bool pass = valid && key >= T;
                if (MODE == 2 && pass)
                    for (int t = 0; t < MAX_DEPTH; ++t) pass = pass && seeds[t] != id;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(cnt, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (pass) {
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < SURV_CAP) { sk[pos] = key; si[pos] = id; }
                    }
                }
            });
        }
    }
}

// ---- 3b. sort the survivors, write the top-k ---------------------------------------------------------------
__global__ void __launch_bounds__(512)
sort_survivors_kernel(const int* __restrict__ surv_count, const uint64_t* __restrict__ surv_keys,
                      const int64_t* __restrict__ surv_ids, int k, uint64_t* __restrict__ out_keys,
                      int64_t* __restrict__ out_ids, int* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint64_t* key = reinterpret_cast<uint64_t*>(sm);
    int64_t* id = reinterpret_cast<int64_t*>(sm + (size_t)SURV_CAP * 8);
    const int qi = blockIdx.x, tid = threadIdx.x;
    const int cnt = surv_count[qi];
    if (cnt > SURV_CAP) {                       // overflow: the gated streaming select takes over
        if (tid == 0) gate[qi] = 1;
        return;
    }
    int P = 64;
    while (P < cnt) P <<= 1;
    for (int i = tid; i < P; i += 512) {
        key[i] = i < cnt ? surv_keys[(size_t)qi * SURV_CAP + i] : KEY_EMPTY;
        id[i] = i < cnt ? surv_ids[(size_t)qi * SURV_CAP + i] : ID_EMPTY;
    }
    __syncthreads();
    for (unsigned size = 2; size <= (unsigned)P; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = tid; t < (unsigned)P / 2; t += 512) {
                const unsigned i = 2 * t - (t & (stride - 1));
                const unsigned l = i + stride;
                const bool up = ((i & size) == 0);
                const uint64_t ka = key[i], kb = key[l];
                const int64_t ia = id[i], ib = id[l];
                if (better(kb, ib, ka, ia) == up) { key[i] = kb; id[i] = ib; key[l] = ka; id[l] = ia; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += 512) {
        const bool live = i < cnt && i < P;
        out_keys[(size_t)qi * k + i] = live ? key[i] : KEY_EMPTY;
        out_ids[(size_t)qi * k + i] = live ? id[i] : ID_EMPTY;
    }
}

// ---- 3c. gated fallback: streaming block top-k over recomputed scores ------------------------------------------
// grid (G, nq): block g streams a contiguous run of tiles through the shared-memory selector of select.cuh and writes
// its sorted top-k to cand[(q * G + g) * k ...]; merge_kernel joins the G lists.  MODE 2 also accumulates max R.
template <int MODE>
__global__ void __launch_bounds__(SEL_THREADS)
stream_select_kernel(SelectArgs a, int k, uint64_t* __restrict__ cand_keys, int64_t* __restrict__ cand_ids) {
    __shared__ SelBuf sb;
    __shared__ int64_t seeds[MAX_DEPTH];
    __shared__ uint64_t wscratch[SEL_THREADS / 32];
    const int qi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.gate && !a.gate[qi]) return;
    if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = (MODE == 2 && threadIdx.x < a.depth) ? a.seeds_all[qi * MAX_DEPTH + threadIdx.x] : -1;
    sel_init(sb);
    const Rerank rk = a.rerank(qi);
    const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per;
    const int64_t t1 = t0 + per < a.n_tiles ? t0 + per : a.n_tiles;
    constexpr int ROUND_TILES = SEL_ROUND / SEL_TILE;       // 4 tiles = SEL_ROUND offers per round: an offer never overflows
    uint64_t best = KEY_EMPTY;
    for (int64_t tr = t0; tr < t1; tr += ROUND_TILES) {
        __syncthreads();
        const int cnt_now = sb.count;
        __syncthreads();
        if (cnt_now + SEL_ROUND > SEL_CAP) sel_prune(sb, k);
        const int64_t tile = tr + warp;
        if (warp < ROUND_TILES && tile < t1) {              // warp-uniform: warps 0..3 score one tile each
            double f[FIN_U];
            unsigned valid;
            tile_scores<MODE>(a, rk, qi, tile, lane, f, valid);
#pragma unroll 1
            for (int u = 0; u < FIN_U; ++u) {
                const bool in = (valid >> u) & 1u;
                const int64_t id = a.id_base + tile * SEL_TILE + 32 * u + lane;
                const uint64_t key = in ? dkey(f[u]) : KEY_EMPTY;
                best = key > best ? key : best;
                bool ok = in;
                if (MODE == 2 && ok)
                    for (int t = 0; t < MAX_DEPTH; ++t) ok = ok && seeds[t] != id;
                sel_offer(sb, ok, key, id);
            }
        }
    }
    __syncthreads();
    sel_prune(sb, k);
    const size_t o = ((size_t)qi * gridDim.x + blockIdx.x) * (size_t)k;
    sel_write(sb, k, cand_keys + o, cand_ids + o);
    if (MODE == 2 && a.max_all) {
        __syncthreads();
        block_max_to_global(best, wscratch, &a.max_all[qi]);
    }
}

// ---- near-tie witness ------------------------------------------------------------------------------------
struct WitnessArgs {
    SelectArgs a; int qi;                     // sources of the query's scores
    int second_pass;                          // 1: value = normalised blend R (seeds skipped); 0: combined score
    uint64_t last_key;                        // key of the last entry of the sorted prefix
    const double* max_r; int normalize;       // device scalar (nullable)
    double thresh, inv_thresh;
    uint64_t* table; int64_t n_buckets;
    int* flag;
};

__device__ __forceinline__ double witness_value(int normalize, double max_r, double raw) {
    return (normalize && max_r > 0.0) ? __ddiv_rn(raw, max_r) : raw;             // webui.py:210-211
}

__global__ void __launch_bounds__(256)
witness_kernel(WitnessArgs w) {
    __shared__ int64_t seeds[MAX_DEPTH];
    const SelectArgs& a = w.a;
    if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = (w.second_pass && threadIdx.x < a.depth) ? a.seeds_all[w.qi * MAX_DEPTH + threadIdx.x] : -1;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const double max_r = w.max_r ? *w.max_r : 0.0;
    const Rerank rk = a.rerank(w.qi);
    const double v_last = w.last_key == ~0ull ? 1.0 : witness_value(w.normalize, max_r, dkey_inv(w.last_key));
    const uint64_t neg_inf = dkey(-INFINITY);
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = warp_id; tile < a.n_tiles; tile += n_warps) {
        if (*((volatile int*)w.flag)) return;                   // somebody found one
        double f[FIN_U];
        unsigned valid;
        if (w.second_pass) tile_scores<2>(a, rk, w.qi, tile, lane, f, valid);
        else tile_scores<1>(a, rk, w.qi, tile, lane, f, valid);
#pragma unroll 1
        for (int u = 0; u < FIN_U; ++u) {
            if (!((valid >> u) & 1u)) continue;
            const double raw = f[u];
            const uint64_t key = dkey(raw);
            if (key > w.last_key || key <= neg_inf) continue;       // inside the prefix region / masked
            const int64_t id = a.id_base + tile * SEL_TILE + 32 * u + lane;
            bool seed = false;
            for (int t = 0; t < MAX_DEPTH; ++t) seed = seed || (seeds[t] == id);
            if (seed) continue;
            const double v = witness_value(w.normalize, max_r, raw);
            const double off = (v_last - v) * w.inv_thresh;
            if (!(off >= 0.0) || off >= (double)w.n_buckets) continue;
            const uint64_t vk = dkey(v);
            const uint64_t old = atomicMax(reinterpret_cast<unsigned long long*>(&w.table[(int64_t)off]), (unsigned long long)vk);
            if (old != KEY_EMPTY && old != vk) {
                const double o = dkey_inv(old);
                const double d = o > v ? __dsub_rn(o, v) : __dsub_rn(v, o);
                if (d != 0.0 && d < w.thresh) atomicExch(w.flag, 1);
            }
        }
    }
}

// ---- exact last resort / seams: materialise keys or scores of one query ----------------------------------------
// keys[i] / ids[i] for every doc i of the shard (second_pass: R with the PRF seeds blanked); n_pad >= n slots
__global__ void __launch_bounds__(256)
fill_keys_kernel(SelectArgs a, int qi, int second_pass, uint64_t* __restrict__ keys, int64_t* __restrict__ ids) {
    __shared__ int64_t seeds[MAX_DEPTH];
    if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = (second_pass && threadIdx.x < a.depth) ? a.seeds_all[qi * MAX_DEPTH + threadIdx.x] : -1;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const Rerank rk = a.rerank(qi);
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = warp_id; tile < a.n_tiles; tile += n_warps) {
        double f[FIN_U];
        unsigned valid;
        if (second_pass) tile_scores<2>(a, rk, qi, tile, lane, f, valid);
        else tile_scores<1>(a, rk, qi, tile, lane, f, valid);
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            if (!((valid >> u) & 1u)) continue;
            const int64_t i = tile * SEL_TILE + 32 * u + lane;
            uint64_t key = dkey(f[u]);
            int64_t id = a.id_base + i;
            for (int t = 0; t < MAX_DEPTH; ++t)
                if (seeds[t] == id) { key = KEY_EMPTY; id = ID_EMPTY; }
            keys[i] = key;
            ids[i] = id;
        }
    }
}

// out[i] = combined score of doc i (ais_final_scores / ais_debug_read seam)
__global__ void __launch_bounds__(256)
materialize_finals_kernel(FinSrc S, int qi, int64_t n_tiles, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = warp_id; tile < n_tiles; tile += n_warps) {
        double f[FIN_U];
        unsigned valid;
        tile_finals(S, qi, tile, lane, f, valid);
#pragma unroll
        for (int u = 0; u < FIN_U; ++u)
            if ((valid >> u) & 1u) out[tile * FIN_TILE + 32 * u + lane] = f[u];
    }
}

// per-tile extremes of one column of the row store (static per index; pass 2, column mode)
__global__ void __launch_bounds__(256)
column_bounds_kernel(const float* __restrict__ col, int64_t n, int64_t n_tiles, float* __restrict__ col_lo, float* __restrict__ col_hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = warp_id; tile < n_tiles; tile += n_warps) {
        float lo = INFINITY, hi = -INFINITY;
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const int64_t d = tile * FIN_TILE + 32 * u + lane;
            if (d < n) { const float x = col[d]; lo = fminf(lo, x); hi = fmaxf(hi, x); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { col_lo[tile] = lo; col_hi[tile] = hi; }
    }
}

}  // namespace ais
