// Streaming threshold select (the fast path of the top-k stages) and the near-tie witness pass.
//
// Exact top-k of N keys in three light passes, none of which synchronises inside its loop:
//   1. segmax:    the docs are cut into S <= 2048 segments; one warp per segment streams its docs and records
//                 the segment's best key (pass 1: done by the BM25 tile kernel's combine phase, bm25.cuh).
//   2. threshold: T = k-th largest segment maximum.  At least k docs (one per such segment) have
//                 key >= T, so the global top-k is contained in {key >= T}; with docs spread over the
//                 segments |{key >= T}| ~ -S ln(1 - k/S), i.e. barely more than k.
//   3. collect:   passes 1 also record the best key of every 256-doc tile; the collect pass reads those
//                 (8 B per 256 docs) and visits only tiles whose best key reaches T - a few hundred of the
//                 39 063 tiles of a 10 M-doc shard - appending every doc with key >= T to a survivor list
//                 (warp-aggregated atomics); then ONE block sorts the survivors (key desc, doc id asc) and
//                 writes the top-k.
// If the survivors exceed SURV_CAP (top docs clustered in few segments) the query's gate flag is raised
// and the buffer-based kernels of select.cuh, launched right behind and gated on that flag, redo it.
//
// The witness pass settles filter_searched_result's "is there a second near-tie anywhere?" question
//  without sorting: two DISTINCT scores closer than DIFF_FILTER_THRESH anywhere below the
// returned prefix imply an adjacent near-tie there, and two such scores falling into one bucket of width
// thresh are found with a single atomicMax per doc.
#pragma once
#include "select.cuh"

namespace ais {

constexpr int SEG_MAX = 2048;            // segments per query
constexpr int SEG_WARPS = 8;             // warps (= segments) per block
constexpr int SEG_MIN_DOCS = 64;
constexpr int SURV_CAP = 4096;
constexpr int COLLECT_THREADS = 256;
constexpr int SEL_TILE = 256;            // docs per tile of the tile-maximum table (= BM25_SUB)

// ---- score functors: key of doc i, false if the doc is not a candidate ---------------------------
struct ScoreFinal {        // stored combined scores
    const double* fin;
    __device__ __forceinline__ bool operator()(int64_t i, int64_t, uint64_t& key) const {
        key = dkey(fin[i]);
        return true;
    }
    __device__ __forceinline__ double val(int64_t i) const { return fin[i]; }
    __device__ __forceinline__ bool is_seed(int64_t) const { return false; }
};
struct ScoreRerank {       // pass 2: webui.py:208 blend; the PRF seeds are not candidates (webui.py:217)
    const double* fin; const float* rer; float cq; CombineParams cp; const int64_t* seeds; int depth;   // rer[i] * cq: see RerankF
    __device__ __forceinline__ bool operator()(int64_t i, int64_t id, uint64_t& key) const {
        const double r = __dadd_rn(__dmul_rn(cp.wo, fin[i]), (double)__fmul_rn(cp.wr, __fmul_rn(rer[i], cq)));
        key = dkey(r);
        return true;                 // seeds are filtered by is_seed() only for docs that would otherwise qualify
    }
    __device__ __forceinline__ double val(int64_t i) const {
        return __dadd_rn(__dmul_rn(cp.wo, fin[i]), (double)__fmul_rn(cp.wr, __fmul_rn(rer[i], cq)));
    }
    __device__ __forceinline__ bool is_seed(int64_t id) const {
        bool hit = false;
        for (int t = 0; t < depth; ++t) hit = hit || (seeds[t] == id);
        return hit;
    }
};

struct SelectArgs {        // everything the three kernels share
    const float* sim; double* fin; const float* rer;
    int64_t rer_qstride;        // ld, or 0 when `rer` is the shared column buffer
    const float* rer_scale;     // null (scale 1) or the queries' scalars [q * DIM]
    int64_t n, ld, id_base;
    CombineParams cp;
    const double* maxes;        // [nq][2] (mode 0)
    const int64_t* seeds_all;   // [nq][MAX_DEPTH] (mode 2)
    int depth;
    int mode;                   // 0 combine (pass 1), 1 stored finals, 2 rerank blend (pass 2)
    int n_seg; int64_t seg_len; // a segment = seg_len / SEL_TILE whole tiles
    uint64_t* seg_max;          // [nq][SEG_MAX]
    uint64_t* tile_max;         // [nq][tile_ld]: best key of every SEL_TILE docs, seeds included (an upper bound: skip filter only)
    int64_t tile_ld, n_tiles;
    uint64_t* max_all;          // [nq] atomicMax over ALL docs (mode 2: max R) or null
    uint64_t* thr;              // [nq]
    int* surv_count;            // [nq]
    uint64_t* surv_keys;        // [nq][SURV_CAP]
    int64_t* surv_ids;
    int* gate;                  // [nq] raised when the survivors overflow
};

template <int MODE, typename Body>
__device__ __forceinline__ void with_functor(const SelectArgs& a, int qi, const int64_t* seeds_smem, Body&& body) {
    if (MODE == 1) {
        ScoreFinal f{a.fin + (size_t)qi * a.ld};
        body(f);
    } else {
        ScoreRerank f{a.fin + (size_t)qi * a.ld, a.rer + (size_t)qi * a.rer_qstride,
                      a.rer_scale ? a.rer_scale[(size_t)qi * DIM] : 1.0f, a.cp, seeds_smem, a.depth};
        body(f);
    }
}

// 64-bit warp maximum with two redux.sync instead of five shuffle rounds
__device__ __forceinline__ uint64_t warp_max_u64_redux(uint64_t v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((uint64_t)mh << 32) | ml;
}

// ---- 1. segment maxima (+ tile maxima) ---------------------------------------------------------------
// The stream runs in the double domain (one fmax per doc; an order-preserving key costs a dozen integer instructions
// and made this kernel issue-bound at 71 % with DRAM at 51 %): keys are formed once per tile / segment.  A NaN score
// (only possible with a zero weight times -inf, or NaN rows) sorts first like dkey(NaN) does: it is tracked by a flag.
constexpr uint64_t KEY_NAN = 0xFFF8000000000000ull;

template <int MODE>
__global__ void __launch_bounds__(32 * SEG_WARPS)
segmax_kernel(SelectArgs a) {
    __shared__ int64_t seeds[MAX_DEPTH];
    __shared__ uint64_t wall[SEG_WARPS];
    const int qi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE == 2) {
        if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = a.seeds_all[qi * MAX_DEPTH + threadIdx.x];
        __syncthreads();
    }
    const int seg = blockIdx.x * SEG_WARPS + warp;
    uint64_t all_best = KEY_EMPTY;
    if (seg < a.n_seg) {
        const int64_t tiles_per_seg = a.seg_len / SEL_TILE;
        const int64_t t0 = (int64_t)seg * tiles_per_seg;
        const int64_t t1 = t0 + tiles_per_seg < a.n_tiles ? t0 + tiles_per_seg : a.n_tiles;
        uint64_t* tmax = a.tile_max + (int64_t)qi * a.tile_ld;
        double bestd = -INFINITY;              // best non-seed score of the segment so far (valid once `has`)
        bool has = false, best_nan = false;
        with_functor<MODE>(a, qi, seeds, [&](auto& f) {
#pragma unroll 1
            for (int64_t tile = t0; tile < t1; ++tile) {
                const int64_t lo = tile * SEL_TILE;
                const int64_t hi = lo + SEL_TILE < a.n ? lo + SEL_TILE : a.n;
                double tb = -INFINITY;
                bool tnan = false, tany = false;
                // the PRF seeds are not candidates (webui.py:217); at most `depth` of the shard's tiles hold one, so
                // the per-doc seed test runs only there (a lane-private "new best?" test is NOT rare: a lane meets
                // only 160 docs per segment, and some lane of the warp finds a new best in nearly every tile)
                const bool seeded = MODE == 2 && __any_sync(0xffffffffu, lane < a.depth && seeds[lane] >= a.id_base + lo &&
                                                                             seeds[lane] < a.id_base + hi);
                if (hi - lo == SEL_TILE && !seeded) {
                    // all 8 docs of a lane are loaded before any arithmetic (16 loads in flight per lane)
                    double rr[SEL_TILE / 32];
#pragma unroll
                    for (int u = 0; u < SEL_TILE / 32; ++u) rr[u] = f.val(lo + 32 * u + lane);
                    double m = rr[0];
                    bool nn = rr[0] != rr[0];
#pragma unroll
                    for (int u = 1; u < SEL_TILE / 32; ++u) { m = fmax(m, rr[u]); nn = nn || (rr[u] != rr[u]); }
                    tb = m;
                    tnan = nn;
                    tany = true;
                    bestd = fmax(bestd, m);
                    best_nan = best_nan || nn;
                    has = true;
                } else {
                    for (int64_t i = lo + lane; i < hi; i += 32) {
                        const double r = f.val(i);
                        tnan = tnan || (r != r);
                        tb = fmax(tb, r);
                        tany = true;
                        if ((r > bestd || !has || r != r) && !f.is_seed(a.id_base + i)) {
                            if (r != r) best_nan = true; else if (r > bestd || !has) bestd = r;
                            has = true;
                        }
                    }
                }
                uint64_t tbest = tnan ? KEY_NAN : (tany ? dkey(tb) : KEY_EMPTY);
                tbest = warp_max_u64_redux(tbest);
                if (lane == 0) tmax[tile] = tbest;
                all_best = tbest > all_best ? tbest : all_best;
            }
        });
        uint64_t best = best_nan ? KEY_NAN : (has ? dkey(bestd) : KEY_EMPTY);
        best = warp_max_u64(best);
        if (lane == 0) a.seg_max[(size_t)qi * SEG_MAX + seg] = best;
    }
    if (a.max_all) {
        if (lane == 0) wall[warp] = all_best;              // already warp-uniform
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t m = wall[0];
            for (int w = 1; w < SEG_WARPS; ++w) m = wall[w] > m ? wall[w] : m;
            if (m != KEY_EMPTY) atomicMax(reinterpret_cast<unsigned long long*>(&a.max_all[qi]), (unsigned long long)m);
        }
    }
}

// ---- 2. threshold = k-th largest segment maximum ------------------------------------------------------
__global__ void __launch_bounds__(256)
threshold_kernel(const uint64_t* __restrict__ seg_max, int n_seg, int k, uint64_t* __restrict__ thr,
                 int* __restrict__ surv_count, int* __restrict__ gate) {
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
        thr[qi] = t != KEY_EMPTY ? t : 1ull;          // fewer than k non-empty segments: every doc qualifies
        surv_count[qi] = 0;
        gate[qi] = 0;
    }
}

// ---- 3a. collect the survivors -------------------------------------------------------------------------
// One lane per tile reads the tile's best key; the warp then walks the (rare) tiles that reach the threshold.
template <int MODE>
__global__ void __launch_bounds__(COLLECT_THREADS)
collect_kernel(SelectArgs a) {
    __shared__ int64_t seeds[MAX_DEPTH];
    const int qi = blockIdx.y, lane = threadIdx.x & 31;
    if (MODE == 2) {
        if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = a.seeds_all[qi * MAX_DEPTH + threadIdx.x];
        __syncthreads();
    }
    const uint64_t T = a.thr[qi];
    int* cnt = a.surv_count + qi;
    uint64_t* sk = a.surv_keys + (size_t)qi * SURV_CAP;
    int64_t* si = a.surv_ids + (size_t)qi * SURV_CAP;
    const uint64_t* tmax = a.tile_max + (int64_t)qi * a.tile_ld;
    const int64_t warp_id = ((int64_t)blockIdx.x * COLLECT_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * COLLECT_THREADS) >> 5;
    with_functor<MODE>(a, qi, seeds, [&](auto& f) {
        for (int64_t tb = warp_id * 32; tb < a.n_tiles; tb += n_warps * 32) {
            const int64_t my_tile = tb + lane;
            unsigned todo = __ballot_sync(0xffffffffu, my_tile < a.n_tiles && tmax[my_tile] >= T);
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t lo = (tb + b) * SEL_TILE;
#pragma unroll 1
                for (int u = 0; u < SEL_TILE / 32; ++u) {
                    const int64_t i = lo + u * 32 + lane;
                    uint64_t key = KEY_EMPTY;
                    if (i < a.n) f(i, a.id_base + i, key);
                    const bool pass = i < a.n && key >= T && !f.is_seed(a.id_base + i);
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
    });
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
    if (cnt > SURV_CAP) {                       // overflow: the gated buffer-select kernels take over
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

// ---- near-tie witness ------------------------------------------------------------------------------------
struct WitnessArgs {
    const double* fin; const float* rer;      // rows of the query
    const float* rer_scale;                   // null (1) or the query's scalar (column mode)
    int64_t n, id_base;
    CombineParams cp;
    int second_pass;                          // 1: value = normalised blend R (seeds skipped); 0: combined score
    const int64_t* seeds; int depth;
    uint64_t last_key;                        // key of the last entry of the sorted prefix
    const double* max_r; int normalize;          // device scalar (nullable)
    double thresh, inv_thresh;
    uint64_t* table; int64_t n_buckets;
    int* flag;
};

__device__ __forceinline__ double witness_value(int normalize, double max_r, double raw) {
    return (normalize && max_r > 0.0) ? __ddiv_rn(raw, max_r) : raw;             // webui.py:210-211
}

__global__ void __launch_bounds__(256)
witness_kernel(WitnessArgs a) {
    __shared__ int64_t seeds[MAX_DEPTH];
    if (threadIdx.x < MAX_DEPTH) seeds[threadIdx.x] = (a.second_pass && threadIdx.x < a.depth) ? a.seeds[threadIdx.x] : -1;
    __syncthreads();
    const double max_r = a.max_r ? *a.max_r : 0.0;
    const float cq = a.rer_scale ? *a.rer_scale : 1.0f;
    const double v_last = a.last_key == ~0ull ? 1.0 : witness_value(a.normalize, max_r, dkey_inv(a.last_key));
    const uint64_t neg_inf = dkey(-INFINITY);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        if (*((volatile int*)a.flag)) return;                   // somebody found one
        double raw = a.fin[i];
        if (a.second_pass) raw = __dadd_rn(__dmul_rn(a.cp.wo, raw), (double)__fmul_rn(a.cp.wr, __fmul_rn(a.rer[i], cq)));
        const uint64_t key = dkey(raw);
        if (key > a.last_key || key <= neg_inf) continue;       // inside the prefix region / masked
        const int64_t id = a.id_base + i;
        bool seed = false;
        for (int t = 0; t < MAX_DEPTH; ++t) seed = seed || (seeds[t] == id);
        if (seed) continue;
        const double v = witness_value(a.normalize, max_r, raw);
        const double off = (v_last - v) * a.inv_thresh;
        if (!(off >= 0.0) || off >= (double)a.n_buckets) continue;
        const uint64_t vk = dkey(v);
        const uint64_t old = atomicMax(reinterpret_cast<unsigned long long*>(&a.table[(int64_t)off]), (unsigned long long)vk);
        if (old != KEY_EMPTY && old != vk) {
            const double o = dkey_inv(old);
            const double d = o > v ? __dsub_rn(o, v) : __dsub_rn(v, o);
            if (d != 0.0 && d < a.thresh) atomicExch(a.flag, 1);
        }
    }
}

}  // namespace ais
