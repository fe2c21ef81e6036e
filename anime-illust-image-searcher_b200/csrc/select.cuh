// Score combination, mask-aware top-k selection, candidate merge, the PRF centroid and the
// filter_searched_result tail - everything of webui.py:376-383, :189-253 and :63-80 that is O(N).
//
// Ordering contract everywhere: (score descending, doc id ascending) == the reference's stable
// sorts of enumerate(...) by -score (webui.py:191-192, :237).  Scores travel as order-preserving
// uint64 keys (common.cuh), so selection is exact on the fp64 values.
//
// Block-level top-k: a shared-memory candidate buffer of SEL_CAP entries with a running threshold
// (the current k-th best).  Each round offers SEL_ROUND new items; items better than the
// threshold are appended with one warp-aggregated atomic per warp; when the buffer could overflow
// it is bitonic-sorted in place (best first), cut to k and the threshold tightened.  Because
// SEL_ROUND <= SEL_CAP - SEL_KMAX an offer never overflows.  After the stream the buffer holds the
// exact top-k, sorted.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int SEL_CAP = 2048;
constexpr int SEL_KMAX = 1024;
constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 4;
constexpr int SEL_ROUND = SEL_THREADS * SEL_ITEMS;  // 1024 <= SEL_CAP - SEL_KMAX

struct SelBuf {
    uint64_t key[SEL_CAP];
    int64_t id[SEL_CAP];
    int count;
    int have_thr;
    uint64_t thr_key;
    int64_t thr_id;
};

__device__ __forceinline__ void sel_init(SelBuf& sb) {
    if (threadIdx.x == 0) {
        sb.count = 0;
        sb.have_thr = 0;
        sb.thr_key = 0;
        sb.thr_id = ID_EMPTY;
    }
    __syncthreads();
}

// Sort the buffer best-first, keep the best k, tighten the threshold.  All threads; ends synced.
__device__ void sel_prune(SelBuf& sb, int k) {
    const int tid = threadIdx.x;
    const int cnt = sb.count;
    for (int i = cnt + tid; i < SEL_CAP; i += SEL_THREADS) {
        sb.key[i] = KEY_EMPTY;
        sb.id[i] = ID_EMPTY;
    }
    __syncthreads();
    for (unsigned size = 2; size <= SEL_CAP; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = tid; t < SEL_CAP / 2; t += SEL_THREADS) {
                const unsigned i = 2 * t - (t & (stride - 1));
                const unsigned l = i + stride;
                const bool up = ((i & size) == 0);
                const uint64_t ka = sb.key[i], kb = sb.key[l];
                const int64_t ia = sb.id[i], ib = sb.id[l];
                const bool l_better = better(kb, ib, ka, ia);
                if (l_better == up) {
                    sb.key[i] = kb; sb.id[i] = ib;
                    sb.key[l] = ka; sb.id[l] = ia;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        const int c = cnt < k ? cnt : k;
        sb.count = c;
        if (c == k) {
            sb.have_thr = 1;
            sb.thr_key = sb.key[k - 1];
            sb.thr_id = sb.id[k - 1];
        }
    }
    __syncthreads();
}

// Offer one item per thread (all 32 lanes of every warp must call this together).
__device__ __forceinline__ void sel_offer(SelBuf& sb, bool valid, uint64_t key, int64_t id) {
    const bool pass = valid && (!sb.have_thr || better(key, id, sb.thr_key, sb.thr_id));
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&sb.count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            sb.key[pos] = key;
            sb.id[pos] = id;
        }
    }
}

// Stream items [lo, hi) through the selector.  F: bool operator()(int64 idx, uint64& key, int64& id)
template <typename F>
__device__ void sel_stream(SelBuf& sb, int64_t lo, int64_t hi, int k, F& f) {
    for (int64_t base = lo; base < hi; base += SEL_ROUND) {
        __syncthreads();
        const int cnt_now = sb.count;      // read by everyone BEFORE anyone may append again
        __syncthreads();
        if (cnt_now + SEL_ROUND > SEL_CAP) sel_prune(sb, k);
        uint64_t keys[SEL_ITEMS];
        int64_t ids[SEL_ITEMS];
        bool ok[SEL_ITEMS];
#pragma unroll
        for (int it = 0; it < SEL_ITEMS; ++it) {
            const int64_t idx = base + it * SEL_THREADS + threadIdx.x;
            ok[it] = false;
            keys[it] = 0;
            ids[it] = 0;
            if (idx < hi) ok[it] = f(idx, keys[it], ids[it]);
        }
#pragma unroll
        for (int it = 0; it < SEL_ITEMS; ++it) sel_offer(sb, ok[it], keys[it], ids[it]);
    }
    __syncthreads();
    sel_prune(sb, k);
}

__device__ __forceinline__ void sel_write(const SelBuf& sb, int k, uint64_t* out_keys, int64_t* out_ids) {
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
        const bool live = i < sb.count;
        out_keys[i] = live ? sb.key[i] : KEY_EMPTY;
        out_ids[i] = live ? sb.id[i] : ID_EMPTY;
    }
}

__device__ __forceinline__ void block_max_to_global(uint64_t v, uint64_t* wscratch, uint64_t* dst) {
    v = warp_max_u64(v);
    if ((threadIdx.x & 31) == 0) wscratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SEL_THREADS / 32; ++w) v = wscratch[w] > v ? wscratch[w] : v;
        atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)v);
    }
}

// ---- weights of the combine (webui.py:376-383, done by bm25_combine_kernel) and of the PRF blend (webui.py:208) ----
struct CombineParams {
    double wb;   // BM25_WEIGHT     (python float * float64 array -> fp64 multiply)
    float wd;    // DOC2VEC_WEIGHT  (python float * float32 array -> fp32 multiply)
    double wo;   // ORIGINAL_SCORE_WEIGHT
    float wr;    // RERANKED_SCORE_WEIGHT (fp32 multiply, same reason)
};

// ---- merge candidate lists.  Entry (list, query, pos) sits at list*list_stride + query*q_stride + pos:
//      all-gathered lists [n_lists][nq][k]: list_stride = nq*k, q_stride = k;
//      a kernel's per-block lists [nq][grid][k]: list_stride = k, q_stride = grid*k.
struct CandF {
    const uint64_t* keys;
    const int64_t* ids;
    int64_t list_stride, q_stride;
    int qi, k, list0;
    __device__ __forceinline__ bool operator()(int64_t i, uint64_t& key, int64_t& id) const {
        const int64_t list = i / k, pos = i - list * k;
        const size_t o = (size_t)((list + list0) * list_stride + (int64_t)qi * q_stride + pos);
        key = keys[o];
        id = ids[o];
        return key != KEY_EMPTY;
    }
};
// grid (n_groups, nq): block (g, q) merges lists [g*group, min(n_lists, (g+1)*group)) of query q into
// the sorted top-k_out written at out[(q*n_groups + g)*out_stride ...] (+ its count).  A two-level
// tree (group = 16, then one block per query) keeps the serial part short.
__global__ void __launch_bounds__(SEL_THREADS)
merge_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ ids, int n_lists, int group,
             int64_t list_stride, int64_t q_stride, int k_in, int k_out,
             uint64_t* __restrict__ out_keys, int64_t* __restrict__ out_ids, int64_t out_stride,
             int32_t* __restrict__ out_count, const int* __restrict__ gate) {
    __shared__ SelBuf sb;
    const int g = blockIdx.x, qi = blockIdx.y;
    if (gate && !gate[qi]) return;
    const int list0 = g * group;
    const int my_lists = (n_lists - list0 < group) ? (n_lists - list0) : group;
    sel_init(sb);
    CandF f{keys, ids, list_stride, q_stride, qi, k_in, list0};
    const int64_t total = (int64_t)my_lists * k_in;
    if (total > 0) sel_stream(sb, 0, total, k_out, f);
    else { __syncthreads(); sel_prune(sb, k_out); }
    const size_t o = ((size_t)qi * gridDim.x + g) * (size_t)out_stride;
    sel_write(sb, k_out, out_keys + o, out_ids + o);
    if (threadIdx.x == 0 && out_count) out_count[qi * gridDim.x + g] = sb.count;
}

// How much of a merged list is certainly the true global order?  Every doc a shard did NOT return scores at most that
// shard's last returned key, so merged entries strictly above the largest such key (over the lists that came back full)
// are exact; so are the first k (the classic top-k-of-top-k argument).  The count is cut to that prefix: the filter
// then treats the rest as unknown instead of wrong.  Shards can therefore return SHORT lists (k ~ 1024 / n_shards)
// and the merged prefix is still ~1024 deep.
__global__ void __launch_bounds__(256)
prefix_bound_kernel(const uint64_t* __restrict__ keys, int n_lists, int64_t list_stride, int64_t q_stride, int k,
                    const uint64_t* __restrict__ merged, int k_out, int32_t* __restrict__ count) {
    __shared__ uint64_t s_bound[8];
    __shared__ int s_cnt[8];
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t b = KEY_EMPTY;
    for (int l = tid; l < n_lists; l += blockDim.x) {
        const uint64_t last = keys[(size_t)(l * list_stride + (int64_t)qi * q_stride + (k - 1))];
        b = last > b ? last : b;
    }
    b = warp_max_u64(b);
    if (lane == 0) s_bound[wid] = b;
    __syncthreads();
    b = s_bound[0];
    for (int w = 1; w < 8; ++w) b = s_bound[w] > b ? s_bound[w] : b;
    const int m = count[qi];
    int c = 0;
    for (int i = tid; i < m; i += blockDim.x) c += merged[(size_t)qi * k_out + i] > b;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_cnt[wid] = c;
    __syncthreads();
    if (tid == 0) {
        int exact = 0;
        for (int w = 0; w < 8; ++w) exact += s_cnt[w];
        const int classic = k < m ? k : m;
        count[qi] = exact > classic ? exact : classic;
    }
}

// ---- PRF: stored rows of the top docs, centroid, re-query vector (webui.py:195-205) ----------
// rows_out [nq][depth][DIM]: the stored row if the doc lives on this shard, else zeros.
__global__ void gather_top_rows_kernel(const float* __restrict__ rows, int64_t n, int64_t id_base,
                                       const int64_t* __restrict__ top_ids_all, int depth, float* __restrict__ rows_out) {
    const int qi = blockIdx.y, t = blockIdx.x;
    const int64_t id = top_ids_all[qi * MAX_DEPTH + t] - id_base;
    float* dst = rows_out + ((size_t)qi * depth + t) * DIM;
    for (int j = threadIdx.x; j < DIM; j += blockDim.x) dst[j] = (id >= 0 && id < n) ? rows[id * DIM + j] : 0.0f;
}

// numpy's pairwise float64 sum for n < 128 contiguous items (what np.average's wgt.sum() does)
__device__ inline double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// One block per query.  collapse = 1 reproduces the reference exactly: the (300,2) array of
// (index, value) pairs is averaged, divided by its Frobenius norm INCLUDING the index column,
// the indices are round()-ed (all become 0) and gensim's sparse2full keeps the LAST value for id 0
// -> q2 = [ c_299 / ||c||, 0, ..., 0 ].  collapse = 0 gives the full unit centroid.
__global__ void __launch_bounds__(320)
prf_query_kernel(const float* __restrict__ rows_in,        // [nq][depth][DIM]
                 const double* __restrict__ top_scores_all, // [nq][MAX_DEPTH]
                 int depth, int collapse, float* __restrict__ q2_out, int32_t* __restrict__ status) {
    __shared__ double w[MAX_DEPTH];
    __shared__ double val[DIM];
    __shared__ double icol[DIM];
    __shared__ double scal[2];
    const int qi = blockIdx.x, tid = threadIdx.x;
    if (tid < depth) w[tid] = top_scores_all[qi * MAX_DEPTH + tid];
    __syncthreads();
    bool bad = false;
    for (int t = 0; t < depth; ++t) bad = bad || !isfinite(w[t]);
    const double scl = np_pairwise_sum(w, depth);
    float* q2 = q2_out + (size_t)qi * DIM;
    if (bad || scl == 0.0) {
        if (tid < DIM) q2[tid] = 0.0f;
        if (tid == 0 && status[qi] == 0) status[qi] = bad ? 1 /*AIS_Q_NAN_WEIGHTS*/ : 2 /*AIS_Q_ZERO_WEIGHT_SUM*/;
        return;
    }
    if (tid < DIM) {
        double c, idxcol;
        // np.multiply(a, wgt).sum(axis=0) / scl : sequential over the `depth` docs
        const float* r = rows_in + (size_t)qi * depth * DIM + tid;
        c = __dmul_rn((double)r[0], w[0]);
        idxcol = __dmul_rn((double)tid, w[0]);
        for (int t = 1; t < depth; ++t) {
            c = __dadd_rn(c, __dmul_rn((double)r[(size_t)t * DIM], w[t]));
            idxcol = __dadd_rn(idxcol, __dmul_rn((double)tid, w[t]));
        }
        c = __ddiv_rn(c, scl);
        idxcol = __ddiv_rn(idxcol, scl);
        val[tid] = c;
        icol[tid] = idxcol;
    }
    __syncthreads();
    if (tid == 0) {
        double ss = 0.0;
        if (collapse) {
            // Frobenius norm over BOTH columns of the (300,2) array (webui.py:201)
            for (int j = 0; j < DIM; ++j) {
                ss = __dadd_rn(ss, __dmul_rn(icol[j], icol[j]));
                ss = __dadd_rn(ss, __dmul_rn(val[j], val[j]));
            }
        } else {
            for (int j = 0; j < DIM; ++j) ss = __dadd_rn(ss, __dmul_rn(val[j], val[j]));
        }
        scal[0] = sqrt(ss);
    }
    __syncthreads();
    const double fro = scal[0];
    if (tid < DIM) val[tid] = __ddiv_rn(val[tid], fro);
    __syncthreads();
    if (tid == 0) {
        // gensim unitvec: length = sqrt(sum(val**2)) over the list entries, sequential
        double ss = 0.0;
        for (int j = 0; j < DIM; ++j) ss = __dadd_rn(ss, __dmul_rn(val[j], val[j]));
        scal[1] = sqrt(ss);
    }
    __syncthreads();
    const double len = scal[1];
    if (!(len > 0.0)) {
        if (tid < DIM) q2[tid] = 0.0f;
        if (tid == 0 && status[qi] == 0) status[qi] = (len == 0.0) ? 3 /*AIS_Q_ZERO_VECTOR*/ : 1;
        return;
    }
    if (tid < DIM) {
        const double u = (len != 1.0) ? __ddiv_rn(val[tid], len) : val[tid];
        if (collapse) q2[tid] = (tid == 0) ? (float)((len != 1.0) ? __ddiv_rn(val[DIM - 1], len) : val[DIM - 1]) : 0.0f;
        else q2[tid] = (float)u;
    }
}

// ---- tail: filter_searched_result (webui.py:63-80) on the sorted prefix -----------------------
struct TailParams {
    double thresh;      // DIFF_FILTER_THRESH
    int topn;
    int depth;          // number of pinned top docs (0 in the no-PRF branch)
    int normalize;      // 1: divide rest scores by max_r when max_r > 0 (webui.py:210-211)
    int64_t n_total;    // docs in the whole index (all shards)
};

__device__ __forceinline__ double tail_value(int64_t i, int depth, const uint64_t* rest_keys, double max_r, int normalize) {
    if (i < depth) return 1.0;                              // webui.py:222
    double v = dkey_inv(rest_keys[i - depth]);
    if (normalize && max_r > 0.0) v = __ddiv_rn(v, max_r);
    return v;
}

// one block per query.  rest_*: sorted (best first) non-top docs, rest_count[q] valid entries.
__global__ void __launch_bounds__(SEL_THREADS)
tail_kernel(const uint64_t* __restrict__ rest_keys_all, const int64_t* __restrict__ rest_ids_all, int64_t rest_stride,
            const int32_t* __restrict__ rest_count, const int64_t* __restrict__ rest_count64,
            const int64_t* __restrict__ top_ids_all, const double* __restrict__ max_r_all, TailParams tp,
            const int32_t* __restrict__ witness,      // nullable; [q] = 1: a near-tie is known to exist below the prefix
            int64_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_count,
            int32_t* __restrict__ out_ambiguous, uint64_t* __restrict__ out_last_key) {
    __shared__ unsigned long long s_first, s_second, s_nonpos;
    const int qi = blockIdx.x, tid = threadIdx.x;
    const uint64_t* rest_keys = rest_keys_all + (size_t)qi * rest_stride;
    const int64_t* rest_ids = rest_ids_all + (size_t)qi * rest_stride;
    const int64_t m = rest_count64 ? rest_count64[qi] : (int64_t)rest_count[qi];
    const double max_r = max_r_all ? max_r_all[qi] : 0.0;
    const int depth = tp.depth;
    const int64_t len = depth + m;                          // sorted prefix we hold
    const bool complete = (len >= tp.n_total);              // it is the whole list
    const unsigned long long INF = ~0ull;
    if (tid == 0) { s_first = INF; s_second = INF; s_nonpos = INF; }
    __syncthreads();

    // adjacent differences; exact zeros ignored; "found" = diff < thresh (webui.py:66-73)
    const int64_t CH = 65536;
    for (int64_t c0 = 0; c0 < len - 1; c0 += CH) {
        const int64_t c1 = (c0 + CH < len - 1) ? c0 + CH : len - 1;
        for (int pass = 0; pass < 2; ++pass) {
            const unsigned long long first = s_first;
            if (pass == 1 && first == INF) break;
            for (int64_t i = c0 + tid; i < c1; i += SEL_THREADS) {
                const double a = tail_value(i, depth, rest_keys, max_r, tp.normalize);
                const double b = tail_value(i + 1, depth, rest_keys, max_r, tp.normalize);
                double d = __dsub_rn(a, b);
                if (d == 0.0) d = INFINITY;
                if (d < tp.thresh) {
                    if (pass == 0) atomicMin(&s_first, (unsigned long long)i);
                    else if ((unsigned long long)i > first) atomicMin(&s_second, (unsigned long long)i);
                }
            }
            __syncthreads();
        }
        if (s_second != INF) break;
        __syncthreads();
    }
    __syncthreads();
    // first index (within topn) whose score is not > 0  (webui.py:80 keeps score > 0 only)
    const int64_t lim = len < tp.topn ? len : tp.topn;
    for (int64_t i = tid; i < lim; i += SEL_THREADS) {
        const double a = tail_value(i, depth, rest_keys, max_r, tp.normalize);
        if (!(a > 0.0)) atomicMin(&s_nonpos, (unsigned long long)i);
    }
    __syncthreads();
    const unsigned long long first = s_first, second = s_second;
    const int64_t npos = (s_nonpos == INF) ? lim : (int64_t)s_nonpos;
    int64_t t;
    bool ambiguous = false;
    if (second != INF) t = (int64_t)second;                       // webui.py:76-77
    else if (first != INF) {
        if (complete) t = (int64_t)first;                         // webui.py:74-75
        else { t = lim; ambiguous = (int64_t)first < npos && !(witness && witness[qi]); }   // a 2nd point may exist beyond the prefix
    } else t = complete ? len : lim;
    int64_t cnt = t < npos ? t : npos;
    if (cnt > lim) cnt = lim;
    // webui.py:78,80: divide by the maximum of the list (== 1.0 when top docs are pinned at 1.0)
    const double max_val = (len > 0) ? tail_value(0, depth, rest_keys, max_r, tp.normalize) : 1.0;
    for (int64_t i = tid; i < cnt; i += SEL_THREADS) {
        const double a = tail_value(i, depth, rest_keys, max_r, tp.normalize);
        out_scores[(size_t)qi * tp.topn + i] = __ddiv_rn(a, max_val);
        out_ids[(size_t)qi * tp.topn + i] = (i < depth) ? top_ids_all[qi * MAX_DEPTH + i] : rest_ids[i - depth];
    }
    if (tid == 0) {
        out_count[qi] = (int32_t)cnt;
        out_ambiguous[qi] = ambiguous ? 1 : 0;
        if (out_last_key) out_last_key[qi] = m > 0 ? rest_keys[m - 1] : ~0ull;
    }
}

// ---- fallback: sort ALL keys of a shard (ambiguous filter outcome; keys from fill_keys_kernel, select2.cuh) ----
constexpr int GS_TILE = 2048;     // elements sorted per block in shared memory
constexpr int GS_THREADS = 256;

__device__ __forceinline__ void cmpswap_best_first(uint64_t& ka, int64_t& ia, uint64_t& kb, int64_t& ib, bool up) {
    const bool l_better = better(kb, ib, ka, ia);
    if (l_better == up) {
        uint64_t tk = ka; ka = kb; kb = tk;
        int64_t ti = ia; ia = ib; ib = ti;
    }
}

// all (size, stride) steps with stride < GS_TILE for sizes in [size_lo, size_hi], inside shared memory
__global__ void __launch_bounds__(GS_THREADS)
bitonic_local_kernel(uint64_t* __restrict__ keys, int64_t* __restrict__ ids, int64_t n_pad, unsigned long long size_lo,
                     unsigned long long size_hi) {
    __shared__ uint64_t sk[GS_TILE];
    __shared__ int64_t si[GS_TILE];
    const int64_t base = (int64_t)blockIdx.x * GS_TILE;
    for (int i = threadIdx.x; i < GS_TILE; i += GS_THREADS) { sk[i] = keys[base + i]; si[i] = ids[base + i]; }
    __syncthreads();
    for (unsigned long long size = size_lo; size <= size_hi; size <<= 1) {
        unsigned long long s0 = size >> 1;
        if (s0 >= GS_TILE) s0 = GS_TILE >> 1;
        for (unsigned long long stride = s0; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < GS_TILE / 2; t += GS_THREADS) {
                const unsigned i = 2 * t - (t & ((unsigned)stride - 1));
                const unsigned l = i + (unsigned)stride;
                const bool up = (((unsigned long long)(base + i) & size) == 0);
                cmpswap_best_first(sk[i], si[i], sk[l], si[l], up);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < GS_TILE; i += GS_THREADS) { keys[base + i] = sk[i]; ids[base + i] = si[i]; }
}

// one (size, stride) step with stride >= GS_TILE, straight in global memory
__global__ void bitonic_global_kernel(uint64_t* __restrict__ keys, int64_t* __restrict__ ids, int64_t n_pad,
                                      unsigned long long size, unsigned long long stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad / 2) return;
    const int64_t i = 2 * t - (t & (int64_t)(stride - 1));
    const int64_t l = i + (int64_t)stride;
    const bool up = (((unsigned long long)i & size) == 0);
    uint64_t ka = keys[i], kb = keys[l];
    int64_t ia = ids[i], ib = ids[l];
    const bool l_better = better(kb, ib, ka, ia);
    if (l_better == up) { keys[i] = kb; ids[i] = ib; keys[l] = ka; ids[l] = ia; }
}

__global__ void count_live_kernel(const uint64_t* __restrict__ keys, int64_t n_pad, int64_t* __restrict__ count) {
    // sorted best-first: live entries form a prefix; count them
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const bool live = keys[i] != KEY_EMPTY;
    const bool next_live = (i + 1 < n_pad) ? (keys[i + 1] != KEY_EMPTY) : false;
    if (live && !next_live) *count = i + 1;
    if (i == 0 && !live) *count = 0;
}

}  // namespace ais
