// ais_b200 engine: the C ABI of include/ais_b200.h over the sm_100a kernels in scan.cuh,
// bm25.cuh and select.cuh.  One engine == one GPU == one contiguous shard of the documents.
//
// Per query batch (<= max_batch queries share every pass over the doc vectors):
//   stage_score    bm25_score_kernel (fp64, bit-exact, per-tile records)  +  one scan kernel (fp32 / 3xTF32 dot,
//                  one pass over the rows) -> this shard's {max bm25, max dot}                      webui.py:352,374
//   stage_combine  bm25_combine_kernel: final = 0.5*bm25/max + 0.5*dot/max; streaming select          webui.py:376-383,191-195
//   stage_top      global top-`depth` docs (PRF seeds), optional row gather  webui.py:193-199
//   stage_requery  re-query vector (host callback or device centroid); its scores: column 0 times a scalar for the
//                  reference's collapsed centroid, else a 2nd dense scan; R = 0.7*final + 0.3*rer, max(R),
//                  streaming select without the seeds                              webui.py:200-217
//   stage_finish   merge, normalise, filter_searched_result, [:topn]         webui.py:219-246,63-80
// ais_search chains the stages on one GPU; a doc-sharded caller puts its collectives between them.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "../../include/ais_b200.h"
#include "bm25.cuh"
#include "build.cuh"
#include "common.cuh"
#include "scan.cuh"
#include "scan_tc.cuh"
#include "scan_pair.cuh"
#include "select.cuh"
#include "select2.cuh"

using namespace ais;
static_assert(SEL_TILE == BM25_SUB, "the BM25 sub-tile is the tile of the tile-maximum table");

namespace {

thread_local char g_err[768] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(AIS_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)
#define TRY(call)                  \
    do {                           \
        int _s = (call);           \
        if (_s != AIS_OK) return _s; \
    } while (0)

#define LAUNCHED(e) do { (e)->kernel_launches++; CK(cudaGetLastError()); } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int MERGE_GROUP = 16;
constexpr int SEL_MIN_CHUNK = 4096;

// ---- tiny helper kernels ----------------------------------------------------------------------
__global__ void init_keys_kernel(uint32_t* maxs, uint64_t* maxb, uint64_t* maxr, int32_t* status, int nq, int which) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    if (which & 1) { maxs[i] = fkey(-INFINITY); maxb[i] = dkey(-INFINITY); status[i] = 0; }
    if (which & 2) { maxs[i] = fkey(-INFINITY); maxr[i] = dkey(-INFINITY); }
}
__global__ void maxes_kernel(const uint64_t* maxb, const uint32_t* maxs, int nq, double* maxes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    maxes[2 * i] = dkey_inv(maxb[i]);
    maxes[2 * i + 1] = (double)fkey_inv(maxs[i]);
}
__global__ void maxr_kernel(const uint64_t* maxr, int nq, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = dkey_inv(maxr[i]);
}
// merged top candidates [nq][k_stride] -> top_ids / top_scores [nq][MAX_DEPTH]
__global__ void top_unpack_kernel(const uint64_t* keys, const int64_t* ids, int k_stride, int depth, int64_t* top_ids,
                                  double* top_scores) {
    const int qi = blockIdx.x, t = threadIdx.x;
    if (t >= MAX_DEPTH) return;
    int64_t id = ID_EMPTY;
    double sc = -INFINITY;
    if (t < depth) {
        const uint64_t k = keys[(size_t)qi * k_stride + t];
        if (k != KEY_EMPTY) { id = ids[(size_t)qi * k_stride + t]; sc = dkey_inv(k); }
    }
    top_ids[qi * MAX_DEPTH + t] = id;
    top_scores[qi * MAX_DEPTH + t] = sc;
}
__global__ void keys_from_scores_kernel(const double* scores, int64_t n, uint64_t* keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = dkey(scores[i]);
}
// a single sorted candidate list per query: "merging" is a copy plus a count of the live entries
__global__ void copy_list_kernel(const uint64_t* keys, const int64_t* ids, int k, uint64_t* out_keys, int64_t* out_ids,
                                 int32_t* out_count) {
    const int qi = blockIdx.x;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t kk = keys[(size_t)qi * k + i];
        out_keys[(size_t)qi * k + i] = kk;
        out_ids[(size_t)qi * k + i] = ids[(size_t)qi * k + i];
        mine += kk != KEY_EMPTY;
    }
    if (mine) atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0 && out_count) out_count[qi] = cnt;
}
__global__ void fill_empty_kernel(uint64_t* keys, int64_t* ids, int64_t lo, int64_t hi) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi) { keys[i] = KEY_EMPTY; ids[i] = ID_EMPTY; }
}

}  // namespace

struct ais_engine {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    ais_params p;
    int64_t first_doc = 0, n_total = -1;

    // index
    Buf rows;  int64_t n_vec = 0, cap_vec = 0;
    Buf post_ptr, post_doc, post_tf, idf, kd, g1, doc_len;
    Buf post_len, kd_tab, g1_tab;  int64_t max_doc_len = -1;   // doc length per posting (uint16) + K_d / quotient per length: the
                                                               // score kernel's per-posting lookup is an L1 hit instead of a gather
    double avgdl = 0.0;
    bool has_tf = false;
    int32_t n_vocab = 0;
    int64_t n_bm25 = -1, n_post = 0, n_built = -1;

    // per-batch work
    int qt_cap = 0;
    int64_t ld = 0;
    Buf sim, scratch64, fin_ext, rer, d_q, d_q2, d_qt, maxs_key, maxb_key, maxr_key, maxes_own, maxr_own;
    Buf top_ids, top_scores, status, rows_own;
    Buf blk_keys, blk_ids, grp_keys, grp_ids, cand_keys, cand_ids, rest_keys, rest_ids, rest_count;
    Buf p1_keys, p1_ids;  int p1_k = 0;       // this shard's pass-1 candidates [nq][p1_k] (the pass-2 threshold starts from them)
    Buf out_ids, out_scores, out_count, out_amb;
    Buf fs_keys, fs_ids, fs_count;
    Buf bm25_slices;
    int bm25_t_cap = 1;
    Buf q_nreq;                // [qt_cap] number of required terms per query
    Buf q_idf;                 // [qt_cap][MAX_TERMS] idf of every query term
    Buf qnorm_tab;  bool qnorm_ready = false;   // [qt_cap] per-query constants of the combine, formed once per batch (finals.cuh QNorm)
    Buf tile_hdr;              // [qt_cap][tile_ld][8] bitmap of the docs with a BM25 record (bm25.cuh)
    Buf tile_off;              // [qt_cap][tile_ld] first record slot of the tile, relative to rec_base[q]
    Buf rec_val, rec_pos;      // record pools: fp64 values / positions inside the tile
    Buf rec_base;              // [qt_cap] int64 pool offset of the query's records
    int64_t* h_rec_base = nullptr;
    Buf term_bits, slot_terms;  // bitmap path: presence bitmaps [n_slots][n_tiles][32 B] of the batch's distinct terms / their ids
    int32_t* h_slot_terms = nullptr;
    int n_slots = 0;
    bool use_bits = false;      // current batch: BM25 from the term bitmaps (tf == 1 index, bitmaps fit) - else per-tile records
    bool ieee_div = false;      // AIS_IEEE_DIV=1: __ddiv_rn for bm25 / max (cross-check of the FMA-corrected quotient)
    int combine_occ = 5;        // AIS_COMBINE_OCC: the same for bm25_combine_kernel (5 / 6 / 8)
    int score_occ = 6;          // AIS_SCORE_OCC: resident 256-thread blocks per SM the score kernel is compiled for (4 / 5 / 6)
    bool want_bits = false;     // AIS_BM25_BITMAP=1: the bitmap path where it applies (default: per-tile records)
    int64_t bitmap_cap_bytes = 4LL << 30;     // AIS_BM25_BITMAP_MB
    int64_t bitmap_batches = 0;
    std::vector<int64_t> h_post_ptr;          // host copy of post_ptr: sizes the record pool of a batch (sum of df per query)
    Buf tile_max;  int64_t tile_ld = 0;       // [qt_cap][tile_ld] best combined key per 256-doc tile (select2.cuh)
    Buf tile_max2;             // the same for the blend R of a dense re-query (pass 2)
    Buf col_lo, col_hi;        // [tile_ld] extreme values of the cached column per tile (pass 2, column mode)
    const double* cur_maxes = nullptr;        // device [nq][2] global maxima of the current batch (caller- or engine-owned)
    bool ext_fin = false;      // current batch: combined scores were supplied (ais_rerank), not computed
    bool bound_ok = false;     // current batch, pass 2: AIS_TILE_BOUND=1 and the tile upper bound on R is valid
    bool ub_valid = false;     // current batch, pass 2: the tile upper bound on R is valid (column mode, weights >= 0, pass-1 list kept)
    bool no_skip = false;      // AIS_NO_TILE_SKIP=1: the pass-2 maxima kernel visits every tile
    int64_t skip_passes = 0;
    Buf seg_max, sel_thr, surv_count, surv_keys, surv_ids, gate, witness, last_keys, wit_table;
    uint64_t* h_last_keys = nullptr;
    int sel_k_cap = 0, out_topn_cap = 0;
    int cur_nq = 0;
    bool cur_prf = false;      // second pass ran for the current batch
    // pinned host staging
    float* h_q = nullptr;  QueryTerms* h_qt = nullptr;  float* h_q2 = nullptr;
    int64_t* h_top_ids = nullptr;  double* h_top_scores = nullptr;
    int64_t* h_out_ids = nullptr;  double* h_out_scores = nullptr;  int32_t* h_small = nullptr;  // count|status|amb
    int h_out_cap = 0;

    // stats
    int64_t column_scan_launches = 0, last_tiles_per_seg = 0, bound_passes = 0, pair_scan_launches = 0;
    Buf colbuf;                    // rows[.][col_comp] as a compact array (column mode of the PRF re-query)
    int col_comp = -1;  const void* col_rows_ptr = nullptr;  int64_t col_n = -1;
    bool rer_column = false;       // current batch: rer[q][d] = colbuf[d] * d_q2[q][col_comp], never materialised
    bool requery_dense = false;    // AIS_REQUERY_DENSE=1: always run the dense scan for the PRF re-query
    int sel_deep = 576;            // AIS_SELECT_DEPTH: cap on the prefix length the select stages ask for beyond need (-1: none).
                                   // 10 M docs, batch 256, same box: 512 28.64 ms, 640 28.83, 768 29.16, 1024 29.88 per step -
                                   // deeper lists cost select / pass-2 work, shorter ones bring witness passes back (384: +1.3 ms)
    bool no_bound = true;          // AIS_TILE_BOUND=1: pass 2 of the collapsed re-query skips tiles by an upper bound on R
                                   // (exact, but only selective when BM25 separates the top docs; measured on the benchmark:
                                   // the pass-1-candidate threshold lets > 4096 survivors through for sim-dominated queries)
    int64_t scan_launches = 0, kernel_launches = 0, fullsort_fallbacks = 0, bytes_device = 0;
    bool profiling = false;
    bool use_mma = true;       // >= 5 queries per pass: tensor-core scan (3xTF32); AIS_SCAN_SIMT=1 keeps the fp32 SIMT kernel
    bool tc_wide = true;       // > 32 queries left: 64 queries per tcgen05 pass (AIS_SCAN_TC_WIDE=0: 32)
    int tc_min = 9;            // >= tc_min queries left in a batch: tcgen05 scan, 32 or 64 queries per pass (AIS_SCAN_TC_MIN; 0 = off)
    Buf qsplit;                // [64][300] hi | lo images of the queries of one tcgen05 pass
    CUtensorMap tm_rows, tm_q[3];       // tm_q[0]: 32 queries per pass, tm_q[1]: 64, tm_q[2]: 64 in boxes of 32 rows (CTA pair)
    int scan_pair = 9;                  // AIS_SCAN_PAIR: ring depth (8 / 9) of scan_pair_kernel; 0 = single-CTA scan_tc_kernel<64>
    const void* tm_rows_ptr = nullptr;  int64_t tm_rows_n = -1;  const void* tm_q_ptr = nullptr;
    double kind_ms[AIS_N_KINDS] = {0};          // summed CUDA-event time per kernel class (profiling on)
    int64_t kind_launches[AIS_N_KINDS] = {0};
    struct PendingEv { int kind; cudaEvent_t a, b; };
    std::vector<PendingEv> ev_pending;
    std::vector<cudaEvent_t> ev_free;

    int64_t n() const { return n_vec > 0 ? n_vec : (n_bm25 > 0 ? n_bm25 : 0); }
    int64_t total() const { return n_total >= 0 ? n_total : n_vec; }
};

namespace {

int dev_alloc(ais_engine* e, Buf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return AIS_OK;
    if (b.p) { CK(cudaFree(b.p)); e->bytes_device -= (int64_t)b.cap; b.p = nullptr; b.cap = 0; }
    CK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    e->bytes_device += (int64_t)bytes;
    return AIS_OK;
}
void dev_free(ais_engine* e, Buf& b) {
    if (b.p) { cudaFree(b.p); e->bytes_device -= (int64_t)b.cap; }
    b.p = nullptr; b.cap = 0;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int next_pow2_int(int v) { int r = 1; while (r < v) r <<= 1; return r; }

// blocks per query of the gated streaming select; bounded so that its candidate lists [q][G][k] stay below 2^25 entries
int sel_blocks_cap(const ais_engine* e, int k) {
    int64_t cap = (1LL << 25) / ((int64_t)(e->qt_cap > 0 ? e->qt_cap : 1) * k);
    if (cap > 4LL * e->sm_count) cap = 4LL * e->sm_count;
    return (int)(cap < 4 ? 4 : cap);
}
int sel_blocks(const ais_engine* e, int k) {
    int64_t g = (e->n() + SEL_MIN_CHUNK - 1) / SEL_MIN_CHUNK;
    const int64_t cap = sel_blocks_cap(e, k);
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
int merge_groups(int n_lists) { return (n_lists + MERGE_GROUP - 1) / MERGE_GROUP; }

int check_loaded(const ais_engine* e) {
    if (e->n_bm25 < 0) return fail(AIS_ERR_NOT_LOADED, "BM25 index not loaded (ais_load_bm25)");
    if (e->n_vec != e->n_bm25)
        return fail(AIS_ERR_NOT_LOADED, "doc vectors cover %lld docs but the BM25 index covers %lld",
                    (long long)e->n_vec, (long long)e->n_bm25);
    return AIS_OK;
}

// Per-posting doc lengths and the per-LENGTH tables of K_d and the tf == 1 quotient (both depend on the doc only through
// its length, webui.py:145): bm25_score_kernel then needs no dependent gather into the per-doc arrays.  Lengths beyond
// 65535 keep the per-doc arrays (post_len stays empty).
int build_len_tables(ais_engine* e, int64_t n_docs, bool postings_changed) {
    e->max_doc_len = -1;
    if (n_docs <= 0 || e->n_post <= 0) return AIS_OK;
    Buf d_max;
    TRY(dev_alloc(e, d_max, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d_max.p, 0, sizeof(unsigned long long), e->stream));
    max_len_kernel<<<(unsigned)((n_docs + 255) / 256), 256, 0, e->stream>>>(e->doc_len.as<int64_t>(), n_docs, d_max.as<unsigned long long>());
    LAUNCHED(e);
    unsigned long long mx = 0;
    CK(cudaMemcpyAsync(&mx, d_max.p, sizeof(mx), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    dev_free(e, d_max);
    if (mx >= 65536ull) return AIS_OK;
    TRY(dev_alloc(e, e->kd_tab, (size_t)(mx + 1) * sizeof(double)));
    TRY(dev_alloc(e, e->g1_tab, (size_t)(mx + 1) * sizeof(double)));
    kd_table_kernel<<<(unsigned)((mx + 256) / 256), 256, 0, e->stream>>>((int64_t)mx, e->avgdl, e->p.k1, e->p.b, 1.0 - e->p.b, e->p.k1 + 1.0,
                                                                        e->kd_tab.as<double>(), e->g1_tab.as<double>());
    LAUNCHED(e);
    if (postings_changed || e->post_len.cap < (size_t)e->n_post * sizeof(uint16_t)) {
        TRY(dev_alloc(e, e->post_len, (size_t)e->n_post * sizeof(uint16_t)));
        post_len_kernel<<<(unsigned)((e->n_post + 255) / 256), 256, 0, e->stream>>>(e->post_doc.as<int32_t>(), e->n_post, e->doc_len.as<int64_t>(),
                                                                                   e->post_len.as<uint16_t>());
        LAUNCHED(e);
    }
    CK(cudaStreamSynchronize(e->stream));
    e->max_doc_len = (int64_t)mx;
    return AIS_OK;
}

int ensure_work(ais_engine* e) {
    const int qt = next_pow2_int(e->p.max_batch);
    const int64_t nmax = e->n_vec > e->n_bm25 ? e->n_vec : e->n_bm25;
    const int64_t ld = ((nmax + 63) / 64) * 64 + 64;
    if (qt <= e->qt_cap && ld <= e->ld) return AIS_OK;
    const int q = qt > e->qt_cap ? qt : e->qt_cap;
    const int64_t l = ld > e->ld ? ld : e->ld;
    // per doc and query only the fp32 dot score stays resident; the re-query scores (rer) exist only for a dense
    // re-query and the combined scores are never stored (finals.cuh)
    TRY(dev_alloc(e, e->sim, (size_t)q * l * sizeof(float)));
    TRY(dev_alloc(e, e->d_q, (size_t)q * DIM * sizeof(float)));
    TRY(dev_alloc(e, e->d_q2, (size_t)q * DIM * sizeof(float)));
    TRY(dev_alloc(e, e->d_qt, (size_t)q * sizeof(QueryTerms)));
    TRY(dev_alloc(e, e->maxs_key, (size_t)q * sizeof(uint32_t)));
    TRY(dev_alloc(e, e->maxb_key, (size_t)q * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->maxr_key, (size_t)q * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->maxes_own, (size_t)q * 2 * sizeof(double)));
    TRY(dev_alloc(e, e->maxr_own, (size_t)q * sizeof(double)));
    TRY(dev_alloc(e, e->top_ids, (size_t)q * MAX_DEPTH * sizeof(int64_t)));
    TRY(dev_alloc(e, e->top_scores, (size_t)q * MAX_DEPTH * sizeof(double)));
    TRY(dev_alloc(e, e->status, (size_t)q * sizeof(int32_t)));
    TRY(dev_alloc(e, e->rows_own, (size_t)q * MAX_DEPTH * DIM * sizeof(float)));
    TRY(dev_alloc(e, e->out_count, (size_t)q * sizeof(int32_t)));
    TRY(dev_alloc(e, e->out_amb, (size_t)q * sizeof(int32_t)));
    TRY(dev_alloc(e, e->seg_max, (size_t)q * SEG_MAX * sizeof(uint64_t)));
    const int64_t tl = (l + SEL_TILE - 1) / SEL_TILE;
    TRY(dev_alloc(e, e->tile_max, (size_t)q * tl * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->tile_hdr, (size_t)q * tl * 8 * sizeof(uint32_t)));
    TRY(dev_alloc(e, e->tile_off, (size_t)q * tl * sizeof(uint32_t)));
    TRY(dev_alloc(e, e->q_nreq, (size_t)q * sizeof(int32_t)));
    TRY(dev_alloc(e, e->q_idf, (size_t)q * MAX_TERMS * sizeof(double)));
    TRY(dev_alloc(e, e->qnorm_tab, (size_t)q * sizeof(QNorm)));
    TRY(dev_alloc(e, e->rec_base, (size_t)q * sizeof(int64_t)));
    TRY(dev_alloc(e, e->slot_terms, (size_t)q * MAX_TERMS * sizeof(int32_t)));
    TRY(dev_alloc(e, e->col_lo, (size_t)tl * sizeof(float)));
    TRY(dev_alloc(e, e->col_hi, (size_t)tl * sizeof(float)));
    e->col_comp = -1;                  // the per-tile column extremes are re-derived with the cached column
    e->tile_ld = tl;
    TRY(dev_alloc(e, e->sel_thr, (size_t)q * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->surv_count, (size_t)q * sizeof(int)));
    TRY(dev_alloc(e, e->surv_keys, (size_t)q * SURV_CAP * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->surv_ids, (size_t)q * SURV_CAP * sizeof(int64_t)));
    TRY(dev_alloc(e, e->gate, (size_t)q * sizeof(int)));
    TRY(dev_alloc(e, e->witness, (size_t)q * sizeof(int32_t)));
    TRY(dev_alloc(e, e->last_keys, (size_t)q * sizeof(uint64_t)));
    if (q > e->qt_cap) {
        if (e->h_q) { cudaFreeHost(e->h_q); cudaFreeHost(e->h_qt); cudaFreeHost(e->h_q2); cudaFreeHost(e->h_top_ids);
                      cudaFreeHost(e->h_top_scores); cudaFreeHost(e->h_small); cudaFreeHost(e->h_last_keys);
                      cudaFreeHost(e->h_rec_base); cudaFreeHost(e->h_slot_terms); }
        CK(cudaMallocHost((void**)&e->h_rec_base, (size_t)q * sizeof(int64_t)));
        CK(cudaMallocHost((void**)&e->h_slot_terms, (size_t)q * MAX_TERMS * sizeof(int32_t)));
        CK(cudaMallocHost((void**)&e->h_q, (size_t)q * DIM * sizeof(float)));
        CK(cudaMallocHost((void**)&e->h_q2, (size_t)q * DIM * sizeof(float)));
        CK(cudaMallocHost((void**)&e->h_qt, (size_t)q * sizeof(QueryTerms)));
        CK(cudaMallocHost((void**)&e->h_top_ids, (size_t)q * MAX_DEPTH * sizeof(int64_t)));
        CK(cudaMallocHost((void**)&e->h_top_scores, (size_t)q * MAX_DEPTH * sizeof(double)));
        CK(cudaMallocHost((void**)&e->h_small, (size_t)q * 3 * sizeof(int32_t)));
        CK(cudaMallocHost((void**)&e->h_last_keys, (size_t)q * sizeof(uint64_t)));
        e->sel_k_cap = 0;       // candidate buffers are sized per query too
        e->out_topn_cap = 0;
    }
    e->qt_cap = q;
    e->ld = l;
    return AIS_OK;
}

int ensure_sel(ais_engine* e, int k) {
    if (k <= e->sel_k_cap) return AIS_OK;
    const size_t q = (size_t)e->qt_cap;
    const size_t G = (size_t)sel_blocks_cap(e, k);
    const size_t ng = (size_t)merge_groups((int)(4 * e->sm_count)) + 64;
    TRY(dev_alloc(e, e->blk_keys, q * G * k * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->blk_ids, q * G * k * sizeof(int64_t)));
    TRY(dev_alloc(e, e->grp_keys, q * ng * k * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->grp_ids, q * ng * k * sizeof(int64_t)));
    TRY(dev_alloc(e, e->cand_keys, q * k * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->cand_ids, q * k * sizeof(int64_t)));
    TRY(dev_alloc(e, e->rest_keys, q * k * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->rest_ids, q * k * sizeof(int64_t)));
    TRY(dev_alloc(e, e->rest_count, q * sizeof(int32_t)));
    e->sel_k_cap = k;
    return AIS_OK;
}

int ensure_out(ais_engine* e, int topn) {
    if (topn <= e->out_topn_cap) return AIS_OK;
    const size_t q = (size_t)e->qt_cap;
    TRY(dev_alloc(e, e->out_ids, q * topn * sizeof(int64_t)));
    TRY(dev_alloc(e, e->out_scores, q * topn * sizeof(double)));
    if (e->h_out_ids) { cudaFreeHost(e->h_out_ids); cudaFreeHost(e->h_out_scores); }
    CK(cudaMallocHost((void**)&e->h_out_ids, q * topn * sizeof(int64_t)));
    CK(cudaMallocHost((void**)&e->h_out_scores, q * topn * sizeof(double)));
    e->out_topn_cap = topn;
    return AIS_OK;
}


// CUDA-event bracket around the launches of one kernel class on the engine's stream (only while profiling is on);
// ais_get_stats sums the pairs per class - bench.py's roofline.kernels[] comes from here.
struct ProfScope {
    ais_engine* e; int kind; cudaEvent_t a = nullptr;
    ProfScope(ais_engine* e_, int kind_) : e(e_), kind(kind_) {
        if (!e->profiling) return;
        if (!e->ev_free.empty()) { a = e->ev_free.back(); e->ev_free.pop_back(); }
        else if (cudaEventCreate(&a) != cudaSuccess) { a = nullptr; return; }
        cudaEventRecord(a, e->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEvent_t b = nullptr;
        if (!e->ev_free.empty()) { b = e->ev_free.back(); e->ev_free.pop_back(); }
        else if (cudaEventCreate(&b) != cudaSuccess) { e->ev_free.push_back(a); return; }
        cudaEventRecord(b, e->stream);
        e->ev_pending.push_back({kind, a, b});
        e->kind_launches[kind]++;
    }
};

CombineParams combine_params(const ais_engine* e) {
    CombineParams cp;
    cp.wb = e->p.bm25_weight;
    cp.wd = (float)e->p.doc2vec_weight;
    cp.wo = e->p.original_score_weight;
    cp.wr = (float)e->p.reranked_score_weight;
    return cp;
}

template <int QT>
int launch_scan_t(ais_engine* e, const float* d_q, int nq, float* out, uint32_t* max_keys) {
    const int64_t n_tiles = (e->n_vec + TILE_ROWS - 1) / TILE_ROWS;
    int grid = (int)(n_tiles < e->sm_count ? n_tiles : e->sm_count);
    ProfScope prof(e, AIS_KIND_SCAN);
    scan_kernel<QT><<<grid, ScanCfg<QT>::THREADS, scan_smem_bytes<QT>(), e->stream>>>(
        e->rows.as<float>(), e->n_vec, d_q, out, e->ld, max_keys, nq, 1);
    LAUNCHED(e);
    e->scan_launches++;
    return AIS_OK;
}

template <int QT>
int launch_scan_mma_t(ais_engine* e, const float* d_q, int nq, float* out, uint32_t* max_keys) {
    const int64_t n_tiles = (e->n_vec + TILE_ROWS - 1) / TILE_ROWS;
    int grid = (int)(n_tiles < e->sm_count ? n_tiles : e->sm_count);
    ProfScope prof(e, AIS_KIND_SCAN);
    scan_mma_kernel<QT><<<grid, MMA_THREADS, scan_mma_smem_bytes<QT>(), e->stream>>>(e->rows.as<float>(), e->n_vec, d_q, out, e->ld,
                                                                                    max_keys, nq);
    LAUNCHED(e);
    e->scan_launches++;
    return AIS_OK;
}

// ---- tcgen05 scan: tensor maps (driver entry point fetched through the runtime, no libcuda link) --------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// [n_rows][300] fp32 row-major -> boxes of [box_rows][32 columns] landing in the K-major SWIZZLE_128B layout
int make_row_tmap(CUtensorMap* tm, const void* ptr, int64_t n_rows, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(AIS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)DIM, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, elem,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AIS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return AIS_OK;
}

template <int N, int MAIN, int CROSS, int RAW, int ASTG, int ASUB, int NBUF>
void launch_tc_variant(ais_engine* e, int grid, float* out, uint32_t* max_keys, int nq) {
    scan_tc_kernel<N, MAIN, CROSS, RAW, ASTG, ASUB, NBUF><<<grid, TC_THREADS, tc_smem_bytes(N, RAW), e->stream>>>(e->tm_rows, e->tm_q[N == 64], e->n_vec, out,
                                                                                           e->ld, max_keys, nq);
}

// one tcgen05 pass over the rows for up to 32 (wide = false) or 64 (wide = true) queries
int launch_scan_tc(ais_engine* e, const float* d_q, int nq, bool wide, float* out, uint32_t* max_keys) {
    if (e->n_vec >= (1LL << 31)) return fail(AIS_ERR_INVALID, "tcgen05 scan: shard larger than 2^31 docs");
    const int n_pass = wide ? 64 : 32;
    TRY(dev_alloc(e, e->qsplit, (size_t)2 * TC_N_MAX * DIM * sizeof(float)));
    if (e->tm_q_ptr != e->qsplit.p) {
        TRY(make_row_tmap(&e->tm_q[0], e->qsplit.p, 2 * 32, 32));
        TRY(make_row_tmap(&e->tm_q[1], e->qsplit.p, 2 * 64, 64));
        TRY(make_row_tmap(&e->tm_q[2], e->qsplit.p, 2 * 64, TCP_NH));
        e->tm_q_ptr = e->qsplit.p;
    }
    if (e->tm_rows_ptr != e->rows.p || e->tm_rows_n != e->n_vec) {
        TRY(make_row_tmap(&e->tm_rows, e->rows.p, e->n_vec, TC_M));
        e->tm_rows_ptr = e->rows.p;
        e->tm_rows_n = e->n_vec;
    }
    split_queries_kernel<<<(n_pass * DIM + 255) / 256, 256, 0, e->stream>>>(d_q, nq, n_pass, e->qsplit.as<float>());
    LAUNCHED(e);
    const int64_t n_tiles = (e->n_vec + TC_M - 1) / TC_M;
    const int grid = (int)(n_tiles < e->sm_count ? n_tiles : e->sm_count);
    ProfScope prof(e, AIS_KIND_SCAN);
    if (wide && e->scan_pair > 0 && e->sm_count >= 2) {
        // CTA pairs: half of the query images per SM, a ring of 8 / 9 landing stages (scan_pair.cuh)
        const int64_t n_steps = (n_tiles + 1) / 2;
        const int pairs = (int)(n_steps < e->sm_count / 2 ? n_steps : e->sm_count / 2);
        if (e->scan_pair >= 9)
            scan_pair_kernel<9><<<2 * pairs, TC_THREADS, tcp_smem_bytes(9), e->stream>>>(e->tm_rows, e->tm_q[2], e->n_vec, out, e->ld, max_keys, nq);
        else
            scan_pair_kernel<8><<<2 * pairs, TC_THREADS, tcp_smem_bytes(8), e->stream>>>(e->tm_rows, e->tm_q[2], e->n_vec, out, e->ld, max_keys, nq);
        e->pair_scan_launches++;
    } else if (wide) launch_tc_variant<64, 2, 1, 4, 2, 1, 2>(e, grid, out, max_keys, nq);
    else launch_tc_variant<32, 3, 1, 6, 4, 1, 2>(e, grid, out, max_keys, nq);
    LAUNCHED(e);
    e->scan_launches++;
    return AIS_OK;
}

int launch_scan_one(ais_engine* e, const float* d_q, int nq, float* out, uint32_t* max_keys) {
    if (nq <= 1) return launch_scan_t<1>(e, d_q, nq, out, max_keys);
    if (nq <= 2) return launch_scan_t<2>(e, d_q, nq, out, max_keys);
    if (nq <= 4) return launch_scan_t<4>(e, d_q, nq, out, max_keys);
    if (e->use_mma) return nq <= 8 ? launch_scan_mma_t<8>(e, d_q, nq, out, max_keys) : launch_scan_mma_t<16>(e, d_q, nq, out, max_keys);
    if (nq <= 8) return launch_scan_t<8>(e, d_q, nq, out, max_keys);
    return launch_scan_t<16>(e, d_q, nq, out, max_keys);
}

struct RerRef { const float* base; int64_t qstride; const float* scale; };
RerRef rer_ref(const ais_engine* e) {
    if (e->rer_column) return {e->colbuf.as<float>(), 0, e->d_q2.as<float>() + e->col_comp};
    return {e->rer.as<float>(), e->ld, nullptr};
}

// column mode of the re-query: make sure colbuf holds column `comp` of the current rows (extracted once per index),
// together with its extreme values per 256-doc tile (the pass-2 collect bounds the blend R of a tile with them)
int prepare_column(ais_engine* e, int comp) {
    TRY(dev_alloc(e, e->colbuf, (size_t)(e->n_vec > 0 ? e->n_vec : 1) * sizeof(float)));
    if (e->col_comp != comp || e->col_rows_ptr != e->rows.p || e->col_n != e->n_vec) {
        if (e->n_vec > 0) {
            extract_column_kernel<<<(unsigned)((e->n_vec + 255) / 256), 256, 0, e->stream>>>(e->rows.as<float>(), e->n_vec, comp,
                                                                                        e->colbuf.as<float>());
            LAUNCHED(e);
            const int64_t n_tiles = (e->n_vec + SEL_TILE - 1) / SEL_TILE;
            int64_t blocks = (n_tiles + 7) / 8;
            if (blocks > 8LL * e->sm_count) blocks = 8LL * e->sm_count;
            column_bounds_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(e->colbuf.as<float>(), e->n_vec, n_tiles,
                                                                         e->col_lo.as<float>(), e->col_hi.as<float>());
            LAUNCHED(e);
        }
        e->col_comp = comp; e->col_rows_ptr = e->rows.p; e->col_n = e->n_vec;
    }
    return AIS_OK;
}

// queries with a single non-zero component `comp` (the reference's PRF re-query): one sector per doc
int launch_scan_column(ais_engine* e, const float* d_q, int nq, int comp, float* out, uint32_t* max_keys) {
    if (e->n_vec == 0) return AIS_OK;
    int64_t gx = (e->n_vec + COL_THREADS - 1) / COL_THREADS;
    if (gx > 8LL * e->sm_count) gx = 8LL * e->sm_count;
    column_scan_kernel<<<dim3((unsigned)gx, (unsigned)((nq + COL_QC - 1) / COL_QC)), COL_THREADS, 0, e->stream>>>(
        e->rows.as<float>(), e->n_vec, comp, d_q, nq, out, e->ld, max_keys);
    LAUNCHED(e);
    return AIS_OK;
}

// one pass over the rows per MAX_QT queries (the query buffer is zero-padded to a power of two >= nq)
int launch_scan(ais_engine* e, const float* d_q, int nq, float* out, uint32_t* max_keys) {
    if (e->n_vec == 0) return AIS_OK;
    for (int q0 = 0; q0 < nq;) {
        const int left = nq - q0;
        const bool tc = e->tc_min > 0 && left >= e->tc_min;
        const bool wide = tc && e->tc_wide && left > 32;
        const int cap = tc ? (wide ? 64 : 32) : MAX_QT;
        const int m = left < cap ? left : cap;
        if (tc) TRY(launch_scan_tc(e, d_q + (size_t)q0 * DIM, m, wide, out + (size_t)q0 * e->ld, max_keys + q0));
        else TRY(launch_scan_one(e, d_q + (size_t)q0 * DIM, m, out + (size_t)q0 * e->ld, max_keys + q0));
        q0 += m;
    }
    return AIS_OK;
}

int set_scan_attrs() {
    CK(cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes<1>()));
    CK(cudaFuncSetAttribute(scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes<2>()));
    CK(cudaFuncSetAttribute(scan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes<4>()));
    CK(cudaFuncSetAttribute(scan_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes<8>()));
    CK(cudaFuncSetAttribute(scan_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes<16>()));
    CK(cudaFuncSetAttribute(scan_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_mma_smem_bytes<8>()));
    CK(cudaFuncSetAttribute(scan_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_mma_smem_bytes<16>()));
    CK(cudaFuncSetAttribute(scan_tc_kernel<32, 3, 1, 6, 4, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(32, 6)));
    CK(cudaFuncSetAttribute(scan_tc_kernel<64, 2, 1, 4, 2, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(64, 4)));
    CK(cudaFuncSetAttribute(scan_pair_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp_smem_bytes(8)));
    CK(cudaFuncSetAttribute(scan_pair_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp_smem_bytes(9)));
    CK(cudaFuncSetAttribute(sort_survivors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SURV_CAP * 16));
    return AIS_OK;
}

// Query records -> pinned staging -> device.  Also sizes the BM25 record pool of the batch: the records of a tile
// (docs whose BM25 value is not the query's default) sit at the offset "postings of the query's terms before the tile",
// so a query's region holds the sum of its terms' document frequencies.
int upload_queries(ais_engine* e, const ais_query* qs, int nq, bool with_vec, bool with_terms) {
    for (int i = 0; i < nq; ++i) {
        if (with_vec) {
            if (!qs[i].vec) return fail(AIS_ERR_INVALID, "query %d: vec is NULL", i);
            memcpy(e->h_q + (size_t)i * DIM, qs[i].vec, DIM * sizeof(float));
        }
        if (with_terms) {
            if (qs[i].n_terms < 0 || qs[i].n_terms > MAX_TERMS)
                return fail(AIS_ERR_INVALID, "query %d: n_terms %d outside [0, %d]", i, qs[i].n_terms, MAX_TERMS);
            QueryTerms& t = e->h_qt[i];
            t.n_terms = qs[i].n_terms;
            t.n_required = 0;
            for (int j = 0; j < qs[i].n_terms; ++j) {
                t.term[j] = qs[i].term_ids[j];
                t.slot[j] = -1;
                t.weight[j] = qs[i].weights[j];
                t.n_required += qs[i].weights[j] > e->p.require_magic;        // webui.py:161 (1000 itself is NOT required)
            }
        }
    }
    if (with_vec) {
        const int padded = next_pow2_int(nq);
        for (int i = nq; i < padded; ++i) memset(e->h_q + (size_t)i * DIM, 0, DIM * sizeof(float));
        CK(cudaMemcpyAsync(e->d_q.p, e->h_q, (size_t)padded * DIM * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    }
    if (with_terms) {
        // the distinct terms of the batch get one presence bitmap each (bitmap path of the BM25 side; tf == 1 indexes)
        e->n_slots = 0;
        e->use_bits = false;
        const int64_t n_tiles = ((e->n_bm25 > 0 ? e->n_bm25 : 0) + SEL_TILE - 1) / SEL_TILE;
        if (!e->has_tf && e->want_bits && n_tiles > 0) {
            std::unordered_map<int32_t, int32_t> slot_of;
            for (int i = 0; i < nq; ++i)
                for (int j = 0; j < qs[i].n_terms; ++j) {
                    const int32_t t = qs[i].term_ids[j];
                    if (t < 0 || t >= e->n_vocab) continue;
                    auto it = slot_of.find(t);
                    if (it == slot_of.end()) {
                        it = slot_of.emplace(t, e->n_slots).first;
                        e->h_slot_terms[e->n_slots++] = t;
                    }
                    e->h_qt[i].slot[j] = it->second;
                }
            e->use_bits = (int64_t)e->n_slots * n_tiles * 32 <= e->bitmap_cap_bytes;
        }
        CK(cudaMemcpyAsync(e->d_qt.p, e->h_qt, (size_t)nq * sizeof(QueryTerms), cudaMemcpyHostToDevice, e->stream));
        if (e->use_bits) {
            if (e->n_slots > 0) {
                TRY(dev_alloc(e, e->term_bits, (size_t)e->n_slots * n_tiles * 32));
                CK(cudaMemcpyAsync(e->slot_terms.p, e->h_slot_terms, (size_t)e->n_slots * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
            }
            return AIS_OK;                               // no record pool on this path
        }
        int64_t total = 0;

        for (int i = 0; i < nq; ++i) {
            int64_t touched = 0;
            for (int j = 0; j < qs[i].n_terms; ++j) {
                const int32_t t = qs[i].term_ids[j];
                if (t >= 0 && t < e->n_vocab) touched += e->h_post_ptr[(size_t)t + 1] - e->h_post_ptr[(size_t)t];
            }
            if (touched >= (1LL << 32))
                return fail(AIS_ERR_UNSUPPORTED, "query %d: its terms hold %lld postings on this shard (32-bit record offsets)", i, (long long)touched);
            e->h_rec_base[i] = total;
            total += touched;
        }
        if ((size_t)total * sizeof(double) > e->rec_val.cap) {          // grow with headroom: batches differ in size
            const size_t want = (size_t)total + (size_t)total / 4 + 1024;
            TRY(dev_alloc(e, e->rec_val, want * sizeof(double)));
            TRY(dev_alloc(e, e->rec_pos, want));
        }
        CK(cudaMemcpyAsync(e->rec_base.p, e->h_rec_base, (size_t)nq * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
    }
    return AIS_OK;
}

// where the combined scores of the current batch come from (finals.cuh)
FinSrc fin_src(const ais_engine* e, const double* d_maxes) {
    FinSrc S;
    memset(&S, 0, sizeof(S));
    S.fin_ext = e->ext_fin ? e->fin_ext.as<double>() : nullptr;
    S.sim = e->sim.as<float>();
    S.ld = e->ld;
    S.use_bits = e->use_bits ? 1 : 0;
    S.ieee_div = e->ieee_div ? 1 : 0;
    S.B.bits = e->term_bits.as<uint8_t>();
    S.B.n_tiles = (e->n() + SEL_TILE - 1) / SEL_TILE;
    S.B.queries = e->d_qt.as<QueryTerms>();
    S.B.q_idf = e->q_idf.as<double>();
    S.B.g1 = e->g1.as<double>();
    S.B.magic = e->p.require_magic;
    S.tile_hdr = e->tile_hdr.as<uint32_t>();
    S.tile_off = e->tile_off.as<uint32_t>();
    S.tile_ld = e->tile_ld;
    S.rec_val = e->rec_val.as<double>();
    S.rec_pos = e->rec_pos.as<uint8_t>();
    S.rec_base = e->rec_base.as<int64_t>();
    S.maxes = d_maxes;
    S.n_required = e->q_nreq.as<int32_t>();
    S.qtab = (e->qnorm_ready && d_maxes == e->cur_maxes) ? e->qnorm_tab.as<QNorm>() : nullptr;
    S.wb = e->p.bm25_weight;
    S.wd = (float)e->p.doc2vec_weight;
    S.n = e->n();
    return S;
}

Bm25Args bm25_args(ais_engine* e, int64_t n_sub) {
    Bm25Args a;
    memset(&a, 0, sizeof(a));
    a.slices = e->bm25_slices.as<int64_t>();
    a.t_cap = e->bm25_t_cap;
    a.n_sub = n_sub;
    a.post_doc = e->post_doc.as<int32_t>();
    a.post_tf = e->has_tf ? e->post_tf.as<int32_t>() : nullptr;
    a.kd = e->kd.as<double>();
    a.g1 = e->g1.as<double>();
    a.post_len = e->max_doc_len >= 0 ? e->post_len.as<uint16_t>() : nullptr;
    a.kd_tab = e->kd_tab.as<double>();
    a.g1_tab = e->g1_tab.as<double>();
    a.n = e->n_bm25;
    a.queries = e->d_qt.as<QueryTerms>();
    a.q_idf = e->q_idf.as<double>();
    a.magic = e->p.require_magic;
    a.k1p1 = e->p.k1 + 1.0;
    a.ld = e->ld;
    a.tile_hdr = e->tile_hdr.as<uint32_t>();
    a.tile_off = e->tile_off.as<uint32_t>();
    a.rec_val = e->rec_val.as<double>();
    a.rec_pos = e->rec_pos.as<uint8_t>();
    a.rec_base = e->rec_base.as<int64_t>();
    a.tile_ld = e->tile_ld;
    return a;
}

// the BM25 side of a batch: slice table, per-tile records, per-query maximum (dense scores only for the
// compute_bm25_scores seam)
int launch_bm25_max(ais_engine* e, int nq, double* dense_out) {
    if (e->n_bm25 <= 0) return AIS_OK;
    const int64_t n_sub = (e->n_bm25 + BM25_SUB - 1) / BM25_SUB;
    if (e->use_bits) {
        // bitmap path: one streaming pass over the posting list of every DISTINCT term of the batch, then the per-query
        // maxima from the bitmaps (the records / slice table of the general path below do not exist on this path)
        e->bitmap_batches++;
        {
            ProfScope prof(e, AIS_KIND_BM25_SLICES);
            query_idf_kernel<<<(nq * MAX_TERMS + 255) / 256, 256, 0, e->stream>>>(e->d_qt.as<QueryTerms>(), nq, e->idf.as<double>(),
                                                                                 e->n_vocab, e->q_idf.as<double>(), e->q_nreq.as<int32_t>());
            LAUNCHED(e);
            if (e->n_slots > 0) {
                CK(cudaMemsetAsync(e->term_bits.p, 0, (size_t)e->n_slots * n_sub * 32, e->stream));
                int64_t longest = 0;
                for (int s = 0; s < e->n_slots; ++s) {
                    const int64_t df = e->h_post_ptr[(size_t)e->h_slot_terms[s] + 1] - e->h_post_ptr[(size_t)e->h_slot_terms[s]];
                    longest = df > longest ? df : longest;
                }
                int64_t chunks = (longest + BITMAP_THREADS * 8 - 1) / (BITMAP_THREADS * 8);
                if (chunks > 2LL * e->sm_count) chunks = 2LL * e->sm_count;
                if (chunks < 1) chunks = 1;
                bm25_bitmap_kernel<<<dim3((unsigned)chunks, (unsigned)e->n_slots), BITMAP_THREADS, 0, e->stream>>>(
                    e->post_ptr.as<int64_t>(), e->post_doc.as<int32_t>(), e->slot_terms.as<int32_t>(), n_sub, e->term_bits.as<uint32_t>());
                LAUNCHED(e);
            }
        }
        ProfScope prof(e, AIS_KIND_BM25_SCORE);
        const FinSrc S = fin_src(e, nullptr);
        bm25_max_bits_kernel<<<dim3((unsigned)n_sub, (unsigned)((nq + BITQ_WARPS - 1) / BITQ_WARPS)), 32 * BITQ_WARPS, 0, e->stream>>>(
            S.B, e->n_bm25, nq, e->maxb_key.as<uint64_t>(), dense_out, e->ld);
        LAUNCHED(e);
        return AIS_OK;
    }
    int t_cap = 1;
    for (int q = 0; q < nq; ++q) t_cap = e->h_qt[q].n_terms > t_cap ? e->h_qt[q].n_terms : t_cap;
    e->bm25_t_cap = t_cap;
    TRY(dev_alloc(e, e->bm25_slices, (size_t)e->qt_cap * t_cap * (n_sub + 1) * sizeof(int64_t)));
    {
    ProfScope prof(e, AIS_KIND_BM25_SLICES);
    bm25_slices_kernel<<<dim3((unsigned)((n_sub + 1 + 127) / 128), (unsigned)(nq * t_cap)), 128, 0, e->stream>>>(
        e->post_ptr.as<int64_t>(), e->post_doc.as<int32_t>(), e->n_vocab, e->d_qt.as<QueryTerms>(), e->idf.as<double>(), t_cap, n_sub,
        e->bm25_slices.as<int64_t>(), e->q_nreq.as<int32_t>(), e->q_idf.as<double>());
    LAUNCHED(e);
    }
    ProfScope prof(e, AIS_KIND_BM25_SCORE);
    Bm25Args a = bm25_args(e, n_sub);
    a.max_keys = e->maxb_key.as<uint64_t>();
    a.dense_out = dense_out;
    const dim3 gs((unsigned)((n_sub + BM25_WARPS - 1) / BM25_WARPS), (unsigned)nq);
    if (e->score_occ >= 8) bm25_score_kernel<8><<<gs, BM25_THREADS, BM25_SMEM, e->stream>>>(a);
    else if (e->score_occ >= 6) bm25_score_kernel<6><<<gs, BM25_THREADS, BM25_SMEM, e->stream>>>(a);
    else if (e->score_occ == 5) bm25_score_kernel<5><<<gs, BM25_THREADS, BM25_SMEM, e->stream>>>(a);
    else bm25_score_kernel<4><<<gs, BM25_THREADS, BM25_SMEM, e->stream>>>(a);
    LAUNCHED(e);
    return AIS_OK;
}

// two-level merge of candidate lists into out[nq][k_out] (sorted best first, KEY_EMPTY padded)
int merge_lists(ais_engine* e, const uint64_t* keys, const int64_t* ids, int n_lists, int64_t list_stride,
                int64_t q_stride, int k_in, int k_out, int nq, uint64_t* out_keys, int64_t* out_ids, int32_t* out_count,
                const int* gate = nullptr) {
    if (n_lists > 2 * MERGE_GROUP) {
        const int ng = merge_groups(n_lists);
        if ((size_t)e->qt_cap * ng * k_out * sizeof(uint64_t) > e->grp_keys.cap)
            return fail(AIS_ERR_INVALID, "too many candidate lists to merge (%d)", n_lists);
        merge_kernel<<<dim3(ng, nq), SEL_THREADS, 0, e->stream>>>(keys, ids, n_lists, MERGE_GROUP, list_stride, q_stride,
                                                                 k_in, k_out, e->grp_keys.as<uint64_t>(),
                                                                 e->grp_ids.as<int64_t>(), k_out, nullptr, gate);
        LAUNCHED(e);
        merge_kernel<<<dim3(1, nq), SEL_THREADS, 0, e->stream>>>(e->grp_keys.as<uint64_t>(), e->grp_ids.as<int64_t>(), ng, ng,
                                                                k_out, (int64_t)ng * k_out, k_out, k_out, out_keys, out_ids,
                                                                k_out, out_count, gate);
        LAUNCHED(e);
    } else {
        merge_kernel<<<dim3(1, nq), SEL_THREADS, 0, e->stream>>>(keys, ids, n_lists, n_lists, list_stride, q_stride, k_in,
                                                                k_out, out_keys, out_ids, k_out, out_count, gate);
        LAUNCHED(e);
    }
    return AIS_OK;
}

// everything the select kernels share for the current batch; pass2: the blend R with the re-query scores
SelectArgs select_args(ais_engine* e, bool pass2) {
    SelectArgs a;
    memset(&a, 0, sizeof(a));
    a.S = fin_src(e, e->cur_maxes);
    if (pass2) {
        const RerRef rr = rer_ref(e);
        a.rer = rr.base; a.rer_qstride = rr.qstride; a.rer_scale = rr.scale;
    }
    a.n = e->n();
    a.id_base = e->first_doc;
    a.cp = combine_params(e);
    a.seeds_all = e->top_ids.as<int64_t>();
    a.depth = pass2 ? e->p.prf_depth : 0;
    // segments = groups of whole tiles
    a.n_tiles = (a.n + SEL_TILE - 1) / SEL_TILE;
    const int64_t tiles_per_seg = a.n_tiles > 0 ? (a.n_tiles + SEG_MAX - 1) / SEG_MAX : 1;
    a.tiles_per_seg = (int)tiles_per_seg;
    a.n_seg = (int)((a.n_tiles + tiles_per_seg - 1) / tiles_per_seg);
    e->last_tiles_per_seg = tiles_per_seg;
    a.seg_max = e->seg_max.as<uint64_t>();
    a.tile_max = e->tile_max.as<uint64_t>();
    a.col_lo = e->col_lo.as<float>();
    a.col_hi = e->col_hi.as<float>();
    a.max_all = pass2 ? e->maxr_key.as<uint64_t>() : nullptr;
    a.thr = e->sel_thr.as<uint64_t>();
    a.surv_count = e->surv_count.as<int>();
    a.surv_keys = e->surv_keys.as<uint64_t>();
    a.surv_ids = e->surv_ids.as<int64_t>();
    a.gate = e->gate.as<int>();
    return a;
}

// Exact local top-k of one scoring stage -> out[nq][k] (sorted best first, KEY_EMPTY padded).
//   mode 0: pass 1 from the dot scores + BM25 records (bm25_combine_kernel emits the tile / segment maxima),
//   mode 1: pass 1 from supplied combined scores (ais_rerank), 2: pass 2, the blend R.
// Fast path: tile / segment maxima -> threshold -> collect -> sort (select2.cuh); the streaming buffer select
// follows, gated per query on the overflow flag the fast path raises.
int local_select(ais_engine* e, int mode, int nq, int k, uint64_t* d_keys, int64_t* d_ids) {
    if (k < 1 || k > SEL_KMAX) return fail(AIS_ERR_INVALID, "k %d outside [1, %d]", k, SEL_KMAX);
    TRY(ensure_sel(e, k));
    const int64_t n = e->n();
    SelectArgs a = select_args(e, mode == 2);
    const bool bound = mode == 2 && e->bound_ok;         // pass 2 (AIS_TILE_BOUND=1): per-tile upper bound in the collect, no maxima pass
    if (bound) e->bound_passes++;
    // pass 2 on the records path: rerank_max_kernel (seeds not excluded from its segment maxima -> `depth` more segments
    // counted by the threshold); `skip`: with the tile bound on R, from the threshold of the pass-1 candidates
    const bool fast = !e->use_bits && !e->ext_fin && !getenv("AIS_NO_RERANK_MAX");     // records path: two-class tile evaluation
    const bool records2 = mode == 2 && !bound && fast;
    const bool skip = records2 && e->ub_valid && !e->no_skip;
    if (n > 0 && !bound) {
        ProfScope prof(e, AIS_KIND_COMBINE);
        CK(cudaMemsetAsync(e->seg_max.p, 0, (size_t)nq * SEG_MAX * sizeof(uint64_t), e->stream));
        if (mode == 0) {
            CombineArgs c;
            c.S = a.S;
            c.n_sub = a.n_tiles;
            c.seg_max = a.seg_max; c.seg_stride = SEG_MAX; c.tiles_per_seg = a.tiles_per_seg;
            c.tile_max = a.tile_max;
            if (e->use_bits)
                bm25_combine_bits_kernel<<<dim3((unsigned)a.n_tiles, (unsigned)((nq + BITQ_WARPS - 1) / BITQ_WARPS)), 32 * BITQ_WARPS, 0, e->stream>>>(c, nq);
            else {
                const dim3 gc((unsigned)((a.n_tiles + BM25C_WARPS - 1) / BM25C_WARPS), (unsigned)nq);
                if (e->combine_occ >= 8) bm25_combine_kernel<8><<<gc, BM25C_THREADS, 0, e->stream>>>(c);
                else if (e->combine_occ >= 6) bm25_combine_kernel<6><<<gc, BM25C_THREADS, 0, e->stream>>>(c);
                else bm25_combine_kernel<5><<<gc, BM25C_THREADS, 0, e->stream>>>(c);
            }
            LAUNCHED(e);
        } else {
            if (mode == 2) {                                 // the tile table of R is separate: pass 1's stays valid
                TRY(dev_alloc(e, e->tile_max2, (size_t)e->qt_cap * e->tile_ld * sizeof(uint64_t)));
                a.tile_max = e->tile_max2.as<uint64_t>();
            }
            const dim3 g1((unsigned)a.n_seg, (unsigned)((nq + SEG_WARPS - 1) / SEG_WARPS));
            if (mode == 1) segmax_kernel<1><<<g1, 32 * SEG_WARPS, 0, e->stream>>>(a, nq);
            else if (records2) {                             // records path: the light pass over the scaled records
                const dim3 g2((unsigned)((a.n_tiles + 31) / 32), (unsigned)((nq + SEG_WARPS - 1) / SEG_WARPS));
                if (skip) {                                  // threshold from the pass-1 candidates first: tiles that cannot reach it are skipped
                    rerank_threshold_kernel<<<nq, 256, 0, e->stream>>>(e->p1_keys.as<uint64_t>(), e->p1_ids.as<int64_t>(), e->p1_k, a, k);
                    LAUNCHED(e);
                    rerank_max_kernel<1><<<g2, 32 * SEG_WARPS, 0, e->stream>>>(a, nq, e->tile_max.as<uint64_t>());
                    e->skip_passes++;
                } else {
                    rerank_max_kernel<0><<<g2, 32 * SEG_WARPS, 0, e->stream>>>(a, nq, nullptr);
                }
            } else segmax_kernel<2><<<g1, 32 * SEG_WARPS, 0, e->stream>>>(a, nq);
            LAUNCHED(e);
        }
    }
    ProfScope prof(e, AIS_KIND_SELECT);
    if (bound) {
        rerank_threshold_kernel<<<nq, 256, 0, e->stream>>>(e->p1_keys.as<uint64_t>(), e->p1_ids.as<int64_t>(), e->p1_k, a, k);
        LAUNCHED(e);
    } else {
        if (n <= 0) CK(cudaMemsetAsync(e->seg_max.p, 0, (size_t)nq * SEG_MAX * sizeof(uint64_t), e->stream));
        threshold_kernel<<<nq, 256, 0, e->stream>>>(a.seg_max, a.n_seg, records2 ? k + e->p.prf_depth : k, a.thr, a.surv_count, a.gate,
                                                    skip ? 1 : 0);
        LAUNCHED(e);
    }
    if (n > 0) {
        int64_t cb = (a.n_tiles + COLLECT_THREADS - 1) / COLLECT_THREADS;     // one lane per tile
        if (cb > 2LL * e->sm_count) cb = 2LL * e->sm_count;
        const dim3 g3((unsigned)cb, (unsigned)nq);
        if (bound) collect_kernel<2, 1><<<g3, COLLECT_THREADS, 0, e->stream>>>(a);
        else if (fast && mode != 2) collect_fast_kernel<1><<<g3, COLLECT_THREADS, 0, e->stream>>>(a);
        else if (fast) collect_fast_kernel<2><<<g3, COLLECT_THREADS, 0, e->stream>>>(a);
        else if (mode != 2) collect_kernel<1, 0><<<g3, COLLECT_THREADS, 0, e->stream>>>(a);
        else collect_kernel<2, 0><<<g3, COLLECT_THREADS, 0, e->stream>>>(a);
        LAUNCHED(e);
    }
    sort_survivors_kernel<<<nq, 512, SURV_CAP * 16, e->stream>>>(a.surv_count, a.surv_keys, a.surv_ids, k, d_keys, d_ids, a.gate);
    LAUNCHED(e);
    // gated fallback (runs only for queries whose survivors overflowed)
    const int G = sel_blocks(e, k);
    if (mode != 2) stream_select_kernel<1><<<dim3(G, nq), SEL_THREADS, 0, e->stream>>>(a, k, e->blk_keys.as<uint64_t>(), e->blk_ids.as<int64_t>());
    else stream_select_kernel<2><<<dim3(G, nq), SEL_THREADS, 0, e->stream>>>(a, k, e->blk_keys.as<uint64_t>(), e->blk_ids.as<int64_t>());
    LAUNCHED(e);
    return merge_lists(e, e->blk_keys.as<uint64_t>(), e->blk_ids.as<int64_t>(), G, k, (int64_t)G * k, k, k, nq, d_keys, d_ids,
                       nullptr, e->gate.as<int>());
}

// combined scores of one query of the current batch -> e->scratch64 [n] (test seams only)
int materialize_finals(ais_engine* e, int qi, const double* d_maxes) {
    TRY(dev_alloc(e, e->scratch64, (size_t)e->ld * sizeof(double)));
    if (e->n() == 0) return AIS_OK;
    const int64_t n_tiles = (e->n() + SEL_TILE - 1) / SEL_TILE;
    int64_t blocks = (n_tiles + 7) / 8;
    if (blocks > 8LL * e->sm_count) blocks = 8LL * e->sm_count;
    materialize_finals_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(fin_src(e, d_maxes), qi, n_tiles, e->scratch64.as<double>());
    LAUNCHED(e);
    return AIS_OK;
}

// ------------------------------------------------------------------------------------------------
int do_score(ais_engine* e, const ais_query* qs, int nq, double* d_maxes) {
    TRY(check_loaded(e));
    TRY(ensure_work(e));
    if (nq < 1 || nq > e->p.max_batch) return fail(AIS_ERR_INVALID, "nq %d outside [1, max_batch=%d]", nq, e->p.max_batch);
    TRY(upload_queries(e, qs, nq, true, true));
    init_keys_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(e->maxs_key.as<uint32_t>(), e->maxb_key.as<uint64_t>(),
                                             e->maxr_key.as<uint64_t>(), e->status.as<int32_t>(), nq, 1);
    LAUNCHED(e);
    TRY(launch_bm25_max(e, nq, nullptr));
    TRY(launch_scan(e, e->d_q.as<float>(), nq, e->sim.as<float>(), e->maxs_key.as<uint32_t>()));
    maxes_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(e->maxb_key.as<uint64_t>(), e->maxs_key.as<uint32_t>(), nq, d_maxes);
    LAUNCHED(e);
    e->cur_nq = nq;
    e->cur_prf = false;
    e->rer_column = false;
    e->ext_fin = false;
    e->bound_ok = false;
    e->qnorm_ready = false;
    e->p1_k = 0;
    return AIS_OK;
}

// pass 1: this shard's best k docs by combined score.  d_maxes: the GLOBAL maxima [nq][2] (must stay valid until the
// batch is finished: the later stages recompute combined scores from them).  The list is also kept inside the engine:
// the pass-2 threshold of the collapsed re-query starts from it.
int do_combine(ais_engine* e, int nq, const double* d_maxes, int k, uint64_t* d_keys, int64_t* d_ids) {
    e->cur_maxes = d_maxes;
    e->qnorm_ready = false;
    if (!e->ext_fin) {                                   // the per-query constants of webui.py:376-383, once per batch
        qnorm_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(fin_src(e, d_maxes), nq, e->qnorm_tab.as<QNorm>());
        LAUNCHED(e);
        e->qnorm_ready = true;
    }
    TRY(local_select(e, e->ext_fin ? 1 : 0, nq, k, d_keys, d_ids));
    TRY(dev_alloc(e, e->p1_keys, (size_t)e->qt_cap * SEL_KMAX * sizeof(uint64_t)));
    TRY(dev_alloc(e, e->p1_ids, (size_t)e->qt_cap * SEL_KMAX * sizeof(int64_t)));
    CK(cudaMemcpyAsync(e->p1_keys.p, d_keys, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(e->p1_ids.p, d_ids, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToDevice, e->stream));
    e->p1_k = k;
    return AIS_OK;
}

int do_top(ais_engine* e, int nq, int n_lists, int k, const uint64_t* d_keys, const int64_t* d_ids, int64_t* out_top_ids,
           double* out_top_scores, float* d_rows) {
    const int depth = e->p.prf_depth;
    TRY(ensure_sel(e, k > depth ? k : depth));
    ProfScope prof(e, AIS_KIND_REQUERY);
    const uint64_t* mk = d_keys;
    const int64_t* mi = d_ids;
    int stride = k;
    if (n_lists > 1) {
        TRY(merge_lists(e, d_keys, d_ids, n_lists, (int64_t)nq * k, k, k, depth, nq, e->rest_keys.as<uint64_t>(),
                        e->rest_ids.as<int64_t>(), nullptr));
        mk = e->rest_keys.as<uint64_t>();
        mi = e->rest_ids.as<int64_t>();
        stride = depth;
    } else if (k < depth) {
        return fail(AIS_ERR_INVALID, "stage_top: k %d < prf_depth %d", k, depth);
    }
    top_unpack_kernel<<<nq, 32, 0, e->stream>>>(mk, mi, stride, depth, e->top_ids.as<int64_t>(), e->top_scores.as<double>());
    LAUNCHED(e);
    if (d_rows) {
        gather_top_rows_kernel<<<dim3(depth, nq), 128, 0, e->stream>>>(e->rows.as<float>(), e->n(), e->first_doc,
                                                                      e->top_ids.as<int64_t>(), depth, d_rows);
        LAUNCHED(e);
    }
    if (out_top_ids || out_top_scores) {
        CK(cudaMemcpyAsync(e->h_top_ids, e->top_ids.p, (size_t)nq * MAX_DEPTH * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->h_top_scores, e->top_scores.p, (size_t)nq * MAX_DEPTH * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        for (int q = 0; q < nq; ++q)
            for (int t = 0; t < depth; ++t) {
                if (out_top_ids) out_top_ids[q * depth + t] = e->h_top_ids[q * MAX_DEPTH + t];
                if (out_top_scores) out_top_scores[q * depth + t] = e->h_top_scores[q * MAX_DEPTH + t];
            }
    }
    return AIS_OK;
}

int do_requery_select(ais_engine* e, int nq, int k, uint64_t* d_keys, int64_t* d_ids) {
    // the max is re-accumulated (idempotent under atomicMax)
    return local_select(e, 2, nq, k, d_keys, d_ids);
}

int do_requery(ais_engine* e, int nq, const float* q2_host, const float* d_rows, int prf_mode, int k, double* d_max_r,
               uint64_t* d_keys, int64_t* d_ids) {
    const int depth = e->p.prf_depth;
    // Does every re-query vector have at most one non-zero component, the same one for the whole batch?  The collapsed
    // centroid of the reference always does (component 0); a caller-supplied vector is inspected.
    int comp = -1;
    {
    ProfScope prof(e, AIS_KIND_REQUERY);
    if (q2_host) {
        comp = -2;                                               // -2: no non-zero seen yet
        for (int q = 0; q < nq && comp != -1; ++q)
            for (int j = 0; j < DIM; ++j)
                if (q2_host[(size_t)q * DIM + j] != 0.0f) {
                    if (comp == -2) comp = j;
                    else if (comp != j) { comp = -1; break; }
                }
        if (comp == -2) comp = 0;
        const int padded = next_pow2_int(nq);
        memcpy(e->h_q2, q2_host, (size_t)nq * DIM * sizeof(float));
        for (int i = nq; i < padded; ++i) memset(e->h_q2 + (size_t)i * DIM, 0, DIM * sizeof(float));
        CK(cudaMemcpyAsync(e->d_q2.p, e->h_q2, (size_t)padded * DIM * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    } else {
        if (!d_rows) return fail(AIS_ERR_INVALID, "stage_requery: neither q2 nor d_rows given");
        if (prf_mode != AIS_PRF_STORED_ROWS && prf_mode != AIS_PRF_STORED_ROWS_FULL)
            return fail(AIS_ERR_INVALID, "stage_requery: device re-query needs a STORED_ROWS prf_mode");
        const int padded = next_pow2_int(nq);
        if (padded > nq)
            CK(cudaMemsetAsync(e->d_q2.as<float>() + (size_t)nq * DIM, 0, (size_t)(padded - nq) * DIM * sizeof(float), e->stream));
        prf_query_kernel<<<nq, 320, 0, e->stream>>>(d_rows, e->top_scores.as<double>(), depth,
                                                   prf_mode == AIS_PRF_STORED_ROWS ? 1 : 0, e->d_q2.as<float>(),
                                                   e->status.as<int32_t>());
        LAUNCHED(e);
        if (prf_mode == AIS_PRF_STORED_ROWS) comp = 0;          // [c, 0, ..., 0] by construction (or all zero)
    }
    if (e->requery_dense) comp = -1;
    init_keys_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(e->maxs_key.as<uint32_t>(), e->maxb_key.as<uint64_t>(),
                                             e->maxr_key.as<uint64_t>(), e->status.as<int32_t>(), nq, 2);
    LAUNCHED(e);
    e->rer_column = comp >= 0;
    if (comp >= 0) {                                     // no per-query array: col[d] * c_q on the fly
        TRY(prepare_column(e, comp));
        e->column_scan_launches++;
    } else {                                             // dense re-query: the only case that holds a second score array
        TRY(dev_alloc(e, e->rer, (size_t)e->qt_cap * e->ld * sizeof(float)));
        TRY(launch_scan(e, e->d_q2.as<float>(), nq, e->rer.as<float>(), e->maxs_key.as<uint32_t>()));
    }
    // The per-tile bound on R (select2.cuh, collect_kernel<2, 1>) needs the column form, non-negative blend weights
    // (every rounding of wo * fin + wr * (col * c) is then monotone) and this shard's pass-1 candidates.
    e->ub_valid = comp >= 0 && e->p.original_score_weight > 0.0 && e->p.reranked_score_weight >= 0.0 && e->p1_k > 0 && e->n() > 0;
    e->bound_ok = e->ub_valid && !e->no_bound;
    }
    TRY(do_requery_select(e, nq, k, d_keys, d_ids));
    maxr_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(e->maxr_key.as<uint64_t>(), nq, d_max_r);
    LAUNCHED(e);
    e->cur_prf = true;
    return AIS_OK;
}

int copy_results(ais_engine* e, int nq, int topn, int64_t* out_ids, double* out_scores, int32_t* out_counts,
                 int32_t* out_status, int32_t* out_amb, uint64_t* out_last_keys) {
    CK(cudaMemcpyAsync(e->h_out_ids, e->out_ids.p, (size_t)nq * topn * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_out_scores, e->out_scores.p, (size_t)nq * topn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_small, e->out_count.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_small + e->qt_cap, e->status.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_small + 2 * e->qt_cap, e->out_amb.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_last_keys, e->last_keys.p, (size_t)nq * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int q = 0; q < nq; ++q) {
        if (out_last_keys) out_last_keys[q] = e->h_last_keys[q];
        const int st = e->h_small[e->qt_cap + q];
        const int cnt = st == AIS_Q_OK ? e->h_small[q] : 0;
        if (out_counts) out_counts[q] = cnt;
        if (out_status) out_status[q] = st;
        if (out_amb) out_amb[q] = st == AIS_Q_OK ? e->h_small[2 * e->qt_cap + q] : 0;
        if (out_ids) memcpy(out_ids + (size_t)q * topn, e->h_out_ids + (size_t)q * topn, (size_t)cnt * sizeof(int64_t));
        if (out_scores) memcpy(out_scores + (size_t)q * topn, e->h_out_scores + (size_t)q * topn, (size_t)cnt * sizeof(double));
    }
    return AIS_OK;
}

// d_max_r == NULL: the no-PRF branch (webui.py:247-253) - candidates are sorted finals, no pinned docs
int do_finish(ais_engine* e, int nq, int n_lists, int k, const uint64_t* d_keys, const int64_t* d_ids, const double* d_max_r,
              const int32_t* d_witness, int topn, int64_t* out_ids, double* out_scores, int32_t* out_counts, int32_t* out_status,
              int32_t* out_amb, uint64_t* out_last_keys) {
    if (topn < 1) return fail(AIS_ERR_INVALID, "topn must be >= 1");
    const int k_out = n_lists == 1 ? k : std::min<int64_t>(SEL_KMAX, (int64_t)n_lists * k);
    TRY(ensure_sel(e, k_out));
    TRY(ensure_out(e, topn));
    ProfScope prof(e, AIS_KIND_TAIL);
    if (n_lists == 1) {
        copy_list_kernel<<<nq, 128, 0, e->stream>>>(d_keys, d_ids, k, e->rest_keys.as<uint64_t>(), e->rest_ids.as<int64_t>(),
                                                   e->rest_count.as<int32_t>());
        LAUNCHED(e);
    } else {
        TRY(merge_lists(e, d_keys, d_ids, n_lists, (int64_t)nq * k, k, k, k_out, nq, e->rest_keys.as<uint64_t>(),
                        e->rest_ids.as<int64_t>(), e->rest_count.as<int32_t>()));
        if (k_out > k) {
            prefix_bound_kernel<<<nq, 256, 0, e->stream>>>(d_keys, n_lists, (int64_t)nq * k, k, k, e->rest_keys.as<uint64_t>(),
                                                          k_out, e->rest_count.as<int32_t>());
            LAUNCHED(e);
        }
    }
    TailParams tp;
    tp.thresh = e->p.diff_filter_thresh;
    tp.topn = topn;
    tp.depth = d_max_r ? e->p.prf_depth : 0;
    tp.normalize = d_max_r ? 1 : 0;
    tp.n_total = e->total();
    tail_kernel<<<nq, SEL_THREADS, 0, e->stream>>>(e->rest_keys.as<uint64_t>(), e->rest_ids.as<int64_t>(), k_out,
                                                  e->rest_count.as<int32_t>(), nullptr, e->top_ids.as<int64_t>(), d_max_r, tp,
                                                  d_witness, e->out_ids.as<int64_t>(), e->out_scores.as<double>(),
                                                  e->out_count.as<int32_t>(), e->out_amb.as<int32_t>(), e->last_keys.as<uint64_t>());
    LAUNCHED(e);
    return copy_results(e, nq, topn, out_ids, out_scores, out_counts, out_status, out_amb, out_last_keys);
}

constexpr int64_t WITNESS_BUCKETS = 1LL << 22;

// For every ambiguous query: is there a pair of distinct scores closer than the threshold anywhere at or
// below the last prefix entry?  d_witness[q] = 1 if this shard holds such a pair (sufficient, not necessary).
int do_witness(ais_engine* e, int nq, const int32_t* amb, const uint64_t* last_keys, int second_pass, const double* d_max_r,
               int32_t* d_witness) {
    ProfScope prof(e, AIS_KIND_WITNESS);
    CK(cudaMemsetAsync(d_witness, 0, (size_t)nq * sizeof(int32_t), e->stream));
    if (e->n() == 0) return AIS_OK;
    TRY(dev_alloc(e, e->wit_table, (size_t)WITNESS_BUCKETS * sizeof(uint64_t)));
    for (int q = 0; q < nq; ++q) {
        if (!amb[q]) continue;
        CK(cudaMemsetAsync(e->wit_table.p, 0, (size_t)WITNESS_BUCKETS * sizeof(uint64_t), e->stream));
        WitnessArgs w;
        w.a = select_args(e, second_pass != 0);
        w.a.depth = second_pass ? e->p.prf_depth : 0;
        w.qi = q;
        w.second_pass = second_pass ? 1 : 0;
        w.last_key = last_keys[q];
        w.max_r = d_max_r ? d_max_r + q : nullptr;
        w.normalize = d_max_r ? 1 : 0;
        w.thresh = e->p.diff_filter_thresh;
        w.inv_thresh = 1.0 / e->p.diff_filter_thresh;
        w.table = e->wit_table.as<uint64_t>();
        w.n_buckets = WITNESS_BUCKETS;
        w.flag = d_witness + q;
        int64_t blocks = (w.a.n_tiles + 7) / 8;
        if (blocks > 8LL * e->sm_count) blocks = 8LL * e->sm_count;
        witness_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(w);
        LAUNCHED(e);
    }
    return AIS_OK;
}

int64_t sort_capacity(int64_t n) {
    int64_t p = GS_TILE;
    while (p < n) p <<= 1;
    return p;
}

// sort n_pad (power of two >= GS_TILE) entries best-first
int bitonic_sort(ais_engine* e, uint64_t* keys, int64_t* ids, int64_t n_pad) {
    const unsigned blocks = (unsigned)(n_pad / GS_TILE);
    bitonic_local_kernel<<<blocks, GS_THREADS, 0, e->stream>>>(keys, ids, n_pad, 2ull, (unsigned long long)GS_TILE);
    LAUNCHED(e);
    for (unsigned long long size = 2ull * GS_TILE; size <= (unsigned long long)n_pad; size <<= 1) {
        for (unsigned long long stride = size >> 1; stride >= (unsigned long long)GS_TILE; stride >>= 1) {
            bitonic_global_kernel<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, e->stream>>>(keys, ids, n_pad, size, stride);
            LAUNCHED(e);
        }
        bitonic_local_kernel<<<blocks, GS_THREADS, 0, e->stream>>>(keys, ids, n_pad, size, size);
        LAUNCHED(e);
    }
    return AIS_OK;
}

// write this shard's keys of query qi (pass 2: R, seeds blanked; pass 1: finals) into caller arrays [n_local]
int do_export_keys(ais_engine* e, int qi, int second_pass, uint64_t* d_keys, int64_t* d_ids) {
    if (e->n() == 0) return AIS_OK;
    SelectArgs a = select_args(e, second_pass != 0);
    int64_t blocks = (a.n_tiles + 7) / 8;
    if (blocks > 8LL * e->sm_count) blocks = 8LL * e->sm_count;
    fill_keys_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(a, qi, second_pass ? 1 : 0, d_keys, d_ids);
    LAUNCHED(e);
    return AIS_OK;
}

// full sort of n_entries keys (all shards' exports, concatenated by the caller into arrays of
// ais_sort_capacity(n_entries) slots) and the exact tail for query qi
int do_sort_finish(ais_engine* e, int qi, uint64_t* d_keys, int64_t* d_ids, int64_t n_entries, const double* d_max_r, int topn,
                   int64_t* out_ids, double* out_scores, int32_t* out_count, int32_t* out_status) {
    TRY(ensure_out(e, topn));
    TRY(dev_alloc(e, e->fs_count, sizeof(int64_t) * 2));
    ProfScope prof(e, AIS_KIND_WITNESS);
    const int64_t n_pad = sort_capacity(n_entries);
    if (n_pad > n_entries) {
        fill_empty_kernel<<<(unsigned)((n_pad - n_entries + 255) / 256), 256, 0, e->stream>>>(d_keys, d_ids, n_entries, n_pad);
        LAUNCHED(e);
    }
    TRY(bitonic_sort(e, d_keys, d_ids, n_pad));
    CK(cudaMemsetAsync(e->fs_count.p, 0, sizeof(int64_t), e->stream));
    count_live_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, e->stream>>>(d_keys, n_pad, e->fs_count.as<int64_t>());
    LAUNCHED(e);
    TailParams tp;
    tp.thresh = e->p.diff_filter_thresh;
    tp.topn = topn;
    tp.depth = d_max_r ? e->p.prf_depth : 0;
    tp.normalize = d_max_r ? 1 : 0;
    tp.n_total = e->total();
    // single-query launch: offset every per-query array to qi
    tail_kernel<<<1, SEL_THREADS, 0, e->stream>>>(d_keys, d_ids, 0, nullptr, e->fs_count.as<int64_t>(),
                                                 e->top_ids.as<int64_t>() + (size_t)qi * MAX_DEPTH,
                                                 d_max_r ? d_max_r + qi : nullptr, tp, nullptr, e->out_ids.as<int64_t>(),
                                                 e->out_scores.as<double>(), e->out_count.as<int32_t>(),
                                                 e->out_amb.as<int32_t>(), nullptr);
    LAUNCHED(e);
    e->fullsort_fallbacks++;
    CK(cudaMemcpyAsync(e->h_out_ids, e->out_ids.p, (size_t)topn * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_out_scores, e->out_scores.p, (size_t)topn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_small, e->out_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_small + 1, e->status.as<int32_t>() + qi, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const int st = e->h_small[1];
    const int cnt = st == AIS_Q_OK ? e->h_small[0] : 0;
    if (out_count) *out_count = cnt;
    if (out_status) *out_status = st;
    if (out_ids) memcpy(out_ids, e->h_out_ids, (size_t)cnt * sizeof(int64_t));
    if (out_scores) memcpy(out_scores, e->h_out_scores, (size_t)cnt * sizeof(double));
    return AIS_OK;
}

// numpy's float64 pairwise sum of < 128 items (np.average's wgt.sum())
double np_sum_host(const double* a, int n) {
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// How many candidates a select stage asks for.  The stages NEED `need` (topn + 1: the filter looks one entry past the
// cut), but a longer exact prefix costs next to nothing (a few hundred more tiles recomputed, a larger survivor sort) and
// lets tail_kernel see the second near-tie of filter_searched_result (webui.py:66-77) inside the prefix instead of
// sending the query to the witness pass: at 10 M docs the best 1 024 scores hold ~10 gaps below 1e-6.  `cap`: the longest
// list the stage may return.  Kept below half the segment count so that the segment-maximum threshold stays selective.
int select_depth(const ais_engine* e, int need, int cap) {
    const int64_t n_tiles = (e->n() + SEL_TILE - 1) / SEL_TILE;
    const int64_t tps = n_tiles > 0 ? (n_tiles + SEG_MAX - 1) / SEG_MAX : 1;
    const int64_t n_seg = (n_tiles + tps - 1) / tps;
    int64_t deep = n_seg / 2;
    if (e->sel_deep >= 0 && deep > e->sel_deep) deep = e->sel_deep;
    int k = need > deep ? need : (int)deep;
    if (k > cap) k = cap;
    return k < 1 ? 1 : k;
}

// One batch on one GPU.  qs == NULL: the combined scores were supplied (ais_rerank, e->fin_ext).
int run_batch(ais_engine* e, const ais_query* qs, int q_index0, int nq, int topn, int prf_mode, ais_infer_cb cb, void* ctx,
              int64_t* out_ids, double* out_scores, int32_t* out_counts, int32_t* out_status) {
    const int depth = e->p.prf_depth;
    TRY(check_loaded(e));
    TRY(ensure_work(e));            // every buffer exists BEFORE its pointer is taken
    TRY(ensure_sel(e, SEL_KMAX));
    double* maxes = e->maxes_own.as<double>();
    if (qs) TRY(do_score(e, qs, nq, maxes));
    const bool prf = prf_mode != AIS_PRF_OFF && e->total() > depth;     // webui.py:193 `len(sims) > 10`
    std::vector<int32_t> amb(nq, 0), status(nq, 0);
    std::vector<uint64_t> last_keys(nq, 0);
    auto any_amb = [&]() { bool any = false; for (int q = 0; q < nq; ++q) any = any || amb[q]; return any; };
    uint64_t* ck = e->cand_keys.as<uint64_t>();
    int64_t* ci = e->cand_ids.as<int64_t>();
    // a result longer than the selector can return: the exact full-sort path below serves every query
    const bool too_long = prf ? (topn + 1 - depth > SEL_KMAX - depth) : (topn + 1 > SEL_KMAX);
    if (!prf) {
        const int k = too_long ? 1 : select_depth(e, topn + 1, SEL_KMAX);
        TRY(do_combine(e, nq, maxes, k, ck, ci));
        if (!too_long) {
            TRY(do_finish(e, nq, 1, k, ck, ci, nullptr, nullptr, topn, out_ids, out_scores, out_counts, out_status, amb.data(),
                          last_keys.data()));
            if (any_amb()) {       // one near-tie inside the prefix: look for a second one anywhere below it
                TRY(do_witness(e, nq, amb.data(), last_keys.data(), 0, nullptr, e->witness.as<int32_t>()));
                TRY(do_finish(e, nq, 1, k, ck, ci, nullptr, e->witness.as<int32_t>(), topn, out_ids, out_scores, out_counts,
                              out_status, amb.data(), nullptr));
            }
        }
    } else {
        // pass 2 returns k2 docs next to the `depth` pinned seeds; pass 1 returns the seeds and as many docs again
        // (the pass-2 threshold of the collapsed re-query is the k2-th best blend among them)
        int k2 = topn + 1 - depth;
        k2 = too_long ? SEL_KMAX - depth : select_depth(e, k2 < 1 ? 1 : k2, SEL_KMAX - depth);
        const int k1 = depth + k2;
        TRY(do_combine(e, nq, maxes, k1, ck, ci));
        const bool host_q2 = prf_mode == AIS_PRF_CALLBACK;
        std::vector<int64_t> top_ids;
        std::vector<double> top_scores;
        std::vector<float> q2;
        if (host_q2) {
            if (!cb) return fail(AIS_ERR_INVALID, "AIS_PRF_CALLBACK needs a callback");
            top_ids.resize((size_t)nq * depth);
            top_scores.resize((size_t)nq * depth);
            q2.assign((size_t)nq * DIM, 0.0f);
        }
        TRY(do_top(e, nq, 1, k1, ck, ci, host_q2 ? top_ids.data() : nullptr, host_q2 ? top_scores.data() : nullptr,
                   host_q2 ? nullptr : e->rows_own.as<float>()));
        if (host_q2) {
            bool any_bad = false;
            for (int q = 0; q < nq; ++q) {
                const double* w = &top_scores[(size_t)q * depth];
                bool bad = false;
                for (int t = 0; t < depth; ++t) bad = bad || !isfinite(w[t]);
                if (bad) status[q] = AIS_Q_NAN_WEIGHTS;                       // webui.py:200-203
                else if (np_sum_host(w, depth) == 0.0) status[q] = AIS_Q_ZERO_WEIGHT_SUM;
                else if (cb(ctx, q_index0 + q, &top_ids[(size_t)q * depth], w, depth, &q2[(size_t)q * DIM]) != 0) {
                    status[q] = AIS_Q_CALLBACK_FAILED;
                    memset(&q2[(size_t)q * DIM], 0, DIM * sizeof(float));
                }
                any_bad = any_bad || status[q] != 0;
            }
            if (any_bad)
                CK(cudaMemcpyAsync(e->status.p, status.data(), (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        }
        double* maxr = e->maxr_own.as<double>();
        TRY(do_requery(e, nq, host_q2 ? q2.data() : nullptr, host_q2 ? nullptr : e->rows_own.as<float>(), prf_mode, k2, maxr, ck, ci));
        if (!too_long) {
            TRY(do_finish(e, nq, 1, k2, ck, ci, maxr, nullptr, topn, out_ids, out_scores, out_counts, out_status, amb.data(),
                          last_keys.data()));
            if (any_amb()) {
                TRY(do_witness(e, nq, amb.data(), last_keys.data(), 1, maxr, e->witness.as<int32_t>()));
                TRY(do_finish(e, nq, 1, k2, ck, ci, maxr, e->witness.as<int32_t>(), topn, out_ids, out_scores, out_counts,
                              out_status, amb.data(), nullptr));
            }
        }
    }
    if (too_long) {
        TRY(ensure_out(e, topn));
        for (int q = 0; q < nq; ++q) amb[q] = 1;
    }
    // exact fallback: the filter outcome depends on scores beyond the selected prefix (or topn exceeds what the selector
    // returns) -> sort every doc's key
    for (int q = 0; q < nq; ++q) {
        if (!amb[q]) continue;
        const int64_t cap = sort_capacity(e->n());
        TRY(dev_alloc(e, e->fs_keys, (size_t)cap * sizeof(uint64_t)));
        TRY(dev_alloc(e, e->fs_ids, (size_t)cap * sizeof(int64_t)));
        TRY(do_export_keys(e, q, prf ? 1 : 0, e->fs_keys.as<uint64_t>(), e->fs_ids.as<int64_t>()));
        TRY(do_sort_finish(e, q, e->fs_keys.as<uint64_t>(), e->fs_ids.as<int64_t>(), e->n(), prf ? e->maxr_own.as<double>() : nullptr,
                           topn, out_ids ? out_ids + (size_t)q * topn : nullptr, out_scores ? out_scores + (size_t)q * topn : nullptr,
                           out_counts ? out_counts + q : nullptr, out_status ? out_status + q : nullptr));
    }
    return AIS_OK;
}

int need_whole_index(const ais_engine* e) {
    if (e->total() != e->n() || e->first_doc != 0)
        return fail(AIS_ERR_INVALID, "this engine holds a shard (%lld of %lld docs): drive it through the ais_stage_* calls",
                    (long long)e->n(), (long long)e->total());
    return AIS_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* ais_last_error(void) { return g_err; }
int ais_abi_version(void) { return AIS_ABI_VERSION; }
int ais_max_select_k(void) { return SEL_KMAX; }
int64_t ais_sort_capacity(int64_t n_entries) { return sort_capacity(n_entries); }

void ais_default_params(ais_params* p) {
    if (!p) return;
    p->k1 = 1.5;
    p->b = 0.75;
    p->bm25_weight = 0.5;
    p->doc2vec_weight = 0.5;
    p->original_score_weight = 0.7;
    p->reranked_score_weight = 0.3;
    p->diff_filter_thresh = 1e-6;
    p->require_magic = 1000.0;
    p->prf_depth = 10;
    p->max_batch = 1;
}

static int validate_params(const ais_params* p) {
    if (p->prf_depth < 1 || p->prf_depth > MAX_DEPTH) return fail(AIS_ERR_INVALID, "prf_depth %d outside [1, %d]", p->prf_depth, MAX_DEPTH);
    if (p->max_batch < 1 || p->max_batch > MAX_BATCH) return fail(AIS_ERR_INVALID, "max_batch %d outside [1, %d]", p->max_batch, MAX_BATCH);
    return AIS_OK;
}

int ais_create(ais_engine** out, int device_id, const ais_params* p) {
    if (!out) return fail(AIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(AIS_ERR_CUDA, "no CUDA device available (%s): the ais_b200 engine has no CPU path", cudaGetErrorString(ce));
    if (device_id < 0 || device_id >= count) return fail(AIS_ERR_INVALID, "device_id %d outside [0, %d)", device_id, count);
    ais_params pp;
    ais_default_params(&pp);
    if (p) pp = *p;
    TRY(validate_params(&pp));
    DeviceGuard g(device_id);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10)
        return fail(AIS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device_id, prop.major, prop.minor);
    ais_engine* e = new ais_engine();
    e->device = device_id;
    e->sm_count = prop.multiProcessorCount;
    e->p = pp;
    cudaError_t se = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete e; return fail(AIS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(se)); }
    e->stream = e->own_stream;
    const char* simt = getenv("AIS_SCAN_SIMT");
    e->use_mma = !(simt && simt[0] == '1');
    if (const char* tcm = getenv("AIS_SCAN_TC_MIN")) e->tc_min = atoi(tcm);
    if (const char* tcw = getenv("AIS_SCAN_TC_WIDE")) e->tc_wide = atoi(tcw) != 0;
    if (const char* sp = getenv("AIS_SCAN_PAIR")) e->scan_pair = atoi(sp);
    if (const char* rd = getenv("AIS_REQUERY_DENSE")) e->requery_dense = atoi(rd) != 0;
    if (const char* nb = getenv("AIS_TILE_BOUND")) e->no_bound = atoi(nb) == 0;
    if (const char* ns = getenv("AIS_NO_TILE_SKIP")) e->no_skip = atoi(ns) != 0;

    if (const char* sd = getenv("AIS_SELECT_DEPTH")) e->sel_deep = atoi(sd);
    if (const char* fr = getenv("AIS_BM25_BITMAP")) e->want_bits = atoi(fr) != 0;
    if (const char* so = getenv("AIS_SCORE_OCC")) e->score_occ = atoi(so);
    if (const char* co = getenv("AIS_COMBINE_OCC")) e->combine_occ = atoi(co);
    if (const char* dv = getenv("AIS_IEEE_DIV")) e->ieee_div = atoi(dv) != 0;
    if (const char* bm = getenv("AIS_BM25_BITMAP_MB")) e->bitmap_cap_bytes = (int64_t)atoll(bm) << 20;

    int s = set_scan_attrs();
    if (s != AIS_OK) { cudaStreamDestroy(e->own_stream); delete e; return s; }
    if (e->scan_pair > 0) {
        // the CTA-pair scan needs two SMs of one TPC per cluster: ask the runtime whether such clusters can be resident at
        // all (MIG slices, odd SM counts); otherwise the single-CTA kernel serves the 64-query passes
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2, 1, 1);
        cfg.blockDim = dim3(TC_THREADS, 1, 1);
        cfg.dynamicSmemBytes = e->scan_pair >= 9 ? tcp_smem_bytes(9) : tcp_smem_bytes(8);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        const cudaError_t ce = e->scan_pair >= 9 ? cudaOccupancyMaxActiveClusters(&n_clusters, scan_pair_kernel<9>, &cfg)
                                                  : cudaOccupancyMaxActiveClusters(&n_clusters, scan_pair_kernel<8>, &cfg);
        if (ce != cudaSuccess || n_clusters < 1) { (void)cudaGetLastError(); e->scan_pair = 0; }
    }
    *out = e;
    return AIS_OK;
}

int ais_destroy(ais_engine* e) {
    if (!e) return AIS_OK;
    DeviceGuard g(e->device);
    cudaStreamSynchronize(e->stream);
    for (Buf* b : {&e->rows, &e->post_ptr, &e->post_doc, &e->post_tf, &e->idf, &e->kd, &e->g1, &e->doc_len, &e->sim, &e->scratch64, &e->fin_ext, &e->p1_keys, &e->p1_ids, &e->q_idf, &e->tile_off, &e->rec_val, &e->rec_pos, &e->rec_base, &e->qnorm_tab, &e->post_len, &e->kd_tab, &e->g1_tab, &e->term_bits, &e->slot_terms, &e->tile_max2, &e->col_lo, &e->col_hi,
                   &e->rer, &e->d_q, &e->d_q2, &e->d_qt, &e->maxs_key, &e->maxb_key, &e->maxr_key, &e->maxes_own, &e->maxr_own,
                   &e->top_ids, &e->top_scores, &e->status, &e->rows_own, &e->blk_keys, &e->blk_ids, &e->grp_keys, &e->grp_ids,
                   &e->cand_keys, &e->cand_ids, &e->rest_keys, &e->rest_ids, &e->rest_count, &e->out_ids, &e->out_scores,
                   &e->out_count, &e->out_amb, &e->fs_keys, &e->fs_ids, &e->fs_count, &e->seg_max, &e->tile_max, &e->tile_hdr, &e->q_nreq, &e->colbuf, &e->sel_thr, &e->surv_count,
                   &e->surv_keys, &e->surv_ids, &e->gate, &e->witness, &e->last_keys, &e->wit_table, &e->bm25_slices, &e->qsplit})
        dev_free(e, *b);
    for (void* h : {(void*)e->h_q, (void*)e->h_qt, (void*)e->h_q2, (void*)e->h_top_ids, (void*)e->h_top_scores,
                    (void*)e->h_out_ids, (void*)e->h_out_scores, (void*)e->h_small, (void*)e->h_last_keys, (void*)e->h_rec_base, (void*)e->h_slot_terms})
        if (h) cudaFreeHost(h);
    for (const ais_engine::PendingEv& p : e->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (cudaEvent_t ev : e->ev_free) cudaEventDestroy(ev);
    cudaStreamDestroy(e->own_stream);
    delete e;
    return AIS_OK;
}

int ais_set_params(ais_engine* e, const ais_params* p) {
    if (!e || !p) return fail(AIS_ERR_INVALID, "NULL argument");
    TRY(validate_params(p));
    DeviceGuard g(e->device);
    const bool kd_stale = (p->k1 != e->p.k1 || p->b != e->p.b) && e->n_bm25 > 0;
    e->p = *p;
    if (kd_stale) {
        kd_kernel<<<(unsigned)((e->n_bm25 + 255) / 256), 256, 0, e->stream>>>(e->doc_len.as<int64_t>(), e->n_bm25, e->avgdl, e->p.k1,
                                                                             e->p.b, 1.0 - e->p.b, e->p.k1 + 1.0, e->kd.as<double>(),
                                                                             e->g1.as<double>());
        LAUNCHED(e);
        TRY(build_len_tables(e, e->n_bm25, false));
    }
    return AIS_OK;
}

int ais_set_stream(ais_engine* e, void* cuda_stream) {
    if (!e) return fail(AIS_ERR_INVALID, "NULL engine");
    DeviceGuard g(e->device);
    CK(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return AIS_OK;
}

int ais_set_shard(ais_engine* e, int64_t first_doc_id, int64_t n_total_docs) {
    if (!e) return fail(AIS_ERR_INVALID, "NULL engine");
    if (first_doc_id < 0 || n_total_docs < 0) return fail(AIS_ERR_INVALID, "negative shard bounds");
    e->first_doc = first_doc_id;
    e->n_total = n_total_docs;
    return AIS_OK;
}

int ais_reserve_docs(ais_engine* e, int64_t n_docs) {
    if (!e || n_docs < 0) return fail(AIS_ERR_INVALID, "bad argument");
    DeviceGuard g(e->device);
    if (n_docs <= e->cap_vec) return AIS_OK;
    Buf nb;
    TRY(dev_alloc(e, nb, (size_t)n_docs * ROW_BYTES));
    if (e->n_vec > 0) CK(cudaMemcpyAsync(nb.p, e->rows.p, (size_t)e->n_vec * ROW_BYTES, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    dev_free(e, e->rows);
    e->rows = nb;
    e->cap_vec = n_docs;
    return AIS_OK;
}

int ais_load_vectors(ais_engine* e, const float* rows, int64_t n, int32_t dim, int64_t first_row) {
    if (e) e->col_comp = -1;                 // cached column of the row store (PRF re-query) is stale now
    if (!e || (!rows && n > 0) || n < 0 || first_row < 0) return fail(AIS_ERR_INVALID, "bad argument");
    if (dim != DIM) return fail(AIS_ERR_UNSUPPORTED, "vector dimension %d: only %d (genmodel.py:16 VECTOR_LENGTH) is built", dim, DIM);
    if (first_row > e->n_vec) return fail(AIS_ERR_INVALID, "first_row %lld leaves a gap after %lld loaded rows", (long long)first_row, (long long)e->n_vec);
    DeviceGuard g(e->device);
    if (first_row + n > e->cap_vec) {
        int64_t want = e->cap_vec + e->cap_vec / 2;
        if (want < first_row + n) want = first_row + n;
        TRY(ais_reserve_docs(e, want));
    }
    if (n > 0)
        CK(cudaMemcpyAsync(e->rows.as<float>() + (size_t)first_row * DIM, rows, (size_t)n * ROW_BYTES, cudaMemcpyDefault, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (first_row + n > e->n_vec) e->n_vec = first_row + n;
    return AIS_OK;
}

int ais_vectors_device_ptr(ais_engine* e, int64_t n_docs, float** out_rows) {
    if (e) e->col_comp = -1;                 // the caller is about to (re)write rows in place
    if (!e || !out_rows || n_docs < 0) return fail(AIS_ERR_INVALID, "bad argument");
    TRY(ais_reserve_docs(e, n_docs));
    e->n_vec = n_docs;
    *out_rows = e->rows.as<float>();
    return AIS_OK;
}

int ais_load_bm25(ais_engine* e, const int64_t* post_ptr, const int32_t* post_doc, const int32_t* post_tf, int32_t n_terms,
                  int64_t n_docs, const double* idf, const int64_t* doc_len, double avgdl) {
    if (!e || !post_ptr || !idf || (!doc_len && n_docs > 0) || n_terms < 0 || n_docs < 0) return fail(AIS_ERR_INVALID, "bad argument");
    if (n_docs >= (1LL << 31)) return fail(AIS_ERR_UNSUPPORTED, "a shard holds at most 2^31-1 docs (int32 local doc ids)");
    DeviceGuard g(e->device);
    TRY(dev_alloc(e, e->post_ptr, (size_t)(n_terms + 1) * sizeof(int64_t)));
    CK(cudaMemcpyAsync(e->post_ptr.p, post_ptr, (size_t)(n_terms + 1) * sizeof(int64_t), cudaMemcpyDefault, e->stream));
    e->h_post_ptr.assign((size_t)n_terms + 1, 0);
    CK(cudaMemcpyAsync(e->h_post_ptr.data(), e->post_ptr.p, (size_t)(n_terms + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const int64_t n_post = e->h_post_ptr[(size_t)n_terms];
    for (int32_t t = 0; t < n_terms; ++t)
        if (e->h_post_ptr[(size_t)t + 1] < e->h_post_ptr[(size_t)t]) return fail(AIS_ERR_INVALID, "post_ptr is not monotonic at term %d", t);
    if (e->h_post_ptr[0] != 0) return fail(AIS_ERR_INVALID, "post_ptr[0] must be 0");
    if (n_post < 0) return fail(AIS_ERR_INVALID, "post_ptr[n_terms] is negative");
    if (n_post > 0 && !post_doc) return fail(AIS_ERR_INVALID, "post_doc is NULL");
    TRY(dev_alloc(e, e->post_doc, (size_t)n_post * sizeof(int32_t)));
    if (n_post > 0) CK(cudaMemcpyAsync(e->post_doc.p, post_doc, (size_t)n_post * sizeof(int32_t), cudaMemcpyDefault, e->stream));
    e->has_tf = post_tf != nullptr;
    if (post_tf) {
        TRY(dev_alloc(e, e->post_tf, (size_t)n_post * sizeof(int32_t)));
        if (n_post > 0) CK(cudaMemcpyAsync(e->post_tf.p, post_tf, (size_t)n_post * sizeof(int32_t), cudaMemcpyDefault, e->stream));
    }
    TRY(dev_alloc(e, e->idf, (size_t)n_terms * sizeof(double)));
    if (n_terms > 0) CK(cudaMemcpyAsync(e->idf.p, idf, (size_t)n_terms * sizeof(double), cudaMemcpyDefault, e->stream));
    TRY(dev_alloc(e, e->doc_len, (size_t)n_docs * sizeof(int64_t)));
    TRY(dev_alloc(e, e->kd, (size_t)n_docs * sizeof(double)));
    TRY(dev_alloc(e, e->g1, (size_t)n_docs * sizeof(double)));
    if (n_docs > 0) {
        CK(cudaMemcpyAsync(e->doc_len.p, doc_len, (size_t)n_docs * sizeof(int64_t), cudaMemcpyDefault, e->stream));
        // webui.py:145  k1 * (1 - b + b * (dl / bm25_avgdl)), same operation order, no contraction
        kd_kernel<<<(unsigned)((n_docs + 255) / 256), 256, 0, e->stream>>>(e->doc_len.as<int64_t>(), n_docs, avgdl, e->p.k1, e->p.b,
                                                                          1.0 - e->p.b, e->p.k1 + 1.0, e->kd.as<double>(),
                                                                          e->g1.as<double>());
        LAUNCHED(e);
    }
    if (n_post > 0 && n_terms > 0) {                      // the kernels index per-tile arrays with these ids: check them once
        TRY(dev_alloc(e, e->scratch64, sizeof(double)));
        CK(cudaMemsetAsync(e->scratch64.p, 0, sizeof(int), e->stream));
        validate_postings_kernel<<<4 * e->sm_count, 256, 0, e->stream>>>(e->post_ptr.as<int64_t>(), e->post_doc.as<int32_t>(), n_terms,
                                                                        n_docs, e->scratch64.as<int>());
        LAUNCHED(e);
        int bad_term = 0;
        CK(cudaMemcpyAsync(&bad_term, e->scratch64.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (bad_term > 0) {
            e->n_vocab = 0; e->n_bm25 = 0; e->n_post = 0;
            return fail(AIS_ERR_INVALID, "posting list of term %d: doc ids must lie in [0, n_docs) and ascend strictly", bad_term - 1);
        }
    }
    CK(cudaStreamSynchronize(e->stream));
    e->avgdl = avgdl;
    e->n_vocab = n_terms;
    e->n_bm25 = n_docs;
    e->n_post = n_post;
    return build_len_tables(e, n_docs, true);
}

// ---- BM25 index build (genmodel.py:51-99) -------------------------------------------------------------------
int ais_build_bm25(ais_engine* e, const int64_t* seq_ptr, const int32_t* seq_ids, int64_t n_docs, int32_t n_terms,
                   int64_t* out_df, int64_t* out_doc_len) {
    if (!e || !seq_ptr || n_docs < 0 || n_terms < 1) return fail(AIS_ERR_INVALID, "bad argument");
    if (n_docs >= (1LL << 31)) return fail(AIS_ERR_UNSUPPORTED, "a shard holds at most 2^31-1 docs (int32 local doc ids)");
    DeviceGuard g(e->device);
    Buf d_ptr, d_ids, d_cnt, d_df, d_err;
    int st = AIS_OK;
    auto body = [&]() -> int {
        TRY(dev_alloc(e, d_ptr, (size_t)(n_docs + 1) * sizeof(int64_t)));
        CK(cudaMemcpyAsync(d_ptr.p, seq_ptr, (size_t)(n_docs + 1) * sizeof(int64_t), cudaMemcpyDefault, e->stream));
        int64_t n_tok = 0;
        CK(cudaMemcpyAsync(&n_tok, (const char*)d_ptr.p + (size_t)n_docs * sizeof(int64_t), sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (n_tok < 0 || (n_tok > 0 && !seq_ids)) return fail(AIS_ERR_INVALID, "bad token sequence");
        TRY(dev_alloc(e, d_ids, (size_t)n_tok * sizeof(int32_t)));
        if (n_tok > 0) CK(cudaMemcpyAsync(d_ids.p, seq_ids, (size_t)n_tok * sizeof(int32_t), cudaMemcpyDefault, e->stream));
        int64_t docs_per_cta = 1024;
        const int64_t max_ctas = 4LL * e->sm_count;
        if ((n_docs + docs_per_cta - 1) / docs_per_cta > max_ctas) docs_per_cta = (n_docs + max_ctas - 1) / max_ctas;
        const int n_ctas = (int)((n_docs + docs_per_cta - 1) / docs_per_cta);
        TRY(dev_alloc(e, d_cnt, (size_t)(n_ctas > 0 ? n_ctas : 1) * n_terms * sizeof(int32_t)));
        TRY(dev_alloc(e, d_df, (size_t)n_terms * sizeof(int64_t)));
        TRY(dev_alloc(e, d_err, sizeof(int)));
        TRY(dev_alloc(e, e->doc_len, (size_t)n_docs * sizeof(int64_t)));
        TRY(dev_alloc(e, e->post_ptr, (size_t)(n_terms + 1) * sizeof(int64_t)));
        CK(cudaMemsetAsync(d_cnt.p, 0, (size_t)(n_ctas > 0 ? n_ctas : 1) * n_terms * sizeof(int32_t), e->stream));
        CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), e->stream));
        if (n_ctas > 0) {
            build_count_kernel<<<n_ctas, BUILD_THREADS, 0, e->stream>>>(d_ptr.as<int64_t>(), d_ids.as<int32_t>(), n_docs, docs_per_cta,
                                                                       n_terms, d_cnt.as<int32_t>(), e->doc_len.as<int64_t>(), d_err.as<int>());
            LAUNCHED(e);
        }
        build_scan_kernel<<<(n_terms + 127) / 128, 128, 0, e->stream>>>(d_cnt.as<int32_t>(), n_ctas, n_terms, d_df.as<int64_t>());
        LAUNCHED(e);
        build_ptr_kernel<<<1, 32, 0, e->stream>>>(d_df.as<int64_t>(), n_terms, e->post_ptr.as<int64_t>());
        LAUNCHED(e);
        int err = 0;
        int64_t n_post = 0;
        CK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(&n_post, (const char*)e->post_ptr.p + (size_t)n_terms * sizeof(int64_t), sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (err == 1) return fail(AIS_ERR_UNSUPPORTED, "a doc has more than %d tags (one sort chunk of the builder)", BUILD_MAX_DOC_TAGS);
        if (err == 2) return fail(AIS_ERR_INVALID, "term id outside [0, %d)", n_terms);
        TRY(dev_alloc(e, e->post_doc, (size_t)n_post * sizeof(int32_t)));
        TRY(dev_alloc(e, e->post_tf, (size_t)n_post * sizeof(int32_t)));
        if (n_ctas > 0) {
            build_fill_kernel<<<n_ctas, BUILD_THREADS, 0, e->stream>>>(d_ptr.as<int64_t>(), d_ids.as<int32_t>(), n_docs, docs_per_cta, n_terms,
                                                                      d_cnt.as<int32_t>(), e->post_ptr.as<int64_t>(),
                                                                      e->post_doc.as<int32_t>(), e->post_tf.as<int32_t>());
            LAUNCHED(e);
        }
        if (out_df) CK(cudaMemcpyAsync(out_df, d_df.p, (size_t)n_terms * sizeof(int64_t), cudaMemcpyDefault, e->stream));
        if (out_doc_len && n_docs > 0)
            CK(cudaMemcpyAsync(out_doc_len, e->doc_len.p, (size_t)n_docs * sizeof(int64_t), cudaMemcpyDefault, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->has_tf = true;
        e->n_vocab = n_terms;
        e->n_post = n_post;
        e->n_bm25 = -1;                 // scoring needs ais_finish_bm25 (IDF table, avgdl) first
        e->n_built = n_docs;
        return AIS_OK;
    };
    st = body();
    for (Buf* b : {&d_ptr, &d_ids, &d_cnt, &d_df, &d_err}) dev_free(e, *b);
    return st;
}

int ais_finish_bm25(ais_engine* e, const double* idf, double avgdl) {
    if (!e || !idf) return fail(AIS_ERR_INVALID, "NULL argument");
    if (e->n_built < 0) return fail(AIS_ERR_NOT_LOADED, "ais_build_bm25 has not run");
    DeviceGuard g(e->device);
    const int64_t n_docs = e->n_built;
    TRY(dev_alloc(e, e->idf, (size_t)e->n_vocab * sizeof(double)));
    CK(cudaMemcpyAsync(e->idf.p, idf, (size_t)e->n_vocab * sizeof(double), cudaMemcpyDefault, e->stream));
    TRY(dev_alloc(e, e->kd, (size_t)n_docs * sizeof(double)));
    TRY(dev_alloc(e, e->g1, (size_t)n_docs * sizeof(double)));
    if (n_docs > 0) {
        kd_kernel<<<(unsigned)((n_docs + 255) / 256), 256, 0, e->stream>>>(e->doc_len.as<int64_t>(), n_docs, avgdl, e->p.k1, e->p.b,
                                                                          1.0 - e->p.b, e->p.k1 + 1.0, e->kd.as<double>(),
                                                                          e->g1.as<double>());
        LAUNCHED(e);
    }
    e->h_post_ptr.assign((size_t)e->n_vocab + 1, 0);
    CK(cudaMemcpyAsync(e->h_post_ptr.data(), e->post_ptr.p, (size_t)(e->n_vocab + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->avgdl = avgdl;
    e->n_bm25 = n_docs;
    return build_len_tables(e, n_docs, true);
}

int ais_export_postings(ais_engine* e, int64_t* post_ptr, int32_t* post_doc, int32_t* post_tf) {
    if (!e || !post_ptr) return fail(AIS_ERR_INVALID, "NULL argument");
    if (!e->post_ptr.p) return fail(AIS_ERR_NOT_LOADED, "no posting lists");
    DeviceGuard g(e->device);
    CK(cudaMemcpyAsync(post_ptr, e->post_ptr.p, (size_t)(e->n_vocab + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    if (post_doc && e->n_post > 0) CK(cudaMemcpyAsync(post_doc, e->post_doc.p, (size_t)e->n_post * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (post_tf && e->n_post > 0) {
        if (!e->has_tf) return fail(AIS_ERR_INVALID, "the index was loaded without tf (every tf is 1)");
        CK(cudaMemcpyAsync(post_tf, e->post_tf.p, (size_t)e->n_post * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

// ---- seams ----------------------------------------------------------------------------------------
int ais_dot_scores(ais_engine* e, const float* q, float* out) {
    if (!e || !q || !out) return fail(AIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    if (e->n_vec <= 0) return fail(AIS_ERR_NOT_LOADED, "no doc vectors loaded");
    TRY(ensure_work(e));
    memcpy(e->h_q, q, DIM * sizeof(float));
    CK(cudaMemcpyAsync(e->d_q.p, e->h_q, DIM * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    init_keys_kernel<<<1, 64, 0, e->stream>>>(e->maxs_key.as<uint32_t>(), e->maxb_key.as<uint64_t>(), e->maxr_key.as<uint64_t>(),
                                             e->status.as<int32_t>(), 1, 1);
    LAUNCHED(e);
    TRY(launch_scan(e, e->d_q.as<float>(), 1, e->sim.as<float>(), e->maxs_key.as<uint32_t>()));
    CK(cudaMemcpyAsync(out, e->sim.p, (size_t)e->n_vec * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

int ais_bm25_scores(ais_engine* e, const int32_t* term_ids, const double* weights, int32_t n_terms, double* out) {
    if (!e || !out || (n_terms > 0 && (!term_ids || !weights))) return fail(AIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    if (e->n_bm25 < 0) return fail(AIS_ERR_NOT_LOADED, "BM25 index not loaded");
    TRY(ensure_work(e));
    ais_query q{nullptr, term_ids, weights, n_terms};
    TRY(upload_queries(e, &q, 1, false, true));
    init_keys_kernel<<<1, 64, 0, e->stream>>>(e->maxs_key.as<uint32_t>(), e->maxb_key.as<uint64_t>(), e->maxr_key.as<uint64_t>(),
                                             e->status.as<int32_t>(), 1, 1);
    LAUNCHED(e);
    TRY(dev_alloc(e, e->scratch64, (size_t)e->ld * sizeof(double)));
    TRY(launch_bm25_max(e, 1, e->scratch64.as<double>()));
    CK(cudaMemcpyAsync(out, e->scratch64.p, (size_t)e->n_bm25 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

int ais_final_scores(ais_engine* e, const ais_query* q, double* out) {
    if (!e || !q || !out) return fail(AIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    TRY(check_loaded(e));
    TRY(ensure_work(e));
    TRY(ensure_sel(e, 1));
    TRY(do_score(e, q, 1, e->maxes_own.as<double>()));
    e->cur_maxes = e->maxes_own.as<double>();
    TRY(materialize_finals(e, 0, e->maxes_own.as<double>()));
    CK(cudaMemcpyAsync(out, e->scratch64.p, (size_t)e->n() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

// ---- the fused path -----------------------------------------------------------------------------------
int ais_search(ais_engine* e, const ais_query* queries, int32_t n_queries, int32_t topn, int32_t prf_mode, ais_infer_cb cb,
               void* cb_ctx, int64_t* out_ids, double* out_scores, int32_t* out_counts, int32_t* out_status) {
    if (!e || (n_queries > 0 && !queries)) return fail(AIS_ERR_INVALID, "NULL argument");
    if (topn < 1) return fail(AIS_ERR_INVALID, "topn must be >= 1");
    if (prf_mode < AIS_PRF_CALLBACK || prf_mode > AIS_PRF_OFF) return fail(AIS_ERR_INVALID, "unknown prf_mode %d", prf_mode);
    DeviceGuard g(e->device);
    TRY(check_loaded(e));
    TRY(need_whole_index(e));
    for (int q0 = 0; q0 < n_queries; q0 += e->p.max_batch) {
        const int nq = n_queries - q0 < e->p.max_batch ? n_queries - q0 : e->p.max_batch;
        TRY(run_batch(e, queries + q0, q0, nq, topn, prf_mode, cb, cb_ctx, out_ids ? out_ids + (size_t)q0 * topn : nullptr,
                      out_scores ? out_scores + (size_t)q0 * topn : nullptr, out_counts ? out_counts + q0 : nullptr,
                      out_status ? out_status + q0 : nullptr));
    }
    return AIS_OK;
}

int ais_rerank(ais_engine* e, const double* final_scores, int32_t topn, int32_t prf_mode, ais_infer_cb cb, void* cb_ctx,
               int64_t* out_ids, double* out_scores, int32_t* out_count, int32_t* out_status) {
    if (!e || !final_scores) return fail(AIS_ERR_INVALID, "NULL argument");
    if (topn < 1) return fail(AIS_ERR_INVALID, "topn must be >= 1");
    DeviceGuard g(e->device);
    TRY(check_loaded(e));
    TRY(need_whole_index(e));
    TRY(ensure_work(e));
    TRY(dev_alloc(e, e->fin_ext, (size_t)e->ld * sizeof(double)));
    CK(cudaMemcpyAsync(e->fin_ext.p, final_scores, (size_t)e->n() * sizeof(double), cudaMemcpyDefault, e->stream));
    e->ext_fin = true;
    e->rer_column = false;
    e->bound_ok = false;
    e->p1_k = 0;
    CK(cudaMemsetAsync(e->status.p, 0, sizeof(int32_t), e->stream));
    e->cur_nq = 1;
    return run_batch(e, nullptr, 0, 1, topn, prf_mode, cb, cb_ctx, out_ids, out_scores, out_count, out_status);
}

// filter_searched_result(sorted_scores) webui.py:63-80 on a caller-supplied, already sorted list
int ais_filter_sorted(ais_engine* e, const int64_t* ids, const double* scores, int64_t n, int64_t* out_ids, double* out_scores,
                      int64_t* out_count) {
    if (!e || !out_count || (n > 0 && (!ids || !scores || !out_ids || !out_scores))) return fail(AIS_ERR_INVALID, "NULL argument");
    if (n < 0 || n >= (1LL << 31)) return fail(AIS_ERR_INVALID, "list length outside [0, 2^31)");
    *out_count = 0;
    if (n == 0) return AIS_OK;
    DeviceGuard g(e->device);
    Buf d_sc, d_oi, d_os;
    int st = AIS_OK;
    auto body = [&]() -> int {
        TRY(dev_alloc(e, e->fs_keys, (size_t)n * sizeof(uint64_t)));
        TRY(dev_alloc(e, e->fs_ids, (size_t)n * sizeof(int64_t)));
        TRY(dev_alloc(e, e->fs_count, sizeof(int64_t) * 2));
        TRY(dev_alloc(e, d_sc, (size_t)n * sizeof(double)));
        TRY(dev_alloc(e, d_oi, (size_t)n * sizeof(int64_t)));
        TRY(dev_alloc(e, d_os, (size_t)n * sizeof(double)));
        TRY(dev_alloc(e, e->out_count, sizeof(int32_t)));
        TRY(dev_alloc(e, e->out_amb, sizeof(int32_t)));
        CK(cudaMemcpyAsync(d_sc.p, scores, (size_t)n * sizeof(double), cudaMemcpyDefault, e->stream));
        CK(cudaMemcpyAsync(e->fs_ids.p, ids, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, e->stream));
        CK(cudaMemcpyAsync(e->fs_count.p, &n, sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
        keys_from_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d_sc.as<double>(), n, e->fs_keys.as<uint64_t>());
        LAUNCHED(e);
        TailParams tp;
        tp.thresh = e->p.diff_filter_thresh;
        tp.topn = (int)n;
        tp.depth = 0;
        tp.normalize = 0;
        tp.n_total = n;
        tail_kernel<<<1, SEL_THREADS, 0, e->stream>>>(e->fs_keys.as<uint64_t>(), e->fs_ids.as<int64_t>(), 0, nullptr,
                                                     e->fs_count.as<int64_t>(), nullptr, nullptr, tp, nullptr, d_oi.as<int64_t>(),
                                                     d_os.as<double>(), e->out_count.as<int32_t>(), e->out_amb.as<int32_t>(), nullptr);
        LAUNCHED(e);
        int32_t cnt = 0;
        CK(cudaMemcpyAsync(&cnt, e->out_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (cnt > 0) {
            CK(cudaMemcpyAsync(out_ids, d_oi.p, (size_t)cnt * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(out_scores, d_os.p, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
            CK(cudaStreamSynchronize(e->stream));
        }
        *out_count = cnt;
        return AIS_OK;
    };
    st = body();
    dev_free(e, d_sc);
    dev_free(e, d_oi);
    dev_free(e, d_os);
    return st;
}

// ---- staged form ----------------------------------------------------------------------------------------
int ais_stage_score(ais_engine* e, const ais_query* queries, int32_t nq, double* d_maxes) {
    if (!e || !queries || !d_maxes) return fail(AIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    return do_score(e, queries, nq, d_maxes);
}
int ais_stage_combine(ais_engine* e, int32_t nq, const double* d_maxes, int32_t k, uint64_t* d_cand_keys, int64_t* d_cand_ids) {
    if (!e || !d_maxes || !d_cand_keys || !d_cand_ids) return fail(AIS_ERR_INVALID, "NULL argument");
    if (nq != e->cur_nq) return fail(AIS_ERR_INVALID, "nq %d differs from the scored batch (%d)", nq, e->cur_nq);
    DeviceGuard g(e->device);
    return do_combine(e, nq, d_maxes, k, d_cand_keys, d_cand_ids);
}
int ais_stage_top(ais_engine* e, int32_t nq, int32_t n_lists, int32_t k, const uint64_t* d_cand_keys, const int64_t* d_cand_ids,
                  int64_t* out_top_ids, double* out_top_scores, float* d_rows) {
    if (!e || !d_cand_keys || !d_cand_ids || n_lists < 1) return fail(AIS_ERR_INVALID, "bad argument");
    if (nq != e->cur_nq) return fail(AIS_ERR_INVALID, "nq %d differs from the scored batch (%d)", nq, e->cur_nq);
    DeviceGuard g(e->device);
    return do_top(e, nq, n_lists, k, d_cand_keys, d_cand_ids, out_top_ids, out_top_scores, d_rows);
}
int ais_stage_set_status(ais_engine* e, int32_t nq, const int32_t* status) {
    if (!e || !status || nq != e->cur_nq) return fail(AIS_ERR_INVALID, "bad argument");
    DeviceGuard g(e->device);
    CK(cudaMemcpyAsync(e->status.p, status, (size_t)nq * sizeof(int32_t), cudaMemcpyDefault, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}
int ais_stage_requery(ais_engine* e, int32_t nq, const float* q2, const float* d_rows, int32_t prf_mode, int32_t k,
                      double* d_max_r, uint64_t* d_cand_keys, int64_t* d_cand_ids) {
    if (!e || !d_max_r || !d_cand_keys || !d_cand_ids) return fail(AIS_ERR_INVALID, "NULL argument");
    if (nq != e->cur_nq) return fail(AIS_ERR_INVALID, "nq %d differs from the scored batch (%d)", nq, e->cur_nq);
    DeviceGuard g(e->device);
    return do_requery(e, nq, q2, d_rows, prf_mode, k, d_max_r, d_cand_keys, d_cand_ids);
}
int ais_stage_requery_select(ais_engine* e, int32_t nq, int32_t k, uint64_t* d_cand_keys, int64_t* d_cand_ids) {
    if (!e || !d_cand_keys || !d_cand_ids) return fail(AIS_ERR_INVALID, "NULL argument");
    if (nq != e->cur_nq || !e->cur_prf) return fail(AIS_ERR_INVALID, "no second pass to re-select from");
    DeviceGuard g(e->device);
    return do_requery_select(e, nq, k, d_cand_keys, d_cand_ids);
}
int ais_stage_finish(ais_engine* e, int32_t nq, int32_t n_lists, int32_t k, const uint64_t* d_cand_keys, const int64_t* d_cand_ids,
                     const double* d_max_r, const int32_t* d_witness, int32_t topn, int64_t* out_ids, double* out_scores,
                     int32_t* out_counts, int32_t* out_status, int32_t* out_ambiguous, uint64_t* out_last_keys) {
    if (!e || !d_cand_keys || !d_cand_ids || n_lists < 1) return fail(AIS_ERR_INVALID, "bad argument");
    if (nq != e->cur_nq) return fail(AIS_ERR_INVALID, "nq %d differs from the scored batch (%d)", nq, e->cur_nq);
    DeviceGuard g(e->device);
    return do_finish(e, nq, n_lists, k, d_cand_keys, d_cand_ids, d_max_r, d_witness, topn, out_ids, out_scores, out_counts,
                     out_status, out_ambiguous, out_last_keys);
}
int ais_stage_witness(ais_engine* e, int32_t nq, const int32_t* ambiguous, const uint64_t* last_keys, int32_t second_pass,
                      const double* d_max_r, int32_t* d_witness) {
    if (!e || !ambiguous || !last_keys || !d_witness) return fail(AIS_ERR_INVALID, "NULL argument");
    if (nq != e->cur_nq) return fail(AIS_ERR_INVALID, "nq %d differs from the scored batch (%d)", nq, e->cur_nq);
    DeviceGuard g(e->device);
    return do_witness(e, nq, ambiguous, last_keys, second_pass, d_max_r, d_witness);
}
int ais_stage_export_keys(ais_engine* e, int32_t query, int32_t second_pass, uint64_t* d_keys, int64_t* d_ids) {
    if (!e || !d_keys || !d_ids || query < 0 || query >= e->cur_nq) return fail(AIS_ERR_INVALID, "bad argument");
    DeviceGuard g(e->device);
    return do_export_keys(e, query, second_pass, d_keys, d_ids);
}
int ais_stage_sort_finish(ais_engine* e, int32_t query, uint64_t* d_keys, int64_t* d_ids, int64_t n_entries, const double* d_max_r,
                          int32_t topn, int64_t* out_ids, double* out_scores, int32_t* out_count, int32_t* out_status) {
    if (!e || !d_keys || !d_ids || query < 0 || query >= e->cur_nq || n_entries < 0 || topn < 1)
        return fail(AIS_ERR_INVALID, "bad argument");
    DeviceGuard g(e->device);
    return do_sort_finish(e, query, d_keys, d_ids, n_entries, d_max_r, topn, out_ids, out_scores, out_count, out_status);
}

// test seam: read back a per-doc work array of the current batch (which: 0 sim fp32, 1 bm25 fp64, 2 combined fp64, 3 rer fp32)
int ais_debug_read(ais_engine* e, int32_t which, int32_t query, void* out) {
    if (!e || !out || query < 0 || query >= e->qt_cap || which < 0 || which > 3) return fail(AIS_ERR_INVALID, "bad argument");
    DeviceGuard g(e->device);
    const size_t n = (size_t)e->n();
    if (which == 3) TRY(dev_alloc(e, e->rer, (size_t)e->qt_cap * e->ld * sizeof(float)));
    if (which == 3 && e->rer_column && e->n_vec > 0)        // column mode keeps no rer array: materialise this query's row
        TRY(launch_scan_column(e, e->d_q2.as<float>() + (size_t)query * DIM, 1, e->col_comp, e->rer.as<float>() + (size_t)query * e->ld,
                               e->maxs_key.as<uint32_t>() + query));
    if (which == 2) {
        if (!e->cur_maxes) return fail(AIS_ERR_INVALID, "no combined scores: no batch has been combined yet");
        TRY(materialize_finals(e, query, e->cur_maxes));
    }
    const void* src = which == 0 ? (const void*)(e->sim.as<float>() + (size_t)query * e->ld)
                    : (which == 1 || which == 2) ? (const void*)(e->scratch64.as<double>())
                                 : (const void*)(e->rer.as<float>() + (size_t)query * e->ld);
    CK(cudaMemcpyAsync(out, src, n * ((which == 0 || which == 3) ? 4 : 8), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

#ifdef AIS_TC_TRACE
extern "C" int ais_debug_tc_trace(ais_engine* e, long long* out /*[4][256][4]*/) {
    DeviceGuard g(e->device);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 4 * 256 * 4));
    return AIS_OK;
}
#endif

// ---- introspection ----------------------------------------------------------------------------------------
int ais_set_profiling(ais_engine* e, int on) {
    if (!e) return fail(AIS_ERR_INVALID, "NULL engine");
    e->profiling = on != 0;
    return AIS_OK;
}
static int drain_events(ais_engine* e) {
    if (e->ev_pending.empty()) return AIS_OK;
    CK(cudaStreamSynchronize(e->stream));
    for (const ais_engine::PendingEv& p : e->ev_pending) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, p.a, p.b));
        e->kind_ms[p.kind] += ms;
        e->ev_free.push_back(p.a);
        e->ev_free.push_back(p.b);
    }
    e->ev_pending.clear();
    return AIS_OK;
}
int ais_get_stats(ais_engine* e, ais_stats* out) {
    if (!e || !out) return fail(AIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    TRY(drain_events(e));
    out->n_docs = e->n_vec;
    out->n_postings = e->n_post;
    out->dim = DIM;
    out->n_terms = e->n_vocab;
    out->scan_launches = e->scan_launches;
    out->scan_ms_total = e->kind_ms[AIS_KIND_SCAN];
    for (int k = 0; k < AIS_N_KINDS; ++k) { out->kind_ms[k] = e->kind_ms[k]; out->kind_launches[k] = e->kind_launches[k]; }
    out->kernel_launches = e->kernel_launches;
    out->fullsort_fallbacks = e->fullsort_fallbacks;
    out->bytes_device = e->bytes_device;
    out->column_scan_launches = e->column_scan_launches;
    out->tiles_per_seg = e->last_tiles_per_seg;
    out->pair_scan_launches = e->pair_scan_launches;
    out->bound_passes = e->bound_passes + e->skip_passes;
    out->bitmap_batches = e->bitmap_batches;
    return AIS_OK;
}
int ais_reset_stats(ais_engine* e) {
    if (!e) return fail(AIS_ERR_INVALID, "NULL engine");
    DeviceGuard g(e->device);
    TRY(drain_events(e));
    e->scan_launches = e->kernel_launches = e->fullsort_fallbacks = e->column_scan_launches = e->bound_passes = e->bitmap_batches = e->skip_passes = e->pair_scan_launches = 0;
    for (int k = 0; k < AIS_N_KINDS; ++k) { e->kind_ms[k] = 0.0; e->kind_launches[k] = 0; }
    return AIS_OK;
}
int ais_synchronize(ais_engine* e) {
    if (!e) return fail(AIS_ERR_INVALID, "NULL engine");
    DeviceGuard g(e->device);
    CK(cudaStreamSynchronize(e->stream));
    return AIS_OK;
}

}  // extern "C"
