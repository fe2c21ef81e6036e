/*
 * ais_b200.h - C ABI of the B200-native query-time scoring engine for
 * ryogrid/anime-illust-image-searcher (the webui.py search path).
 *
 * The reference has no FFI: its seam is the set of Python callables webui.py uses
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it
 * replaces (paths under /root/reference).  The binding a maintainer would add to
 * webui.py is shown in INTEGRATION.md (ctypes, ~30 lines).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - every function returns an int status (AIS_OK == 0); ais_last_error() returns a
 *     thread-local message for the last non-zero status.
 *   - the caller owns every buffer it passes; the engine owns all device memory it
 *     allocates.  "host-or-device" pointers may be either (CUDA unified addressing).
 *   - one engine == one GPU == one shard of the documents (one process per GPU;
 *     multi-GPU search exchanges the small per-query records through the caller's
 *     collective library between the ais_stage_* calls).
 *   - calls on one engine must not overlap in time (webui.py is single-threaded per
 *     session; the Python shim holds a lock).
 *   - doc ids are 0-based rows of tags-wd-tagger_doc2vec_idx.csv (webui.py:592).
 */
#ifndef AIS_B200_H
#define AIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIS_ABI_VERSION 1

/* status codes */
#define AIS_OK 0
#define AIS_ERR_INVALID 1       /* bad argument */
#define AIS_ERR_CUDA 2          /* CUDA runtime failure (message has the CUDA error string) */
#define AIS_ERR_NOT_LOADED 3    /* index not staged yet */
#define AIS_ERR_UNSUPPORTED 4   /* e.g. vector dimension other than 300 (genmodel.py:16) */
#define AIS_ERR_CALLBACK 5      /* the PRF re-inference callback failed */

/* per-query status written by ais_search / ais_stage_finish */
#define AIS_Q_OK 0
#define AIS_Q_NAN_WEIGHTS 1     /* PRF weights hold -inf/NaN: reference raises ValueError
                                   "cannot convert float NaN to integer" (webui.py:202-203) */
#define AIS_Q_ZERO_WEIGHT_SUM 2 /* reference: ZeroDivisionError from np.average (webui.py:200) */
#define AIS_Q_ZERO_VECTOR 3     /* reference: gensim unitvec assertion inside index[.] (webui.py:205) */
#define AIS_Q_CALLBACK_FAILED 4 /* callback returned non-zero; see ais_search */

/* Engine constants.  Defaults are the reference's module constants, all "modifiable"
 * per webui.py:51-52; they are honoured at call time (ais_set_params). */
typedef struct ais_params {
    double k1;                      /* 1.5   webui.py:126 */
    double b;                       /* 0.75  webui.py:127 */
    double bm25_weight;             /* 0.5   BM25_WEIGHT            webui.py:51 */
    double doc2vec_weight;          /* 0.5   DOC2VEC_WEIGHT         webui.py:52 */
    double original_score_weight;   /* 0.7   ORIGINAL_SCORE_WEIGHT  webui.py:55 */
    double reranked_score_weight;   /* 0.3   RERANKED_SCORE_WEIGHT  webui.py:56 */
    double diff_filter_thresh;      /* 1e-6  DIFF_FILTER_THRESH     webui.py:58 */
    double require_magic;           /* 1000  REQUIRE_TAG_MAGIC_NUMBER webui.py:60 */
    int32_t prf_depth;              /* 10    webui.py:193-195 */
    int32_t max_batch;              /* queries per engine batch (1..256); up to 64 of them share one pass over the doc vectors
                                     * (ais_search takes any n_queries and cuts the list into engine batches itself) */
} ais_params;

/* One weighted tag query, already parsed by the host (webui.py:354-371 stays Python). */
typedef struct ais_query {
    const float* vec;        /* [dim] host; dense fp32 query as gensim hands it to numpy.dot:
                                sparse2full(unitvec(normalize_and_apply_weight_doc2vec(q))) webui.py:349-352 */
    const int32_t* term_ids; /* [n_terms] host; dict keys in insertion order (webui.py:355,374) */
    const double* weights;   /* [n_terms] host; 1000+W = required, negative = exclude */
    int32_t n_terms;         /* <= AIS_MAX_TERMS */
} ais_query;

#define AIS_MAX_TERMS 64

/* How the pseudo-relevance-feedback re-query vector (webui.py:193-205) is obtained. */
typedef enum ais_prf_mode {
    AIS_PRF_CALLBACK = 0,        /* host callback re-infers the top docs (reference behaviour,
                                    webui.py:182-187 calls gensim infer_vector) */
    AIS_PRF_STORED_ROWS = 1,     /* device: weighted mean of the STORED rows of the top docs, then the
                                    reference's (collapsing) normalisation webui.py:200-205 */
    AIS_PRF_STORED_ROWS_FULL = 2,/* device: as above but the un-collapsed ("intended") centroid */
    AIS_PRF_OFF = 3              /* skip the re-rank pass: sorted(final) -> filter -> [:topn] */
} ais_prf_mode;

/* PRF callback: given the top-`depth` docs of query `query_index` (global doc ids, their
 * combined scores = the np.average weights), write the dense fp32 re-query vector [dim] exactly
 * as gensim would build it from weighted_mean_vec_with_docid (webui.py:200-205).
 * Return 0 on success; any other value marks the query AIS_Q_CALLBACK_FAILED. */
typedef int (*ais_infer_cb)(void* ctx, int32_t query_index, const int64_t* doc_ids,
                            const double* scores, int32_t depth, float* out_query);

typedef struct ais_engine ais_engine;

/* kernel classes of the per-class CUDA-event timers (ais_set_profiling) */
#define AIS_N_KINDS 8
#define AIS_KIND_SCAN 0         /* dense doc-vector scans (index[vec], webui.py:352,205) */
#define AIS_KIND_BM25_SLICES 1  /* posting-list slice table of the batch */
#define AIS_KIND_BM25_SCORE 2   /* compute_bm25_scores (webui.py:119-172) -> per-tile records */
#define AIS_KIND_COMBINE 3      /* normalise + combine (webui.py:376-383) -> tile / segment maxima */
#define AIS_KIND_SELECT 4       /* threshold, collect, survivor sort (the sorts of webui.py:191-192,237) */
#define AIS_KIND_REQUERY 5      /* PRF seeds, centroid, re-query scores (webui.py:193-205) */
#define AIS_KIND_TAIL 6         /* merge, filter_searched_result, result copy (webui.py:63-80,219-246) */
#define AIS_KIND_WITNESS 7      /* near-tie witness pass / exact full-sort fallback */

typedef struct ais_stats {
    int64_t n_docs;             /* docs in this shard */
    int64_t n_postings;
    int32_t dim;
    int32_t n_terms;
    int64_t scan_launches;      /* doc-vector scan kernel launches since the last reset */
    double scan_ms_total;       /* their summed CUDA-event duration (profiling on) */
    int64_t kernel_launches;    /* all kernels launched by the engine since the last reset */
    int64_t fullsort_fallbacks; /* times the ambiguous-filter fallback sorted the whole shard */
    int64_t bytes_device;       /* device memory held by the engine */
    int64_t column_scan_launches; /* re-query passes served by the single-component column scan */
    int64_t tiles_per_seg;      /* 256-doc tiles per select segment in the last batch (> 1 from ~525 k docs per shard) */
    double kind_ms[AIS_N_KINDS];        /* summed CUDA-event time per kernel class since the last reset (profiling on) */
    int64_t kind_launches[AIS_N_KINDS]; /* timed brackets per class */
    int64_t bound_passes;       /* second passes served by the per-tile bound on the blend instead of a streaming pass */
    int64_t bitmap_batches;     /* batches whose BM25 side ran on the term-bitmap path (tf == 1 index) instead of per-tile records */
    int64_t pair_scan_launches; /* of scan_launches: 64-query passes on CTA pairs (scan_pair_kernel, tcgen05 cta_group::2) */
} ais_stats;

const char* ais_last_error(void);
int ais_abi_version(void);
void ais_default_params(ais_params* p);

/* --- lifecycle --------------------------------------------------------------------- */
int ais_create(ais_engine** out, int device_id, const ais_params* p /* NULL = defaults */);
int ais_destroy(ais_engine* e);
int ais_set_params(ais_engine* e, const ais_params* p);
/* Run on a caller-owned CUDA stream (cudaStream_t as void*); NULL = the engine's own (non-blocking)
 * stream.  To run on the legacy default stream pass its explicit handle cudaStreamLegacy (0x1). */
int ais_set_stream(ais_engine* e, void* cuda_stream);
/* This engine holds docs [first_doc_id, first_doc_id + n_local) of an index of n_total docs. */
int ais_set_shard(ais_engine* e, int64_t first_doc_id, int64_t n_total_docs);

/* --- index staging: replaces load_model() webui.py:649-689 --------------------------- */
/* Rows [first_row, first_row+n) of the fp32 doc-vector matrix (gensim Similarity shards,
 * genmodel.py:168-175; rows are stored RAW, never normalised).  host-or-device pointer.
 * Appendable shard by shard. */
int ais_load_vectors(ais_engine* e, const float* rows, int64_t n, int32_t dim, int64_t first_row);
/* Pre-size the row store (avoids re-allocation while gensim shards are appended one by one). */
int ais_reserve_docs(ais_engine* e, int64_t n_docs);
/* Declare n_docs rows resident and hand out the device pointer of the row store [n_docs x 300] so a
 * caller that already has the rows on the device (or generates them there) can write them in place.
 * The rows must be final before the next search: the engine caches derived data (column 0 for the
 * collapsed PRF re-query); call this again (or ais_load_vectors) after changing rows. */
int ais_vectors_device_ptr(ais_engine* e, int64_t n_docs, float** out_rows);
/* The BM25 index of genmodel.py:51-99 (bm25_corpus / bm25_idf / bm25_doc_lengths / bm25_avgdl) as
 * tag-major posting lists with LOCAL doc ids ascending inside each list.  post_tf may be NULL
 * (every tf == 1).  idf[t] = 0 for terms absent from bm25_idf (webui.py:140).  host-or-device. */
int ais_load_bm25(ais_engine* e, const int64_t* post_ptr /*[n_terms+1]*/, const int32_t* post_doc,
                  const int32_t* post_tf, int32_t n_terms, int64_t n_docs, const double* idf /*[n_terms]*/,
                  const int64_t* doc_len /*[n_docs]*/, double avgdl);

/* --- native reader of the reference's `bm25_corpus` file (host code, no GPU needed) --------------------------------
 * genmodel.py:84-85 pickles a Python list of N dicts {term id: tf}; load_model() (webui.py:680) unpickles it into N
 * Python dicts on every cold start.  These two calls stream the pickle's opcodes straight into doc-major CSR arrays
 * instead: scan counts docs / entries, fill writes row_ptr [n_docs+1], term_ids [nnz], tfs [nnz] (host arrays; dict
 * insertion order kept).  A pickle outside the list-of-int-dicts opcode subset returns AIS_ERR_UNSUPPORTED (the caller
 * falls back to pickle.load); the message is in ais_pickle_last_error(). */
int ais_pickle_csr_scan(const char* path, int64_t* out_n_docs, int64_t* out_nnz);
int ais_pickle_csr_fill(const char* path, int64_t n_docs, int64_t nnz, int64_t* row_ptr, int32_t* term_ids, int32_t* tfs);
const char* ais_pickle_last_error(void);

/* --- index build: gen_and_save_bm25_index genmodel.py:51-99 on the GPU ------------------------ */
/* From the docs' term-id sequences (doc-major CSR, csv tag order, repeats allowed; the host has already
 * mapped tags through dictionary.token2id and dropped unknown ones, genmodel.py:59-61) build, on the device:
 * doc lengths (genmodel.py:69), document frequencies (genmodel.py:72-73) and the tag-major posting lists with
 * tf (the bm25_corpus dicts of genmodel.py:64-68, transposed).  out_df [n_terms] / out_doc_len [n_docs] are
 * optional host-or-device outputs.  IDF (genmodel.py:79-82 uses numpy.log) and avgdl (numpy.mean) stay on the
 * host for bit-exact parity: pass them to ais_finish_bm25, which makes the index searchable. */
int ais_build_bm25(ais_engine* e, const int64_t* seq_ptr /*[n_docs+1]*/, const int32_t* seq_ids, int64_t n_docs,
                   int32_t n_terms, int64_t* out_df, int64_t* out_doc_len);
int ais_finish_bm25(ais_engine* e, const double* idf /*[n_terms], 0 where df == 0*/, double avgdl);
/* posting lists as staged / built (host arrays; post_ptr [n_terms+1], post_doc / post_tf [post_ptr[n_terms]]) */
int ais_export_postings(ais_engine* e, int64_t* post_ptr, int32_t* post_doc, int32_t* post_tf);

/* --- test seams (full score vectors; they defeat fusion and are not the fast path) ---- */
/* index[vec]  (webui.py:352, :205): out[N] fp32 = rows . q ; q is the dense unit query. */
int ais_dot_scores(ais_engine* e, const float* q /*[dim] host*/, float* out /*[N] host*/);
/* compute_bm25_scores(query_weights=...) webui.py:119-172: out[N] fp64, -inf where masked. */
int ais_bm25_scores(ais_engine* e, const int32_t* term_ids, const double* weights, int32_t n_terms,
                    double* out /*[N] host*/);
/* 0.5*bm25/max + 0.5*sim/max of webui.py:376-383 for one query: out[N] fp64 (host). */
int ais_final_scores(ais_engine* e, const ais_query* q, double* out);

/* --- the fused path: find_similar_documents(new_doc, topn) webui.py:345-390 ------------ */
/* Scores every doc of the shard for each query (BM25 + dot + masks + max-normalised combine),
 * runs the PRF re-rank, applies filter_searched_result (webui.py:63-80) and returns up to topn
 * (doc id, score) pairs per query, best first.  out_ids/out_scores are [n_queries x topn] host
 * arrays, out_counts / out_status [n_queries]. */
int ais_search(ais_engine* e, const ais_query* queries, int32_t n_queries, int32_t topn, int32_t prf_mode,
               ais_infer_cb cb, void* cb_ctx, int64_t* out_ids, double* out_scores, int32_t* out_counts,
               int32_t* out_status);
/* get_doc2vec_based_reranked_scores(final_scores, topn) webui.py:189-253 on caller-supplied
 * combined scores (host fp64[N]). */
int ais_rerank(ais_engine* e, const double* final_scores, int32_t topn, int32_t prf_mode, ais_infer_cb cb,
               void* cb_ctx, int64_t* out_ids, double* out_scores, int32_t* out_count, int32_t* out_status);

/* filter_searched_result(sorted_scores) webui.py:63-80 on a caller-supplied list of n (doc id, score)
 * pairs sorted best first (host-or-device arrays).  Writes *out_count <= n pairs (host arrays [n]). */
int ais_filter_sorted(ais_engine* e, const int64_t* ids, const double* scores, int64_t n, int64_t* out_ids,
                      double* out_scores, int64_t* out_count);

/* --- staged form of ais_search for doc-sharded multi-GPU search ------------------------- */
/* All d_* arguments are DEVICE pointers owned by the caller (e.g. torch tensors) so that the
 * caller can run its collectives (NCCL all-reduce MAX / all-gather) on them between stages, on
 * the stream given to ais_set_stream.  nq <= max_batch.  Record layout:
 *   maxes   double [nq][2]      {max bm25, max dot} of this shard           (all-reduce MAX)
 *   cand    {uint64 key, int64 global doc id} as two arrays [nq][k]        (all-gather)
 *   key     order-preserving image of the fp64 score (larger key = better); 0 = empty slot
 *   rows    float  [nq][depth][dim]  stored rows of the top docs, zeros if not on this shard
 *                                                                          (all-reduce SUM)   */
int ais_stage_score(ais_engine* e, const ais_query* queries, int32_t nq, double* d_maxes);
int ais_stage_combine(ais_engine* e, int32_t nq, const double* d_maxes, int32_t k, uint64_t* d_cand_keys,
                      int64_t* d_cand_ids);
/* merge `n_lists` candidate lists per query ([n_lists][nq][k], as all-gathered) into the global
 * top-`depth` kept inside the engine; optionally export them (host arrays, may be NULL) and
 * gather their stored rows that live on this shard into d_rows (may be NULL). */
int ais_stage_top(ais_engine* e, int32_t nq, int32_t n_lists, int32_t k, const uint64_t* d_cand_keys,
                  const int64_t* d_cand_ids, int64_t* out_top_ids, double* out_top_scores, float* d_rows);
/* re-query vectors: either host vectors [nq][dim] (callback / host path) or, when q2 == NULL,
 * built on device from d_rows (the all-reduced stored rows) per prf_mode.  Then the second scan,
 * the 0.7/0.3 blend, this shard's max and its top-k candidates (top docs excluded). */
int ais_stage_requery(ais_engine* e, int32_t nq, const float* q2, const float* d_rows, int32_t prf_mode,
                      int32_t k, double* d_max_r /*[nq]*/, uint64_t* d_cand_keys, int64_t* d_cand_ids);
/* host-side PRF (callback mode in a sharded caller): per-query status decided on the host
 * (AIS_Q_NAN_WEIGHTS ...) before ais_stage_requery; such queries return count 0. */
int ais_stage_set_status(ais_engine* e, int32_t nq, const int32_t* status);
/* merge the all-gathered second-pass candidates, normalise by the (all-reduced) max, apply
 * filter_searched_result and write the results (host arrays).  d_max_r == NULL selects the
 * no-PRF branch (webui.py:247-253): the candidates are the sorted combined scores of
 * ais_stage_combine and no docs are pinned.  A query whose filter outcome depends on scores
 * beyond the k candidates gets out_ambiguous[q] = 1: the caller runs ais_stage_witness, repeats this
 * call with the witness flags, and sorts everything (ais_stage_export_keys + ais_stage_sort_finish)
 * for the queries that are still ambiguous. */
int ais_stage_finish(ais_engine* e, int32_t nq, int32_t n_lists, int32_t k, const uint64_t* d_cand_keys,
                     const int64_t* d_cand_ids, const double* d_max_r, const int32_t* d_witness /* nullable */,
                     int32_t topn, int64_t* out_ids, double* out_scores, int32_t* out_counts, int32_t* out_status,
                     int32_t* out_ambiguous, uint64_t* out_last_keys /* nullable, [nq] */);
/* Cheap resolution of an ambiguous outcome: for every query with ambiguous[q] != 0 (host array) look, on
 * this shard, for two DISTINCT scores closer than DIFF_FILTER_THRESH at or below the last entry of the
 * sorted prefix (last_keys[q], host array as returned by ais_stage_finish) - such a pair implies an adjacent
 * near-tie below the prefix.  d_witness [nq] (device, int32) is set to 1 where one is found; the caller
 * all-reduces it (MAX) and repeats ais_stage_finish with it.  Sufficient, not necessary: a query that
 * stays ambiguous goes to the full sort below. */
int ais_stage_witness(ais_engine* e, int32_t nq, const int32_t* ambiguous, const uint64_t* last_keys,
                      int32_t second_pass, const double* d_max_r, int32_t* d_witness);
/* re-select second-pass candidates with another k without re-scanning (k <= ais_max_select_k()). */
int ais_stage_requery_select(ais_engine* e, int32_t nq, int32_t k, uint64_t* d_cand_keys, int64_t* d_cand_ids);
int ais_max_select_k(void);
/* Exact fallback for an ambiguous filter outcome: every shard exports the keys of ALL its docs for
 * one query (second_pass = 1: the blended R with the pinned top docs blanked; 0: the combined scores)
 * into caller arrays [n_local]; the caller concatenates the shards' exports (all-gather) into device
 * arrays of ais_sort_capacity(n_entries) slots and one engine sorts them and applies the exact
 * filter_searched_result. */
int ais_stage_export_keys(ais_engine* e, int32_t query, int32_t second_pass, uint64_t* d_keys, int64_t* d_ids);
int64_t ais_sort_capacity(int64_t n_entries);
int ais_stage_sort_finish(ais_engine* e, int32_t query, uint64_t* d_keys, int64_t* d_ids, int64_t n_entries,
                          const double* d_max_r /* NULL: no-PRF branch */, int32_t topn, int64_t* out_ids,
                          double* out_scores, int32_t* out_count, int32_t* out_status);

/* test seam: per-doc work array of the current batch (which: 0 dot fp32, 1 bm25 fp64, 2 combined fp64,
 * 3 re-query dot fp32) for one query of the batch -> host array [n_local]. */
int ais_debug_read(ais_engine* e, int32_t which, int32_t query, void* out);

/* --- introspection ----------------------------------------------------------------------- */
int ais_set_profiling(ais_engine* e, int on);   /* CUDA-event timing of every scan launch */
int ais_get_stats(ais_engine* e, ais_stats* out);
int ais_reset_stats(ais_engine* e);
int ais_synchronize(ais_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* AIS_B200_H */
