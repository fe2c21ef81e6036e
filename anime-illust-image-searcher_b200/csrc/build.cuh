// BM25 index build on the GPU: what genmodel.py:51-99 (gen_and_save_bm25_index) produces - per-doc term
// frequencies, doc lengths, document frequencies - in the layout the scoring kernels read: tag-major posting
// lists with ascending doc ids and the tf of every posting.
//
// Stable counting sort by term, without any inter-CTA waiting:
//   count   CTA b owns a contiguous range of docs.  One warp per doc removes duplicate tags (tf = multiplicity)
//           and bumps count[b][term] once per distinct term (atomics on the CTA's own row of the table).
//   scan    one thread per term: exclusive prefix over the CTAs -> start of CTA b's range inside term t's list
//           (+ the document frequency and, after a scan over terms on the host side of the ABI, post_ptr).
//   fill    CTA b walks its docs in order, chunk by chunk: the chunk's (term, doc) pairs are sorted in shared
//           memory (bitonic), every run of equal terms is appended at the CTA's cursor for that term.  Chunks are
//           processed sequentially by the same CTA, so every posting list comes out in ascending doc order.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int BUILD_THREADS = 256;
constexpr int BUILD_PAIRS = 2048;            // (term, doc) pairs sorted per chunk
constexpr int BUILD_MAX_DOC_TAGS = BUILD_PAIRS;   // longest tag list of one doc: a doc never spans sort chunks (the tagger emits ~30;
                                                  // round 1 stopped at 256 - the per-warp scratch lists are gone)

// Distinct terms of one doc with their multiplicities, by one warp, 32 tokens at a time.  emit(term, tf, first, mask) is
// called by ALL lanes for every group of 32 tokens: `first` = this lane's token is the first occurrence of its term in
// the doc (genmodel.py:64-66 counts into a dict), `mask` = ballot of `first`.  Returns the number of distinct terms.
template <class F>
__device__ inline int warp_distinct(const int32_t* __restrict__ ids, int len, F emit) {
    const int lane = threadIdx.x & 31;
    int n_out = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const int32_t t = i < len ? ids[i] : -1;
        bool first = i < len;
        int tf = 0;
        if (i < len) {
            for (int j = 0; j < len; ++j) {              // O(len^2 / 32) per doc; len ~ 30
                const int32_t u = ids[j];
                if (u == t) { if (j < i) first = false; ++tf; }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, first);
        emit(t, tf, first, m);
        n_out += __popc(m);
    }
    return n_out;
}

// count[b][t] += 1 per distinct term of every doc of CTA b; doc_len[d] = number of (known) tags incl. repeats
__global__ void __launch_bounds__(BUILD_THREADS)
build_count_kernel(const int64_t* __restrict__ seq_ptr, const int32_t* __restrict__ seq_ids, int64_t n_docs, int64_t docs_per_cta,
                   int32_t n_terms, int32_t* __restrict__ count, int64_t* __restrict__ doc_len, int* __restrict__ error) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t lo = (int64_t)blockIdx.x * docs_per_cta;
    const int64_t hi = lo + docs_per_cta < n_docs ? lo + docs_per_cta : n_docs;
    int32_t* row = count + (size_t)blockIdx.x * n_terms;
    for (int64_t d = lo + warp; d < hi; d += BUILD_THREADS / 32) {
        const int64_t a = seq_ptr[d];
        const int64_t len64 = seq_ptr[d + 1] - a;
        if (lane == 0) doc_len[d] = len64;                             // genmodel.py:69
        if (len64 < 0 || len64 > BUILD_MAX_DOC_TAGS) { if (lane == 0) atomicExch(error, 1); continue; }
        warp_distinct(seq_ids + a, (int)len64, [&](int32_t t, int, bool first, unsigned) {
            if (first) {
                if (t < 0 || t >= n_terms) atomicExch(error, 2);
                else atomicAdd(&row[t], 1);                              // genmodel.py:72-73
            }
        });
    }
}

// per term: df = sum over CTAs; count[b][t] becomes the exclusive prefix over b
__global__ void build_scan_kernel(int32_t* __restrict__ count, int n_ctas, int32_t n_terms, int64_t* __restrict__ df) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    int64_t run = 0;
    for (int b = 0; b < n_ctas; ++b) {
        const int32_t c = count[(size_t)b * n_terms + t];
        count[(size_t)b * n_terms + t] = (int32_t)run;
        run += c;
    }
    df[t] = run;
}

__global__ void __launch_bounds__(BUILD_THREADS)
build_fill_kernel(const int64_t* __restrict__ seq_ptr, const int32_t* __restrict__ seq_ids, int64_t n_docs, int64_t docs_per_cta,
                  int32_t n_terms, int32_t* __restrict__ cursor /* count table after the scan */,
                  const int64_t* __restrict__ post_ptr, int32_t* __restrict__ post_doc, int32_t* __restrict__ post_tf) {
    __shared__ uint64_t key[BUILD_PAIRS];          // term << 32 | doc (global id < 2^31)
    __shared__ int32_t val[BUILD_PAIRS];           // tf
    __shared__ int s_n;
    __shared__ int64_t s_next;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t lo = (int64_t)blockIdx.x * docs_per_cta;
    const int64_t hi = lo + docs_per_cta < n_docs ? lo + docs_per_cta : n_docs;
    int32_t* row = cursor + (size_t)blockIdx.x * n_terms;
    int64_t d0 = lo;
    while (d0 < hi) {
        // chunk = longest run of docs whose token count fits the sort buffer (a doc never spans chunks)
        if (tid == 0) {
            int64_t d = d0;
            int64_t tokens = 0;
            while (d < hi) {
                const int64_t len = seq_ptr[d + 1] - seq_ptr[d];
                if (tokens + len > BUILD_PAIRS && d > d0) break;
                tokens += len;
                ++d;
                if (tokens >= BUILD_PAIRS) break;
            }
            s_next = d;
            s_n = 0;
        }
        __syncthreads();
        const int64_t d1 = s_next;
        for (int64_t d = d0 + warp; d < d1; d += BUILD_THREADS / 32) {
            const int64_t a = seq_ptr[d];
            const int len = (int)(seq_ptr[d + 1] - a);
            // docs longer than a sort chunk were rejected by build_count_kernel's error flag; the order of the pairs inside
            // the chunk is irrelevant (sorted by (term, doc) below), so every group of 32 tokens reserves its own slots
            warp_distinct(seq_ids + a, len, [&](int32_t t, int tf, bool first, unsigned m) {
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (first) {
                    const int slot = base + __popc(m & ((1u << lane) - 1u));
                    key[slot] = ((uint64_t)(uint32_t)t << 32) | (uint64_t)(uint32_t)d;
                    val[slot] = tf;
                }
            });
        }
        __syncthreads();
        const int np = s_n;
        int P = 32;
        while (P < np) P <<= 1;
        for (int i = np + tid; i < P; i += BUILD_THREADS) { key[i] = ~0ull; val[i] = 0; }
        __syncthreads();
        for (unsigned size = 2; size <= (unsigned)P; size <<= 1) {
            for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
                for (unsigned t = tid; t < (unsigned)P / 2; t += BUILD_THREADS) {
                    const unsigned i = 2 * t - (t & (stride - 1));
                    const unsigned l = i + stride;
                    const bool up = ((i & size) == 0);
                    const uint64_t x = key[i], y = key[l];
                    if ((x > y) == up) { key[i] = y; key[l] = x; const int32_t v = val[i]; val[i] = val[l]; val[l] = v; }   // ascending
                }
                __syncthreads();
            }
        }
        // every element finds the start of its run of equal terms; the run's length is added to the cursor once
        for (int i = tid; i < np; i += BUILD_THREADS) {
            const uint32_t t = (uint32_t)(key[i] >> 32);
            int s = i;
            while (s > 0 && (uint32_t)(key[s - 1] >> 32) == t) --s;
            const int64_t pos = post_ptr[t] + row[t] + (i - s);
            post_doc[pos] = (int32_t)(uint32_t)key[i];
            post_tf[pos] = val[i];
        }
        __syncthreads();
        for (int i = tid; i < np; i += BUILD_THREADS) {
            const uint32_t t = (uint32_t)(key[i] >> 32);
            if (i == 0 || (uint32_t)(key[i - 1] >> 32) != t) {
                int e = i + 1;
                while (e < np && (uint32_t)(key[e] >> 32) == t) ++e;
                row[t] += e - i;                       // single writer per (CTA, term)
            }
        }
        __syncthreads();
        d0 = d1;
    }
}

__global__ void build_ptr_kernel(const int64_t* __restrict__ df, int32_t n_terms, int64_t* __restrict__ post_ptr) {
    // V is ~1e4: a single thread's serial scan is a few microseconds
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0;
        for (int t = 0; t < n_terms; ++t) { post_ptr[t] = run; run += df[t]; }
        post_ptr[n_terms] = run;
    }
}

}  // namespace ais
