// Shared device helpers for the ais_b200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ais {

constexpr int DIM = 300;                         // VECTOR_LENGTH, genmodel.py:16
constexpr int ROW_F4 = DIM / 4;                  // 75 float4 per stored row
constexpr int ROW_BYTES = DIM * 4;               // 1200 B, 16-B aligned rows
constexpr int TILE_ROWS = 32;                    // one row per lane
constexpr int TILE_BYTES = TILE_ROWS * ROW_BYTES;  // 38400 B per TMA bulk copy

constexpr int MAX_TERMS = 64;                    // AIS_MAX_TERMS
constexpr int MAX_QT = 16;                       // queries sharing one pass over the doc vectors (one scan launch)
constexpr int MAX_BATCH = 256;                   // queries per engine batch (BASELINE configs[2]); up to 64 share one pass over the rows
constexpr int MAX_DEPTH = 16;                    // PRF depth upper bound (reference uses 10)

// ---- order-preserving integer images of scores -------------------------------------------
// larger score <=> larger key; key 0 is reserved for "empty slot" (every real score, -inf
// included, maps to a key > 0).  -0.0 is folded onto +0.0 so that ties compare equal like in
// Python's sort key (webui.py:192).
__host__ __device__ __forceinline__ uint64_t dkey(double x) {
    uint64_t b;
#ifdef __CUDA_ARCH__
    b = (uint64_t)__double_as_longlong(x);
#else
    memcpy(&b, &x, 8);
#endif
    if (b == 0x8000000000000000ull) b = 0;
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dkey_inv(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double x;
#ifdef __CUDA_ARCH__
    x = __longlong_as_double((long long)b);
#else
    memcpy(&x, &b, 8);
#endif
    return x;
}
__host__ __device__ __forceinline__ uint32_t fkey(float x) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(x);
#else
    memcpy(&b, &x, 4);
#endif
    if (b == 0x80000000u) b = 0;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float fkey_inv(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float x;
#ifdef __CUDA_ARCH__
    x = __uint_as_float(b);
#else
    memcpy(&x, &b, 4);
#endif
    return x;
}

struct QueryTerms {  // device-resident, one per query of the batch (webui.py:354-371: {term id: weight} in dict order)
    int32_t n_terms;
    int32_t n_required;                          // terms with weight > REQUIRE_TAG_MAGIC_NUMBER (webui.py:161); host-filled
    int32_t term[MAX_TERMS];
    int32_t slot[MAX_TERMS];                     // index of the term among the DISTINCT terms of the batch (-1: unknown term)
    double weight[MAX_TERMS];
};

constexpr uint64_t KEY_EMPTY = 0ull;

constexpr int64_t ID_EMPTY = 0x7FFFFFFFFFFFFFFFll;

// (key desc, id asc): the reference's stable sort by -score over enumerate(...) (webui.py:191-192,237)
__device__ __forceinline__ bool better(uint64_t ka, int64_t ia, uint64_t kb, int64_t ib) {
    return ka > kb || (ka == kb && ia < ib);
}

#ifdef __CUDACC__
// ---- mbarrier / TMA bulk-copy PTX ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// 64-bit warp maximum with two redux.sync instead of five shuffle rounds
__device__ __forceinline__ uint64_t warp_max_u64_redux(uint64_t v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((uint64_t)mh << 32) | ml;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}
#endif

}  // namespace ais
