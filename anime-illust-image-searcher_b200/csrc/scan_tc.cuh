// scan_tc_kernel: the doc-vector scan for 17..32 queries per pass on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, A operand and accumulators in TMEM, rows staged by TMA).
//
// Same contract as scan_kernel / scan_mma_kernel (scan.cuh): sim[q][d] = rows[d,:] . query[q,:] for the
// gensim call sites webui.py:352 and webui.py:205, every stored row read from HBM once per launch, plus
// the per-query maximum webui.py:377 needs.  Per tile the work is a [128 docs x 300] x [300 x 32 queries]
// contraction; to stay at fp32-level accuracy (TF32 keeps 11 significant bits) both operands are split
// hi + lo and three products are accumulated (3xTF32: hi*hi + hi*lo + lo*hi; lo*lo ~ 2^-22 relative is
// dropped).
//
// Shared-memory bandwidth (128 B/clk/SM) is the scarce resource next to HBM: an SS-mode MMA re-reads its
// 4 KB A slice from shared memory for every instruction, three times per k-step here, and a split pass
// through shared memory reads and writes every element again (first version: 3.9 ms per launch, no pipe
// saturated).  So the doc rows cross shared memory exactly once: TMA writes a box, a split warp reads it
// into registers, and BOTH split images go to TMEM (tcgen05.st), from where the MMA takes its A operand;
// only the 32 queries (B operand, 1 KB per instruction) are read from shared memory by the tensor core.
//
// Pipeline of one persistent CTA per SM (448 threads):
//   warp 12  producer   one lane streams [128 rows x 32 k] boxes (16 KB) of the row matrix through a 2-D
//                       tensor map with 128-byte swizzle into a 9-stage ring (144 KB in flight); the 10th
//                       k-block is zero-filled by TMA beyond column 300, rows beyond n are zero-filled too.
//   warps 4-11 split    warp w owns TMEM lanes 32*(w%4).. = 32 docs of the box (lane = doc; with the
//                       128-byte swizzle the eight 16-B chunks of eight neighbouring rows sit in eight
//                       distinct bank groups, so the row-wise LDS.128 are conflict-free); the two warps of
//                       a lane quarter alternate over the k-blocks: x -> hi = rn_tf32(x), lo = rn_tf32(x - hi)
//                       -> tcgen05.st into one of 4 A stages (64 TMEM columns each); the shared-memory stage
//                       is released as soon as it has been read.
//   warp 13  MMA        one lane issues 3 tcgen05.mma (M=128, N=32, K=8, A from TMEM) per k-step: hi*hi into
//                       one of 3 K-range accumulators (the tensor core's accumulation rounds toward zero;
//                       short chains keep that bias below fp32 rounding), hi*lo and lo*hi into a 4th; commits
//                       release the A stages; accumulators double-buffered (2 x 128 columns).
//   warps 0-3 epilogue  tcgen05.ld (lane = doc), fp32 sum of the 4 accumulators, coalesced stores of
//                       sim[q][128 docs], running per-query maxima.
// The queries arrive pre-split (split_queries_kernel) as [hi(32) | lo(32)][300] and are loaded once per CTA.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ais {

constexpr int TC_M = 128;                       // docs per tile (UMMA M)
constexpr int TC_N_MAX = 64;                    // queries per pass (UMMA N): 32 or 64
constexpr int TC_KB = 32;                       // fp32 elements per k-block = 128 B = the swizzle span
constexpr int TC_NKB = 10;                      // ceil(300 / 32); the last block carries 12 live columns
constexpr int TC_A_BYTES = TC_M * 128;          // 16 KB per (stage)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_EPI_WARPS = 8, TC_SPLIT_WARPS = 8;        // both: two warps per TMEM lane quarter
constexpr int TC_PRODUCER_WARP = TC_EPI_WARPS + TC_SPLIT_WARPS, TC_MMA_WARP = TC_PRODUCER_WARP + 1;
constexpr int TC_THREADS = 32 * (TC_MMA_WARP + 1);

// shared memory: [B hi: 10 k-blocks x N rows x 128 B | B lo | landing ring | barriers | TMEM base]
constexpr int tc_smem_bytes(int n_q, int raw_stages) {                 // + 1024: manual alignment of the base
    return 2 * TC_NKB * n_q * 128 + raw_stages * TC_A_BYTES + (2 * raw_stages + 2 * 8 + 2 + 2 + 1) * 8 + 16 + 1024;
}

// UMMA instruction descriptor: D fp32 (bit 4), A/B TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc(int n_q) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_q >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
// shared-memory matrix descriptor, K-major SWIZZLE_128B: stride between 8-row groups 1024 B, version 1 (Blackwell)
constexpr uint32_t TC_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
    return ((uint64_t)TC_DESC_HI << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16));
}
// D[tmem] (+)= A[tmem: lane = row, column = k] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_2d_hint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// a wait that cannot hang the GPU: a protocol error traps instead of spinning forever (~2 s of spinning; no printf
// here - its argument buffer would put a stack frame and spills into every role of the kernel)
__device__ __forceinline__ void mbar_wait_guarded(uint32_t bar, uint32_t parity, int tag) {
    uint32_t spins = 0;
    long long t0 = 0;
    (void)tag;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFu) == 0) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) __trap();
        }
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Optional per-role timeline of CTA 0 (build with -DAIS_TC_TRACE, read with ais_debug_tc_trace): slot layout
// trace[role][index][0..3]; roles: 0 epilogue (per tile: acc_full seen, loads done, stores done), 1 MMA (per k-block:
// a_full seen, issued), 2 split warp 4 (per iteration: full_raw seen, computed, a_empty seen, arrived), 3 producer.
#ifdef AIS_TC_TRACE
__device__ long long g_tc_trace[4][256][4];
#ifndef AIS_TC_TRACE_BLOCK
#define AIS_TC_TRACE_BLOCK 0
#endif
#define TC_TRACE(role, idx, slot) do { if (blockIdx.x == AIS_TC_TRACE_BLOCK && (idx) < 256) g_tc_trace[role][idx][slot] = clock64(); } while (0)
#else
#define TC_TRACE(role, idx, slot) do { } while (0)
#endif

__device__ __forceinline__ float rn_tf32(float x) {          // round to nearest TF32 (11 significant bits), two integer ops
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// queries [nq][300] -> qsplit [2 * n_pass][300]: rows 0..n_pass-1 hi, then n_pass rows lo (zero rows beyond nq)
__global__ void split_queries_kernel(const float* __restrict__ q, int nq, int n_pass, float* __restrict__ qsplit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pass * DIM) return;
    const float v = (i / DIM) < nq ? q[i] : 0.0f;
    const float hi = rn_tf32(v);
    qsplit[i] = hi;
    qsplit[n_pass * DIM + i] = rn_tf32(v - hi);
}

// TC_N: queries per pass; TC_MAIN: hi*hi accumulators (the k-steps rotate over them); TC_CROSS: accumulators of the
// cross terms; TC_RAW_STAGES: depth of the shared-memory landing ring; TC_A_STAGES: A-operand stages in TMEM
// ([hi 32 columns | lo 32 columns] each); TC_NBUF: accumulator buffers (2: the epilogue of a tile overlaps the MMAs
// of the next one).  Measured on B200, 10 M docs (profiles/r01_d):
//   N = 32: <32,3,1,6,4,2>: 4 accumulators x 32 columns x 2 buffers = 256 columns + 4 A stages; 80 KB of queries,
//           6 x 16 KB ring: 2.10 ms per pass = 12.0 GB read + 1.28 GB written at 6.3 TB/s (0.97 of the copy peak).
//   N = 64: <64,2,1,4,2,2>: 3 accumulators x 64 columns x 2 buffers = 384 columns + 2 A stages; 160 KB of queries,
//           4 x 16 KB ring: 2.67 ms per pass (5.45 TB/s of total traffic).
// What did NOT matter (each tried): ring depth 4/6/9 at N = 32, 2 vs 4 vs 5 A stages, rotating vs K-range accumulators,
// one vs two accumulator buffers at equal instruction count.  At N = 64 (round 2, 10 M docs): <64,1,1,4,4,1,2> (one main
// accumulator, FOUR A stages) 2.654 ms vs 2.676 ms for <64,2,1,4,2,1,2>, <64,2,1,4,4,2,2> (half-k-block A stages)
// 2.94 ms - so the A hand-over is not what holds the 64-query kernel at 0.69 of the copy peak.  The CTA timeline
// (profiles/r01_d_scan_tc64_cta_timeline.txt) shows the split warps waiting 800-1400 clocks for full_raw: 4 x 16 KB per
// SM in flight is too little at the loaded HBM latency (~1.8 us); the 160 KB of query images leave no room for more.
// Halving the query images per SM (tcgen05 cta_group::2: each CTA of a pair holds 32 of the 64 query rows) is the
// open route to a 8-9 stage ring.  What did: the MMA issue sequence (unrolled, uniform registers:
// 3.8 -> 2.2 ms) and the instruction count of the split and epilogue warps, which share four issue slots.
template <int TC_N, int TC_MAIN, int TC_CROSS, int TC_RAW_STAGES, int TC_A_STAGES, int TC_A_SUB, int TC_NBUF>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_q, int64_t n,
               float* __restrict__ out, int64_t ld, uint32_t* __restrict__ max_keys, int nq_live) {
    constexpr int TC_NACC = TC_MAIN + TC_CROSS;
    constexpr int TC_ACC_COLS = TC_NACC * TC_N;
    static_assert(TC_CROSS == 1 || TC_CROSS == 2 || TC_CROSS == 4, "cross accumulators");
    constexpr int TC_A_COL0 = TC_NBUF * TC_ACC_COLS;   // TMEM columns: [acc buffer(s) | A stages]
    // an A stage holds 1 / TC_A_SUB of a k-block: [hi | lo] x TC_SUB_K columns; the hand-over of a stage (commit ->
    // split warp -> tcgen05.st -> MMA warp) takes ~900 clocks whatever its size, so what counts is how many are in flight
    constexpr int TC_SUB_K = TC_KB / TC_A_SUB;                         // k columns per A stage
    constexpr int TC_SUB_STEPS = TC_SUB_K / 8;                         // MMA k-steps per A stage
    constexpr int TC_SUBS_PER_TILE = (TC_NKB - 1) * TC_A_SUB + (TC_A_SUB == 1 ? 1 : TC_A_SUB / 2);   // the last k-block has 2 live k-steps
    static_assert(TC_A_SUB == 1 || TC_A_SUB == 2, "A sub-stages per k-block");
    static_assert(TC_A_COL0 + TC_A_STAGES * 2 * TC_SUB_K <= TC_TMEM_COLS, "TMEM budget");
    static_assert(TC_A_STAGES <= 8 && (TC_NBUF == 1 || TC_NBUF == 2), "barrier slots");
    constexpr int TC_B_BYTES = TC_N * 128;          // one k-block of the hi (or lo) query image
    constexpr uint32_t TC_IDESC = tc_idesc(TC_N);
    constexpr int TC_OFF_BHI = 0, TC_OFF_BLO = TC_NKB * TC_B_BYTES, TC_OFF_RAW = 2 * TC_NKB * TC_B_BYTES;
    constexpr int TC_OFF_BAR = TC_OFF_RAW + TC_RAW_STAGES * TC_A_BYTES;
    constexpr int TC_N_BARS = 2 * TC_RAW_STAGES + 2 * TC_A_STAGES + 2 + 2 + 1;
    extern __shared__ unsigned char smem_unaligned[];
    const uint32_t base = (smem_u32(smem_unaligned) + 1023u) & ~1023u;
    unsigned char* gbase = smem_unaligned + (base - smem_u32(smem_unaligned));
    const uint32_t bar0 = base + TC_OFF_BAR;
    const uint32_t full_raw = bar0, empty_raw = bar0 + 8 * TC_RAW_STAGES;
    const uint32_t a_full = bar0 + 16 * TC_RAW_STAGES, a_empty = a_full + 8 * TC_A_STAGES;
    const uint32_t acc_full = a_empty + 8 * TC_A_STAGES, acc_empty = acc_full + 16, b_full = acc_empty + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + TC_OFF_BAR + TC_N_BARS * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < TC_RAW_STAGES; ++s) {
            mbar_init(full_raw + 8 * s, 1);
            mbar_init(empty_raw + 8 * s, TC_SPLIT_WARPS / 2);
        }
        for (int s = 0; s < TC_A_STAGES; ++s) {
            mbar_init(a_full + 8 * s, TC_SPLIT_WARPS / 2);
            mbar_init(a_empty + 8 * s, 1);
        }
        for (int b = 0; b < TC_NBUF; ++b) {
            mbar_init(acc_full + 8 * b, 1);
            mbar_init(acc_empty + 8 * b, TC_EPI_WARPS);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + TC_OFF_BAR + TC_N_BARS * 8),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles = (int)((n + TC_M - 1) / TC_M);           // 32-bit tile arithmetic (n < 2^38 docs): fewer live registers
    // every CTA owns a CONTIGUOUS run of tiles: its 32 / 64 output streams sim[q][...] then advance sequentially
    // (512 B per tile and query), which the DRAM write path likes better than 148 CTAs hopping through each stream
    const int tiles_per_cta = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile0 = (int)blockIdx.x * tiles_per_cta;
    const int my_tiles = tile0 < n_tiles ? (n_tiles - tile0 < tiles_per_cta ? n_tiles - tile0 : tiles_per_cta) : 0;
    const int total_it = my_tiles * TC_NKB;

    if (warp == TC_PRODUCER_WARP) {
        // ---------------- producer ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_rows)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q)) : "memory");
            mbar_arrive_expect_tx(b_full, 2 * TC_NKB * TC_B_BYTES);
            for (int kb = 0; kb < TC_NKB; ++kb) {
                tma_2d(base + TC_OFF_BHI + kb * TC_B_BYTES, &tm_q, kb * TC_KB, 0, b_full);
                tma_2d(base + TC_OFF_BLO + kb * TC_B_BYTES, &tm_q, kb * TC_KB, TC_N, b_full);
            }
            int it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int row0 = (int)((tile0 + t) * TC_M);
                for (int kb = 0; kb < TC_NKB; ++kb, ++it) {
                    const int s = it % TC_RAW_STAGES;
                    mbar_wait_guarded(empty_raw + 8 * s, ((it / TC_RAW_STAGES) & 1) ^ 1, 0);
                    TC_TRACE(3, it, 0);
                    mbar_arrive_expect_tx(full_raw + 8 * s, TC_A_BYTES);
                    tma_2d(base + TC_OFF_RAW + s * TC_A_BYTES, &tm_rows, kb * TC_KB, row0, full_raw + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp == TC_MMA_WARP) {
        // ---------------- MMA issuer ----------------
        // The whole warp walks the loop (warp-uniform control flow keeps the descriptors in uniform registers);
        // one elected lane issues.  At N = 32 an MMA occupies the tensor pipe for 16 clocks only, so the issue
        // sequence itself must stay short: everything is unrolled and the descriptors advance by constants.
        mbar_wait_guarded(b_full, 0, 1);
        tc_fence_after();
        const uint64_t b_hi0 = tc_desc(base + TC_OFF_BHI), b_lo0 = tc_desc(base + TC_OFF_BLO);
        int it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t % TC_NBUF;
            mbar_wait_guarded(acc_empty + 8 * buf, ((t / TC_NBUF) & 1) ^ 1, 2);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * TC_ACC_COLS;
#pragma unroll
            for (int kb = 0; kb < TC_NKB; ++kb, ++it) {
                const uint64_t b_hi = b_hi0 + (uint64_t)(kb * (TC_B_BYTES >> 4)), b_lo = b_lo0 + (uint64_t)(kb * (TC_B_BYTES >> 4));
                const int nsub = kb == TC_NKB - 1 ? (TC_A_SUB + 1) / 2 : TC_A_SUB;   // 300 = 9*32 + 12 -> two k-steps of 8 in the last block
#pragma unroll
                for (int h = 0; h < TC_A_SUB; ++h) {
                    if (h >= nsub) break;
                    const int hs = t * TC_SUBS_PER_TILE + kb * TC_A_SUB + h;
                    const int as = hs % TC_A_STAGES;
                    mbar_wait_guarded(a_full + 8 * as, (hs / TC_A_STAGES) & 1, 3);
                    if (lane == 0) TC_TRACE(1, hs, 0);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + TC_A_COL0 + as * 2 * TC_SUB_K, a_lo = a_hi + TC_SUB_K;
                    if (elect_one()) {
#pragma unroll
                        for (int k2 = 0; k2 < TC_SUB_STEPS; ++k2) {
                            const int ks = h * TC_SUB_STEPS + k2;                       // k-step inside the k-block
                            if (kb < TC_NKB - 1 || ks < 2) {
                                // Accumulators rotate with the k-step (consecutive MMAs target different TMEM tiles)
                                const int g = kb * 4 + ks;                              // k-step of the tile, 0..37
                                const int cm = g % TC_MAIN;
                                const int ca = TC_MAIN + (TC_CROSS == 4 ? (g & 1) * 2 : 0);
                                const int cb = TC_CROSS == 1 ? ca : ca + 1;
                                const bool first_c = g < (TC_CROSS == 4 ? 2 : 1);
                                // + 2: 32 B along K inside the swizzle span, in 16-B units
                                tc_mma_ts(acc + cm * TC_N, a_hi + k2 * 8, b_hi + 2 * ks, TC_IDESC, g < TC_MAIN ? 0u : 1u);
                                tc_mma_ts(acc + ca * TC_N, a_hi + k2 * 8, b_lo + 2 * ks, TC_IDESC, first_c ? 0u : 1u);
                                tc_mma_ts(acc + cb * TC_N, a_lo + k2 * 8, b_hi + 2 * ks, TC_IDESC, (first_c && TC_CROSS != 1) ? 0u : 1u);
                            }
                        }
                        tc_commit(a_empty + 8 * as);
                        if (kb == TC_NKB - 1 && h == nsub - 1) tc_commit(acc_full + 8 * buf);
                        TC_TRACE(1, hs, 1);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= TC_EPI_WARPS) {
        // ---------------- split warps: box -> registers -> hi | lo in TMEM ----------------
        const int g = (warp - TC_EPI_WARPS) >> 2;                            // the two warps of a lane quarter take even / odd k-blocks
        const int quarter = warp & 3;                                        // TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int r = quarter * 32 + lane;                                   // row of the box = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + TC_A_COL0;
        for (int it = g; it < total_it; it += 2) {
            const int s = it % TC_RAW_STAGES;
            const int t = it / TC_NKB, kb = it - t * TC_NKB;
            mbar_wait_guarded(full_raw + 8 * s, (it / TC_RAW_STAGES) & 1, 4);
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 0);
            const unsigned char* rowp = gbase + TC_OFF_RAW + s * TC_A_BYTES + r * 128;
            float4 x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float v[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // hi = RN_tf32(x) (two integer ops); lo = x - hi exactly (<= 13 significant bits), which the tensor
                    // core truncates to TF32 itself: |error| <= 2^-21 |x|, sign-symmetric.  Three instructions per element:
                    // the split warps share their issue slots with the epilogue, and every instruction here counts.
                    const float h = rn_tf32(v[j]);
                    hi[4 * c + j] = __float_as_uint(h);
                    lo[4 * c + j] = __float_as_uint(v[j] - h);
                }
            }
            __syncwarp();
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 1);
            if (lane == 0) mbar_arrive(empty_raw + 8 * s);                   // the box is in registers: refill the stage
            const int nsub = kb == TC_NKB - 1 ? (TC_A_SUB + 1) / 2 : TC_A_SUB;
#pragma unroll
            for (int h = 0; h < TC_A_SUB; ++h) {
                if (h >= nsub) break;
                const int hs = t * TC_SUBS_PER_TILE + kb * TC_A_SUB + h;
                const int as = hs % TC_A_STAGES;
                mbar_wait_guarded(a_empty + 8 * as, ((hs / TC_A_STAGES) & 1) ^ 1, 5);
                if (warp == TC_EPI_WARPS && lane == 0 && h == 0) TC_TRACE(2, it / 2, 2);
                tc_fence_after();
                const uint32_t ta = lane_addr + as * 2 * TC_SUB_K;
                if constexpr (TC_A_SUB == 1) {
                    tc_st16(ta, reinterpret_cast<const uint32_t(&)[16]>(hi[0]));
                    tc_st16(ta + 16, reinterpret_cast<const uint32_t(&)[16]>(hi[16]));
                    tc_st16(ta + 32, reinterpret_cast<const uint32_t(&)[16]>(lo[0]));
                    tc_st16(ta + 48, reinterpret_cast<const uint32_t(&)[16]>(lo[16]));
                } else {
                    tc_st16(ta, reinterpret_cast<const uint32_t(&)[16]>(hi[16 * h]));
                    tc_st16(ta + 16, reinterpret_cast<const uint32_t(&)[16]>(lo[16 * h]));
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full + 8 * as);
                if (warp == TC_EPI_WARPS && lane == 0 && h == nsub - 1) TC_TRACE(2, it / 2, 3);
            }
        }
    } else {
        // ---------------- epilogue warps: TMEM -> registers -> sim[q][doc] ----------------
        // Lane = doc; warp w reads TMEM lanes 32 * (w % 4) .. and the query columns of half w / 4.  A tcgen05.ld round
        // trip costs ~700 clocks while the MMAs keep TMEM busy (measured with the CTA timeline, tools/tc_trace.py:
        // 8 dependent rounds per tile made the epilogue, 6000 clocks, the longest stage of the pipeline), so a warp
        // issues all the loads of 16 query columns at once - one or two rounds per tile.  Per (doc, query): the fp32
        // sum of the accumulators, one coalesced store (a warp writes 128 contiguous bytes of sim[q]) and one FMNMX
        // into the lane's running maximum of that query; the lanes' maxima meet once, at the end.  Column chunks
        // beyond the live queries of the pass are skipped; inside the last live chunk the zero-vector padding columns
        // are stored too - the work arrays hold a multiple of 16 rows.
        constexpr int QW = TC_N / 2;                                         // query columns per epilogue warp
        constexpr int CH = 16;
        const int quarter = warp & 3, half = warp >> 2;
        float lmax[QW];
#pragma unroll
        for (int q = 0; q < QW; ++q) lmax[q] = -INFINITY;
        int n_chunks = (nq_live - half * QW + CH - 1) / CH;                  // live chunks of this warp's columns
        n_chunks = n_chunks < 0 ? 0 : (n_chunks > QW / CH ? QW / CH : n_chunks);
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t % TC_NBUF;
            mbar_wait_guarded(acc_full + 8 * buf, (t / TC_NBUF) & 1, 6);
            if (warp == 0 && lane == 0) TC_TRACE(0, t, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TC_ACC_COLS + half * QW;
            const int64_t row = (int64_t)(tile0 + t) * TC_M + quarter * 32 + lane;
            const bool live = row < n;
            float* orow = out + row + (int64_t)(half * QW) * ld;
            if (n_chunks == 0) {                                             // nothing to read: hand the buffer back at once
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
            }
#pragma unroll
            for (int c = 0; c < QW / CH; ++c) {
                if (c >= n_chunks) break;                                    // warp-uniform
                uint32_t a[TC_NACC][CH];
#pragma unroll
                for (int m = 0; m < TC_NACC; ++m) tc_ld16(taddr + m * TC_N + c * CH, a[m]);
                tc_wait_ld();
                if (c == n_chunks - 1) {                                     // the tile is out of TMEM: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
                    if (warp == 0 && lane == 0) TC_TRACE(0, t, 1);
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    float x = __uint_as_float(a[0][j]);
#pragma unroll
                    for (int m = 1; m < TC_MAIN; ++m) x += __uint_as_float(a[m][j]);
                    float y = __uint_as_float(a[TC_MAIN][j]);
#pragma unroll
                    for (int m = TC_MAIN + 1; m < TC_NACC; ++m) y += __uint_as_float(a[m][j]);
                    const float v = x + y;                                   // the hi*hi sums, then the small cross terms
                    const int q = c * CH + j;
                    if (live) {
                        orow[(int64_t)q * ld] = v;
                        lmax[q] = fmaxf(lmax[q], v);
                    }
                }
            }
            if (warp == 0 && lane == 0) TC_TRACE(0, t, 2);
        }
#pragma unroll
        for (int q = 0; q < QW; ++q) {
            const float m = warp_max(lmax[q]);
            if (lane == 0 && half * QW + q < nq_live) atomicMax(&max_keys[half * QW + q], fkey(m));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace ais
