// Streaming doc-vector scan: sim[q][d] = rows[d,:] . query[q,:]  (fp32), plus per-query max.
//
// Replaces gensim Similarity.__getitem__ -> numpy.dot(shard, q) at the reference's call sites
// webui.py:352 (first pass) and webui.py:205 (PRF re-query).  HBM-bound: every stored row
// (1200 B) is read exactly once per launch and shared by the QT queries of the pass.
//
// Layout / schedule (B200): persistent CTAs, one per SM.  Warp 0 is the producer: one elected
// lane streams 32-row tiles (38 400 contiguous bytes) into a 5-stage shared-memory ring with 1-D
// TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) - ~190 KB in flight per SM, no
// register staging.  One (two when the pass carries several queries) consumer warp is bound to every stage; inside a tile each LANE owns one
// row: with a row stride of 300 words a quarter-warp's eight LDS.128 touch eight distinct 16-B bank
// groups (300/4 = 75 is odd), so the row reads are conflict-free and need no shuffles; the query
// float4 is a broadcast read.  Four partial sums per (row, query) keep the FMA chains short.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int SCAN_STAGES = 5;

// warps per stage: the queries of a pass are split between them (QT = 1 needs only one)
template <int QT> struct ScanCfg {
    static constexpr int WPS = QT >= 2 ? 2 : 1;
    static constexpr int QH = QT / WPS;                          // queries per consumer warp
    static constexpr int CONSUMERS = SCAN_STAGES * WPS;
    static constexpr int THREADS = 32 * (1 + CONSUMERS);
};

template <int QT>
constexpr size_t scan_smem_bytes() {
    return (size_t)SCAN_STAGES * TILE_BYTES + (size_t)QT * ROW_BYTES + 2 * SCAN_STAGES * sizeof(uint64_t);
}

// Ring protocol.  Stage s carries this CTA's tiles s, s+S, s+2S, ... ; its consumer warps are BOUND to
// the stage, so each of them meets the phases of full[s] strictly in order (an mbarrier parity wait can
// only tell adjacent phases apart - a warp that ran a phase ahead would fall through the wait).
template <int QT>
__global__ void __launch_bounds__(ScanCfg<QT>::THREADS, 1)
scan_kernel(const float* __restrict__ rows, int64_t n, const float* __restrict__ queries,  // [QT][DIM]
            float* __restrict__ out, int64_t ld,                                             // [QT][ld]
            uint32_t* __restrict__ max_keys,                                                 // [QT] fkey images
            int nq_live, int evict_first) {
    using Cfg = ScanCfg<QT>;
    constexpr int QH = Cfg::QH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    float* qs = reinterpret_cast<float*>(smem_raw + (size_t)SCAN_STAGES * TILE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SCAN_STAGES * TILE_BYTES + (size_t)QT * ROW_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + SCAN_STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    for (int i = tid; i < QT * DIM; i += Cfg::THREADS) qs[i] = queries[i];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SCAN_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::WPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;

    if (warp == 0) {
        // ---------------- producer: one lane drives the TMA ring ----------------
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t stage0 = smem_u32(stage_base);
            int s = 0;
            uint32_t round = 0;                                   // how many times the ring has wrapped
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1u);
                const int64_t row0 = tile * TILE_ROWS;
                const int64_t left = n - row0;
                const uint32_t bytes = (uint32_t)((left < TILE_ROWS ? left : TILE_ROWS) * ROW_BYTES);
                mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                if (evict_first)
                    bulk_g2s_hint(stage0 + s * TILE_BYTES, rows + row0 * DIM, bytes, full0 + 8 * s, pol);
                else
                    bulk_g2s(stage0 + s * TILE_BYTES, rows + row0 * DIM, bytes, full0 + 8 * s);
                if (++s == SCAN_STAGES) { s = 0; ++round; }
            }
        }
        return;
    }

    // ---------------- consumers: lane-per-row dot products out of shared memory ----------------
    const int cw = warp - 1;
    const int s = cw % SCAN_STAGES;                               // the stage this warp is bound to
    const int q0 = (cw / SCAN_STAGES) * QH;                       // its first query
    const float4* q4 = reinterpret_cast<const float4*>(qs) + q0 * ROW_F4;
    const float4* rp = reinterpret_cast<const float4*>(stage_base + (size_t)s * TILE_BYTES) + lane * ROW_F4;
    float lmax[QH];
#pragma unroll
    for (int qi = 0; qi < QH; ++qi) lmax[qi] = -INFINITY;

    uint32_t phase = 0;
    for (int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x; tile < n_tiles;
         tile += (int64_t)SCAN_STAGES * gridDim.x, phase ^= 1u) {
        mbar_wait(full0 + 8 * s, phase);

        float4 acc[QH];
#pragma unroll
        for (int qi = 0; qi < QH; ++qi) acc[qi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
        for (int j = 0; j < ROW_F4; ++j) {
            const float4 x = rp[j];
#pragma unroll
            for (int qi = 0; qi < QH; ++qi) {
                const float4 w = q4[qi * ROW_F4 + j];
                acc[qi].x = fmaf(x.x, w.x, acc[qi].x);
                acc[qi].y = fmaf(x.y, w.y, acc[qi].y);
                acc[qi].z = fmaf(x.z, w.z, acc[qi].z);
                acc[qi].w = fmaf(x.w, w.w, acc[qi].w);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);               // the stage may be refilled now
        const int64_t row = tile * TILE_ROWS + lane;
        const bool live = row < n;
#pragma unroll
        for (int qi = 0; qi < QH; ++qi) {
            const float v = (acc[qi].x + acc[qi].y) + (acc[qi].z + acc[qi].w);
            if (live && q0 + qi < nq_live) {
                out[(int64_t)(q0 + qi) * ld + row] = v;
                lmax[qi] = fmaxf(lmax[qi], v);
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < QH; ++qi) {
        const float m = warp_max(lmax[qi]);
        if (lane == 0 && q0 + qi < nq_live) atomicMax(&max_keys[q0 + qi], fkey(m));
    }
}

// =================================================================================================
// scan_mma_kernel<QT> (QT = 8, 16): the same pass for many queries on the tensor cores.
//
// At >= 8 queries per pass the lane-per-row kernel above is bound by shared-memory wavefronts and FP32
// issue (profiles/r01_b: 3.7 ms vs 1.75 ms at one query), while the work is a real [32 x 300] x [300 x QT]
// contraction per tile.  Here every consumer warp owns 16 rows of a tile and issues
// mma.sync.m16n8k8 TF32 instructions with fp32 accumulation.  To stay at fp32-level accuracy (TF32
// keeps 10 mantissa bits) both operands are split hi + lo (3xTF32: lo*hi + hi*lo + hi*hi; the dropped
// lo*lo term is ~2^-22 relative): the query split is done once per CTA into shared memory, the row split
// in registers right after the LDS, so no extra shared-memory traffic is spent on it.
// Fragment loads are conflict-free: row stride 300 words -> (gid*12 + tig) mod 32 distinct; query stride
// 308 words -> (gid*20 + tig) mod 32 distinct.
// =================================================================================================
constexpr int MMA_STAGES = 4;
constexpr int MMA_CONSUMERS = 2 * MMA_STAGES;          // two warps per stage, 16 rows each
constexpr int MMA_THREADS = 32 * (1 + MMA_CONSUMERS);
constexpr int QSTRIDE = 308;                           // padded query row (words): K = 304 zero-padded + bank skew
constexpr int K_STEPS = 38;                            // 38 * 8 = 304 >= 300

template <int QT>
constexpr size_t scan_mma_smem_bytes() {
    return (size_t)MMA_STAGES * TILE_BYTES + 2 * (size_t)QT * QSTRIDE * sizeof(float) + 2 * MMA_STAGES * sizeof(uint64_t);
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int QT>
__global__ void __launch_bounds__(MMA_THREADS, 1)
scan_mma_kernel(const float* __restrict__ rows, int64_t n, const float* __restrict__ queries,  // [QT][DIM]
                float* __restrict__ out, int64_t ld, uint32_t* __restrict__ max_keys, int nq_live) {
    constexpr int NT = QT / 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    float* q_hi = reinterpret_cast<float*>(smem_raw + (size_t)MMA_STAGES * TILE_BYTES);
    float* q_lo = q_hi + QT * QSTRIDE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(q_lo + QT * QSTRIDE);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + MMA_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < QT * QSTRIDE; i += MMA_THREADS) {
        const int q = i / QSTRIDE, k = i - q * QSTRIDE;
        const float v = k < DIM ? queries[q * DIM + k] : 0.0f;
        const float hi = __uint_as_float(to_tf32(v));
        q_hi[i] = hi;
        q_lo[i] = __uint_as_float(to_tf32(v - hi));
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MMA_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 2);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t stage0 = smem_u32(stage_base);
            int s = 0;
            uint32_t round = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1u);
                const int64_t row0 = tile * TILE_ROWS;
                const int64_t left = n - row0;
                const uint32_t bytes = (uint32_t)((left < TILE_ROWS ? left : TILE_ROWS) * ROW_BYTES);
                mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                bulk_g2s_hint(stage0 + s * TILE_BYTES, rows + row0 * DIM, bytes, full0 + 8 * s, pol);
                if (++s == MMA_STAGES) { s = 0; ++round; }
            }
        }
        return;
    }

    const int cw = warp - 1;
    const int s = cw % MMA_STAGES;            // bound stage
    const int half = cw / MMA_STAGES;         // rows [16*half, 16*half + 16) of the tile
    const int gid = lane >> 2, tig = lane & 3;
    const float* tile_rows = reinterpret_cast<const float*>(stage_base + (size_t)s * TILE_BYTES) + (half * 16 + gid) * DIM + tig;
    const float* bh = q_hi + gid * QSTRIDE + tig;
    const float* bl = q_lo + gid * QSTRIDE + tig;
    float lmax[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) lmax[t][0] = lmax[t][1] = -INFINITY;

    uint32_t phase = 0;
    for (int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x; tile < n_tiles;
         tile += (int64_t)MMA_STAGES * gridDim.x, phase ^= 1u) {
        mbar_wait(full0 + 8 * s, phase);
        float acc[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.0f;

#pragma unroll 2
        for (int ks = 0; ks < K_STEPS; ++ks) {
            const int k0 = ks * 8;
            float av[4];
            av[0] = tile_rows[k0];
            av[1] = tile_rows[8 * DIM + k0];
            const bool in = (k0 + 4 + tig) < DIM;                  // the last step runs past the 300 columns
            av[2] = in ? tile_rows[k0 + 4] : 0.0f;
            av[3] = in ? tile_rows[8 * DIM + k0 + 4] : 0.0f;
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // hi = row value rounded to TF32 with two full-rate integer ops (cvt.rna.tf32 issues at a quarter
                // of that rate); lo = the exact remainder, truncated to TF32 by the tensor core itself
                ah[i] = (__float_as_uint(av[i]) + 0x1000u) & 0xffffe000u;
                al[i] = __float_as_uint(av[i] - __uint_as_float(ah[i]));
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const uint32_t bh0 = __float_as_uint(bh[t * 8 * QSTRIDE + k0]);
                const uint32_t bh1 = __float_as_uint(bh[t * 8 * QSTRIDE + k0 + 4]);
                const uint32_t bl0 = __float_as_uint(bl[t * 8 * QSTRIDE + k0]);
                const uint32_t bl1 = __float_as_uint(bl[t * 8 * QSTRIDE + k0 + 4]);
                // the three partial products of this k-step start from zero and are added to the running sums
                // with ordinary fp32 adds: the tensor core's internal accumulation rounds toward zero, and that
                // bias would otherwise build up over the 38 steps
                float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                mma_tf32(d, al, bh0, bh1);
                mma_tf32(d, ah, bl0, bl1);
                mma_tf32(d, ah, bh0, bh1);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[t][i] += d[i];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        const int64_t r0 = tile * TILE_ROWS + half * 16 + gid;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = t * 8 + 2 * tig + c;
                if (q < nq_live) {
                    if (r0 < n) { out[(int64_t)q * ld + r0] = acc[t][c]; lmax[t][c] = fmaxf(lmax[t][c], acc[t][c]); }
                    if (r0 + 8 < n) { out[(int64_t)q * ld + r0 + 8] = acc[t][2 + c]; lmax[t][c] = fmaxf(lmax[t][c], acc[t][2 + c]); }
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float m = lmax[t][c];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            const int q = t * 8 + 2 * tig + c;
            if (gid == 0 && q < nq_live) atomicMax(&max_keys[q], fkey(m));
        }
}

// =================================================================================================
// column_scan_kernel: the scan for query vectors with ONE non-zero component.
//
// The reference's PRF re-query (webui.py:200-205) is such a vector: the (300,2) array of (index, value) pairs is
// normalised as a whole and its indices are round()-ed to 0, gensim's sparse2full keeps the last value for id 0, so
// index[.] receives [c, 0, ..., 0] and its scores are rows[d][0] * c - numpy's fp32 dot product of a
// row with that vector is exactly RN(rows[d][0] * c), every other product being +-0.  Reading one 32-byte sector per
// doc instead of the 1200-byte row gives the same bits for 1/37 of the traffic, shared by all queries of the batch.
// =================================================================================================
constexpr int COL_QC = 16;            // queries per block row
constexpr int COL_THREADS = 256;

__global__ void __launch_bounds__(COL_THREADS)
column_scan_kernel(const float* __restrict__ rows, int64_t n, int comp, const float* __restrict__ queries,  // [nq][DIM]
                   int nq, float* __restrict__ out, int64_t ld, uint32_t* __restrict__ max_keys) {
    __shared__ float c[COL_QC];
    const int q0 = blockIdx.y * COL_QC;
    const int live = nq - q0 < COL_QC ? nq - q0 : COL_QC;
    if (threadIdx.x < COL_QC) c[threadIdx.x] = threadIdx.x < live ? queries[(size_t)(q0 + threadIdx.x) * DIM + comp] : 0.0f;
    __syncthreads();
    float lmax[COL_QC];
#pragma unroll
    for (int q = 0; q < COL_QC; ++q) lmax[q] = -INFINITY;
    const int64_t stride = (int64_t)gridDim.x * COL_THREADS;
    for (int64_t d = (int64_t)blockIdx.x * COL_THREADS + threadIdx.x; d < n; d += stride) {
        const float x = __ldg(rows + d * DIM + comp);
#pragma unroll
        for (int q = 0; q < COL_QC; ++q) {
            if (q < live) {
                const float v = __fmul_rn(x, c[q]);
                out[(int64_t)(q0 + q) * ld + d] = v;
                lmax[q] = fmaxf(lmax[q], v);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < COL_QC; ++q) {
        const float m = warp_max(lmax[q]);
        if ((threadIdx.x & 31) == 0 && q < live && m > -INFINITY) atomicMax(&max_keys[q0 + q], fkey(m));
    }
}

// column `comp` of the row matrix as a compact fp32 array (40 MB at 10 M docs: L2-resident for all queries of a batch).
// The pass-2 kernels form rer[d] = RN(col[d] * c_q) on the fly, so the collapsed re-query needs no per-query array at all.
__global__ void extract_column_kernel(const float* __restrict__ rows, int64_t n, int comp, float* __restrict__ col) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) col[d] = __ldg(rows + d * DIM + comp);
}

}  // namespace ais
