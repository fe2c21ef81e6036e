// Streaming doc-vector scan: sim[q][d] = rows[d,:] . query[q,:]  (fp32), plus per-query max.
//
// Replaces gensim Similarity.__getitem__ -> numpy.dot(shard, q) at the reference's call sites
// webui.py:352 (first pass) and webui.py:205 (PRF re-query).  HBM-bound: every stored row
// (1200 B) is read exactly once per launch and shared by the QT queries of the pass.
//
// Layout / schedule (B200): persistent CTAs, one per SM.  Warp 0 is the producer: one elected
// lane streams 32-row tiles (38 400 contiguous bytes) into a 5-stage shared-memory ring with 1-D
// TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) - ~190 KB in flight per SM, no
// register staging.  Eight consumer warps take tiles round-robin; inside a tile each LANE owns one
// row: with a row stride of 300 words a quarter-warp's eight LDS.128 touch eight distinct 16-B bank
// groups (300/4 = 75 is odd), so the row reads are conflict-free and need no shuffles; the query
// float4 is a broadcast read.  Four partial sums per (row, query) keep the FMA chains short.
#pragma once
#include "common.cuh"

namespace ais {

constexpr int SCAN_STAGES = 5;
constexpr int SCAN_CONSUMERS = 8;
constexpr int SCAN_THREADS = 32 * (1 + SCAN_CONSUMERS);

template <int QT>
constexpr size_t scan_smem_bytes() {
    return (size_t)SCAN_STAGES * TILE_BYTES + (size_t)QT * ROW_BYTES + 2 * SCAN_STAGES * sizeof(uint64_t);
}

template <int QT>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
scan_kernel(const float* __restrict__ rows, int64_t n, const float* __restrict__ queries,  // [QT][DIM]
            float* __restrict__ out, int64_t ld,                                             // [QT][ld]
            uint32_t* __restrict__ max_keys,                                                 // [QT] fkey images
            int nq_live, int evict_first) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    float* qs = reinterpret_cast<float*>(smem_raw + (size_t)SCAN_STAGES * TILE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SCAN_STAGES * TILE_BYTES + (size_t)QT * ROW_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + SCAN_STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    for (int i = tid; i < QT * DIM; i += SCAN_THREADS) qs[i] = queries[i];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SCAN_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
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
            for (int64_t i = 0;; ++i) {
                const int64_t tile = (int64_t)blockIdx.x + i * (int64_t)gridDim.x;
                if (tile >= n_tiles) break;
                const int s = (int)(i % SCAN_STAGES);
                if (i >= SCAN_STAGES) mbar_wait(empty0 + 8 * s, (uint32_t)(((i / SCAN_STAGES) - 1) & 1));
                const int64_t row0 = tile * TILE_ROWS;
                const int64_t left = n - row0;
                const uint32_t bytes = (uint32_t)((left < TILE_ROWS ? left : TILE_ROWS) * ROW_BYTES);
                mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                if (evict_first)
                    bulk_g2s_hint(stage0 + s * TILE_BYTES, rows + row0 * DIM, bytes, full0 + 8 * s, pol);
                else
                    bulk_g2s(stage0 + s * TILE_BYTES, rows + row0 * DIM, bytes, full0 + 8 * s);
            }
        }
        return;
    }

    // ---------------- consumers: lane-per-row dot products out of shared memory ----------------
    const int cw = warp - 1;
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    float lmax[QT];
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) lmax[qi] = -INFINITY;

    for (int64_t i = cw;; i += SCAN_CONSUMERS) {
        const int64_t tile = (int64_t)blockIdx.x + i * (int64_t)gridDim.x;
        if (tile >= n_tiles) break;
        const int s = (int)(i % SCAN_STAGES);
        mbar_wait(full0 + 8 * s, (uint32_t)((i / SCAN_STAGES) & 1));

        const float4* rp = reinterpret_cast<const float4*>(stage_base + (size_t)s * TILE_BYTES) + lane * ROW_F4;
        float4 acc[QT];
#pragma unroll
        for (int qi = 0; qi < QT; ++qi) acc[qi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
        for (int j = 0; j < ROW_F4; ++j) {
            const float4 x = rp[j];
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                const float4 w = q4[qi * ROW_F4 + j];
                acc[qi].x = fmaf(x.x, w.x, acc[qi].x);
                acc[qi].y = fmaf(x.y, w.y, acc[qi].y);
                acc[qi].z = fmaf(x.z, w.z, acc[qi].z);
                acc[qi].w = fmaf(x.w, w.w, acc[qi].w);
            }
        }
        const int64_t row = tile * TILE_ROWS + lane;
        const bool live = row < n;
#pragma unroll
        for (int qi = 0; qi < QT; ++qi) {
            const float v = (acc[qi].x + acc[qi].y) + (acc[qi].z + acc[qi].w);
            if (live && qi < nq_live) {
                out[(int64_t)qi * ld + row] = v;
                lmax[qi] = fmaxf(lmax[qi], v);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
        const float m = warp_max(lmax[qi]);
        if (lane == 0 && qi < nq_live) atomicMax(&max_keys[qi], fkey(m));
    }
}

}  // namespace ais
