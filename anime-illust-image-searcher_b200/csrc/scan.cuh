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

}  // namespace ais
