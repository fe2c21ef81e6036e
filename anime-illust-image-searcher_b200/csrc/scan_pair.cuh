// scan_pair_kernel: the 64-queries-per-pass doc-vector scan on a CTA PAIR (tcgen05 cta_group::2).
//
// Same contract and arithmetic as scan_tc_kernel<64,...> (scan_tc.cuh; webui.py:352 / :205): sim[q][d] = rows[d,:] .
// query[q,:] as 3xTF32 (hi*hi into two rotating accumulators, hi*lo + lo*hi into a third), every stored row read from
// HBM once per launch, per-query maxima for webui.py:377.  The single-CTA kernel keeps both split images of all 64
// queries in shared memory (160 KB), which leaves 4 x 16 KB for the landing ring - too few bytes in flight at the loaded
// HBM latency (profiles/r01_d_scan_tc64_cta_timeline.txt: the split warps wait 800-1400 clocks for data; 0.69 of the
// copy peak).  Here the two CTAs of a cluster form one UMMA of M = 256 docs x N = 64 queries: each CTA holds HALF of
// the query rows (32 of 64; the tensor cores of the pair exchange their B halves over the pair link), 80 KB instead of
// 160, and the ring grows to TCP_RAW stages of 16 KB.
//
// Per CTA the roles are those of scan_tc_kernel (producer / split warps / MMA issuer / epilogue warps), with the
// barriers placed as follows (leader = cluster rank 0, the only CTA that issues tcgen05.mma):
//   full_raw / empty_raw   local   own TMA producer <-> own split warps
//   a_full[s]              LEADER  split warps of BOTH CTAs arrive (the peer's by a remote mbarrier arrive)
//   a_empty[s], acc_full   both    tcgen05.commit ... multicast::cluster with mask 0b11
//   acc_empty[b]           LEADER  epilogue warps of both CTAs arrive
//   b_peer                 LEADER  the peer tells that its half of the query images has landed
// Tiles: pair p owns a contiguous run of tile pairs; in step j CTA r works on tile 2 * (first + j) + r.  Rows beyond n
// are zero-filled by TMA and masked in the epilogue, so both CTAs always run the same number of steps.
#pragma once
#include "scan_tc.cuh"

namespace ais {

constexpr int TCP_N = 64;                        // queries per pass (UMMA N)
constexpr int TCP_NH = TCP_N / 2;                // query rows held by one CTA
constexpr int TCP_MAIN = 2, TCP_CROSS = 1, TCP_NACC = TCP_MAIN + TCP_CROSS;
constexpr int TCP_ACC_COLS = TCP_NACC * TCP_N;   // 192 columns per accumulator buffer
constexpr int TCP_NBUF = 2;
constexpr int TCP_A_STAGES = 2;                  // [hi 32 | lo 32] columns each
constexpr int TCP_A_COL0 = TCP_NBUF * TCP_ACC_COLS;
static_assert(TCP_A_COL0 + TCP_A_STAGES * 2 * TC_KB <= TC_TMEM_COLS, "TMEM budget");
constexpr int TCP_B_BYTES = TCP_NH * 128;        // one k-block of this CTA's half of the hi (or lo) query image

constexpr int tcp_smem_bytes(int raw_stages) {
    return 2 * TC_NKB * TCP_B_BYTES + raw_stages * TC_A_BYTES + (2 * raw_stages + 2 * TCP_A_STAGES + 2 * TCP_NBUF + 2) * 8 + 16 + 1024;
}
// UMMA instruction descriptor of the pair: M = 256
__host__ __device__ constexpr uint32_t tcp_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCP_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the address of `bar` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// What is handed over lives in TMEM and is ordered by tcgen05.wait / tcgen05.fence::before_thread_sync; the arrive itself
// keeps the default .release.cta (the form CUTLASS' ClusterBarrier::arrive(cta_id) uses).  A .release.cluster arrive cost
// ~2000 clocks per hand-over here (CTA timeline, round 2): it drains the whole SM's memory traffic first.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// once per kernel (the peer's query images have landed in ITS shared memory, written there by TMA, and the leader's
// tensor core is about to read them): generic-memory data crosses CTAs here, so this pair is cluster-scoped
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if ((++spins & 0x3FFu) == 0) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
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
// waits on a barrier other CTAs arrive on; traps instead of spinning forever (see mbar_wait_guarded)
__device__ __forceinline__ void mbar_wait_cluster_guarded(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 0x3FFu) == 0) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void tcp_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at offset `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void tcp_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

template <int TCP_RAW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_pair_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_q, int64_t n,
                 float* __restrict__ out, int64_t ld, uint32_t* __restrict__ max_keys, int nq_live) {
    constexpr uint32_t IDESC = tcp_idesc();
    constexpr int OFF_BHI = 0, OFF_BLO = TC_NKB * TCP_B_BYTES, OFF_RAW = 2 * TC_NKB * TCP_B_BYTES;
    constexpr int OFF_BAR = OFF_RAW + TCP_RAW * TC_A_BYTES;
    constexpr int N_BARS = 2 * TCP_RAW + 2 * TCP_A_STAGES + 2 * TCP_NBUF + 2;
    extern __shared__ unsigned char smem_unaligned[];
    const uint32_t base = (smem_u32(smem_unaligned) + 1023u) & ~1023u;        // the same offset in both CTAs of the pair
    unsigned char* gbase = smem_unaligned + (base - smem_u32(smem_unaligned));
    const uint32_t bar0 = base + OFF_BAR;
    const uint32_t full_raw = bar0, empty_raw = bar0 + 8 * TCP_RAW;
    const uint32_t a_full = bar0 + 16 * TCP_RAW, a_empty = a_full + 8 * TCP_A_STAGES;
    const uint32_t acc_full = a_empty + 8 * TCP_A_STAGES, acc_empty = acc_full + 8 * TCP_NBUF;
    const uint32_t b_full = acc_empty + 8 * TCP_NBUF, b_peer = b_full + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + N_BARS * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (tid == 0) {
        for (int s = 0; s < TCP_RAW; ++s) {
            mbar_init(full_raw + 8 * s, 1);
            mbar_init(empty_raw + 8 * s, TC_SPLIT_WARPS / 2);
        }
        for (int s = 0; s < TCP_A_STAGES; ++s) {
            mbar_init(a_full + 8 * s, TC_SPLIT_WARPS);               // leader's copy counts: 4 warps per stage from each CTA
            mbar_init(a_empty + 8 * s, 1);
        }
        for (int b = 0; b < TCP_NBUF; ++b) {
            mbar_init(acc_full + 8 * b, 1);
            mbar_init(acc_empty + 8 * b, 2 * TC_EPI_WARPS);          // leader's copy counts: the epilogue warps of both CTAs
        }
        mbar_init(b_full, 1);
        mbar_init(b_peer, 1);
        fence_mbar_init();
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + OFF_BAR + N_BARS * 8),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // both CTAs' barriers exist before anyone arrives on a remote one
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // the leader's copies of the barriers both CTAs arrive on
    const uint32_t a_full_leader = mapa_shared(a_full, 0), acc_empty_leader = mapa_shared(acc_empty, 0);

    const int n_tiles = (int)((n + TC_M - 1) / TC_M);
    const int n_steps_all = (n_tiles + 1) / 2;                                  // tile pairs
    const int n_pairs = (int)gridDim.x / 2, pair = (int)blockIdx.x / 2;
    const int steps_per_pair = (n_steps_all + n_pairs - 1) / n_pairs;
    const int step0 = pair * steps_per_pair;
    const int my_steps = step0 < n_steps_all ? (n_steps_all - step0 < steps_per_pair ? n_steps_all - step0 : steps_per_pair) : 0;
    const int total_it = my_steps * TC_NKB;

    if (warp == TC_PRODUCER_WARP) {
        // ---------------- producer: this CTA's half of the query images, then its own row tiles ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_rows)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q)) : "memory");
            mbar_arrive_expect_tx(b_full, 2 * TC_NKB * TCP_B_BYTES);
            for (int kb = 0; kb < TC_NKB; ++kb) {
                tma_2d(base + OFF_BHI + kb * TCP_B_BYTES, &tm_q, kb * TC_KB, (int)rank * TCP_NH, b_full);
                tma_2d(base + OFF_BLO + kb * TCP_B_BYTES, &tm_q, kb * TC_KB, TCP_N + (int)rank * TCP_NH, b_full);
            }
            int it = 0;
            for (int t = 0; t < my_steps; ++t) {
                const int row0 = (2 * (step0 + t) + (int)rank) * TC_M;          // beyond n: the box is zero-filled
                for (int kb = 0; kb < TC_NKB; ++kb, ++it) {
                    const int s = it % TCP_RAW;
                    mbar_wait_guarded(empty_raw + 8 * s, ((it / TCP_RAW) & 1) ^ 1, 0);
                    TC_TRACE(3, it, 0);
                    mbar_arrive_expect_tx(full_raw + 8 * s, TC_A_BYTES);
                    tma_2d(base + OFF_RAW + s * TC_A_BYTES, &tm_rows, kb * TC_KB, row0, full_raw + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp == TC_MMA_WARP) {
        mbar_wait_guarded(b_full, 0, 1);
        if (!leader) {
            // the peer only reports that its query half has landed; the leader issues every MMA of the pair
            fence_proxy_async();                      // TMA (async proxy) wrote the images; the release below publishes them
            if (lane == 0) mbar_arrive_release_cluster(mapa_shared(b_peer, 0));
            __syncwarp();
        } else {
            mbar_wait_acquire_cluster(b_peer, 0);
            tc_fence_after();
            const uint64_t b_hi0 = tc_desc(base + OFF_BHI), b_lo0 = tc_desc(base + OFF_BLO);
            int it = 0;
            for (int t = 0; t < my_steps; ++t) {
                const int buf = t % TCP_NBUF;
                mbar_wait_cluster_guarded(acc_empty + 8 * buf, ((t / TCP_NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * TCP_ACC_COLS;
#pragma unroll
                for (int kb = 0; kb < TC_NKB; ++kb, ++it) {
                    const uint64_t b_hi = b_hi0 + (uint64_t)(kb * (TCP_B_BYTES >> 4)), b_lo = b_lo0 + (uint64_t)(kb * (TCP_B_BYTES >> 4));
                    const int as = it % TCP_A_STAGES;
                    mbar_wait_cluster_guarded(a_full + 8 * as, (it / TCP_A_STAGES) & 1);
                    if (lane == 0) TC_TRACE(1, it, 0);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + TCP_A_COL0 + as * 2 * TC_KB, a_lo = a_hi + TC_KB;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (kb < TC_NKB - 1 || ks < 2) {                    // 300 = 9 * 32 + 12: two k-steps in the last block
                                const int g = kb * 4 + ks;
                                const int cm = g % TCP_MAIN;
                                tcp_mma_ts(acc + cm * TCP_N, a_hi + ks * 8, b_hi + 2 * ks, IDESC, g < TCP_MAIN ? 0u : 1u);
                                tcp_mma_ts(acc + TCP_MAIN * TCP_N, a_hi + ks * 8, b_lo + 2 * ks, IDESC, g < 1 ? 0u : 1u);
                                tcp_mma_ts(acc + TCP_MAIN * TCP_N, a_lo + ks * 8, b_hi + 2 * ks, IDESC, 1u);
                            }
                        }
                        tcp_commit_both(a_empty + 8 * as);
                        if (kb == TC_NKB - 1) tcp_commit_both(acc_full + 8 * buf);
                        TC_TRACE(1, it, 1);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= TC_EPI_WARPS) {
        // ---------------- split warps: box -> registers -> hi | lo in this CTA's TMEM ----------------
        const int g = (warp - TC_EPI_WARPS) >> 2;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + TCP_A_COL0;
        for (int it = g; it < total_it; it += 2) {
            const int s = it % TCP_RAW;
            mbar_wait_guarded(full_raw + 8 * s, (it / TCP_RAW) & 1, 4);
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 0);
            const unsigned char* rowp = gbase + OFF_RAW + s * TC_A_BYTES + r * 128;
            float4 x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float v[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float h = rn_tf32(v[j]);
                    hi[4 * c + j] = __float_as_uint(h);
                    lo[4 * c + j] = __float_as_uint(v[j] - h);
                }
            }
            __syncwarp();
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 1);
            if (lane == 0) mbar_arrive(empty_raw + 8 * s);
            const int as = it % TCP_A_STAGES;
            mbar_wait_cluster_guarded(a_empty + 8 * as, ((it / TCP_A_STAGES) & 1) ^ 1);       // arrives by the leader's commit
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 2);
            tc_fence_after();
            const uint32_t ta = lane_addr + as * 2 * TC_KB;
            tc_st16(ta, reinterpret_cast<const uint32_t(&)[16]>(hi[0]));
            tc_st16(ta + 16, reinterpret_cast<const uint32_t(&)[16]>(hi[16]));
            tc_st16(ta + 32, reinterpret_cast<const uint32_t(&)[16]>(lo[0]));
            tc_st16(ta + 48, reinterpret_cast<const uint32_t(&)[16]>(lo[16]));
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_full_leader + 8 * as);
            if (warp == TC_EPI_WARPS && lane == 0) TC_TRACE(2, it / 2, 3);
        }
    } else {
        // ---------------- epilogue warps: TMEM -> registers -> sim[q][doc] (see scan_tc_kernel) ----------------
        constexpr int QW = TCP_N / 2;
        constexpr int CH = 16;
        const int quarter = warp & 3, half = warp >> 2;
        float lmax[QW];
#pragma unroll
        for (int q = 0; q < QW; ++q) lmax[q] = -INFINITY;
        int n_chunks = (nq_live - half * QW + CH - 1) / CH;
        n_chunks = n_chunks < 0 ? 0 : (n_chunks > QW / CH ? QW / CH : n_chunks);
        for (int t = 0; t < my_steps; ++t) {
            const int buf = t % TCP_NBUF;
            mbar_wait_cluster_guarded(acc_full + 8 * buf, (t / TCP_NBUF) & 1);                // arrives by the leader's commit
            if (warp == 0 && lane == 0) TC_TRACE(0, t, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TCP_ACC_COLS + half * QW;
            const int64_t row = (int64_t)(2 * (step0 + t) + (int)rank) * TC_M + quarter * 32 + lane;
            const bool live = row < n;
            float* orow = out + row + (int64_t)(half * QW) * ld;
            if (n_chunks == 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8 * buf);
            }
#pragma unroll
            for (int c = 0; c < QW / CH; ++c) {
                if (c >= n_chunks) break;
                uint32_t a[TCP_NACC][CH];
#pragma unroll
                for (int m = 0; m < TCP_NACC; ++m) tc_ld16(taddr + m * TCP_N + c * CH, a[m]);
                tc_wait_ld();
                if (c == n_chunks - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8 * buf);
                    if (warp == 0 && lane == 0) TC_TRACE(0, t, 1);
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    float x = __uint_as_float(a[0][j]);
#pragma unroll
                    for (int m = 1; m < TCP_MAIN; ++m) x += __uint_as_float(a[m][j]);
                    const float v = x + __uint_as_float(a[TCP_MAIN][j]);
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

    // nobody leaves while the other CTA may still read this CTA's shared memory / TMEM or arrive on its barriers
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace ais
