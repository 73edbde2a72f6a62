// The RISE residual tower for small batches on CTA PAIRS: the transposed kernel of rise_trunk_t.cuh, one board per
// cluster of two CTAs, each streaming HALF of the weights.
//
// rise_trunk_t.cuh is bound by shared-memory bandwidth: with one board per CTA every weight byte is written to shared
// memory by the copy engine and read back by the tensor core for only 64 columns.  Here the two CTAs of a cluster split
// that traffic without changing a single sum:
//   * chunk pair gc belongs to CTA gc & 1: that CTA streams W1 of the pair, runs MMA1 and the depthwise stage, and
//     ends up with H2(gc) in its own shared memory; one bulk copy over distributed shared memory puts the same 16 KB into
//     the partner's H2 buffer;
//   * BOTH CTAs run MMA2 for EVERY pair, in pair order, but each only for its half of the 256 output channels
//     (CTA r: channels 128 r .. 128 r + 127 -- one 32 KB unit of W2 per pair instead of two);
//   * the block epilogue produces the new X for the CTA's own 128 channels (two of the tile's four K panels) and a
//     second bulk copy hands those 16 KB to the partner; the pooled sums of the squeeze-excitation are exchanged as
//     128 floats, the two small FCs run redundantly in both CTAs.
// Per CTA the stream is 2 units per pair instead of 4.  Accumulation orders are those of the one-CTA kernel: results
// are bit-identical (tests/test_net_gpu.py).
// Synchronisation across the pair: tcgen05.commit with .multicast::cluster (an H2 buffer / the X tile may be rewritten
// once BOTH tensor cores are done with it), complete_tx of the DSMEM copies on the receiver's mbarriers, one
// cluster-scope mbarrier for the pooled sums.  H2 buffer b always belongs to CTA b & 1 (four buffers).
// Warp roles: 0, 18, 19 = weight producers, 1 = MMA issuer + TMEM owner, 2..17 = compute, 20 = exchange (DSMEM copies).
#pragma once
#include "rise_trunk_args.h"
#include "rise_trunk_t.cuh"

namespace ara {

constexpr int kRtcProducers = 3;
constexpr int kRtcThreads = (kRttComputeWarps + 2 + kRtcProducers - 1 + 1) * 32;  // + the exchange warp
constexpr int kRtcExchangeWarp = kRttComputeWarps + 2 + kRtcProducers - 1;
constexpr int kRtcRing = 3;
constexpr int kRtcBufs = 4;                                    // H2 buffers (buffer b: written by CTA b & 1)
constexpr int kRtcOffX = 0;                                    // [4 slabs][64 rows][128 B]
constexpr int kRtcOffH2 = kRtcOffX + 32768;                    // [4 buffers][2 slabs][64 rows][128 B]
constexpr int kRtcOffW = kRtcOffH2 + kRtcBufs * 16384;         // [3 slots][32 KB]
constexpr int kRtcOffSe = kRtcOffW + kRtcRing * kTrunkTUnit;   // SE scratch (fp32): part[1024] pool[256] hid[128] scale[256] poolpart[512]
constexpr int kRtcOffB2 = kRtcOffSe + (1024 + 256 + 128 + 256 + 512) * 4;
constexpr int kRtcOffBar = kRtcOffB2 + 2 * 512;
constexpr int kRtcSmemBytes = kRtcOffBar + 512 + 1024;
static_assert(kRtcSmemBytes <= 232448, "cluster trunk kernel shared memory exceeds the sm_100 limit");
constexpr uint32_t kRtcColD1 = 0, kRtcColD2 = 128;
constexpr int kRtcTmemCols = 256;

__device__ __forceinline__ uint32_t rtc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// arrives on the mbarrier at this offset in every CTA of `mask` once the previously issued tcgen05.mma have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// shared memory of this CTA -> shared memory of another CTA of the cluster; completes `bytes` on the receiver's mbarrier
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kRtcThreads, 1) rise_trunk_c_kernel(const __grid_constant__ TrunkArgs args) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + kRtcOffX;
    uint8_t* sH2 = smem + kRtcOffH2;
    uint8_t* sW = smem + kRtcOffW;
    float* sPart = reinterpret_cast<float*>(smem + kRtcOffSe);
    float* sPool = sPart + 1024;
    float* sHid = sPool + 256;
    float* sScale = sHid + 128;
    float* sPoolPart = sScale + 256;  // [4 groups of 16 squares][128 own channels]
    float* sB2all = reinterpret_cast<float*>(smem + kRtcOffB2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRtcOffBar);
    uint64_t* x_ready = bars + 0;       // 2: the exchange warp (own panels stored) + the armed expect_tx (partner's panels)
    uint64_t* x_local = bars + 1;       // 16 compute warps: own panels of the tile stored
    uint64_t* blk_done = bars + 2;      // 2 commits (both CTAs): every MMA of the block has completed
    uint64_t* pool_bar = bars + 3;      // 8 warps (4 here, 4 in the partner): pooled sums of all 256 channels present
    uint64_t* w_full = bars + 4;        // [3]
    uint64_t* w_empty = bars + 7;       // [3]
    uint64_t* d1_full = bars + 10;      // [2]
    uint64_t* d1_empty = bars + 12;     // [2]  8 warps
    uint64_t* h2_written = bars + 14;   // [4]  8 warps (own buffers)
    uint64_t* h2_full = bars + 18;      // [4]  1: the exchange warp (own buffers) / the armed expect_tx (partner's)
    uint64_t* h2_free = bars + 22;      // [4]  2 commits (both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = rtc_cluster_rank(), partner = rank ^ 1u;
    const int board = blockIdx.x >> 1;
    const int n_blocks = args.n_blocks;
    // (both CTAs of a cluster take the same decision)
    if (args.boards_dev != nullptr && board >= *args.boards_dev) return;

    if (warp == 0 && lane == 0) {
        mbar_init(x_ready, 2);
        mbar_init(x_local, kRttComputeWarps);
        mbar_init(blk_done, 2);
        mbar_init(pool_bar, 8);
        for (int i = 0; i < kRtcRing; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 8);
        }
        for (int i = 0; i < kRtcBufs; ++i) {
            mbar_init(&h2_written[i], 8);
            mbar_init(&h2_full[i], 1);
            mbar_init(&h2_free[i], 2);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kRtcTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();  // the partner's barriers exist before anything arrives on them
    pdl_wait();

    if (warp == 0 || (warp >= kRttComputeWarps + 2 && warp < kRtcExchangeWarp)) {
        // ---------------------------------------------------------------- producers: unit u of this CTA's stream: producer u % 3
        if (lane == 0) {
            const int n_units = args.c_units[rank];
            const int* seq = args.c_seq[rank];
            for (int u = warp == 0 ? 0 : warp - (kRttComputeWarps + 1); u < n_units; u += kRtcProducers) {
                const uint32_t s = static_cast<uint32_t>(u) % kRtcRing;
                mbar_wait_relaxed(&w_empty[s], ((static_cast<uint32_t>(u) / kRtcRing) & 1) ^ 1);
                mbar_arrive_expect_tx(&w_full[s], kTrunkTUnit);
                bulk_load_1d(sW + s * kTrunkTUnit, args.t_img + static_cast<size_t>(__ldg(seq + u)) * kTrunkTUnit, kTrunkTUnit, &w_full[s]);
            }
        }
    } else if (warp == kRtcExchangeWarp) {
        // ---------------------------------------------------------------- exchange: this CTA's H2 buffers and X panels to the partner
        if (lane == 0) {
            for (int b = 0; b < n_blocks; ++b) {
                const TrunkBlock& B = args.blk[b];
                const int P = (B.n_chunks + 1) >> 1;
                // the tile of block b: own panels (2 rank, 2 rank + 1) are stored, the partner gets a copy
                mbar_wait(x_local, b & 1);
                mbar_arrive(x_ready);
                bulk_copy_to_cta(cluster_map(sX + rank * 16384, partner), sX + rank * 16384, 16384, cluster_map(x_ready, partner));
                for (int i = 0; i < P; ++i) {
                    const uint32_t gc = static_cast<uint32_t>(B.pair0 + i);
                    if ((gc & 1u) != rank) continue;
                    const uint32_t buf = gc % kRtcBufs, k = gc / kRtcBufs;
                    mbar_wait(&h2_written[buf], k & 1);
                    mbar_arrive(&h2_full[buf]);
                    bulk_copy_to_cta(cluster_map(sH2 + buf * 16384, partner), sH2 + buf * 16384, 16384, cluster_map(&h2_full[buf], partner));
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0);
        const uint32_t aX = smem_u32(sX), aH2 = smem_u32(sH2), aW = smem_u32(sW);
        uint32_t useq = 0, n_own = 0;
        RT_PROF_DECL();
        if (lane == 0) {  // what the partner will send: its X panels of the first tile, its H2 buffers
            mbar_arrive_expect_tx(x_ready, 16384);
            for (uint32_t bf = 0; bf < kRtcBufs; ++bf)
                if ((bf & 1u) != rank) mbar_arrive_expect_tx(&h2_full[bf], 16384);
        }
        __syncwarp();
        auto next_unit = [&]() -> uint32_t {
            const uint32_t s = useq % kRtcRing;
            mbar_wait(&w_full[s], (useq / kRtcRing) & 1);
            tc_fence_after();
            return aW + s * kTrunkTUnit;
        };
        auto release_unit = [&]() {
            if (lane == 0) umma_commit(&w_empty[useq % kRtcRing]);
            __syncwarp();
            ++useq;
        };
        for (int b = 0; b < n_blocks; ++b) {
            const TrunkBlock& B = args.blk[b];
            const int P = (B.n_chunks + 1) >> 1;
            const bool odd = (B.n_chunks & 1) != 0;
            if (B.se_type != 0) useq += 4;  // the block's squeeze-excitation units: consumed by the compute warps
            RT_PROF(0);
            mbar_wait(x_ready, b & 1);
            if (lane == 0 && b + 1 < n_blocks) mbar_arrive_expect_tx(x_ready, 16384);  // the partner's panels of the next tile
            __syncwarp();
            RT_PROF(1);  // wait for the X tile (block boundary)
            tc_fence_after();
            auto mma2 = [&](int i) {
                const uint32_t gc = static_cast<uint32_t>(B.pair0 + i), buf = gc % kRtcBufs, k = gc / kRtcBufs;
                RT_PROF(0);
                mbar_wait(&h2_full[buf], k & 1);
                if (lane == 0 && (buf & 1u) != rank) mbar_arrive_expect_tx(&h2_full[buf], 16384);  // the buffer's next use
                __syncwarp();
                RT_PROF(2);  // wait for H2 (compute warps / the partner)
                const int slabs = (odd && i == P - 1) ? 1 : 2;
                const uint32_t a = next_unit();
                RT_PROF(3);  // wait for the weight stream
                if (lane == 0) {
                    for (int s = 0; s < slabs; ++s)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_f16_ss(tmem_base + kRtcColD2, umma_desc_k_sw128(a + s * 16384 + kk * 32, 1024),
                                        umma_desc_k_sw128(aH2 + buf * 16384 + s * 8192 + kk * 32, 1024), idesc,
                                        (i == 0 && s == 0 && kk == 0) ? 0u : 1u);
                }
                release_unit();
                if (lane == 0) umma_commit_mc(&h2_free[buf], 3);
                __syncwarp();
            };
            for (int i = 0; i < P; ++i) {
                const uint32_t gc = static_cast<uint32_t>(B.pair0 + i);
                if ((gc & 1u) == rank) {
                    const uint32_t g = n_own & 1, n = n_own >> 1;
                    RT_PROF(0);
                    mbar_wait(&d1_empty[g], (n & 1) ^ 1);
                    RT_PROF(4);  // wait for a free D1 accumulator
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t a = next_unit();
                        RT_PROF(3);
                        if (lane == 0) {
#pragma unroll
                            for (int s = 0; s < 2; ++s)
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16_ss(tmem_base + kRtcColD1 + g * 64, umma_desc_k_sw128(a + s * 16384 + kk * 32, 1024),
                                                umma_desc_k_sw128(aX + (u * 2 + s) * 8192 + kk * 32, 1024), idesc,
                                                (u == 0 && s == 0 && kk == 0) ? 0u : 1u);
                        }
                        release_unit();
                    }
                    if (lane == 0) umma_commit(&d1_full[g]);
                    __syncwarp();
                    ++n_own;
                }
                if (i >= kTrunkCLag) mma2(i - kTrunkCLag);
            }
            for (int i = P > kTrunkCLag ? P - kTrunkCLag : 0; i < P; ++i) mma2(i);
            if (lane == 0) umma_commit_mc(blk_done, 3);
            __syncwarp();
            // the kernels behind this one may be scheduled now (programmatic dependent launch): triggered late -- by this
            // one thread, at the last block -- so that their thread blocks do not sit on SMs while the tower still runs
            if (lane == 0 && b == n_blocks - 1) pdl_launch_dependents();
        }
        RT_PROF(0);
        RT_PROF_FLUSH(0);
    } else {
        // ---------------------------------------------------------------- compute warps
        const int cw = warp - 2;   // 0..15
        const int q = warp & 3;    // TMEM lane quadrant this warp may access
        const int g4 = cw >> 2;    // epilogue role: squares 16 g4 .. 16 g4 + 15 of own channel 32 q + lane
        const int dhalf = g4 & 1;  // depthwise role: output rows 4 dhalf .. 4 dhalf + 3; group cw >> 3
        const int dgrp = cw >> 3;
        const int tid = cw * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int eo = q * 32 + lane;                       // own channel index 0..127
        const int ec = static_cast<int>(rank) * 128 + eo;   // channel of the tile
        const int esq0 = g4 * 16;
        uint8_t* ex = sX + (ec >> 6) * 8192 + esq0 * 128 + (ec & 7) * 2;
        const uint32_t ec3s = static_cast<uint32_t>(((ec & 63) >> 3) << 4);
        const int dch = q * 32 + lane;                      // depthwise role: channel of the pair
        const uint32_t dc3s = static_cast<uint32_t>(((dch & 63) >> 3) << 4);
        uint32_t n_own = 0, n_se = 0;
        RT_PROF_DECL();

        // Squeeze-excitation of block `Bn` on this thread's 16 tile values (own channel): returns the channel's scale.
        // Pooling order of rise_trunk.cuh: 16-square sums in butterfly order, ((s0 + s1) + (s2 + s3)).
        auto se_scale = [&](const TrunkBlock& Bn, const uint32_t (&xp)[8]) -> float {
            const uint32_t u0 = static_cast<uint32_t>(Bn.se_seq0c[rank]);
            auto unit_ptr = [&](uint32_t j) -> const uint8_t* { return sW + ((u0 + j) % kRtcRing) * kTrunkTUnit; };
            auto unit_wait = [&](uint32_t j) { mbar_wait(&w_full[(u0 + j) % kRtcRing], ((u0 + j) / kRtcRing) & 1); };
            auto unit_free = [&](uint32_t j) { mbar_arrive(&w_empty[(u0 + j) % kRtcRing]); };
            {
                float v[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 a = rt_unpack(xp[i]);
                    v[2 * i] = a.x, v[2 * i + 1] = a.y;
                }
                sPoolPart[g4 * 128 + eo] = rtt_tree16(v);
            }
            rtt_bar_sync(2);
            if (tid < 128) {  // own channels: to both CTAs
                const float pooled = ((sPoolPart[tid] + sPoolPart[128 + tid]) + (sPoolPart[256 + tid] + sPoolPart[384 + tid])) * (1.0f / 64.0f);
                sPool[rank * 128 + tid] = pooled;
                st_cluster_f32(cluster_map(sPool + rank * 128 + tid, partner), pooled);
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(cluster_map(pool_bar, rank));
                    mbar_arrive_cluster(cluster_map(pool_bar, partner));
                }
            }
            mbar_wait_cluster(pool_bar, n_se & 1);
            ++n_se;
            RT_PROF(10);  // SE: pooling
            if (Bn.se_type == 1) {
                {   // fc1 (256 -> 128): 8 K-groups of 32 x 64 output pairs; matrix [256][128] fp16 = units 0, 1
                    const int kg = tid >> 6, jp = tid & 63;
                    unit_wait(kg >> 2);
                    const __half2* w = reinterpret_cast<const __half2*>(unit_ptr(kg >> 2) + ((kg & 3) * 32) * 256) + jp;
                    const float* p0 = sPool + kg * 32;
                    float a0 = 0.0f, c0 = 0.0f;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float2 wf = __half22float2(w[k * 64]);
                        a0 = fmaf(wf.x, p0[k], a0);
                        c0 = fmaf(wf.y, p0[k], c0);
                    }
                    sPart[kg * 128 + 2 * jp] = a0;
                    sPart[kg * 128 + 2 * jp + 1] = c0;
                }
                RT_PROF(11);  // SE: fc1
                rtt_bar_sync(2);
                if (tid == 0) unit_free(0), unit_free(1);
                if (tid < 128) {
                    const float* qq = sPart + tid;
                    sHid[tid] = fmaxf(((qq[0] + qq[128]) + (qq[256] + qq[384])) + ((qq[512] + qq[640]) + (qq[768] + qq[896])), 0.0f);
                }
                rtt_bar_sync(1);
                RT_PROF(12);  // SE: hidden layer
                {   // fc2 (128 -> 256): 4 K-groups of 32 x 128 output pairs; matrix [128][256] fp16 = units 2, 3
                    const int kg = tid >> 7, cp = tid & 127;
                    unit_wait(2 + (kg >> 1));
                    const __half2* w = reinterpret_cast<const __half2*>(unit_ptr(2 + (kg >> 1)) + ((kg & 1) * 32) * 512) + cp;
                    const float* h0 = sHid + kg * 32;
                    float a0 = 0.0f, c0 = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 wf = __half22float2(w[j * 128]);
                        a0 = fmaf(wf.x, h0[j], a0);
                        c0 = fmaf(wf.y, h0[j], c0);
                    }
                    sPart[kg * 256 + 2 * cp] = a0;
                    sPart[kg * 256 + 2 * cp + 1] = c0;
                }
                RT_PROF(13);  // SE: fc2
                rtt_bar_sync(2);
                if (tid == 0) unit_free(2), unit_free(3);
                if (tid < 256) {
                    const float* qq = sPart + tid;
                    sScale[tid] = rt_hard_sigmoid((qq[0] + qq[256]) + (qq[512] + qq[768]));
                }
            } else {
                {   // 256 -> 256: 4 K-groups of 64 x 128 output pairs; matrix [256][256] fp16 = units 0 .. 3
                    const int kg = tid >> 7, cp = tid & 127;
                    unit_wait(kg);
                    const __half2* w = reinterpret_cast<const __half2*>(unit_ptr(kg)) + cp;
                    const float* p0 = sPool + kg * 64;
                    float a0 = 0.0f, c0 = 0.0f;
#pragma unroll 32
                    for (int k = 0; k < 64; ++k) {
                        const float2 wf = __half22float2(w[k * 128]);
                        a0 = fmaf(wf.x, p0[k], a0);
                        c0 = fmaf(wf.y, p0[k], c0);
                    }
                    sPart[kg * 256 + 2 * cp] = a0;
                    sPart[kg * 256 + 2 * cp + 1] = c0;
                    // the ring has three slots: a quarter's unit is released as soon as its 128 threads are done with it
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + kg) : "memory");
                    if (cp == 0) unit_free(kg);
                }
                rtt_bar_sync(2);
                if (tid < 256) {
                    const float* qq = sPart + tid;
                    sScale[tid] = rt_hard_sigmoid(__ldg(Bn.se_b + tid) + ((qq[0] + qq[256]) + (qq[512] + qq[768])));
                }
            }
            rtt_bar_sync(1);
            return sScale[ec];
        };
        // this thread's 16 tile values: from / to the X tile (element i = square esq0 + i of channel ec)
        auto load_tile = [&](uint32_t (&xp)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t lo = *reinterpret_cast<const uint16_t*>(ex + (2 * i) * 128 + (ec3s ^ (((2 * i) & 7) << 4)));
                const uint32_t hi = *reinterpret_cast<const uint16_t*>(ex + (2 * i + 1) * 128 + (ec3s ^ (((2 * i + 1) & 7) << 4)));
                xp[i] = lo | (hi << 16);
            }
        };
        auto store_tile = [&](const uint32_t (&xp)[8], bool scaled, float sc) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t v = xp[i];
                if (scaled) {
                    const float2 f = rt_unpack(v);
                    v = rt_pack(f.x * sc, f.y * sc);
                }
                *reinterpret_cast<uint16_t*>(ex + (2 * i) * 128 + (ec3s ^ (((2 * i) & 7) << 4))) = static_cast<uint16_t>(v & 0xffffu);
                *reinterpret_cast<uint16_t*>(ex + (2 * i + 1) * 128 + (ec3s ^ (((2 * i + 1) & 7) << 4))) = static_cast<uint16_t>(v >> 16);
            }
        };
        auto hand_over = [&]() {  // own panels stored: the exchange warp passes them on
            rt_fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_local);
        };

        {   // stem output, own panels -> the X tile (16-byte pieces into the swizzled K-major layout)
            const uint4* src = reinterpret_cast<const uint4*>(args.x_in + static_cast<size_t>(board) * 64 * 256);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int p = tid + i * 512;        // piece: row p / 16, 16-byte column p % 16 of the own half row
                const int r = p >> 4, c16 = p & 15;
                const uint4 v = __ldg(src + r * 32 + rank * 16 + c16);
                *reinterpret_cast<uint4*>(sX + (rank * 2 + (c16 >> 3)) * 8192 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = v;
            }
            rtt_bar_sync(1);  // from here on a thread only touches its own elements of the tile (channel ec, 16 squares)
            if (args.blk[0].se_type != 0) {  // squeeze-excitation on the tower input
                uint32_t xp[8];
                load_tile(xp);
                const float sc = se_scale(args.blk[0], xp);
                store_tile(xp, true, sc);
            }
            hand_over();
        }
        RT_PROF(0);  // X load (+ SE of the first block)
        for (int b = 0; b < n_blocks; ++b) {
            const TrunkBlock& B = args.blk[b];
            const int P = (B.n_chunks + 1) >> 1;
            const bool odd = (B.n_chunks & 1) != 0;
            const bool last = b == n_blocks - 1;
            float* sB2 = sB2all + (b & 1) * 128;
            if (tid < 128) sB2[tid] = __ldg(B.b2 + rank * 128 + tid);
            for (int i = 0; i < P; ++i) {
                const uint32_t gc = static_cast<uint32_t>(B.pair0 + i);
                if ((gc & 1u) != rank) continue;
                const uint32_t own = n_own++;
                if (static_cast<int>(own & 1) != dgrp) continue;
                const uint32_t n = own >> 1, buf = gc % kRtcBufs, k = gc / kRtcBufs;
                const bool idle = odd && i == P - 1 && q >= 2;  // the padded half of an odd last pair: nothing to compute
                const uint8_t* aux = args.t_aux + static_cast<size_t>(gc) * kTrunkTAux;
                float b1 = 0.0f, bd = 0.0f;
                uint32_t wp[13];
                if (!idle) {  // the channel's vectors: in flight while the tensor core works on D1
                    b1 = __ldg(reinterpret_cast<const float*>(aux) + dch);
                    bd = __ldg(reinterpret_cast<const float*>(aux + 512) + dch);
                    const uint16_t* wd = reinterpret_cast<const uint16_t*>(aux + 1024) + dch;  // [k*k][128]
                    const int kk = B.ksize * B.ksize;
#pragma unroll
                    for (int j = 0; j < 13; ++j) {
                        const uint32_t lo = 2 * j < kk ? __ldg(wd + (2 * j) * 128) : 0u;
                        const uint32_t hi = 2 * j + 1 < kk ? __ldg(wd + (2 * j + 1) * 128) : 0u;
                        wp[j] = lo | (hi << 16);
                    }
                }
                mbar_wait(&d1_full[dgrp], n & 1);
                RT_PROF(2);  // wait for D1 (tensor core)
                tc_fence_after();
                uint32_t h1[32];
                if (!idle) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tmem_base + lane_addr + kRtcColD1 + dgrp * 64 + hh * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            h1[hh * 16 + j] = rt_pack(fmaxf(__uint_as_float(v[2 * j]) + b1, 0.0f), fmaxf(__uint_as_float(v[2 * j + 1]) + b1, 0.0f));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d1_empty[dgrp]);
                RT_PROF(3);  // TMEM read-out, bias, relu
                mbar_wait(&h2_free[buf], (k & 1) ^ 1);  // both tensor cores are done with the buffer's previous content
                RT_PROF(4);  // wait for the H2 buffer
                if (!idle) {
                    uint8_t* h2 = sH2 + buf * 16384 + (dch >> 6) * 8192 + (dch & 7) * 2;
                    if (B.ksize == 3) {
                        if (dhalf == 0) rtt_depthwise<3, 0>(h1, wp, bd, h2, dc3s);
                        else rtt_depthwise<3, 4>(h1, wp, bd, h2, dc3s);
                    } else {
                        if (dhalf == 0) rtt_depthwise<5, 0>(h1, wp, bd, h2, dc3s);
                        else rtt_depthwise<5, 4>(h1, wp, bd, h2, dc3s);
                    }
                }
                rt_fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h2_written[buf]);
                RT_PROF(5);  // depthwise + H2 write
            }
            // ---- block epilogue, own channels: X <- (D2 + b2) + X (the last block: to global memory), then the next block's SE
            uint32_t xo[8];
            load_tile(xo);  // the block input (this thread's own elements): the residual
            mbar_wait(blk_done, b & 1);  // both CTAs: every MMA of the block is complete (the partner's tile may be rewritten too)
            RT_PROF(6);  // wait for D2
            tc_fence_after();
            rtt_bar_sync(1);  // b2 of this block is visible
            uint32_t xp[8];
            {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + lane_addr + kRtcColD2 + esq0, v);
                tmem_ld_wait();
                const float b2 = sB2[eo];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 r = rt_unpack(xo[i]);
                    xp[i] = rt_pack((__uint_as_float(v[2 * i]) + b2) + r.x, (__uint_as_float(v[2 * i + 1]) + b2) + r.y);
                }
            }
            tc_fence_before();
            RT_PROF(7);  // block epilogue: D2 + b2 + X
            if (!last && args.blk[b + 1].se_type != 0) {
                const float sc = se_scale(args.blk[b + 1], xp);
                RT_PROF(8);  // squeeze-excitation of the next block
                store_tile(xp, true, sc);
            } else {
                store_tile(xp, false, 1.0f);
            }
            if (last) {
                rtt_bar_sync(1);
                uint4* dst = reinterpret_cast<uint4*>(args.out + static_cast<size_t>(board) * 64 * 256);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int p = tid + i * 512;
                    const int r = p >> 4, c16 = p & 15;
                    dst[r * 32 + rank * 16 + c16] =
                        *reinterpret_cast<const uint4*>(sX + (rank * 2 + (c16 >> 3)) * 8192 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
                }
            } else {
                hand_over();
            }
            RT_PROF(9);  // tile store + hand-over
        }
        if (warp == 2) RT_PROF_FLUSH(1);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves while its partner may still copy into its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kRtcTmemCols>(tmem_base);
    }
#endif
}

}  // namespace ara
