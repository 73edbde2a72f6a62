// One RISE bottleneck residual block as ONE kernel (builder_util.py:437-475, _BottlekneckResidualBlock):
//     y = x + BN(conv1x1_{Cop->256}( relu(BN(dw_kxk( relu(BN(conv1x1_{256->Cop}(x))) ))) ))
// One CTA = two boards (128 rows).  The operating channels are processed in chunks of 64:
//     MMA1  D1[128x64]  = X[128x256] . W1_chunk^T        tcgen05, A = resident X tile, B = TMA-streamed weights
//     epi1  relu(D1 + b1) -> H1 (smem, fp16)             8 compute warps, tcgen05.ld
//     dw    depthwise kxk over the two 8x8 boards, + bd, relu -> H2 (smem, written directly in the swizzled
//           K-major layout the tensor core reads)        CUDA cores, fp32 accumulate
//     MMA2  D2[128x256] += H2[128x64] . W2_chunk^T        tcgen05, accumulator stays in TMEM across all chunks
// then   y = D2 + b2 + X (residual from the smem-resident X tile) -> global.
// X is read from HBM/L2 once, the Cop-wide intermediates never leave the SM, and the block is one launch instead of
// three.  TMEM: D1 double-buffered (2 x 64 columns) + D2 (256 columns).
#pragma once
#include "rise_block_args.h"
#include "sm100_prims.cuh"

namespace ara {

constexpr int kRbThreads = 320;  // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2..9 compute
constexpr int kRbX = 65536, kRbW1 = 32768, kRbW2 = 32768, kRbH2 = 16384, kRbH1 = 16384;
constexpr int kRbOffW1 = kRbX, kRbOffW2 = kRbOffW1 + 2 * kRbW1, kRbOffH2 = kRbOffW2 + kRbW2, kRbOffH1 = kRbOffH2 + 2 * kRbH2;
constexpr int kRbOffWd = kRbOffH1 + kRbH1;            // 25 * 64 floats
constexpr int kRbOffBias = kRbOffWd + 25 * 64 * 4;    // b1 chunk, bd chunk: 2 * 64 floats
constexpr int kRbOffBar = kRbOffBias + 2 * 64 * 4;
constexpr int kRbSmemBytes = kRbOffBar + 256 + 1024;

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kRbThreads, 1)
rise_block_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                  const __grid_constant__ CUtensorMap tm_w2, const RiseBlockArgs args) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment by offset arithmetic on the shared array itself: a pointer -> integer -> pointer round trip would
    // make every access below a GENERIC load/store (LD.E / ST.E) instead of LDS / STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem;
    uint8_t* sW1 = smem + kRbOffW1;
    uint8_t* sW2 = smem + kRbOffW2;
    uint8_t* sH2 = smem + kRbOffH2;
    uint8_t* sH1 = smem + kRbOffH1;
    float* sWd = reinterpret_cast<float*>(smem + kRbOffWd);
    float* sBd = reinterpret_cast<float*>(smem + kRbOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRbOffBar);
    uint64_t* x_full = bars + 0;
    uint64_t* w1_full = bars + 1;   // [2]
    uint64_t* w1_empty = bars + 3;  // [2]
    uint64_t* w2_full = bars + 5;
    uint64_t* w2_empty = bars + 6;
    uint64_t* d1_full = bars + 7;    // [2]
    uint64_t* d1_empty = bars + 9;   // [2]
    uint64_t* h2_full = bars + 11;   // [2]
    uint64_t* h2_empty = bars + 13;  // [2]
    uint64_t* d2_full = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x;
    const int nch = args.n_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w1);
        tma_prefetch_desc(&tm_w2);
        mbar_init(x_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 8);
            mbar_init(&h2_full[i], 8);
            mbar_init(&h2_empty[i], 1);
        }
        mbar_init(w2_full, 1);
        mbar_init(w2_empty, 1);
        mbar_init(d2_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(x_full, kRbX);
            for (int p = 0; p < 4; ++p) tma_load_4d(sX + p * 16384, &tm_x, x_full, p * 64, 0, 0, m_tile * 2);
            for (int j = 0; j < nch; ++j) {
                const int s = j & 1;
                mbar_wait(&w1_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&w1_full[s], kRbW1);
                for (int p = 0; p < 4; ++p) tma_load_2d(sW1 + s * kRbW1 + p * 8192, &tm_w1, &w1_full[s], p * 64, j * 64);
                mbar_wait(w2_empty, (j & 1) ^ 1);
                mbar_arrive_expect_tx(w2_full, kRbW2);
                tma_load_2d(sW2, &tm_w2, w2_full, j * 64, 0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc1 = umma_idesc_f16(128, 64, 0);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, 256, 0);
        const uint32_t aX = smem_u32(sX), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aH2 = smem_u32(sH2);
        auto mma2 = [&](int i) {
            const int s = i & 1;
            mbar_wait(&h2_full[s], (i >> 1) & 1);
            mbar_wait(w2_full, i & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_base + 128, umma_desc_k_sw128(aH2 + s * kRbH2 + k * 32, 1024),
                                umma_desc_k_sw128(aW2 + k * 32, 1024), idesc2, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(&h2_empty[s]);
                umma_commit(w2_empty);
            }
            __syncwarp();
        };
        mbar_wait(x_full, 0);
        for (int j = 0; j < nch; ++j) {
            const int s = j & 1;
            mbar_wait(&w1_full[s], (j >> 1) & 1);
            mbar_wait(&d1_empty[s], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + s * 64, umma_desc_k_sw128(aX + kb * 16384 + k * 32, 1024),
                                    umma_desc_k_sw128(aW1 + s * kRbW1 + kb * 8192 + k * 32, 1024), idesc1,
                                    (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit(&w1_empty[s]);
                umma_commit(&d1_full[s]);
            }
            __syncwarp();
            if (j >= 1) mma2(j - 1);
        }
        mma2(nch - 1);
        if (lane == 0) umma_commit(d2_full);
        __syncwarp();
    } else {
        const int cw = warp - 2;       // 0..7
        const int grp = warp & 3;      // TMEM lane group this warp may access
        const int hf = cw >> 2;        // channel half of the 64-wide chunk handled by this thread
        const int r = grp * 32 + lane; // row of the 128-row tile
        const int tid = cw * 32 + lane;
        const int sq = r & 63, brd = r & 64;
        const int y = sq >> 3, x = sq & 7;
        const int K = args.ksize, R = K >> 1, KK = K * K;
        const uint32_t lane_addr = static_cast<uint32_t>(grp * 32) << 16;
        for (int j = 0; j < nch; ++j) {
            const int s = j & 1;
            // ---- epilogue 1: D1 -> relu(+b1) -> H1
            mbar_wait(&d1_full[s], (j >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + s * 64 + hf * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[s]);
            named_bar_sync(1, 256);  // every thread is done with the previous chunk's H1 / depthwise vectors
            // per-chunk vectors: depthwise weights and bias of these 64 channels
            for (int i = tid; i < KK * 64; i += 256) sWd[i] = __ldg(args.wd + (i >> 6) * args.cpad + j * 64 + (i & 63));
            if (tid < 64) sBd[tid] = __ldg(args.bd + j * 64 + tid);
            {
                const float4* b1p = reinterpret_cast<const float4*>(args.b1 + j * 64 + hf * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 ba = __ldg(b1p + q * 2), bb = __ldg(b1p + q * 2 + 1);
                    const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float f0 = fmaxf(__uint_as_float(v[q * 8 + e * 2]) + bias[e * 2], 0.0f);
                        const float f1 = fmaxf(__uint_as_float(v[q * 8 + e * 2 + 1]) + bias[e * 2 + 1], 0.0f);
                        oh[e] = __floats2half2_rn(f0, f1);
                    }
                    const int chunk = (hf * 4 + q) ^ (r & 7);
                    *reinterpret_cast<uint4*>(sH1 + r * 128 + chunk * 16) = o;
                }
            }
            named_bar_sync(2, 256);  // H1 and the per-chunk vectors are complete
            // ---- depthwise k x k on the two boards held in H1
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = sBd[hf * 32 + c];
            for (int dy = -R; dy <= R; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy > 7) continue;
                for (int dx = -R; dx <= R; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx > 7) continue;
                    const int rn = brd | (yy * 8 + xx);
                    const float* wt = sWd + ((dy + R) * K + (dx + R)) * 64 + hf * 32;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int chunk = (hf * 4 + q) ^ (rn & 7);
                        const uint4 hv = *reinterpret_cast<const uint4*>(sH1 + rn * 128 + chunk * 16);
                        const __half2* hh = reinterpret_cast<const __half2*>(&hv);
                        const float4 w0 = *reinterpret_cast<const float4*>(wt + q * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(wt + q * 8 + 4);
                        const float2 a0 = __half22float2(hh[0]), a1 = __half22float2(hh[1]), a2 = __half22float2(hh[2]),
                                     a3 = __half22float2(hh[3]);
                        acc[q * 8 + 0] = fmaf(a0.x, w0.x, acc[q * 8 + 0]);
                        acc[q * 8 + 1] = fmaf(a0.y, w0.y, acc[q * 8 + 1]);
                        acc[q * 8 + 2] = fmaf(a1.x, w0.z, acc[q * 8 + 2]);
                        acc[q * 8 + 3] = fmaf(a1.y, w0.w, acc[q * 8 + 3]);
                        acc[q * 8 + 4] = fmaf(a2.x, w1.x, acc[q * 8 + 4]);
                        acc[q * 8 + 5] = fmaf(a2.y, w1.y, acc[q * 8 + 5]);
                        acc[q * 8 + 6] = fmaf(a3.x, w1.z, acc[q * 8 + 6]);
                        acc[q * 8 + 7] = fmaf(a3.y, w1.w, acc[q * 8 + 7]);
                    }
                }
            }
            // ---- H2 (A operand of MMA2) in the 128B-swizzled K-major layout
            mbar_wait(&h2_empty[s], ((j >> 1) & 1) ^ 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    oh[e] = __floats2half2_rn(fmaxf(acc[q * 8 + e * 2], 0.0f), fmaxf(acc[q * 8 + e * 2 + 1], 0.0f));
                const int chunk = (hf * 4 + q) ^ (r & 7);
                *reinterpret_cast<uint4*>(sH2 + s * kRbH2 + r * 128 + chunk * 16) = o;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_full[s]);
        }
        // ---- final epilogue: y = D2 + b2 + x
        mbar_wait(d2_full, 0);
        tc_fence_after();
        const int m = m_tile * 128 + r;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = hf * 128 + cc * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + 128 + c0, v);
            tmem_ld_wait();
            if (m < args.M) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = c0 + q * 8;
                    const int panel = c >> 6, chunk = ((c & 63) >> 3) ^ (r & 7);
                    const uint4 xv = *reinterpret_cast<const uint4*>(sX + panel * 16384 + r * 128 + chunk * 16);
                    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.b2 + c));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.b2 + c + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xr = __half22float2(xh[e]);
                        oh[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + e * 2]) + bb[e * 2] + xr.x,
                                                  __uint_as_float(v[q * 8 + e * 2 + 1]) + bb[e * 2 + 1] + xr.y);
                    }
                    *reinterpret_cast<uint4*>(args.out + static_cast<size_t>(m) * 256 + c) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
#endif
}

}  // namespace ara
