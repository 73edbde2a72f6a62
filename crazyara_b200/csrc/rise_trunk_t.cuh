// The RISE residual tower for SMALL batches: one board per CTA, channels in the tensor core's M dimension.
//
// rise_trunk.cuh puts the board's 64 squares into M: with one board per CTA that is UMMA M=64 (half rate), and its CUDA
// core stages (depthwise through shared memory, two block-wide barriers per chunk) bound the kernel.  This kernel
// transposes the problem -- same arithmetic, same summation orders, bit-identical results (tests/test_net_gpu.py):
//
//     MMA1  D1^T[128 op-channels x 64 squares] = W1_pair[128 x 256] . X^T      A = weights (smem), B = X tile (smem),
//                                                                               M = 128: full rate
//     dw    one THREAD owns one operating channel: its 8x8 plane arrives as the 64 columns of its TMEM lane, the
//           depthwise k x k runs entirely in registers (FHFMA, fp32 accumulate, tap order of rt_depthwise), the result
//           goes to H2[square][channel] (128B-swizzled K-major, the B operand of MMA2).  No H1 in shared memory, no
//           block-wide barrier in the chunk loop.
//     MMA2  D2^T[256 x 64] += W2_pair[256 x 128] . H2^T                        two M = 128 halves, accumulated in TMEM
//                                                                               over the pairs of the block
//     block epilogue: thread = (output channel, 32 squares): X <- (D2 + b2) + X in the shared-memory X tile
//     squeeze-excitation: the pooled sum of a channel is thread-local (same summation tree as the butterfly of
//     rise_trunk.cuh), the two small FCs as there.
//
// Operating channels are processed in PAIRS of 64-channel chunks (an odd last chunk is padded with zero weights).
// Two groups of eight compute warps (two per TMEM lane quadrant: a channel's output rows 0-3 and 4-7) each own a D1
// accumulator and an H2 buffer and work on alternate pairs; MMA2 of a pair is issued kTrunkTLag pairs behind its MMA1, which is the order in which the
// weights stream: 32 KB units (one bulk copy each, three issuing threads -- the per-SM copy engine moves >100 B/clk in
// that regime, tools/micro/l2_ingest.cu) through a 4-slot ring.  With one board per CTA the tensor core re-reads the
// weights from shared memory for only 64 columns: the kernel is bound by shared-memory bandwidth (copy-engine writes +
// operand reads, ~350 KB per pair), not by the tensor pipe.
// Warp roles: 0, 18, 19 = weight producers (units by sequence number modulo 3), 1 = MMA issuer + TMEM owner,
// 2..17 = compute.  TMEM columns: D1 2 x 64, D2 2 x 64.
#pragma once
#include "rise_trunk_args.h"
#include "rise_trunk.cuh"

namespace ara {

constexpr int kRttGroups = 2;         // depthwise groups (8 warps each: two per TMEM lane quadrant)
constexpr int kRttComputeWarps = 16;
#if !defined(ARA_RTT_PRODUCERS)
#define ARA_RTT_PRODUCERS 3
#endif
constexpr int kRttProducers = ARA_RTT_PRODUCERS;
constexpr int kRttThreads = (kRttComputeWarps + 2 + kRttProducers - 1) * 32;  // warp 0 + warps 18.. = producers, warp 1 = MMA
constexpr int kRttRing = 4;
constexpr int kRttOffX = 0;                                   // [4 slabs][64 rows][128 B]
constexpr int kRttOffH2 = kRttOffX + 32768;                   // [2 groups][2 slabs][64 rows][128 B]
constexpr int kRttOffW = kRttOffH2 + kRttGroups * 16384;      // [4 slots][32 KB]
constexpr int kRttOffSe = kRttOffW + kRttRing * kTrunkTUnit;  // SE scratch (fp32): part[1024] pool[256] hid[128] scale[256] poolpart[512]
constexpr int kRttOffB2 = kRttOffSe + (1024 + 256 + 128 + 256 + 512) * 4;
constexpr int kRttOffBar = kRttOffB2 + 2 * 1024;
constexpr int kRttSmemBytes = kRttOffBar + 512 + 1024;
static_assert(kRttSmemBytes <= 232448, "transposed trunk kernel shared memory exceeds the sm_100 limit");
constexpr uint32_t kRttColD1 = 0, kRttColD2 = 128;
constexpr int kRttTmemCols = 512;

__device__ __forceinline__ void rtt_bar_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kRttComputeWarps * 32) : "memory");
}

// acc += x.half[HX] * w.half[HW] (one FHFMA; the half selectors are compile-time)
template <int HX, int HW>
__device__ __forceinline__ void rtt_fhfma(float& acc, uint32_t x, uint32_t w) {
    if (HX == 0 && HW == 0)
        asm("{\n\t.reg .f16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %1;\n\tmov.b32 {wl, wh}, %2;\n\tfma.rn.f32.f16 %0, xl, wl, %0;\n\t}" : "+f"(acc) : "r"(x), "r"(w));
    else if (HX == 0 && HW == 1)
        asm("{\n\t.reg .f16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %1;\n\tmov.b32 {wl, wh}, %2;\n\tfma.rn.f32.f16 %0, xl, wh, %0;\n\t}" : "+f"(acc) : "r"(x), "r"(w));
    else if (HX == 1 && HW == 0)
        asm("{\n\t.reg .f16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %1;\n\tmov.b32 {wl, wh}, %2;\n\tfma.rn.f32.f16 %0, xh, wl, %0;\n\t}" : "+f"(acc) : "r"(x), "r"(w));
    else
        asm("{\n\t.reg .f16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %1;\n\tmov.b32 {wl, wh}, %2;\n\tfma.rn.f32.f16 %0, xh, wh, %0;\n\t}" : "+f"(acc) : "r"(x), "r"(w));
}
// acc += x.half[sel >> 1] * w.half[sel & 1] (one FHFMA; `sel` is a compile-time constant after unrolling)
__device__ __forceinline__ void rtt_fhfma_sel(float& acc, uint32_t x, uint32_t w, int sel) {
    switch (sel) {
        case 0: rtt_fhfma<0, 0>(acc, x, w); break;
        case 1: rtt_fhfma<0, 1>(acc, x, w); break;
        case 2: rtt_fhfma<1, 0>(acc, x, w); break;
        default: rtt_fhfma<1, 1>(acc, x, w); break;
    }
}
// depthwise k x k of one channel's 8x8 plane (h1: 32 packed pairs of horizontally adjacent squares; wp: the k*k taps,
// row-major, two per register), relu, fp16, to H2[square][channel].  Per output the taps are added in rt_depthwise's
// order (column offset outer, row offset inner); the eight outputs of a row are independent chains, interleaved.
template <int K, int Y0>
__device__ __forceinline__ void rtt_depthwise(const uint32_t (&h1)[32], const uint32_t (&wp)[13], float bd, uint8_t* h2, uint32_t c3s) {
    constexpr int R = K / 2;
#pragma unroll
    for (int y = Y0; y < Y0 + 4; ++y) {  // (the channel's other four rows: the partner warp)
        float acc[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) acc[x] = bd;
#pragma unroll
        for (int dxi = 0; dxi < K; ++dxi) {
#pragma unroll
            for (int dyi = 0; dyi < K; ++dyi) {
                const int yy = y + dyi - R;
                if (yy < 0 || yy > 7) continue;
                const int t = dyi * K + dxi;  // weight index: row-major k x k
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const int xx = x + dxi - R;
                    if (xx < 0 || xx > 7) continue;
                    const int sq = yy * 8 + xx;
                    rtt_fhfma_sel(acc[x], h1[sq >> 1], wp[t >> 1], (sq & 1) * 2 + (t & 1));
                }
            }
        }
#pragma unroll
        for (int x = 0; x < 8; ++x)
            *reinterpret_cast<__half*>(h2 + (y * 8 + x) * 128 + (c3s ^ (x << 4))) = __float2half_rn(fmaxf(acc[x], 0.0f));
    }
}

// sum of 16 values in the order of the lane butterfly of rise_trunk.cuh (partners 8, 4, 2, 1 apart)
__device__ __forceinline__ float rtt_tree16(const float (&x)[16]) {
    float a[8], b[4], c[2];
#pragma unroll
    for (int l = 0; l < 8; ++l) a[l] = x[l] + x[l + 8];
#pragma unroll
    for (int l = 0; l < 4; ++l) b[l] = a[l] + a[l + 4];
#pragma unroll
    for (int l = 0; l < 2; ++l) c[l] = b[l] + b[l + 2];
    return c[0] + c[1];
}

__global__ void __launch_bounds__(kRttThreads, 1) rise_trunk_t_kernel(const __grid_constant__ TrunkArgs args) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + kRttOffX;
    uint8_t* sH2 = smem + kRttOffH2;
    uint8_t* sW = smem + kRttOffW;
    float* sPart = reinterpret_cast<float*>(smem + kRttOffSe);
    float* sPool = sPart + 1024;
    float* sHid = sPool + 256;
    float* sScale = sHid + 128;
    float* sPoolPart = sScale + 256;  // [2 square halves][256]
    float* sB2all = reinterpret_cast<float*>(smem + kRttOffB2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRttOffBar);
    uint64_t* x_ready = bars + 0;
    uint64_t* d2_full = bars + 1;
    uint64_t* w_full = bars + 2;     // [4]
    uint64_t* w_empty = bars + 6;    // [4]
    uint64_t* d1_full = bars + 10;   // [2]
    uint64_t* d1_empty = bars + 12;  // [2]
    uint64_t* h2_full = bars + 14;   // [2]
    uint64_t* h2_empty = bars + 16;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int board = blockIdx.x;
    const int n_blocks = args.n_blocks;
    if (args.boards_dev != nullptr && board >= *args.boards_dev) return;

    if (warp == 0 && lane == 0) {
        mbar_init(x_ready, kRttComputeWarps);
        mbar_init(d2_full, 1);
        for (int i = 0; i < kRttRing; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < kRttGroups; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 8);
            mbar_init(&h2_full[i], 8);
            mbar_init(&h2_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kRttTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0 || warp >= kRttComputeWarps + 2) {
        // ---------------------------------------------------------------- producers: unit u belongs to producer u % 3
        if (lane == 0) {
            const int n_units = args.t_units;
            for (int u = warp == 0 ? 0 : warp - (kRttComputeWarps + 1); u < n_units; u += kRttProducers) {
                const uint32_t s = static_cast<uint32_t>(u) % kRttRing;
                mbar_wait_relaxed(&w_empty[s], ((static_cast<uint32_t>(u) / kRttRing) & 1) ^ 1);
                mbar_arrive_expect_tx(&w_full[s], kTrunkTUnit);
                bulk_load_1d(sW + s * kTrunkTUnit, args.t_img + static_cast<size_t>(__ldg(args.t_seq + u)) * kTrunkTUnit, kTrunkTUnit,
                             &w_full[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0);
        const uint32_t aX = smem_u32(sX), aH2 = smem_u32(sH2), aW = smem_u32(sW);
        uint32_t useq = 0;
        RT_PROF_DECL();
        auto next_unit = [&]() -> uint32_t {  // waits for the next unit of the stream, returns its shared-memory address
            const uint32_t s = useq % kRttRing;
            mbar_wait(&w_full[s], (useq / kRttRing) & 1);
            tc_fence_after();
            return aW + s * kTrunkTUnit;
        };
        auto release_unit = [&]() {
            if (lane == 0) umma_commit(&w_empty[useq % kRttRing]);
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
            RT_PROF(1);  // wait for the X tile (block boundary)
            tc_fence_after();
            auto mma2 = [&](int i) {
                const uint32_t gc = static_cast<uint32_t>(B.pair0 + i), g = gc % kRttGroups, n = gc / kRttGroups;
                RT_PROF(0);
                mbar_wait(&h2_full[g], n & 1);
                RT_PROF(2);  // wait for H2 (compute warps)
                const int slabs = (odd && i == P - 1) ? 1 : 2;
                for (int h = 0; h < 2; ++h) {
                    const uint32_t a = next_unit();
                    RT_PROF(3);  // wait for the weight stream
                    if (lane == 0) {
                        for (int s = 0; s < slabs; ++s)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ss(tmem_base + kRttColD2 + h * 64, umma_desc_k_sw128(a + s * 16384 + k * 32, 1024),
                                            umma_desc_k_sw128(aH2 + g * 16384 + s * 8192 + k * 32, 1024), idesc,
                                            (i == 0 && s == 0 && k == 0) ? 0u : 1u);
                    }
                    release_unit();
                }
                if (lane == 0) umma_commit(&h2_empty[g]);
                __syncwarp();
            };
            for (int i = 0; i < P; ++i) {
                const uint32_t gc = static_cast<uint32_t>(B.pair0 + i), g = gc % kRttGroups, n = gc / kRttGroups;
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
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ss(tmem_base + kRttColD1 + g * 64, umma_desc_k_sw128(a + s * 16384 + k * 32, 1024),
                                            umma_desc_k_sw128(aX + (u * 2 + s) * 8192 + k * 32, 1024), idesc,
                                            (u == 0 && s == 0 && k == 0) ? 0u : 1u);
                    }
                    release_unit();
                }
                if (lane == 0) umma_commit(&d1_full[g]);
                __syncwarp();
                if (i >= kTrunkTLag) mma2(i - kTrunkTLag);
            }
            for (int i = P > kTrunkTLag ? P - kTrunkTLag : 0; i < P; ++i) mma2(i);
            if (lane == 0) umma_commit(d2_full);
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
        const int g = cw >> 2;     // epilogue role: (output half, square half)
        const int dg = cw >> 3;    // depthwise role: group dg, output rows 4 dhalf .. 4 dhalf + 3 of the channel's plane
        const int dhalf = g & 1;
        const int tid = cw * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        // epilogue / SE role: output channel ec, squares esq0 .. esq0 + 31 of the X tile
        const int eh = g & 1, ehalf = g >> 1;
        const int ec = eh * 128 + q * 32 + lane;
        const int esq0 = ehalf * 32;
        uint8_t* ex = sX + (ec >> 6) * 8192 + esq0 * 128 + (ec & 7) * 2;
        const uint32_t ec3s = static_cast<uint32_t>(((ec & 63) >> 3) << 4);
        // depthwise role: channel dch of the pair (chunk dch / 64), H2 buffer of the group
        const int dch = q * 32 + lane;
        uint8_t* h2 = sH2 + dg * 16384 + (dch >> 6) * 8192 + (dch & 7) * 2;
        const uint32_t dc3s = static_cast<uint32_t>(((dch & 63) >> 3) << 4);
        RT_PROF_DECL();

        // Squeeze-excitation of block `Bn` on the tile values xh (this thread's channel, its 32 squares, fp16): returns
        // the channel's scale.  Pooling order of rise_trunk.cuh: 16-square sums in butterfly order, ((s0 + s1) + (s2 + s3)).
        auto se_scale = [&](const TrunkBlock& Bn, const uint32_t (&xp)[16]) -> float {
            // (xp: the 32 fp16 tile values, two per register)  The FC matrices arrive through the weight ring: units
            // se_seq0 .. se_seq0 + 3 of the stream
            const uint32_t u0 = static_cast<uint32_t>(Bn.se_seq0);
            auto unit_ptr = [&](uint32_t j) -> const uint8_t* { return sW + ((u0 + j) % kRttRing) * kTrunkTUnit; };
            auto unit_wait = [&](uint32_t j) { mbar_wait(&w_full[(u0 + j) % kRttRing], ((u0 + j) / kRttRing) & 1); };
            auto unit_free = [&](uint32_t j) { mbar_arrive(&w_empty[(u0 + j) % kRttRing]); };
            {
                float v0[16], v1[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 a = rt_unpack(xp[i]), c = rt_unpack(xp[8 + i]);
                    v0[2 * i] = a.x, v0[2 * i + 1] = a.y, v1[2 * i] = c.x, v1[2 * i + 1] = c.y;
                }
                sPoolPart[ehalf * 256 + ec] = rtt_tree16(v0) + rtt_tree16(v1);
            }
            rtt_bar_sync(2);
            if (tid < 256) sPool[tid] = (sPoolPart[tid] + sPoolPart[256 + tid]) * (1.0f / 64.0f);
            rtt_bar_sync(1);
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
                }
                rtt_bar_sync(2);
                if (tid == 0) unit_free(0), unit_free(1), unit_free(2), unit_free(3);
                if (tid < 256) {
                    const float* qq = sPart + tid;
                    sScale[tid] = rt_hard_sigmoid(__ldg(Bn.se_b + tid) + ((qq[0] + qq[256]) + (qq[512] + qq[768])));
                }
            }
            rtt_bar_sync(1);
            return sScale[ec];
        };
        // this thread's 32 tile values: from / to the X tile (element i = square esq0 + i)
        auto load_tile = [&](uint32_t (&xp)[16]) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t lo = *reinterpret_cast<const uint16_t*>(ex + (2 * i) * 128 + (ec3s ^ (((2 * i) & 7) << 4)));
                const uint32_t hi = *reinterpret_cast<const uint16_t*>(ex + (2 * i + 1) * 128 + (ec3s ^ (((2 * i + 1) & 7) << 4)));
                xp[i] = lo | (hi << 16);
            }
        };
        auto store_tile = [&](const uint32_t (&xp)[16], bool scaled, float sc) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t v = xp[i];
                if (scaled) {
                    const float2 f = rt_unpack(v);
                    v = rt_pack(f.x * sc, f.y * sc);
                }
                *reinterpret_cast<uint16_t*>(ex + (2 * i) * 128 + (ec3s ^ (((2 * i) & 7) << 4))) = static_cast<uint16_t>(v & 0xffffu);
                *reinterpret_cast<uint16_t*>(ex + (2 * i + 1) * 128 + (ec3s ^ (((2 * i + 1) & 7) << 4))) = static_cast<uint16_t>(v >> 16);
            }
        };
        auto hand_over = [&]() {
            rt_fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_ready);
        };

        {   // stem output -> the X tile (16-byte pieces into the swizzled K-major layout)
            const uint4* src = reinterpret_cast<const uint4*>(args.x_in + static_cast<size_t>(board) * 64 * 256);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = tid + i * 512;        // piece: row p / 32, 16-byte column p % 32
                const int r = p >> 5, c16 = p & 31;
                const uint4 v = __ldg(src + p);
                *reinterpret_cast<uint4*>(sX + (c16 >> 3) * 8192 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = v;
            }
            rtt_bar_sync(1);  // from here on a thread only touches its own elements of the tile (channel ec, 32 squares)
            if (args.blk[0].se_type != 0) {  // squeeze-excitation on the tower input
                uint32_t xp[16];
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
            float* sB2 = sB2all + (b & 1) * 256;
            if (tid < 256) sB2[tid] = __ldg(B.b2 + tid);
            {
                for (int i = 0; i < P; ++i) {
                    const uint32_t gc = static_cast<uint32_t>(B.pair0 + i);
                    if (static_cast<int>(gc % kRttGroups) != dg) continue;
                    const uint32_t n = gc / kRttGroups;
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
                    mbar_wait(&d1_full[dg], n & 1);
                    RT_PROF(2);  // wait for D1 (tensor core)
                    tc_fence_after();
                    uint32_t h1[32];
                    if (!idle) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t v[32];
                            tmem_ld_32x32b_x32(tmem_base + lane_addr + kRttColD1 + dg * 64 + hh * 32, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                h1[hh * 16 + j] = rt_pack(fmaxf(__uint_as_float(v[2 * j]) + b1, 0.0f), fmaxf(__uint_as_float(v[2 * j + 1]) + b1, 0.0f));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d1_empty[dg]);
                    RT_PROF(3);  // TMEM read-out, bias, relu
                    mbar_wait(&h2_empty[dg], (n & 1) ^ 1);
                    RT_PROF(4);  // wait for the group's H2 buffer
                    if (!idle) {
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
                    if (lane == 0) mbar_arrive(&h2_full[dg]);
                    RT_PROF(5);  // depthwise + H2 write
                }
            }
            // ---- block epilogue: X <- (D2 + b2) + X (the last block: also to global memory), then the next block's SE
            uint32_t xo[16];
            load_tile(xo);  // the block input (this thread's own elements): the residual
            mbar_wait(d2_full, b & 1);
            RT_PROF(6);  // wait for D2
            tc_fence_after();
            rtt_bar_sync(1);  // b2 of this block (written by the first 256 threads before the pair loop) is visible
            uint32_t xp[16];
            {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + lane_addr + kRttColD2 + eh * 64 + esq0, v);
                tmem_ld_wait();
                const float b2 = sB2[ec];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
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
                for (int i = 0; i < 4; ++i) {
                    const int p = tid + i * 512;
                    const int r = p >> 5, c16 = p & 31;
                    dst[p] = *reinterpret_cast<const uint4*>(sX + (c16 >> 3) * 8192 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
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
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kRttTmemCols>(tmem_base);
    }
#endif
}

}  // namespace ara
