// glibc 2.39 single-precision powf / logf, restated operation by operation so that the device (and the 1-lane host
// emulation) produce the bits the reference engine gets from std::pow / std::log on an x86-64 host:
//   apply_temperature  util/blazeutil.h:78-88   (blaze::pow on a float vector -> scalar powf per element)
//   get_dirichlet_noise util/blazeutil.h:113-124 (std::gamma_distribution<float>: logf, powf inside libstdc++)
// Algorithm: sysdeps/ieee754/flt-32/e_powf.c, e_logf.c (S. Nagy, ARM optimized-routines): 16-entry log2 / log tables,
// 32-entry exp2 table, all arithmetic in double, one final rounding to float.  The x86-64 libm dispatches (ifunc) to
// the variants compiled with -mfma -mavx2 on every FMA machine, where each `a*b+c` of the C source is ONE fused
// multiply-add; the sequence below is the one of those variants (read off `objdump -d libm.so.6`, __powf_fma /
// __logf_fma), written with explicit fma() so that neither nvcc (-fmad) nor g++ (-ffp-contract) decides.
// Tables: tools/extract_glibc_flt32_tables.py.  Pinned against the live libm by tests/test_glibc_flt32.py
// (exhaustive over x for the default exponent 1/1.7, random (x, y) pairs, all special-case branches).
#pragma once
#include <math.h>
#include <stdint.h>

#include "warp_ctx.cuh"

namespace ara {
namespace glibc {

#define ARA_POWF_LOG2_TAB {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2, 0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2, 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3, 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4, 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3, 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}
#define ARA_POWF_LOG2_POLY {0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0}
#define ARA_EXP2F_TAB {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, 0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, 0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, 0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, 0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, 0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, 0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, 0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}
#define ARA_EXP2F_SHIFT_SCALED 0x1.8000000000000p+47
#define ARA_EXP2F_POLY {0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1}
#define ARA_LOGF_TAB {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3, 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4, 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2, 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}
#define ARA_LOGF_LN2 0x1.62e42fefa39efp-1
#define ARA_LOGF_POLY {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2}

static const double h_powf_log2_tab[32] = ARA_POWF_LOG2_TAB;
static const double h_powf_log2_poly[5] = ARA_POWF_LOG2_POLY;
static const uint64_t h_exp2f_tab[32] = ARA_EXP2F_TAB;
static const double h_exp2f_poly[3] = ARA_EXP2F_POLY;
static const double h_logf_tab[32] = ARA_LOGF_TAB;
static const double h_logf_poly[3] = ARA_LOGF_POLY;
#if defined(__CUDACC__)
__device__ const double d_powf_log2_tab[32] = ARA_POWF_LOG2_TAB;
__device__ const double d_powf_log2_poly[5] = ARA_POWF_LOG2_POLY;
__device__ const uint64_t d_exp2f_tab[32] = ARA_EXP2F_TAB;
__device__ const double d_exp2f_poly[3] = ARA_EXP2F_POLY;
__device__ const double d_logf_tab[32] = ARA_LOGF_TAB;
__device__ const double d_logf_poly[3] = ARA_LOGF_POLY;
#endif
#if defined(__CUDA_ARCH__)
#define ARA_GLIBC_TAB(name) d_##name
#define ARA_FMA(a, b, c) __fma_rn((a), (b), (c))
#define ARA_DMUL(a, b) __dmul_rn((a), (b))
#define ARA_DADD(a, b) __dadd_rn((a), (b))
#else
#define ARA_GLIBC_TAB(name) h_##name
#define ARA_FMA(a, b, c) fma((a), (b), (c))
#define ARA_DMUL(a, b) ((a) * (b))
#define ARA_DADD(a, b) ((a) + (b))
#endif

ARA_HD uint32_t f2u(float f) {
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
}
ARA_HD float u2f(uint32_t u) {
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
}
ARA_HD uint64_t d2u(double d) {
    union { double d; uint64_t u; } c;
    c.d = d;
    return c.u;
}
ARA_HD double u2d(uint64_t u) {
    union { double d; uint64_t u; } c;
    c.u = u;
    return c.d;
}

// e_logf.c: __logf.  Domain handled bit-exactly: every float (negative / NaN inputs return NaN without errno).
ARA_HD float logf_(float x) {
    const double* T = ARA_GLIBC_TAB(logf_tab);
    const double* A = ARA_GLIBC_TAB(logf_poly);
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0) return -INFINITY;              // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                 // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return u2f(0x7fc00000u);  // x < 0 or NaN
        ix = f2u(x * 8388608.0f);                        // subnormal: normalise (0x1p23f)
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = static_cast<int32_t>(tmp) >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = T[2 * i], logc = T[2 * i + 1];
    const double z = static_cast<double>(u2f(iz));
    const double r = ARA_FMA(z, invc, -1.0);
    const double y0 = ARA_FMA(static_cast<double>(k), ARA_LOGF_LN2, logc);
    const double r2 = ARA_DMUL(r, r);
    double y = ARA_FMA(A[1], r, A[2]);
    y = ARA_FMA(A[0], r2, y);
    y = ARA_FMA(y, r2, ARA_DADD(y0, r));
    return static_cast<float>(y);
}

// e_powf.c: __powf.  The search only raises priors / uniform variates, i.e. finite x >= 0, finite y > 0; those and the
// zero / subnormal / underflow / overflow branches follow glibc bit for bit.  x < 0, NaN and infinite operands return
// what IEEE pow returns for them but are not exercised by the engine.
ARA_HD float powf_(float x, float y) {
    const double* T = ARA_GLIBC_TAB(powf_log2_tab);
    const double* A = ARA_GLIBC_TAB(powf_log2_poly);
    const uint64_t* E = ARA_GLIBC_TAB(exp2f_tab);
    const double* C = ARA_GLIBC_TAB(exp2f_poly);
    uint32_t ix = f2u(x);
    const uint32_t iy = f2u(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || 2u * iy - 1u >= 2u * 0x7f800000u - 1u) {
        // y is 0, inf or NaN
        if (2u * iy - 1u >= 2u * 0x7f800000u - 1u) {
            if (2u * iy == 0) return 1.0f;
            if (ix == 0x3f800000u) return 1.0f;
            if (2u * ix > 2u * 0x7f800000u || 2u * iy > 2u * 0x7f800000u) return x + y;
            if (2u * ix == 2u * 0x3f800000u) return 1.0f;
            if ((2u * ix < 2u * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;  // |x|<1 && y==inf or |x|>1 && y==-inf
            return y * y;
        }
        // x is 0, inf or NaN (sign of the result for odd integer y is not needed: priors are >= 0)
        if (2u * ix - 1u >= 2u * 0x7f800000u - 1u) {
            float x2 = x * x;
            if (2u * ix == 0 && (iy & 0x80000000u)) return INFINITY;
            return (iy & 0x80000000u) ? 1.0f / x2 : x2;
        }
        if (ix & 0x80000000u) return u2f(0x7fc00000u);  // x < 0: not on the engine's path
        if (ix < 0x00800000u) {                          // subnormal x: normalise
            ix = f2u(x * 8388608.0f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    // log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = static_cast<int32_t>(top) >> 23;
    const double invc = T[2 * i], logc = T[2 * i + 1];
    const double z = static_cast<double>(u2f(iz));
    const double r = ARA_FMA(z, invc, -1.0);
    const double y0 = ARA_DADD(logc, static_cast<double>(k));
    const double r2 = ARA_DMUL(r, r);
    double yy = ARA_FMA(A[0], r, A[1]);
    const double p = ARA_FMA(A[2], r, A[3]);
    const double r4 = ARA_DMUL(r2, r2);
    double q = ARA_FMA(A[4], r, y0);
    q = ARA_FMA(p, r2, q);
    const double logx = ARA_FMA(yy, r4, q);
    const double ylogx = ARA_DMUL(static_cast<double>(y), logx);
    if (((d2u(ylogx) >> 47) & 0xffff) >= (0x405f800000000000ULL >> 47)) {  // |y*log2(x)| >= 126
        if (ylogx > 0x1.fffffffd1d571p+6) return INFINITY;                  // __math_oflowf
        if (ylogx <= -150.0) return 0.0f;                                   // __math_uflowf
        if (ylogx < -149.0) return u2f(1u);                                 // __math_may_uflowf: 0x1.4p-75f squared
    }
    // exp2_inline
    const double kd0 = ARA_DADD(ylogx, ARA_EXP2F_SHIFT_SCALED);
    const uint64_t ki = d2u(kd0);
    const double kd = ARA_DADD(kd0, -(ARA_EXP2F_SHIFT_SCALED));
    const double rr = ARA_DADD(ylogx, -kd);
    uint64_t t = E[ki & 31];
    t += ki << 47;
    const double s = u2d(t);
    const double zz = ARA_FMA(C[0], rr, C[1]);
    const double rr2 = ARA_DMUL(rr, rr);
    double w = ARA_FMA(C[2], rr, 1.0);
    w = ARA_FMA(zz, rr2, w);
    w = ARA_DMUL(w, s);
    return static_cast<float>(w);
}

}  // namespace glibc
}  // namespace ara
