// Lane abstraction: the search / rules code is written as "lane-strided loops + warp votes" so that the same source
// runs as one 32-lane warp per tree on the GPU and as a 1-lane emulation on the host (unit tests on a CPU-only box;
// the host build is test scaffolding, never a product fallback).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ARA_HD __host__ __device__ __forceinline__
#else
#define ARA_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define ARA_LANE (static_cast<int>(threadIdx.x) & 31)
#define ARA_WARP_N 32
#define ARA_WARP_SYNC() __syncwarp()
#define ARA_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define ARA_ALL(p) (__all_sync(0xffffffffu, (p)) != 0)
#define ARA_POPC(m) __popc(m)
#define ARA_POPC_BELOW(m) __popc((m) & ((1u << ARA_LANE) - 1u))
#define ARA_SHFL(v, src) __shfl_sync(0xffffffffu, (v), (src))
#define ARA_SHFL_XOR(v, x) __shfl_xor_sync(0xffffffffu, (v), (x))
#define ARA_REDUCE_MAX(v) __reduce_max_sync(0xffffffffu, (v))
#define ARA_REDUCE_MIN(v) __reduce_min_sync(0xffffffffu, (v))
#else
#define ARA_LANE 0
#define ARA_WARP_N 1
#define ARA_WARP_SYNC() ((void)0)
#define ARA_BALLOT(p) ((p) ? 1u : 0u)
#define ARA_ALL(p) (p)
#define ARA_POPC(m) __builtin_popcount(m)
#define ARA_POPC_BELOW(m) 0
#define ARA_SHFL(v, src) (v)
#define ARA_SHFL_XOR(v, x) (v)
#define ARA_REDUCE_MAX(v) (v)
#define ARA_REDUCE_MIN(v) (v)
#endif
