// Host handle of the fused bottleneck-block kernel (rise_block.cuh).
#pragma once
#include "abi_common.h"
#include <cuda.h>

#include "rise_block_args.h"

namespace ara {

struct RiseBlockLayer {
    CUtensorMap tm_x, tm_w1, tm_w2;
    RiseBlockArgs args;
};

// x: [boards_cap, 8, 8, 256] fp16; w1: [w1_rows >= n_chunks*64, 256] fp16 K-major; w2: [>= 256 rows, w2_k = n_chunks*64]
// fp16 K-major; b1p / bdp: [n_chunks*64]; wdp: [k*k][n_chunks*64] (zero padded); b2: [256]; out: [boards*64, 256].
int rise_block_init(RiseBlockLayer* L, const __half* x, int boards_cap, const __half* w1, int w1_rows, const __half* w2,
                    int w2_rows, int w2_k, int c_op, int ksize, const float* b1p, const float* wdp, const float* bdp,
                    const float* b2, __half* out);
int rise_block_launch(const RiseBlockLayer* L, int boards, cudaStream_t stream);

// shared with conv_gemm_host.cu
int make_act_tensor_map(CUtensorMap* m, const __half* act, int boards_cap, int cin);
int make_weight_tensor_map(CUtensorMap* m, const __half* w, int k_total, int rows, int box_rows);

}  // namespace ara
