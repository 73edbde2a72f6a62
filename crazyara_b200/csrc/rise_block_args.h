// Argument block of the fused bottleneck-block kernel (shared by the kernel and its host handle).
#pragma once
#include <cuda_fp16.h>

namespace ara {

struct RiseBlockArgs {
    int M;         // valid rows (= boards * 64)
    int n_chunks;  // ceil(Cop / 64)
    int ksize;     // depthwise kernel: 3 or 5
    int cpad;      // n_chunks * 64: pitch of the per-channel vectors below
    const float* b1;  // [cpad] conv1 bias (BN folded), zero padded
    const float* wd;  // [k*k][cpad] depthwise weights (BN folded), zero padded
    const float* bd;  // [cpad]
    const float* b2;  // [256]
    __half* out;      // [M, 256]
};

}  // namespace ara
