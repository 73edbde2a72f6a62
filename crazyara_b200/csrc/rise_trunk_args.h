// Argument block of the persistent trunk kernel (rise_trunk.cuh): every bottleneck block of the RISE tower.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ara {

constexpr int kTrunkMaxBlocks = 24;

struct TrunkBlock {
    int n_chunks;     // ceil(Cop / 64)
    int ksize;        // depthwise kernel: 3 or 5
    int se_type;      // 0 none, 1 ca_se, 2 eca_se (applied to the block input, in place)
    int row0;         // first row of this block in the stacked W1 matrix == first K column in the stacked W2 matrix
    int aux_off;      // byte offset of the block's first per-chunk record in `aux`
    int aux_bytes;    // record size: (128 + k*k*64) * 4
    const float* b2;      // [256]
    const __half* se_w1t;  // ca_se: [256][128]; eca_se: [256][256] (transposed, fp16 copy owned by the trunk)
    const __half* se_w2t;  // ca_se: [128][256]
    const float* se_b;     // eca_se: [256]
};

struct TrunkArgs {
    int M;         // valid rows (= boards * 64)
    int n_blocks;
    const uint8_t* aux;  // per 64-channel chunk: b1[64] f32 | bd[64] f32 | wd[k*k][64] f32
    __half* out;         // [M, 256]
    unsigned long long* prof;  // profiling builds (-DARA_TRUNK_PROF): [2][16] cycle counters of CTA 0, else null
    TrunkBlock blk[kTrunkMaxBlocks];
};

}  // namespace ara
