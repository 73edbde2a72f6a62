// Host handle of the persistent trunk kernel (rise_trunk.cuh): stacks the weights of every bottleneck block into the
// two matrices the kernel streams through its TMA rings and builds the per-chunk vector records.
#pragma once
#include <cuda.h>

#include <vector>

#include "abi_common.h"
#include "rise_trunk_args.h"

namespace ara {

struct TrunkBlockHost {
    int c_op = 0, ksize = 3, se_type = 0;
    std::vector<float> w1;  // [c_op][256]   conv1x1 256 -> c_op (BN folded)
    std::vector<float> b1;  // [c_op]
    std::vector<float> wd;  // [c_op][k*k]   depthwise (BN folded)
    std::vector<float> bd;  // [c_op]
    std::vector<float> w2;  // [256][c_op]   conv1x1 c_op -> 256 (BN folded)
    const float* b2 = nullptr;      // device [256]
    const float* se_w1t = nullptr;  // device, see TrunkBlock
    const float* se_w2t = nullptr;
    const float* se_b = nullptr;
};

struct RiseTrunk {
    TrunkArgs args;
    void* d_w1 = nullptr;
    void* d_w2 = nullptr;
    void* d_timg = nullptr;  // rise_trunk_t.cuh: pair images, vector records, unit order
    void* d_taux = nullptr;
    void* d_tseq = nullptr;
    void* d_cseq = nullptr;  // rise_trunk_c.cuh: unit order of the two cluster ranks
    int sm_count = 148;
    void* d_prof = nullptr;  // [2][16] cycle counters, written only by -DARA_TRUNK_PROF builds
    std::vector<void*> d_se;  // fp16 copies of the squeeze-excitation matrices
};

// x_in: [boards_cap, 8, 8, 256] fp16 (stem output); out: [boards*64, 256] fp16 (may alias x_in: every CTA reads its
// own rows before it writes them)
int rise_trunk_init(RiseTrunk* T, const std::vector<TrunkBlockHost>& blocks, const __half* x_in, int boards_cap, __half* out);
// x_in (optional): another stem-output buffer than the one given to rise_trunk_init (a second input / output set)
int rise_trunk_launch(const RiseTrunk* T, int boards, cudaStream_t stream, const int* boards_dev = nullptr, const __half* x_in = nullptr);
void rise_trunk_destroy(RiseTrunk* T);

}  // namespace ara
