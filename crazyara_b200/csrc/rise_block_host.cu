#include "rise_block_host.h"

#include "rise_block.cuh"

#include <cstring>

namespace ara {

int rise_block_init(RiseBlockLayer* L, const __half* x, int boards_cap, const __half* w1, int w1_rows, const __half* w2,
                    int w2_rows, int w2_k, int c_op, int ksize, const float* b1p, const float* wdp, const float* bdp,
                    const float* b2, __half* out) {
    memset(L, 0, sizeof(*L));
    const int n_chunks = (c_op + 63) / 64;
    if (ksize != 3 && ksize != 5) return set_error("rise_block_init: depthwise kernel %d unsupported", ksize);
    if (w1_rows < n_chunks * 64 || w2_k != n_chunks * 64 || w2_rows < 256)
        return set_error("rise_block_init: weight padding mismatch (w1_rows %d, w2_rows %d, w2_k %d, chunks %d)", w1_rows, w2_rows,
                         w2_k, n_chunks);
    if (make_act_tensor_map(&L->tm_x, x, boards_cap, 256)) return -1;
    if (make_weight_tensor_map(&L->tm_w1, w1, 256, w1_rows, 64)) return -1;
    if (make_weight_tensor_map(&L->tm_w2, w2, w2_k, w2_rows, 256)) return -1;
    L->args.M = 0;
    L->args.n_chunks = n_chunks;
    L->args.ksize = ksize;
    L->args.cpad = n_chunks * 64;
    L->args.b1 = b1p;
    L->args.wd = wdp;
    L->args.bd = bdp;
    L->args.b2 = b2;
    L->args.out = out;
    return 0;
}

int rise_block_launch(const RiseBlockLayer* L, int boards, cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        ARA_CUDA_OK(cudaFuncSetAttribute(rise_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRbSmemBytes));
        attr_done = true;
    }
    RiseBlockArgs a = L->args;
    a.M = boards * 64;
    ARA_CUDA_OK(launch_pdl(rise_block_kernel, dim3((boards + 1) / 2), dim3(kRbThreads), kRbSmemBytes, stream, L->tm_x, L->tm_w1,
                           L->tm_w2, a));
    return 0;
}

}  // namespace ara
