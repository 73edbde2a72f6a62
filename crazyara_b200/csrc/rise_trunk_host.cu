#include "rise_trunk_host.h"

#include <cstring>

#include "rise_block_host.h"  // make_act_tensor_map / make_weight_tensor_map
#include "rise_trunk.cuh"

namespace ara {

int rise_trunk_init(RiseTrunk* T, const std::vector<TrunkBlockHost>& blocks, const __half* x_in, int boards_cap, __half* out) {
    const int nb = static_cast<int>(blocks.size());
    if (nb < 1 || nb > kTrunkMaxBlocks) return set_error("rise_trunk_init: %d blocks unsupported (max %d)", nb, kTrunkMaxBlocks);
    memset(&T->args, 0, sizeof(T->args));
    int rows = 0;
    size_t aux_bytes = 0;
    for (int i = 0; i < nb; ++i) {
        const TrunkBlockHost& h = blocks[i];
        if (h.ksize != 3 && h.ksize != 5) return set_error("rise_trunk_init: depthwise kernel %d unsupported", h.ksize);
        if (h.c_op < 1) return set_error("rise_trunk_init: block %d has no operating channels", i);
        TrunkBlock& B = T->args.blk[i];
        B.n_chunks = (h.c_op + 63) / 64;
        B.ksize = h.ksize;
        B.se_type = h.se_type;
        B.row0 = rows;
        B.aux_off = static_cast<int>(aux_bytes);
        B.aux_bytes = 512 + h.ksize * h.ksize * 128;  // b1[64] f32 | bd[64] f32 | wd[k*k][64] f16
        B.b2 = h.b2;
        B.se_b = h.se_b;
        if (h.se_type != 0) {  // fp16 copies of the squeeze-excitation matrices (the kernel is bound by their traffic)
            const size_t n1 = h.se_type == 1 ? 256 * 128 : 256 * 256, n2 = h.se_type == 1 ? 128 * 256 : 0;
            std::vector<float> f(n1 + n2);
            ARA_CUDA_OK(cudaMemcpy(f.data(), h.se_w1t, n1 * 4, cudaMemcpyDeviceToHost));
            if (n2) ARA_CUDA_OK(cudaMemcpy(f.data() + n1, h.se_w2t, n2 * 4, cudaMemcpyDeviceToHost));
            std::vector<__half> hh(n1 + n2);
            for (size_t k = 0; k < hh.size(); ++k) hh[k] = __float2half_rn(f[k]);
            void* d = nullptr;
            ARA_CUDA_OK(cudaMalloc(&d, hh.size() * sizeof(__half)));
            ARA_CUDA_OK(cudaMemcpy(d, hh.data(), hh.size() * sizeof(__half), cudaMemcpyHostToDevice));
            T->d_se.push_back(d);
            B.se_w1t = static_cast<const __half*>(d);
            B.se_w2t = n2 ? static_cast<const __half*>(d) + n1 : nullptr;
        }
        rows += B.n_chunks * 64;
        aux_bytes += static_cast<size_t>(B.n_chunks) * B.aux_bytes;
    }
    T->args.n_blocks = nb;
    T->args.out = out;
    // W1 stacked by rows: [rows][256]; W2 stacked along K: [256][rows]; both fp16, zero padded per block to 64
    std::vector<__half> w1(static_cast<size_t>(rows) * 256, __float2half(0.0f));
    std::vector<__half> w2(static_cast<size_t>(256) * rows, __float2half(0.0f));
    std::vector<uint8_t> aux(aux_bytes, 0);
    for (int i = 0; i < nb; ++i) {
        const TrunkBlockHost& h = blocks[i];
        const TrunkBlock& B = T->args.blk[i];
        const int kk = h.ksize * h.ksize;
        for (int c = 0; c < h.c_op; ++c)
            for (int k = 0; k < 256; ++k)
                w1[(static_cast<size_t>(B.row0) + c) * 256 + k] = __float2half_rn(h.w1[static_cast<size_t>(c) * 256 + k]);
        for (int n = 0; n < 256; ++n)
            for (int c = 0; c < h.c_op; ++c)
                w2[static_cast<size_t>(n) * rows + B.row0 + c] = __float2half_rn(h.w2[static_cast<size_t>(n) * h.c_op + c]);
        for (int j = 0; j < B.n_chunks; ++j) {
            uint8_t* rec = aux.data() + B.aux_off + static_cast<size_t>(j) * B.aux_bytes;
            float* rec_f = reinterpret_cast<float*>(rec);
            __half* rec_w = reinterpret_cast<__half*>(rec + 512);
            for (int cc = 0; cc < 64; ++cc) {
                const int c = j * 64 + cc;
                if (c >= h.c_op) break;
                rec_f[cc] = h.b1[c];
                rec_f[64 + cc] = h.bd[c];
                for (int q = 0; q < kk; ++q) rec_w[q * 64 + cc] = __float2half_rn(h.wd[static_cast<size_t>(c) * kk + q]);
            }
        }
    }
    ARA_CUDA_OK(cudaMalloc(&T->d_w1, w1.size() * sizeof(__half)));
    ARA_CUDA_OK(cudaMalloc(&T->d_w2, w2.size() * sizeof(__half)));
    ARA_CUDA_OK(cudaMalloc(&T->d_aux, aux.size()));
    ARA_CUDA_OK(cudaMemcpy(T->d_w1, w1.data(), w1.size() * sizeof(__half), cudaMemcpyHostToDevice));
    ARA_CUDA_OK(cudaMemcpy(T->d_w2, w2.data(), w2.size() * sizeof(__half), cudaMemcpyHostToDevice));
    ARA_CUDA_OK(cudaMemcpy(T->d_aux, aux.data(), aux.size(), cudaMemcpyHostToDevice));
    T->args.aux = static_cast<const uint8_t*>(T->d_aux);
    ARA_CUDA_OK(cudaMalloc(&T->d_prof, 32 * sizeof(unsigned long long)));
    ARA_CUDA_OK(cudaMemset(T->d_prof, 0, 32 * sizeof(unsigned long long)));
    T->args.prof = static_cast<unsigned long long*>(T->d_prof);
    if (make_act_tensor_map(&T->tm_x, x_in, boards_cap, 256)) return -1;
    if (make_weight_tensor_map(&T->tm_w1, static_cast<const __half*>(T->d_w1), 256, rows, 64)) return -1;
    if (make_weight_tensor_map(&T->tm_w2, static_cast<const __half*>(T->d_w2), rows, 256, 128)) return -1;
    ARA_CUDA_OK(cudaFuncSetAttribute(rise_trunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSmemBytes));
    return 0;
}

int rise_trunk_launch(const RiseTrunk* T, int boards, cudaStream_t stream) {
    TrunkArgs a = T->args;
    a.M = boards * 64;
    ARA_CUDA_OK(launch_pdl(rise_trunk_kernel, dim3((boards + 1) / 2), dim3(kRtThreads), kRtSmemBytes, stream, T->tm_x,
                           T->tm_w1, T->tm_w2, a));
    return 0;
}

void rise_trunk_destroy(RiseTrunk* T) {
    if (T->d_w1) cudaFree(T->d_w1);
    if (T->d_w2) cudaFree(T->d_w2);
    if (T->d_aux) cudaFree(T->d_aux);
    if (T->d_prof) cudaFree(T->d_prof);
    for (void* p : T->d_se) cudaFree(p);
    T->d_se.clear();
    T->d_w1 = T->d_w2 = T->d_aux = T->d_prof = nullptr;
}

}  // namespace ara
