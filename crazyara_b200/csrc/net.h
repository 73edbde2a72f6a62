// RISE network instance on one GPU stream: weights resident in HBM, activation buffers, launch sequence.
// Mirrors the role of the reference's NeuralNetAPI/TensorrtAPI (engine/src/nn/neuralnetapi.h:148-311,
// engine/src/nn/tensorrtapi.cpp:160-237) but the "engine" is our own kernel sequence.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "conv_gemm_host.h"
#include "rise_block_host.h"
#include "rise_trunk_host.h"

namespace ara {

struct BlockDesc {
    int c_op;
    int kernel;   // 3 or 5
    int se_type;  // 0 none, 1 ca_se, 2 eca_se
};

struct NetHeader {
    int in_channels;
    int policy_channels;
    int n_blocks;
    int channels;        // 256
    int value_channels;  // 8
    int value_fc;        // 256
    int wdl_mode;        // 1: value head with WDL + plys-to-end auxiliary outputs
    int input_version;   // e.g. 10 = v1.0, 30 = v3.0
};

struct BlockW {
    float *se_w1t = nullptr, *se_w2t = nullptr, *se_b = nullptr;
    __half* w1 = nullptr;
    float* b1 = nullptr;
    float* wd = nullptr;
    float* bd = nullptr;
    __half* w2 = nullptr;
    float* b2 = nullptr;
    ConvLayer conv1, conv2;
    float* wd_pad = nullptr;  // [k*k][ceil64(c_op)] for the fused block kernel
    RiseBlockLayer fused;
};

class Net {
   public:
    Net() = default;
    ~Net();
    int init(const char* blob_path, int device, int batch);
    // host-buffer API (reference NeuralNetAPI::predict semantics, synchronous)
    int predict(const float* planes_host, int n, float* value_host, float* prob_host, float* aux_host);
    // device-resident API: input already in in_h (NHWC fp16), outputs stay in d_value / d_prob
    // boards_dev (optional): device-side count (<= n) of the input rows that really hold positions -- the launch is
    // sized for n, thread blocks of the rows beyond the count leave at once; the pointer is baked into the CUDA graph
    int forward_device(int n, cudaStream_t stream, const int* boards_dev = nullptr);
    int forward_from_f32_device(int n, cudaStream_t stream);  // converts d_in_f32 -> in_h first

    NetHeader hdr{};
    std::vector<BlockDesc> blocks;
    int device = 0;
    int batch = 0;      // max boards per call
    int batch_cap = 0;  // even, >= 2
    int cin_pad = 0;
    int ldp = 0;  // padded policy channels (multiple of 32)
    int n_labels() const { return hdr.policy_channels * 64; }
    int n_aux() const { return hdr.wdl_mode ? 4 : 0; }
    int kernels_per_forward(bool from_f32) const;
    int trunk_cycles(unsigned long long* out32);  // -DARA_TRUNK_PROF builds: per-role cycle counters of CTA 0
    cudaStream_t stream = nullptr;
    // the value head runs beside the policy head: forked side stream, joined before the forward ends
    cudaStream_t head_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool fork_heads = true;  // ARA_NET_FORK_HEADS=0: both heads in sequence on one stream

    // device buffers
    float* d_in_f32 = nullptr;   // [batch, C, 64]
    __half* d_in_h = nullptr;    // [batch_cap, 64, cin_pad]
    __half* d_x[2] = {nullptr, nullptr};  // trunk ping-pong [batch_cap*64, 256]
    __half* d_h1 = nullptr;      // [batch_cap*64, max_cop]
    __half* d_h2 = nullptr;
    __half* d_p1 = nullptr;      // [batch_cap*64, 256]
    float* d_logits = nullptr;   // [batch_cap*64, ldp]
    float* d_prob = nullptr;     // [batch, L]
    float* d_value = nullptr;    // [batch]
    float* d_aux = nullptr;      // [batch, 4]
    long long launches = 0;      // kernels launched so far (bench bookkeeping)
    bool use_graph = true;
    bool use_fused = false;  // ARA_FUSED_BLOCKS=1: one kernel per bottleneck block (rise_block.cuh)
    bool use_trunk = true;   // the whole residual tower as one persistent kernel (rise_trunk.cuh); ARA_TRUNK=0 disables

   private:
    int enqueue(int n, cudaStream_t s, bool from_f32, const int* boards_dev = nullptr);
    std::vector<void*> allocs_;
    __half* stem_w = nullptr;
    float* stem_b = nullptr;
    ConvLayer stem_conv;
    std::vector<BlockW> bw_;
    RiseTrunk trunk_;
    float *vh_wv = nullptr, *vh_bv = nullptr, *vh_w1t = nullptr, *vh_b1 = nullptr, *vh_w2 = nullptr, *vh_b2 = nullptr;
    float *vh_wdl_w = nullptr, *vh_wdl_b = nullptr, *vh_plys_w = nullptr, *vh_plys_b = nullptr;
    __half *pol_w1 = nullptr, *pol_w2 = nullptr;
    float* pol_b1 = nullptr;
    ConvLayer pol_conv1, pol_conv2;
    std::map<int, cudaGraphExec_t> graphs_[3];  // plain, from fp32 input, with a device-side count
    const int* count_ptr_ = nullptr;            // the pointer the graphs_[2] entries were captured with
    template <typename T>
    int dalloc(T** p, size_t count);
    int upload_conv_w(const float* w, int n_out, int cin, int ksize, __half** dst, int* rows);
    int upload_f32(const float* src, size_t count, size_t padded, float** dst);
};

}  // namespace ara
