/* ara_b200.h -- C-ABI of the B200-native (sm_100a) leaf-evaluation + MCTS engine.
 *
 * Every entry point is what the reference's C++ seam for this hot path would bind (file:line refer to
 * QueensGambit/CrazyAra, engine/src/...).  Plain pointers and sizes only; all functions return 0 on success
 * and -1 on failure with a message retrievable through ara_last_error() (thread local).  There is no CPU
 * fallback: creation fails on anything that is not an sm_100 device.
 */
#ifndef ARA_B200_H
#define ARA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ara_net_s* ara_net_t;

/* Last error message of the calling thread ("" if none). */
const char* ara_last_error(void);

/* ---- Neural network seam: replaces NeuralNetAPI (nn/neuralnetapi.h:148-311) / TensorrtAPI (nn/tensorrtapi.cpp).
 *
 * ara_net_create  <-> TensorrtAPI::TensorrtAPI + NeuralNetAPI::initialize (nn/tensorrtapi.cpp:43-63,
 *                     nn/neuralnetapi.cpp:93-99): loads an ARAB2001 weight blob (crazyara_b200/weights.py), binds
 *                     device buffers and one CUDA stream on `device`, fixed maximum batch size.
 * ara_net_shape   <-> get_nb_input_values_total / get_nb_policy_values / get_nb_auxiliary_outputs /
 *                     is_policy_map / get_version / get_batch_size (nn/neuralnetapi.h:116-293).
 * ara_net_predict <-> NeuralNetAPI::predict(float* inputPlanes, float* valueOutput, float* probOutputs,
 *                     float* auxiliaryOutputs) (nn/neuralnetapi.h:237, nn/tensorrtapi.cpp:195-237):
 *                     planes [n, C, 8, 8] fp32 host -> value [n], prob [n, L] (softmax over ALL L labels, policy-map
 *                     order channel*64+square), aux [n, A].  Synchronous.  n <= batch_size rows are evaluated
 *                     (the reference always runs the full batch; n < B just skips the unused rows).
 */
ara_net_t ara_net_create(const char* weights_path, int device, int batch_size);
void ara_net_destroy(ara_net_t net);
int ara_net_shape(ara_net_t net, int* in_channels, int* n_labels, int* n_aux, int* is_policy_map, int* input_version,
                  int* batch_size);
int ara_net_predict(ara_net_t net, const float* planes, int n, float* value, float* prob, float* aux);

/* Device-resident variant (inputs already in HBM; outputs stay in HBM): planes_dev [n, C, 8, 8] fp32 device
 * pointer, or NULL to evaluate the net's own NHWC fp16 input buffer filled by ara_encode_planes_device. */
int ara_net_forward_device(ara_net_t net, const float* planes_dev, int n, float** value_dev, float** prob_dev);

/* Number of CUDA kernels this net has launched so far (bench bookkeeping). */
long long ara_net_launch_count(ara_net_t net);

/* ---- debug / unit-test entries (one tcgen05 convolution layer on caller-provided device buffers) */
int ara_debug_conv(const void* act_half, int boards_cap, int boards, int cin, const void* w_half, int w_rows, int n_out,
                   int ksize, const float* bias, int relu, const void* residual, int ldr, void* out_half,
                   float* out_f32, int ldo, int bn, void* stream);
int ara_debug_choose_bn(int boards, int n_out);

#ifdef __cplusplus
}
#endif
#endif /* ARA_B200_H */
