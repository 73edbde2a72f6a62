// Shared host-side plumbing for the C-ABI: last-error string, CUDA error checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>

namespace ara {

std::string& last_error_ref();
int set_error(const char* fmt, ...);

#define ARA_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::ara::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor in the stream is
// still draining; kernels call pdl_wait() before touching the predecessor's outputs.  ARA_NO_PDL=1 disables it.
bool pdl_enabled();
// While one lives on the calling thread, launch_pdl launches (and captures) plain kernels.  A search with Threads = 2 runs
// the network on a second stream beside the tree kernels; thread blocks of early-launched network kernels parked at
// their griddepcontrol.wait then delay the tree stream's launches (measured: 29.8 ms per headline search with the
// attribute, 21.4 ms without), so that mode captures the network without it.
struct PdlSuspend {
    explicit PdlSuspend(bool on);
    ~PdlSuspend();
    PdlSuspend(const PdlSuspend&) = delete;
    PdlSuspend& operator=(const PdlSuspend&) = delete;

   private:
    bool on_;
};
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ara
