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

}  // namespace ara
