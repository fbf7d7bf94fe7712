// Shared host-side helpers of the C-ABI library: error text, launch accounting, checked launches.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/orienmask_b200.h"

namespace om {

char* error_buffer();                       // thread-local, 512 bytes
int32_t fail(int32_t code, const char* fmt, ...);
void count_launch(int n = 1);

inline int32_t check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    count_launch();
    return OM_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace om

#define OM_CUDA_TRY(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) return om::fail(OM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
