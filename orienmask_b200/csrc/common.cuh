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

// Launch with programmatic dependent launch (PDL): the grid may start while its predecessor in the stream drains;
// the kernel must execute `griddepcontrol.wait` (pdl_wait()) before it touches memory the predecessor reads or
// writes.  ORIENMASK_B200_PDL=0 turns the attribute off (plain stream order) for A/B measurements.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace om

#ifdef __CUDACC__
// Device side of PDL: block until every prerequisite grid has completed and its writes are visible, then let the
// next grid in the stream begin its own prologue.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

#define OM_CUDA_TRY(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) return om::fail(OM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
