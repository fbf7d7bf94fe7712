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

// Debug trace (om_debug_trace): when armed, every conv-engine launch gets the next 4 x uint64 record of a device buffer and stamps
// %globaltimer into it: [0] first CTA start (min), [1] first "dependencies resolved" (min over CTAs of the time griddepcontrol.wait
// returned), [2] last CTA end (max), [3] last CTA start (max).  nullptr (the normal case) costs one predicated branch per CTA.
unsigned long long* trace_next();

}  // namespace om

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_start(unsigned long long* rec) {
    if (rec != nullptr) { const unsigned long long t = trace_now(); atomicMin(rec, t); atomicMax(rec + 3, t); }
}
__device__ __forceinline__ void trace_dep(unsigned long long* rec) { if (rec != nullptr) atomicMin(rec + 1, trace_now()); }
__device__ __forceinline__ void trace_end(unsigned long long* rec) { if (rec != nullptr) atomicMax(rec + 2, trace_now()); }

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
