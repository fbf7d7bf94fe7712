// Shared host-side helpers of the C-ABI library: error text, launch accounting, checked launches.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/orienmask_b200.h"

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (the persistent kernels decode tile coordinates per tile: six runtime integer
// divisions of ~25 instructions each were 22 % of the warp instructions of the narrow memory-bound layers, ncu source view of conv2.0).
// d >= 2: l = ceil(log2 d), mul = ceil(2^(31+l) / d) < 2^32, q = umulhi(n, mul) >> (l - 1); exact because mul * d - 2^(31+l) < 2^l.
struct FastDiv { uint32_t mul, shift, one, d; };

namespace om {

inline FastDiv make_fastdiv(int d) {
    FastDiv f = {0u, 0u, 1u, (uint32_t)d};
    if (d <= 1) return f;
    int l = 0;
    while ((1ll << l) < d) ++l;
    f.mul = (uint32_t)(((1ull << (31 + l)) + (unsigned long long)d - 1) / (unsigned long long)d);
    f.shift = (uint32_t)(l - 1);
    f.one = 0u;
    return f;
}

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
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
    const int q = (int)(__umulhi((uint32_t)n, f.mul) >> f.shift);
    return f.one ? n : q;
}
// Packed fp32 pairs (Blackwell FADD2 / FMUL2: one issue slot for two IEEE fp32 operations, bit-identical to the scalar forms).
__device__ __forceinline__ void add2(float& a, float& b, float x, float y) {
    asm("{\n\t.reg .b64 p, q;\n\tmov.b64 p, {%0, %1};\n\tmov.b64 q, {%2, %3};\n\tadd.rn.f32x2 p, p, q;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "+f"(a), "+f"(b) : "f"(x), "f"(y));
}
__device__ __forceinline__ void mul2(float& o0, float& o1, float a, float b, float x) {
    asm("{\n\t.reg .b64 p, q;\n\tmov.b64 p, {%2, %3};\n\tmov.b64 q, {%4, %4};\n\tmul.rn.f32x2 p, p, q;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=f"(o0), "=f"(o1) : "f"(a), "f"(b), "f"(x));
}
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
