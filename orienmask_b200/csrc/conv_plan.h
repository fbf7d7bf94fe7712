// Internal interface between the C-ABI entry points (api.cu) and the two convolution engines.
#pragma once
#include <cuda_runtime.h>
#include <cstring>

#include "../../include/orienmask_b200.h"

namespace om {

// fp16 tcgen05 engine on CTA pairs, cta_group::2 (conv_tc2.cu)
int32_t tc2_plan_create(const om_conv_desc& d, void** out);
int32_t tc2_plan_run(const void* plan, cudaStream_t stream, void* output = nullptr);
void tc2_plan_destroy(void* plan);
// debug / tests: the planner's decisions for one layer (see om_debug_conv_plan_info in api.cu for the field order)
void tc2_plan_info(const void* plan, int32_t* info24);

// tensor-core first layer (conv_stem_tc.cu)
int32_t stem_tc_run(const float* image, const float* weights, const float* bias, void* output, int batch, int h, int w, int rows,
                    int out_s2d, cudaStream_t stream, int split = 0);

// first two layers in one launch: stem 3 -> 32 + stride-2 32 -> 64 (stem_fused.cu)
bool stem_fused_supported(int h, int w, const void* image);
int32_t stem_fused_run(const float* image, const float* w27, const float* b1, const void* w2, const float* b2, void* out, int batch, int h,
                       int w, int out_rows, cudaStream_t stream);

// fused DarkNet block, C = 32 (dark_block.cu)
bool dark_block_supported(int cin, int cmid, int width, int rows);
int32_t dark_block_run(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, void* out, int batch, int height,
                       int width, int rows, int out_s2d, cudaStream_t stream);

// fp32 FFMA parity engine (conv_f32.cu)
int32_t f32_conv_run(const om_conv_desc& d, cudaStream_t stream);

}  // namespace om
