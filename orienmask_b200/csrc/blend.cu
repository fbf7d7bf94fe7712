// Visualiser mask blend: utils/visualizer.py:46-100 of the reference.
//
//   reference                                                      here
//   ---------                                                      ----
//   InferenceVisualizer._recover_shape_segm  utils/visualizer.py:122-127   crop the padding, F.interpolate(bilinear,
//                                                                            align_corners=False) to the image size, NOT rounded
//   all_mask.sum(2).sum(1).argsort()         utils/visualizer.py:69        soft-mask areas (om_mask_areas), smallest drawn first
//   plot_all_mask                            utils/visualizer.py:95-100    image = image * prod_k(1 - a*m_k) + sum_k c_k*(m_k*a) * prod_{j<k}(1 - a*m_j)
//
// The reference materialises K resized float masks, a [K,h,w,3] colour tensor and a cumulative product (for K = 100 at
// 480x640: ~0.6 GB of temporaries); here one thread owns one image pixel, walks the instances in drawing order, evaluates the
// bilinear blend of four mask bytes and keeps the running product and sum in registers.  HBM-bound: reads K*H*W mask bytes
// (through L2) and the fp32 image once, writes the image once.
//
// Arithmetic: bilinear as in prep.cu / rle.cu (ATen's expression, single-rounded fp32); per instance, in drawing order k:
//   cm = (m * colour) * alpha;  term = (k == 0) ? cm : cm * cum;  cum = cum * (1 - alpha * m)
// and image = image * cum_last + term_0 + (term_1 + term_2 + ...) -- the reference adds the k >= 1 terms with torch.sum, whose
// order is not specified, so results agree to fp32 rounding of that sum, not bit for bit.
#include "common.cuh"

namespace {

__device__ __forceinline__ void src_index(float scale, int d, int n_in, int& i0, int& i1, float& l0, float& l1) {
    float s = __fmaf_rn(scale, __fadd_rn((float)d, 0.5f), -0.5f);
    s = fmaxf(s, 0.0f);
    i0 = min((int)s, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    l1 = __fsub_rn(s, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

struct Tap { long long o00, o01, o10, o11; float lx0, lx1, ly0, ly1; };

__device__ __forceinline__ Tap make_tap(const om_blend_config& c, int y, int x) {
    Tap t;
    int y0, y1, x0, x1;
    src_index((float)c.crop_h / (float)c.out_h, y, c.crop_h, y0, y1, t.ly0, t.ly1);
    src_index((float)c.crop_w / (float)c.out_w, x, c.crop_w, x0, x1, t.lx0, t.lx1);
    const long long r0 = (long long)(c.top + y0) * c.mask_w + c.left, r1 = (long long)(c.top + y1) * c.mask_w + c.left;
    t.o00 = r0 + x0; t.o01 = r0 + x1; t.o10 = r1 + x0; t.o11 = r1 + x1;
    return t;
}

__device__ __forceinline__ float soft_mask(const unsigned char* m, const Tap& t) {
    const float top = __fmaf_rn(t.lx0, (float)__ldg(m + t.o00), __fmul_rn(t.lx1, (float)__ldg(m + t.o01)));
    const float bot = __fmaf_rn(t.lx0, (float)__ldg(m + t.o10), __fmul_rn(t.lx1, (float)__ldg(m + t.o11)));
    return __fmaf_rn(t.ly0, top, __fmul_rn(t.ly1, bot));
}

// areas[k] = sum over the image of the resized (soft) mask k; double accumulation so that the result does not depend on the
// order of the partial sums (the drawing order is an argsort of these)
__global__ void __launch_bounds__(256) mask_area_kernel(om_blend_config c, const unsigned char* __restrict__ mask, double* __restrict__ acc) {
    const int k = blockIdx.y;
    const unsigned char* m = mask + (long long)k * c.mask_h * c.mask_w;
    const long long n = (long long)c.out_h * c.out_w;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / c.out_w), x = (int)(i - (long long)y * c.out_w);
        s += (double)soft_mask(m, make_tap(c, y, x));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(acc + k, t);
    }
}

__global__ void area_finish_kernel(const double* __restrict__ acc, float* __restrict__ areas, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) areas[i] = (float)acc[i];
}

__global__ void __launch_bounds__(256) mask_blend_kernel(om_blend_config c, const unsigned char* __restrict__ mask, int k,
                                                         const int* __restrict__ order, const float* __restrict__ colors,
                                                         float* __restrict__ image) {
    extern __shared__ float s_col[];                  // [k][4]: colour of the j-th drawn instance, then its index
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const int id = order[j];
        s_col[4 * j + 0] = colors[3 * id + 0]; s_col[4 * j + 1] = colors[3 * id + 1]; s_col[4 * j + 2] = colors[3 * id + 2];
        s_col[4 * j + 3] = __int_as_float(id);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)c.out_h * c.out_w) return;
    const int y = (int)(i / c.out_w), x = (int)(i - (long long)y * c.out_w);
    const Tap t = make_tap(c, y, x);
    const long long plane = (long long)c.mask_h * c.mask_w;
    float cum = 1.0f, first[3] = {0.f, 0.f, 0.f}, rest[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < k; ++j) {
        const int id = __float_as_int(s_col[4 * j + 3]);
        const float m = soft_mask(mask + (long long)id * plane, t);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float cm = __fmul_rn(__fmul_rn(m, s_col[4 * j + ch]), c.alpha);           // (mask * colour) * alpha
            if (j == 0) first[ch] = cm; else rest[ch] = __fadd_rn(rest[ch], __fmul_rn(cm, cum));
        }
        cum = __fmul_rn(cum, __fsub_rn(1.0f, __fmul_rn(c.alpha, m)));                       // cumprod of (1 - alpha * mask)
    }
    float* px = image + i * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float v = __fadd_rn(__fmul_rn(px[ch], cum), first[ch]);                               // image.mul_(cum[-1]).add_(cm[0])
        if (k > 1) v = __fadd_rn(v, rest[ch]);                                                // image.add_(sum_k>=1 ...)
        px[ch] = v;
    }
}

int32_t check_cfg(const om_blend_config* c, const char* who) {
    if (!c) return om::fail(OM_ERR_INVALID, "%s: null config", who);
    if (c->mask_h < 1 || c->mask_w < 1 || c->crop_h < 1 || c->crop_w < 1 || c->out_h < 1 || c->out_w < 1 || c->top < 0 || c->left < 0 ||
        c->top + c->crop_h > c->mask_h || c->left + c->crop_w > c->mask_w)
        return om::fail(OM_ERR_INVALID, "%s: the crop window does not fit the %dx%d mask", who, c->mask_h, c->mask_w);
    return OM_OK;
}

}  // namespace

extern "C" int32_t om_mask_areas(const om_blend_config* cfg, const uint8_t* mask, int32_t k, double* scratch, float* areas, void* stream) {
    int32_t rc = check_cfg(cfg, "om_mask_areas");
    if (rc) return rc;
    if (!mask || !scratch || !areas || k < 1 || k > 65535) return om::fail(OM_ERR_INVALID, "om_mask_areas: null argument or k out of range");
    OM_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * k, (cudaStream_t)stream));
    const long long n = (long long)cfg->out_h * cfg->out_w;
    const unsigned gx = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
    mask_area_kernel<<<dim3(gx < 1 ? 1 : gx, (unsigned)k), 256, 0, (cudaStream_t)stream>>>(*cfg, mask, scratch);
    rc = om::check_launch("mask_area_kernel");
    if (rc) return rc;
    area_finish_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scratch, areas, k);
    return om::check_launch("area_finish_kernel");
}

extern "C" int32_t om_mask_blend(const om_blend_config* cfg, const uint8_t* mask, int32_t k, const int32_t* order, const float* colors,
                                 float* image, void* stream) {
    int32_t rc = check_cfg(cfg, "om_mask_blend");
    if (rc) return rc;
    if (!mask || !order || !colors || !image || k < 1 || k > 2048) return om::fail(OM_ERR_INVALID, "om_mask_blend: null argument or k out of range");
    const long long n = (long long)cfg->out_h * cfg->out_w;
    mask_blend_kernel<<<(unsigned)((n + 255) / 256), 256, (size_t)k * 16, (cudaStream_t)stream>>>(*cfg, mask, k, order, colors, image);
    return om::check_launch("mask_blend_kernel");
}
