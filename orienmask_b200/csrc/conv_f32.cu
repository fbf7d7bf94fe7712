// fp32 parity engine: FFMA implicit-GEMM convolution and the 3-channel stem.
//
// This engine exists so that the full path can be checked against the reference's fp32 CPU output
// inside the reference's own reorder noise (SURVEY §7 "hard parts": an fp16 tensor-core forward
// cannot meet the 1e-3 score tolerance on synthetic weights).  It shares the activation layout,
// the folded-BN weights and every epilogue option with the tcgen05 engine (conv_tc.cu), and is the
// bring-up reference for it on the GPU.  Same math as model/base.py:104-137 with BN folded:
// y = leaky(conv(x, W') + b' [+ upsampled partial]) [+ residual].
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, APAD = 4;

struct F32Params {
    int batch, in_h, in_w, in_rows, out_h, out_w, out_rows, cin, cout, cout_pad, cout_stride;
    int ksize, stride, leaky, out_kind, up_rows, in_s2d, out_s2d;
    const float* in; const float* w; const float* bias; const float* residual; const float* upadd;
    float* out;
};

__global__ void __launch_bounds__(256) conv_f32_kernel(const F32Params p) {
    __shared__ __align__(16) float As[BK][BM + APAD];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M = p.batch * p.out_h * p.out_w;
    const int pad = p.ksize / 2;

    // A loader: one pixel, 4 consecutive channels
    const int pa = tid >> 2, kq = tid & 3;
    const int ma = m0 + pa;
    int an = 0, aoy = 0, aox = 0;
    const bool a_ok = ma < M;
    if (a_ok) { an = ma / (p.out_h * p.out_w); const int r = ma - an * p.out_h * p.out_w; aoy = r / p.out_w; aox = r - aoy * p.out_w; }
    // B loader: one k row, 4 consecutive couts
    const int kb = tid >> 4, cq = tid & 15;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
        const int r = tap / p.ksize, s = tap - r * p.ksize;
        const int iy = aoy * p.stride + r - pad, ix = aox * p.stride + s - pad;
        const bool in_ok = a_ok && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
        size_t apix = (size_t)(an * p.in_rows + iy) * p.in_w + ix;
        if (p.in_s2d) {                                    // parity-split input: plane 2*(Y&1)+(x&1), then (Y>>1, x>>1)
            const int Yi = an * p.in_rows + iy;
            apix = (size_t)(2 * (Yi & 1) + (ix & 1)) * ((size_t)p.batch * p.in_rows / 2 * (p.in_w / 2)) + (size_t)(Yi >> 1) * (p.in_w / 2) + (ix >> 1);
        }
        const float* arow = p.in + apix * p.cin + kq * 4;
        const float* brow = p.w + ((size_t)tap * p.cin + kb) * p.cout_pad + n0 + cq * 4;
        const bool b_ok = n0 + cq * 4 < p.cout_pad;
        for (int c0 = 0; c0 < p.cin; c0 += BK) {
            const float4 av = in_ok ? *reinterpret_cast<const float4*>(arow + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 bv = b_ok ? *reinterpret_cast<const float4*>(brow + (size_t)c0 * p.cout_pad) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            As[kq * 4 + 0][pa] = av.x; As[kq * 4 + 1][pa] = av.y; As[kq * 4 + 2][pa] = av.z; As[kq * 4 + 3][pa] = av.w;
            *reinterpret_cast<float4*>(&Bs[kb][cq * 4]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int n = m / (p.out_h * p.out_w);
        const int rr = m - n * p.out_h * p.out_w;
        const int oy = rr / p.out_w, ox = rr - oy * p.out_w;
        const size_t pix = (size_t)(n * p.out_rows + oy) * p.out_w + ox;
        size_t opix = pix;
        if (p.out_s2d) {
            const int Yo = n * p.out_rows + oy;
            opix = (size_t)(2 * (Yo & 1) + (ox & 1)) * ((size_t)p.batch * p.out_rows / 2 * (p.out_w / 2)) + (size_t)(Yo >> 1) * (p.out_w / 2) + (ox >> 1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= p.cout) continue;
            float v = acc[i][j];
            if (p.upadd) v += p.upadd[((size_t)(n * p.up_rows + (oy >> 1)) * (p.out_w >> 1) + (ox >> 1)) * p.cout + c];
            if (p.out_kind != OM_OUT_PARTIAL && p.bias) v += p.bias[c];
            if (p.leaky) v = v > 0.f ? v : 0.1f * v;
            if (p.out_kind == OM_OUT_NCHW) {
                p.out[((size_t)(n * p.cout + c) * p.out_h + oy) * p.out_w + ox] = v;
            } else {
                if (p.residual) v += p.residual[pix * p.cout_stride + c];
                p.out[opix * p.cout_stride + c] = v;
            }
        }
    }
}

// Stem: one thread per output pixel, all `COUT` channels; weights broadcast from shared memory.
// SPLIT (OM_PREC_SPLIT): the fp32 result is stored as an fp16 pair, [.., 2 * COUT] halves per pixel (hi | lo).
template <typename OutT, int COUT, bool SPLIT = false>
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                   const float* __restrict__ bias, OutT* __restrict__ out,
                                                   int batch, int h, int wd, int rows, int out_s2d) {
    __shared__ float sw[27 * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)batch * h * wd) return;
    const int n = (int)(idx / ((long long)h * wd));
    const int r = (int)(idx - (long long)n * h * wd);
    const int y = r / wd, x = r - y * wd;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = x + kx - 1;
            const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < wd;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v = ok ? __ldg(img + ((size_t)(n * 3 + ci) * h + iy) * wd + ix) : 0.f;
                const float* wr = sw + ((ky * 3 + kx) * 3 + ci) * COUT;
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = acc[c] > 0.f ? acc[c] : 0.1f * acc[c];
    size_t opix = (size_t)(n * rows + y) * wd + x;
    if (out_s2d) {
        const int Yo = n * rows + y;
        opix = (size_t)(2 * (Yo & 1) + (x & 1)) * ((size_t)batch * rows / 2 * (wd / 2)) + (size_t)(Yo >> 1) * (wd / 2) + (x >> 1);
    }
    OutT* o = out + opix * (SPLIT ? 2 * COUT : COUT);
    if constexpr (SPLIT) {
#pragma unroll
        for (int c = 0; c < COUT; c += 8) {
            uint4 hv, lv;
            __half2* hh = reinterpret_cast<__half2*>(&hv);
            __half2* lh = reinterpret_cast<__half2*>(&lv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hh[j] = __floats2half2_rn(acc[c + 2 * j], acc[c + 2 * j + 1]);
                const float2 back = __half22float2(hh[j]);
                lh[j] = __floats2half2_rn(acc[c + 2 * j] - back.x, acc[c + 2 * j + 1] - back.y);
            }
            *reinterpret_cast<uint4*>(o + c) = hv;
            *reinterpret_cast<uint4*>(o + COUT + c) = lv;
        }
    } else if constexpr (sizeof(OutT) == 2) {
#pragma unroll
        for (int c = 0; c < COUT; c += 8) {
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[c + 2 * j], acc[c + 2 * j + 1]);
            *reinterpret_cast<uint4*>(o + c) = ov;
        }
    } else {
#pragma unroll
        for (int c = 0; c < COUT; c += 4)
            *reinterpret_cast<float4*>(o + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    }
}

}  // namespace

namespace om {

int32_t f32_conv_run(const om_conv_desc& d, cudaStream_t stream) {
    if (d.cin % BK) return fail(OM_ERR_INVALID, "fp32 engine needs cin %% 16 == 0 (got %d)", d.cin);
    F32Params p;
    p.batch = d.batch; p.in_h = d.in_h; p.in_w = d.in_w; p.in_rows = d.in_rows;
    p.out_h = d.out_h; p.out_w = d.out_w; p.out_rows = d.out_rows; p.cin = d.cin; p.cout = d.cout;
    p.cout_pad = (d.cout + 3) / 4 * 4; p.cout_stride = d.cout_stride;
    p.ksize = d.ksize; p.stride = d.stride; p.leaky = d.leaky; p.out_kind = d.out_kind; p.up_rows = d.up_rows;
    p.in_s2d = d.in_s2d; p.out_s2d = d.out_s2d;
    p.in = reinterpret_cast<const float*>(d.input); p.w = reinterpret_cast<const float*>(d.weights);
    p.bias = d.bias; p.residual = reinterpret_cast<const float*>(d.residual); p.upadd = d.upadd;
    p.out = reinterpret_cast<float*>(d.output);
    const int M = d.batch * d.out_h * d.out_w;
    dim3 grid(ceil_div(M, BM), ceil_div(d.cout, BN));
    conv_f32_kernel<<<grid, 256, 0, stream>>>(p);
    return check_launch("conv_f32_kernel");
}

}  // namespace om

extern "C" int32_t om_stem_conv(int32_t precision, const float* image, const float* weights, const float* bias, void* output,
                                int32_t batch, int32_t h, int32_t w, int32_t rows, int32_t cout, int32_t out_s2d, void* stream) {
    if (!image || !weights || !bias || !output) return om::fail(OM_ERR_INVALID, "om_stem_conv: null argument");
    if (cout != 32) return om::fail(OM_ERR_UNSUPPORTED, "om_stem_conv: cout must be 32 (got %d)", cout);
    if (batch < 1 || h < 1 || w < 1 || rows <= h) return om::fail(OM_ERR_INVALID, "om_stem_conv: bad geometry");
    if (out_s2d && (rows % 2 || w % 2)) return om::fail(OM_ERR_INVALID, "om_stem_conv: out_s2d needs even rows and width");
    const long long total = (long long)batch * h * w;
    const int blocks = (int)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == OM_PREC_F16) {
        const char* sel = getenv("ORIENMASK_B200_STEM");
        if (!(sel && sel[0] == 'f') && h % 4 == 0 && w % 32 == 0)          // default: tensor-core stem (conv_stem_tc.cu)
            return om::stem_tc_run(image, weights, bias, output, batch, h, w, rows, out_s2d, st);
        stem_kernel<__half, 32><<<blocks, 256, 0, st>>>(image, weights, bias, reinterpret_cast<__half*>(output), batch, h, w, rows, out_s2d);
    } else if (precision == OM_PREC_SPLIT) {
        const char* sel = getenv("ORIENMASK_B200_STEM");
        if (!(sel && sel[0] == 'f') && h % 4 == 0 && w % 32 == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0)   // tensor cores, hi + lo pairs
            return om::stem_tc_run(image, weights, bias, output, batch, h, w, rows, out_s2d, st, 1);
        stem_kernel<__half, 32, true><<<blocks, 256, 0, st>>>(image, weights, bias, reinterpret_cast<__half*>(output), batch, h, w, rows, out_s2d);
    }
    else if (precision == OM_PREC_F32)
        stem_kernel<float, 32><<<blocks, 256, 0, st>>>(image, weights, bias, reinterpret_cast<float*>(output), batch, h, w, rows, out_s2d);
    else
        return om::fail(OM_ERR_INVALID, "om_stem_conv: unknown precision %d", precision);
    return om::check_launch("stem_kernel");
}
