// First layer (3 -> 32, 3x3, stride 1, BN folded, LeakyReLU) on the sm_100a tensor cores, straight from the
// caller's fp32 NCHW image to the fp16 padded-row NHWC activation the rest of the engine consumes.
//
// GEMM view: M = pixels, K = 27 taps*channels (zero-padded to 32), N = 32.  The tensor work is trivial (two
// K=16 tcgen05 MMAs per 128 pixels); what matters is that the layer moves 12 B in + 64 B out per pixel and
// nothing else, so the CUDA cores only build the im2col rows: a CTA stages a 3 x 6 x 34 fp32 patch in shared
// memory (coalesced), each of its 128 threads gathers the 27 taps of one pixel into a 64-byte SWIZZLE_64B
// K-major row, one elected lane issues the MMAs, and every thread drains its accumulator row from TMEM
// (bias + LeakyReLU, fp16) with two 32-byte stores.  Several CTAs per SM overlap gather, MMA and stores.
// The patch is staged by TMA (cp.async.bulk.tensor.4d over the [B][3][H][W] image, box {40, 6, 3, 1} at (x0 - 4, y0 - 1): the image
// border is the tensor map's out-of-bounds zero fill), double-buffered behind two mbarriers, one tile ahead of the gather; the
// register-prefetched __ldg version it replaces (612 bounds-checked loads + shared-memory stores per tile) is kept as `TMA = false`.
// The FFMA version of this layer (conv_f32.cu, kept for the fp32 parity engine) needs 864 FMAs per pixel and
// was the single slowest launch of the forward pass (profiles/r01_layers_events_v2d.md).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kTileW = 32, kTileH = 4;       // 128 pixels: warp = tile row, lane = column
constexpr int kPatchW = 40;                  // 34 used columns (x0 - 1 .. x0 + 32) inside a box that starts at x0 - 4
constexpr int kPatchX = 3;                   // ... because a TMA box must start on a 16-byte boundary of the innermost dimension
constexpr int kCout = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// 16-byte chunk c (0..3) of 64-byte row r in a SWIZZLE_64B K-major tile (8-row atoms of 512 B).
__device__ __forceinline__ uint32_t swz64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ uint64_t kmajor_desc_64b(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

struct __align__(128) PatchStage { float v[3][kTileH + 2][kPatchW]; };      // sizeof rounds up to a multiple of 128: a TMA destination

// SPLIT (OM_PREC_SPLIT, the tensor-core parity mode): image values and weights as fp16 hi + lo pairs, three MMA passes per tile
// (A_lo * W_hi, A_hi * W_lo, A_hi * W_hi: corrections first, see conv_tc2.cu), weights pre-scaled by a power of two so that W_lo stays a
// normal fp16 number, output written as hi | lo halves ([.., 2 * 32] per pixel).
template <bool TMA, bool SPLIT>
__global__ void __launch_bounds__(128, (TMA && !SPLIT) ? 8 : 1) stem_tc_kernel(const __grid_constant__ CUtensorMap map_img, const float* __restrict__ img,
                                                      const float* __restrict__ w27, const float* __restrict__ bias, __half* __restrict__ out,
                                                      int batch, int h, int wd, int rows, int out_s2d, unsigned long long* trace,
                                                      FastDiv fd_tpi, FastDiv fd_tx) {
    if (threadIdx.x == 0) trace_start(trace);
    constexpr int kParts = SPLIT ? 2 : 1;                    // [0] = hi (or the plain fp16 value), [1] = lo
    __shared__ __align__(1024) uint8_t s_a[2][kParts][128 * 64];   // im2col rows, K-major SWIZZLE_64B (double-buffered)
    __shared__ __align__(1024) uint8_t s_b[kParts][kCout * 64];    // weights [cout][k], same layout
    __shared__ unsigned int s_amax;
    __shared__ PatchStage s_patch2[TMA ? 2 : 1];             // TMA: two stages, each the box {36, 6, 3} as it lands (128-byte aligned)
    __shared__ __align__(16) float s_bias[kCout];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ __align__(8) uint64_t s_pfull[2];             // TMA: patch stage filled
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // weights: w27 is [27][32] (tap-major: (ky*3+kx)*3+ci, then cout); B[n][k] = w27[k][n], k >= 27 -> 0
    float wscale = 1.0f, acc_scale = 1.0f;
    if (SPLIT) {                                             // power of two that puts max|w| in [2^13, 2^14)
        if (tid == 0) s_amax = 0u;
        __syncthreads();
        float m = 0.0f;
        for (int i = tid; i < 27 * kCout; i += 128) m = fmaxf(m, fabsf(w27[i]));
        atomicMax(&s_amax, __float_as_uint(m));
        __syncthreads();
        const float amax = __uint_as_float(s_amax);
        int sh = amax > 0.0f ? 13 - (int)floorf(log2f(amax)) : 0;
        sh = sh < -24 ? -24 : (sh > 40 ? 40 : sh);
        wscale = exp2f((float)sh);
        acc_scale = exp2f((float)-sh);
    }
    for (int i = tid; i < kCout * 4; i += 128) {
        const int n = i >> 2, c = i & 3;
        __half v[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 8 + e;
            const float w = (k < 27 ? w27[k * kCout + n] : 0.0f) * wscale;
            v[e] = __float2half(w);
            l[e] = __float2half(w - __half2float(v[e]));
        }
        *reinterpret_cast<uint4*>(s_b[0] + swz64(n, c)) = *reinterpret_cast<const uint4*>(v);
        if (SPLIT) *reinterpret_cast<uint4*>(s_b[kParts - 1] + swz64(n, c)) = *reinterpret_cast<const uint4*>(l);
    }
    if (tid < kCout) s_bias[tid] = bias[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_pfull[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_pfull[1])));
        if (TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_img));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    pdl_wait();                                        // the previous step's last kernels may still read/write our buffers
    if (threadIdx.x == 0) trace_dep(trace);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kCout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = kmajor_desc_64b(smem_u32(s_b[0]));
    const uint64_t bdesc_lo = kmajor_desc_64b(smem_u32(s_b[kParts - 1]));

    const int tiles_x = wd / kTileW, tiles_y = h / kTileH;
    const int tiles_per_image = tiles_x * tiles_y;
    // tile -> (image, tile row, tile column) by multiply-high (common.cuh): two runtime divisions per thread and tile were a quarter of
    // the kernel's instructions
    const FastDiv d_tpi = fd_tpi, d_tx = fd_tx;
    const long long total = (long long)batch * tiles_per_image;
    uint32_t phase[2] = {0, 0};
    constexpr int kPatchElems = 3 * (kTileH + 2) * 34;               // 612 = 4.78 per thread
    constexpr int kPerThread = (kPatchElems + 127) / 128;
    // patch of a tile: rows y0-1 .. y0+4, columns x0-1 .. x0+32 of the three planes (zero outside the image);
    // the loads for tile i+1 are issued before tile i is processed, so their latency hides behind its work
    // per-thread patch slots (tile-independent): plane/row/column inside the patch, offset inside the image, smem index
    int p_off[kPerThread], p_y[kPerThread], p_x[kPerThread], p_s[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
        const int i = tid + q * 128;
        const int ci = i / ((kTileH + 2) * 34);
        const int rem = i - ci * (kTileH + 2) * 34;
        const int py = rem / 34, px = rem - py * 34;
        p_y[q] = i < kPatchElems ? py - 1 : -(1 << 20);                 // out-of-range slot -> never valid
        p_x[q] = px - 1;
        p_off[q] = (ci * h + py - 1) * wd + px - 1;
        p_s[q] = (i / 34) * kPatchW + (i % 34) + kPatchX;
    }
    const int itotal = (int)total;
    // TMA: one thread asks for the whole patch of a tile; rows / columns outside the image arrive as zeros
    constexpr uint32_t kPatchBytes = 3 * (kTileH + 2) * kPatchW * 4;
    auto tma_patch = [&](int tile, int stage) {
        const int n = fdiv(tile, d_tpi);
        const int r = tile - n * tiles_per_image;
        const int ty = fdiv(r, d_tx), tx = r - ty * tiles_x;
        const uint32_t bar = smem_u32(&s_pfull[stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPatchBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(&s_patch2[stage].v[0][0][0])), "l"(&map_img), "r"(bar), "r"(tx * kTileW - (kPatchX + 1)), "r"(ty * kTileH - 1), "r"(0), "r"(n) : "memory");
    };
    auto fetch = [&](int tile, float (&regs)[kPerThread]) {
        const int n = fdiv(tile, d_tpi);
        const int r = tile - n * tiles_per_image;
        const int ty = fdiv(r, d_tx), tx = r - ty * tiles_x;
        const int y0 = ty * kTileH, x0 = tx * kTileW;
        const float* base = img + (size_t)n * 3 * h * wd + (size_t)y0 * wd + x0;
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int iy = y0 + p_y[q], ix = x0 + p_x[q];
            regs[q] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(base + p_off[q]) : 0.0f;
        }
    };
    // epilogue of the tile whose accumulator sits in TMEM stage `buf`: this thread's pixel, 32 channels
    auto epilogue = [&](int buf, int tile) {
        const int n = fdiv(tile, d_tpi);
        const int r = tile - n * tiles_per_image;
        const int ty = fdiv(r, d_tx), tx = r - ty * tiles_x;
        mbar_wait(&s_bar[buf], phase[buf]);
        phase[buf] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t acc[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]),
              "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]), "=r"(acc[15]),
              "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]),
              "=r"(acc[24]), "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 32)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int Yo = n * rows + ty * kTileH + warp, xo = tx * kTileW + lane;
        size_t opix = (size_t)Yo * wd + xo;
        if (out_s2d)                                       // parity-split output for the stride-2 consumer (om_conv_desc)
            opix = (size_t)(2 * (Yo & 1) + (xo & 1)) * ((size_t)batch * rows / 2 * (wd / 2)) + (size_t)(Yo >> 1) * (wd / 2) + (xo >> 1);
        __half* o = out + opix * (kParts * kCout);
#pragma unroll
        for (int i = 0; i < 32; i += 16) {
            uint32_t wv[8], wl[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float a = __uint_as_float(acc[i + 2 * q]), b = __uint_as_float(acc[i + 2 * q + 1]);
                if (SPLIT) mul2(a, b, a, b, acc_scale);                // power of two: exact
                const float2 bv = *reinterpret_cast<const float2*>(&s_bias[i + 2 * q]);
                add2(a, b, bv.x, bv.y);
                float ma, mb;
                mul2(ma, mb, a, b, 0.1f);                      // LeakyReLU(0.1) = max(v, 0.1 v)
                a = fmaxf(a, ma);
                b = fmaxf(b, mb);
                const __half2 hv = __floats2half2_rn(a, b);
                wv[q] = *reinterpret_cast<const uint32_t*>(&hv);
                if (SPLIT) {
                    const float2 back = __half22float2(hv);
                    const __half2 lv = __floats2half2_rn(a - back.x, b - back.y);
                    wl[q] = *reinterpret_cast<const uint32_t*>(&lv);
                }
            }
            st_global_256(o + i, wv);
            if (SPLIT) st_global_256(o + kCout + i, wl);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    float pre[kPerThread];
    if (!TMA && (int)blockIdx.x < itotal) fetch((int)blockIdx.x, pre);
    if (TMA && tid == 0 && (int)blockIdx.x < itotal) tma_patch((int)blockIdx.x, 0);
    int buf = 0;
    int prev = -1;
    int pstage = 0;
    uint32_t pphase[2] = {0, 0};
    // software pipeline: the MMA of tile i is in flight while the epilogue of tile i-1 runs
    for (int tile = blockIdx.x; tile < itotal; tile += gridDim.x) {
        float (*s_patch)[kTileH + 2][kPatchW] = s_patch2[TMA ? pstage : 0].v;
        if (TMA) {
            // 1. the next tile's patch into the other stage (its last readers -- the gather of the previous tile -- are behind the
            //    second __syncthreads of the previous iteration), then wait for this tile's
            if (tid == 0 && tile + (int)gridDim.x < itotal) tma_patch(tile + (int)gridDim.x, pstage ^ 1);
            mbar_wait(&s_pfull[pstage], pphase[pstage]);
            pphase[pstage] ^= 1;
        } else {
            // 1. stage the prefetched patch, start the next tile's loads
#pragma unroll
            for (int q = 0; q < kPerThread; ++q)
                if (tid + q * 128 < kPatchElems) (&s_patch[0][0][0])[p_s[q]] = pre[q];
            if (tile + (int)gridDim.x < itotal) fetch(tile + (int)gridDim.x, pre);
            __syncthreads();
        }
        // 2. im2col row of pixel (y0 + warp, x0 + lane): k = (ky*3 + kx)*3 + ci.  s_a[buf] was last read by the MMA of
        //    tile i-2, whose completion every thread observed in the epilogue of tile i-2.
        {
            float v[32];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) v[(ky * 3 + kx) * 3 + ci] = s_patch[ci][warp + ky][lane + kx + kPatchX];
#pragma unroll
            for (int k = 27; k < 32; ++k) v[k] = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 q, ql;
                __half2* hq = reinterpret_cast<__half2*>(&q);
                __half2* lq = reinterpret_cast<__half2*>(&ql);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hq[e] = __floats2half2_rn(v[c * 8 + 2 * e], v[c * 8 + 2 * e + 1]);
                    if (SPLIT) {
                        const float2 back = __half22float2(hq[e]);
                        lq[e] = __floats2half2_rn(v[c * 8 + 2 * e] - back.x, v[c * 8 + 2 * e + 1] - back.y);
                    }
                }
                *reinterpret_cast<uint4*>(s_a[buf][0] + swz64(tid, c)) = q;
                if (SPLIT) *reinterpret_cast<uint4*>(s_a[buf][kParts - 1] + swz64(tid, c)) = ql;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                   // also: TMEM stage `buf` was drained by every thread (epilogue i-2)
        // 3. D[128 x 32] = A[128 x 32] * B^T into TMEM stage `buf`
        if (warp == 0) {
            if (elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = kmajor_desc_64b(smem_u32(s_a[buf][0]));
                const uint64_t adesc_lo = kmajor_desc_64b(smem_u32(s_a[buf][kParts - 1]));
                // plain: one pass; split: (A_lo, W_hi), (A_hi, W_lo), (A_hi, W_hi)
#pragma unroll
                for (int pass = SPLIT ? 0 : 2; pass < 3; ++pass) {
                    const uint64_t ad = pass == 0 ? adesc_lo : adesc;
                    const uint64_t bd = pass == 1 ? bdesc_lo : bdesc;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t acc = (pass == (SPLIT ? 0 : 2) && k == 0) ? 0u : 1u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem + (uint32_t)(buf * 32)), "l"(ad + (uint64_t)(2 * k)), "l"(bd + (uint64_t)(2 * k)), "r"(idesc), "r"(acc) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[buf])) : "memory");
            }
            __syncwarp();
        }
        // 4. epilogue of the previous tile while this tile's MMA runs
        if (prev >= 0) epilogue(buf ^ 1, prev);
        prev = tile;
        buf ^= 1;
        pstage ^= 1;
    }
    if (prev >= 0) epilogue(buf ^ 1, prev);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) trace_end(trace);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

}  // namespace

namespace om {

int32_t stem_tc_run(const float* image, const float* weights, const float* bias, void* output, int batch, int h, int w, int rows,
                    int out_s2d, cudaStream_t stream, int split) {
    if (h % kTileH || w % kTileW) return fail(OM_ERR_INVALID, "tensor-core stem needs h %% 4 == 0 and w %% 32 == 0 (got %dx%d)", h, w);
    const long long tiles = (long long)batch * (h / kTileH) * (w / kTileW);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)sms * 5;               // one resident wave (96 registers x 128 threads -> 5 CTAs per SM)
    if (grid > tiles) grid = tiles;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const char* sel = getenv("ORIENMASK_B200_STEM");
    bool use_tma = !(sel && sel[0] == 'l');            // ORIENMASK_B200_STEM=ldg: the register-prefetch version (A/B)
    if (use_tma) {
        typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeTiledFn fn = nullptr;
        if (!fn) {
            void* ptr = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                fn = reinterpret_cast<EncodeTiledFn>(ptr);
        }
        if (!fn) return fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        if (reinterpret_cast<uintptr_t>(image) & 15) use_tma = false;          // TMA needs a 16-byte aligned base; the __ldg version takes anything
        else {
            // the caller's fp32 NCHW image as a 4-d tensor [B][3][H][W]; box = one tile's patch; out-of-bounds elements read as zero
            cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, 3, (cuuint64_t)batch};
            cuuint64_t str[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)3 * h * w * 4};
            cuuint32_t box[4] = {(cuuint32_t)kPatchW, (cuuint32_t)(kTileH + 2), 3, 1};
            cuuint32_t ones[4] = {1, 1, 1, 1};
            CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(image), dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled(stem image) failed with CUresult %d", (int)r);
        }
    }
    if (use_tma) {
        // 58 registers and 23.8 KB of shared memory per CTA: up to 8 CTAs per SM (8 x 64 TMEM columns = the whole TMEM)
        const char* ce = getenv("ORIENMASK_B200_STEM_CTAS");
        const int per_sm = (ce && atoi(ce) >= 1 && atoi(ce) <= 8) ? atoi(ce) : (split ? 5 : 8);     // split: 42 KB of shared memory per CTA
        grid = (long long)sms * per_sm;
        if (grid > tiles) grid = tiles;
    }
    __half* o = reinterpret_cast<__half*>(output);
    unsigned long long* tr = trace_next();
    if (split && !use_tma) return fail(OM_ERR_UNSUPPORTED, "the split-precision tensor-core stem needs a 16-byte aligned image (TMA)");
    if (split)
        OM_CUDA_TRY(launch_pdl(stem_tc_kernel<true, true>, dim3((unsigned)grid), dim3(128), 0, stream, map, image, weights, bias, o, batch, h, w, rows, out_s2d, tr, make_fastdiv((h / kTileH) * (w / kTileW)), make_fastdiv(w / kTileW)));
    else if (use_tma)
        OM_CUDA_TRY(launch_pdl(stem_tc_kernel<true, false>, dim3((unsigned)grid), dim3(128), 0, stream, map, image, weights, bias, o, batch, h, w, rows, out_s2d, tr, make_fastdiv((h / kTileH) * (w / kTileW)), make_fastdiv(w / kTileW)));
    else
        OM_CUDA_TRY(launch_pdl(stem_tc_kernel<false, false>, dim3((unsigned)grid), dim3(128), 0, stream, map, image, weights, bias, o, batch, h, w, rows, out_s2d, tr, make_fastdiv((h / kTileH) * (w / kTileW)), make_fastdiv(w / kTileW)));
    return check_launch("stem_tc_kernel");
}

}  // namespace om
