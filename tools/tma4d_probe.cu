// Probe: which 4-d tiled TMA box configurations over an fp32 [B][3][H][W] image load correctly on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/tma4d_probe tools/tma4d_probe.cu -lcuda && tools/_bin/tma4d_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int bw, int bh, int bc, int x, int y, int c, int n, int align_off, int prefetch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* dst = reinterpret_cast<float*>(smem + align_off);
    const unsigned bytes = (unsigned)(bw * bh * bc * 4);
    if (threadIdx.x == 0) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(&map));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(&map), "r"(b), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
    }
    __syncthreads();
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    long long t0 = clock64();
    while (true) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        if (ok) break;
        if (clock64() - t0 > 200000000ll) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    }
    for (int i = threadIdx.x; i < bw * bh * bc; i += blockDim.x) out[i] = dst[i];
}

int main() {
    const int B = 2, H = 64, W = 96;
    std::vector<float> h((size_t)B * 3 * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) + 1.0f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 16);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn fn = (Fn)p;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    // finding (B200, driver 580): the innermost start coordinate must keep the row start 16-byte aligned -- x = -1 (fp32) faults with
    // "illegal instruction", x = -4 is fine; negative / out-of-range coordinates in the outer dimensions are fine (zero fill)
    struct Case { int bw, bh, bc, x, y, c, n, off, pf; } cases[] = {
        {32, 6, 3, 0, 0, 0, 0, 0, 0}, {40, 6, 3, -4, -1, 0, 1, 0, 0}, {40, 6, 3, 28, 3, 0, 1, 0, 1}, {40, 6, 3, 60, 59, 0, 1, 2944, 0},
        {40, 6, 3, 92, 61, 0, 0, 0, 0}, {32, 6, 3, -1, -1, 0, 1, 0, 0}};
    for (const Case& k : cases) {
        CUtensorMap map; memset(&map, 0, sizeof(map));
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)k.bw, (cuuint32_t)k.bh, (cuuint32_t)k.bc, 1};
        cuuint32_t ones[4] = {1, 1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemset(o, 0, 1 << 16);
        probe<<<1, 128, 32768>>>(map, o, k.bw, k.bh, k.bc, k.x, k.y, k.c, k.n, k.off, k.pf);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> got((size_t)k.bw * k.bh * k.bc);
        int bad = 0;
        if (e == cudaSuccess) {
            cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
            for (int c = 0; c < k.bc; ++c) for (int yy = 0; yy < k.bh; ++yy) for (int xx = 0; xx < k.bw; ++xx) {
                const int gx = k.x + xx, gy = k.y + yy;
                const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(((size_t)k.n * 3 + c) * H + gy) * W + gx] : 0.0f;
                if (got[((size_t)c * k.bh + yy) * k.bw + xx] != want) ++bad;
            }
        }
        printf("box {%d,%d,%d} at (%d,%d,%d,%d) smem+%d prefetch %d: encode %d, run %s, mismatches %d%s\n", k.bw, k.bh, k.bc, k.x, k.y, k.c, k.n, k.off, k.pf, (int)r,
               cudaGetErrorString(e), bad, (e == cudaSuccess && got[0] == -12345.f) ? " (TIMEOUT)" : "");
        if (e != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 16); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
                                cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); }
    }
    return 0;
}
