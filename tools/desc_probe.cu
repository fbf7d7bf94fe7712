// Probe: does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor accept a start address that is a
// multiple of 128 B but not of 1024 B (a pixel-shifted view of a TMA-swizzled halo tile), and with which
// base_offset / stride-byte-offset semantics?  Decides whether one halo tile in shared memory can feed all
// nine taps of a 3x3 convolution.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/desc_probe tools/desc_probe.cu && tools/_bin/desc_probe
//
// Shared memory holds P pixels x 64 fp16 channels (128 B per pixel) written exactly as TMA SWIZZLE_128B
// writes them: 16-byte chunk c of pixel p lives at p*128 + ((c ^ (p & 7)) << 4) from a 1024-aligned base.
// Value(p, c*8+e) = (p % 32) * 64 + c*8+e (exact in fp16).  B = 64x64 identity, so D[m][n] = A[m][n] as the
// tensor core saw it.  Expectation for a working shifted view: D[m][n] = value(shift + (m/8)*pitch + m%8, n).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kPixels = 512;

__global__ void __launch_bounds__(128, 1) probe(float* out, int shift, int pitch, int use_base_offset) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __half* A = reinterpret_cast<__half*>(smem);                         // kPixels * 128 B
    __half* B = reinterpret_cast<__half*>(smem + kPixels * 128);         // 64 rows x 128 B, swizzled the same way
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kPixels * 128 + 64 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kPixels * 8; i += 128) {
        const int p = i >> 3, c = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((float)((p % 32) * 64 + c * 8 + e));
        *reinterpret_cast<uint4*>(smem + p * 128 + ((c ^ (p & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((c * 8 + e) == n ? 1.0f : 0.0f);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(B) + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(A) + (uint32_t)shift * 128u;
        const uint32_t b_addr = smem_u32(B);
        const uint64_t sbo_a = (uint64_t)(pitch * 128) >> 4, sbo_b = 1024 >> 4;
        const uint64_t boff = use_base_offset ? (uint64_t)((a_addr >> 7) & 7) : 0;
        const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo_a << 32) | (1ull << 46) | (boff << 49) | (2ull << 61);
        const uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo_b << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k != 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
            if (clock64() - t0 > 2000000000ll) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) out[tid * 64 + c0 + e] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * sizeof(float));
    static float h[128 * 64];
    const size_t smem = 1024 + kPixels * 128 + 64 * 128 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int pitches[] = {8, 10, 16, 18};
    for (int pi = 0; pi < 4; ++pi)
        for (int bo = 0; bo < 2; ++bo)
            for (int shift = 0; shift < 12; ++shift) {
                const int pitch = pitches[pi];
                cudaMemset(d, 0xff, sizeof(h));
                probe<<<1, 128, smem>>>(d, shift, pitch, bo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pitch %d base_off %d shift %d: CUDA error %s\n", pitch, bo, shift, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                int bad = 0, first_bad = -1;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 64; ++n) {
                        const int p = shift + (m / 8) * pitch + (m % 8);
                        const float want = (float)((p % 32) * 64 + n);
                        if (h[m * 64 + n] != want) { ++bad; if (first_bad < 0) first_bad = m * 64 + n; }
                    }
                printf("pitch %2d base_off %d shift %2d: %s", pitch, bo, shift, bad ? "MISMATCH" : "ok");
                if (bad) {
                    const int m = first_bad / 64;
                    printf(" (%d bad; row %d got pixel/chan:", bad, m);
                    for (int n = 0; n < 64; n += 8) { const int v = (int)h[m * 64 + n]; printf(" %d/%d", v / 64, v % 64); }
                    printf(")");
                }
                printf("\n");
            }
    return 0;
}
