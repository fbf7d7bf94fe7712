// First TWO layers in one launch: backbone.conv1 (3 -> 32, 3x3, stride 1) and backbone.conv2.0 (32 -> 64, 3x3, stride 2), BN folded,
// LeakyReLU after each (model/backbone/darknet.py:41-43, model/base.py:113-128), straight from the caller's fp32 NCHW image to the fp16
// padded-row NHWC activation at half resolution.
//
// Why: unfused, the stem writes its 32-channel full-resolution output (606 MB at bs 32 / 544x544, the largest tensor of the forward)
// and the stride-2 layer reads it straight back: 1.2 GB of the ~1.6 GB these two launches move, 217 + 193 us in the pipelined forward
// (tools/timeline.py) for layers whose inputs and outputs proper are 114 MB + 303 MB.  Fused, the stem output of a tile never leaves
// the SM.
//
// One tile = 8 x 16 output pixels of the stride-2 layer = a 17 x 33 region of stem pixels = a 19 x 35 patch of the image:
//   1. TMA: the fp32 patch, box {24, 35, 3, 1} of the [B][3][H][W] image at (2*x0 - 4, 2*y0 - 2) (a TMA box must start on a 16-byte
//      boundary of its innermost dimension; rows / columns outside the image arrive as zeros), double-buffered one tile ahead;
//   2. CUDA cores: the im2col rows of the 561 stem pixels (27 taps -> K = 32, fp16, 64-byte SWIZZLE_64B K-major rows), row r = the
//      entry of the parity-plane layout of step 4 the pixel will occupy; a thread gathers two adjacent pixels from aligned 8-byte loads;
//   3. MMA 1: five M = 128 tcgen05 MMAs x 2 K-steps against the resident stem weights -> TMEM (5 x 32 columns);
//   4. epilogue 1: + b1, LeakyReLU, ZERO where the stem pixel lies outside the image (the stride-2 layer pads the stem OUTPUT), fp16,
//      written over the dead im2col rows in exactly the layout conv_tc2.cu's parity-plane halo mode consumes: four (row parity,
//      column parity) planes of 9 x 17 pixels x 64 bytes, SWIZZLE_64B;
//   5. MMA 2: nine taps = nine shifted UMMA descriptors (plane 2*(r != 1) + (s != 1), offset ((r != 0), (s != 0)) inside it, stride-byte
//      offset = one plane row of 9 pixels), K = 32 per tap, resident W2 -> [128 x 64] in TMEM (over the drained columns of MMA 1);
//   6. epilogue 2: + b2, LeakyReLU, fp16, two 32-byte stores per thread.
// Same operands and the same accumulation order as the two launches it replaces (one K = 32 chain per stem pixel; tap-major K loop
// of the stride-2 layer), so the result is bit-identical to them: tests/test_gpu_forward.py::test_c_engine_matches_the_python_schedule
// compares the C engine (this kernel) with the Python-scheduled twin (stem_tc_kernel + conv_tc2_kernel).
//
// Like dark_block.cu a tile runs its phases in sequence on a STREAM of 256 threads; one CTA per SM carries three streams that share the
// resident weights and own their patch stages, operand tile, TMEM columns and barriers, so the phases of different tiles overlap.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int kC1 = 32, kC2 = 64;            // stem / stride-2 layer output channels
constexpr int kTW = 8, kTH = 16;             // output tile of the stride-2 layer (128 pixels = one M = 128 MMA)
constexpr int kSH = 2 * kTH + 1;             // stem region: 17 columns x 33 rows
constexpr int kM1 = 5;                       // MMA 1: 5 x (M = 128) = 640 rows = 4 parity planes x 160 entries (153 used)
constexpr int kPW = 24, kPH = kSH + 2;       // image patch: 24 columns from 2*x0 - 4 (19 used, from 2*x0 - 2), 35 rows from 2*y0 - 2
constexpr int kPX = 2;                       // first used patch column
constexpr int kPatchBytes = 3 * kPH * kPW * 4;                    // 10080
constexpr int kPatchStage = ((kPatchBytes + 1023) / 1024) * 1024;  // 10240
constexpr int kPlaneW = kTW + 1, kPlaneH = kTH + 1;               // parity plane box 9 x 17
constexpr int kPlane = ((kPlaneW * kPlaneH * 64 + 1023) / 1024) * 1024;   // 10240 B
constexpr int kABytes = kM1 * 128 * 64;      // 40960: im2col rows, then (aliased) the four planes of the stem output
static_assert(4 * kPlane <= kABytes, "the stem output tile overwrites the im2col rows");
constexpr int kStreams = 3;
constexpr int kStreamThreads = 256;
constexpr int kThreads = kStreams * kStreamThreads;
constexpr int kStreamBytes = 2 * kPatchStage + kABytes;            // 61440
constexpr int kW1Bytes = kC1 * 64;           // [32 rows (cout)][32 k] fp16, SWIZZLE_64B
constexpr int kW2Bytes = 9 * kC2 * 64;       // [9][64 rows (cout)][32 k] fp16, SWIZZLE_64B
constexpr int kTmemCols = kM1 * kC1;         // 160 per stream
static_assert(kStreams * kTmemCols <= 512, "TMEM columns");
constexpr int kSmem = 1024 + kStreams * kStreamBytes + kW1Bytes + kW2Bytes + 1024;
static_assert(kSmem <= 227 * 1024, "shared memory");

struct FusedParams {
    FastDiv d_tpi, d_tx;                     // tiles per image, tiles per tile row
    int tiles_x, tiles_per_image, total;     // per image: (W/2 / 8) x (H/2 / 16)
    int h, w;                                // image size (stem resolution)
    int out_rows;                            // rows per image of the padded-row output (half resolution)
    const float* w27; const float* b1;       // stem weights [27][32] fp32 (tap-major, BN folded) and bias
    const __half* w2; const float* b2;       // engine layout of OM_PREC_F16: [9][64][32]
    __half* out;
    unsigned long long* trace;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    const long long t0 = clock64();
    for (uint32_t spins = 1;; ++spins) {               // bounded: a protocol bug traps instead of hanging the GPU
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if ((spins & 63u) == 0u && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, uint32_t sbo_bytes) {   // K-major, 64-byte rows
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void stream_sync(int stream) {                 // named barrier of one stream (0 is __syncthreads)
    asm volatile("bar.sync %0, %1;" ::"r"(stream + 1), "r"(kStreamThreads) : "memory");
}
// 16-byte chunk c (0..3) of 64-byte row r in a SWIZZLE_64B K-major tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// Debug phase log (om_debug_phase_log): stream 0 of CTA 0 stamps %globaltimer at the phase boundaries of its first tiles.
__device__ unsigned long long* g_phase_log = nullptr;
__device__ __forceinline__ void phase_tick(int tile_no, int slot, bool who) {
    if (g_phase_log != nullptr && who && tile_no < 16) g_phase_log[tile_no * 8 + slot] = trace_now();
}

__global__ void __launch_bounds__(kThreads, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap map_img, const FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (threadIdx.x == 0) trace_start(p.trace);
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stream = threadIdx.x / kStreamThreads;
    uint8_t* s_patch = smem + stream * kStreamBytes;     // [2][kPatchStage]  fp32 [3][35][24]                    (per stream)
    uint8_t* s_a = s_patch + 2 * kPatchStage;            // [kABytes]         im2col rows, then the stem output planes (per stream)
    uint8_t* s_w1 = smem + kStreams * kStreamBytes;      // [32][64 B]                                            (shared)
    uint8_t* s_w2 = s_w1 + kW1Bytes;                     // [9][64][64 B]
    float* s_b1 = reinterpret_cast<float*>(s_w2 + kW2Bytes);      // [32]
    float* s_b2 = s_b1 + kC1;                                      // [64]
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(s_b2 + kC2);  // per stream: p_full[2], d1, d2
    uint64_t* bars = bars_all + 4 * stream;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_all + 4 * kStreams);
    const int tid = threadIdx.x % kStreamThreads, warp = tid >> 5, lane = tid & 31;    // position inside the stream

    // ---- prologue (overlaps the previous launch's tail under PDL): weights, barriers, TMEM ----
    for (int i = threadIdx.x; i < kC1 * 4; i += kThreads) {              // W1: B[n][k] = w27[k][n], k >= 27 -> 0
        const int n = i >> 2, c = i & 3;
        __half v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 8 + e;
            v[e] = __float2half(k < 27 ? p.w27[k * kC1 + n] : 0.0f);
        }
        *reinterpret_cast<uint4*>(s_w1 + swz64(n, c)) = *reinterpret_cast<const uint4*>(v);
    }
    for (int i = threadIdx.x; i < 9 * kC2 * 4; i += kThreads) {          // W2: row R = tap * 64 + n, chunk c
        const int R = i >> 2, c = i & 3;
        *reinterpret_cast<uint4*>(s_w2 + swz64(R, c)) = __ldg(reinterpret_cast<const uint4*>(p.w2 + R * 32 + c * 8));
    }
    if (threadIdx.x < kC1) s_b1[threadIdx.x] = p.b1[threadIdx.x];
    if (threadIdx.x < kC2) s_b2[threadIdx.x] = p.b2[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * kStreams; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars_all[i])));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_img));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the weight tiles above are read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)(stream * kTmemCols);
    uint64_t* p_full = bars;
    uint64_t* bar_d1 = bars + 2;
    uint64_t* bar_d2 = bars + 3;
    pdl_wait();                                          // the image is the previous launch's output; our output is read by the previous forward
    if (threadIdx.x == 0) trace_dep(p.trace);

    auto decode = [&](int tile, int& n, int& ty, int& tx) {
        n = fdiv(tile, p.d_tpi);
        const int r = tile - n * p.tiles_per_image;
        ty = fdiv(r, p.d_tx);
        tx = r - ty * p.tiles_x;
    };
    auto load_patch = [&](int tile, int stage) {
        int n, ty, tx;
        decode(tile, n, ty, tx);
        const uint32_t bar = smem_u32(&p_full[stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kPatchBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(s_patch + stage * kPatchStage)), "l"(&map_img), "r"(bar), "r"(2 * tx * kTW - 4), "r"(2 * ty * kTH - 2), "r"(0), "r"(n) : "memory");
    };
    // UMMA instruction descriptors: c_format F32 (bit 4), a/b F16, K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(kC1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(kC2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t quad = warp & 3, half = warp >> 2;

    const int first = (int)blockIdx.x * kStreams + stream, step = (int)gridDim.x * kStreams;
    if (tid == 0 && first < p.total) load_patch(first, 0);
    int stage = 0;
    uint32_t pphase[2] = {0, 0}, d1phase = 0, d2phase = 0;
    const bool logger = blockIdx.x == 0 && threadIdx.x == 0;
    int tile_no = -1;
    for (int tile = first; tile < p.total; tile += step) {
        int n, ty, tx;
        decode(tile, n, ty, tx);
        ++tile_no;
        phase_tick(tile_no, 0, logger);
        // the next tile's patch into the other stage: its last readers (the gather of the previous tile) are behind the barriers that
        // ended the previous iteration
        if (tid == 0 && tile + step < p.total) load_patch(tile + step, stage ^ 1);
        mbar_wait(&p_full[stage], pphase[stage]);
        pphase[stage] ^= 1;
        phase_tick(tile_no, 1, logger);
        // ---- im2col rows.  Row r = plane * 160 + pr: the stem pixel that ends up at entry pr = jj * 9 + ii of parity plane `plane`
        // (exactly where epilogue 1 will write it, so MMA 1 row r <-> plane row r), k = (ky*3 + kx)*3 + ci.  A thread gathers a PAIR of
        // horizontally adjacent stem pixels -- the odd-column one at entry ii and the even-column one at entry ii + 1 of the same stem
        // row -- from two aligned 8-byte loads per (ky, ci): columns X-1 .. X+2 serve both (a stride-2 scalar gather would waste half
        // of every shared-memory wavefront; the LSU data pipe, not the tensor pipe or HBM, bounds this kernel). ----
        {
            const float* pt = reinterpret_cast<const float*>(s_patch + stage * kPatchStage);
#pragma unroll 1
            for (int it = tid; it < kSH * kPlaneW; it += kStreamThreads) {   // 33 stem rows x 9 pair slots
                const int j = it / kPlaneW, ii = it - j * kPlaneW;          // stem row j of the region, odd-column entry ii
                const int prow = 1 - (j & 1), jj = (j + 1) >> 1;            // row parity of the stem row, its entry row in the planes
                const int r_odd = (2 * prow + 1) * 160 + jj * kPlaneW + ii;      // X odd:  plane 2*prow + 1, entry ii
                const int r_even = (2 * prow) * 160 + jj * kPlaneW + ii + 1;     // X even: plane 2*prow,     entry ii + 1 (ii + 1 <= 8)
                const float* base = pt + j * kPW + 2 * ii + kPX;            // patch row j + ky; columns 2*ii + 2 .. + 5 = X-1 .. X+2
                float va[32], vb[32];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        const float2 lo = *reinterpret_cast<const float2*>(base + (ci * kPH + ky) * kPW);
                        const float2 hi = *reinterpret_cast<const float2*>(base + (ci * kPH + ky) * kPW + 2);
                        va[(ky * 3 + 0) * 3 + ci] = lo.x; va[(ky * 3 + 1) * 3 + ci] = lo.y; va[(ky * 3 + 2) * 3 + ci] = hi.x;
                        vb[(ky * 3 + 0) * 3 + ci] = lo.y; vb[(ky * 3 + 1) * 3 + ci] = hi.x; vb[(ky * 3 + 2) * 3 + ci] = hi.y;
                    }
#pragma unroll
                for (int k = 27; k < 32; ++k) { va[k] = 0.0f; vb[k] = 0.0f; }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 q;
                    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
                    for (int e = 0; e < 4; ++e) hq[e] = __floats2half2_rn(va[c * 8 + 2 * e], va[c * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(s_a + swz64(r_odd, c)) = q;
                }
                if (ii + 1 < kPlaneW) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 q;
                        __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
                        for (int e = 0; e < 4; ++e) hq[e] = __floats2half2_rn(vb[c * 8 + 2 * e], vb[c * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(s_a + swz64(r_even, c)) = q;
                    }
                }
            }
            // entries no stem pixel maps to (row 0 of the even-row planes, column 0 of the even-column planes, the 7 rows past each
            // plane) keep whatever the previous tile left there: finite fp16 values whose MMA 1 results are never used, and MMA 2 never
            // reads them
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        stream_sync(stream);
        phase_tick(tile_no, 2, logger);
        // ---- MMA 1: stem pixels [640 x 32] = rows [640 x 32] * W1^T -> TMEM columns g*32 .. +31 for rows g*128 .. +127 ----
        if (warp == 0) {
            if (elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t as = smem_u32(s_a);
                const uint64_t bd = desc_sw64(smem_u32(s_w1), 512u);
#pragma unroll
                for (int g = 0; g < kM1; ++g) {
                    const uint64_t ad = desc_sw64(as + (uint32_t)(g * 128 * 64), 512u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma(tmem + (uint32_t)(g * kC1), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, k != 0);
                }
                umma_commit(bar_d1);
            }
            __syncwarp();
        }
        mbar_wait(bar_d1, d1phase);
        d1phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        phase_tick(tile_no, 3, logger);
        // ---- epilogue 1: row r = g*128 + quad*32 + lane = plane * 160 + pr -> 64 bytes at entry pr of its parity plane, i.e. over the
        // im2col row it came from.  plane = 2*(row odd) + (column odd); entry (jj, ii) of a plane is the stem pixel
        // (2*(y0 - 1 + jj) + row parity, 2*(x0 - 1 + ii) + column parity), y0 / x0 the tile origin in output pixels. ----
        for (int g = (int)half; g < kM1; g += 2) {
            const int r = g * 128 + (int)quad * 32 + lane;
            uint32_t v[32];
            tmem_ld32(tmem + ((quad * 32u) << 16) + (uint32_t)(g * kC1), v);
            const int plane = r / 160, pr = r - plane * 160;
            if (pr < kPlaneW * kPlaneH) {
                const int jj = pr / kPlaneW, ii = pr - jj * kPlaneW;
                const int Y = 2 * (ty * kTH - 1 + jj) + (plane >> 1), X = 2 * (tx * kTW - 1 + ii) + (plane & 1);
                const bool inside = Y >= 0 && Y < p.h && X >= 0 && X < p.w;
                uint32_t w[16];
#pragma unroll
                for (int q = 0; q < 16; q += 2) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_b1 + 2 * q);          // warp-uniform: smem broadcast
                    float a = __uint_as_float(v[2 * q]), b = __uint_as_float(v[2 * q + 1]);
                    float c = __uint_as_float(v[2 * q + 2]), d = __uint_as_float(v[2 * q + 3]);
                    add2(a, b, bv.x, bv.y); add2(c, d, bv.z, bv.w);
                    float ma, mb, mc, md;
                    mul2(ma, mb, a, b, 0.1f); mul2(mc, md, c, d, 0.1f);
                    a = fmaxf(a, ma); b = fmaxf(b, mb); c = fmaxf(c, mc); d = fmaxf(d, md);
                    const __half2 h0 = inside ? __floats2half2_rn(a, b) : __floats2half2_rn(0.0f, 0.0f);
                    const __half2 h1 = inside ? __floats2half2_rn(c, d) : __floats2half2_rn(0.0f, 0.0f);
                    w[q] = *reinterpret_cast<const uint32_t*>(&h0);
                    w[q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(s_a + swz64(r, c)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        stream_sync(stream);                             // every accumulator of MMA 1 is drained, every plane is written
        phase_tick(tile_no, 4, logger);
        // ---- MMA 2: nine taps over the four planes (shifted descriptors), resident W2 -> TMEM columns 0 .. 63 ----
        if (warp == 0) {
            if (elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ys = smem_u32(s_a), ws = smem_u32(s_w2);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int r = tap / 3, s = tap % 3;
                    const uint32_t off = (uint32_t)((2 * (r != 1) + (s != 1)) * kPlane + ((r != 0) * kPlaneW + (s != 0)) * 64);
                    const uint64_t ad = desc_sw64(ys + off, kPlaneW * 64u);
                    const uint64_t bd = desc_sw64(ws + (uint32_t)tap * (kC2 * 64), 512u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc2, (tap | k) != 0);
                }
                umma_commit(bar_d2);
            }
            __syncwarp();
        }
        mbar_wait(bar_d2, d2phase);
        d2phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        phase_tick(tile_no, 5, logger);
        // ---- epilogue 2: output pixel m = quad*32 + lane, channels half*32 .. +31: + b2, LeakyReLU, fp16, store ----
        {
            const int m = (int)(quad * 32) + lane;
            const int oy = m / kTW, ox = m - oy * kTW;
            uint32_t v[32];
            tmem_ld32(tmem + ((quad * 32u) << 16) + half * 32u, v);
            const size_t opix = ((size_t)n * p.out_rows + (size_t)(ty * kTH + oy)) * (size_t)(p.w >> 1) + (size_t)(tx * kTW + ox);
            __half* o = p.out + opix * kC2 + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 16) {
                uint32_t wv[8];
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const int ch = i + 2 * q;
                    const float4 bv = *reinterpret_cast<const float4*>(s_b2 + half * 32 + ch);
                    float a = __uint_as_float(v[ch]), b = __uint_as_float(v[ch + 1]);
                    float c = __uint_as_float(v[ch + 2]), d = __uint_as_float(v[ch + 3]);
                    add2(a, b, bv.x, bv.y); add2(c, d, bv.z, bv.w);
                    float ma, mb, mc, md;
                    mul2(ma, mb, a, b, 0.1f); mul2(mc, md, c, d, 0.1f);
                    a = fmaxf(a, ma); b = fmaxf(b, mb); c = fmaxf(c, mc); d = fmaxf(d, md);
                    const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
                    wv[q] = *reinterpret_cast<const uint32_t*>(&h0);
                    wv[q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                }
                st_global_256(o + i, wv);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        stream_sync(stream);                             // patch[stage], the operand tile and the accumulators of this stream are free again
        phase_tick(tile_no, 6, logger);
        stage ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) trace_end(p.trace);
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace

namespace om {

int32_t stem_fused_set_phase_log(void* dev_ptr) {
    OM_CUDA_TRY(cudaMemcpyToSymbol(g_phase_log, &dev_ptr, sizeof(void*)));
    return OM_OK;
}

bool stem_fused_supported(int h, int w, const void* image) {
    const char* e = getenv("ORIENMASK_B200_FUSED_STEM");
    if (e && e[0] == '0') return false;
    return h % (2 * kTH) == 0 && w % (2 * kTW) == 0 && w % 4 == 0 && (image == nullptr || (reinterpret_cast<uintptr_t>(image) & 15) == 0);
}

// image [batch, 3, h, w] fp32 -> out [batch * out_rows, w / 2, 64] fp16 padded-row NHWC (rows h / 2 .. out_rows - 1 of an image are never
// written: they stay zero).  w27 / b1: the stem's folded fp32 weights [27][32] and bias; w2 / b2: the stride-2 layer's OM_PREC_F16
// weights [9][64][32] and bias.
int32_t stem_fused_run(const float* image, const float* w27, const float* b1, const void* w2, const float* b2, void* out, int batch, int h,
                       int w, int out_rows, cudaStream_t stream) {
    if (!stem_fused_supported(h, w, image)) return fail(OM_ERR_UNSUPPORTED, "fused stem needs h %% 32 == 0, w %% 16 == 0 and a 16-byte aligned image");
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
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, 3, (cuuint64_t)batch};
    cuuint64_t str[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)3 * h * w * 4};
    cuuint32_t box[4] = {(cuuint32_t)kPW, (cuuint32_t)kPH, 3, 1};
    cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(image), dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled(fused stem image) failed with CUresult %d", (int)r);
    FusedParams p;
    p.tiles_x = (w / 2) / kTW;
    p.tiles_per_image = p.tiles_x * ((h / 2) / kTH);
    p.total = batch * p.tiles_per_image;
    p.d_tpi = make_fastdiv(p.tiles_per_image); p.d_tx = make_fastdiv(p.tiles_x);
    p.h = h; p.w = w; p.out_rows = out_rows;
    p.w27 = w27; p.b1 = b1; p.w2 = reinterpret_cast<const __half*>(w2); p.b2 = b2;
    p.out = reinterpret_cast<__half*>(out);
    p.trace = trace_next();
    OM_CUDA_TRY(cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = sms;                                    // one CTA of kStreams tile streams per SM
    if (grid * kStreams > p.total) grid = (p.total + kStreams - 1) / kStreams;
    OM_CUDA_TRY(launch_pdl(stem_fused_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmem, stream, map, p));
    return check_launch("stem_fused_kernel");
}

}  // namespace om
