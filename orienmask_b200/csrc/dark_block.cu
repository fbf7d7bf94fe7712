// Fused DarkNet residual block (model/backbone/darknet.py:6-15) for the first stage, C = 32:
//
//     out = x + leaky(bn(conv3x3(leaky(bn(conv1x1(x))))))          x: 64 channels, the 1x1 squeezes to 32, the 3x3 expands to 64
//
// in ONE launch per block instead of two (`backbone.conv2.1.conv.{0,1}` at 272x272 for a 544x544 input: the two most memory-bound
// launches of the forward after the stem).  Unfused, a pixel of this block moves x (128 B read), y (64 B written, 64 B re-read 9x through
// L2/halo), x again as the residual (128 B) and the output (128 B); fused it moves x once and the output once:
//
//   1. TMA:  the x halo of an 8 x 16 output tile, box {64 ch, 10, 18} -> 180 rows of 128 B in SWIZZLE_128B shared memory (the image
//            border is the tensor map's out-of-bounds zero fill; rows between images are zero in memory -- "padded-row NHWC");
//   2. MMA1: y_halo[256 x 32] = x_halo[256 x 64] * W1^T as two M = 128 tcgen05 MMAs x 4 K-steps (rows 180..255 read past the box: their
//            results are never used); accumulators in TMEM;
//   3. epilogue 1: + b1, LeakyReLU, ZERO where the halo pixel lies outside the image (the 3x3 pads y, not x), fp16, written to shared
//            memory in exactly the SWIZZLE_64B halo layout the 3x3 of conv_tc2.cu consumes (pixel = 64 B, halo row = 10 pixels);
//   4. MMA2: nine taps = nine shifted UMMA descriptors over that tile (start + (r * 10 + s) pixels, stride-byte-offset = one halo row:
//            the tensor core applies the swizzle to absolute shared-memory address bits), K = 32 per tap, resident W2 -> [128 x 64];
//   5. epilogue 2: + b2, LeakyReLU, + x (the residual, read from the x halo still in shared memory), fp16, 2 x 32-byte stores.
//
// Same arithmetic, same order as the two unfused launches (tap-major K loop, fp32 bias add, leaky, fp32 residual add, one rounding), so
// the result is bit-identical to them -- tests/test_gpu_forward.py::test_c_engine_matches_the_python_schedule compares the C engine (this
// kernel) with the Python-scheduled twin (two conv_tc2 launches) head for head.
//
// A tile runs its phases in sequence on a STREAM of 256 threads (8 warps); one CTA per SM carries kStreams = 3 independent streams that
// share the resident weights (40 KB) and own their x stages (2 x 23 KB), y tile (12 KB), 128 TMEM columns, mbarriers and a named barrier
// each -- the phases of different tiles overlap across streams, and the x halo of a stream's next tile is in flight during its current
// one.  (Two separate CTAs of one stream each -- the first version -- duplicate the weights: 100 KB per CTA, 2 per SM: 242 us; three
// streams: see DESIGN.md.)  Only C = 32 fits this form: with C = 64 (stage conv3) the resident W2 alone is 147 KB.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int kC = 32;                       // squeezed channels; the block's input / output have 2 * kC
constexpr int kTW = 8, kTH = 16;             // output tile (128 pixels = one M = 128 MMA)
constexpr int kHW = kTW + 2, kHH = kTH + 2;  // halo 10 x 18 = 180 pixels
constexpr int kHalo = kHW * kHH;
#ifndef OM_BLOCK_STREAMS
#define OM_BLOCK_STREAMS 3
#endif
#ifndef OM_BLOCK_XSTAGES
#define OM_BLOCK_XSTAGES 2
#endif
constexpr int kStreams = OM_BLOCK_STREAMS;   // independent tile pipelines per CTA
constexpr int kXStages = OM_BLOCK_XSTAGES;   // x halo stages per stream (1: the next halo is requested after epilogue 2)
constexpr int kStreamThreads = 256;
constexpr int kThreads = kStreams * kStreamThreads;
constexpr int kXStage = ((kHalo * 128 + 1023) / 1024) * 1024;      // 23552 B
constexpr int kYBytes = ((kHalo * 64 + 1023) / 1024) * 1024;       // 12288 B
constexpr int kW1Bytes = kC * 128;                                 // [32 rows (cout)][64 k] fp16, SWIZZLE_128B
constexpr int kW2Bytes = 9 * 2 * kC * 64;                          // [9][64 rows (cout)][32 k] fp16, SWIZZLE_64B
constexpr int kStreamBytes = kXStages * kXStage + kYBytes;
constexpr int kSmem = 1024 + kStreams * kStreamBytes + kW1Bytes + kW2Bytes + 1024;

struct BlockParams {
    FastDiv d_tiles_x, d_rows;
    int tiles_x, tiles_y;
    int width, height, rows, total_rows;     // output == input geometry (stride 1)
    int out_s2d;
    long long s2d_plane;
    const __half* w1; const __half* w2;      // engine layout of OM_PREC_F16: [1][32][64] and [9][64][32]
    const float* b1; const float* b2;
    __half* out;
    unsigned long long* trace;               // om_debug_trace record of this launch, or nullptr
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
        uint32_t ok;                                   // fast path: no clock reads when the phase has already completed
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    const long long t0 = clock64();
    for (uint32_t spins = 1;; ++spins) {               // suspend-time hint: the hardware parks the warp instead of re-issuing the poll
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if ((spins & 63u) == 0u && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {          // K-major, 128-byte rows, dense 8-row groups
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
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

__global__ void __launch_bounds__(kThreads, 1)
dark_block_kernel(const __grid_constant__ CUtensorMap map_x, const BlockParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (threadIdx.x == 0) trace_start(p.trace);
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stream = threadIdx.x / kStreamThreads;
    uint8_t* s_x = smem + stream * kStreamBytes;         // [2][kXStage]   x halo, SWIZZLE_128B rows of 128 B   (per stream)
    uint8_t* s_y = s_x + kXStages * kXStage;             // [kYBytes]      y halo, SWIZZLE_64B rows of 64 B     (per stream)
    uint8_t* s_w1 = smem + kStreams * kStreamBytes;      // [32][128 B]                                         (shared)
    uint8_t* s_w2 = s_w1 + kW1Bytes;                     // [9][64][64 B]
    float* s_b1 = reinterpret_cast<float*>(s_w2 + kW2Bytes);     // [32]
    float* s_b2 = s_b1 + kC;                                      // [64]
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(s_b2 + 2 * kC);  // per stream: x_full[2], d1, d2
    uint64_t* bars = bars_all + 4 * stream;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_all + 4 * kStreams);
    const int tid = threadIdx.x % kStreamThreads, warp = tid >> 5, lane = tid & 31;    // position inside the stream

    // ---- prologue (overlaps the previous layer's tail under PDL): weights (not produced by the previous layer), barriers, TMEM ----
    for (int i = threadIdx.x; i < kC * 8; i += kThreads) {               // W1: row n, 16-byte chunk c
        const int n = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(s_w1 + n * 128 + ((c ^ (n & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(p.w1 + n * 64 + c * 8));
    }
    for (int i = threadIdx.x; i < 9 * 2 * kC * 4; i += kThreads) {       // W2: row R = tap * 64 + n, chunk c
        const int R = i >> 2, c = i & 3;
        *reinterpret_cast<uint4*>(s_w2 + R * 64 + ((c ^ ((R >> 1) & 3)) << 4)) = __ldg(reinterpret_cast<const uint4*>(p.w2 + R * 32 + c * 8));
    }
    if (threadIdx.x < kC) s_b1[threadIdx.x] = p.b1[threadIdx.x];
    if (threadIdx.x < 2 * kC) s_b2[threadIdx.x] = p.b2[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * kStreams; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars_all[i])));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {                              // 128 columns per stream, allocated as one power of two
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        static_assert(kStreams * 128 <= 512, "128 TMEM columns per stream");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the weight tiles above are read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)(stream * 128);
    uint64_t* x_full = bars;
    uint64_t* bar_d1 = bars + 2;
    uint64_t* bar_d2 = bars + 3;
    pdl_wait();                                                           // x is the previous layer's output
    if (threadIdx.x == 0) trace_dep(p.trace);

    const int total = p.tiles_x * p.tiles_y;
    constexpr uint32_t kXBytes = kHalo * 128;
    auto load_x = [&](int tile, int stage) {
        const int ty = fdiv(tile, p.d_tiles_x), tx = tile - ty * p.tiles_x;
        const uint32_t bar = smem_u32(&x_full[stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kXBytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(s_x + stage * kXStage)), "l"(&map_x), "r"(bar), "r"(0), "r"(tx * kTW - 1), "r"(ty * kTH - 1) : "memory");
    };
    // UMMA instruction descriptors: c_format F32 (bit 4), a/b F16, K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * kC) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t w1desc = desc_sw128(smem_u32(s_w1));
    const uint32_t quad = warp & 3, half = warp >> 2;

    const int first = (int)blockIdx.x * kStreams + stream, step = (int)gridDim.x * kStreams;
    if (tid == 0 && first < total) load_x(first, 0);
    int stage = 0;
    uint32_t xphase[2] = {0, 0}, d1phase = 0, d2phase = 0;
    // MMA 1 of a tile: y_halo = x_halo * W1^T (rows 0..127 -> TMEM columns 0..31, rows 128..255 -> columns 32..63).  Issued by warp 0 as
    // soon as the accumulator is free -- for the NEXT tile that is right after MMA 2 of the current one was issued, so it runs on the
    // tensor pipe under epilogue 2 and its completion latency is off the critical path (ncu: the wait for it was 10 % of all warp time).
    auto mma1 = [&](int st) {
        mbar_wait(&x_full[st], xphase[st]);            // whole warp (the epilogues wait on the same phase later: already complete then)
        if (elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t xa = smem_u32(s_x + st * kXStage);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t ad = desc_sw128(xa + h * 128 * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(tmem + (uint32_t)(h * kC), ad + (uint64_t)(2 * k), w1desc + (uint64_t)(2 * k), idesc1, k != 0);
            }
            umma_commit(bar_d1);
        }
        __syncwarp();
    };
    if (kXStages == 2 && warp == 0 && first < total) mma1(0);
    for (int tile = first; tile < total; tile += step) {
        const int ty = fdiv(tile, p.d_tiles_x), tx = tile - ty * p.tiles_x;
        const int x0 = tx * kTW, y0 = ty * kTH;
        // the next tile's halo into the other stage: its last readers (epilogue 2 of the previous tile) are behind the barrier that ended
        // the previous iteration
        if (kXStages == 2 && tid == 0 && tile + step < total) load_x(tile + step, stage ^ 1);
        mbar_wait(&x_full[stage], xphase[stage]);      // (the residual of epilogue 2 reads x with ordinary loads)
        xphase[stage] ^= 1;
        if (kXStages == 1 && warp == 0) mma1(0);       // one stage: MMA 1 cannot run ahead of its halo
        mbar_wait(bar_d1, d1phase);                    // two stages: MMA 1 of this tile was issued one tile ago
        d1phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: halo pixel r = half * 128 + quad * 32 + lane -> 64 bytes of y in the SWIZZLE_64B halo tile ----
        {
            const int r = (int)(half * 128 + quad * 32) + lane;
            if ((int)(half * 128 + quad * 32) < kHalo) {                 // warp-uniform: warps 6, 7 own no halo row
                uint32_t v[32];
                tmem_ld32(tmem + ((quad * 32u) << 16) + half * kC, v);
                if (r < kHalo) {
                    const int hy = r / kHW, hx = r - hy * kHW;
                    const int gx = x0 - 1 + hx, Y = y0 - 1 + hy;
                    const int img = Y >= 0 ? fdiv(Y, p.d_rows) : 0;
                    const bool inside = gx >= 0 && gx < p.width && Y >= 0 && Y < p.total_rows && (Y - img * p.rows) < p.height;
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
                        *reinterpret_cast<uint4*>(s_y + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of y -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        stream_sync(stream);
        // ---- MMA 2: nine taps over the y halo (shifted descriptors), resident W2 -> TMEM columns 64..127 ----
        if (warp == 0) {
            if (elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ys = smem_u32(s_y), ws = smem_u32(s_w2);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t ad = desc_sw64(ys + (uint32_t)((tap / 3) * kHW + (tap % 3)) * 64u, kHW * 64u);
                    const uint64_t bd = desc_sw64(ws + (uint32_t)tap * (2 * kC * 64), 512u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma(tmem + 2 * kC, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc2, (tap | k) != 0);
                }
                umma_commit(bar_d2);
            }
            __syncwarp();
            // epilogue 1 of this tile has drained accumulator 1 (the barrier above): start the next tile's MMA 1 behind MMA 2
            if (kXStages == 2 && tile + step < total) mma1(stage ^ 1);
        }
        mbar_wait(bar_d2, d2phase);
        d2phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: output pixel m = quad * 32 + lane, channels half * 32 .. + 31: + b2, LeakyReLU, + x, fp16, store ----
        {
            const int m = (int)(quad * 32) + lane;
            const int oy = m / kTW, ox = m - oy * kTW;
            uint32_t v[32];
            tmem_ld32(tmem + ((quad * 32u) << 16) + 2 * kC + half * 32, v);
            const int gx = x0 + ox, Y = y0 + oy;
            const int img = fdiv(Y, p.d_rows), yin = Y - img * p.rows;
            const bool valid = gx < p.width && Y < p.total_rows && yin < p.height;
            const int rr = (oy + 1) * kHW + ox + 1;                       // this pixel inside the x halo
            const uint8_t* xrow = s_x + stage * kXStage + rr * 128;
            if (valid) {
                size_t opix = (size_t)Y * p.width + gx;
                if (p.out_s2d) opix = (size_t)(2 * (Y & 1) + (gx & 1)) * (size_t)p.s2d_plane + (size_t)(Y >> 1) * (p.width >> 1) + (gx >> 1);
                __half* o = p.out + opix * (2 * kC) + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 16) {
                    uint32_t wv[8];
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        const int chunk = (int)half * 4 + (i >> 3) + c8;                       // 16-byte chunk of the 128-byte x row
                        const uint4 rv = *reinterpret_cast<const uint4*>(xrow + ((chunk ^ (rr & 7)) << 4));
                        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ch = i + c8 * 8 + 2 * q;
                            const float2 bv = *reinterpret_cast<const float2*>(s_b2 + half * 32 + ch);
                            float a = __uint_as_float(v[ch]), b = __uint_as_float(v[ch + 1]);
                            add2(a, b, bv.x, bv.y);
                            float ma, mb;
                            mul2(ma, mb, a, b, 0.1f);
                            a = fmaxf(a, ma); b = fmaxf(b, mb);
                            const float2 rf = __half22float2(rh[q]);
                            add2(a, b, rf.x, rf.y);
                            const __half2 hv = __floats2half2_rn(a, b);
                            wv[c8 * 4 + q] = *reinterpret_cast<const uint32_t*>(&hv);
                        }
                    }
                    st_global_256(o + i, wv);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        stream_sync(stream);                           // x[stage], y and both accumulators of this stream are free again
        if (kXStages == 2) stage ^= 1;
        else if (tid == 0 && tile + step < total) load_x(tile + step, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) trace_end(p.trace);
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace

namespace om {

bool dark_block_supported(int cin, int cmid, int width, int rows) {
    const char* e = getenv("ORIENMASK_B200_FUSED_BLOCK");
    if (e && e[0] == '0') return false;
    return cin == 2 * kC && cmid == kC && width >= kTW && rows > 0;
}

// x [batch * rows, width, 64] fp16 padded-row NHWC -> out (same geometry; parity-split when out_s2d); w1 / w2 / b1 / b2: the engine's
// OM_PREC_F16 weight layout of the block's two convolutions (BN folded).
int32_t dark_block_run(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, void* out, int batch, int height,
                       int width, int rows, int out_s2d, cudaStream_t stream) {
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
    cuuint64_t dims[3] = {(cuuint64_t)(2 * kC), (cuuint64_t)width, (cuuint64_t)batch * rows};
    cuuint64_t str[2] = {(cuuint64_t)(2 * kC) * 2, (cuuint64_t)width * (2 * kC) * 2};
    cuuint32_t box[3] = {(cuuint32_t)(2 * kC), (cuuint32_t)kHW, (cuuint32_t)kHH};
    cuuint32_t ones[3] = {1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(x), dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled(fused block) failed with CUresult %d", (int)r);
    BlockParams p;
    p.tiles_x = (width + kTW - 1) / kTW;
    p.tiles_y = (batch * rows + kTH - 1) / kTH;
    p.d_tiles_x = make_fastdiv(p.tiles_x); p.d_rows = make_fastdiv(rows);
    p.width = width; p.height = height; p.rows = rows; p.total_rows = batch * rows;
    p.out_s2d = out_s2d; p.s2d_plane = (long long)batch * rows / 2 * (width / 2);
    p.w1 = reinterpret_cast<const __half*>(w1); p.w2 = reinterpret_cast<const __half*>(w2); p.b1 = b1; p.b2 = b2;
    p.out = reinterpret_cast<__half*>(out);
    p.trace = trace_next();
    OM_CUDA_TRY(cudaFuncSetAttribute(dark_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = sms;                                    // one CTA of kStreams tile streams per SM
    const long long tiles = (long long)p.tiles_x * p.tiles_y;
    if (grid * kStreams > tiles) grid = (tiles + kStreams - 1) / kStreams;
    OM_CUDA_TRY(launch_pdl(dark_block_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmem, stream, map, p));
    return check_launch("dark_block_kernel");
}

}  // namespace om
