// fp16 implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B/64B-swizzled shared memory, persistent warp-specialised CTAs.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[tap, cout, cin]
//
// * A (activations, padded-row NHWC fp16, see include/orienmask_b200.h): one 3-D tensor map
//   {C, W, rows}; a CTA tile is th full-width-tw row segments = th*tw <= 128 output pixels, and every
//   filter tap is the same TMA box shifted by (dx, dy).  Horizontal padding = TMA out-of-bounds
//   zero fill, vertical padding = the zero rows between images.  Stride-2 layers use four
//   parity-split views of the input (even/odd rows x even/odd columns) so that each tap is again a
//   dense box.
// * B (weights [tap][cout_pad][cin] fp16, K-major): 2-D tensor map {cin, taps*cout_pad}.
// * warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
//   warps 2..5 = epilogue (TMEM -> registers -> bias / up-add / LeakyReLU / residual -> global).
//   Two accumulator stages in TMEM let the epilogue of tile i overlap the main loop of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int kThreads = 192;            // 6 warps
constexpr int kEpiWarp0 = 2;
constexpr int kBlockM = 128;
constexpr int kMaxStages = 8;

struct TcParams {
    int tiles_x, tiles_y, tiles_n;       // tile grid; linear tile id = (ty * tiles_x + tx) * tiles_n + tn
    int tw, th;                          // pixel tile (tw * th <= 128)
    int taps, stride;                    // 1 or 9; 1 or 2
    int k_chunks;                        // cin / BK
    int block_n;                         // UMMA N (multiple of 16, <= 256)
    int cout_pad;                        // weight rows per tap
    int stages;
    int tmem_cols;                       // allocated columns (power of two >= 2*block_n, <= 512)
    uint32_t idesc;
    // epilogue
    int out_h, out_w, out_rows, total_rows;   // valid rows per image, width, rows_per_image, B*rows_per_image
    int cout, cout_stride, leaky, out_kind;
    int up_rows;
    const float* bias;
    const __half* residual;
    const float* upadd;
    void* output;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
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

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address and byte
// offsets in 16-byte units, version 1 (Blackwell), layout 2 = SWIZZLE_128B / 4 = SWIZZLE_64B.
// Rows are BK*2 bytes apart inside an 8-row swizzle atom; atoms follow each other every 8 rows.
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t sbo = (8 * BK * 2) >> 4;                  // 1024 B (BK=64) or 512 B (BK=32)
    constexpr uint64_t layout = (BK == 64) ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
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

struct TileCoord { int tx, ty, tn; };
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
    TileCoord t;
    t.tn = tile % p.tiles_n;
    const int r = tile / p.tiles_n;
    t.tx = r % p.tiles_x;
    t.ty = r / p.tiles_x;
    return t;
}

template <int BK>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
               const __grid_constant__ CUtensorMap map_b, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][A tile | B tile] (1024-aligned), then barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = kBlockM * BK * 2;
    const int b_bytes = p.block_n * BK * 2;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;                         // [stages]
    uint64_t* empty_bar = bars + kMaxStages;           // [stages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [tiles_n * block_n] (whole padded cout)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int k_iters = p.taps * p.k_chunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b));
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.tiles_n * p.block_n; i += kThreads)
        s_bias[i] = (p.bias != nullptr && i < p.cout) ? p.bias[i] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)(p.tw * p.th * BK * 2 + b_bytes);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                const int x0 = t.tx * p.tw, y0 = t.ty * p.th, n0 = t.tn * p.block_n;
                for (int tap = 0; tap < p.taps; ++tap) {
                    int dx = 0, dy = 0;
                    const CUtensorMap* ma = &map_a0;
                    if (p.taps == 9) {
                        const int r = tap / 3, s = tap - r * 3;
                        if (p.stride == 1) { dx = s - 1; dy = r - 1; }
                        else {
                            dx = (s == 0) ? -1 : 0; dy = (r == 0) ? -1 : 0;
                            const int sel = ((r != 1) ? 2 : 0) + ((s != 1) ? 1 : 0);   // odd row / odd column views
                            ma = sel == 0 ? &map_a0 : sel == 1 ? &map_a1 : sel == 2 ? &map_a2 : &map_a3;
                        }
                    }
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_3d(sa, ma, &full_bar[stage], kc * BK, x0 + dx, y0 + dy);
                        tma_load_2d(sa + a_bytes, &map_b, &full_bar[stage], kc * BK, tap * p.cout_pad + n0);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
                for (int kit = 0; kit < k_iters; ++kit) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_kmajor_desc<BK>(sa);
                    const uint64_t bdesc = make_kmajor_desc<BK>(sa + a_bytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)       // +32 bytes (2 x 16B units) per K=16 step inside the swizzle atom
                        umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (kit | k) != 0);
                    umma_commit(&empty_bar[stage]);         // frees the smem slot once these MMAs have read it
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[as]);                // accumulator complete -> epilogue
                as ^= 1; if (as == 0) aphase ^= 1;
            }
        }
    } else {
        // ===== epilogue: 4 warps, TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        const int m = quad * 32 + lane;                     // accumulator row = pixel inside the tile
        int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            const int my = m / p.tw, mx = m - my * p.tw;
            const int Y = t.ty * p.th + my, x = t.tx * p.tw + mx;
            const int img = Y / p.out_rows, y = Y - img * p.out_rows;
            const bool valid = (m < p.tw * p.th) && (Y < p.total_rows) && (y < p.out_h);
            const int n0 = t.tn * p.block_n;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.block_n);
            const size_t pix = (size_t)Y * p.out_w + x;
            const float* up = nullptr;
            if (p.upadd != nullptr && valid)
                up = p.upadd + ((size_t)(img * p.up_rows + (y >> 1)) * (p.out_w >> 1) + (x >> 1)) * p.cout;
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);          // warp-collective: every lane participates
                if (!valid) continue;
                const int cg = n0 + c0;                      // first global channel of this chunk
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                if (up != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 u = *reinterpret_cast<const float4*>(up + cg + i);
                        f[i] += u.x; f[i + 1] += u.y; f[i + 2] += u.z; f[i + 3] += u.w;
                    }
                }
                if (p.out_kind != OM_OUT_PARTIAL) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] += s_bias[cg + i];
                }
                if (p.leaky) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.0f ? f[i] : 0.1f * f[i];
                }
                if (p.out_kind == OM_OUT_ACT) {
                    __half* o = reinterpret_cast<__half*>(p.output) + pix * p.cout_stride + cg;
                    if (p.residual != nullptr) {
                        const __half* r = p.residual + pix * p.cout_stride + cg;
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            const uint4 rv = *reinterpret_cast<const uint4*>(r + i);
                            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 rf = __half22float2(rh[j]);
                                f[i + 2 * j] += rf.x; f[i + 2 * j + 1] += rf.y;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 ov;
                        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                        for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                        *reinterpret_cast<uint4*>(o + i) = ov;
                    }
                } else if (p.out_kind == OM_OUT_PARTIAL) {
                    float* o = reinterpret_cast<float*>(p.output) + pix * p.cout_stride + cg;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                } else {   // OM_OUT_NCHW: dense fp32 [B, cout, H, W]; lanes hold consecutive pixels -> coalesced per channel
                    float* o = reinterpret_cast<float*>(p.output) + ((size_t)img * p.cout * p.out_h + y) * p.out_w + x;
                    const size_t plane = (size_t)p.out_h * p.out_w;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (cg + i < p.cout) o[(size_t)(cg + i) * plane] = f[i];
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[as]);
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int32_t encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, int bk) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return OM_OK;
}

struct TcPlan {
    CUtensorMap map_a[4];
    CUtensorMap map_b;
    TcParams p;
    int bk;
    int grid;
    size_t smem;
};

int pick_tile_w(int w) {
    int best = 1, best_fill = 0;
    for (int tw = 1; tw <= w && tw <= kBlockM; ++tw) {
        if (w % tw) continue;
        const int fill = tw * (kBlockM / tw);
        if (fill >= best_fill) { best_fill = fill; best = tw; }
    }
    return best;
}

}  // namespace

namespace om {

int32_t tc_plan_create(const om_conv_desc& d, void** out) {
    if (d.cin % 32) return fail(OM_ERR_INVALID, "fp16 engine needs cin %% 32 == 0 (got %d)", d.cin);
    if (d.out_kind != OM_OUT_NCHW && (d.cout % 32 || d.cout_stride % 8 || d.cout_stride < d.cout))
        return fail(OM_ERR_INVALID, "fp16 engine needs cout %% 32 == 0 and an aligned channel pitch for NHWC outputs");
    if (d.upadd && (d.out_w % 2 || d.out_kind == OM_OUT_NCHW)) return fail(OM_ERR_INVALID, "upadd needs an even width and an NHWC output");
    if (d.stride == 2 && (d.in_rows != 2 * d.out_rows || d.in_w != 2 * d.out_w || d.ksize != 3))
        return fail(OM_ERR_INVALID, "stride-2 layers must be 3x3 with in_rows == 2*out_rows and in_w == 2*out_w");
    if (d.stride == 1 && (d.in_rows != d.out_rows || d.in_w != d.out_w))
        return fail(OM_ERR_INVALID, "stride-1 layers need identical input/output geometry");
    TcPlan* plan = new TcPlan();
    memset(plan, 0, sizeof(TcPlan));
    TcParams& p = plan->p;
    const int bk = (d.cin % 64 == 0) ? 64 : 32;
    plan->bk = bk;
    int cout_pad = (d.cout + 15) / 16 * 16;
    if (cout_pad < 32) cout_pad = 32;
    int bn = cout_pad;
    if (bn > 256) {
        bn = 256;
        while (cout_pad % bn) bn -= 16;
    }
    p.block_n = bn; p.cout_pad = cout_pad; p.tiles_n = cout_pad / bn;
    p.tw = pick_tile_w(d.out_w); p.th = kBlockM / p.tw;
    p.tiles_x = d.out_w / p.tw;
    p.total_rows = d.batch * d.out_rows;
    p.tiles_y = (p.total_rows + p.th - 1) / p.th;
    p.taps = d.ksize * d.ksize; p.stride = d.stride; p.k_chunks = d.cin / bk;
    const int stage_bytes = (kBlockM + bn) * bk * 2;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { delete plan; return fail(OM_ERR_INVALID, "tile does not fit shared memory"); }
    p.stages = stages;
    int cols = 32;
    while (cols < 2 * bn) cols <<= 1;
    p.tmem_cols = cols;
    // cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b F16, K-major, N>>3 at bit 17, M>>4 at bit 24
    p.idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
    p.out_h = d.out_h; p.out_w = d.out_w; p.out_rows = d.out_rows;
    p.cout = d.cout; p.cout_stride = d.cout_stride; p.leaky = d.leaky; p.out_kind = d.out_kind;
    p.up_rows = d.up_rows; p.bias = d.bias; p.residual = reinterpret_cast<const __half*>(d.residual);
    p.upadd = d.upadd; p.output = d.output;

    const size_t esz = 2;
    int32_t rc = OM_OK;
    if (d.stride == 1) {
        cuuint64_t dims[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.in_w, (cuuint64_t)d.batch * d.in_rows};
        cuuint64_t str[2] = {(cuuint64_t)d.cin * esz, (cuuint64_t)d.in_w * d.cin * esz};
        cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)p.tw, (cuuint32_t)p.th};
        rc = encode(&plan->map_a[0], d.input, 3, dims, str, box, bk);
        for (int i = 1; i < 4 && rc == OM_OK; ++i) plan->map_a[i] = plan->map_a[0];
    } else {
        for (int sel = 0; sel < 4 && rc == OM_OK; ++sel) {          // sel = 2*odd_row + odd_col
            const int py = sel >> 1, px = sel & 1;
            cuuint64_t dims[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.in_w / 2, (cuuint64_t)d.batch * d.in_rows / 2};
            cuuint64_t str[2] = {(cuuint64_t)2 * d.cin * esz, (cuuint64_t)2 * d.in_w * d.cin * esz};
            cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)p.tw, (cuuint32_t)p.th};
            const char* base = reinterpret_cast<const char*>(d.input) + ((size_t)py * d.in_w + px) * d.cin * esz;
            rc = encode(&plan->map_a[sel], base, 3, dims, str, box, bk);
        }
    }
    if (rc == OM_OK) {
        cuuint64_t dims[2] = {(cuuint64_t)d.cin, (cuuint64_t)p.taps * cout_pad};
        cuuint64_t str[1] = {(cuuint64_t)d.cin * esz};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
        rc = encode(&plan->map_b, d.weights, 2, dims, str, box, bk);
    }
    if (rc != OM_OK) { delete plan; return rc; }

    plan->smem = 1024 + (size_t)stages * stage_bytes + (2 * kMaxStages + 4) * 8 + 16 + (size_t)cout_pad * 4;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    plan->grid = tiles < sms ? tiles : sms;
    // the attribute is per kernel function, not per launch: always opt in to the full 227 KB
    constexpr int kMaxSmem = 227 * 1024;
    if (plan->smem > (size_t)kMaxSmem) {
        const size_t need = plan->smem;
        delete plan;
        return fail(OM_ERR_INVALID, "tile needs %zu bytes of shared memory", need);
    }
    cudaError_t e = bk == 64
        ? cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
        : cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) { delete plan; return fail(OM_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); }
    *out = plan;
    return OM_OK;
}

int32_t tc_plan_run(const void* vp, cudaStream_t stream) {
    const TcPlan* plan = reinterpret_cast<const TcPlan*>(vp);
    if (plan->bk == 64)
        conv_tc_kernel<64><<<plan->grid, kThreads, plan->smem, stream>>>(plan->map_a[0], plan->map_a[1], plan->map_a[2], plan->map_a[3], plan->map_b, plan->p);
    else
        conv_tc_kernel<32><<<plan->grid, kThreads, plan->smem, stream>>>(plan->map_a[0], plan->map_a[1], plan->map_a[2], plan->map_a[3], plan->map_b, plan->p);
    return check_launch("conv_tc_kernel");
}

void tc_plan_destroy(void* vp) { delete reinterpret_cast<TcPlan*>(vp); }

}  // namespace om
