// fp16 implicit-GEMM convolution on CTA pairs: tcgen05.mma.cta_group::2 (UMMA M = 256 over two SMs), operands
// staged by TMA into swizzled shared memory, accumulators in TMEM, persistent warp-specialised CTAs.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[tap, cout, cin]
//
// Why pairs: with one CTA per 128 x 256 tile the shared-memory fill (16 KB of activations + 32 KB of weights per 512
// tensor-core cycles) bounds the 3x3 layers, not the tensor pipe (profiles/r01_ncu_conv_v1.txt).  A CTA pair
// computes two vertically adjacent pixel tiles against the same weight tile; each CTA stages its own activations
// plus HALF of the weight rows and the cta_group::2 MMA reads both halves.
//
// Roles (19 warps / CTA): warp 0 operand producer (TMA), warp 1 TMEM allocator + MMA issuer (only the leader CTA
// issues), warp 2 addend producer (TMA: residual or up-add chunks), warps 3..18 epilogue in four groups of four.
// Operand staging has three shapes (see the comment at the role dispatch): per-tap boxes, one halo box per chunk that
// feeds all nine taps of a 3x3 stride-1 layer, and -- when all weights of a small layer fit -- resident weights.
// Epilogue: TMEM -> registers -> (+ fp32 up-add) + bias -> LeakyReLU -> (+ fp16 residual) -> 32-byte global stores;
// up to four accumulator stages in TMEM let the main loop run ahead of it.  Every launch uses programmatic dependent
// launch: everything before pdl_wait() overlaps the previous layer's tail.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int kEpiGroups = 4;             // epilogue groups of 4 warps (power of two)
constexpr int kThreads = (3 + 4 * kEpiGroups) * 32;   // producer, MMA, addend producer, kEpiGroups x 4 epilogue warps
constexpr int kEpiWarp0 = 3;
constexpr int kBlockM = 128;             // pixels per CTA tile (UMMA M = 256 per pair)
constexpr int kMaxStages = 8;
constexpr int kStageRows = 128;          // staging tile rows
constexpr int kStageBytes = kStageRows * 128;
constexpr int kMaxResBufs = 2 * kEpiGroups;   // addend staging buffers: one lane per epilogue group, up to 2 deep
constexpr int kMaxAcc = 8;                    // accumulator stages in TMEM (narrow tiles: 512 columns / N)

struct Tc2Params {
    FastDiv d_tiles_n, d_tiles_x, d_out_rows, d_tw, d_flat_hw, d_out_w, d_k_chunks;
    int tiles_x, pairs_y, tiles_n;       // pair grid; linear pair id = (py * tiles_x + tx) * tiles_n + tn
    int tw, th;                          // pixel tile (tw * th <= 128)
    int taps, stride;
    int k_chunks;                        // K-loop chunks per tap: cin / BK (split precision: 3 * cin / BK, see below)
    int a_chunks;                        // BK-channel chunks of one pixel in memory: cin / BK (split precision: 2 * cin / BK, hi | lo)
    int split_kr;                        // split precision (OM_PREC_SPLIT): cin / BK, else 0.  Activations and weights are stored as
                                         // fp16 hi | lo halves along the channel axis and the K loop of a tap walks 3 * kr chunks:
                                         // (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) -- three MMAs into the same fp32 accumulator
    float acc_scale;                     // split precision: the weights were scaled by a power of two to keep W_lo normal; undone here
    int pix_stride, lo_off;              // split precision: fp16 elements per output pixel (2 * cout_stride) and offset of the lo half
    int block_n, half_n;                 // UMMA N and the rows of it each CTA stages
    int cout_pad;
    int halo;                            // 3x3 stride 1 with tw == 8: one halo box per chunk feeds all nine taps
    int halo_s2;                         // halo mode for a stride-2 3x3 over a parity-split input: one (tw+1) x (th+1) box per parity plane
                                         // and chunk (4 boxes instead of 9 tap boxes); tap (r, s) reads plane 2*(r!=1) + (s!=1) at offset
                                         // ((r!=0), (s!=0)) inside the box
    int halo_pitch;                      // pixels per row of a halo box: tw + 2 (stride 1) or tw + 1 (stride 2)
    int halo_planes;                     // boxes per chunk: 1 or 4
    int b_resident;                      // halo mode, all weights of the layer fit: loaded once, no block ring at all
    int n_sub;                           // (tap, chunk) blocks per pipeline stage
    int stages;                          // depth of the block ring
    int h_stages, h_stage_bytes, h_chunk_bytes;   // halo ring (halo mode): one stage = the halos of all chunks of a tile
    int a_box_pixels;                    // pixels per activation TMA box
    int a_sub_bytes;                     // bytes reserved per per-tap activation block
    int tmem_cols;
    int acc_stages;                      // accumulator stages in TMEM (2..4): narrow tiles let the MMA run further ahead of the epilogue
    uint32_t idesc;
    // epilogue
    int out_h, out_w, out_rows, total_rows;
    int cout, leaky, out_kind, up_rows;
    int has_res;                         // epilogue addend staged by TMA: 0 none, 1 residual (fp16, same resolution, after the
                                         // activation), 2 up-add (fp32 half-resolution partial sum, before bias and activation)
    int res_depth;                       // addend staging slots per epilogue group (1..2)
    int res_bufs, res_bufs_log2;         // addend staging buffers (kEpiGroups * res_depth): chunk number `seq` of the CTA's chunk sequence uses
                                         // buffer seq % res_bufs (phase (seq / res_bufs) & 1) and is consumed by epilogue group seq % kEpiGroups.
                                         // (Tried: 64-column chunks of the tensor-bound 68x68 residual layers in TWO buffers shared by the
                                         // four groups, to halve the staging's TMA rows -- no gain, in-situ A/B; a buffer shared across
                                         // groups also needs an extra wait, because a parity wait is only meaningful within one phase.)
    int res_buf_bytes;                   // one staged 32-column chunk: 128 rows x 64 B (fp16 residual) or x 128 B (fp32 up-add)
    int up_bw, up_bh;                    // up-add box: tw/2 + 1 by th/2 + 1 source pixels (covers odd tile origins)
    int chunk_cols;                      // accumulator columns per staged chunk (64 fp16 / 32 fp32 / 32 narrow fp16)
    int row_bytes;                       // bytes per staged row (128 or 64)
    int cout_stride;
    int out_s2d;                         // write the fp16 output parity-split (see om_conv_desc)
    int flat;                            // pixel tile = 128 consecutive REAL output pixels in (image, y, x) order, gathered by an
                                         // im2col-mode tensor map (no pad rows, no partially filled tiles; see tc2_plan_create)
    int flat_hw, flat_total, pad;        // out_h * out_w, batch * out_h * out_w, ksize / 2
    int res_direct;                      // flat mode: the fp16 residual is read by the epilogue threads themselves (prefetched one
                                         // chunk ahead) -- im2col-mode TMA costs ~6.5 cycles per pixel whatever the row size, and the
                                         // flat layers are bounded by exactly that rate (profiles/r01_flat_tiles.md)
    const void* residual;
    long long s2d_plane;                 // pixels per parity plane of the output
    const float* bias;
    const float* upadd;
    void* output;
    unsigned long long* trace;           // om_debug_trace record of this launch, or nullptr
    int h_halves;                        // split precision with ONE halo stage: its lo half (read by pass 0 only) and its hi half (passes 1, 2)
                                         // have their own barrier pairs, so the next tile's lo half loads under passes 1-2 of this tile and
                                         // its hi half under ITS pass 0: the halo load of a tile is no longer exposed (see the planner)
    int epi_sleep_ns;                    // back-off of the epilogue warps' wait for an accumulator (see mbar_wait_relaxed)
    int hint_a, hint_w, hint_res, hint_out;   // L2 eviction priority of the activation / weight / residual loads and of the fp16 output stores
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t num_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// One lane of a converged warp (always the same one).  The single-thread instructions (TMA, tcgen05.mma/commit,
// expect_tx) are issued under this predicate while the surrounding loops stay warp-uniform: inside a
// `lane == 0` branch ptxas cannot keep descriptors/addresses in uniform registers and wraps every
// UTCHMMA/UTMALDG in an ELECT/R2UR waterfall loop (~20 instructions per MMA), which made the issuing thread --
// not the tensor pipe -- the bottleneck.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Execution barrier over the CTA pair.  Relaxed arrive: the default release form compiles to MEMBAR.ALL.GPU and makes
// every thread wait until all of its global stores are visible device-wide (~0.5 us at kernel exit); what the two
// uses need is weaker -- barrier-init visibility comes from fence.mbarrier_init.release.cluster, and at exit only
// "the peer no longer touches my shared memory / TMEM" matters.
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default (.release.cta) semantics: a cluster-scope release would drain every global store in flight first
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.  The slow path passes a suspend-time hint:
// the warp is parked by the hardware until the phase completes or the hint expires instead of re-issuing try_wait (with 15 of
// the 19 warps waiting on something at any time, polling took half of the issue slots of the memory-bound layers away from
// the four epilogue warps that had work: ncu, 1x1 64->32 @272: 3850 warp instructions per tile at IPC 2).
__device__ unsigned int g_wait_hint_ns = 20000;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    {
        uint32_t ok;                                   // fast path: no clock reads when the phase has already completed
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    const unsigned int hint = g_wait_hint_ns;
    const long long t0 = clock64();
    for (uint32_t spins = 1;; ++spins) {               // the clock is read every 64th round only: this loop was 19 % of conv2.0's instructions
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(hint) : "memory");
        if (ok) return;
        if ((spins & 63u) == 0u && clock64() - t0 > 4000000000ll) __trap();
    }
}

// The epilogue warps' wait for the next accumulator lasts most of a tile (7 us on the tensor-bound layers) and has a whole tile of
// slack: between polls they sleep.  (The hinted try_wait above returns after ~150 cycles whatever the hint says; with 16 warps waiting
// the polls were 47 % of the warp instructions of a tensor-bound layer -- ncu source view of backbone.conv4.1.conv.1.)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int sleep_ns) {
    if (sleep_ns <= 0) { mbar_wait(bar, parity); return; }
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (uint32_t spins = 1;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep((unsigned)sleep_ns);
        if ((spins & 63u) == 0u && clock64() - t0 > 4000000000ll) __trap();
    }
}

// Operand loads of a CTA pair: data lands in the issuing CTA, completion bytes go to the LEADER's barrier.
// L2 eviction-priority policies (createpolicy): 0 none, 1 evict_first (data that is dead after this access), 2 evict_last (data
// the next layers read again).  Every operand load / output store carries one; see the planner (l2 hints) for who gets what.
// 0 = no hint at all (NOT the same as an explicit evict_normal policy: with that on every load the memory-bound 136x136 layers
// measured 20 % slower, in-situ A/B) -- the wrappers below issue the plain instruction when the policy word is 0.
__device__ __forceinline__ uint64_t l2_policy(int kind) {
    uint64_t p = 0;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, uint64_t pol) {
    if (pol == 0)
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, uint64_t pol) {
    if (pol == 0)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t pol) {
    if (pol == 0)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
            ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
// im2col-mode loads (tensor map from cuTensorMapEncodeIm2col): `pixelsPerColumn` pixels starting at input position (w, h) of
// image n, shifted by the filter tap (ow, oh), walking x -> y -> image inside the map's bounding box with its traversal
// stride; positions outside the image read zeros.
__device__ __forceinline__ void tma_load_im2col_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c, int w, int h, int n,
                                                     int ow, int oh, uint64_t pol) {
    if (pol == 0)
        asm volatile(
            "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_im2col(void* dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h, int n) {
    const uint16_t z = 0;
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %7};"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(z) : "memory");
}
// One 32-byte store per thread: a whole sector, so the L1 -> L2 write traffic is not inflated by partial sectors.
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void st_global_256_hint(void* ptr, const uint32_t (&w)[8], uint64_t pol) {
    if (pol == 0) { st_global_256(ptr, w); return; }
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;"
                 ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(pol) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
// Arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address and byte
// offsets in 16-byte units, version 1 (Blackwell), layout 2 = SWIZZLE_128B / 4 = SWIZZLE_64B.
// `sbo_bytes` is the distance between consecutive 8-row groups: 8 rows for a dense tile, one halo row
// (tw + 2 pixels) when the eight rows of a group are the eight pixels of one tile row inside a halo.
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
    constexpr uint64_t layout = (BK == 64) ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
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

// Byte offset of 16-byte chunk `c` of row `r` in a TMA-swizzled staging tile (128B or 64B rows).
__device__ __forceinline__ uint32_t swz(int r, int c, int row_bytes) {
    return row_bytes == 128 ? (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)) : (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

struct AMaps { CUtensorMap m[4]; };       // activation views: [0] everything for stride 1; 2*odd_row + odd_col parity views for stride 2

// Debug timeline (tools/_bin/tiny_layer.py): when set, cluster 0 records %globaltimer at fixed points of one launch.
__device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ void tick(int slot, bool who) {
    if (g_timeline != nullptr && who && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[slot] = t;
    }
}

struct PairCoord { int tx, py, tn; };
__device__ __forceinline__ PairCoord decode_pair(const Tc2Params& p, int pair) {
    PairCoord t;
    const int r = fdiv(pair, p.d_tiles_n);
    t.tn = pair - r * p.tiles_n;
    t.py = fdiv(r, p.d_tiles_x);
    t.tx = r - t.py * p.tiles_x;
    return t;
}

// Flat mode: first output pixel (x, y, image) of pixel tile `tm`; tiles past the end (odd tile count) re-read tile 0 and
// are masked in the epilogue.
struct FlatOrigin { int x, y, n; };
__device__ __forceinline__ FlatOrigin flat_origin(const Tc2Params& p, int tm) {
    int q0 = tm * kBlockM;
    if (q0 >= p.flat_total) q0 = 0;
    FlatOrigin o;
    o.n = fdiv(q0, p.d_flat_hw);
    const int rem = q0 - o.n * p.flat_hw;
    o.y = fdiv(rem, p.d_out_w);
    o.x = rem - o.y * p.out_w;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue.  kEpiGroups groups of 4 warps (TMEM lane quadrant = warp % 4); the 32-column chunks of the accumulators
// are dealt round-robin to the groups over this CTA's whole chunk sequence (chunk g -> group g % kEpiGroups, which is
// also the addend staging lane it reads), so kEpiGroups chunks are in flight.  The memory-bound layers are bounded
// by how fast TMEM is drained, transformed and stored: with one warp per scheduler the dependent-issue latency of
// the ~250 instructions per 32 columns was the limit (ncu: epilogue warps busy, tensor pipe 18 % active), hence
// four warps per scheduler and one specialised instance per (output kind, addend kind).
//   KIND: 0 fp16 NHWC activation, 1 fp32 NHWC (partial sums), 2 fp32 NCHW (heads)
//   ADD : 0 none, 1 residual (fp16 chunk staged by TMA, added after the activation), 2 up-add (fp32 half-resolution
//         chunk staged by TMA, added before bias and activation), 3 residual read straight from global memory by the
//         thread that owns the pixel, one 32-column chunk ahead (flat mode)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg_res32(uint4 (&r)[4], const __half* src) {      // 32 fp16 = 64 bytes of one pixel
    // two 32-byte loads (LDG.256): whole sectors per lane -- with the L1 carved down to nothing, four 16-byte loads fetched every
    // sector twice from L2
#pragma unroll
    for (int i = 0; i < 4; i += 2)
        asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[i].x), "=r"(r[i].y), "=r"(r[i].z), "=r"(r[i].w), "=r"(r[i + 1].x), "=r"(r[i + 1].y), "=r"(r[i + 1].z), "=r"(r[i + 1].w)
                     : "l"(src + 8 * i) : "memory");
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct EpiCtx {
    uint32_t tmem_base, leader_tmem_empty0, rank, s_bias_addr;
    uint64_t* tmem_full; uint64_t* res_full; uint64_t* res_empty;
    const uint8_t* res_buf; const float* s_bias;
    int warp, lane, first_pair, pair_step, num_pairs;
};

template <int KIND, int ADD, bool FLAT, bool SPLIT>
__device__ __forceinline__ void epilogue_loop(const Tc2Params& p, const EpiCtx& c) {
    const int quad = c.warp & 3;
    const int grp = (c.warp - kEpiWarp0) >> 2;
    const int m = quad * 32 + c.lane;                   // accumulator row = pixel inside the tile
    const int n_chunks = p.block_n / p.chunk_cols;
    const int my = fdiv(m, p.d_tw), mx = m - my * p.tw;
    const bool in_tile = m < p.tw * p.th;
    const uint64_t pol_out = l2_policy(p.hint_out);
    int as = 0; uint32_t aphase = 0;
    int seq0 = 0;                                        // sequence number of the tile's first chunk in this CTA's chunk sequence
    int g0 = 0;                                          // ... mod kEpiGroups
    for (int pair = c.first_pair; pair < c.num_pairs; pair += c.pair_step) {
        if (pair == c.first_pair) pdl_wait();          // while the first accumulator is still being produced
        const int j_first = (grp - g0) & (kEpiGroups - 1);
        if (j_first >= n_chunks) {
            // Narrow tiles (N = 32 / 64) give work to one or two of the four groups; the others only keep the accumulator
            // barrier in step, without the tile's coordinate arithmetic (with all 16 warps doing it, a memory-bound 1x1
            // 64 -> 32 tile cost 3850 warp instructions at IPC 2: issue-bound, not HBM-bound).
            mbar_wait_relaxed(&c.tmem_full[as], aphase, p.epi_sleep_ns);
            tc_fence_after();
            tc_fence_before();
            __syncwarp();
            if (c.lane == 0) mbar_arrive_cluster(c.leader_tmem_empty0 + (uint32_t)(as * 8));
            g0 = (g0 + n_chunks) & (kEpiGroups - 1);
            seq0 += n_chunks;
            if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
            continue;
        }
        const PairCoord t = decode_pair(p, pair);
        int x0 = t.tx * p.tw, y0 = (2 * t.py + (int)c.rank) * p.th;
        int Y = y0 + my, x = x0 + mx;
        int img = fdiv(Y, p.d_out_rows), y = Y - img * p.out_rows;
        bool valid = in_tile && (Y < p.total_rows) && (y < p.out_h) && (x < p.out_w);
        if (FLAT) {                                     // accumulator row m = output pixel (2 * py + rank) * 128 + m in (image, y, x) order
            const int q = (2 * t.py + (int)c.rank) * kBlockM + m;
            valid = q < p.flat_total;
            img = fdiv(q, p.d_flat_hw);
            const int rem = q - img * p.flat_hw;
            y = fdiv(rem, p.d_out_w); x = rem - y * p.out_w;
            Y = img * p.out_rows + y;
        }
        const int n0 = t.tn * p.block_n;
        uint4 rr[4] = {};
        const __half* rsrc = nullptr;
        if (ADD == 3) {
            rsrc = reinterpret_cast<const __half*>(p.residual) + ((size_t)Y * p.out_w + x) * (SPLIT ? p.pix_stride : p.cout_stride) + n0;
            if (!SPLIT && valid && j_first < n_chunks) ldg_res32(rr, rsrc + j_first * 32);
        }
        mbar_wait_relaxed(&c.tmem_full[as], aphase, p.epi_sleep_ns);
        if (pair == c.first_pair) tick(6, c.warp == kEpiWarp0 && c.lane == 0);
        tc_fence_after();
        const uint32_t taddr = c.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.block_n);
        const size_t pix = (size_t)Y * p.out_w + x;
        const float* up = nullptr;                       // fallback up-add straight from global (geometry TMA cannot box)
        if (ADD == 0 && KIND != 2 && p.upadd != nullptr && valid)
            up = p.upadd + ((size_t)(img * p.up_rows + (y >> 1)) * (p.out_w >> 1) + (x >> 1)) * p.cout;
        const int up_row = ((Y >> 1) - (y0 >> 1)) * p.up_bw + ((x >> 1) - (x0 >> 1));   // source pixel inside the staged box
        for (int j = j_first; j < n_chunks; j += kEpiGroups) {
            const int seq = seq0 + j;
            const int rb = seq & (p.res_bufs - 1);
            if (ADD == 1 || ADD == 2) {
                mbar_wait(&c.res_full[rb], (uint32_t)(seq >> p.res_bufs_log2) & 1u);
            }
            for (int c0 = 0; c0 < p.chunk_cols; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(j * p.chunk_cols + c0), v);
                if (j + kEpiGroups >= n_chunks && c0 + 32 >= p.chunk_cols) {   // this group's part of the accumulator is drained
                    tc_fence_before();
                    __syncwarp();
                    if (c.lane == 0) mbar_arrive_cluster(c.leader_tmem_empty0 + (uint32_t)(as * 8));
                }
                const int cg = n0 + j * p.chunk_cols + c0;    // first global channel of these 32 columns
                float f[32];
    #pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                if (SPLIT) {
    #pragma unroll
                    for (int i = 0; i < 32; i += 2) mul2(f[i], f[i + 1], f[i], f[i + 1], p.acc_scale);   // power of two (x gain fix)
                }
                if (ADD == 2) {
                    const uint8_t* ubuf = c.res_buf + rb * p.res_buf_bytes;
                    if (in_tile) {
    #pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 u = *reinterpret_cast<const float4*>(ubuf + swz(up_row, i >> 2, 128));
                            add2(f[i], f[i + 1], u.x, u.y); add2(f[i + 2], f[i + 3], u.z, u.w);
                        }
                    }
                } else if (up != nullptr) {
    #pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(up + cg + i));
                        add2(f[i], f[i + 1], u.x, u.y); add2(f[i + 2], f[i + 3], u.z, u.w);
                    }
                }
    #pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bv = lds_f4(c.s_bias_addr + (uint32_t)(cg + i) * 4u);          // warp-uniform: smem broadcast
                    add2(f[i], f[i + 1], bv.x, bv.y); add2(f[i + 2], f[i + 3], bv.z, bv.w);
                }
                if (p.leaky) {
    #pragma unroll
                    for (int i = 0; i < 32; i += 2) {                                       // LeakyReLU(0.1) = max(f, 0.1 f)
                        float m0, m1;
                        mul2(m0, m1, f[i], f[i + 1], 0.1f);
                        f[i] = fmaxf(f[i], m0); f[i + 1] = fmaxf(f[i + 1], m1);
                    }
                }
                if (KIND == 0) {
                    if (ADD == 1) {
                        const uint8_t* rbuf = c.res_buf + rb * p.res_buf_bytes;
    #pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            const uint4 rv = *reinterpret_cast<const uint4*>(rbuf + swz(m, (c0 >> 3) + (i >> 3), p.row_bytes));
                            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
    #pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 rf = __half22float2(rh[q]);
                                add2(f[i + 2 * q], f[i + 2 * q + 1], rf.x, rf.y);
                            }
                        }
                    }
                    if (ADD == 3 && SPLIT) {
                        if (valid) {                         // residual = hi + lo, both halves read by the thread that owns the pixel
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                ldg_res32(rr, rsrc + half * p.lo_off + j * 32);
#pragma unroll
                                for (int i = 0; i < 32; i += 8) {
                                    const __half2* rh = reinterpret_cast<const __half2*>(&rr[i >> 3]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 rf = __half22float2(rh[q]);
                                        add2(f[i + 2 * q], f[i + 2 * q + 1], rf.x, rf.y);
                                    }
                                }
                            }
                        }
                    }
                    if (ADD == 3 && !SPLIT) {
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                const __half2* rh = reinterpret_cast<const __half2*>(&rr[i >> 3]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 rf = __half22float2(rh[q]);
                                    add2(f[i + 2 * q], f[i + 2 * q + 1], rf.x, rf.y);
                                }
                            }
                            if (j + kEpiGroups < n_chunks) ldg_res32(rr, rsrc + (j + kEpiGroups) * 32);
                        }
                    }
                    if (valid && SPLIT) {                    // hi = fp16(f), lo = fp16(f - hi): the pair carries ~22 significant bits
                        size_t opix = pix;
                        if (p.out_s2d) opix = (size_t)(2 * (Y & 1) + (x & 1)) * (size_t)p.s2d_plane + (size_t)(Y >> 1) * (p.out_w >> 1) + (x >> 1);
                        __half* o = reinterpret_cast<__half*>(p.output) + opix * p.pix_stride + cg;
    #pragma unroll
                        for (int i = 0; i < 32; i += 16) {
                            uint32_t wh[8], wl[8];
    #pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const __half2 h = __floats2half2_rn(f[i + 2 * q], f[i + 2 * q + 1]);
                                const float2 hb = __half22float2(h);
                                const __half2 l = __floats2half2_rn(f[i + 2 * q] - hb.x, f[i + 2 * q + 1] - hb.y);
                                wh[q] = *reinterpret_cast<const uint32_t*>(&h);
                                wl[q] = *reinterpret_cast<const uint32_t*>(&l);
                            }
                            st_global_256(o + i, wh);
                            st_global_256(o + p.lo_off + i, wl);
                        }
                    }
                    if (valid && !SPLIT) {
                        size_t opix = pix;
                        if (p.out_s2d) opix = (size_t)(2 * (Y & 1) + (x & 1)) * (size_t)p.s2d_plane + (size_t)(Y >> 1) * (p.out_w >> 1) + (x >> 1);
                        __half* o = reinterpret_cast<__half*>(p.output) + opix * p.cout_stride + cg;
    #pragma unroll
                        for (int i = 0; i < 32; i += 16) {       // 2 x 32-byte stores: every store fills whole sectors
                            uint32_t w[8];
    #pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const __half2 h = __floats2half2_rn(f[i + 2 * q], f[i + 2 * q + 1]);
                                w[q] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            st_global_256_hint(o + i, w, pol_out);
                        }
                    }
                } else if (KIND == 1) {
                    if (valid) {
                        float* o = reinterpret_cast<float*>(p.output) + pix * p.cout_stride + cg;
    #pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            uint32_t w[8];
    #pragma unroll
                            for (int q = 0; q < 8; ++q) w[q] = __float_as_uint(f[i + q]);
                            st_global_256(o + i, w);
                        }
                    }
                } else if (valid) {   // fp32 NCHW [B, cout, H, W]; lanes hold consecutive pixels of a row
                    float* o = reinterpret_cast<float*>(p.output) + ((size_t)img * p.cout * p.out_h + y) * p.out_w + x;
                    const size_t plane = (size_t)p.out_h * p.out_w;
    #pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (cg + i < p.cout) o[(size_t)(cg + i) * plane] = f[i];
                }
            }
            if (ADD == 1 || ADD == 2) {
                __syncwarp();
                if (c.lane == 0) mbar_arrive(&c.res_empty[rb]);
            }
        }
        g0 = (g0 + n_chunks) & (kEpiGroups - 1);
        seq0 += n_chunks;
        if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
    }
    tick(7, c.warp == kEpiWarp0 && c.lane == 0);
}


template <int BK, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ AMaps maps_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_res, const Tc2Params p) {
    const CUtensorMap& map_a0 = maps_a.m[0];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    tick(0, threadIdx.x == 0);
    if (threadIdx.x == 0) trace_start(p.trace);
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_sub = p.half_n * BK * 2;                                // one (tap, chunk) block of this CTA's weight rows
    const int a_sub = p.a_sub_bytes;                                    // per-tap mode: the matching activation box (0 in halo mode)
    const int sub_bytes = a_sub + b_sub;
    const int stage_bytes = p.n_sub * sub_bytes;
    uint8_t* h_ring = smem;                                             // [h_stages][h_stage_bytes]   (halo mode only)
    uint8_t* s_ring = smem + (size_t)p.h_stages * p.h_stage_bytes;      // [stages][n_sub][A box | B block]
    uint8_t* res_buf = s_ring + (size_t)p.stages * stage_bytes;         // [2][kStageBytes] (has_res only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(res_buf + (p.has_res ? p.res_bufs * p.res_buf_bytes : 0));
    uint64_t* s_full = bars;                           // [stages]     (the leader's copies of the full barriers are the live ones)
    uint64_t* s_empty = bars + kMaxStages;             // [stages]
    uint64_t* h_full = bars + 2 * kMaxStages;          // [h_stages]
    uint64_t* h_empty = bars + 3 * kMaxStages;         // [h_stages]
    uint64_t* tmem_full = bars + 4 * kMaxStages;       // [kMaxAcc]
    uint64_t* tmem_empty = tmem_full + kMaxAcc;        // [kMaxAcc]  (leader's copy: 16 warp arrivals from both CTAs)
    uint64_t* res_full = tmem_empty + kMaxAcc;         // [kMaxResBufs]  buffer = group + 2 * slot
    uint64_t* res_empty = res_full + kMaxResBufs;      // [kMaxResBufs]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + kMaxResBufs);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);            // [cout_pad]; L1 is carved down to nothing here, so
                                                                        // per-use global bias loads would each pay an L2 round trip

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int num_pairs = p.tiles_x * p.pairs_y * p.tiles_n;
    const int first_pair = (int)cluster_id_x(), pair_step = (int)num_clusters_x();
    const int n_groups = p.b_resident ? 0 : p.taps * p.k_chunks / p.n_sub;   // stages per tile

    if (warp == 0) {                                   // one barrier pair per lane: the ~60 inits are not a serial prologue
        if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0));
        if (lane == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b));
        if (lane < p.stages) { mbar_init(&s_full[lane], 1); mbar_init(&s_empty[lane], 1); }
        if (lane >= 8 && lane - 8 < (p.h_halves ? 2 : p.h_stages)) { mbar_init(&h_full[lane - 8], 1); mbar_init(&h_empty[lane - 8], 1); }
        if (lane >= 16 && lane < 16 + kMaxAcc) { mbar_init(&tmem_full[lane - 16], 1); mbar_init(&tmem_empty[lane - 16], 8 * kEpiGroups); }
        if (lane >= 24 && lane - 24 < kMaxResBufs) { mbar_init(&res_full[lane - 24], 1); mbar_init(&res_empty[lane - 24], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tick(11, lane == 0);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        tick(12, lane == 0);
    }
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads)
        s_bias[i] = (p.bias != nullptr && p.out_kind != OM_OUT_PARTIAL && i < p.cout) ? p.bias[i] : 0.0f;
    tick(13, threadIdx.x == 64);
    tc_fence_before();
    __syncthreads();
    tick(14, threadIdx.x == 0);
    cluster_sync();                                    // peer barriers are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tick(1, threadIdx.x == 0);
    // pdl_wait() (griddepcontrol.wait) is executed per role, as late as possible: the producer right before its first
    // activation load (weights are not produced by the previous layer), the addend producer before its first load, the
    // epilogue warps before their first store.                                        // everything above overlapped the previous layer's tail

    // The K loop of a tile is the sequence of (tap, 64-channel chunk) blocks, tap-major.  A pipeline stage holds
    // n_sub consecutive blocks behind ONE full/empty barrier pair, so the single MMA-issuing thread pays one wait
    // and one commit per n_sub * BK/16 MMAs (the issuer, not the tensor pipe, bounds narrow-N layers otherwise).
    //   per-tap mode: block = activation box {BK, tw, th} shifted by the tap + weight block;
    //   halo mode (3x3 stride 1, tw = 8): block = weight block only; the activations of the whole tile are ONE
    //   {BK, tw+2, th+2} halo box per chunk in a second ring, and a tap is a shifted descriptor over it (the
    //   tensor core applies the 128B/64B swizzle to absolute shared-memory address bits, so a start address
    //   shifted by whole pixels with a row pitch of tw+2 pixels addresses the TMA-written halo correctly --
    //   tools/desc_probe.cu, profiles/r01_desc_probe.txt).
    if (warp == 0) {
        {
            // ===== operand producer (both CTAs).  The whole warp runs the loop; lane j issues the TMA loads of block j of
            // the stage, so the ~80 instructions of tap decoding / coordinate arithmetic per block run in parallel
            // across lanes instead of serialising on one thread (which bounded the layers with small boxes). =====
            int st = 0; uint32_t s_phase = 0;
            int hs = 0; uint32_t h_phase = 0;
            const uint64_t pol_a = l2_policy(p.hint_a), pol_w = l2_policy(p.hint_w);
            const uint32_t s_tx = 2u * (uint32_t)p.n_sub * (uint32_t)((p.halo ? 0 : p.a_box_pixels * BK * 2) + b_sub);
            const uint32_t h_tx = 2u * (uint32_t)(p.a_chunks * p.halo_planes) * (uint32_t)(p.a_box_pixels * BK * 2);
            if (p.b_resident) {                               // the whole weight tensor (this CTA's half of the rows), once
                const uint32_t lb = mapa(smem_u32(&s_full[0]), 0);
                if (lane == 0 && rank == 0) mbar_expect_tx(&s_full[0], 2u * (uint32_t)(p.taps * p.k_chunks * b_sub));
                if (lane < p.taps * p.k_chunks) {
                    const int tap = lane / p.k_chunks, kc = lane - tap * p.k_chunks;
                    tma_load_2d_pair(s_ring + (size_t)lane * b_sub, &map_b, lb, kc * BK, tap * p.cout_pad + (int)rank * p.half_n, pol_w);
                }
                __syncwarp();
            }
            for (int pair = first_pair; pair < num_pairs; pair += pair_step) {
                const PairCoord t = decode_pair(p, pair);
                const int x0 = t.tx * p.tw, y0 = (2 * t.py + (int)rank) * p.th;
                const int n0 = t.tn * p.block_n + (int)rank * p.half_n;
                FlatOrigin fo = {0, 0, 0};
                if (p.flat) fo = flat_origin(p, 2 * t.py + (int)rank);
                const int fw = fo.x * p.stride - p.pad, fh = fo.y * p.stride - p.pad;
                // half-stage mode: barrier pair 1 = the lo chunks [kr, 2 kr) of the one stage, pair 0 = the hi chunks [0, kr); h_phase toggles per tile
                auto load_half = [&](int half) {
                    mbar_wait(&h_empty[half], h_phase ^ 1);
                    const uint32_t lb = mapa(smem_u32(&h_full[half]), 0);
                    if (lane == 0 && rank == 0) mbar_expect_tx(&h_full[half], h_tx >> 1);
                    if (lane < p.split_kr) {
                        const int kc = half * p.split_kr + lane;
                        tma_load_3d_pair(h_ring + (size_t)kc * p.h_chunk_bytes, &maps_a.m[0], lb, kc * BK, x0 - 1, y0 - 1, pol_a);
                    }
                    __syncwarp();
                };
                if (p.halo && p.h_halves) {
                    if (pair == first_pair) { pdl_wait(); tick(2, lane == 0); if (lane == 0) trace_dep(p.trace); }
                    load_half(1);                                  // lo first: pass 0 reads it; the hi half follows in front of pass 1 (below)
                } else if (p.halo) {
                    mbar_wait(&h_empty[hs], h_phase ^ 1);
                    const uint32_t lb = mapa(smem_u32(&h_full[hs]), 0);
                    if (pair == first_pair) { pdl_wait(); tick(2, lane == 0); if (lane == 0) trace_dep(p.trace); }
                    if (lane == 0 && rank == 0) mbar_expect_tx(&h_full[hs], h_tx);
                    if (lane < p.a_chunks * p.halo_planes) {       // buffer order: [plane][chunk]; plane coordinates == output coordinates
                        const int plane = lane / p.a_chunks, kc = lane - plane * p.a_chunks;
                        tma_load_3d_pair(h_ring + (size_t)hs * p.h_stage_bytes + (size_t)lane * p.h_chunk_bytes, &maps_a.m[plane], lb, kc * BK, x0 - 1, y0 - 1, pol_a);
                    }
                    __syncwarp();
                    if (++hs == p.h_stages) { hs = 0; h_phase ^= 1; }
                }
                for (int g = 0; g < n_groups; ++g) {
                    if (p.h_halves && 3 * g == n_groups) { load_half(0); h_phase ^= 1; }     // the hi half, in front of the first block of pass 1
                    mbar_wait(&s_empty[st], s_phase ^ 1);
                    const uint32_t lb = mapa(smem_u32(&s_full[st]), 0);
                    uint8_t* sb = s_ring + (size_t)st * stage_bytes;
                    const int i = g * p.n_sub + lane;                // (tap, chunk) block of this lane, tap-major
                    int tap = fdiv(i, p.d_k_chunks), kc = i - tap * p.k_chunks;
                    int ka = kc, kb = kc;                            // channel chunk of the activation / weight block in memory
                    if (SPLIT) {
                        // pass-major K loop: (A_lo, W_hi) over all taps, then (A_hi, W_lo), then (A_hi, W_hi).  tcgen05 accumulates
                        // with truncation (every MMA loses ~half an ulp of the RUNNING SUM, toward zero), so the two correction
                        // passes run while the accumulator is still ~2^-11 of its final magnitude and only the hi*hi chain pays
                        const int per_pass = p.taps * p.split_kr;
                        const int pass = i / per_pass, r = i - pass * per_pass;
                        tap = r / p.split_kr;
                        kc = r - tap * p.split_kr;
                        ka = (pass == 0 ? p.split_kr : 0) + kc;
                        kb = (pass == 1 ? p.split_kr : 0) + kc;
                    }
                    int dx = 0, dy = 0, sel = 0;
                    int tap_r = 0, tap_s = 0;
                    if (!p.halo && p.taps == 9) {
                        const int r = tap / 3, s = tap - r * 3;
                        tap_r = r; tap_s = s;
                        if (p.stride == 1) { dx = s - 1; dy = r - 1; }
                        else {
                            dx = (s == 0) ? -1 : 0; dy = (r == 0) ? -1 : 0;
                            sel = ((r != 1) ? 2 : 0) + ((s != 1) ? 1 : 0);       // odd row / odd column views
                        }
                    }
                    if (!p.halo && pair == first_pair && g == 0) { pdl_wait(); tick(2, lane == 0); if (lane == 0) trace_dep(p.trace); }   // first activation load
                    if (lane == 0 && rank == 0) mbar_expect_tx(&s_full[st], s_tx);
                    if (lane < p.n_sub) {
                        if (p.flat) tma_load_im2col_pair(sb + (size_t)lane * sub_bytes, &map_a0, lb, ka * BK, fw, fh, fo.n, tap_s, tap_r, pol_a);
                        else if (!p.halo) tma_load_3d_pair(sb + (size_t)lane * sub_bytes, &maps_a.m[sel], lb, ka * BK, x0 + dx, y0 + dy, pol_a);
                        tma_load_2d_pair(sb + (size_t)lane * sub_bytes + a_sub, &map_b, lb, kb * BK, tap * p.cout_pad + n0, pol_w);
                    }
                    __syncwarp();
                    if (pair == first_pair && g == 0) tick(3, lane == 0);
                    if (++st == p.stages) { st = 0; s_phase ^= 1; }
                }
            }
            tick(8, lane == 0);
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA; whole warp runs the loop, one elected lane issues) =====
            // The loop body is descriptor arithmetic only (64-bit adds of 16-byte units on precomputed bases): with
            // ~50 integer instructions per (tap, chunk) block the single issuing thread, at one dependent instruction
            // every few cycles, was slower than the MMAs of a narrow tile (N = 64: 32 cycles per MMA).
            int st = 0; uint32_t s_phase = 0;
            int hs = 0; uint32_t h_phase = 0;
            int as = 0; uint32_t aphase = 0;
            constexpr uint32_t pixel_bytes = BK * 2;
            constexpr uint64_t kLayout = (BK == 64) ? 2 : 4;
            const uint64_t hi_b = ((uint64_t)((8 * pixel_bytes) >> 4) << 32) | (1ull << 16) | (1ull << 46) | (kLayout << 61);
            const uint64_t hi_a = p.halo ? (((uint64_t)(((uint32_t)p.halo_pitch * pixel_bytes) >> 4) << 32) | (1ull << 16) | (1ull << 46) | (kLayout << 61)) : hi_b;
            // descriptor offset (16-byte units) of filter tap (r, s) inside the tile's halo stage
            const uint32_t plane16 = (uint32_t)p.a_chunks * ((uint32_t)p.h_chunk_bytes >> 4);
            auto tap_off = [&](int r, int s) -> uint32_t {
                if (p.halo_s2) return (uint32_t)(2 * (r != 1) + (s != 1)) * plane16 + (uint32_t)((r != 0) * p.halo_pitch + (s != 0)) * (pixel_bytes >> 4);
                return (uint32_t)(r * p.halo_pitch + s) * (pixel_bytes >> 4);
            };
            const uint32_t s_ring_addr = smem_u32(s_ring), h_ring_addr = smem_u32(h_ring);
            const uint32_t sub16 = (uint32_t)sub_bytes >> 4, asub16 = (uint32_t)a_sub >> 4, bsub16 = (uint32_t)b_sub >> 4;
            const uint32_t hchunk16 = (uint32_t)p.h_chunk_bytes >> 4;
            if (p.b_resident) mbar_wait(&s_full[0], 0);
            for (int pair = first_pair; pair < num_pairs; pair += pair_step) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
                uint64_t ad_tile = 0;
                if (p.halo) {
                    mbar_wait(&h_full[p.h_halves ? 1 : hs], h_phase);          // (half-stage mode: the lo half; the hi half is awaited in front of pass 1)
                    ad_tile = hi_a | (uint64_t)(((h_ring_addr + (uint32_t)(hs * p.h_stage_bytes)) & 0x3FFFF) >> 4);
                }
                tc_fence_after();
                if (p.b_resident) {                       // halo + resident weights: no ring, constant offsets
                    if (elect_one()) {
                        uint64_t bd = hi_b | (uint64_t)((s_ring_addr & 0x3FFFF) >> 4);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            uint64_t ad = ad_tile + (uint64_t)tap_off(tap / 3, tap % 3);
                            for (int kc = 0; kc < p.k_chunks; ++kc) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_f16_pair(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), p.idesc, (tap | kc | k) != 0 ? 1u : 0u);
                                ad += hchunk16; bd += bsub16;
                            }
                        }
                    }
                    __syncwarp();
                } else {
                    uint32_t accumulate = 0;
                    uint32_t tap_off16 = p.halo ? tap_off(0, 0) : 0; int tap_r = 0, tap_s = 0, kc = 0;   // halo: descriptor offset of the current tap, its row / column, chunk
                    int pass = 0;                                             // split precision: see the producer
                    const int kc_end = SPLIT ? p.split_kr : p.k_chunks;
                    for (int g = 0; g < n_groups; ++g) {
                        if (p.h_halves && 3 * g == n_groups) {             // pass 0 is issued: the lo half is free once it completes; pass 1 needs the hi half
                            if (elect_one()) umma_commit_pair(&h_empty[1]);
                            __syncwarp();
                            mbar_wait(&h_full[0], h_phase);
                            tc_fence_after();
                        }
                        mbar_wait(&s_full[st], s_phase);
                        if (pair == first_pair && g == 0) tick(4, lane == 0);
                        tc_fence_after();
                        const uint32_t sb16 = ((s_ring_addr + (uint32_t)(st * stage_bytes)) & 0x3FFFF) >> 4;
                        uint64_t bd = hi_b | (uint64_t)(sb16 + asub16);
                        uint64_t ad_ring = hi_a | (uint64_t)sb16;
                        for (int j = 0; j < p.n_sub; ++j) {
                            const int ka = (SPLIT && pass == 0) ? kc + p.split_kr : kc;            // split: pass 0 reads A_lo, passes 1 and 2 A_hi
                            const uint64_t ad = p.halo ? ad_tile + (uint64_t)(tap_off16 + (uint32_t)ka * hchunk16) : ad_ring;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_f16_pair(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), p.idesc, (k == 0) ? accumulate : 1u);
                            }
                            accumulate = 1;
                            bd += sub16; ad_ring += sub16;
                            if (++kc == kc_end) {           // next tap
                                kc = 0;
                                if (++tap_s == 3) { tap_s = 0; ++tap_r; }
                                if (SPLIT && tap_r == 3) { tap_r = 0; ++pass; }
                                if (p.halo) tap_off16 = tap_off(tap_r, tap_s);
                            }
                        }
                        if (elect_one()) umma_commit_pair(&s_empty[st]);
                        __syncwarp();
                        if (++st == p.stages) { st = 0; s_phase ^= 1; }
                    }
                }
                if (p.halo && p.h_halves) {
                    if (elect_one()) umma_commit_pair(&h_empty[0]);
                    h_phase ^= 1;
                } else if (p.halo) {
                    if (elect_one()) umma_commit_pair(&h_empty[hs]);
                    if (++hs == p.h_stages) { hs = 0; h_phase ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&tmem_full[as]);
                __syncwarp();
                if (pair == first_pair) tick(5, lane == 0);
                if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp == 2) {
        if (p.has_res) {
            // ===== residual producer: the 64-channel chunks of this CTA's output tile, in epilogue order =====
            // chunk g of this CTA's chunk sequence goes to epilogue group g & 1; that group's k-th chunk (k = g >> 1)
            // uses its slot k % res_depth, so the loads run res_depth chunks ahead of each group (the DRAM latency of
            // the addend was the top stall of the memory-bound residual layers with a single slot)
            pdl_wait();
            const uint64_t pol_r = l2_policy(p.hint_res);
            int seq = 0;
            const int n_chunks = p.block_n / p.chunk_cols;
            const uint32_t bytes = p.has_res == 1 ? (uint32_t)(p.tw * p.th * p.row_bytes) : (uint32_t)(p.up_bw * p.up_bh * 128);
            for (int pair = first_pair; pair < num_pairs; pair += pair_step) {
                const PairCoord t = decode_pair(p, pair);
                int x0 = t.tx * p.tw, y0 = (2 * t.py + (int)rank) * p.th;
                if (p.has_res == 2) { x0 >>= 1; y0 >>= 1; }          // half-resolution source of the nearest x2 up-sampling
                FlatOrigin fo = {0, 0, 0};
                if (p.flat) fo = flat_origin(p, 2 * t.py + (int)rank);
                for (int j = 0; j < n_chunks; ++j) {
                    const int rb = seq & (p.res_bufs - 1);
                    mbar_wait(&res_empty[rb], ((uint32_t)(seq >> p.res_bufs_log2) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(&res_full[rb], bytes);
                        if (p.flat) tma_load_im2col(res_buf + rb * p.res_buf_bytes, &map_res, &res_full[rb], t.tn * p.block_n + j * p.chunk_cols, fo.x, fo.y, fo.n);
                        else tma_load_3d(res_buf + rb * p.res_buf_bytes, &map_res, &res_full[rb], t.tn * p.block_n + j * p.chunk_cols, x0, y0, pol_r);
                    }
                    __syncwarp();
                    ++seq;
                }
            }
        }
    } else {
        EpiCtx c;
        c.tmem_base = tmem_base; c.leader_tmem_empty0 = mapa(smem_u32(&tmem_empty[0]), 0); c.rank = rank;
        c.tmem_full = tmem_full; c.res_full = res_full; c.res_empty = res_empty;
        c.res_buf = res_buf; c.s_bias = s_bias; c.s_bias_addr = smem_u32(s_bias);
        c.warp = warp; c.lane = lane; c.first_pair = first_pair; c.pair_step = pair_step; c.num_pairs = num_pairs;
        if (p.flat) {
            if (p.out_kind == OM_OUT_PARTIAL) epilogue_loop<1, 0, true, SPLIT>(p, c);
            else if (p.res_direct) epilogue_loop<0, 3, true, SPLIT>(p, c);
            else epilogue_loop<0, 0, true, SPLIT>(p, c);
        }
        else if (p.out_kind == OM_OUT_NCHW) epilogue_loop<2, 0, false, SPLIT>(p, c);
        else if (p.out_kind == OM_OUT_PARTIAL) { if (p.has_res == 2) epilogue_loop<1, 2, false, SPLIT>(p, c); else epilogue_loop<1, 0, false, SPLIT>(p, c); }
        else if (p.res_direct) epilogue_loop<0, 3, false, SPLIT>(p, c);
        else if (!SPLIT && p.has_res == 1) epilogue_loop<0, 1, false, false>(p, c);      // split precision reads its residual directly
        else if (p.has_res == 2) epilogue_loop<0, 2, false, SPLIT>(p, c);
        else epilogue_loop<0, 0, false, SPLIT>(p, c);
    }

    tick(9, threadIdx.x == 0);
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                    // nobody leaves while the peer may still touch its smem / TMEM
    tick(10, threadIdx.x == 0);
    if (threadIdx.x == 0) trace_end(p.trace);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ORIENMASK_B200_PLAN_DRYRUN=1 (tests/test_planner.py): om_conv_create plans a layer without a GPU -- 148 SMs assumed, tensor maps
// left zeroed after checking the arguments against the documented limits of cuTensorMapEncode*, no kernel attribute set.  A plan
// made this way is never launched (tc2_plan_run refuses).
bool plan_dryrun() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ORIENMASK_B200_PLAN_DRYRUN"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int sm_count() {
    if (plan_dryrun()) return 148;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// The argument rules of cuTensorMapEncodeTiled / Im2col without interleave (CUDA driver API reference).
int32_t check_map_args(const char* who, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       int inner_elems, int esz, int inner_bytes) {
    if (rank < 1 || rank > 5) return om::fail(OM_ERR_INVALID, "%s: rank %d", who, rank);
    if (reinterpret_cast<uintptr_t>(base) & 15) return om::fail(OM_ERR_INVALID, "%s: base address not 16-byte aligned", who);
    for (int i = 0; i < rank; ++i)
        if (dims[i] < 1 || dims[i] > (1ull << 32)) return om::fail(OM_ERR_INVALID, "%s: dim %d = %llu", who, i, (unsigned long long)dims[i]);
    for (int i = 0; i + 1 < rank; ++i)
        if (strides_bytes[i] % 16 || strides_bytes[i] >= (1ull << 40))
            return om::fail(OM_ERR_INVALID, "%s: stride %d = %llu bytes", who, i, (unsigned long long)strides_bytes[i]);
    if (inner_bytes != 128 && inner_bytes != 64) return om::fail(OM_ERR_INVALID, "%s: swizzle span %d", who, inner_bytes);
    if (inner_elems < 1 || inner_elems > 256 || inner_elems * esz > inner_bytes || (inner_elems * esz) % 16)
        return om::fail(OM_ERR_INVALID, "%s: inner box of %d x %d bytes does not fit the %d-byte swizzle span", who, inner_elems, esz, inner_bytes);
    return OM_OK;
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

int32_t encode(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, int inner_bytes) {
    if (plan_dryrun()) {
        memset(map, 0, sizeof(*map));
        for (int i = 0; i < rank; ++i)
            if (box[i] < 1 || box[i] > 256) return om::fail(OM_ERR_INVALID, "tiled map: box dim %d = %u", i, box[i]);
        return check_map_args("tiled map", base, rank, dims, strides_bytes, (int)box[0], dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2, inner_bytes);
    }
    EncodeTiledFn fn = get_encode();
    if (!fn) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return OM_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC tensor [batch][h (image pitch `rows` rows)][w][c (pixel pitch `c_stride`)] as an im2col map: `channels` channels of 128 output
// pixels per load; k x k filter with padding k/2 and traversal stride `stride` (bounding-box corners as cuDNN / CUTLASS fprop).
int32_t encode_im2col(CUtensorMap* map, const void* base, int c, int c_stride, int w, int h, int rows, int batch, int channels, int ksize,
                      int stride, int inner_bytes) {
    if (plan_dryrun()) {
        memset(map, 0, sizeof(*map));
        cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
        cuuint64_t str[3] = {(cuuint64_t)c_stride * 2, (cuuint64_t)w * c_stride * 2, (cuuint64_t)rows * w * c_stride * 2};
        if (kBlockM > 1024 || stride < 1 || stride > 8) return om::fail(OM_ERR_INVALID, "im2col map: pixels per column / traversal stride");
        return check_map_args("im2col map", base, 4, dims, str, channels, 2, inner_bytes);
    }
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(ptr);
    }
    if (!fn) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
    const size_t esz = 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
    cuuint64_t str[3] = {(cuuint64_t)c_stride * esz, (cuuint64_t)w * c_stride * esz, (cuuint64_t)rows * w * c_stride * esz};
    const int pad = ksize / 2;
    int lower[2] = {-pad, -pad}, upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
    cuuint32_t trav[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, str, lower, upper, (cuuint32_t)channels,
                    (cuuint32_t)kBlockM, trav, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return om::fail(OM_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
    // same fix-up CUTLASS applies (copy_traits_sm90_im2col.hpp) for drivers <= 13.1 on tensors smaller than 128 KB
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (size_t)batch * rows * w * c_stride * esz < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return OM_OK;
}

struct Tc2Plan {
    AMaps maps;
    CUtensorMap map_b, map_res;
    Tc2Params p;
    int bk;
    bool split;
    int grid;
    size_t smem;
};

// Pixel tile of <= 128 = tw x th with tw | width.  NHWC outputs: the fullest tile, discounting tiles narrower
// than 8 pixels (their TMA boxes degenerate into many 128..512-byte rows and they share little halo in L2).
// NCHW heads: the widest row segment among tiles that are at least 75 % full, so that the per-channel fp32
// stores of a warp form few long runs without idling a large part of the epilogue lanes.
int pick_tile_w(int w, bool widest) {
    int best = 1, best_score = -1;
    for (int tw = 1; tw <= w && tw <= kBlockM; ++tw) {
        if (w % tw) continue;
        const int fill = tw * (kBlockM / tw);
        int score;
        if (widest) score = (fill * 4 >= kBlockM * 3) ? 1000 + tw : fill;
        else score = fill * (tw < 8 ? tw : 8);
        if (score >= best_score) { best_score = score; best = tw; }
    }
    return best;
}

}  // namespace

namespace om {

int32_t tc2_set_wait_hint(unsigned int ns);

static int32_t plan_create(const om_conv_desc& d, void** out, bool allow_halo, bool* halo_used) {
    if (d.cin % 32) return fail(OM_ERR_INVALID, "fp16 engine needs cin %% 32 == 0 (got %d)", d.cin);
    const bool split = d.precision == OM_PREC_SPLIT;
    const int cin_mem = split ? 2 * d.cin : d.cin;          // fp16 elements per input pixel (split precision: hi | lo)
    if (split && d.out_kind == OM_OUT_ACT && d.cout_stride != d.cout)
        return fail(OM_ERR_INVALID, "split precision needs a dense activation output (cout_stride == cout)");
    if (d.out_kind != OM_OUT_NCHW && (d.cout % 32 || d.cout_stride % 16 || d.cout_stride < d.cout))
        return fail(OM_ERR_INVALID, "fp16 engine needs cout %% 32 == 0 and an aligned channel pitch for NHWC outputs");
    if (d.upadd && (d.out_w % 2 || d.out_kind == OM_OUT_NCHW)) return fail(OM_ERR_INVALID, "upadd needs an even width and an NHWC output");
    if (d.stride == 2 && (d.in_w != 2 * d.out_w || d.ksize != 3))
        return fail(OM_ERR_INVALID, "stride-2 layers must be 3x3 with in_w == 2*out_w");
    if (d.stride == 1 && (d.in_rows != d.out_rows || d.in_w != d.out_w))
        return fail(OM_ERR_INVALID, "stride-1 layers need identical input/output geometry");
    {
        static bool hint_set = false;
        const char* wh = getenv("ORIENMASK_B200_WAIT_HINT");
        if (!hint_set && wh && atoi(wh) >= 0 && atoi(wh) <= 1000000) { tc2_set_wait_hint((unsigned int)atoi(wh)); }
        hint_set = true;
    }
    Tc2Plan* plan = new Tc2Plan();
    memset(plan, 0, sizeof(Tc2Plan));
    Tc2Params& p = plan->p;
    const int bk = (d.cin % 64 == 0) ? 64 : 32;
    plan->bk = bk;
    const int cout_pad = (d.cout + 31) / 32 * 32;       // a CTA pair splits the N tile in halves of a multiple of 16 rows
    int bn = cout_pad;
    int bn_cap = 256;
    if (getenv("ORIENMASK_B200_BN") && d.out_w <= atoi(getenv("ORIENMASK_B200_BN_MAXW") ? getenv("ORIENMASK_B200_BN_MAXW") : "0")) {
        const int v = atoi(getenv("ORIENMASK_B200_BN"));   // experiment: narrower N tiles for the layers that quantise badly
        if (v == 32 || v == 64 || v == 128 || v == 256) bn_cap = v;
    }
    if (bn > bn_cap) {
        bn = bn_cap;
        while (bn > 32 && cout_pad % bn) bn -= 32;
    }
    if (!getenv("ORIENMASK_B200_BN") && !(getenv("ORIENMASK_B200_NARROW_N") && getenv("ORIENMASK_B200_NARROW_N")[0] == '0')) {
        // Small batches: when the layer has fewer pixel-tile pairs than the GPU has SM pairs, most SMs idle while a few grind through a
        // 256-wide tile.  Split N further (128 / 64 columns) as long as everything still fits one wave: the same MMA work spreads over
        // 2-4x the SMs.  Model per wave: a fixed ~3000 cycles (first loads, epilogue, exit) + K steps * N/2 tensor cycles.
        // bs 1 at 544x544, forward + post-process: 1.45 -> 1.12 ms (A/B over a fixed cap, profiles/r02_latency_bs1_narrow_n.txt).
        const int clusters = sm_count() / 2;
        const long long m_pairs = (((long long)d.batch * d.out_h * d.out_w + kBlockM - 1) / kBlockM + 1) / 2;
        if (m_pairs * (cout_pad / bn) < clusters) {
            const long long steps = (long long)d.ksize * d.ksize * (d.cin / 16) * (split ? 3 : 1);
            long long best_t = -1;
            int best_bn = bn;
            for (int cand = bn; cand >= 32; cand >>= 1) {          // (N = 32 where it still fits one wave: bs 1, 17x17 3x3 layers -10 %, A/B)
                if (cout_pad % cand || cand % 32) continue;
                const long long pairs = m_pairs * (cout_pad / cand);
                const long long waves = (pairs + clusters - 1) / clusters;
                const long long t = waves * (3000 + steps * (cand / 2));
                if (best_t < 0 || t < best_t) { best_t = t; best_bn = cand; }
            }
            bn = best_bn;
        }
    }
    if (bn % 32) { delete plan; return fail(OM_ERR_INVALID, "padded cout %d cannot be split over a CTA pair", cout_pad); }
    p.block_n = bn; p.half_n = bn / 2; p.cout_pad = cout_pad; p.tiles_n = cout_pad / bn;
    const char* halo_env = getenv("ORIENMASK_B200_HALO");
    p.halo = d.ksize == 3 && d.stride == 1 && d.out_kind != OM_OUT_NCHW && (d.out_w % 8 == 0 || d.out_w >= 64) &&
             !(halo_env && halo_env[0] == '0') && allow_halo;
    *halo_used = p.halo != 0;
    // Flat tiles: every layer that is not served by a halo box, a parity-split input or a TMA-staged up-add computes tiles of 128
    // consecutive real pixels (image, y, x) gathered by im2col-mode TMA: no pad rows and no partially filled tiles, which at
    // 17x17 / 34x34 is the difference between 3 and 2 (5 and 4) waves of CTA pairs.
    const char* flat_env = getenv("ORIENMASK_B200_FLAT");
    p.flat = !p.halo && !d.in_s2d && d.upadd == nullptr && d.out_kind != OM_OUT_NCHW && !d.out_s2d && !(flat_env && flat_env[0] == '0');
    p.flat_hw = d.out_h * d.out_w; p.flat_total = d.batch * p.flat_hw; p.pad = d.ksize / 2;
    if (p.flat && !(flat_env && flat_env[0] == '2')) {
        // ... but im2col-mode loads cost ~6.5 cycles per pixel, which the memory-bound layers with many waves of tiles cannot
        // afford: keep rectangular boxes when quantisation is not the problem (more than ~10 waves of pairs)
        const int sms = sm_count();
        const long long flat_pairs = ((p.flat_total + kBlockM - 1) / kBlockM + 1) / 2;
        // (a stride-2 layer whose input and output row pitches are not 2:1 has no rectangular alternative: it stays flat)
        if (flat_pairs * (cout_pad / bn) > 10ll * (sms / 2) && !(d.stride == 2 && d.in_rows != 2 * d.out_rows)) p.flat = 0;
    }
    if (d.stride == 2 && !p.flat && d.in_rows != 2 * d.out_rows) {      // rectangular / parity-plane boxes address rows of the whole batch
        delete plan;
        return fail(OM_ERR_INVALID, "this stride-2 layer needs in_rows == 2*out_rows (only im2col-gathered tiles are free of it)");
    }
    p.halo_s2 = d.ksize == 3 && d.stride == 2 && d.in_s2d && d.out_kind != OM_OUT_NCHW && (d.out_w % 8 == 0 || d.out_w >= 64) &&
                !(halo_env && halo_env[0] == '0') && !getenv("ORIENMASK_B200_NO_HALO_S2");
    if (p.halo_s2) {
        const size_t chunk = ((size_t)(9 * 17) * bk * 2 + 1023) / 1024 * 1024;
        const size_t halo2 = 2 * 4 * (size_t)((split ? 2 : 1) * (d.cin / bk)) * chunk;          // two stages of four plane boxes per chunk
        if (!split) {
            // only where all weights stay resident next to two halo stages (the 32 -> 64 layer): with a weight ring the four plane
            // boxes leave too few stages and the layer gets slower than with per-tap boxes (64 -> 128 @136: 113 -> 165 us)
            const size_t need = halo2 + (size_t)9 * (d.cin / bk) * (bn / 2) * bk * 2 + 8 * 1024;
            if (cout_pad / bn != 1 || need > 227 * 1024) p.halo_s2 = 0;
        } else {
            // split precision (no resident weights): per-tap boxes make this layer TMA-row bound -- 27 (tap, pass) blocks x (128 + 32)
            // rows per tile = 15 000 cycles of TMA at ~3.5 cycles per row against 1 700 cycles of MMA (993 us for backbone.conv2.0 in
            // the parity-mode timeline); the plane boxes are 1 224 rows per tile.  Needs room for a ring of >= 8 weight blocks.
            const size_t need = halo2 + (size_t)8 * (bn / 2) * bk * 2 + 8 * 1024;
            if (need > 227 * 1024) p.halo_s2 = 0;
        }
    }
    if (p.halo_s2) p.halo = 1;
    p.halo_planes = p.halo_s2 ? 4 : 1;
    if (p.halo) { p.tw = 8; p.th = 16; p.tiles_x = (d.out_w + 7) / 8; }
    else if (p.flat) { p.tw = kBlockM; p.th = 1; p.tiles_x = 1; }
    else { p.tw = pick_tile_w(d.out_w, d.out_kind == OM_OUT_NCHW); p.th = kBlockM / p.tw; p.tiles_x = d.out_w / p.tw; }
    p.total_rows = d.batch * d.out_rows;
    const int tiles_y = p.flat ? (p.flat_total + kBlockM - 1) / kBlockM : (p.total_rows + p.th - 1) / p.th;
    p.pairs_y = (tiles_y + 1) / 2;
    p.d_tiles_n = om::make_fastdiv(p.tiles_n); p.d_tiles_x = om::make_fastdiv(p.tiles_x); p.d_out_rows = om::make_fastdiv(d.out_rows);
    p.d_tw = om::make_fastdiv(p.tw); p.d_flat_hw = om::make_fastdiv(p.flat_hw); p.d_out_w = om::make_fastdiv(d.out_w);
    p.taps = d.ksize * d.ksize; p.stride = d.stride;
    p.k_chunks = (split ? 3 : 1) * (d.cin / bk); p.a_chunks = (split ? 2 : 1) * (d.cin / bk); p.split_kr = split ? d.cin / bk : 0;
    p.d_k_chunks = om::make_fastdiv(p.k_chunks);
    p.acc_scale = (split && d.acc_scale != 0.0f) ? d.acc_scale : 1.0f;
    if (split && !getenv("ORIENMASK_B200_NO_GAIN_FIX")) {
        // tcgen05 truncates its fp32 running sum: every MMA into a non-empty accumulator loses a fraction of an ulp toward zero, which
        // for a sum that grows from zero is, on average, a common gain error of -0.262 * 2^-24 per MMA step of the hi*hi pass (measured
        // for K = 64 .. 9216, tools/split_probe.py, profiles/r02_split_probe.json: -0.258 .. -0.30 per step, the same for both
        // tensor-core engines).  Undoing the mean halves the layer error; what remains is its element-dependent part.
        const int steps = d.ksize * d.ksize * (d.cin / 16);
        p.acc_scale *= 1.0f + 0.262f * (float)steps * 5.9604645e-8f;
    }
    p.pix_stride = split ? 2 * d.cout_stride : d.cout_stride; p.lo_off = d.cout_stride;
    {
        // the fp16 residual: staged by TMA for the memory-bound layers (its DRAM latency must be covered several chunks ahead),
        // read by the epilogue threads themselves in flat mode and (experiment: ORIENMASK_B200_RESDIRECT=1) on tensor-bound tiles
        const char* rd = getenv("ORIENMASK_B200_RESDIRECT");
        const int cycles = d.ksize * d.ksize * (d.cin / bk) * (bk / 16) * (bn / 2);
        const bool direct = p.flat || split || (rd && rd[0] == '1' && cycles >= 8192);
        p.res_direct = (d.residual != nullptr && direct) ? 1 : 0;
        p.residual = d.residual;
        p.has_res = (d.residual != nullptr && !direct) ? 1 : 0;
    }
    p.up_bw = p.tw / 2 + 1; p.up_bh = p.th / 2 + 1;
    if (d.upadd != nullptr && !p.has_res && d.up_rows * 2 == d.out_rows && d.cout % 32 == 0 && p.up_bw * p.up_bh <= kStageRows)
        p.has_res = 2;
    // staged chunk: fp16 residual 64 columns = 128-byte rows (64-byte TMA rows measured slower), everything else 32 columns
    // (only for the memory-bound small-K layers: the tensor-bound ones would rather keep the shared memory for operand stages)
    const int tile_cycles_est = p.taps * p.k_chunks * (bk / 16) * (bn / 2);
    p.chunk_cols = (p.has_res == 1 && bn >= 64 && tile_cycles_est < 8192) ? 64 : 32; p.row_bytes = p.has_res == 1 ? p.chunk_cols * 2 : 128;
    p.res_buf_bytes = kStageRows * p.row_bytes;
    p.halo_pitch = p.halo_s2 ? p.tw + 1 : p.tw + 2;
    p.a_box_pixels = p.halo_s2 ? (p.tw + 1) * (p.th + 1) : p.halo ? (p.tw + 2) * (p.th + 2) : p.tw * p.th;
    p.h_chunk_bytes = p.halo ? ((p.a_box_pixels * bk * 2 + 1023) / 1024) * 1024 : 0;
    p.h_stage_bytes = p.h_chunk_bytes * p.a_chunks * p.halo_planes;
    p.h_stages = p.halo ? 2 : 0;
    if (p.halo && getenv("ORIENMASK_B200_HSTAGES")) {
        const int v = atoi(getenv("ORIENMASK_B200_HSTAGES"));
        if (v >= 2 && v <= 4) p.h_stages = v;
    }
    // per-tap activation block: the tw*th-row box rounded up to the 1024-byte swizzle period (the MMA reads 128 rows; rows
    // beyond the box alias the following weight block and only feed accumulator rows the epilogue masks)
    p.a_sub_bytes = p.halo ? 0 : ((p.tw * p.th * bk * 2 + 1023) / 1024) * 1024;
    const int sub_bytes = p.a_sub_bytes + p.half_n * bk * 2;
    // addend prefetch depth: memory-bound tiles (few tensor cycles per tile) need the addend several chunks ahead
    {
        const int tile_cycles = p.taps * p.k_chunks * (bk / 16) * (bn / 2);
        p.res_depth = !p.has_res ? 1 : (tile_cycles < 8192 && p.res_buf_bytes <= 8192) ? 2 : 1;   // kEpiGroups chunks are in flight anyway
        const char* rd_env = getenv("ORIENMASK_B200_RESDEPTH");
        if (p.has_res && rd_env && atoi(rd_env) >= 1 && atoi(rd_env) <= 2) p.res_depth = atoi(rd_env);
    }
    p.res_bufs = kEpiGroups * p.res_depth;
    p.res_bufs_log2 = p.res_bufs == 4 ? 2 : 3;
    const int epi_bytes = p.has_res ? p.res_bufs * p.res_buf_bytes : 0;
    constexpr int kMaxSmem = 227 * 1024;
    const int fixed = 1024 + epi_bytes + (4 * kMaxStages + 2 * kMaxAcc + 2 * kMaxResBufs) * 8 + 16 + cout_pad * 4;
    {
        const int total_b = p.taps * p.k_chunks * p.half_n * bk * 2;
        const char* br_env = getenv("ORIENMASK_B200_BRES");
        p.b_resident = p.halo && p.tiles_n == 1 && total_b <= 96 * 1024 && !(br_env && br_env[0] == '0') && !split;
        if (p.b_resident) {
            int hs = (kMaxSmem - fixed - total_b) / p.h_stage_bytes;
            p.h_stages = hs > 4 ? 4 : hs;
            if (p.h_stages < 2) p.b_resident = 0, p.h_stages = 2;
        }
    }
    if (split && p.halo && p.h_stages == 2 && !(getenv("ORIENMASK_B200_HSTAGES")) &&
        kMaxSmem - fixed - 2 * p.h_stage_bytes < 4 * sub_bytes && kMaxSmem - fixed - p.h_stage_bytes >= 4 * sub_bytes) {
        // split precision doubles the halo (hi | lo chunks): two halo stages of a 128-channel layer leave room for two weight blocks only,
        // and the MMA warp then waits on a barrier round trip per block (ncu, 3x3 128->256 @136x136: tensor pipe 50 % active).  One halo
        // stage + a deep weight ring keeps the tensor pipe fed; the price is one exposed halo load per tile (~2 us of ~28 us).
        p.h_stages = 1;
    }
    const int ring_budget = kMaxSmem - fixed - p.h_stages * p.h_stage_bytes;
    if (ring_budget < 2 * sub_bytes) { delete plan; return fail(OM_ERR_INVALID, "tile does not fit shared memory"); }
    // blocks per stage: enough MMA cycles behind each barrier round trip (>= ~1024 tensor cycles; one block issues
    // BK/16 MMAs of N/2 cycles each), while keeping at least three stages in flight
    {
        const int total = p.taps * p.k_chunks;
        const int block_cycles = (bk / 16) * (bn / 2);
        const char* ns_env = getenv("ORIENMASK_B200_NSUB");
        int best = 1;
        for (int n = 1; n <= total && n <= 32; ++n) {     // lane j of the producer warp issues block j of a stage
            if (total % n) continue;
            const int st = ring_budget / (n * sub_bytes);
            if (st < 3 && n > 1 && !(n == total && st >= 2)) break;   // a whole tile per stage may run double-buffered
            best = n;
            if (n * block_cycles >= 1024) break;
        }
        if (ns_env && atoi(ns_env) > 0 && atoi(ns_env) <= 32 && total % atoi(ns_env) == 0 && ring_budget / (atoi(ns_env) * sub_bytes) >= 2) best = atoi(ns_env);
        p.n_sub = best;
        p.stages = ring_budget / (best * sub_bytes);
        if (p.stages > kMaxStages) p.stages = kMaxStages;
        if (getenv("ORIENMASK_B200_MAXSTAGES") && atoi(getenv("ORIENMASK_B200_MAXSTAGES")) >= 2 && p.stages > atoi(getenv("ORIENMASK_B200_MAXSTAGES")))
            p.stages = atoi(getenv("ORIENMASK_B200_MAXSTAGES"));           // experiment: sensitivity to the depth of the block ring
        if (p.stages < 2) { delete plan; return fail(OM_ERR_INVALID, "tile does not fit shared memory"); }
        if (p.b_resident) { p.n_sub = total; p.stages = 1; }          // the "ring" is the resident weight tensor
        // one halo stage of a split-precision layer: lo and hi halves behind their own barriers, when the passes are whole groups
        const char* hh = getenv("ORIENMASK_B200_HHALVES");
        p.h_halves = (split && p.halo && !p.halo_s2 && p.h_stages == 1 && !p.b_resident && (p.taps * p.split_kr) % p.n_sub == 0 &&
                      !(hh && hh[0] == '0')) ? 1 : 0;
    }
    // 4 accumulators at most: 2 / 4 / 8 in flight measured identical on the narrow memory-bound layers (ORIENMASK_B200_ACC, up to
    // kMaxAcc, for experiments), and a smaller TMEM allocation lets the next layer's prologue allocate earlier
    p.acc_stages = 512 / bn >= 4 ? 4 : 512 / bn;
    if (getenv("ORIENMASK_B200_ACC") && atoi(getenv("ORIENMASK_B200_ACC")) >= 2 && atoi(getenv("ORIENMASK_B200_ACC")) <= kMaxAcc &&
        atoi(getenv("ORIENMASK_B200_ACC")) * bn <= 512)
        p.acc_stages = atoi(getenv("ORIENMASK_B200_ACC"));
    int cols = 32;
    while (cols < p.acc_stages * bn) cols <<= 1;
    p.tmem_cols = cols;
    // cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b F16, K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 per pair)
    p.idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((2 * kBlockM) >> 4) << 24);
    p.out_h = d.out_h; p.out_w = d.out_w; p.out_rows = d.out_rows;
    p.cout = d.cout; p.leaky = d.leaky; p.out_kind = d.out_kind;
    p.up_rows = d.up_rows; p.bias = d.bias; p.upadd = d.upadd;
    p.cout_stride = d.cout_stride; p.output = d.output;
    p.out_s2d = d.out_s2d; p.s2d_plane = (long long)d.batch * d.out_rows / 2 * (d.out_w / 2);
    {
        const char* se = getenv("ORIENMASK_B200_EPI_SLEEP");
        p.epi_sleep_ns = se ? atoi(se) : 250;        // time-neutral in the interleaved A/B (100 .. 1000 ns), +0.1 .. 0.7 % in 60-step runs under the
        if (p.epi_sleep_ns < 0 || p.epi_sleep_ns > 100000) p.epi_sleep_ns = 0;   // power cap (fewer instructions per joule-limited step)
    }
    {
        // L2 eviction priorities (experiment: ORIENMASK_B200_L2HINT=mode)
        const char* he = getenv("ORIENMASK_B200_L2HINT");
        const int mode = he ? atoi(he) : 0;
        const bool res_layer = d.residual != nullptr && d.out_kind == OM_OUT_ACT;
        const bool squeeze = d.ksize == 1 && d.stride == 1 && d.out_kind == OM_OUT_ACT && d.cout < d.cin && d.upadd == nullptr;
        if (mode == 1 || mode == 2) { if (res_layer) p.hint_a = 1; }
        if (mode == 1 || mode == 3) { if (res_layer) p.hint_out = 2; if (squeeze) p.hint_a = 2; }
        if (mode == 4) { if (res_layer) { p.hint_a = 1; p.hint_res = 1; } }
        if (mode == 5) { p.hint_w = 2; }
        if (mode == 6) { if (res_layer) { p.hint_a = 1; } p.hint_w = 2; }
    }

    const size_t esz = 2;
    int32_t rc = OM_OK;
    const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (p.flat) {
        rc = encode_im2col(&plan->maps.m[0], d.input, cin_mem, cin_mem, d.in_w, d.in_h, d.in_rows, d.batch, bk, d.ksize, d.stride, bk * 2);
        for (int i = 1; i < 4 && rc == OM_OK; ++i) plan->maps.m[i] = plan->maps.m[0];
    } else if (d.stride == 1) {
        cuuint64_t dims[3] = {(cuuint64_t)cin_mem, (cuuint64_t)d.in_w, (cuuint64_t)d.batch * d.in_rows};
        cuuint64_t str[2] = {(cuuint64_t)cin_mem * esz, (cuuint64_t)d.in_w * cin_mem * esz};
        cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(p.halo ? p.tw + 2 : p.tw), (cuuint32_t)(p.halo ? p.th + 2 : p.th)};
        rc = encode(&plan->maps.m[0], f16, d.input, 3, dims, str, box, bk * 2);
        for (int i = 1; i < 4 && rc == OM_OK; ++i) plan->maps.m[i] = plan->maps.m[0];
    } else {
        for (int sel = 0; sel < 4 && rc == OM_OK; ++sel) {          // sel = 2*odd_row + odd_col
            const int py = sel >> 1, px = sel & 1;
            cuuint64_t dims[3] = {(cuuint64_t)cin_mem, (cuuint64_t)d.in_w / 2, (cuuint64_t)d.batch * d.in_rows / 2};
            cuuint64_t str[2] = {(cuuint64_t)2 * cin_mem * esz, (cuuint64_t)2 * d.in_w * cin_mem * esz};
            cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(p.halo_s2 ? p.tw + 1 : p.tw), (cuuint32_t)(p.halo_s2 ? p.th + 1 : p.th)};
            const char* base = reinterpret_cast<const char*>(d.input) + ((size_t)py * d.in_w + px) * cin_mem * esz;
            if (d.in_s2d) {                                   // dense parity planes: every tap is a contiguous box
                str[0] = (cuuint64_t)cin_mem * esz; str[1] = (cuuint64_t)(d.in_w / 2) * cin_mem * esz;
                base = reinterpret_cast<const char*>(d.input) + (size_t)sel * ((size_t)d.batch * d.in_rows / 2 * (d.in_w / 2)) * cin_mem * esz;
            }
            rc = encode(&plan->maps.m[sel], f16, base, 3, dims, str, box, bk * 2);
        }
    }
    if (rc == OM_OK) {
        cuuint64_t dims[2] = {(cuuint64_t)cin_mem, (cuuint64_t)p.taps * cout_pad};
        cuuint64_t str[1] = {(cuuint64_t)cin_mem * esz};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)p.half_n};
        rc = encode(&plan->map_b, f16, d.weights, 2, dims, str, box, bk * 2);
    }
    if (rc == OM_OK && p.has_res == 2) {
        cuuint64_t dims[3] = {(cuuint64_t)d.cout, (cuuint64_t)d.out_w / 2, (cuuint64_t)d.batch * d.up_rows};
        cuuint64_t str[2] = {(cuuint64_t)d.cout * 4, (cuuint64_t)(d.out_w / 2) * d.cout * 4};
        cuuint32_t box[3] = {32u, (cuuint32_t)p.up_bw, (cuuint32_t)p.up_bh};
        rc = encode(&plan->map_res, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.upadd, 3, dims, str, box, 128);
    }
    if (rc == OM_OK && p.has_res == 1 && p.flat) {
        rc = encode_im2col(&plan->map_res, d.residual, d.cout, d.cout_stride, d.out_w, d.out_h, d.out_rows, d.batch, p.chunk_cols, 1, 1, p.row_bytes);
    } else if (rc == OM_OK && p.has_res == 1) {
        cuuint64_t dims[3] = {(cuuint64_t)d.cout, (cuuint64_t)d.out_w, (cuuint64_t)d.batch * d.out_rows};
        cuuint64_t str[2] = {(cuuint64_t)d.cout_stride * esz, (cuuint64_t)d.out_w * d.cout_stride * esz};
        cuuint32_t box[3] = {(cuuint32_t)p.chunk_cols, (cuuint32_t)p.tw, (cuuint32_t)p.th};
        rc = encode(&plan->map_res, f16, d.residual, 3, dims, str, box, p.row_bytes);
    }
    if (rc == OM_OK && !p.has_res) plan->map_res = plan->maps.m[0];
    if (rc != OM_OK) { delete plan; return rc; }

    plan->smem = (size_t)fixed + (size_t)p.h_stages * p.h_stage_bytes + (size_t)p.stages * p.n_sub * sub_bytes;
    const int sms = sm_count();
    const int pairs = p.tiles_x * p.pairs_y * p.tiles_n;
    const int clusters = pairs < sms / 2 ? pairs : sms / 2;
    plan->grid = 2 * clusters;
    if (plan->smem > (size_t)kMaxSmem || p.tmem_cols > 512 || plan->grid < 2) {       // cannot happen by construction; the dry-run sweep proves it
        const size_t need = plan->smem;
        const int cols = p.tmem_cols;
        delete plan;
        return fail(OM_ERR_INVALID, "plan out of budget: %zu bytes of shared memory, %d TMEM columns", need, cols);
    }
    plan->split = split;
    cudaError_t e = cudaSuccess;
    if (!plan_dryrun()) {
        if (split) e = bk == 64 ? cudaFuncSetAttribute(conv_tc2_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
                                : cudaFuncSetAttribute(conv_tc2_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        else e = bk == 64 ? cudaFuncSetAttribute(conv_tc2_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
                          : cudaFuncSetAttribute(conv_tc2_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    }
    if (e != cudaSuccess) { delete plan; return fail(OM_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); }
    *out = plan;
    return OM_OK;
}

int32_t tc2_plan_create(const om_conv_desc& d, void** out) {
    bool halo_used = false;
    int32_t rc = plan_create(d, out, true, &halo_used);
    // The halo box of a wide-K 3x3 layer (cin >= 256: 94 KB per stage) does not fit next to a staged residual; such a layer is
    // planned without the halo (flat or per-tap tiles, the modes the 17x17 / 34x34 layers use).  Found by tests/test_planner.py:
    // widths whose stride-16 map is a multiple of 8 wide (640, 1024, 1280, ...) could not be planned at all.
    if (rc == OM_ERR_INVALID && halo_used && strstr(om::error_buffer(), "does not fit shared memory") != nullptr)
        rc = plan_create(d, out, false, &halo_used);
    return rc;
}

// `output` != nullptr redirects the result (the epilogue stores through a plain pointer, no tensor map): the model's head layers
// write into tensors allocated per call, so results handed to the caller are never overwritten by the next forward.
int32_t tc2_plan_run(const void* vp, cudaStream_t stream, void* output) {
    const Tc2Plan* plan = reinterpret_cast<const Tc2Plan*>(vp);
    if (plan_dryrun()) return fail(OM_ERR_UNSUPPORTED, "plans made under ORIENMASK_B200_PLAN_DRYRUN cannot be launched");
    Tc2Params p = plan->p;
    p.trace = trace_next();
    if (output != nullptr) {
        if (p.has_res || p.res_direct) return fail(OM_ERR_INVALID, "om_conv_run_to: layers with a residual write in place");
        p.output = output;
    }
    if (plan->split) {
        if (plan->bk == 64)
            OM_CUDA_TRY(launch_pdl(conv_tc2_kernel<64, true>, dim3(plan->grid), dim3(kThreads), plan->smem, stream, plan->maps, plan->map_b, plan->map_res, p));
        else
            OM_CUDA_TRY(launch_pdl(conv_tc2_kernel<32, true>, dim3(plan->grid), dim3(kThreads), plan->smem, stream, plan->maps, plan->map_b, plan->map_res, p));
    } else if (plan->bk == 64)
        OM_CUDA_TRY(launch_pdl(conv_tc2_kernel<64, false>, dim3(plan->grid), dim3(kThreads), plan->smem, stream, plan->maps, plan->map_b, plan->map_res, p));
    else
        OM_CUDA_TRY(launch_pdl(conv_tc2_kernel<32, false>, dim3(plan->grid), dim3(kThreads), plan->smem, stream, plan->maps, plan->map_b, plan->map_res, p));
    return check_launch("conv_tc2_kernel");
}

void tc2_plan_destroy(void* vp) { delete reinterpret_cast<Tc2Plan*>(vp); }

void tc2_plan_info(const void* vp, int32_t* info) {
    const Tc2Plan* plan = reinterpret_cast<const Tc2Plan*>(vp);
    const Tc2Params& p = plan->p;
    const int32_t v[24] = {p.halo, p.flat, p.halo_s2, p.b_resident, p.tw, p.th, p.block_n, p.tiles_n, p.stages, p.n_sub, p.h_stages,
                           p.acc_stages, p.has_res, p.res_direct, (int32_t)plan->smem, plan->grid,
                           p.tiles_x * p.pairs_y * p.tiles_n, p.taps, p.k_chunks, plan->bk, p.tmem_cols, p.cout, p.out_h, p.out_w};
    memcpy(info, v, sizeof(v));
}

int32_t tc2_set_wait_hint(unsigned int ns) {
    OM_CUDA_TRY(cudaMemcpyToSymbol(g_wait_hint_ns, &ns, sizeof(ns)));
    return OM_OK;
}

int32_t tc2_set_timeline(void* dev_ptr) {
    OM_CUDA_TRY(cudaMemcpyToSymbol(g_timeline, &dev_ptr, sizeof(void*)));
    return OM_OK;
}

}  // namespace om
