// Detections -> COCO segmentation format: the step right after the post-process in test.py / infer.py -j.
//
//   reference                                                            here
//   ---------                                                            ----
//   COCOMetrics._recover_shape_segm   eval/coco_eval.py:191-205          crop the padding, flips, F.interpolate(bilinear,
//                                                                         align_corners=False) to the original size, round()
//   COCOMetrics._to_segm_coco_format  eval/coco_eval.py:108-127          per instance: mask.cpu().numpy() (a D2H copy and a sync each),
//                                                                         maskUtils.encode(np.asfortranarray(mask)), counts.decode()
//   pycocotools maskApi.c (cocoapi, unpinned `pycocotools` of requirements.txt; not vendored by the reference):
//     rleEncode    column-major run lengths, first run counts zeros            -> phase 1-3 below
//     rleToString  5-bit groups, LEB128-like, delta against counts[i-2] for i>2 -> phase 4 below
//
// One CTA per instance.  The resized mask never exists: a thread walks one output column, evaluates the bilinear blend of
// the four source mask bytes, thresholds it exactly like round() (> 0.5; a tie rounds to even = 0) and only records where
// the column-major bit stream changes.  Output per instance: the run-length counts and the compressed COCO string --
// a few hundred bytes leave the GPU instead of H*W bytes per instance.
//
// Bilinear arithmetic: the same single-rounded fp32 sequence as prep.cu / oracle/prep_oracle.py (ATen's expression):
// src = max(fma(scale, d + 0.5, -0.5), 0); l1 = src - i0; l0 = 1 - l1; v = fma(l0y, fma(l0x, a, l1x*b), l1y * fma(l0x, c, l1x*d)).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void src_index(float scale, int d, int n_in, int& i0, int& i1, float& l0, float& l1) {
    float s = __fmaf_rn(scale, __fadd_rn((float)d, 0.5f), -0.5f);
    s = fmaxf(s, 0.0f);
    i0 = min((int)s, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    l1 = __fsub_rn(s, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

struct ColCtx {
    const unsigned char* c0;       // source column x0 (already cropped / flipped), row stride `pitch`
    const unsigned char* c1;       // source column x1
    float lx0, lx1;
};

__device__ __forceinline__ int resized_bit(const ColCtx& c, long long r0, long long r1, float ly0, float ly1) {
    const float a = (float)__ldg(c.c0 + r0), b = (float)__ldg(c.c1 + r0);
    const float cc = (float)__ldg(c.c0 + r1), d = (float)__ldg(c.c1 + r1);
    const float top = __fmaf_rn(c.lx0, a, __fmul_rn(c.lx1, b));
    const float bot = __fmaf_rn(c.lx0, cc, __fmul_rn(c.lx1, d));
    const float v = __fmaf_rn(ly0, top, __fmul_rn(ly1, bot));
    return v > 0.5f ? 1 : 0;                     // torch.round(): half to even, and v is in [0, 1]
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                              // protect warp_sums from the previous use
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const int s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

// y table in shared memory: source rows (flip applied, relative to the crop) and the weights, per output row
struct YTab { int r0, r1; float l0, l1; };

// One output column as 32-row words of the column-major bit stream.  For every word the CHANGE mask (bit i set = the stream
// changes at row 32*w + i) goes to `tw[w * kThreads]` (shared memory, one column per thread) and the number of changes is
// returned, so the second pass only iterates set bits instead of re-evaluating the blend.  Rows outside [ya, yb] are zero
// without any load; consecutive output rows share source rows (y1 of one row is y0 of the next when resizing by ~1x), so the
// horizontal blend of a source row is reused instead of reloaded.
__device__ __forceinline__ int column_changes(const ColCtx& c, bool occupied, int prev, int oh, int ya, int yb, const YTab* ytab,
                                              long long pitch, unsigned int* tw) {
    const int nw = (oh + 31) >> 5;
    int cnt = 0;
    int rA = -1, rB = -1; float vA = 0.f, vB = 0.f;         // two cached source rows and their horizontal blends
    for (int w = 0; w < nw; ++w) {
        unsigned int word = 0;
        const int y_lo = w << 5, y_hi = min(y_lo + 32, oh);
        if (occupied && y_lo <= yb && y_hi > ya) {
            for (int y = max(y_lo, ya); y < min(y_hi, yb + 1); ++y) {
                const YTab t = ytab[y];
                float top, bot;
                if (t.r0 == rA) top = vA; else if (t.r0 == rB) top = vB;
                else top = __fmaf_rn(c.lx0, (float)__ldg(c.c0 + t.r0 * pitch), __fmul_rn(c.lx1, (float)__ldg(c.c1 + t.r0 * pitch)));
                if (t.r1 == t.r0) bot = top; else if (t.r1 == rA) bot = vA; else if (t.r1 == rB) bot = vB;
                else bot = __fmaf_rn(c.lx0, (float)__ldg(c.c0 + t.r1 * pitch), __fmul_rn(c.lx1, (float)__ldg(c.c1 + t.r1 * pitch)));
                rA = t.r0; vA = top; rB = t.r1; vB = bot;
                const float v = __fmaf_rn(t.l0, top, __fmul_rn(t.l1, bot));
                word |= (v > 0.5f ? 1u : 0u) << (y - y_lo);          // torch.round(): half to even, and v is in [0, 1]
            }
        }
        unsigned int change = word ^ ((word << 1) | (unsigned int)prev);
        if (y_hi - y_lo < 32) change &= (1u << (y_hi - y_lo)) - 1u;          // the column ends inside this word
        tw[w * kThreads] = change;
        cnt += __popc(change);
        prev = (int)(word >> 31);                           // rows past `oh` in the last word are 0: no column follows inside it
    }
    return cnt;
}

__global__ void __launch_bounds__(kThreads) mask_rle_kernel(const om_rle_image* __restrict__ images, int max_inst, int cap, int str_cap,
                                                            int change_off,
                                                            unsigned int* __restrict__ counts, int* __restrict__ n_counts,
                                                            unsigned char* __restrict__ str, int* __restrict__ str_len) {
    extern __shared__ __align__(16) unsigned char rle_smem[];
    __shared__ int warp_sums[kThreads / 32];
    __shared__ int s_ya, s_yb;
    const int b = blockIdx.y, k = blockIdx.x;
    const om_rle_image im = images[b];
    const long long inst = (long long)b * max_inst + k;
    if (k >= im.count) return;
    const int oh = im.out_h, ow = im.out_w;
    YTab* ytab = reinterpret_cast<YTab*>(rle_smem);                                  // [oh]
    unsigned char* col_any = rle_smem + (size_t)oh * sizeof(YTab);                   // [mask_w]  any pixel set in this mask column
    unsigned char* row_any = col_any + ((im.mask_w + 15) & ~15);                     // [mask_h]  ... in this mask row
    unsigned int* changes = reinterpret_cast<unsigned int*>(rle_smem + change_off) + threadIdx.x;   // [(max_oh+31)/32][kThreads]
    const unsigned char* mk = im.mask + (long long)k * im.mask_h * im.mask_w;
    const unsigned char* m = mk + (long long)im.top * im.mask_w + im.left;           // crop origin
    const long long pitch = im.mask_w;

    // ---- phase 0: which rows / columns of the (cropped) source mask hold any set pixel.  Instance masks cover a few
    // percent of the image: everything outside the occupied band is zero and is never walked. ----
    for (int i = threadIdx.x; i < im.mask_w + im.mask_h; i += kThreads) (i < im.mask_w ? col_any[i] : row_any[i - im.mask_w]) = 0;
    if (threadIdx.x == 0) { s_ya = oh; s_yb = -1; }
    __syncthreads();
    const int segs = im.mask_w >> 4;
    if ((im.mask_w & 15) == 0 && (reinterpret_cast<uintptr_t>(mk) & 15) == 0 && segs <= kThreads) {
        const int rows_per_pass = kThreads / segs;
        if ((int)threadIdx.x < rows_per_pass * segs) {
            const int seg = threadIdx.x % segs;
            uint4 acc = make_uint4(0, 0, 0, 0);
            for (int r = im.top + (int)threadIdx.x / segs; r < im.top + im.crop_h; r += rows_per_pass) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(mk + (long long)r * pitch) + seg);
                if (v.x | v.y | v.z | v.w) row_any[r] = 1;
                acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
            }
            const unsigned int w4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if ((w4[i >> 2] >> ((i & 3) * 8)) & 0xffu) col_any[seg * 16 + i] = 1;
        }
    } else {
        for (long long i = threadIdx.x; i < (long long)im.crop_h * im.crop_w; i += kThreads) {
            const int r = (int)(i / im.crop_w), cx = (int)(i - (long long)r * im.crop_w);
            if (__ldg(m + (long long)r * pitch + cx)) { row_any[im.top + r] = 1; col_any[im.left + cx] = 1; }
        }
    }
    __syncthreads();
    const float scale_y = (float)im.crop_h / (float)oh, scale_x = (float)im.crop_w / (float)ow;
    {
        int ya = oh, yb = -1;
        for (int y = threadIdx.x; y < oh; y += kThreads) {
            int y0, y1; float l0, l1;
            src_index(scale_y, y, im.crop_h, y0, y1, l0, l1);
            if (im.vflip) { y0 = im.crop_h - 1 - y0; y1 = im.crop_h - 1 - y1; }
            ytab[y].r0 = y0; ytab[y].r1 = y1; ytab[y].l0 = l0; ytab[y].l1 = l1;
            if (row_any[im.top + y0] | row_any[im.top + y1]) { ya = min(ya, y); yb = max(yb, y); }
        }
        if (yb >= 0) { atomicMin(&s_ya, ya); atomicMax(&s_yb, yb); }
    }
    __syncthreads();
    const int ya = s_ya, yb = s_yb;               // occupied output rows (empty mask: ya > yb)
    unsigned int* out = counts + inst * cap;
    int n_trans = 0;                              // transitions found so far (all threads hold the same value)
    for (int x_base = 0; x_base < ow && yb >= 0; x_base += kThreads) {
        const int x = x_base + threadIdx.x;
        ColCtx c; c.c0 = m; c.c1 = m; c.lx0 = 1.f; c.lx1 = 0.f;
        int cnt = 0, first_prev = 0;
        bool occ = false;
        if (x < ow) {
            int x0, x1;
            src_index(scale_x, x, im.crop_w, x0, x1, c.lx0, c.lx1);
            if (im.hflip) { x0 = im.crop_w - 1 - x0; x1 = im.crop_w - 1 - x1; }
            c.c0 = m + x0; c.c1 = m + x1;
            occ = (col_any[im.left + x0] | col_any[im.left + x1]) != 0;
            // the bit that precedes this column in column-major order: last row of column x - 1 (0 before the first pixel,
            // and 0 whenever the last output row or that column is unoccupied)
            if (x > 0 && yb == oh - 1) {
                ColCtx pc;
                int px0, px1;
                src_index(scale_x, x - 1, im.crop_w, px0, px1, pc.lx0, pc.lx1);
                if (im.hflip) { px0 = im.crop_w - 1 - px0; px1 = im.crop_w - 1 - px1; }
                if (col_any[im.left + px0] | col_any[im.left + px1]) {
                    pc.c0 = m + px0; pc.c1 = m + px1;
                    const YTab t = ytab[oh - 1];
                    first_prev = resized_bit(pc, t.r0 * pitch, t.r1 * pitch, t.l0, t.l1);
                }
            }
            cnt = column_changes(c, occ, first_prev, oh, ya, yb, ytab, pitch, changes);
        }
        int total;
        int w = n_trans + block_exclusive_scan(cnt, warp_sums, total);
        if (cnt > 0) {                            // positions of the set bits, in stream order
            const unsigned int base = (unsigned int)x * (unsigned int)oh;
            const int nw = (oh + 31) >> 5;
            for (int wi = 0; wi < nw; ++wi) {
                unsigned int ch = changes[wi * kThreads];
                while (ch) {
                    const int i = __ffs(ch) - 1;
                    ch &= ch - 1;
                    if (w < cap) out[w] = base + (unsigned int)(wi * 32 + i);
                    ++w;
                }
            }
        }
        n_trans += total;
    }
    // positions -> run lengths in place: counts = [p0, p1 - p0, ..., N - p_last]  (rleEncode).  Chunks are converted from
    // the end so that a chunk only ever reads positions that have not been overwritten yet.
    const int n = n_trans + 1;
    if (threadIdx.x == 0) n_counts[inst] = n;
    if (n > cap) { if (threadIdx.x == 0) str_len[inst] = -1; return; }      // the caller retries with cap >= n_counts
    __syncthreads();                              // positions written by other threads of this CTA
    const unsigned int N = (unsigned int)oh * (unsigned int)ow;
    for (int base = ((n - 1) / kThreads) * kThreads; base >= 0; base -= kThreads) {
        const int i = base + threadIdx.x;
        unsigned int cur = 0, prv = 0;
        if (i < n) {
            cur = i < n_trans ? out[i] : N;
            prv = i > 0 ? out[i - 1] : 0u;
        }
        __syncthreads();
        if (i < n) out[i] = cur - prv;
    }
    __syncthreads();
    // rleToString: x = counts[i] - (i > 2 ? counts[i-2] : 0); 5 bits per character, bit 5 = "more", sign-extended groups
    unsigned char* sp = str + inst * str_cap;
    int written = 0;
    for (int base = 0; base < n; base += kThreads) {
        const int i = base + threadIdx.x;
        long long x = 0;
        int len = 0;
        if (i < n) {
            x = (long long)out[i];
            if (i > 2) x -= (long long)out[i - 2];
            long long t = x;
            bool more = true;
            while (more) {
                const int c = (int)(t & 0x1f);
                t >>= 5;
                more = (c & 0x10) ? (t != -1) : (t != 0);
                ++len;
            }
        }
        int total;
        int w = written + block_exclusive_scan(len, warp_sums, total);
        if (i < n && w + len <= str_cap) {
            bool more = true;
            while (more) {
                int c = (int)(x & 0x1f);
                x >>= 5;
                more = (c & 0x10) ? (x != -1) : (x != 0);
                if (more) c |= 0x20;
                sp[w++] = (unsigned char)(c + 48);
            }
        }
        written += total;
    }
    if (threadIdx.x == 0) str_len[inst] = written <= str_cap ? written : -1;
}

}  // namespace

extern "C" int32_t om_mask_rle(const om_rle_image* images, int32_t batch, int32_t max_inst, int32_t max_out_h, int32_t max_mask_h,
                               int32_t max_mask_w, int32_t cap,
                               int32_t str_cap, uint32_t* counts, int32_t* n_counts, uint8_t* str, int32_t* str_len, void* stream) {
    if (!images || !counts || !n_counts || !str || !str_len) return om::fail(OM_ERR_INVALID, "om_mask_rle: null argument");
    if (batch < 1 || batch > 65535 || max_inst < 1 || cap < 1 || str_cap < 1 || max_out_h < 1 || max_mask_h < 1 || max_mask_w < 1)
        return om::fail(OM_ERR_INVALID, "om_mask_rle: non-positive size (or batch > 65535)");
    // per-CTA shared memory: row table of the tallest output + occupancy flags of the largest mask (host copy of the sizes)
    const size_t change_off = ((size_t)max_out_h * 16 + (size_t)((max_mask_w + 15) & ~15) + (size_t)max_mask_h + 15) & ~(size_t)15;
    const size_t smem = change_off + (size_t)((max_out_h + 31) / 32) * kThreads * 4;
    if (smem > 200 * 1024) return om::fail(OM_ERR_UNSUPPORTED, "om_mask_rle: sizes need %zu bytes of shared memory per CTA (max 204800)", smem);
    if (smem > 48 * 1024)        // per device (a process may drive several GPUs): set on every call that needs it, like allow_smem() in post.cu
        OM_CUDA_TRY(cudaFuncSetAttribute(mask_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    mask_rle_kernel<<<dim3((unsigned)max_inst, (unsigned)batch), kThreads, smem, (cudaStream_t)stream>>>(
        images, max_inst, cap, str_cap, (int)change_off, counts, n_counts, str, str_len);
    return om::check_launch("mask_rle_kernel");
}
