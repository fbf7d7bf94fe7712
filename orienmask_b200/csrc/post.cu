// Post-process kernels: box/score decode + radix top-k select, class-wise greedy NMS, mask assembly.
//
// Compiled with -fmad=false: every arithmetic step that decides a comparison (scores, IoU, mask
// thresholds) is a single-rounded fp32 operation in the order the reference performs it
// (eval/orienmask_yolo_postprocess.py:126-166, eval/function.py:94-97, eval/src/nms_cpu.cpp:17-59);
// the only fused operations are the explicit __fmaf_rn calls of the bilinear x4 interpolation,
// which reproduce ATen's CPU kernel bit-for-bit (see oracle/post_oracle.py).
//
// All kernels are HBM/L2-bound integer+fp32 work: coalesced plane-strided reads of the NCHW heads,
// 16-byte streaming stores for the masks, warp shuffles/ballots for scans, no tensor cores.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kSelThreads = 256;
constexpr int kHistBins = 2048;
constexpr int kL0Bins = 256;                  // level-0 digit: 8 bits (every block of the head pass flushes its non-empty bins)

struct PostDev {
    int num_scales, C, H, W, h4, w4;
    int gh[OM_MAX_SCALES], gw[OM_MAX_SCALES], na[OM_MAX_SCALES], aidx[OM_MAX_SCALES][4];
    int pred_off[OM_MAX_SCALES + 1];
    int n_pred;
    int total_anchors;
    float aw_px[OM_MAX_ANCHORS], ah_px[OM_MAX_ANCHORS];
    int anchor_scale[OM_MAX_ANCHORS], anchor_slot[OM_MAX_ANCHORS];   // anchor -> (scale, index inside scale)
    float conf_thresh, nms_thresh, orien_thresh;
    int nms_pre, nms_post, NP;
    const float* bbox[OM_MAX_SCALES];
    long long bstride[OM_MAX_SCALES];
    const float* orien[OM_MAX_SCALES];
    long long ostride[OM_MAX_SCALES];
};

struct SelState {
    unsigned hist[2][kHistBins];
    unsigned prefix, k_rem, total, take_all, n_out, n_list, n_edge, done;
};

// Radix digits of a score.  Scores are fp32 in (conf_thresh, 1], so their bit patterns span only
// [bits(conf_thresh), 0x3F800000]: the digits are cut from rel = key - kmin (nb significant bits, nb <= 30), not from
// the raw pattern (whose top 11 bits are the sign, the exponent and two mantissa bits: ~30 distinct values).
// bin0 = rel >> s0 (8 bits), bin1 = (rel >> s1) & m1 (<= 11 bits), bin2 = rel & m2 (<= 11 bits).
struct KeyBins {
    unsigned kmin;
    int s0, s1;
    unsigned m1, m2;
};

constexpr int kTailThreads = 512;

__device__ __forceinline__ float sigmoidf_rn(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void locate_pred(const PostDev& d, int n, int& s, int& a, int& cell, int& plane) {
    s = 0;
#pragma unroll
    for (int i = 1; i < OM_MAX_SCALES; ++i)
        if (i < d.num_scales && n >= d.pred_off[i]) s = i;
    int r = n - d.pred_off[s];
    plane = d.gh[s] * d.gw[s];
    a = r / plane;
    cell = r - a * plane;
}

// One warp: walk a 2048-bin histogram from the top bin down until k_rem elements are covered.
// Returns (to every lane) the bin that holds the k_rem-th largest element, the number of elements still to take
// from that bin, and the total count.
__device__ __noinline__ void scan_top(const unsigned* h, unsigned k_rem, int& bin_out, unsigned& rem_out, unsigned& total_out) {
    const int lane = threadIdx.x & 31;
    constexpr int per = kHistBins / 32;
    unsigned sum = 0;
    const int top = kHistBins - 1 - lane * per;          // this lane covers bins top, top-1, ..., top-per+1
    for (int i = 0; i < per; ++i) sum += h[top - i];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    total_out = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned excl = incl - sum;
    int bin = 0;
    unsigned rem = 0;
    const bool mine = excl < k_rem && k_rem <= incl;     // at most one lane
    if (mine) {
        unsigned above = excl;
        bin = top;
        for (int i = 0; i < per; ++i) {
            bin = top - i;
            const unsigned c = h[bin];
            if (above + c >= k_rem) break;
            above += c;
        }
        rem = k_rem - above;
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int src = who ? __ffs(who) - 1 : 0;
    bin_out = __shfl_sync(0xffffffffu, bin, src);
    rem_out = __shfl_sync(0xffffffffu, rem, src);
}

// Pass 1 (the only pass over the head tensors): one thread per prediction, 80 plane-strided (coalesced across
// the warp) class reads each.  Every (prediction, class) whose score clears conf_thresh is appended to the
// image's candidate list (score bits + flat index, unordered; slots handed out per warp and round of 8 classes) and counted in the level-0 radix histogram; the remaining radix levels and the collection
// run over that list, not over the heads, so sigmoid/exp are evaluated once per pair.  Logits that cannot clear the
// threshold even with sigma(obj) = 1 skip the sigmoid entirely (margin 0.01 in logit space >> any rounding of the
// fp32 sigmoid).  The last block of an image to finish walks the level-0 histogram (no separate launch).
__global__ void __launch_bounds__(kSelThreads) conf_compact_kernel(PostDev d, KeyBins kb, SelState* st, unsigned* keys, unsigned* flats,
                                                                   long long list_cap, float reject_logit) {
    __shared__ unsigned sh[kHistBins];
    __shared__ unsigned s_last;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    SelState& S = st[b];
    for (int i = threadIdx.x; i < kL0Bins; i += kSelThreads) sh[i] = 0;
    __syncthreads();
    const int n = blockIdx.x * kSelThreads + threadIdx.x;
    const float* pc = nullptr;
    int plane = 0;
    float obj = 0.f, rej = reject_logit;
    bool live = false;
    if (n < d.n_pred) {
        int s, a, cell;
        locate_pred(d, n, s, a, cell, plane);
        const float* p = d.bbox[s] + (long long)b * d.bstride[s] + (long long)(a * (5 + d.C)) * plane + cell;
        const float ol = __ldg(p + 4 * plane);
        if (ol >= reject_logit) {
            obj = sigmoidf_rn(ol);
            live = true;
            // per-prediction bound: sigma(x) * obj > conf_thresh needs x > logit(conf_thresh / obj); 0.05 in logit space
            // is a relative margin of >= 5e-4 on sigma(x) for conf_thresh / obj < 0.99 -- far above any fp32 rounding
            // of the sigmoid, the quotient or the product -- so only pairs that certainly fail skip the sigmoid
            const float r = d.conf_thresh / obj;
            const float t = r < 0.99f ? logf(r / (1.0f - r)) - 0.05f : 4.0f;      // sigma(4) = 0.982 < 0.99 (1 - 5e-4)
            rej = fmaxf(rej, t);                                                   // NaN / -inf (conf_thresh <= 0): keeps rej
        }
        pc = p + 5 * plane;
    }
    unsigned* kout = keys + (long long)b * list_cap;
    unsigned* fout = flats + (long long)b * list_cap;
    for (int c0 = 0; c0 < d.C; c0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (live && c0 + i < d.C) ? __ldg(pc + (long long)(c0 + i) * plane) : -1e30f;
        unsigned okmask = 0, kbits[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            kbits[i] = 0;
            if (v[i] >= rej) {
                const float conf = sigmoidf_rn(v[i]) * obj;
                if (live && conf > d.conf_thresh) okmask |= 1u << i;
                kbits[i] = __float_as_uint(conf);
            }
        }
        // list slots of the round: one warp scan of the per-lane counts and one atomic on the image's list length per
        // warp; the warp then owns a contiguous range of the list and there is no block-level synchronisation at all
        const unsigned mine = __popc(okmask);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned wtotal = __shfl_sync(0xffffffffu, incl, 31);
        if (wtotal) {                                                   // warp-uniform
            unsigned slot = 0;
            if (lane == 31) slot = atomicAdd(&S.n_list, wtotal);
            slot = __shfl_sync(0xffffffffu, slot, 31) + incl - mine;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if ((okmask >> i) & 1u) {
                    kout[slot] = kbits[i];
                    fout[slot] = (unsigned)(n * d.C + c0 + i);
                    ++slot;
                    atomicAdd(&sh[(kbits[i] - kb.kmin) >> kb.s0], 1u);
                }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kL0Bins; i += kSelThreads)
        if (sh[i]) atomicAdd(&S.hist[0][i], sh[i]);
    // last block of this image: level-0 scan
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&S.done, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < kHistBins; i += kSelThreads) sh[i] = i < kL0Bins ? __ldcg(&S.hist[0][i]) : 0u;
    __syncthreads();
    if (threadIdx.x < 32) {
        int bin;
        unsigned rem, total;
        scan_top(sh, (unsigned)d.nms_pre, bin, rem, total);
        if (lane == 0) {
            S.total = total;
            if (total <= (unsigned)d.nms_pre) { S.take_all = 1; }
            else { S.prefix = (unsigned)bin; S.k_rem = rem; }
        }
    }
}

// Pass 2, over the candidate list (a fixed number of blocks per image strides over it): candidates above the
// level-0 boundary bin are certain members of the top-k and go straight to the output; candidates inside it are
// counted in the level-1 histogram and copied to the image's edge list, the only thing the tail kernel still reads.
__global__ void __launch_bounds__(kSelThreads) select_edge_kernel(PostDev d, KeyBins kb, SelState* st, const unsigned* keys, const unsigned* flats,
                                                                  long long list_cap, uint2* edge, uint2* raw) {
    __shared__ unsigned sh[kHistBins];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    SelState& S = st[b];
    const unsigned n_list = S.n_list;
    const unsigned stride = gridDim.x * kSelThreads;
    if (blockIdx.x * kSelThreads >= n_list) return;
    const unsigned take_all = S.take_all, prefix = S.prefix;
    for (int i = threadIdx.x; i < kHistBins; i += kSelThreads) sh[i] = 0;
    __syncthreads();
    const unsigned* kin = keys + (long long)b * list_cap;
    const unsigned* fin = flats + (long long)b * list_cap;
    uint2* eout = edge + (long long)b * list_cap;
    constexpr int U = 4;
    for (unsigned base0 = blockIdx.x * kSelThreads; base0 < n_list; base0 += U * stride) {        // block-uniform trip count
        const unsigned i0 = base0 + threadIdx.x;
        unsigned key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned i = i0 + u * stride;
            key[u] = i < n_list ? __ldg(kin + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned i = i0 + u * stride;
            const bool valid = i < n_list;
            const unsigned rel = key[u] - kb.kmin;
            const unsigned b0 = rel >> kb.s0;
            const bool sure = valid && (take_all || b0 > prefix);
            const bool onedge = valid && !take_all && b0 == prefix;
            if (sure) {
                const unsigned slot = atomicAdd(&S.n_out, 1u);
                if (slot < (unsigned)d.nms_pre) raw[(long long)b * d.nms_pre + slot] = make_uint2(key[u], fin[i]);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, onedge);
            if (bal == 0) continue;                                     // warp-uniform
            const int leader = __ffs(bal) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&S.n_edge, (unsigned)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (onedge) {
                eout[base + __popc(bal & ((1u << lane) - 1u))] = make_uint2(key[u], fin[i]);
                atomicAdd(&sh[(rel >> kb.s1) & kb.m1], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistBins; i += kSelThreads)
        if (sh[i]) atomicAdd(&S.hist[1][i], sh[i]);
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* key, int NP) {
    for (int size = 2; size <= NP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < NP / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) { key[lo] = b; key[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// One CTA per image, the tail of the selection: level-1 boundary from the histogram of the edge pass, level-2
// histogram and boundary over the edge list (normally a few dozen entries: one of 256 level-0 bins), collection
// of the edge candidates above the threshold (ties at the threshold: as many as are still needed), then order the
// <= nms_pre (key, flat) pairs the way the reference orders its candidates, decode their boxes and emit the
// candidate arrays.
__global__ void __launch_bounds__(kTailThreads) select_tail_kernel(PostDev d, KeyBins kb, SelState* st, const uint2* edge, long long list_cap,
                                                                   uint2* raw, int* cand_count, float* cand_det, int* cand_cls, int* cand_pred) {
    extern __shared__ unsigned long long skey[];
    __shared__ unsigned sh[kHistBins];
    __shared__ unsigned s_prefix, s_krem, s_T, s_need, s_eq, s_out;
    const int b = blockIdx.x;
    SelState& S = st[b];
    const bool take_all = S.take_all != 0;
    if (!take_all) {
        const unsigned n_edge = S.n_edge;
        const uint2* ein = edge + (long long)b * list_cap;
        if (threadIdx.x < 32) {
            int bin;
            unsigned rem, total;
            scan_top(S.hist[1], S.k_rem, bin, rem, total);
            if (threadIdx.x == 0) {
                s_prefix = (S.prefix << (kb.s0 - kb.s1)) | (unsigned)bin;
                s_krem = rem;
                s_eq = 0;
                s_out = S.n_out;
            }
        }
        for (int i = threadIdx.x; i < kHistBins; i += kTailThreads) sh[i] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        constexpr int U = 4;                               // independent loads in flight (the list is long only when scores tie en masse)
        for (unsigned i0 = threadIdx.x; i0 < n_edge; i0 += U * kTailThreads) {
            unsigned key[U];
#pragma unroll
            for (int u = 0; u < U; ++u) key[u] = i0 + u * kTailThreads < n_edge ? ein[i0 + u * kTailThreads].x : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned rel = key[u] - kb.kmin;
                if (i0 + u * kTailThreads < n_edge && (rel >> kb.s1) == prefix) atomicAdd(&sh[rel & kb.m2], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            int bin;
            unsigned rem, total;
            scan_top(sh, s_krem, bin, rem, total);
            if (threadIdx.x == 0) {
                s_T = (prefix << kb.s1) | (unsigned)bin;
                s_need = rem;
            }
        }
        __syncthreads();
        const unsigned T = s_T, need_eq = s_need;
        for (unsigned i0 = threadIdx.x; i0 < n_edge; i0 += U * kTailThreads) {
            uint2 e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) e[u] = i0 + u * kTailThreads < n_edge ? ein[i0 + u * kTailThreads] : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * kTailThreads >= n_edge) continue;
                const unsigned rel = e[u].x - kb.kmin;
                bool take = rel > T;
                if (!take && rel == T) take = atomicAdd(&s_eq, 1u) < need_eq;
                if (take) {
                    const unsigned slot = atomicAdd(&s_out, 1u);
                    if (slot < (unsigned)d.nms_pre) raw[(long long)b * d.nms_pre + slot] = e[u];
                }
            }
        }
        __syncthreads();                                   // raw[] of this image is complete and visible to the block
    } else {
        if (threadIdx.x == 0) s_out = S.n_out;
        __syncthreads();
    }
    const int n = min((int)s_out, d.nms_pre);
    for (int i = threadIdx.x; i < d.NP; i += blockDim.x) {
        unsigned long long v = ~0ull;
        if (i < n) {
            const uint2 r = raw[(long long)b * d.nms_pre + i];
            v = take_all ? (((unsigned long long)r.y << 32) | r.x)
                         : (((unsigned long long)(~r.x) << 32) | r.y);
        }
        skey[i] = v;
    }
    bitonic_sort_u64(skey, d.NP);
    if (threadIdx.x == 0) cand_count[b] = n;
    for (int p = threadIdx.x; p < d.nms_pre; p += blockDim.x) {
        float* out = cand_det + ((long long)b * d.nms_pre + p) * 5;
        if (p >= n) {
            out[0] = out[1] = out[2] = out[3] = out[4] = 0.f;
            cand_cls[(long long)b * d.nms_pre + p] = 0;
            cand_pred[(long long)b * d.nms_pre + p] = 0;
            continue;
        }
        const unsigned long long v = skey[p];
        const unsigned key = take_all ? (unsigned)(v & 0xffffffffu) : ~(unsigned)(v >> 32);
        const unsigned flat = take_all ? (unsigned)(v >> 32) : (unsigned)(v & 0xffffffffu);
        const int pred = flat / d.C, cls = flat - pred * d.C;
        int s, a, cell, plane;
        locate_pred(d, pred, s, a, cell, plane);
        const int gy = cell / d.gw[s], gx = cell - gy * d.gw[s];
        const float* q = d.bbox[s] + (long long)b * d.bstride[s] + (long long)(a * (5 + d.C)) * plane + cell;
        const int anchor = d.aidx[s][a];
        const float aw = d.aw_px[anchor] / (float)d.W;            // normalized_anchors (:18-20)
        const float ah = d.ah_px[anchor] / (float)d.H;
        out[0] = (sigmoidf_rn(q[0]) + (float)gx) / (float)d.gw[s];
        out[1] = (sigmoidf_rn(q[plane]) + (float)gy) / (float)d.gh[s];
        out[2] = expf(q[2 * plane]) * aw;
        out[3] = expf(q[3 * plane]) * ah;
        out[4] = __uint_as_float(key);
        cand_cls[(long long)b * d.nms_pre + p] = cls;
        cand_pred[(long long)b * d.nms_pre + p] = pred;
    }
}

struct NmsArgs {
    const float* dets;        // [batch, cap, 5]
    const int* cls;           // [batch, cap] or null (no class offset)
    const int* pred;          // [batch, cap] or null
    const int* counts;        // [batch] or null -> n
    int n, cap, NP, nms_post;
    float thr;
    int mode;                 // OM_NMS_CPU: nms_cpu.cpp (>= suppresses, areas from corners, ascending-index result);
                              // OM_NMS_CUDA: nms_kernel.cu (> suppresses, areas w*h, score-descending result)
    // batched outputs
    int* det_count; float* det; long long* det_cls; int* det_anchor; int* det_keep;
    float* records;                 // optional [batch, nms_post*6 + 1]: (cx, cy, w, h, score, cls) per slot, then the count -- the row a rank all-gathers
    // stand-alone outputs
    long long* keep; int* keep_count;
};

__device__ __forceinline__ unsigned orderable(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// block-wide stream compaction of byte flags (n <= 1024): slot[i] = number of set flags before i.
__device__ int compact_slots(const unsigned char* flag, int n, int* slot, int* chunk_off) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nchunks = (n + 31) >> 5;
    __syncthreads();
    for (int c = warp; c < nchunks; c += nwarps) {
        const int i = c * 32 + lane;
        const unsigned bal = __ballot_sync(0xffffffffu, i < n && flag[i]);
        if (i < n) slot[i] = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) chunk_off[c] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {
        int v = lane < nchunks ? chunk_off[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        chunk_off[lane] = incl - v;
        if (lane == 31) chunk_off[32] = incl;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] += chunk_off[i >> 5];
    __syncthreads();
    return chunk_off[32];
}

// One CTA per box list.  Warp-cooperative: all warps fill the n x n/32 suppression bit matrix,
// warp 0 then resolves the greedy chain with the removed-set held one word per lane.
template <bool BATCHED>
__global__ void __launch_bounds__(512) nms_kernel(NmsArgs a, PostDev d) {
    extern __shared__ unsigned long long smem64[];
    const int b = blockIdx.x;
    const int n = a.counts ? min(a.counts[b], a.cap) : a.n;
    const int cap = a.cap;
    unsigned long long* skey = smem64;                               // [NP]
    float* gx1 = reinterpret_cast<float*>(skey + a.NP);              // [cap] x5
    float* gy1 = gx1 + cap; float* gx2 = gy1 + cap; float* gy2 = gx2 + cap; float* gar = gy2 + cap;
    int* slot = reinterpret_cast<int*>(gar + cap);                   // [cap]
    int* chunk_off = slot + cap;                                     // [33]
    unsigned* mask = reinterpret_cast<unsigned*>(chunk_off + 36);    // [cap * nw]
    const int nw = (cap + 31) >> 5;
    unsigned char* kept_rank = reinterpret_cast<unsigned char*>(mask + (size_t)cap * nw);   // [cap]
    unsigned char* kept_pos = kept_rank + cap;                                                // [cap]
    const float* dets = a.dets + (long long)b * cap * 5;
    const int* cls = a.cls ? a.cls + (long long)b * cap : nullptr;

    for (int i = threadIdx.x; i < a.NP; i += blockDim.x)
        skey[i] = i < n ? (((unsigned long long)(~orderable(dets[i * 5 + 4])) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    {   // om_decode_select hands its candidates over in exactly this order unless it took every candidate: sort only if needed
        int unsorted = 0;
        for (int i = threadIdx.x; i + 1 < a.NP; i += blockDim.x) unsorted |= skey[i] > skey[i + 1];
        if (__syncthreads_or(unsorted)) bitonic_sort_u64(skey, a.NP); // rank -> position, score desc / index asc
    }

    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int p = (int)(skey[r] & 0xffffffffu);
        const float off = cls ? (float)cls[p] * 2.0f : 0.0f;         // function.py:94-97 (max_coordinate 1.5 + 0.5)
        const float x = dets[p * 5 + 0] + off, y = dets[p * 5 + 1] + off;
        const float hw = dets[p * 5 + 2] / 2.0f, hh = dets[p * 5 + 3] / 2.0f;
        const float x1 = x - hw, y1 = y - hh, x2 = x + hw, y2 = y + hh;
        gx1[r] = x1; gy1[r] = y1; gx2[r] = x2; gy2[r] = y2;
        gar[r] = a.mode == OM_NMS_CUDA ? dets[p * 5 + 2] * dets[p * 5 + 3]          // nms_kernel.cu:20-21: Sa = a[2] * a[3]
                                       : (x2 - x1) * (y2 - y1);                      // nms_cpu.cpp:22
        kept_rank[r] = 0; kept_pos[r] = 0;
    }
    __syncthreads();

    // suppression bit matrix, upper triangle: one warp per (32-row block, 32-column word); lane = row, so the five
    // shared-memory reads of box j are warp-wide broadcasts and the word stores hit 32 distinct banks (nw is odd
    // or the rows differ in r * nw mod 32 -- at worst a 2-way conflict), instead of nw-way conflicts on every read
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const int nrb = (n + 31) >> 5;
        for (int item = warp; item < nrb * nw; item += nwarps) {
            const int rb = item / nw, wd = item - rb * nw;
            const int r = rb * 32 + lane;
            if (r >= n) continue;
            unsigned bits = 0;
            if (wd >= rb) {
                const float ix1 = gx1[r], iy1 = gy1[r], ix2 = gx2[r], iy2 = gy2[r], ia = gar[r];
                const int jn = max(0, min(32, n - wd * 32));
                for (int bb = 0; bb < jn; ++bb) {
                    const int j = wd * 32 + bb;
                    const float xx1 = fmaxf(ix1, gx1[j]), yy1 = fmaxf(iy1, gy1[j]);
                    const float xx2 = fminf(ix2, gx2[j]), yy2 = fminf(iy2, gy2[j]);
                    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
                    const float inter = w * h;
                    // disjoint boxes (almost all pairs once the class offset is applied): inter == 0 gives ovr = 0 or NaN,
                    // neither of which reaches a positive threshold -- skip the IEEE division
                    if (inter == 0.0f && a.thr > 0.0f) continue;
                    const float ovr = inter / (ia + gar[j] - inter);
                    if (j > r && (a.mode == OM_NMS_CUDA ? ovr > a.thr : ovr >= a.thr)) bits |= 1u << bb;   // nms_kernel.cu:58 / nms_cpu.cpp:59
                }
            }
            mask[r * nw + wd] = bits;
        }
    }
    __syncthreads();

    // greedy chain, 32 ranks at a time: every lane fetches the chunk's 32 diagonal words (independent shuffles) and
    // resolves the chunk's own dependencies in registers (32 unrolled ALU steps), then the full rows of the kept boxes
    // are OR-ed into the removed set (lane = word) as independent predicated loads -- instead of one dependent
    // shuffle + shared-memory round trip per rank
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned removed = 0;
        const int nchunks = (n + 31) >> 5;
        for (int c = 0; c < nchunks; ++c) {
            const int r = c * 32 + lane;
            const unsigned diag = r < n ? mask[r * nw + c] : 0u;
            unsigned dg[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) dg[i] = __shfl_sync(0xffffffffu, diag, i);
            const int left = n - c * 32;
            const unsigned validbits = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
            unsigned w = __shfl_sync(0xffffffffu, removed, c) | ~validbits;     // ranks past n count as removed
            unsigned keptbits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (!((w >> i) & 1u)) { keptbits |= 1u << i; w |= dg[i]; }
            if (r < n) kept_rank[r] = (unsigned char)((keptbits >> lane) & 1u);
            unsigned acc = 0;
            const unsigned* mrow = mask + (size_t)c * 32 * nw + (lane < nw ? lane : 0);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if ((keptbits >> i) & 1u) acc |= mrow[i * nw];
            if (lane < nw) removed |= acc;
        }
    }
    __syncthreads();

    int total = compact_slots(kept_rank, n, slot, chunk_off);
    // result order: score-descending ranks after a post-NMS top-k (postprocess.py:150-154) and always under nms_kernel.cu semantics
    // (:136-139, order_t.index(keep)); ascending original index otherwise (nms_cpu.cpp:62)
    const bool topk = (BATCHED && total > a.nms_post) || a.mode == OM_NMS_CUDA;
    if (!topk) {
        for (int r = threadIdx.x; r < n; r += blockDim.x)
            if (kept_rank[r]) kept_pos[(int)(skey[r] & 0xffffffffu)] = 1;
        total = compact_slots(kept_pos, n, slot, chunk_off);
    }
    const int n_out = BATCHED ? min(total, a.nms_post) : total;

    if (BATCHED) {
        float* det = a.det + (long long)b * a.nms_post * 5;
        long long* dcls = a.det_cls + (long long)b * a.nms_post;
        int* danc = a.det_anchor + (long long)b * a.nms_post;
        int* dkeep = a.det_keep + (long long)b * a.nms_post;
        const int* pred = a.pred + (long long)b * cap;
        float* rec = a.records ? a.records + (long long)b * (a.nms_post * 6 + 1) : nullptr;
        if (threadIdx.x == 0) { a.det_count[b] = n_out; if (rec) rec[a.nms_post * 6] = (float)n_out; }
        for (int i = threadIdx.x; i < a.nms_post; i += blockDim.x)
            if (i >= n_out) {
                for (int c = 0; c < 5; ++c) det[i * 5 + c] = 0.f;
                dcls[i] = 0; danc[i] = 0; dkeep[i] = 0;
                if (rec) for (int c = 0; c < 6; ++c) rec[i * 6 + c] = 0.f;
            }
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            // e indexes ranks (top-k case, score-descending) or positions (ascending-index case)
            const bool on = topk ? kept_rank[e] : kept_pos[e];
            if (!on || slot[e] >= n_out) continue;
            const int p = topk ? (int)(skey[e] & 0xffffffffu) : e;
            const int o = slot[e];
            for (int c = 0; c < 5; ++c) det[o * 5 + c] = dets[p * 5 + c];
            dcls[o] = cls[p];
            if (rec) { for (int c = 0; c < 5; ++c) rec[o * 6 + c] = dets[p * 5 + c]; rec[o * 6 + 5] = (float)cls[p]; }
            int s, an, cell, plane;
            locate_pred(d, pred[p], s, an, cell, plane);
            danc[o] = d.aidx[s][an];
            dkeep[o] = p;
        }
    } else {
        if (threadIdx.x == 0) *a.keep_count = n_out;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            if (topk) { if (kept_rank[e]) a.keep[slot[e]] = (long long)(skey[e] & 0xffffffffu); }
            else if (kept_pos[e]) a.keep[slot[e]] = e;
        }
    }
}

struct MaskInst { float xc, yc, tw, th; int k; };

// Each thread owns 16 consecutive pixels of one image row: per anchor it rebuilds the x4 bilinear
// up-sampling of that anchor's two orientation channels in registers, then streams one 16-byte
// store per instance of that anchor.  HBM-write bound: K*H*W bytes out, 18*(H/4)*(W/4)*4 bytes in.
__global__ void __launch_bounds__(256) mask_kernel(PostDev d, const int* det_count, const float* det,
                                                   const int* det_anchor, unsigned char* mask) {
    extern __shared__ unsigned char msm[];
    MaskInst* inst = reinterpret_cast<MaskInst*>(msm);               // [nms_post], grouped by anchor
    __shared__ int a_cnt[OM_MAX_ANCHORS], a_start[OM_MAX_ANCHORS + 1], a_fill[OM_MAX_ANCHORS];
    const int b = blockIdx.y;
    const int K = min(det_count[b], d.nms_post);
    if (K == 0) return;
    if (threadIdx.x < OM_MAX_ANCHORS) { a_cnt[threadIdx.x] = 0; a_fill[threadIdx.x] = 0; }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(&a_cnt[det_anchor[(long long)b * d.nms_post + k]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int a = 0; a < OM_MAX_ANCHORS; ++a) { a_start[a] = acc; acc += a_cnt[a]; }
        a_start[OM_MAX_ANCHORS] = acc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int a = det_anchor[(long long)b * d.nms_post + k];
        const float* r = det + ((long long)b * d.nms_post + k) * 5;
        const int s = d.anchor_scale[a];
        const float gw = (float)d.gw[s], gh = (float)d.gh[s];       // grid_sizes (:22-27)
        MaskInst m;
        m.xc = gw * r[0];                                            // :157-158
        m.yc = gh * r[1];
        m.tw = (d.orien_thresh * r[2]) * gw;                         // :162,164  (thresh * size) * grid
        m.th = (d.orien_thresh * r[3]) * gh;
        m.k = k;
        inst[a_start[a] + atomicAdd(&a_fill[a], 1)] = m;
    }
    __syncthreads();

    const int units_per_row = d.W >> 4;
    const int unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= d.H * units_per_row) return;
    const int y = unit / units_per_row;
    const int xb = (unit - y * units_per_row) << 4;
    // source rows (align_corners=False, scale 4): src = max((y+0.5)/4-0.5, 0)
    const float sy = fmaxf((y + 0.5f) * 0.25f - 0.5f, 0.0f);
    const int y0 = min((int)floorf(sy), d.h4 - 1);
    const int y1 = min(y0 + 1, d.h4 - 1);
    const float ly1 = sy - (float)y0, ly0 = 1.0f - ly1;
    int col[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) col[j] = min(max((xb >> 2) - 1 + j, 0), d.w4 - 1);
    const long long img_plane = (long long)d.H * d.W;
    unsigned char* mrow = mask + (long long)b * d.nms_post * img_plane + (long long)y * d.W + xb;

    // pixel-centre grids of :41-43 ((x / W) * nW): they depend on the scale only, and the IEEE division is the most
    // expensive part of a pixel, so they are rebuilt only when the scale changes between consecutive anchors
    float bxs[16], bys = 0.f;
    int cur_scale = -1;
    for (int a = 0; a < d.total_anchors; ++a) {
        const int i0 = a_start[a], i1 = a_start[a + 1];
        if (i0 == i1) continue;                                      // block-uniform
        const int s = d.anchor_scale[a];
        if (s != cur_scale) {                                        // block-uniform
            cur_scale = s;
            bys = ((float)y / (float)d.H) * (float)d.gh[s];
#pragma unroll
            for (int i = 0; i < 16; ++i) bxs[i] = ((float)(xb + i) / (float)d.W) * (float)d.gw[s];
        }
        const float* op = d.orien[s] + (long long)b * d.ostride[s] + (long long)(2 * d.anchor_slot[a]) * d.h4 * d.w4;
        // grid_anchors (:21-25): (px / W) * nW
        const float gax = (d.aw_px[a] / (float)d.W) * (float)d.gw[s];
        const float gay = (d.ah_px[a] / (float)d.H) * (float)d.gh[s];
        const float base_y = bys;
        float px[16], py[16];
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
            const float* r0 = op + (long long)comp * d.h4 * d.w4 + (long long)y0 * d.w4;
            const float* r1 = op + (long long)comp * d.h4 * d.w4 + (long long)y1 * d.w4;
            float v0[6], v1[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) { v0[j] = __ldg(r0 + col[j]); v1[j] = __ldg(r1 + col[j]); }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int j = (i + 2) >> 2;                          // i0 column relative to col[0]
                float lx1 = ((i & 3) == 0) ? 0.625f : ((i & 3) == 1) ? 0.875f : ((i & 3) == 2) ? 0.125f : 0.375f;
                if (xb + i < 2) lx1 = 0.0f;                          // clamped source (src < 0)
                const float lx0 = 1.0f - lx1;
                const float t0 = __fmaf_rn(lx0, v0[j], lx1 * v0[j + 1]);
                const float t1 = __fmaf_rn(lx0, v1[j], lx1 * v1[j + 1]);
                const float up = __fmaf_rn(ly0, t0, ly1 * t1);
                if (comp == 0) {
                    const float base_x = bxs[i];
                    px[i] = (up * gax) / 2.0f + base_x;              // :141-144
                } else {
                    py[i] = (up * gay) / 2.0f + base_y;
                }
            }
        }
        // bounds of the run: fp32 subtraction is monotonic, so fl(min - c) >= t (or fl(max - c) <= -t) for the run
        // extremes proves |fl(p - c)| >= t for every pixel of the run -- an all-zero store without the 16 tests
        float pxmin = px[0], pxmax = px[0], pymin = py[0], pymax = py[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            pxmin = fminf(pxmin, px[i]); pxmax = fmaxf(pxmax, px[i]);
            pymin = fminf(pymin, py[i]); pymax = fmaxf(pymax, py[i]);
        }
        for (int e = i0; e < i1; ++e) {
            const MaskInst m = inst[e];
            if (pxmin - m.xc >= m.tw || pxmax - m.xc <= -m.tw || pymin - m.yc >= m.th || pymax - m.yc <= -m.th) {
                __stcs(reinterpret_cast<uint4*>(mrow + (long long)m.k * img_plane), make_uint4(0u, 0u, 0u, 0u));
                continue;
            }
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned v = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = q * 4 + t;
                    const bool in = (fabsf(px[i] - m.xc) < m.tw) && (fabsf(py[i] - m.yc) < m.th);
                    v |= (in ? 1u : 0u) << (8 * t);
                }
                w[q] = v;
            }
            __stcs(reinterpret_cast<uint4*>(mrow + (long long)m.k * img_plane), make_uint4(w[0], w[1], w[2], w[3]));
        }
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int32_t make_dev(const om_post_config* c, PostDev& d) {
    if (!c) return om::fail(OM_ERR_INVALID, "null post config");
    if (c->num_scales < 1 || c->num_scales > OM_MAX_SCALES) return om::fail(OM_ERR_INVALID, "num_scales %d out of range", c->num_scales);
    if (c->total_anchors < 1 || c->total_anchors > OM_MAX_ANCHORS) return om::fail(OM_ERR_INVALID, "total_anchors %d out of range", c->total_anchors);
    if (c->nms_pre < 1 || c->nms_pre > 1024 || c->nms_post < 1 || c->nms_post > c->nms_pre)
        return om::fail(OM_ERR_INVALID, "need 1 <= nms_post <= nms_pre <= 1024 (got %d, %d)", c->nms_post, c->nms_pre);
    if (c->image_h % 32 || c->image_w % 32 || c->image_h <= 0 || c->image_w <= 0)
        return om::fail(OM_ERR_INVALID, "image size %dx%d must be a positive multiple of 32", c->image_h, c->image_w);
    if (c->num_classes < 1) return om::fail(OM_ERR_INVALID, "num_classes must be positive");
    if (!(c->conf_thresh > 0.f)) return om::fail(OM_ERR_INVALID, "conf_thresh must be > 0");
    memset(&d, 0, sizeof(d));
    d.num_scales = c->num_scales; d.C = c->num_classes; d.H = c->image_h; d.W = c->image_w;
    d.h4 = c->image_h / 4; d.w4 = c->image_w / 4;
    d.total_anchors = c->total_anchors;
    for (int a = 0; a < OM_MAX_ANCHORS; ++a) { d.anchor_scale[a] = 0; d.anchor_slot[a] = 0; }
    int off = 0;
    for (int s = 0; s < c->num_scales; ++s) {
        if (c->anchors_per_scale[s] < 1 || c->anchors_per_scale[s] > 4) return om::fail(OM_ERR_INVALID, "anchors_per_scale out of range");
        d.gh[s] = c->grid_h[s]; d.gw[s] = c->grid_w[s]; d.na[s] = c->anchors_per_scale[s];
        d.pred_off[s] = off;
        off += d.na[s] * d.gh[s] * d.gw[s];
        for (int j = 0; j < d.na[s]; ++j) {
            const int a = c->anchor_index[s][j];
            if (a < 0 || a >= c->total_anchors) return om::fail(OM_ERR_INVALID, "anchor_mask entry %d out of range", a);
            d.aidx[s][j] = a; d.anchor_scale[a] = s; d.anchor_slot[a] = j;
        }
    }
    for (int s = c->num_scales; s <= OM_MAX_SCALES; ++s) d.pred_off[s] = off;
    d.n_pred = off;
    if ((long long)off * c->num_classes >= (1ll << 31)) return om::fail(OM_ERR_INVALID, "too many (prediction, class) pairs");
    for (int a = 0; a < c->total_anchors; ++a) { d.aw_px[a] = c->anchor_w[a]; d.ah_px[a] = c->anchor_h[a]; }
    d.conf_thresh = c->conf_thresh; d.nms_thresh = c->nms_thresh; d.orien_thresh = c->orien_thresh;
    d.nms_pre = c->nms_pre; d.nms_post = c->nms_post; d.NP = next_pow2(c->nms_pre);
    return OM_OK;
}

size_t nms_smem_bytes(int NP, int cap) {
    const int nw = (cap + 31) >> 5;
    return (size_t)NP * 8 + (size_t)cap * 5 * 4 + (size_t)cap * 4 + 36 * 4 + (size_t)cap * nw * 4 + 2 * (size_t)cap + 16;
}

template <typename K>
int32_t allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) OM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return OM_OK;
}

}  // namespace

// workspace layout: [batch] SelState | [batch][nms_pre] raw (key, flat) | [batch][cap] keys | [batch][cap] flats |
// [batch][cap] edge (key, flat); cap = n_pred * num_classes (every pair may clear conf_thresh, e.g. untrained weights)
static size_t ws_align(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int32_t om_post_workspace_bytes(const om_post_config* cfg, int32_t batch, size_t* bytes) {
    PostDev d;
    int32_t rc = make_dev(cfg, d);
    if (rc) return rc;
    if (batch < 1 || !bytes) return om::fail(OM_ERR_INVALID, "bad batch / null output");
    const size_t cap = (size_t)d.n_pred * d.C;
    *bytes = ws_align((size_t)batch * sizeof(SelState)) + ws_align((size_t)batch * cfg->nms_pre * sizeof(uint2)) +
             2 * ws_align((size_t)batch * cap * sizeof(unsigned)) + ws_align((size_t)batch * cap * sizeof(uint2));
    return OM_OK;
}

// Radix digits of the scores in (conf_thresh, 1] (see KeyBins).
static KeyBins make_bins(float conf_thresh) {
    KeyBins kb{};
    unsigned kmin = 0;
    if (conf_thresh > 0.f) memcpy(&kmin, &conf_thresh, 4);
    const unsigned kmax = 0x3F800000u;                               // 1.0f: sigma(cls) * sigma(obj) <= 1
    const unsigned span1 = kmin < kmax ? kmax - kmin : 0u;            // largest rel
    int nb = 0;
    while (nb < 32 && (span1 >> nb) != 0u) ++nb;
    kb.kmin = kmin;
    kb.s0 = nb > 8 ? nb - 8 : 0;
    kb.s1 = nb > 19 ? nb - 19 : 0;
    kb.m1 = (1u << (kb.s0 - kb.s1)) - 1u;
    kb.m2 = (1u << kb.s1) - 1u;
    return kb;
}

extern "C" int32_t om_decode_select(const om_post_config* cfg, const float* const* bbox, const int64_t* bbox_batch_stride,
                                    int32_t batch, void* workspace, int32_t* cand_count, float* cand_det,
                                    int32_t* cand_cls, int32_t* cand_pred, void* stream) {
    PostDev d;
    int32_t rc = make_dev(cfg, d);
    if (rc) return rc;
    if (batch < 1 || !bbox || !bbox_batch_stride || !workspace || !cand_count || !cand_det || !cand_cls || !cand_pred)
        return om::fail(OM_ERR_INVALID, "om_decode_select: null argument or batch < 1");
    for (int s = 0; s < d.num_scales; ++s) {
        if (!bbox[s]) return om::fail(OM_ERR_INVALID, "om_decode_select: bbox[%d] is null", s);
        d.bbox[s] = bbox[s]; d.bstride[s] = bbox_batch_stride[s];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long cap = (long long)d.n_pred * d.C;
    char* ws = reinterpret_cast<char*>(workspace);
    SelState* state = reinterpret_cast<SelState*>(ws);
    ws += ws_align((size_t)batch * sizeof(SelState));
    uint2* raw = reinterpret_cast<uint2*>(ws);
    ws += ws_align((size_t)batch * d.nms_pre * sizeof(uint2));
    unsigned* keys = reinterpret_cast<unsigned*>(ws);
    ws += ws_align((size_t)batch * cap * sizeof(unsigned));
    unsigned* flats = reinterpret_cast<unsigned*>(ws);
    ws += ws_align((size_t)batch * cap * sizeof(unsigned));
    uint2* edge = reinterpret_cast<uint2*>(ws);
    OM_CUDA_TRY(cudaMemsetAsync(state, 0, (size_t)batch * sizeof(SelState), st));
    // sigma(x) <= conf_thresh for x <= logit(conf_thresh): such logits can never yield a candidate
    const double t = (double)d.conf_thresh;
    const float reject_logit = t >= 1.0 ? 3.0e38f : (float)(log(t / (1.0 - t)) - 0.01);
    const KeyBins kb = make_bins(d.conf_thresh);
    dim3 grid(om::ceil_div(d.n_pred, kSelThreads), batch);
    conf_compact_kernel<<<grid, kSelThreads, 0, st>>>(d, kb, state, keys, flats, cap, reject_logit);
    if ((rc = om::check_launch("conf_compact"))) return rc;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_image = (8 * sms + batch - 1) / batch;        // one wave of 8 resident blocks per SM
    if (per_image < 1) per_image = 1;
    const int max_useful = (int)((cap + kSelThreads - 1) / kSelThreads);
    if (per_image > max_useful) per_image = max_useful;
    dim3 lgrid(per_image, batch);
    select_edge_kernel<<<lgrid, kSelThreads, 0, st>>>(d, kb, state, keys, flats, cap, edge, raw);
    if ((rc = om::check_launch("select_edge"))) return rc;
    select_tail_kernel<<<batch, kTailThreads, (size_t)d.NP * 8, st>>>(d, kb, state, edge, cap, raw, cand_count, cand_det, cand_cls, cand_pred);
    return om::check_launch("select_tail");
}

extern "C" int32_t om_batched_nms(const om_post_config* cfg, const int32_t* cand_count, const float* cand_det,
                                  const int32_t* cand_cls, const int32_t* cand_pred, int32_t batch, int32_t* det_count,
                                  float* det, int64_t* det_cls, int32_t* det_anchor, int32_t* det_keep, float* records, void* stream) {
    PostDev d;
    int32_t rc = make_dev(cfg, d);
    if (rc) return rc;
    if (batch < 1 || !cand_count || !cand_det || !cand_cls || !cand_pred || !det_count || !det || !det_cls || !det_anchor || !det_keep)
        return om::fail(OM_ERR_INVALID, "om_batched_nms: null argument or batch < 1");
    NmsArgs a{};
    a.dets = cand_det; a.cls = cand_cls; a.pred = cand_pred; a.counts = cand_count;
    a.n = 0; a.cap = d.nms_pre; a.NP = d.NP; a.nms_post = d.nms_post; a.thr = d.nms_thresh;
    a.mode = cfg->nms_semantics == OM_NMS_CUDA ? OM_NMS_CUDA : OM_NMS_CPU;
    a.det_count = det_count; a.det = det; a.det_cls = reinterpret_cast<long long*>(det_cls);
    a.det_anchor = det_anchor; a.det_keep = det_keep; a.records = records;
    const size_t smem = nms_smem_bytes(a.NP, a.cap);
    if ((rc = allow_smem(nms_kernel<true>, smem))) return rc;
    nms_kernel<true><<<batch, 512, smem, (cudaStream_t)stream>>>(a, d);
    return om::check_launch("nms_kernel<batched>");
}

extern "C" int32_t om_nms(const float* dets, int32_t n, float threshold, int64_t* keep, int32_t* keep_count, void* stream) {
    return om_nms_ex(dets, n, threshold, OM_NMS_CPU, keep, keep_count, stream);
}

extern "C" int32_t om_nms_ex(const float* dets, int32_t n, float threshold, int32_t semantics, int64_t* keep, int32_t* keep_count, void* stream) {
    if (n < 0 || n > 1024) return om::fail(OM_ERR_INVALID, "om_nms: n=%d outside [0, 1024]", n);
    if (semantics != OM_NMS_CPU && semantics != OM_NMS_CUDA) return om::fail(OM_ERR_INVALID, "om_nms: unknown semantics %d", semantics);
    if (!keep_count || (n > 0 && (!dets || !keep))) return om::fail(OM_ERR_INVALID, "om_nms: null argument");
    if (n == 0) {
        OM_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), (cudaStream_t)stream));
        return OM_OK;
    }
    NmsArgs a{};
    a.dets = dets; a.n = n; a.cap = n; a.NP = next_pow2(n); a.nms_post = n; a.thr = threshold; a.mode = semantics;
    a.keep = reinterpret_cast<long long*>(keep); a.keep_count = keep_count;
    PostDev d;
    memset(&d, 0, sizeof(d));
    const size_t smem = nms_smem_bytes(a.NP, a.cap);
    int32_t rc;
    if ((rc = allow_smem(nms_kernel<false>, smem))) return rc;
    nms_kernel<false><<<1, 512, smem, (cudaStream_t)stream>>>(a, d);
    return om::check_launch("nms_kernel<single>");
}

extern "C" int32_t om_mask_assemble(const om_post_config* cfg, const float* const* orien, const int64_t* orien_batch_stride,
                                    const int32_t* det_count, const float* det, const int32_t* det_anchor, int32_t batch,
                                    uint8_t* mask, void* stream) {
    PostDev d;
    int32_t rc = make_dev(cfg, d);
    if (rc) return rc;
    if (batch < 1 || !orien || !orien_batch_stride || !det_count || !det || !det_anchor || !mask)
        return om::fail(OM_ERR_INVALID, "om_mask_assemble: null argument or batch < 1");
    for (int s = 0; s < d.num_scales; ++s) {
        if (!orien[s]) return om::fail(OM_ERR_INVALID, "om_mask_assemble: orien[%d] is null", s);
        d.orien[s] = orien[s]; d.ostride[s] = orien_batch_stride[s];
    }
    if (reinterpret_cast<uintptr_t>(mask) & 15) return om::fail(OM_ERR_INVALID, "om_mask_assemble: mask must be 16-byte aligned");
    const int units = d.H * (d.W >> 4);
    dim3 grid(om::ceil_div(units, 256), batch);
    mask_kernel<<<grid, 256, (size_t)d.nms_post * sizeof(MaskInst), (cudaStream_t)stream>>>(d, det_count, det, det_anchor, mask);
    return om::check_launch("mask_kernel");
}
