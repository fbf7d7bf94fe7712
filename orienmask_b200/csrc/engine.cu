// Whole-network entry points of the C ABI: om_engine_create / om_forward / om_engine_destroy.
//
// The launch schedule of OrienMaskYOLOFPNPlus.forward (model/orienmask_yolo_fpnplus.py:74-90 over DarkNet53.forward,
// model/backbone/darknet.py:47-54) lives HERE: a C caller hands over the reference's state dict as named fp32 device tensors plus one
// workspace and gets the 93-launch (fp16; 95 in the other precisions) forward behind a single call.  What the schedule does (same as the Python-scheduled engine it
// replaces, which tests keep as a bit-exact cross-check):
//   * BatchNorm folded into the convolution weights in fp32 on the device (W' = W * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps),
//     model/base.py:104-137), packed once into the engine layout of the chosen precision (om_conv_desc.weights);
//   * one kernel per ConvBNLeaky (bias + LeakyReLU in the epilogue), the residual add of a DarkNet block fused into its 3x3, in place;
//   * cat([nearest_up(route), x]) -> 1x1 conv evaluated as W_x * x + nearest_up(W_r * route): the coarse product is a small fp32
//     "partial" buffer that the fine layer adds in its epilogue (exact re-association; neither the up-sampled nor the concatenated
//     tensor exists);
//   * a static buffer plan inside the caller's workspace ("padded-row NHWC", include/orienmask_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_plan.h"

namespace {

constexpr int kStageChannels[5] = {32, 64, 128, 256, 512};
constexpr int kStageBlocks[5] = {1, 2, 8, 8, 4};

struct Tensor { const float* data; long long numel; };

struct Buf {
    char* ptr = nullptr;
    int stride = 0, c = 0;
    bool s2d = false;
};

struct Layer {
    std::string name;
    om_conv_desc desc;          // stem: only the geometry fields are meaningful
    om_conv* conv = nullptr;    // null for the stem
    int head_slot = -1;         // 0..2 bbox scales (32, 16, 8), 3 orientation head
    bool is_stem = false;
    const float* stem_w = nullptr;
    const float* stem_b = nullptr;
    bool is_stem2 = false;      // stem + backbone.conv2.0 in one launch (stem_fused.cu): stem_w / stem_b + blk_w2 / blk_b2, desc.output = x
    bool is_block = false;      // fused DarkNet block (dark_block.cu): desc.input = x, desc.output = out, the two weight sets below
    const void* blk_w1 = nullptr; const void* blk_w2 = nullptr;
    const float* blk_b1 = nullptr; const float* blk_b2 = nullptr;
    double flops = 0, bytes = 0;
    char shape[96];
    // small-batch branch concurrency (om_engine.lanes): stream lane of the launch, events it waits for / records
    int lane = 0, wait_ev = -1, record_ev = -1;
};

// Events of the branch schedule: a feature the side lanes consume is ready / a side lane has finished
enum { EV_X4 = 0, EV_N32, EV_N16, EV_N8, EV_SKIPS, EV_HEADS, EV_COUNT };

// ---- weight folding / packing on the device -----------------------------------------------------------------------------------
// One thread per (tap, cout, cin-of-the-slice) element.  Every arithmetic step is a single-rounded fp32 operation in the order the
// Python-scheduled engine (and torch) performs it, so both engines hold bit-identical packed weights.
struct FoldArgs {
    const float* w;          // [cout][cin_full][k][k]
    const float* gamma; const float* beta; const float* mean; const float* var;   // null for a plain conv
    const float* conv_bias;  // plain conv
    int cout, cin_full, col0, cin, kk;
    int precision, cpad;
    float wscale;            // split precision: 2^s
    void* out_w; float* out_bias; unsigned int* amax_bits;
};

__device__ __forceinline__ float bn_scale(const FoldArgs& a, int co) {
    if (a.gamma == nullptr) return 1.0f;
    return __fdiv_rn(a.gamma[co], sqrtf(__fadd_rn(a.var[co], 1e-5f)));
}

__global__ void fold_amax_kernel(FoldArgs a) {
    const long long n = (long long)a.kk * a.cout * a.cin;
    float m = 0.0f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % a.cin);
        const int co = (int)((i / a.cin) % a.cout);
        const int tap = (int)(i / ((long long)a.cin * a.cout));
        const float v = __fmul_rn(a.w[((long long)co * a.cin_full + a.col0 + ci) * a.kk + tap], bn_scale(a, co));
        m = fmaxf(m, fabsf(v));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(a.amax_bits, __float_as_uint(m));
}

__global__ void fold_pack_kernel(FoldArgs a) {
    const long long n = (long long)a.kk * a.cout * a.cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % a.cin);
        const int co = (int)((i / a.cin) % a.cout);
        const int tap = (int)(i / ((long long)a.cin * a.cout));
        const float v = __fmul_rn(a.w[((long long)co * a.cin_full + a.col0 + ci) * a.kk + tap], bn_scale(a, co));
        if (a.precision == OM_PREC_F32) {
            reinterpret_cast<float*>(a.out_w)[((long long)tap * a.cin + ci) * a.cpad + co] = v;
        } else if (a.precision == OM_PREC_F16) {
            reinterpret_cast<__half*>(a.out_w)[((long long)tap * a.cpad + co) * a.cin + ci] = __float2half_rn(v);
        } else {
            const float s = __fmul_rn(v, a.wscale);
            const __half hi = __float2half_rn(s);
            const __half lo = __float2half_rn(__fsub_rn(s, __half2float(hi)));
            __half* row = reinterpret_cast<__half*>(a.out_w) + ((long long)tap * a.cpad + co) * (2 * a.cin);
            row[ci] = hi;
            row[a.cin + ci] = lo;
        }
    }
    if (a.out_bias != nullptr) {
        for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < a.cout; co += gridDim.x * blockDim.x)
            a.out_bias[co] = a.gamma ? __fsub_rn(a.beta[co], __fmul_rn(a.mean[co], bn_scale(a, co))) : a.conv_bias[co];
    }
}

}  // namespace

struct om_engine {
    om_engine_config cfg;
    std::unordered_map<std::string, Tensor> sd;
    std::vector<Layer> layers;
    char* ws = nullptr;
    size_t ws_bytes = 0, used = 0;
    bool sizing = false;        // om_engine_workspace_bytes: walk the schedule, allocate nothing, plan nothing
    cudaStream_t stream = nullptr;
    unsigned int* amax = nullptr;
    int cmul = 1, esz = 2;
    int32_t rc = OM_OK;
    // Small batches (every layer far below one wave of CTAs): the box heads and the skip / partial branch of the stride-4 neck do not
    // sit on the critical path backbone -> necks -> neck4 -> orientation head; they run on two side streams, forked and joined by
    // events (a CUDA-graph capture of om_forward records the same fork / join).  Large batches keep one stream: concurrent 227 KB CTAs
    // do not co-reside and the launches would only compete for SMs (DESIGN finding 14).
    bool lanes = false;
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev[EV_COUNT] = {};
    Buf c1;                     // stem output

    int rows(int stride) const {
        // Rows per image of the padded-row layout (see model.py history): strides 1 / 2 / 4 are tied by the parity-split stride-2
        // layers (rows(in) == 2 * rows(out)) and the stride-4 maps carry ONE pad row; strides 8 / 16 / 32 keep rows(s) == 2 * rows(2s)
        // so that their up-add sources can be staged by TMA.
        if (stride <= 4) return (cfg.height / 4 + 1) * (4 / stride);
        return cfg.height / stride + 32 / stride;
    }

    char* alloc(size_t bytes) {
        const size_t off = (used + 1023) & ~(size_t)1023;
        used = off + bytes;
        if (sizing) return reinterpret_cast<char*>(0x1000000 + off);      // never dereferenced
        if (used > ws_bytes) { if (rc == OM_OK) rc = om::fail(OM_ERR_INVALID, "om_engine_create: workspace of %zu bytes is too small", ws_bytes); return nullptr; }
        return ws + off;
    }

    Buf act(int stride, int channels, bool f32 = false, bool s2d = false) {
        Buf b;
        b.stride = stride; b.c = channels; b.s2d = s2d;
        const size_t elems = (size_t)cfg.batch * rows(stride) * (cfg.width / stride) * channels;
        b.ptr = alloc(f32 ? elems * 4 : elems * (size_t)cmul * esz);
        return b;
    }

    const Tensor* find(const std::string& key, long long numel) {
        if (sizing) return nullptr;
        auto it = sd.find(key);
        if (it == sd.end()) { if (rc == OM_OK) rc = om::fail(OM_ERR_INVALID, "om_engine_create: state dict has no '%s'", key.c_str()); return nullptr; }
        if (it->second.numel != numel) {
            if (rc == OM_OK) rc = om::fail(OM_ERR_INVALID, "om_engine_create: '%s' has %lld elements, expected %lld", key.c_str(), it->second.numel, numel);
            return nullptr;
        }
        return &it->second;
    }

    // Fold + pack the weights of `prefix` (columns [col0, col0 + cin) of its cin_full input channels); returns false on error.
    bool weights(const std::string& prefix, bool cbl, int cout, int cin_full, int col0, int cin, int k, bool want_bias,
                 const void** out_w, const float** out_b, float* acc_scale, int precision) {
        const int kk = k * k;
        const int cpad = precision == OM_PREC_F32 ? (cout + 3) / 4 * 4 : (cout + 31) / 32 * 32;
        const size_t wbytes = precision == OM_PREC_F32 ? (size_t)kk * cin * cpad * 4
                                                       : (size_t)kk * cpad * cin * 2 * (precision == OM_PREC_SPLIT ? 2 : 1);
        char* wbuf = alloc(wbytes);
        float* bbuf = want_bias ? reinterpret_cast<float*>(alloc((size_t)cout * 4)) : nullptr;
        *out_w = wbuf; *out_b = bbuf; *acc_scale = 1.0f;
        if (sizing) return true;
        FoldArgs a = {};
        const std::string wkey = prefix + (cbl ? ".conv_block.0.weight" : ".weight");
        const Tensor* w = find(wkey, (long long)cout * cin_full * kk);
        if (cbl) {
            const Tensor* g = find(prefix + ".conv_block.1.weight", cout);
            const Tensor* b = find(prefix + ".conv_block.1.bias", cout);
            const Tensor* m = find(prefix + ".conv_block.1.running_mean", cout);
            const Tensor* v = find(prefix + ".conv_block.1.running_var", cout);
            if (!w || !g || !b || !m || !v) return false;
            a.gamma = g->data; a.beta = b->data; a.mean = m->data; a.var = v->data;
        } else {
            const Tensor* b = find(prefix + ".bias", cout);
            if (!w || !b) return false;
            a.conv_bias = b->data;
        }
        if (rc != OM_OK || !wbuf) return false;
        a.w = w->data; a.cout = cout; a.cin_full = cin_full; a.col0 = col0; a.cin = cin; a.kk = kk;
        a.precision = precision; a.cpad = cpad; a.wscale = 1.0f; a.out_w = wbuf; a.out_bias = bbuf; a.amax_bits = amax;
        const long long n = (long long)kk * cout * cin;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 2048);
        if (cudaMemsetAsync(wbuf, 0, wbytes, stream) != cudaSuccess) { rc = om::fail(OM_ERR_CUDA, "cudaMemsetAsync(weights)"); return false; }
        if (precision == OM_PREC_SPLIT) {
            // power-of-two scale 2^s that puts max|w| in [2^13, 2^14): W_lo stays a normal fp16 number; undone by acc_scale
            unsigned int bits = 0;
            cudaMemsetAsync(amax, 0, 4, stream);
            fold_amax_kernel<<<blocks, 256, 0, stream>>>(a);
            if (cudaMemcpyAsync(&bits, amax, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
                rc = om::fail(OM_ERR_CUDA, "om_engine_create: weight scale read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
                return false;
            }
            float amaxf;
            memcpy(&amaxf, &bits, 4);
            int s = amaxf > 0.0f ? 13 - (int)std::floor(std::log2((double)amaxf)) : 0;
            s = s < -24 ? -24 : (s > 40 ? 40 : s);
            a.wscale = std::ldexp(1.0f, s);
            *acc_scale = std::ldexp(1.0f, -s);
        }
        fold_pack_kernel<<<blocks, 256, 0, stream>>>(a);
        if (cudaGetLastError() != cudaSuccess) { rc = om::fail(OM_ERR_CUDA, "fold_pack_kernel launch failed"); return false; }
        return true;
    }

    // One convolution launch of the schedule.
    void conv(const std::string& name, const std::string& prefix, bool cbl, const Buf& src, int cin_full, int col0, int cin, int cout,
              const Buf* dst, int k, int stride, bool leaky, int out_kind, const Buf* residual, const Buf* upadd, bool with_bias, int head_slot) {
        if (rc != OM_OK) return;
        Layer L;
        L.name = name;
        om_conv_desc& d = L.desc;
        memset(&d, 0, sizeof(d));
        d.precision = cfg.precision; d.batch = cfg.batch;
        const int si = src.stride, so = si * stride;
        d.in_h = cfg.height / si; d.in_w = cfg.width / si; d.in_rows = rows(si);
        d.out_h = cfg.height / so; d.out_w = cfg.width / so; d.out_rows = rows(so);
        d.cin = cin; d.cout = cout; d.ksize = k; d.stride = stride; d.leaky = leaky ? 1 : 0; d.out_kind = out_kind;
        d.in_s2d = src.s2d ? 1 : 0;
        d.out_s2d = (dst && dst->s2d) ? 1 : 0;
        d.input = src.ptr;
        const void* wp = nullptr; const float* bp = nullptr;
        if (!weights(prefix, cbl, cout, cin_full, col0, cin, k, with_bias, &wp, &bp, &d.acc_scale, cfg.precision)) { if (!sizing) return; }
        d.weights = wp; d.bias = bp;
        if (out_kind == OM_OUT_NCHW) {
            d.cout_stride = cout;
            d.output = reinterpret_cast<void*>(0x1000);          // always redirected by om_forward (om_conv_run_to semantics)
        } else {
            d.cout_stride = dst->c;
            d.output = dst->ptr;
        }
        if (residual) d.residual = residual->ptr;
        if (upadd) { d.upadd = reinterpret_cast<const float*>(upadd->ptr); d.up_rows = rows(upadd->stride); }
        L.head_slot = head_slot;
        L.flops = 2.0 * d.batch * d.out_h * d.out_w * (double)cout * cin * k * k;
        const double e = cfg.precision == OM_PREC_F16 ? 2 : 4;
        L.bytes = (double)d.batch * d.in_h * d.in_w * cin * e + (double)cout * cin * k * k * e +
                  (double)d.batch * d.out_h * d.out_w * cout * (out_kind == OM_OUT_ACT ? e : 4) +
                  (residual ? (double)d.batch * d.out_h * d.out_w * cout * e : 0) + (upadd ? (double)d.batch * d.out_h * d.out_w * cout : 0);
        snprintf(L.shape, sizeof(L.shape), "%dx%d s%d %d->%d @%dx%d%s%s%s", k, k, stride, cin, cout, d.out_h, d.out_w, residual ? " +res" : "",
                 upadd ? " +up" : "", out_kind == OM_OUT_PARTIAL ? " partial" : out_kind == OM_OUT_NCHW ? " nchw" : "");
        if (!sizing) {
            const int32_t r = om_conv_create(&d, &L.conv);
            if (r != OM_OK) { rc = r; return; }
        }
        layers.push_back(L);
    }

    void layers_pop_conv() {
        if (rc != OM_OK || layers.empty()) return;
        if (layers.back().conv) om_conv_destroy(layers.back().conv);
        layers.pop_back();
    }

    // lane / event tags of the layer emitted last (no-ops once an error has stopped the emission)
    Layer scratch_layer;
    Layer& last() { return (rc == OM_OK && !layers.empty()) ? layers.back() : scratch_layer; }

    // Fused residual block x -> x + conv3x3(conv1x1(x)) with C squeezed channels, output into `dst` (never in place: neighbouring tiles
    // read each other's x halo)
    void block(const std::string& prefix, const Buf& x, int c, const Buf& dst) {
        if (rc != OM_OK) return;
        Layer L;
        L.name = prefix + " (fused block)";
        L.is_block = true;
        om_conv_desc& d = L.desc;
        memset(&d, 0, sizeof(d));
        d.precision = cfg.precision; d.batch = cfg.batch;
        d.in_h = d.out_h = cfg.height / x.stride; d.in_w = d.out_w = cfg.width / x.stride; d.in_rows = d.out_rows = rows(x.stride);
        d.cin = d.cout = 2 * c; d.ksize = 3; d.stride = 1; d.leaky = 1; d.out_kind = OM_OUT_ACT; d.cout_stride = 2 * c;
        d.input = x.ptr; d.output = dst.ptr; d.residual = x.ptr; d.out_s2d = dst.s2d ? 1 : 0;
        float sc;
        if (!weights(prefix + ".conv.0", true, c, 2 * c, 0, 2 * c, 1, true, &L.blk_w1, &L.blk_b1, &sc, OM_PREC_F16) && !sizing) return;
        if (!weights(prefix + ".conv.1", true, 2 * c, c, 0, c, 3, true, &L.blk_w2, &L.blk_b2, &sc, OM_PREC_F16) && !sizing) return;
        const double px = (double)d.batch * d.out_h * d.out_w;
        L.flops = 2.0 * px * (2.0 * c * c + 9.0 * c * 2 * c);
        L.bytes = px * (2 * c) * 2 * 2 + (2.0 * c * c + 18.0 * c * c) * 2;        // x read once, output written once, weights
        snprintf(L.shape, sizeof(L.shape), "1x1 %d->%d + 3x3 %d->%d @%dx%d +res fused", 2 * c, c, c, 2 * c, d.out_h, d.out_w);
        layers.push_back(L);
    }

    void cbl(const std::string& prefix, const Buf& src, int cin, int cout, const Buf& dst, int k, int stride = 1, const Buf* residual = nullptr) {
        conv(prefix, prefix, true, src, cin, 0, cin, cout, &dst, k, stride, true, OM_OUT_ACT, residual, nullptr, true, -1);
    }

    // conv_bn_leaky sequence <prefix>.0 .. <prefix>.4 (1x1, 3x3, 1x1, 3x3, 1x1 over c / 2c channels, or the orientation head's
    // 3x3, 1x1, 3x3, 1x1, 3x3); buffers alternate; optional concat-split on the first conv.
    Buf chain(const std::string& prefix, const Buf& src, int cin_full, int col0, int cin, const Buf bufs[2], const int* ks, const int* couts,
              int n, const Buf* first_upadd) {
        Buf cur = src;
        int c_in_full = cin_full, c0 = col0, c_in = cin;
        for (int i = 0; i < n; ++i) {
            const std::string p = prefix + "." + std::to_string(i);
            conv(p, p, true, cur, c_in_full, c0, c_in, couts[i], &bufs[i % 2], ks[i], 1, true, OM_OUT_ACT, nullptr, i == 0 ? first_upadd : nullptr, true, -1);
            cur = bufs[i % 2];
            c_in_full = c_in = couts[i]; c0 = 0;
        }
        return cur;
    }

    // fp32 pre-activation partial W[:, col0:col0+cin] * src (+ nearest-up of a coarser partial)
    Buf partial(const std::string& prefix, int cin_full, int col0, int cin, int cout, const Buf& src, const Buf* upadd) {
        Buf dst = act(src.stride, cout, true);
        const std::string name = prefix + "[" + std::to_string(col0) + ":" + std::to_string(col0 + cin) + "]";
        conv(name, prefix, true, src, cin_full, col0, cin, cout, &dst, 1, 1, false, OM_OUT_PARTIAL, nullptr, upadd, false, -1);
        return dst;
    }

    void head(const std::string& prefix, const Buf& src, int cin, int cout, int slot) {
        conv(prefix, prefix, false, src, cin, 0, cin, cout, nullptr, 1, 1, false, OM_OUT_NCHW, nullptr, nullptr, true, slot);
    }

    void build() {
        const int B = cfg.batch, H = cfg.height, W = cfg.width;
        cmul = cfg.precision == OM_PREC_SPLIT ? 2 : 1;
        esz = cfg.precision == OM_PREC_F32 ? 4 : 2;
        if (!sizing) amax = reinterpret_cast<unsigned int*>(alloc(16)); else alloc(16);
        // stem (model/backbone/darknet.py:41): always the fp32 [27][32] weights; its output is parity-split (read only by conv2.0)
        {
            Layer L;
            L.name = "backbone.conv1"; L.is_stem = true;
            memset(&L.desc, 0, sizeof(L.desc));
            const void* wp = nullptr; const float* bp = nullptr; float sc;
            weights("backbone.conv1", true, 32, 3, 0, 3, 3, true, &wp, &bp, &sc, OM_PREC_F32);
            L.stem_w = reinterpret_cast<const float*>(wp); L.stem_b = bp;
            if (!(cfg.precision == OM_PREC_F16 && om::stem_fused_supported(H, W, nullptr))) c1 = act(1, 32, false, true);   // (fused below: never materialised)
            L.desc.precision = cfg.precision; L.desc.batch = B; L.desc.in_h = L.desc.out_h = H; L.desc.in_w = L.desc.out_w = W;
            L.desc.in_rows = L.desc.out_rows = rows(1); L.desc.cin = 3; L.desc.cout = 32; L.desc.ksize = 3; L.desc.stride = 1; L.desc.leaky = 1;
            L.desc.output = c1.ptr; L.desc.out_s2d = 1;
            L.flops = 2.0 * B * H * W * 32 * 27;
            L.bytes = (double)B * H * W * (3 * 4 + 32 * (cfg.precision == OM_PREC_F16 ? 2 : 4));
            snprintf(L.shape, sizeof(L.shape), "3x3 s1 3->32 @%dx%d stem", H, W);
            layers.push_back(L);
        }
        // fp16 engine: the stem and backbone.conv2.0 run as ONE launch (stem_fused.cu) -- the stem's full-resolution output never
        // reaches memory.  The stem layer emitted above becomes that launch once conv2.0's weights and output buffer exist.
        const bool fuse_stem = cfg.precision == OM_PREC_F16 && om::stem_fused_supported(H, W, nullptr);
        Buf trunk = c1;
        Buf feats[6];            // by log2(stride)
        for (int i = 0; i < 5; ++i) {
            const int c = kStageChannels[i], n = kStageBlocks[i], st = 2 << i;
            const std::string stage = "backbone.conv" + std::to_string(i + 2);
            Buf x = act(st, 2 * c), y = act(st, c);
            if (i == 0 && fuse_stem && rc == OM_OK) {
                Layer& L = layers.back();
                float sc;
                if (!weights(stage + ".0", true, 2 * c, c, 0, c, 3, true, &L.blk_w2, &L.blk_b2, &sc, OM_PREC_F16) && !sizing) return;
                L.name = "backbone.conv1 + conv2.0 (fused)"; L.is_stem = false; L.is_stem2 = true;
                L.desc.output = x.ptr; L.desc.out_s2d = 0; L.desc.out_h = H / 2; L.desc.out_w = W / 2; L.desc.out_rows = rows(2); L.desc.cout = 2 * c;
                L.flops += 2.0 * B * (H / 2) * (W / 2) * (2.0 * c) * c * 9;
                L.bytes = (double)B * H * W * 3 * 4 + (double)B * (H / 2) * (W / 2) * 2 * c * 2 + (27.0 * 32 + 9.0 * 2 * c * c) * 2;
                snprintf(L.shape, sizeof(L.shape), "3x3 s1 3->32 @%dx%d + 3x3 s2 32->64 @%dx%d fused", H, W, H / 2, W / 2);
            } else
            cbl(stage + ".0", trunk, c, 2 * c, x, 3, 2);
            for (int b = 1; b <= n; ++b) {
                const std::string blk = stage + "." + std::to_string(b);
                cbl(blk + ".conv.0", x, 2 * c, c, y, 1);
                if (i == 0 && b == n) {                          // feeds only conv3.0 (stride 2): written parity-split
                    Buf xs = act(st, 2 * c, false, true);
                    if (cfg.precision == OM_PREC_F16 && om::dark_block_supported(2 * c, c, cfg.width / st, rows(st))) {
                        // the whole block in one launch (dark_block.cu): the 1x1 was emitted above as a separate layer -- replace it
                        layers_pop_conv();
                        block(blk, x, c, xs);
                    } else {
                        cbl(blk + ".conv.1", y, c, 2 * c, xs, 3, 1, &x);
                    }
                    x = xs;
                } else {
                    cbl(blk + ".conv.1", y, c, 2 * c, x, 3, 1, &x);   // in place: x += leaky(conv(y))
                }
            }
            trunk = x;
            feats[i + 1] = x;
            if (i == 1) last().record_ev = EV_X4;
        }
        const Buf &x4 = feats[2], &x8 = feats[3], &x16 = feats[4], &x32 = feats[5];
        const int ks[5] = {1, 3, 1, 3, 1};
        auto neck_couts = [](int c, int* out) { out[0] = c; out[1] = 2 * c; out[2] = c; out[3] = 2 * c; out[4] = c; };
        int co[5];

        Buf b32[2] = {act(32, 512), act(32, 1024)};
        neck_couts(512, co);
        Buf neck32 = chain("neck32", x32, 1024, 0, 1024, b32, ks, co, 5, nullptr);
        last().record_ev = EV_N32;
        Buf r32 = act(32, 256);
        cbl("route32.0", neck32, 512, 256, r32, 1);
        Buf p16 = partial("neck16.0", 768, 0, 256, 256, r32, nullptr);
        Buf b16[2] = {act(16, 256), act(16, 512)};
        neck_couts(256, co);
        Buf neck16 = chain("neck16", x16, 768, 256, 512, b16, ks, co, 5, &p16);
        last().record_ev = EV_N16;
        Buf r16 = act(16, 128);
        cbl("route16.0", neck16, 256, 128, r16, 1);
        Buf p8 = partial("neck8.0", 384, 0, 128, 128, r16, nullptr);
        Buf b8[2] = {act(8, 128), act(8, 256)};
        neck_couts(128, co);
        Buf neck8 = chain("neck8", x8, 384, 128, 256, b8, ks, co, 5, &p8);
        last().record_ev = EV_N8;

        const int nb = cfg.num_anchors * (5 + cfg.num_classes);
        const Buf* necks[3] = {&neck32, &neck16, &neck8};
        const int head_c[3] = {512, 256, 128}, head_s[3] = {32, 16, 8};
        for (int j = 0; j < 3; ++j) {
            Buf hb = act(head_s[j], 2 * head_c[j]);
            const std::string hp = "bbox_head" + std::to_string(head_s[j]);
            cbl(hp + ".0", *necks[j], head_c[j], 2 * head_c[j], hb, 3);
            last().lane = 1; last().wait_ev = EV_N32 + j;
            head(hp + ".1", hb, 2 * head_c[j], nb, j);
            last().lane = 1;
            if (j == 2) last().record_ev = EV_HEADS;
        }

        Buf ab4[2] = {act(4, 128), act(4, 256)};
        neck_couts(128, co);
        Buf neck4;
        if (cfg.plus) {
            Buf s32 = act(32, 64), s16 = act(16, 64), s8 = act(8, 64), s4 = act(4, 64);
            cbl("skip32.0", neck32, 512, 64, s32, 1);
            last().lane = 2; last().wait_ev = EV_N32;
            cbl("skip16.0", neck16, 256, 64, s16, 1);
            last().lane = 2; last().wait_ev = EV_N16;
            cbl("skip8.0", neck8, 128, 64, s8, 1);
            cbl("skip4", x4, 128, 64, s4, 1);
            last().lane = 2; last().wait_ev = EV_X4;
            Buf q32 = partial("neck4.0", 256, 0, 64, 128, s32, nullptr);
            last().lane = 2;
            Buf q16 = partial("neck4.0", 256, 64, 64, 128, s16, &q32);
            last().lane = 2; last().record_ev = EV_SKIPS;
            Buf q8 = partial("neck4.0", 256, 128, 64, 128, s8, &q16);
            last().wait_ev = EV_SKIPS;                          // main lane: q16 and (through stream order) s4 are ready
            neck4 = chain("neck4", s4, 256, 192, 64, ab4, ks, co, 5, &q8);
        } else {                                                     // model/orienmask_yolo.py:83: neck4(cat[route8(neck8) up2, x4])
            Buf r8 = act(8, 64);
            cbl("route8.0", neck8, 128, 64, r8, 1);
            Buf p4 = partial("neck4.0", 192, 0, 64, 128, r8, nullptr);
            neck4 = chain("neck4", x4, 192, 64, 128, ab4, ks, co, 5, &p4);
        }
        // orien_head.0-4 alternate 3x3 (128->256) and 1x1 (256->128): neck4 lives in ab4[0], so start on ab4[1]
        const Buf ba4[2] = {ab4[1], ab4[0]};
        const int oks[5] = {3, 1, 3, 1, 3}, oco[5] = {256, 128, 256, 128, 256};
        Buf o = chain("orien_head", neck4, 128, 0, 128, ba4, oks, oco, 5, nullptr);
        head("orien_head.5", o, 256, cfg.num_anchors * 6, 3);
    }
};

static int32_t check_config(const om_engine_config* cfg) {
    if (!cfg) return om::fail(OM_ERR_INVALID, "om_engine: null config");
    if (cfg->precision != OM_PREC_F32 && cfg->precision != OM_PREC_F16 && cfg->precision != OM_PREC_SPLIT)
        return om::fail(OM_ERR_INVALID, "om_engine: unknown precision %d", cfg->precision);
    if (cfg->batch < 1 || cfg->height < 32 || cfg->width < 32 || cfg->height % 32 || cfg->width % 32)
        return om::fail(OM_ERR_INVALID, "om_engine: batch >= 1 and H, W positive multiples of 32 are required (got %d x %d x %d)", cfg->batch,
                        cfg->height, cfg->width);
    if (cfg->num_anchors < 1 || cfg->num_classes < 1) return om::fail(OM_ERR_INVALID, "om_engine: num_anchors / num_classes must be positive");
    return OM_OK;
}

extern "C" int32_t om_engine_workspace_bytes(const om_engine_config* cfg, size_t* bytes) {
    int32_t rc = check_config(cfg);
    if (rc != OM_OK) return rc;
    if (!bytes) return om::fail(OM_ERR_INVALID, "om_engine_workspace_bytes: null output");
    om_engine e;
    e.cfg = *cfg;
    e.sizing = true;
    e.build();
    *bytes = e.used + 2048;          // + alignment slack of the caller's base pointer
    return OM_OK;
}

extern "C" void om_engine_destroy(om_engine* e) {
    if (!e) return;
    for (int i = 0; i < 2; ++i)
        if (e->side[i]) cudaStreamDestroy(e->side[i]);
    for (int i = 0; i < EV_COUNT; ++i)
        if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (Layer& L : e->layers)
        if (L.conv) om_conv_destroy(L.conv);
    delete e;
}

extern "C" int32_t om_engine_create(const om_engine_config* cfg, const om_tensor* weights, int32_t n_weights, void* workspace,
                                    size_t workspace_bytes, void* stream, om_engine** out) {
    int32_t rc = check_config(cfg);
    if (rc != OM_OK) return rc;
    if (!weights || n_weights < 1 || !workspace || !out) return om::fail(OM_ERR_INVALID, "om_engine_create: null argument");
    const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023;       // buffers are carved out 1024-byte aligned
    const size_t slack = base - reinterpret_cast<uintptr_t>(workspace);
    if (workspace_bytes <= slack) return om::fail(OM_ERR_INVALID, "om_engine_create: workspace too small");
    om_engine* e = new om_engine();
    e->cfg = *cfg;
    e->ws = reinterpret_cast<char*>(base);
    e->ws_bytes = workspace_bytes - slack;
    e->stream = (cudaStream_t)stream;
    for (int i = 0; i < n_weights; ++i) {
        if (!weights[i].name || !weights[i].data) { delete e; return om::fail(OM_ERR_INVALID, "om_engine_create: weight %d has a null name or pointer", i); }
        e->sd[weights[i].name] = Tensor{weights[i].data, (long long)weights[i].numel};
    }
    // activations must start zeroed: the pad rows between images are never written
    if (cudaMemsetAsync(workspace, 0, workspace_bytes, e->stream) != cudaSuccess) { delete e; return om::fail(OM_ERR_CUDA, "om_engine_create: cudaMemsetAsync(workspace) failed"); }
    e->build();
    if (e->rc != OM_OK) { rc = e->rc; om_engine_destroy(e); return rc; }
    e->sd.clear();
    {
        const char* le = getenv("ORIENMASK_B200_LANES");
        const long long cells32 = (long long)cfg->batch * (cfg->height / 32) * (cfg->width / 32);
        e->lanes = !(le && le[0] == '0') && (cells32 <= 1200 || (le && le[0] == '1'));       // up to ~bs 4 at 544x544
        if (e->lanes) {
            bool ok = true;
            for (int i = 0; i < 2 && ok; ++i) ok = cudaStreamCreateWithFlags(&e->side[i], cudaStreamNonBlocking) == cudaSuccess;
            for (int i = 0; i < EV_COUNT && ok; ++i) ok = cudaEventCreateWithFlags(&e->ev[i], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) { om_engine_destroy(e); return om::fail(OM_ERR_CUDA, "om_engine_create: cannot create the side streams / events"); }
        }
    }
    *out = e;
    return OM_OK;
}

static int32_t run_layer(const om_engine* e, const Layer& L, const float* image, float* const* bbox, float* orien, cudaStream_t st) {
    if (L.is_stem) {
        if (!image) return om::fail(OM_ERR_INVALID, "om_forward: null image");
        return om_stem_conv(e->cfg.precision, image, L.stem_w, L.stem_b, L.desc.output, e->cfg.batch, e->cfg.height, e->cfg.width, L.desc.in_rows, 32,
                            1, st);
    }
    if (L.is_stem2) {
        if (!image) return om::fail(OM_ERR_INVALID, "om_forward: null image");
        if (reinterpret_cast<uintptr_t>(image) & 15) return om::fail(OM_ERR_INVALID, "om_forward: the fp16 engine needs a 16-byte aligned image");
        return om::stem_fused_run(image, L.stem_w, L.stem_b, L.blk_w2, L.blk_b2, L.desc.output, e->cfg.batch, e->cfg.height, e->cfg.width,
                                  L.desc.out_rows, st);
    }
    if (L.is_block)
        return om::dark_block_run(L.desc.input, L.blk_w1, L.blk_b1, L.blk_w2, L.blk_b2, L.desc.output, e->cfg.batch, L.desc.out_h, L.desc.out_w,
                                  L.desc.out_rows, L.desc.out_s2d, st);
    if (L.head_slot >= 0) {
        void* dst = L.head_slot < 3 ? (bbox ? (void*)bbox[L.head_slot] : nullptr) : (void*)orien;
        if (!dst) return om::fail(OM_ERR_INVALID, "om_forward: null output pointer for head %d", L.head_slot);
        return om_conv_run_to(L.conv, dst, st);
    }
    return om_conv_run(L.conv, st);
}

extern "C" int32_t om_forward(const om_engine* e, const float* image, float* const* bbox, float* orien, void* stream) {
    if (!e) return om::fail(OM_ERR_INVALID, "om_forward: null engine");
    if (!e->lanes) {
        for (const Layer& L : e->layers) {
            const int32_t rc = run_layer(e, L, image, bbox, orien, (cudaStream_t)stream);
            if (rc != OM_OK) return rc;
        }
        return OM_OK;
    }
    cudaStream_t lane[3] = {(cudaStream_t)stream, e->side[0], e->side[1]};
    bool used[3] = {true, false, false}, recorded[EV_COUNT] = {};
    for (const Layer& L : e->layers) {
        cudaStream_t st = lane[L.lane];
        if (L.wait_ev >= 0 && recorded[L.wait_ev]) OM_CUDA_TRY(cudaStreamWaitEvent(st, e->ev[L.wait_ev], 0));
        used[L.lane] = true;
        const int32_t rc = run_layer(e, L, image, bbox, orien, st);
        if (rc != OM_OK) return rc;
        if (L.record_ev >= 0) { OM_CUDA_TRY(cudaEventRecord(e->ev[L.record_ev], st)); recorded[L.record_ev] = true; }
    }
    // join: everything the side lanes did is ordered before whatever the caller enqueues next on `stream`
    if (recorded[EV_HEADS]) OM_CUDA_TRY(cudaStreamWaitEvent(lane[0], e->ev[EV_HEADS], 0));
    if (recorded[EV_SKIPS]) OM_CUDA_TRY(cudaStreamWaitEvent(lane[0], e->ev[EV_SKIPS], 0));
    (void)used;
    return OM_OK;
}

extern "C" int32_t om_engine_layer_count(const om_engine* e) { return e ? (int32_t)e->layers.size() : 0; }

extern "C" int32_t om_engine_layer_info(const om_engine* e, int32_t index, om_layer_info* info) {
    if (!e || !info || index < 0 || index >= (int32_t)e->layers.size()) return om::fail(OM_ERR_INVALID, "om_engine_layer_info: bad argument");
    const Layer& L = e->layers[index];
    memset(info, 0, sizeof(*info));
    snprintf(info->name, sizeof(info->name), "%s", L.name.c_str());
    snprintf(info->shape, sizeof(info->shape), "%s", L.shape);
    info->flops = L.flops; info->bytes = L.bytes; info->head_slot = L.head_slot; info->is_stem = L.is_stem ? 1 : 0;
    info->desc = L.desc;
    info->conv = L.conv;
    return OM_OK;
}

extern "C" int32_t om_engine_run_layer(const om_engine* e, int32_t index, const float* image, float* const* bbox, float* orien, void* stream) {
    if (!e || index < 0 || index >= (int32_t)e->layers.size()) return om::fail(OM_ERR_INVALID, "om_engine_run_layer: bad argument");
    return run_layer(e, e->layers[index], image, bbox, orien, (cudaStream_t)stream);
}

extern "C" int32_t om_engine_run_layers(const om_engine* e, const int32_t* indices, int32_t n, const float* image, float* const* bbox, float* orien,
                                        void* stream) {
    if (!e || !indices || n < 0) return om::fail(OM_ERR_INVALID, "om_engine_run_layers: bad argument");
    for (int32_t i = 0; i < n; ++i) {
        if (indices[i] < 0 || indices[i] >= (int32_t)e->layers.size()) return om::fail(OM_ERR_INVALID, "om_engine_run_layers: index %d out of range", indices[i]);
        const int32_t rc = run_layer(e, e->layers[indices[i]], image, bbox, orien, (cudaStream_t)stream);
        if (rc != OM_OK) return rc;
    }
    return OM_OK;
}
