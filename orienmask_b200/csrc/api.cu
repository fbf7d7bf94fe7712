// C-ABI plumbing: error text, launch accounting, convolution plan objects.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_plan.h"

namespace om {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

char* error_buffer() { return g_err; }

int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += n; }

static unsigned long long* g_trace_base = nullptr;
static int g_trace_cap = 0, g_trace_used = 0;
unsigned long long* trace_next() {
    if (g_trace_base == nullptr || g_trace_used >= g_trace_cap) return nullptr;
    return g_trace_base + 4 * (size_t)(g_trace_used++);
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ORIENMASK_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

}  // namespace om

struct om_conv {
    om_conv_desc desc;
    void* tc_plan;       // CTA-pair tcgen05 plan (OM_PREC_F16 / OM_PREC_SPLIT); null for the FFMA engine
};

extern "C" int32_t om_abi_version(void) { return 8; }
extern "C" const char* om_last_error(void) { return om::error_buffer(); }
extern "C" int64_t om_launch_count(void) { return om::g_launches; }
extern "C" void om_launch_count_reset(void) { om::g_launches = 0; }

extern "C" int32_t om_conv_create(const om_conv_desc* d, om_conv** out) {
    if (!d || !out) return om::fail(OM_ERR_INVALID, "om_conv_create: null argument");
    if (d->precision != OM_PREC_F32 && d->precision != OM_PREC_F16 && d->precision != OM_PREC_SPLIT)
        return om::fail(OM_ERR_INVALID, "unknown precision %d", d->precision);
    if (d->ksize != 1 && d->ksize != 3) return om::fail(OM_ERR_UNSUPPORTED, "ksize must be 1 or 3 (got %d)", d->ksize);
    if (d->stride != 1 && d->stride != 2) return om::fail(OM_ERR_UNSUPPORTED, "stride must be 1 or 2 (got %d)", d->stride);
    if (d->batch < 1 || d->in_h < 1 || d->in_w < 1 || d->out_h < 1 || d->out_w < 1 || d->cin < 1 || d->cout < 1)
        return om::fail(OM_ERR_INVALID, "om_conv_create: non-positive dimension");
    if (d->in_rows <= d->in_h) return om::fail(OM_ERR_INVALID, "in_rows must exceed in_h (zero rows between images)");
    if (d->out_kind != OM_OUT_NCHW && d->out_rows <= d->out_h) return om::fail(OM_ERR_INVALID, "out_rows must exceed out_h");
    if (d->out_kind < OM_OUT_ACT || d->out_kind > OM_OUT_NCHW) return om::fail(OM_ERR_INVALID, "unknown out_kind %d", d->out_kind);
    if (d->out_h != (d->in_h + d->stride - 1) / d->stride || d->out_w != (d->in_w + d->stride - 1) / d->stride)
        return om::fail(OM_ERR_INVALID, "output geometry does not match 'same' padding with stride %d", d->stride);
    if (!d->input || !d->weights || !d->output) return om::fail(OM_ERR_INVALID, "om_conv_create: null tensor pointer");
    if (d->out_kind != OM_OUT_NCHW && d->cout_stride < d->cout)
        return om::fail(OM_ERR_INVALID, "cout_stride %d is smaller than cout %d", d->cout_stride, d->cout);
    if (d->residual && d->out_kind != OM_OUT_ACT) return om::fail(OM_ERR_INVALID, "residual only with activation outputs");
    if (d->upadd && d->up_rows < 1) return om::fail(OM_ERR_INVALID, "upadd needs up_rows");
    if (d->in_s2d && (d->stride != 2 || d->in_rows % 2 || d->in_w % 2)) return om::fail(OM_ERR_INVALID, "in_s2d needs a stride-2 layer over even rows/width");
    if (d->out_s2d && (d->out_kind != OM_OUT_ACT || d->out_rows % 2 || d->out_w % 2 || d->cout_stride != d->cout))
        return om::fail(OM_ERR_INVALID, "out_s2d needs a dense OM_OUT_ACT output over even rows/width");
    if (d->out_s2d && d->residual == d->output && d->residual) return om::fail(OM_ERR_INVALID, "out_s2d cannot be written in place over the residual");
    om_conv* c = new om_conv();
    c->desc = *d;
    c->tc_plan = nullptr;
    if (d->precision != OM_PREC_F32) {
        int32_t rc = om::tc2_plan_create(*d, &c->tc_plan);
        if (rc != OM_OK) { delete c; return rc; }
    }
    *out = c;
    return OM_OK;
}

extern "C" int32_t om_conv_run(const om_conv* c, void* stream) {
    if (!c) return om::fail(OM_ERR_INVALID, "om_conv_run: null plan");
    if (c->tc_plan) return om::tc2_plan_run(c->tc_plan, (cudaStream_t)stream);
    return om::f32_conv_run(c->desc, (cudaStream_t)stream);
}

extern "C" int32_t om_conv_run_to(const om_conv* c, void* output, void* stream) {
    if (!c || !output) return om::fail(OM_ERR_INVALID, "om_conv_run_to: null argument");
    if (c->tc_plan) return om::tc2_plan_run(c->tc_plan, (cudaStream_t)stream, output);
    if (c->desc.residual) return om::fail(OM_ERR_INVALID, "om_conv_run_to: layers with a residual write in place");
    om_conv_desc d = c->desc;
    d.output = output;
    return om::f32_conv_run(d, (cudaStream_t)stream);
}

extern "C" void om_conv_destroy(om_conv* c) {
    if (!c) return;
    if (c->tc_plan) om::tc2_plan_destroy(c->tc_plan);
    delete c;
}

// Debug: device buffer of >= 16 uint64 that cluster 0 of the next conv_tc2 launches fills with %globaltimer stamps (prologue / first
// load / first MMA / first epilogue / exit).
namespace om { int32_t tc2_set_timeline(void* dev_ptr); }
extern "C" int32_t om_debug_conv_timeline(void* dev_ptr) { return om::tc2_set_timeline(dev_ptr); }
// Debug: the planner's decisions for a layer of the CTA-pair engine -- info[24] = halo, flat, halo_s2, b_resident, tw, th,
// block_n, tiles_n, stages, n_sub, h_stages, acc_stages, has_res (1 fp16 residual, 2 staged up-add), res_direct, smem bytes, grid,
// CTA-pair tiles, taps, K chunks, BK, TMEM columns, cout, out_h, out_w.
extern "C" int32_t om_debug_conv_plan_info(const om_conv* c, int32_t* info) {
    if (!c || !info || !c->tc_plan)
        return om::fail(OM_ERR_INVALID, "om_debug_conv_plan_info: not a plan of the CTA-pair engine");
    om::tc2_plan_info(c->tc_plan, info);
    return OM_OK;
}

// Debug: arm (records != NULL) or disarm the launch trace.  `records` is a device buffer of capacity x 4 uint64 that the caller has
// filled with {~0, ~0, 0, 0} per record; every following conv-engine launch (conv_tc2 / stem / fused block) takes the next record
// and stamps %globaltimer into it (first CTA start, dependencies resolved, last CTA end, last CTA start).  Returns the number of
// records handed out since the previous call.  tools/timeline.py turns this into the in-situ timeline of a forward.
extern "C" int32_t om_debug_trace(void* records, int32_t capacity) {
    const int32_t used = om::g_trace_used;
    om::g_trace_base = reinterpret_cast<unsigned long long*>(records);
    om::g_trace_cap = records ? capacity : 0;
    om::g_trace_used = 0;
    return used;
}

// Debug: device buffer of 16 x 8 uint64 that stream 0 of CTA 0 of the fused stem kernel fills with %globaltimer stamps at the phase
// boundaries of its first 16 tiles (tile start, patch arrived, im2col done, MMA 1 done, epilogue 1 done, MMA 2 done, epilogue 2 done).
namespace om { int32_t stem_fused_set_phase_log(void* dev_ptr); }
extern "C" int32_t om_debug_phase_log(void* device_u64x128) { return om::stem_fused_set_phase_log(device_u64x128); }
