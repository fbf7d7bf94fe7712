/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's greedy NMS.  Not product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline legs may load this library.
 *
 * Follows /root/reference/eval/src/nms_cpu.cpp:4-63 (semantics, not text):
 *   - boxes arrive as (cx, cy, w, h, score) rows, fp32                           (:12-16)
 *   - corners are cx -/+ w/2, cy -/+ h/2; area = (x2-x1)*(y2-y1) from corners     (:17-22)
 *   - visit order = scores sorted descending                                      (:24)
 *     (the reference calls an unstable sort; ties are broken here by ascending index,
 *      which is what ATen's CPU sort does in practice)
 *   - a later box is suppressed by a kept earlier box iff
 *         inter / (area_i + area_j - inter) >= threshold                          (:52-59)
 *   - result = indices of un-suppressed boxes in ASCENDING ORIGINAL INDEX order   (:62)
 * All arithmetic is single-rounded fp32 (build with -ffp-contract=off, no -ffast-math).
 *
 * Pinned against the compiled reference (oracle/_ref/ref_nms_cpu.so) by tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float s; int32_t i; } om_key;

static int cmp_desc(const void* a, const void* b) {
    const om_key* x = (const om_key*)a;
    const om_key* y = (const om_key*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* dets: [n,5] fp32 row-major.  keep: out, capacity n.  returns number kept. */
int32_t om_oracle_nms(const float* dets, int32_t n, float threshold, int64_t* keep) {
    if (n <= 0) return 0;
    float* x1 = (float*)malloc(sizeof(float) * 5 * (size_t)n);
    float* y1 = x1 + n; float* x2 = y1 + n; float* y2 = x2 + n; float* area = y2 + n;
    om_key* order = (om_key*)malloc(sizeof(om_key) * (size_t)n);
    uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
    for (int32_t i = 0; i < n; ++i) {
        const float* d = dets + 5 * (size_t)i;
        float hw = d[2] / 2.0f, hh = d[3] / 2.0f;
        x1[i] = d[0] - hw; y1[i] = d[1] - hh;
        x2[i] = d[0] + hw; y2[i] = d[1] + hh;
        area[i] = (x2[i] - x1[i]) * (y2[i] - y1[i]);
        order[i].s = d[4]; order[i].i = i;
    }
    qsort(order, (size_t)n, sizeof(om_key), cmp_desc);
    for (int32_t a = 0; a < n; ++a) {
        int32_t i = order[a].i;
        if (dead[i]) continue;
        for (int32_t b = a + 1; b < n; ++b) {
            int32_t j = order[b].i;
            if (dead[j]) continue;
            float xx1 = x1[i] > x1[j] ? x1[i] : x1[j];
            float yy1 = y1[i] > y1[j] ? y1[i] : y1[j];
            float xx2 = x2[i] < x2[j] ? x2[i] : x2[j];
            float yy2 = y2[i] < y2[j] ? y2[i] : y2[j];
            float w = xx2 - xx1; if (w < 0.0f) w = 0.0f;
            float h = yy2 - yy1; if (h < 0.0f) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (area[i] + area[j] - inter);
            if (ovr >= threshold) dead[j] = 1;
        }
    }
    int32_t k = 0;
    for (int32_t i = 0; i < n; ++i) if (!dead[i]) keep[k++] = i;
    free(x1); free(order); free(dead);
    return k;
}

/*
 * The reference's OTHER native variant, eval/src/nms_kernel.cu (what it runs on CUDA tensors):
 *   - IoU from centre-format boxes: corners cx -/+ w/2, areas Sa = w * h directly            (:13-23)
 *   - boxes visited in score-descending order (scores.sort(0, descending=true), :88-90)
 *   - a later box is suppressed by a kept earlier one iff IoU > threshold (strict)           (:58)
 *   - result = kept boxes in that score-descending order (order_t.index(keep), :136-139)
 * The file cannot be compiled against torch >= 1.11 (THC headers), so unlike om_oracle_nms this restatement is NOT pinned against a
 * binary of the original; single-rounded fp32 like the rest (the original's FMA contraction is unknown).
 */
int32_t om_oracle_nms_cuda(const float* dets, int32_t n, float threshold, int64_t* keep) {
    if (n <= 0) return 0;
    om_key* order = (om_key*)malloc(sizeof(om_key) * (size_t)n);
    uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
    for (int32_t i = 0; i < n; ++i) { order[i].s = dets[5 * (size_t)i + 4]; order[i].i = i; }
    qsort(order, (size_t)n, sizeof(om_key), cmp_desc);
    int32_t k = 0;
    for (int32_t a = 0; a < n; ++a) {
        int32_t i = order[a].i;
        if (dead[i]) continue;
        keep[k++] = i;
        const float* p = dets + 5 * (size_t)i;
        for (int32_t b = a + 1; b < n; ++b) {
            int32_t j = order[b].i;
            if (dead[j]) continue;
            const float* q = dets + 5 * (size_t)j;
            float l1 = p[0] - p[2] / 2, l2 = q[0] - q[2] / 2, r1 = p[0] + p[2] / 2, r2 = q[0] + q[2] / 2;
            float t1 = p[1] - p[3] / 2, t2 = q[1] - q[3] / 2, b1 = p[1] + p[3] / 2, b2 = q[1] + q[3] / 2;
            float left = l1 > l2 ? l1 : l2, right = r1 < r2 ? r1 : r2, top = t1 > t2 ? t1 : t2, bottom = b1 < b2 ? b1 : b2;
            float w = right - left; if (w < 0.f) w = 0.f;
            float h = bottom - top; if (h < 0.f) h = 0.f;
            float inter = w * h;
            float sa = p[2] * p[3], sb = q[2] * q[3];
            if (inter / (sa + sb - inter) > threshold) dead[j] = 1;
        }
    }
    free(order); free(dead);
    return k;
}
