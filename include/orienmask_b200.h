/*
 * orienmask_b200 -- C ABI of the B200-native OrienMask inference hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer named "device" is a CUDA device
 * pointer; every function enqueues work on `stream` (a cudaStream_t passed as void*), performs no
 * allocation and no host synchronisation, and returns OM_OK or a negative error code whose text is
 * available from om_last_error().
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository root, duwt/OrienMask):
 *
 *   om_nms                 eval/src/nms_cuda.cpp:8-17  (pybind `nms(dets, threshold)`), and
 *                          eval/src/nms_cpu.cpp:65-75  whose semantics (>=, ascending indices) it keeps
 *   om_decode_select       eval/orienmask_yolo_postprocess.py:75-114,126-139 (get_boxes, confidence
 *                          filter, pre-NMS top-k) for the whole batch at once
 *   om_batched_nms         eval/function.py:77-103 (batched_nms) + eval/orienmask_yolo_postprocess.py:146-154
 *   om_mask_assemble       eval/orienmask_yolo_postprocess.py:69-72,99,141-144,156-164 (bilinear x4,
 *                          get_orien_grid, per-instance orientation thresholding)
 *   om_preprocess          data/transform.py:444-510 (FastCOCOTransform: permute + Resize + Normalize) and infer.py:21-32 (pad)
 *   om_mask_rle            eval/coco_eval.py:108-127,191-205 (_recover_shape_segm + maskUtils.encode of every instance)
 *   om_mask_areas, om_mask_blend  utils/visualizer.py:46-100,122-127 (resized soft masks, area sort key, alpha blend)
 *   om_stem_conv, om_conv_*  model/base.py:104-137 (ConvBNRelu, BN folded), model/backbone/darknet.py:6-15
 *                          (residual add), model/base.py:95-101 + torch.cat in
 *                          model/orienmask_yolo_fpnplus.py:78-79,85-86 (nearest upsample + concat, done
 *                          algebraically in the consumer's epilogue)
 */
#ifndef ORIENMASK_B200_H_
#define ORIENMASK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OM_OK 0
#define OM_ERR_INVALID (-1)      /* bad argument / unsupported shape */
#define OM_ERR_CUDA (-2)         /* a CUDA runtime / driver call failed */
#define OM_ERR_UNSUPPORTED (-3)  /* valid request that this build cannot serve */

#define OM_MAX_SCALES 4
#define OM_MAX_ANCHORS 16

/* Version of this ABI (bumped on any signature change). */
int32_t om_abi_version(void);
/* Text of the last error raised on the calling thread. */
const char* om_last_error(void);
/* Number of kernel launches enqueued by this library on the calling thread since the last reset. */
int64_t om_launch_count(void);
void om_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------- */
/* Post-process: decode + select + NMS + mask assembly                                           */
/* ------------------------------------------------------------------------------------------- */

/* Constructor arguments of OrienMaskYOLOPostProcess (eval/orienmask_yolo_postprocess.py:9-36). */
typedef struct om_post_config {
    int32_t num_scales;                         /* len(grid_size), <= OM_MAX_SCALES               */
    int32_t num_classes;
    int32_t image_h, image_w;                   /* image_size                                     */
    int32_t grid_h[OM_MAX_SCALES];              /* grid_size[i][0]                                */
    int32_t grid_w[OM_MAX_SCALES];              /* grid_size[i][1]                                */
    int32_t anchors_per_scale[OM_MAX_SCALES];   /* len(anchor_mask[i]), <= 4                      */
    int32_t anchor_index[OM_MAX_SCALES][4];     /* anchor_mask                                    */
    int32_t total_anchors;                      /* len(anchors), <= OM_MAX_ANCHORS                */
    float anchor_w[OM_MAX_ANCHORS];             /* anchors[a][0], pixels                          */
    float anchor_h[OM_MAX_ANCHORS];             /* anchors[a][1], pixels                          */
    float conf_thresh;
    float nms_thresh;                           /* threshold bound into nms_func by the builder   */
    float orien_thresh;
    int32_t nms_pre;                            /* <= 1024                                        */
    int32_t nms_post;                           /* <= nms_pre                                     */
    int32_t nms_semantics;                      /* OM_NMS_CPU (default, 0) or OM_NMS_CUDA         */
} om_post_config;

/* The reference ships TWO native NMS variants that differ at exact ties with the threshold and in result order:
 *   OM_NMS_CPU   eval/src/nms_cpu.cpp:4-63   IoU >= threshold suppresses, areas (x2-x1)*(y2-y1) from the corners, survivors returned as
 *                ascending original indices -- what the reference runs on CPU tensors, and the variant this repo's oracle is pinned to
 *                (the compiled reference file itself, oracle/_ref);
 *   OM_NMS_CUDA  eval/src/nms_kernel.cu:13-23,58-62,136-139   IoU > threshold suppresses, areas w*h, survivors returned in
 *                score-descending order -- what the reference runs on CUDA tensors (eval/function.py:70-74,98-101).  That file does
 *                not build against torch >= 1.11 (THC), so this mode is checked against a restatement only (oracle/nms_oracle.c);
 *                pairs within an ulp of the threshold may round differently from a binary of the original (FMA contraction unknown). */
#define OM_NMS_CPU 0
#define OM_NMS_CUDA 1

/* Bytes of device scratch om_decode_select needs for `batch` images: 16 bytes per (prediction, class) pair and image
 * (candidate keys + flat indices, edge list) plus the radix state -- every pair may clear conf_thresh. */
int32_t om_post_workspace_bytes(const om_post_config* cfg, int32_t batch, size_t* bytes);

/*
 * Box/score decode of every (prediction, class), confidence filter and pre-NMS top-k, for the whole batch.
 *   bbox[s]            device, fp32 NCHW [batch, A_s*(5+C), grid_h[s], grid_w[s]] (model head layout);
 *                      image b starts at bbox[s] + b*bbox_batch_stride[s] (elements)
 *   workspace          device scratch of om_post_workspace_bytes() bytes
 * Outputs (device), rows beyond cand_count[b] are zero-filled:
 *   cand_count [batch]            int32   number of candidates kept (<= nms_pre)
 *   cand_det   [batch,nms_pre,5]  fp32    cx, cy, w, h (normalised), score
 *   cand_cls   [batch,nms_pre]    int32
 *   cand_pred  [batch,nms_pre]    int32   flat prediction index (scale-major, then anchor, y, x)
 * Order inside an image follows the reference: score-descending when more than nms_pre pairs clear
 * conf_thresh (ties: lower (prediction, class) first), otherwise (prediction, class) row-major.
 */
int32_t om_decode_select(const om_post_config* cfg, const float* const* bbox, const int64_t* bbox_batch_stride,
                         int32_t batch, void* workspace, int32_t* cand_count, float* cand_det,
                         int32_t* cand_cls, int32_t* cand_pred, void* stream);

/*
 * Class-wise greedy NMS (centres shifted by cls*2.0 in fp32, IoU >= nms_thresh suppresses) followed by
 * the post-NMS top-k.  Inputs are om_decode_select's outputs.  Outputs (device), zero-filled past det_count[b]:
 *   det_count  [batch]             int32
 *   det        [batch,nms_post,5]  fp32   un-shifted cx, cy, w, h, score
 *   det_cls    [batch,nms_post]    int64
 *   det_anchor [batch,nms_post]    int32  global anchor index of the prediction (selects the orientation map)
 *   det_keep   [batch,nms_post]    int32  index into the candidate list (the reference's `keep`)
 *   records    [batch,nms_post*6+1] fp32  optional (may be NULL): cx, cy, w, h, score, cls per slot, then the count -- the
 *                                         fixed-size row a rank contributes to the multi-GPU all-gather, written by the same
 *                                         kernel so that the collective's source is this buffer
 */
int32_t om_batched_nms(const om_post_config* cfg, const int32_t* cand_count, const float* cand_det,
                       const int32_t* cand_cls, const int32_t* cand_pred, int32_t batch, int32_t* det_count,
                       float* det, int64_t* det_cls, int32_t* det_anchor, int32_t* det_keep, float* records, void* stream);

/*
 * Instance masks: mask[b,k,y,x] = |pix_x - xc| < t*w*nW  &&  |pix_y - yc| < t*h*nH on the x4 bilinear
 * up-sampling of the low-resolution orientation maps (never materialised).
 *   orien[s]   device, fp32 [batch, 2*A_s, image_h/4, image_w/4]; image b at orien[s] + b*orien_batch_stride[s]
 *   mask       device, uint8 (0/1) [batch, nms_post, image_h, image_w]; only rows k < det_count[b] are written
 */
int32_t om_mask_assemble(const om_post_config* cfg, const float* const* orien, const int64_t* orien_batch_stride,
                         const int32_t* det_count, const float* det, const int32_t* det_anchor, int32_t batch,
                         uint8_t* mask, void* stream);

/*
 * Stand-alone greedy NMS over one box list (drop-in for the reference's native `nms(dets, threshold)`).
 *   dets [n,5] device fp32 (cx, cy, w, h, score), n <= 1024
 *   keep [n]   device int64, out: surviving indices in ascending order (nms_cpu.cpp:62); *keep_count device int32
 */
int32_t om_nms(const float* dets, int32_t n, float threshold, int64_t* keep, int32_t* keep_count, void* stream);
/* Same with the variant chosen: OM_NMS_CPU (= om_nms) or OM_NMS_CUDA (keep is then in score-descending order). */
int32_t om_nms_ex(const float* dets, int32_t n, float threshold, int32_t semantics, int64_t* keep, int32_t* keep_count, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Convolution engine                                                                            */
/* ------------------------------------------------------------------------------------------- */

/*
 * Activation layout ("padded-row NHWC"): a feature map of B images, H x W x C, is stored as
 * [B * rows_per_image, W, C] with rows_per_image >= H + 1; rows H..rows_per_image-1 of every image
 * are zero and are never written, so a 3x3 window that leaves an image vertically reads zeros
 * without any bounds logic (horizontal padding comes from TMA out-of-bounds fill).  Stride-2
 * convolutions need rows_per_image(in) == 2 * rows_per_image(out).
 */
#define OM_PREC_F32 0   /* parity engine: fp32 storage, FFMA                                     */
#define OM_PREC_F16 1   /* production engine: fp16 storage, tcgen05 tensor cores, fp32 accumulate */
#define OM_PREC_SPLIT 2 /* parity engine on the tensor cores: every activation and weight is an fp16 pair hi + lo
                           (hi = fp16(v), lo = fp16(v - hi): ~22 significant bits), D = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi as
                           three tcgen05 MMA passes into one fp32 TMEM accumulator (corrections first: the tensor core
                           truncates its running sum).  Activations: [.., W, 2*C] halves (hi | lo along
                           the channel axis); weights [k*k][cout_pad][2*cin] halves (hi | lo), pre-scaled by a power of two so
                           that W_lo stays a normal fp16 number (undone by acc_scale)                                  */

#define OM_OUT_ACT 0      /* padded-row NHWC activation in the engine precision                    */
#define OM_OUT_PARTIAL 1  /* padded-row NHWC fp32 pre-activation partial sum (no bias, no act)     */
#define OM_OUT_NCHW 2     /* dense fp32 NCHW [B, cout, H, W] (model head output), bias, no act     */

typedef struct om_conv_desc {
    int32_t precision;             /* OM_PREC_*                                                   */
    int32_t batch;
    int32_t in_h, in_w, in_rows;   /* input map and its rows_per_image                            */
    int32_t out_h, out_w, out_rows;
    int32_t cin, cout;             /* cin % 32 == 0 (fp16 engine)                                 */
    int32_t cout_stride;           /* channel pitch of the output (>= cout), OM_OUT_ACT/PARTIAL   */
    int32_t ksize;                 /* 1 or 3 (padding ksize/2)                                    */
    int32_t stride;                /* 1 or 2                                                      */
    int32_t leaky;                 /* LeakyReLU(0.1) after bias (+upadd)                          */
    int32_t out_kind;              /* OM_OUT_*                                                    */
    const void* input;             /* device, engine precision                                    */
    const void* weights;           /* device: F16 -> [k*k][cout_pad][cin] half, cout_pad = cout rounded up
                                      to 32;  F32 -> [k*k][cin][cout rounded up to 4] float;
                                      SPLIT -> [k*k][cout_pad][2*cin] half (hi | lo)                */
    const float* bias;             /* device [cout] or NULL                                       */
    const void* residual;          /* device, same layout/precision as an OM_OUT_ACT output, added AFTER the
                                      activation (darknet.py:15), or NULL                         */
    const float* upadd;            /* device fp32 padded-row NHWC [B*up_rows, out_w/2, cout] added BEFORE bias and
                                      activation at (y/2, x/2) (nearest x2 up-sampling), or NULL  */
    int32_t up_rows;
    void* output;
    /* "Parity-split" (space-to-depth) activation layout, used between a layer and a stride-2 consumer that is its only
     * reader: the four (row parity, column parity) sub-images are stored one after the other, each a dense padded-row
     * NHWC tensor [B*rows/2, W/2, C]; plane index = 2*(Y&1) + (x&1), Y the row in the [B*rows] space (rows even).
     * Every tap of the stride-2 convolution is then a dense TMA box instead of a strided gather. */
    int32_t in_s2d;                /* the input (of a stride-2 layer) is parity-split                  */
    int32_t out_s2d;               /* write the OM_OUT_ACT output parity-split (no residual in place)  */
    float acc_scale;               /* OM_PREC_SPLIT: the accumulator is multiplied by this (a power of two, the inverse of the
                                      weight pre-scale) before up-add, bias and activation; 0 means 1  */
} om_conv_desc;

typedef struct om_conv om_conv;

/* Validates the descriptor and precomputes launch geometry and TMA descriptors. */
int32_t om_conv_create(const om_conv_desc* desc, om_conv** out);
int32_t om_conv_run(const om_conv* conv, void* stream);
/* Same launch with the result redirected to `output` (same layout and size as desc.output; layers without a residual): the model
 * wrapper gives every call freshly allocated head tensors, as the reference's forward does. */
int32_t om_conv_run_to(const om_conv* conv, void* output, void* stream);
void om_conv_destroy(om_conv* conv);

/* Debug / tooling (not needed to run the path): the planner's decisions for one layer of the tcgen05 engines --
 * info[24] = halo, flat, halo_s2, b_resident, tw, th, block_n, tiles_n, stages, n_sub, h_stages, acc_stages, has_res (1 fp16 residual staged
 * by TMA, 2 staged up-add), res_direct, smem bytes, grid, CTA-pair tiles, taps, K chunks, BK, TMEM columns, cout, out_h, out_w. */
int32_t om_debug_conv_plan_info(const om_conv* conv, int32_t* info24);
/* Debug: device buffer of >= 16 uint64 that cluster 0 of the following conv launches fills with %globaltimer stamps (NULL turns it off). */
int32_t om_debug_conv_timeline(void* device_u64x16);
/* Launch trace: arm (records != NULL: a device buffer of capacity x 4 uint64, pre-filled with {~0, ~0, 0, 0} per record) or disarm.
 * While armed every conv-engine launch takes the next record and stamps %globaltimer into it: [0] first CTA start, [1] dependencies
 * resolved (griddepcontrol.wait returned), [2] last CTA end, [3] last CTA start.  Returns the records handed out since the previous
 * call.  tools/timeline.py: the in-situ timeline of a pipelined forward (what ncu's serialised launch list cannot show). */
int32_t om_debug_trace(void* records, int32_t capacity);
/* Phase log of the fused stem kernel: stream 0 of CTA 0 stamps %globaltimer at the phase boundaries of its first 16 tiles into
 * [16][8] uint64 (tile start, patch arrived, im2col rows done, MMA 1 done, epilogue 1 done, MMA 2 done, epilogue 2 done). NULL disarms. */
int32_t om_debug_phase_log(void* device_u64x128);

/*
 * First layer (3 -> cout, 3x3, stride 1, BN folded, LeakyReLU) straight from the caller's image.
 *   image    device fp32 NCHW [batch,3,h,w]
 *   weights  device fp32 [27][cout] (tap-major: (ky*3+kx)*3+ci), bias fp32 [cout]; cout == 32
 *   output   padded-row NHWC [batch*rows, w, cout] in `precision` (parity-split when out_s2d, see om_conv_desc);
 *            OM_PREC_SPLIT: [batch*rows, w, 2*cout] halves (hi | lo), computed in fp32 on the CUDA cores
 */
int32_t om_stem_conv(int32_t precision, const float* image, const float* weights, const float* bias, void* output,
                     int32_t batch, int32_t h, int32_t w, int32_t rows, int32_t cout, int32_t out_s2d, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Whole network: OrienMaskYOLOFPNPlus.forward / OrienMaskYOLO.forward behind one call            */
/* (model/orienmask_yolo_fpnplus.py:9-90, model/orienmask_yolo.py:8-86, model/backbone/darknet.py:18-54) */
/* ------------------------------------------------------------------------------------------- */

typedef struct om_engine_config {
    int32_t precision;             /* OM_PREC_*                                                                  */
    int32_t batch, height, width;  /* input [batch, 3, height, width]; height, width multiples of 32              */
    int32_t num_anchors;           /* constructor arguments of the reference model (config/base.py:99-106)        */
    int32_t num_classes;
    int32_t plus;                  /* 1: OrienMaskYOLOFPNPlus (524-key state dict), 0: OrienMaskYOLO (506 keys)   */
} om_engine_config;

/* One entry of the reference state dict: its key (e.g. "backbone.conv2.1.conv.0.conv_block.1.running_var"), the tensor as fp32
 * on the device in the reference's own layout (conv weights OIHW), and its element count (checked against the architecture). */
typedef struct om_tensor {
    const char* name;
    const float* data;
    int64_t numel;
} om_tensor;

typedef struct om_engine om_engine;

/* Bytes of device workspace an engine of this configuration needs: activation buffers, folded + packed weights, biases. */
int32_t om_engine_workspace_bytes(const om_engine_config* cfg, size_t* bytes);
/*
 * Builds the launch schedule of the forward pass: folds BatchNorm into the convolution weights and packs them for `precision` on
 * the device (kernels enqueued on `stream`; the state-dict tensors may be released once that work has completed), lays every
 * activation buffer out inside `workspace` (zero-filled here; at least om_engine_workspace_bytes() bytes, caller-owned, must
 * outlive the engine) and plans every layer (TMA descriptors, tiles).  The one call of the library that synchronises `stream`
 * (OM_PREC_SPLIT reads the per-layer weight scale back); om_forward never does.
 */
int32_t om_engine_create(const om_engine_config* cfg, const om_tensor* weights, int32_t n_weights, void* workspace,
                         size_t workspace_bytes, void* stream, om_engine** out);
/*
 * forward(x) -> ((bbox32, orien32), (bbox16, orien16), (bbox8, orien8)), model/orienmask_yolo_fpnplus.py:74-90.
 *   image    device fp32 NCHW [batch, 3, height, width]; 16-byte aligned for OM_PREC_F16 (its first launch stages the image by TMA)
 *   bbox[3]  device fp32 NCHW [batch, A*(5+C), height/s, width/s] for s = 32, 16, 8
 *   orien    device fp32 NCHW [batch, 6*A, height/4, width/4]: channels [0,2A) belong to stride 32, [2A,4A) to 16, [4A,6A) to 8
 *            (the reference's torch.split(oriens, 2A, dim=1), :88)
 * Enqueues every launch of the schedule on `stream`; no allocation, no synchronisation.
 */
int32_t om_forward(const om_engine* engine, const float* image, float* const* bbox, float* orien, void* stream);
void om_engine_destroy(om_engine* engine);

/* Introspection for tools (per-layer timing, plan tables): the schedule is a list of launches in execution order. */
typedef struct om_layer_info {
    char name[64];                 /* state-dict prefix of the layer ("neck4.0[64:128]" for a concat-split partial)          */
    char shape[96];                /* e.g. "3x3 s1 128->256 @136x136 +res"                                                 */
    double flops, bytes;           /* algorithmic: 2*MAC; activations in + out (+ addends) + weights                       */
    int32_t head_slot;             /* -1, or 0..2 (bbox 32/16/8), 3 (orientation)                                          */
    int32_t is_stem;
    om_conv_desc desc;
    const om_conv* conv;           /* NULL for the stem                                                                    */
} om_layer_info;
int32_t om_engine_layer_count(const om_engine* engine);
int32_t om_engine_layer_info(const om_engine* engine, int32_t index, om_layer_info* info);
/* One launch of the schedule (same arguments as om_forward; only the stem reads `image`, only head layers write outputs). */
int32_t om_engine_run_layer(const om_engine* engine, int32_t index, const float* image, float* const* bbox, float* orien, void* stream);
/* A sub-schedule in one call: the launches `indices[0..n)` back to back on `stream` (tools/insitu.py times the forward with and
 * without a group of layers: what the group costs INSIDE the pipelined forward, which ncu's serialised launch list cannot say). */
int32_t om_engine_run_layers(const om_engine* engine, const int32_t* indices, int32_t n, const float* image, float* const* bbox,
                             float* orien, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Pre-process (the caller side of the path: infer.py:147-151)                                   */
/* ------------------------------------------------------------------------------------------- */

#define OM_SRC_U8 0    /* uint8 HWC (what cv2.imread yields)                                     */
#define OM_SRC_F32 1   /* float32 HWC (what infer.py:148 hands to the transform)                 */

/* FastCOCOTransform(pipeline=[Resize | ShortEdgeResize, Normalize]) followed by pad(): the resized image
 * (resize_h x resize_w, bilinear, align_corners=False) is normalised per channel and placed at
 * (pad_top, pad_left) inside an out_h x out_w canvas filled with pad_value. */
typedef struct om_prep_config {
    int32_t src_h, src_w;          /* source image, layout [batch, src_h, src_w, 3]                  */
    int32_t src_dtype;             /* OM_SRC_*                                                       */
    int32_t resize_h, resize_w;    /* Resize.size / ShortEdgeResize result (data/transform.py:463-494) */
    int32_t pad_top, pad_left;     /* infer.py:25 (centred: (new - old) // 2)                        */
    int32_t out_h, out_w;          /* canvas (multiples of size_divisor)                             */
    float mean[3], std[3];         /* Normalize (data/transform.py:496-507): (v - mean) / std        */
    float pad_value;               /* infer.py:21, applied after normalisation                       */
} om_prep_config;

/*
 *   src   device, uint8 or fp32 [batch, src_h, src_w, 3]; image b at src + b*src_batch_stride (elements)
 *   out   device, fp32 NCHW [batch, 3, out_h, out_w], contiguous (the model input)
 */
int32_t om_preprocess(const om_prep_config* cfg, const void* src, int64_t src_batch_stride, int32_t batch, float* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Detections -> COCO segmentation format (the consumer side of the path: trainer/tester.py:46-50) */
/* ------------------------------------------------------------------------------------------- */

/* One image of a batch: where its instance masks live and how COCOMetrics._recover_shape_segm
 * (eval/coco_eval.py:191-205) maps them back to the original image. */
typedef struct om_rle_image {
    const uint8_t* mask;           /* device, uint8 0/1 [count, mask_h, mask_w] (the post-process output)       */
    int32_t count;                 /* instances of this image (<= max_inst)                                   */
    int32_t mask_h, mask_w;        /* network-input size                                                      */
    int32_t top, left;             /* first row / column kept after removing `collate_pad` and `pad`          */
    int32_t crop_h, crop_w;        /* size of the kept window                                                 */
    int32_t out_h, out_w;          /* sample_info['height'], ['width']: bilinear resize target, then round()  */
    int32_t hflip, vflip;          /* torch.flip on the cropped mask before the resize                        */
} om_rle_image;

/*
 * For every instance k < images[b].count: crop, flip, bilinear-resize (align_corners=False) and round the mask,
 * run-length encode it in column-major order (pycocotools rleEncode: first count = leading zeros) and compress the
 * counts into the COCO string (rleToString).  The resized mask is never materialised.
 *   images      device [batch]
 *   max_out_h, max_mask_h, max_mask_w   maxima over the batch of out_h, mask_h, mask_w (size the per-CTA tables)
 * Outputs (device), instance index = b*max_inst + k:
 *   counts   [batch*max_inst, cap]      uint32 run lengths
 *   n_counts [batch*max_inst]           int32  number of runs; when > cap nothing else is valid for that instance
 *                                              and the caller retries with a larger cap
 *   str      [batch*max_inst, str_cap]  uint8  compressed counts (ASCII, no terminator)
 *   str_len  [batch*max_inst]           int32  bytes in str, or -1 when cap or str_cap was too small
 */
int32_t om_mask_rle(const om_rle_image* images, int32_t batch, int32_t max_inst, int32_t max_out_h, int32_t max_mask_h,
                    int32_t max_mask_w, int32_t cap, int32_t str_cap, uint32_t* counts, int32_t* n_counts, uint8_t* str, int32_t* str_len, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Visualiser mask blend (utils/visualizer.py:46-100, 122-127)                                   */
/* ------------------------------------------------------------------------------------------- */

typedef struct om_blend_config {
    int32_t mask_h, mask_w;        /* network-input size of the instance masks                                 */
    int32_t top, left;             /* pad_info: first row / column kept (utils/visualizer.py:123-124)          */
    int32_t crop_h, crop_w;        /* size of the kept window                                                  */
    int32_t out_h, out_w;          /* the image being drawn on: bilinear resize target (not rounded)           */
    float alpha;                   /* InferenceVisualizer(alpha=...)                                           */
} om_blend_config;

/* areas[i] = sum over the image of the resized soft mask i (utils/visualizer.py:69 sorts by it).
 *   mask    device uint8 0/1 [k, mask_h, mask_w];  scratch  device double [k];  areas  device fp32 [k] */
int32_t om_mask_areas(const om_blend_config* cfg, const uint8_t* mask, int32_t k, double* scratch, float* areas, void* stream);

/* plot_all_mask (utils/visualizer.py:95-100) in place on `image` (device fp32 [out_h, out_w, 3]):
 *   order   device int32 [k]   instance indices in drawing order (ascending area)
 *   colors  device fp32 [k,3]  colour of instance i (not of drawing position) */
int32_t om_mask_blend(const om_blend_config* cfg, const uint8_t* mask, int32_t k, const int32_t* order, const float* colors,
                      float* image, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORIENMASK_B200_H_ */
