"""Shared test helpers: the north-star post-process constants and seeded synthetic head tensors."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# config/base.py:12-16 (ANCHORS_YOLOV4) and :6 (ANCHORS_MASK) of the reference
ANCHORS = [[12, 16], [19, 36], [40, 28], [36, 75], [76, 55], [72, 146], [142, 110], [192, 243], [459, 401]]
ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]


def synthetic_heads(batch, height, width, seed, num_classes=80, dtype=torch.float32):
    """Seeded head tensors shaped like the model output (bbox logits ~N(bias, 1.9), orien ~N(0,1))."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in (32, 16, 8):
        nH, nW = height // s, width // s
        bbox = torch.randn(batch, 3, 5 + num_classes, nH, nW, generator=g) * 1.9
        bbox[:, :, 4] -= 4.0
        bbox[:, :, 5:] -= 2.0
        bbox[:, :, 2:4] *= 0.3
        orien = torch.randn(batch, 6, height // 4, width // 4, generator=g)
        out.append((bbox.view(batch, -1, nH, nW).contiguous().to(dtype), orien.to(dtype)))
    return tuple(out)


def post_config(height, width, conf_thresh=0.005):
    return dict(grid_size=[[height // s, width // s] for s in (32, 16, 8)], image_size=[height, width],
                anchors=ANCHORS, anchor_mask=ANCHOR_MASK, num_classes=80, conf_thresh=conf_thresh,
                nms_pre=400, nms_post=100, orien_thresh=0.3)


def unpack_masks(gold, prefix, b):
    shape = tuple(int(v) for v in gold['%s_maskshape_%d' % (prefix, b)])
    n = int(np.prod(shape))
    return np.unpackbits(gold['%s_maskbits_%d' % (prefix, b)])[:n].reshape(shape).astype(bool)


# ---- helpers for the convolution-engine tests (GPU) ---------------------------------------------
def to_padded(x, rows, dtype):
    """NCHW -> padded-row NHWC [B*rows, W, C] with zero rows after each image."""
    B, C, H, W = x.shape
    out = torch.zeros(B * rows, W, C, dtype=dtype, device=x.device)
    out.view(B, rows, W, C)[:, :H] = x.permute(0, 2, 3, 1).to(dtype)
    return out


def from_padded(t, B, H):
    rows = t.shape[0] // B
    return t.view(B, rows, t.shape[1], t.shape[2])[:, :H].permute(0, 3, 1, 2).float()


def pack_weights(w, precision):
    from orienmask_b200 import _lib
    return _lib.pack_conv_weights(w, precision)


def split_halves(t):
    """fp32 [..., C] -> fp16 [..., 2C]: hi | lo along the last axis (OM_PREC_SPLIT activations)."""
    hi = t.to(torch.float16)
    lo = (t.float() - hi.float()).to(torch.float16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def merge_halves(t):
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


def to_s2d(t):
    """padded-row NHWC [R, W, C] -> parity-split: plane 2*(row&1)+(col&1), each [R/2, W/2, C] (include/orienmask_b200.h)."""
    return torch.stack([t[py::2, px::2] for py in (0, 1) for px in (0, 1)]).contiguous().view(t.shape)


def from_s2d(t):
    R, W, C = t.shape
    planes = t.view(4, R // 2, W // 2, C)
    out = torch.empty_like(t)
    for py in (0, 1):
        for px in (0, 1):
            out[py::2, px::2] = planes[2 * py + px]
    return out


def run_engine_conv(x, w, bias, stride=1, leaky=True, kind=0, residual=None, upadd=None, precision=1, extra_rows=1,
                    in_s2d=False, out_s2d=False):
    """One om_conv_* call on NCHW torch inputs; returns NCHW fp32 (pad rows checked to stay zero)."""
    from orienmask_b200 import _lib
    lib = _lib.lib()
    B, cin, H, W = x.shape
    cout, _, k, _ = w.shape
    Ho, Wo = H // stride, W // stride
    rows_o = Ho + extra_rows
    rows_i = rows_o * stride
    split = precision == _lib.PREC_SPLIT
    adt = torch.float32 if precision == _lib.PREC_F32 else torch.float16
    xin = split_halves(to_padded(x, rows_i, torch.float32)) if split else to_padded(x, rows_i, adt)
    if in_s2d:
        xin = to_s2d(xin)
    wp, acc_scale = pack_weights(w, precision)
    d = _lib.ConvDesc()
    d.precision, d.batch = precision, B
    d.in_h, d.in_w, d.in_rows, d.out_h, d.out_w, d.out_rows = H, W, rows_i, Ho, Wo, rows_o
    d.cin, d.cout, d.cout_stride, d.ksize, d.stride, d.leaky, d.out_kind = cin, cout, cout, k, stride, int(leaky), kind
    d.input, d.weights = xin.data_ptr(), wp.data_ptr()
    d.in_s2d, d.out_s2d = int(in_s2d), int(out_s2d)
    d.acc_scale = acc_scale
    keep = [xin, wp]
    if bias is not None:
        b = bias.float().contiguous()
        keep.append(b)
        d.bias = b.data_ptr()
    if kind == _lib.OUT_NCHW:
        out = torch.full((B, cout, Ho, Wo), float('nan'), dtype=torch.float32, device=x.device)
    elif kind == _lib.OUT_PARTIAL:
        out = torch.zeros(B * rows_o, Wo, cout, dtype=torch.float32, device=x.device)
    else:
        out = torch.zeros(B * rows_o, Wo, cout * (2 if split else 1), dtype=adt, device=x.device)
    d.output = out.data_ptr()
    if residual is not None:
        r = split_halves(to_padded(residual, rows_o, torch.float32)) if split else to_padded(residual, rows_o, adt)
        keep.append(r)
        d.residual = r.data_ptr()
    if upadd is not None:
        up_rows = Ho // 2 + 1
        u = to_padded(upadd, up_rows, torch.float32)
        keep.append(u)
        d.upadd, d.up_rows = u.data_ptr(), up_rows
    h = _lib.c_vp()
    _lib.check(lib.om_conv_create(d, h), 'om_conv_create')
    try:
        _lib.check(lib.om_conv_run(h, _lib.stream_ptr()), 'om_conv_run')
        torch.cuda.synchronize()
    finally:
        lib.om_conv_destroy(h)
    if kind == _lib.OUT_NCHW:
        return out
    if out_s2d:
        out = from_s2d(out)
    if split and kind == _lib.OUT_ACT:
        out = merge_halves(out)
    pad = out.view(B, rows_o, Wo, cout)[:, Ho:]
    assert float(pad.abs().max()) == 0.0, 'padding rows were written'
    return from_padded(out, B, Ho)


def torch_conv_ref(x, w, bias, stride=1, leaky=True, kind=0, residual=None, upadd=None, quantize=False):
    """fp32 torch reference of the same fused op (optionally on fp16-rounded operands)."""
    import torch.nn.functional as F
    if quantize:
        x, w = x.half().float(), w.half().float()
        if residual is not None:
            residual = residual.half().float()
    y = F.conv2d(x.double(), w.double(), None, stride=stride, padding=w.shape[-1] // 2)
    if upadd is not None:
        y = y + F.interpolate(upadd.double(), scale_factor=2, mode='nearest')
    if bias is not None and kind != 1:
        y = y + bias.double().view(1, -1, 1, 1)
    if leaky:
        y = F.leaky_relu(y, 0.1)
    if residual is not None:
        y = y + residual.double()
    return y.float()


# ---- trained-like head tensors: piecewise orientation fields with sharp instance boundaries --------
def trained_like_heads(batch, height, width, seed, num_classes=80, per_image=10):
    """Head tensors shaped like the model output whose ORIENTATION maps look like a trained OrienMask's: inside an instance every pixel's
    offset points at the instance centre (eval/orienmask_yolo_postprocess.py:141-166: pix = orien * grid_anchor / 2 + base lands on the
    centre), outside it is zero -- so a mask boundary is a discontinuity of the field, not a shallow level set of a smooth random one.
    Instances are non-overlapping ellipses laid out on a jittered grid; each gets the anchor closest to its size, a confident
    objectness / class logit at its centre cell and the box regression that reproduces its box.  Returns (heads, instances)."""
    g = np.random.default_rng(seed)
    anchors = np.asarray(ANCHORS, dtype=np.float64)
    heads = []
    for s in (32, 16, 8):
        nH, nW = height // s, width // s
        bbox = np.full((batch, 3, 5 + num_classes, nH, nW), -6.0, dtype=np.float32)
        bbox[:, :, :4] = 0.0
        bbox[:, :, 4] = -9.0
        heads.append([bbox, np.zeros((batch, 6, height // 4, width // 4), dtype=np.float32)])
    cols = int(np.ceil(np.sqrt(per_image)))
    rows = int(np.ceil(per_image / cols))
    yy, xx = np.mgrid[0:height // 4, 0:width // 4]
    py, px = 4.0 * yy + 1.5, 4.0 * xx + 1.5                      # full-resolution position of every low-resolution orientation sample
    instances = []
    for b in range(batch):
        used = set()
        for k in range(per_image):
            cell_h, cell_w = height / rows, width / cols
            r, c = divmod(k, cols)
            w = g.uniform(0.35, 0.8) * cell_w
            h = g.uniform(0.35, 0.8) * cell_h
            cx = (c + 0.5) * cell_w + g.uniform(-0.08, 0.08) * cell_w
            cy = (r + 0.5) * cell_h + g.uniform(-0.08, 0.08) * cell_h
            a = int(np.argmin(np.abs(np.log(anchors[:, 0] / w)) + np.abs(np.log(anchors[:, 1] / h))))
            si = [i for i, m in enumerate(ANCHOR_MASK) if a in m][0]
            aj = ANCHOR_MASK[si].index(a)
            if (si, aj, r, c) in used:
                continue
            used.add((si, aj, r, c))
            s = (32, 16, 8)[si]
            nH, nW = height // s, width // s
            gx, gy = cx / s, cy / s
            ix, iy = int(gx), int(gy)
            cls = int(g.integers(0, num_classes))
            bbox = heads[si][0]
            fx, fy = min(max(gx - ix, 0.02), 0.98), min(max(gy - iy, 0.02), 0.98)
            bbox[b, aj, 0, iy, ix] = np.log(fx / (1 - fx))
            bbox[b, aj, 1, iy, ix] = np.log(fy / (1 - fy))
            bbox[b, aj, 2, iy, ix] = np.log(w / anchors[a, 0])
            bbox[b, aj, 3, iy, ix] = np.log(h / anchors[a, 1])
            bbox[b, aj, 4, iy, ix] = g.uniform(3.0, 7.0)
            bbox[b, aj, 5 + cls, iy, ix] = g.uniform(3.0, 7.0)
            inside = ((px - cx) / (w / 2)) ** 2 + ((py - cy) / (h / 2)) ** 2 < 1.0
            ga_x, ga_y = anchors[a, 0] / width * nW, anchors[a, 1] / height * nH          # grid_anchors (:21-27)
            orien = heads[si][1]
            orien[b, 2 * aj][inside] = ((cx - px[inside]) / width * nW) * 2.0 / ga_x
            orien[b, 2 * aj + 1][inside] = ((cy - py[inside]) / height * nH) * 2.0 / ga_y
            instances.append(dict(image=b, cx=cx, cy=cy, w=w, h=h, anchor=a, cls=cls))
    out = tuple((torch.from_numpy(bb.reshape(batch, -1, bb.shape[-2], bb.shape[-1]).copy()), torch.from_numpy(oo)) for bb, oo in heads)
    return out, instances


# ---- end-to-end agreement with explicit exceptions (SURVEY §8d iii) -------------------------------
def _iou_centre(a, b):
    """IoU of centre-format boxes a [4], b [n,4] (float64; only used to measure margins)."""
    ax1, ay1, ax2, ay2 = a[0] - a[2] / 2, a[1] - a[3] / 2, a[0] + a[2] / 2, a[1] + a[3] / 2
    bx1, by1, bx2, by2 = b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2
    iw = np.clip(np.minimum(ax2, bx2) - np.maximum(ax1, bx1), 0, None)
    ih = np.clip(np.minimum(ay2, by2) - np.maximum(ay1, by1), 0, None)
    inter = iw * ih
    return inter / (a[2] * a[3] + b[:, 2] * b[:, 3] - inter)


def e2e_agreement(ref, padded, b, nms_pre=400, nms_post=100, nms_thr=0.5, score_noise=1e-3, iou_noise=5e-3):
    """Detections of image `b` of an engine result (PaddedDetections) against the oracle's result `ref` for the same image
    (PostProcessOracle.image: forward oracle heads -> post-process oracle), matched by (prediction index, class).

    Returns a report: matched pairs with their worst box / score error and mask IoU, and the list of EXCEPTIONS -- (prediction,
    class) pairs kept by only one side -- each with the margin that explains it, measured on the oracle's own scores and boxes:
    the distance of its score to the pre-NMS top-k cut, to the post-NMS top-k cut, or of the deciding NMS pair's IoU to the
    threshold (directly, or through another exception of the same class it overlaps: a cascade).  An exception whose margins all
    exceed the forward noise (`score_noise`, `iou_noise`) is `explained: False` -- a real disagreement."""
    k = int(padded.count[b].item())
    keep = padded.keep[b, :k].long()
    g_pred = padded.candidates['pred'][b].long()[keep].cpu().numpy()
    g_cls = padded.cls[b, :k].cpu().numpy()
    g_box = padded.det[b, :k].cpu().numpy()
    g_mask = padded.mask[b, :k].cpu().numpy().astype(bool)
    r_key = {(int(p), int(c)): i for i, (p, c) in enumerate(zip(ref['pred'], ref['cls']))}
    g_key = {(int(p), int(c)): i for i, (p, c) in enumerate(zip(g_pred, g_cls))}
    rep = {'reference_detections': len(r_key), 'engine_detections': len(g_key), 'matched': 0, 'max_box_err': 0.0, 'max_score_err': 0.0,
           'min_mask_iou': 1.0, 'mean_mask_iou': 1.0, 'masks_below_0p999': 0, 'aggregate_mask_iou': 1.0, 'max_differing_pixels': 0,
           'masks_off': 0, 'exceptions': []}
    ious, diffs, inter_sum, union_sum = [], [], 0, 0
    for key, gi in g_key.items():
        ri = r_key.get(key)
        if ri is None:
            continue
        rep['matched'] += 1
        rep['max_box_err'] = max(rep['max_box_err'], float(np.abs(ref['bbox'][ri, :4] - g_box[gi, :4]).max()))
        rep['max_score_err'] = max(rep['max_score_err'], float(abs(ref['bbox'][ri, 4] - g_box[gi, 4])))
        u = int((ref['mask'][ri] | g_mask[gi]).sum())
        it = int((ref['mask'][ri] & g_mask[gi]).sum())
        ious.append(it / u if u else 1.0)
        diffs.append(u - it)
        inter_sum, union_sum = inter_sum + it, union_sum + u
    if ious:
        rep['min_mask_iou'], rep['mean_mask_iou'] = min(ious), float(np.mean(ious))
        rep['masks_below_0p999'] = int(sum(i < 0.999 for i in ious))
        # Mask gate (north star: IoU >= 0.999).  A per-mask minimum is not attainable by ANY re-implementation: the reference's own fp32
        # forward against itself in fp64 (head rel-L2 1e-6) flips single pixels and reads min IoU 0.9962 on a 260-pixel mask
        # (profiles/r02_reference_self_noise.json).  So: the IoU over all matched instances of the image (sum of intersections / sum of
        # unions) must be >= 0.999, and a mask below 0.999 counts as OFF only if it also differs in more than 2 pixels.
        rep['aggregate_mask_iou'] = inter_sum / union_sum if union_sum else 1.0
        rep['max_differing_pixels'] = int(max(diffs))
        rep['masks_off'] = int(sum(i < 0.999 and d > 2 for i, d in zip(ious, diffs)))
    # ---- exceptions and their margins (oracle-side quantities only) ----
    cand = ref['cand']
    conf, coord_all = ref['conf'], ref['coord_all']
    C = conf.shape[1]
    cut_pre = float(cand['score'].min()) if len(cand['score']) >= nms_pre else None          # score of the last candidate taken
    kept_scores = np.sort(ref['bbox'][:, 4])[::-1]
    cut_post = float(kept_scores[nms_post - 1]) if len(kept_scores) >= nms_post else None
    only = [(key, 'reference') for key in r_key if key not in g_key] + [(key, 'engine') for key in g_key if key not in r_key]
    rows = []
    for (pred, cls), side in only:
        score = float(conf[pred, cls])
        box = coord_all[pred].astype(np.float64)
        margins = {}
        if cut_pre is not None:
            margins['score_to_pre_nms_cut'] = abs(score - cut_pre)
        if cut_post is not None:
            margins['score_to_post_nms_cut'] = abs(score - cut_post)
        same = (cand['cls'] == cls) & (cand['score'] > score - score_noise) & (cand['pred'] != pred)
        if same.any():
            iou = _iou_centre(box, cand['coord'][same].astype(np.float64))
            margins['nms_iou_to_threshold'] = float(np.abs(iou - nms_thr).min())
        rows.append({'pred': int(pred), 'cls': int(cls), 'kept_by': side, 'oracle_score': score, 'margins': margins, 'box': box})
    for r in rows:
        m = r['margins']
        direct = (m.get('score_to_pre_nms_cut', 1.0) < score_noise or m.get('score_to_post_nms_cut', 1.0) < score_noise
                  or m.get('nms_iou_to_threshold', 1.0) < iou_noise)
        cascade = False
        if not direct:                    # kept / dropped because another exception of its class flipped
            for o in rows:
                if o is not r and o['cls'] == r['cls'] and _iou_centre(r['box'], o['box'][None])[0] >= nms_thr - iou_noise:
                    cascade = True
        # a post-NMS top-k cut moves whenever any other exception enters or leaves the kept set
        if not direct and not cascade and cut_post is not None and len(rows) > 1 and abs(r['oracle_score'] - cut_post) < 10 * score_noise:
            cascade = True
        r['explained'] = bool(direct or cascade)
        r['how'] = 'direct' if direct else 'cascade' if cascade else 'UNEXPLAINED'
    for r in rows:
        r.pop('box')
    rep['exceptions'] = rows
    rep['unexplained'] = int(sum(not r['explained'] for r in rows))
    return rep


# ---- detections -> COCO format fixtures -----------------------------------------------------------
def blob_masks(k, height, width, seed):
    """Seeded instance-like masks: unions of a few ellipses and axis-aligned boxes, bool [k, H, W]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    out = np.zeros((k, height, width), dtype=bool)
    for i in range(k):
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0, height), rng.uniform(0, width)
            ry, rx = rng.uniform(2, height / 3), rng.uniform(2, width / 3)
            if rng.random() < 0.5:
                out[i] |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
            else:
                out[i] |= (np.abs(yy - cy) < ry) & (np.abs(xx - cx) < rx)
    out[k - 1] = False                       # an empty mask
    if k > 1:
        out[k - 2] = True                    # a full mask (first pixel set: leading zero-length run)
    return out


# name -> (mask H, mask W, sample_info); collate_pad = [left, right, top, down, h, w] (infer.py:28-30),
# pad = (top, down, left, right, h, w) (data/transform.py Resize/Pad), as eval/coco_eval.py:160-176,192-197 read them
COCO_INFOS = {
    'plain': (64, 96, {'id': 1, 'height': 45, 'width': 70}),
    'collate': (64, 96, {'id': 2, 'height': 100, 'width': 131, 'collate_pad': [15, 16, 8, 8, 64, 96]}),
    'both': (64, 96, {'id': 3, 'height': 53, 'width': 80, 'collate_pad': [0, 0, 0, 0, 64, 96], 'pad': (4, 6, 10, 0, 64, 96)}),
    'flips': (64, 96, {'id': 4, 'height': 64, 'width': 96, 'hflip': True, 'vflip': True}),
    'hflip_up': (32, 32, {'id': 5, 'height': 75, 'width': 50, 'pad': (0, 3, 2, 1, 32, 32), 'hflip': True}),
}
