"""Host-side mirror of the reference's detections -> COCO format step, backed by one CUDA kernel (``om_mask_rle``).

``COCOMetrics`` keeps the constructor and the methods the callers use (``/root/reference/eval/coco_eval.py:23-74``:
``to_coco_format``, ``update_results``, ``reset``, ``save_as_json``, ``update_from_json``; callers:
``trainer/tester.py:46-50``, ``infer.py:159-163``).  ``to_coco_format(image_info, detections)`` returns the same
``{'bbox': [...], 'segm': [...]}`` lists of dicts.  What differs is where the work happens: the reference copies every
instance mask to the host (one synchronising D2H each) and resizes / encodes it on the CPU; here the crop, flips,
bilinear resize, rounding, run-length encoding and the COCO string compression of all instances of the batch are one
kernel, and only the strings cross PCIe.  ``coco_eval`` hands the collected results to pycocotools' ``COCOeval`` (the AP
accumulation itself is that library's CPU code, outside the path) and fills the attributes ``trainer/tester.py:52-67`` reads.
"""
import contextlib
import ctypes
import json
import os

import numpy as np
import torch

from . import _lib


def _crop_window(info, H, W):
    """(top, left, crop_h, crop_w) after the two crops of eval/coco_eval.py:192-197."""
    top = left = 0
    h, w = H, W
    if info.get('collate_pad') is not None:
        l, r, t, d = info['collate_pad'][:4]
        top, left, h, w = top + t, left + l, h - t - d, w - l - r
    if info.get('pad') is not None:
        t, d, l, r = info['pad'][:4]
        top, left, h, w = top + t, left + l, h - t - d, w - l - r
    return int(top), int(left), int(h), int(w)


def encode_masks(masks, counts, infos, cap=4096):
    """RLE strings of every instance of a batch.

    masks   list (per image) of uint8/bool CUDA tensors [>= counts[b], H, W] (contiguous), or None where counts[b] == 0
    counts  list of instance counts; infos: list of sample_info dicts
    Returns a list (per image) of lists of {'size': [h, w], 'counts': str}.
    """
    B = len(masks)
    max_inst = max(counts) if counts else 0
    if max_inst == 0:
        return [[] for _ in range(B)]
    dev = next(m.device for m in masks if m is not None)
    arr = (_lib.RleImage * B)()
    max_oh = max_mh = max_mw = 1
    keep = []
    for b in range(B):
        r = arr[b]
        r.count = int(counts[b])
        if r.count == 0:
            r.out_h = r.out_w = r.crop_h = r.crop_w = r.mask_h = r.mask_w = 1
            continue
        m = masks[b]
        if not m.is_cuda:
            raise RuntimeError('orienmask_b200 COCO formatting needs CUDA masks (got %s); there is no CPU path' % m.device)
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        if not m.is_contiguous():
            m = m.contiguous()
        keep.append(m)
        H, W = int(m.shape[-2]), int(m.shape[-1])
        top, left, ch, cw = _crop_window(infos[b], H, W)
        if ch < 1 or cw < 1:
            raise ValueError('padding %r removes the whole %dx%d mask' % (infos[b], H, W))
        r.mask, r.mask_h, r.mask_w = m.data_ptr(), H, W
        r.top, r.left, r.crop_h, r.crop_w = top, left, ch, cw
        r.out_h, r.out_w = int(infos[b]['height']), int(infos[b]['width'])
        r.hflip, r.vflip = int(bool(infos[b].get('hflip', False))), int(bool(infos[b].get('vflip', False)))
        max_oh, max_mh, max_mw = max(max_oh, r.out_h), max(max_mh, H), max(max_mw, W)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    lib = _lib.lib()
    with torch.cuda.device(dev):
        while True:
            str_cap = 4 * cap
            rle_counts = torch.empty(B * max_inst, cap, dtype=torch.int32, device=dev)
            n_counts = torch.zeros(B * max_inst, dtype=torch.int32, device=dev)
            rle_str = torch.empty(B * max_inst, str_cap, dtype=torch.uint8, device=dev)
            str_len = torch.zeros(B * max_inst, dtype=torch.int32, device=dev)
            _lib.check(lib.om_mask_rle(_lib.ptr(raw), B, max_inst, max_oh, max_mh, max_mw, cap, str_cap, _lib.ptr(rle_counts), _lib.ptr(n_counts),
                                       _lib.ptr(rle_str), _lib.ptr(str_len), _lib.stream_ptr()), 'om_mask_rle')
            lens = str_len.cpu().numpy()                         # the one sync of the formatting step
            if (lens >= 0).all():
                break
            cap = max(int(n_counts.max().item()), 2 * cap)       # some instance has more runs (or longer text) than reserved
    # only the bytes of the strings cross PCIe: pack them densely on the device (tensor plumbing), one D2H copy
    width = max(int(lens.max()), 1)
    used = torch.arange(width, device=dev)[None, :] < str_len[:, None]
    text = rle_str[:, :width][used].cpu().numpy().tobytes().decode('utf-8')
    ends = np.cumsum(np.maximum(lens, 0))
    out = []
    for b in range(B):
        size = [int(arr[b].out_h), int(arr[b].out_w)]
        row = []
        for k in range(int(counts[b])):
            i = b * max_inst + k
            row.append({'size': size, 'counts': text[ends[i] - lens[i]:ends[i]]})
        out.append(row)
    return out


class COCOMetrics:
    def __init__(self, gt_file, cat2label, with_mask, save_dir):
        self.gt_file = gt_file
        self.cat2label = torch.tensor(cat2label)
        self.with_mask = with_mask
        self.bbox_pred_file = os.path.join(save_dir, 'bbox_prediction.json')
        self.segm_pred_file = os.path.join(save_dir, 'segm_prediction.json')
        self.reset()

    metric_keys = ['AP', 'AP50', 'AP75', 'APS', 'APM', 'APL', 'AR1', 'AR10', 'AR100', 'ARS', 'ARM', 'ARL']

    def reset(self):
        self.bbox_results = []
        self.segm_results = []
        self.bbox_eval_stats, self.segm_eval_stats = [], []
        self.bbox_eval_per_cats_stats, self.segm_eval_per_cats_stats = [], []

    def to_coco_format(self, image_info, detections):
        result = {'bbox': self._to_bbox_coco_format(image_info, detections)}
        if self.with_mask:
            result['segm'] = self._to_segm_coco_format(image_info, detections)
        return result

    def update_results(self, coco_format):
        self.bbox_results += coco_format['bbox']
        if self.with_mask:
            self.segm_results += coco_format['segm']

    def save_as_json(self, filename):
        with open(filename, 'w') as handle:
            json.dump({'bbox': self.bbox_results, 'segm': self.segm_results}, handle)

    def update_from_json(self, filename):
        update = json.load(open(filename))
        self.bbox_results += update['bbox']
        self.segm_results += update['segm']

    def coco_eval(self, per_cats=False):
        """AP / AR accumulation (eval/coco_eval.py:77-106,206-219).  The accumulation is pycocotools' ``COCOeval`` -- a
        third-party CPU library outside the path; this method only hands it the results collected here, so that
        ``trainer/tester.py:52-55`` runs unchanged.  Raises when pycocotools is not installed (it is not in this image)."""
        try:
            from pycocotools.coco import COCO
            from pycocotools.cocoeval import COCOeval
        except ImportError as e:
            raise RuntimeError('COCOMetrics.coco_eval needs pycocotools (AP accumulation is not part of orienmask_b200); the '
                               'results are available through save_as_json(): %s' % e)
        log = {}
        quiet = open(os.devnull, 'w')
        try:
            with contextlib.redirect_stdout(quiet):
                gt = COCO(self.gt_file)
                jobs = [('bbox', self.bbox_pred_file, self.bbox_results)]
                if self.with_mask:
                    jobs.append(('segm', self.segm_pred_file, self.segm_results))
                for kind, path, results in jobs:
                    with open(path, 'w') as handle:
                        json.dump(results, handle)
                    ev = COCOeval(gt, gt.loadRes(path), iouType=kind)
                    ev.evaluate()
                    ev.accumulate()
                    ev.summarize()
                    setattr(self, kind + '_eval_stats', ev.stats)
                    if per_cats:
                        setattr(self, kind + '_eval_per_cats_stats', self._per_category_ap(ev))
                    for key, value in zip(self.metric_keys, ev.stats.tolist()):
                        log['%s_%s' % (kind, key)] = value
        finally:
            quiet.close()
        return log

    def _per_category_ap(self, ev):
        """AP per category in percent: mean of the valid (> -1) precisions over IoU thresholds and recall points, all areas,
        the largest max-dets setting (precision axes: iou, recall, category, area range, max dets)."""
        prec = ev.eval['precision']
        if prec.shape[2] != self.cat2label.numel():
            raise ValueError('%d categories evaluated, %d in cat2label' % (prec.shape[2], self.cat2label.numel()))
        out = []
        for c in range(prec.shape[2]):
            p = prec[:, :, c, 0, -1]
            p = p[p > -1]
            out.append(float(np.mean(p) * 100) if p.size else float('nan'))
        return out

    # ---- eval/coco_eval.py:129-188 -------------------------------------------------------------
    @staticmethod
    def _recover_shape_bbox(bbox, sample_info):
        bx, by, bw, bh = bbox.split(1, dim=-1)
        if sample_info.get('collate_pad') is not None:
            left, right, top, down, h, w = sample_info['collate_pad']
            nh, nw = h - top - down, w - left - right
            bx, by, bw, bh = (bx * w - left) / nw, (by * h - top) / nh, bw * w / nw, bh * h / nh
        if sample_info.get('pad') is not None:
            top, down, left, right, h, w = sample_info['pad']
            nh, nw = h - top - down, w - left - right
            bx, by, bw, bh = (bx * w - left) / nw, (by * h - top) / nh, bw * w / nw, bh * h / nh
        if sample_info.get('hflip', False):
            bx = 1 - bx
        if sample_info.get('vflip', False):
            by = 1 - by
        oh, ow = sample_info['height'], sample_info['width']
        return torch.cat([(bx - bw / 2) * ow, (by - bh / 2) * oh, bw * ow, bh * oh], dim=-1)

    def _to_bbox_coco_format(self, batch_info, detections):
        results = []
        for info, det in zip(batch_info, detections):
            bbox, cls = det['bbox'], det['cls']
            if bbox.numel() == 0:
                continue
            xywh = self._recover_shape_bbox(bbox[:, :4], info).tolist()
            scores = bbox[:, -1].tolist()
            cats = self.cat2label[cls.flatten().cpu()].tolist()
            for box, score, cat in zip(xywh, scores, cats):
                results.append({'image_id': info['id'], 'category_id': cat, 'bbox': box, 'score': score})
        return results

    # ---- eval/coco_eval.py:108-127 on the GPU --------------------------------------------------
    def _to_segm_coco_format(self, batch_info, detections):
        counts = [int(det['bbox'].shape[0]) for det in detections]
        rles = encode_masks([det['mask'] if n else None for det, n in zip(detections, counts)], counts, batch_info)
        results = []
        for info, det, n, enc in zip(batch_info, detections, counts, rles):
            if n == 0:
                continue
            scores = det['bbox'][:, -1].tolist()
            cats = self.cat2label[det['cls'].flatten().cpu()].tolist()
            for rle, score, cat in zip(enc, scores, cats):
                results.append({'image_id': info['id'], 'category_id': cat, 'segmentation': rle, 'score': score})
        return results
