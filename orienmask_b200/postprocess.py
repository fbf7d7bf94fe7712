"""Host-side mirror of the reference post-process interface, backed by the CUDA C-ABI kernels.

``OrienMaskYOLOPostProcess`` keeps the constructor and call signature of
``/root/reference/eval/orienmask_yolo_postprocess.py:8-12,63-64`` and returns the same structure
(``list`` over images of ``{'bbox': float32 [K,5], 'mask': bool [K,H,W], 'cls': int64 [K]}``,
:124,166).  Everything between is four batched kernel families (decode+select, NMS, masks) with a
single host read of the per-image counts at the very end; there is no CPU path.
"""
import ctypes
import functools

import torch

from . import _lib
from .function import batched_nms, nms as plain_nms


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (list, tuple)) else (int(v), int(v))


class PaddedDetections:
    """Fixed-size (padded) batch result living on the device; nothing here forces a host sync."""

    def __init__(self, det, cls, anchor, keep, count, mask):
        self.det, self.cls, self.anchor, self.keep, self.count, self.mask = det, cls, anchor, keep, count, mask
        self.packed = None
        self.count_host = None      # pinned copy of `count`, enqueued right after the NMS kernel (before the mask kernel)
        self.nms_done = None        # event recorded at that point: counts and records are final, the masks are still being written

    def records(self):
        """[B, nms_post, 6] fp32 (cx, cy, w, h, score, cls) -- the unit the multi-GPU gather moves."""
        return torch.cat([self.det, self.cls.to(torch.float32).unsqueeze(-1)], dim=-1)

    def to_list(self):
        """The reference's return value (eval/orienmask_yolo_postprocess.py:124,166): per image a dict of tensors trimmed to its K
        detections.  K has to reach the host -- the one synchronisation of the post-process -- but only the NMS kernel has to have
        finished for it: the mask kernel (most of the post-process time) keeps running while the host slices, and the returned
        mask tensors are ordinary stream-ordered results."""
        if self.nms_done is not None:
            self.nms_done.synchronize()
            counts = self.count_host.tolist()
        else:
            counts = self.count.tolist()
        out = []
        for b, k in enumerate(counts):
            out.append({'bbox': self.det[b, :k], 'mask': self.mask[b, :k].view(torch.bool), 'cls': self.cls[b, :k]})
        return out


class OrienMaskYOLOPostProcess:
    def __init__(self, grid_size, image_size, anchors, anchor_mask, num_classes,
                 conf_thresh=0.05, nms_func=None, nms_pre=400, nms_post=100, orien_thresh=0.3, device=None):
        self.device = torch.device(device) if device is not None else None
        self.grid_size = [(int(g[0]), int(g[1])) for g in grid_size]
        self.image_h, self.image_w = _pair(image_size)
        self.anchors = [(float(a[0]), float(a[1])) for a in anchors]
        self.anchor_mask = [list(int(i) for i in m) for m in anchor_mask]
        self.num_classes = int(num_classes)
        self.conf_thresh, self.nms_pre, self.nms_post = float(conf_thresh), int(nms_pre), int(nms_post)
        self.orien_thresh = float(orien_thresh)
        self.nms = nms_func if nms_func is not None else batched_nms
        self.nms_thresh, self.nms_semantics = self._resolve_nms(self.nms)
        if len(self.grid_size) > _lib.OM_MAX_SCALES or len(self.anchors) > _lib.OM_MAX_ANCHORS:
            raise ValueError('at most %d scales and %d anchors are supported' % (_lib.OM_MAX_SCALES, _lib.OM_MAX_ANCHORS))
        cfg = _lib.PostConfig()
        cfg.num_scales = len(self.grid_size)
        cfg.num_classes = self.num_classes
        cfg.image_h, cfg.image_w = self.image_h, self.image_w
        for s, ((nH, nW), m) in enumerate(zip(self.grid_size, self.anchor_mask)):
            cfg.grid_h[s], cfg.grid_w[s] = nH, nW
            cfg.anchors_per_scale[s] = len(m)
            for j, a in enumerate(m):
                cfg.anchor_index[s][j] = a
        cfg.total_anchors = len(self.anchors)
        for a, (w, h) in enumerate(self.anchors):
            cfg.anchor_w[a], cfg.anchor_h[a] = w, h
        cfg.conf_thresh, cfg.nms_thresh, cfg.orien_thresh = self.conf_thresh, self.nms_thresh, self.orien_thresh
        cfg.nms_pre, cfg.nms_post = self.nms_pre, self.nms_post
        cfg.nms_semantics = self.nms_semantics
        self._cfg = cfg

    @staticmethod
    def _resolve_nms(func):
        """The builder hands over functools.partial(batched_nms, threshold=t) (trainer/builder.py:67-77)."""
        kw = {}
        base = func
        if isinstance(func, functools.partial):
            base, kw = func.func, dict(func.keywords)
        if base is batched_nms and kw.get('normalized', True):
            from .function import _mode
            return float(kw.get('threshold', 0.5)), _mode(kw.get('semantics'))
        raise NotImplementedError(
            'orienmask_b200 fuses class-wise NMS into the batched kernel; nms_func must be '
            'orienmask_b200.function.batched_nms (optionally functools.partial with threshold=...), got %r' % (func,))

    def __call__(self, predict):
        return self.apply(predict)

    def apply(self, predict):
        return self.apply_padded(predict).to_list()

    # -------------------------------------------------------------------------------------------
    def _prep(self, t, channels, h, w, what):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError('orienmask_b200 post-process needs CUDA tensors (%s is %s); there is no CPU path'
                               % (what, getattr(t, 'device', type(t))))
        if tuple(t.shape[1:]) != (channels, h, w):
            raise ValueError('%s has shape %s, expected [B, %d, %d, %d]' % (what, tuple(t.shape), channels, h, w))
        if t.dtype != torch.float32:
            t = t.float()
        if t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w:
            t = t.contiguous()
        return t

    def apply_padded(self, predict):
        lib = _lib.lib()
        cfg = self._cfg
        S = cfg.num_scales
        if len(predict) != S:
            raise ValueError('expected %d scales, got %d' % (S, len(predict)))
        B = int(predict[0][0].shape[0])
        h4, w4 = self.image_h // 4, self.image_w // 4
        bboxes, oriens = [], []
        for s in range(S):
            nH, nW = self.grid_size[s]
            A = len(self.anchor_mask[s])
            bboxes.append(self._prep(predict[s][0], A * (5 + self.num_classes), nH, nW, 'bbox[%d]' % s))
            oriens.append(self._prep(predict[s][1], 2 * A, h4, w4, 'orien[%d]' % s))
        dev = bboxes[0].device
        if self.device is not None and self.device.type == 'cuda' and self.device.index not in (None, dev.index):
            raise RuntimeError('post-process was built for %s but inputs live on %s' % (self.device, dev))
        with torch.cuda.device(dev):
            stream = _lib.stream_ptr()
            bb_ptr = (_lib.c_vp * S)(*[t.data_ptr() for t in bboxes])
            bb_str = (_lib.c_i64 * S)(*[t.stride(0) for t in bboxes])
            or_ptr = (_lib.c_vp * S)(*[t.data_ptr() for t in oriens])
            or_str = (_lib.c_i64 * S)(*[t.stride(0) for t in oriens])
            nbytes = ctypes.c_size_t(0)
            _lib.check(lib.om_post_workspace_bytes(ctypes.byref(cfg), B, ctypes.byref(nbytes)), 'om_post_workspace_bytes')
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            i32 = dict(dtype=torch.int32, device=dev)
            cand_count = torch.empty(B, **i32)
            cand_det = torch.empty(B, self.nms_pre, 5, dtype=torch.float32, device=dev)
            cand_cls = torch.empty(B, self.nms_pre, **i32)
            cand_pred = torch.empty(B, self.nms_pre, **i32)
            _lib.check(lib.om_decode_select(ctypes.byref(cfg), bb_ptr, bb_str, B, _lib.ptr(ws), _lib.ptr(cand_count),
                                            _lib.ptr(cand_det), _lib.ptr(cand_cls), _lib.ptr(cand_pred), stream),
                       'om_decode_select')
            det_count = torch.empty(B, **i32)
            det = torch.empty(B, self.nms_post, 5, dtype=torch.float32, device=dev)
            det_cls = torch.empty(B, self.nms_post, dtype=torch.int64, device=dev)
            det_anchor = torch.empty(B, self.nms_post, **i32)
            det_keep = torch.empty(B, self.nms_post, **i32)
            packed = torch.empty(B, self.nms_post * 6 + 1, dtype=torch.float32, device=dev)
            _lib.check(lib.om_batched_nms(ctypes.byref(cfg), _lib.ptr(cand_count), _lib.ptr(cand_det), _lib.ptr(cand_cls),
                                          _lib.ptr(cand_pred), B, _lib.ptr(det_count), _lib.ptr(det), _lib.ptr(det_cls),
                                          _lib.ptr(det_anchor), _lib.ptr(det_keep), _lib.ptr(packed), stream), 'om_batched_nms')
            count_host = torch.empty(B, dtype=torch.int32, pin_memory=True)
            count_host.copy_(det_count, non_blocking=True)
            nms_done = torch.cuda.Event()
            nms_done.record()
            mask = torch.empty(B, self.nms_post, self.image_h, self.image_w, dtype=torch.uint8, device=dev)
            _lib.check(lib.om_mask_assemble(ctypes.byref(cfg), or_ptr, or_str, _lib.ptr(det_count), _lib.ptr(det),
                                            _lib.ptr(det_anchor), B, _lib.ptr(mask), stream), 'om_mask_assemble')
        out = PaddedDetections(det, det_cls, det_anchor, det_keep, det_count, mask)
        out.packed = packed          # [B, nms_post*6+1] rows written by the NMS kernel: what sharding.gather_detections moves
        out.count_host, out.nms_done = count_host, nms_done
        out.candidates = dict(count=cand_count, det=cand_det, cls=cand_cls, pred=cand_pred)
        out._keepalive = (bboxes, oriens, ws)
        return out
