"""Host-side mirror of the reference's ``InferenceVisualizer`` (``/root/reference/utils/visualizer.py:33-127``).

Same constructor arguments and call signature (``visualizer(detections, image, pad_info) -> uint8 numpy image``).  The GPU
part -- resizing every kept instance mask to the image (bilinear, not rounded), ordering the instances by area and the
alpha blend of ``plot_all_mask`` -- is two kernels (``om_mask_areas``, ``om_mask_blend``) that never materialise the resized
masks; boxes and labels are drawn by the same cv2 calls as the reference, on the host, after the one D2H copy of the image.
The class names / category ids default to the tables of the reference's dataset classes for ``dataset='COCO'`` / ``'VOC'``
(``data/dataset.py:42-62,104-112``); other datasets pass ``classes=`` / ``cat2label=``.
"""
import random

import torch

from . import _lib

# the reference's PALETTE (utils/visualizer.py:10-30): nineteen Material-Design colours, RGB
PALETTE = ((244, 67, 54), (233, 30, 99), (156, 39, 176), (103, 58, 183), (63, 81, 181), (33, 150, 243), (3, 169, 244),
           (0, 188, 212), (0, 150, 136), (76, 175, 80), (139, 195, 74), (205, 220, 57), (255, 235, 59), (255, 193, 7),
           (255, 152, 0), (255, 87, 34), (121, 85, 72), (158, 158, 158), (96, 125, 139))
COCO_CAT_IDS = [i for i in range(1, 91) if i not in (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)]
# label text per dataset name, as the reference's dataset classes spell it (data/dataset.py:42-62,104-112: darknet-style names)
DATASET_LABELS = {
    'COCO': (COCO_CAT_IDS, (
        'person bicycle car motorbike aeroplane bus train truck boat traffic-light fire-hydrant stop-sign parking-meter bench '
        'bird cat dog horse sheep cow elephant bear zebra giraffe backpack umbrella handbag tie suitcase frisbee skis snowboard '
        'sports-ball kite baseball-bat baseball-glove skateboard surfboard tennis-racket bottle wine-glass cup fork knife spoon '
        'bowl banana apple sandwich orange broccoli carrot hot-dog pizza donut cake chair sofa potted-plant bed dining-table '
        'toilet tv-monitor laptop mouse remote keyboard cell-phone microwave oven toaster sink refrigerator book clock vase '
        'scissors teddy-bear hair-drier toothbrush').split()),
    'VOC': (list(range(1, 21)), (
        'aeroplane bicycle bird boat bottle bus car cat chair cow dining-table dog horse motorbike person potted-plant sheep '
        'sofa train tv-monitor').split()),
}


def blend_masks(image, masks, colors, pad_info, alpha):
    """In-place ``plot_all_mask`` of the area-sorted, resized ``masks`` (uint8/bool [K,H,W], CUDA) on ``image`` (fp32 [h,w,3]).
    ``colors``: fp32 [K,3] per instance.  Returns (drawing order, soft-mask areas)."""
    if not (image.is_cuda and masks.is_cuda):
        raise RuntimeError('orienmask_b200 visualiser blend needs CUDA tensors; there is no CPU path')
    if image.dtype != torch.float32 or not image.is_contiguous() or image.dim() != 3 or image.shape[2] != 3:
        raise ValueError('image must be a contiguous float32 [h, w, 3] tensor')
    k = int(masks.shape[0])
    m = masks.view(torch.uint8) if masks.dtype == torch.bool else masks
    m = m.contiguous()
    left, right, top, down = (int(v) for v in pad_info[:4])
    cfg = _lib.BlendConfig()
    cfg.mask_h, cfg.mask_w = int(m.shape[1]), int(m.shape[2])
    cfg.top, cfg.left = top, left
    cfg.crop_h, cfg.crop_w = cfg.mask_h - top - down, cfg.mask_w - left - right
    cfg.out_h, cfg.out_w = int(image.shape[0]), int(image.shape[1])
    cfg.alpha = float(alpha)
    lib = _lib.lib()
    with torch.cuda.device(image.device):
        stream = _lib.stream_ptr()
        scratch = torch.empty(k, dtype=torch.float64, device=image.device)
        areas = torch.empty(k, dtype=torch.float32, device=image.device)
        _lib.check(lib.om_mask_areas(cfg, _lib.ptr(m), k, _lib.ptr(scratch), _lib.ptr(areas), stream), 'om_mask_areas')
        order = areas.argsort().to(torch.int32)                                   # utils/visualizer.py:69
        col = colors.to(device=image.device, dtype=torch.float32).contiguous()
        _lib.check(lib.om_mask_blend(cfg, _lib.ptr(m), k, _lib.ptr(order), _lib.ptr(col), _lib.ptr(image), stream), 'om_mask_blend')
    return order, areas


class InferenceVisualizer:
    def __init__(self, dataset, device, with_mask=True, conf_thresh=0.3, alpha=0.5, line_thickness=1, classes=None, cat2label=None):
        self.dataset = dataset
        ids, names = DATASET_LABELS.get(dataset, (COCO_CAT_IDS, None))          # utils/visualizer.py:36-38: <dataset>Dataset.CAT2LABEL / .CLASSES
        self.cat2label = torch.tensor(cat2label if cat2label is not None else ids, dtype=torch.uint8, device=device)
        self.classes = list(classes) if classes is not None else (names or [str(int(c)) for c in self.cat2label.tolist()])
        self.device = device
        self.with_mask = with_mask
        self.conf_thresh = conf_thresh
        self.alpha = alpha
        self.line_thickness = line_thickness
        self.palette = torch.tensor(PALETTE, dtype=torch.float32, device=device)

    def __call__(self, detections, image, pad_info):
        pred_show = image.clone().contiguous().float()                            # (h, w, 3), utils/visualizer.py:48
        height, width = pred_show.shape[:2]
        bbox, cls = detections['bbox'], detections['cls']
        keep = bbox[:, -1] > self.conf_thresh
        bbox, cls = bbox[keep], cls[keep]
        boxes = []
        if bbox.numel() > 0:
            all_xyxy = self._recover_shape_bbox(bbox[:, :4], width, height, pad_info)
            scores = bbox[:, -1]
            names = [self.classes[int(c)] for c in cls]
            idx = torch.arange(bbox.size(0)) * 5 + random.randint(1, self.palette.size(0))
            colors = self.palette[(idx % self.palette.size(0)).to(self.palette.device)]
            if self.with_mask:
                blend_masks(pred_show, detections['mask'][keep], colors, pad_info, self.alpha)
            boxes = list(zip(all_xyxy.cpu().tolist(), scores.cpu().tolist(), names, colors.cpu().tolist()))
        out = pred_show.round().to(torch.uint8).cpu().numpy()
        for xyxy, score, name, color in boxes:
            self.plot_one_box(xyxy, '%s %.2f' % (name, score), out, color)
        return out

    def plot_one_box(self, bbox, text, image, color):
        import cv2                                                                  # utils/visualizer.py:82-93
        x1, y1, x2, y2 = (int(v) for v in bbox)
        cv2.rectangle(image, (x1, y1), (x2, y2), color, thickness=self.line_thickness)
        face, scale, thick = cv2.FONT_HERSHEY_DUPLEX, 0.4, 1
        tw, th = cv2.getTextSize(text, face, scale, thick)[0]
        cv2.rectangle(image, (x1, y1), (x1 + tw, y1 - th - 4), color, -1)
        cv2.putText(image, text, (x1, y1 - 3), face, scale, [255, 255, 255], thick, cv2.LINE_AA)

    @classmethod
    def _recover_shape_bbox(cls, bbox, width, height, pad_info):
        bx, by, bw, bh = bbox.split(1, dim=-1)                                     # utils/visualizer.py:102-120
        left, right, top, down, h, w = pad_info
        nh, nw = h - top - down, w - left - right
        bx, by, bw, bh = (bx * w - left) / nw, (by * h - top) / nh, bw * w / nw, bh * h / nh
        xyxy = torch.cat([(bx - bw / 2) * width, (by - bh / 2) * height, (bx + bw / 2) * width, (by + bh / 2) * height], dim=-1)
        return xyxy.round().long()
