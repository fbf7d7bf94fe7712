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
