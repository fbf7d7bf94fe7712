"""Golden fixture for the detections -> COCO format step, made by the UNMODIFIED reference here.

    python tests/golden/make_golden_coco.py          (build container only: needs /root/reference)

Runs the reference's own ``COCOMetrics._recover_shape_segm`` / ``_recover_shape_bbox`` (eval/coco_eval.py:146-205,
static methods; torch CPU) on seeded blob masks and boxes for several ``sample_info`` settings and stores inputs and
outputs in ``coco_small.npz``.  (``maskUtils.encode`` itself is pycocotools, which this image does not have: the RLE
codec is pinned by known answers and round trips instead, tests/test_coco_format.py.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import build_ref  # noqa: E402
from tests.common import blob_masks, COCO_INFOS  # noqa: E402


def main():
    build_ref.import_reference()
    from eval.coco_eval import COCOMetrics
    d = {}
    for name, (H, W, info) in COCO_INFOS.items():
        masks = blob_masks(6, H, W, seed=len(name))
        g = torch.Generator().manual_seed(3)
        boxes = torch.rand(6, 4, generator=g) * 0.5 + 0.2
        rec = COCOMetrics._recover_shape_segm(torch.from_numpy(masks), info)
        box = COCOMetrics._recover_shape_bbox(boxes, info)
        print(name, masks.shape, '->', tuple(rec.shape), rec.dtype, int(rec.sum()))
        d[name + '_segm'] = np.packbits(rec.numpy().reshape(-1))
        d[name + '_shape'] = np.asarray(rec.shape, dtype=np.int64)
        d[name + '_boxes_in'] = boxes.numpy()
        d[name + '_boxes_out'] = box.numpy()
    np.savez_compressed(os.path.join(HERE, 'coco_small.npz'), **d)
    print(os.path.getsize(os.path.join(HERE, 'coco_small.npz')))


if __name__ == '__main__':
    main()
