"""Visualiser mask blend: oracle vs the reference's golden output (CPU), CUDA kernels vs oracle (GPU).
Tolerance 1e-3 on 0..255 pixel values: the reference sums the per-instance terms with torch.sum (order unspecified)."""
import numpy as np
import pytest
import torch

from tests.common import GOLDEN, blob_masks
from oracle.visualizer_oracle import blend_oracle


def _case():
    g = np.load(GOLDEN + '/blend_small.npz')
    return g, blob_masks(7, 64, 96, seed=3)[:5]


def test_oracle_matches_reference_golden():
    g, masks = _case()
    out, order, _ = blend_oracle(g['image'], masks, g['colors'], g['pad_info'].tolist(), float(g['alpha']))
    assert order.tolist() == g['order'].tolist()
    assert np.abs(out - g['out']).max() < 1e-3


@pytest.mark.gpu
def test_kernels_match_oracle_and_golden():
    from orienmask_b200.visualizer import blend_masks
    g, masks = _case()
    image = torch.from_numpy(g['image']).cuda().contiguous()
    order, areas = blend_masks(image, torch.from_numpy(masks).cuda(), torch.from_numpy(g['colors']), g['pad_info'].tolist(), float(g['alpha']))
    ref, ref_order, ref_areas = blend_oracle(g['image'], masks, g['colors'], g['pad_info'].tolist(), float(g['alpha']))
    assert order.cpu().tolist() == ref_order.tolist() == g['order'].tolist()
    assert np.allclose(areas.cpu().numpy(), ref_areas, rtol=1e-6)
    assert np.abs(image.cpu().numpy() - ref).max() < 1e-3
    assert np.abs(image.cpu().numpy() - g['out']).max() < 1e-3


@pytest.mark.gpu
def test_visualizer_call_on_post_process_output():
    """InferenceVisualizer(detections, image, pad_info) end to end on real post-process output (utils/visualizer.py:46-80)."""
    import functools
    import random
    import orienmask_b200 as ob
    from tests.common import synthetic_heads, post_config
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'),
                                       **post_config(64, 96, 0.005))
    det = post([(b.cuda(), o.cuda()) for b, o in synthetic_heads(1, 64, 96, seed=5)])[0]
    thr = float(det['bbox'][:, -1].median())
    vis = ob.InferenceVisualizer('COCO', torch.device('cuda:0'), with_mask=True, conf_thresh=thr, alpha=0.6)
    image = torch.rand(50, 75, 3, generator=torch.Generator().manual_seed(1)) * 255
    pad_info = [15, 16, 8, 8, 64, 96]
    random.seed(7)
    shown = vis(det, image.cuda(), pad_info)
    assert shown.shape == (50, 75, 3) and shown.dtype == np.uint8
    # the blend alone, against the oracle with the same colours
    keep = (det['bbox'][:, -1] > thr).cpu().numpy()
    random.seed(7)
    idx = torch.arange(int(keep.sum())) * 5 + random.randint(1, len(ob.visualizer.PALETTE))
    colors = np.asarray(ob.visualizer.PALETTE, np.float32)[(idx % len(ob.visualizer.PALETTE)).numpy()]
    ref, _, _ = blend_oracle(image.numpy(), det['mask'].cpu().numpy()[keep], colors, pad_info, 0.6)
    plain = ob.InferenceVisualizer('COCO', torch.device('cuda:0'), with_mask=True, conf_thresh=thr, alpha=0.6)
    plain.plot_one_box = lambda *a, **k: None                     # compare the blended pixels without boxes / labels
    random.seed(7)
    blended = plain(det, image.cuda(), pad_info)
    assert np.abs(blended.astype(np.int32) - np.round(ref).astype(np.int32)).max() <= 1


def test_default_label_tables_match_the_reference_dataset_classes():
    """InferenceVisualizer('COCO' | 'VOC', ...) labels boxes with the reference's own class names (utils/visualizer.py:36-38)."""
    import os
    import re
    from orienmask_b200.visualizer import DATASET_LABELS, InferenceVisualizer
    assert len(DATASET_LABELS['COCO'][0]) == len(DATASET_LABELS['COCO'][1]) == 80 and len(DATASET_LABELS['VOC'][1]) == 20
    v = InferenceVisualizer('COCO', 'cpu')
    assert v.classes[0] == 'person' and v.classes[79] == 'toothbrush' and int(v.cat2label[79]) == 90
    assert InferenceVisualizer('Other', 'cpu', classes=['a'], cat2label=[7]).classes == ['a']
    src = '/root/reference/data/dataset.py'
    if os.path.isfile(src):                                    # build container: compare with the reference's tables
        text = open(src).read()
        tables = re.findall(r"CLASSES = \[(.*?)\]", text, re.S)
        ids = re.findall(r"CAT2LABEL = \[(.*?)\]", text, re.S)
        for name, names_src, ids_src in zip(('COCO', 'VOC'), tables, ids):
            assert re.findall(r"'([^']+)'", names_src) == DATASET_LABELS[name][1]
            assert [int(t) for t in re.findall(r"\d+", ids_src)] == list(DATASET_LABELS[name][0])
