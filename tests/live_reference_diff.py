"""Differential run of the post-process oracle against the LIVE unmodified reference (build container only: needs
/root/reference) on fresh seeded head tensors -- sizes, seeds and thresholds the committed fixtures do not contain.
Run by tests/test_oracle.py::test_post_oracle_matches_live_reference_on_fresh_cases in a subprocess (importing the
reference shadows the top-level package names ``eval`` / ``model`` / ``utils``).  Prints one JSON line."""
import functools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import build_ref
from oracle.post_oracle import PostProcessOracle
from tests.common import synthetic_heads, post_config, ANCHORS, ANCHOR_MASK
config, ref_model, builder = build_ref.import_reference()
import eval as ref_eval
worst = 0.0
n_cmp = 0
def compare(H, W, seed, thr, num_classes=80):
    global worst, n_cmp
    heads = synthetic_heads(2, H, W, seed=seed, num_classes=num_classes)
    cfg = dict(post_config(H, W, thr), num_classes=num_classes)
    post = ref_eval.OrienMaskYOLOPostProcess(nms_func=functools.partial(ref_eval.batched_nms, threshold=0.5),
                                             device=torch.device('cpu'), **cfg)
    with torch.no_grad():
        ref = post(heads)
    orc = PostProcessOracle(cfg['grid_size'], cfg['image_size'], ANCHORS, ANCHOR_MASK, num_classes, conf_thresh=thr)
    got = orc([(b.numpy(), o.numpy()) for b, o in heads])
    for r, g in zip(ref, got):
        rb, rc, rm = r['bbox'].numpy(), r['cls'].numpy(), r['mask'].numpy()
        assert rb.shape == g['bbox'].shape, (H, W, seed, thr, rb.shape, g['bbox'].shape)
        if rb.shape[0]:
            # the reference's order among equal scores is unspecified: align by (score, class, box)
            key_r = np.lexsort((rb[:, 0], rb[:, 1], rc, -rb[:, 4]))
            key_g = np.lexsort((g['bbox'][:, 0], g['bbox'][:, 1], g['cls'], -g['bbox'][:, 4]))
            assert np.array_equal(rc[key_r], g['cls'][key_g]), (H, W, seed, thr)
            worst = max(worst, float(np.abs(rb[key_r] - g['bbox'][key_g]).max()))
            assert np.array_equal(rm[key_r], g['mask'][key_g]), (H, W, seed, thr, 'mask')
        n_cmp += 1


for (H, W) in ((64, 96), (96, 64), (128, 128)):
    for seed in (11, 12, 13):
        for thr in (0.005, 0.05, 0.3):
            compare(H, W, seed, thr)
for thr in (0.005, 0.05, 0.3):                      # the VOC head layout: 20 classes, 75 channels per scale
    compare(64, 64, 14, thr, num_classes=20)

# ---- forward oracle vs the reference nn.Modules on fresh weights / sizes (both model variants) ----------------------------
from oracle.forward_oracle import forward_oracle  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402
fwd_rel = 0.0
n_fwd = 0
for plus, cls in ((True, ref_model.OrienMaskYOLOFPNPlus), (False, ref_model.OrienMaskYOLO)):
    for wseed, (H, W) in ((3, (64, 96)), (4, (96, 128))):
        sd = synthetic_state_dict(wseed, plus=plus)
        net = cls(3, 80, pretrained=None)
        net.load_state_dict(sd, strict=True)
        net.eval()
        x = synthetic_images(2, H, W, seed=20 + wseed)
        with torch.no_grad():
            ref = net(x)
        got = forward_oracle(sd, x)
        for (rb, ro), (gb, go) in zip(ref, got):
            for r, g in ((rb, gb), (ro, go)):
                assert r.shape == g.shape
                fwd_rel = max(fwd_rel, float((r - g).norm() / r.norm()))
        n_fwd += 1

# ---- the neighbours of the path on fresh cases -------------------------------------------------------------
import data.transform as T  # noqa: E402  (the reference's own data package)
from eval.coco_eval import COCOMetrics as RefMetrics  # noqa: E402
from oracle import prep_oracle, coco_oracle  # noqa: E402
from tests.common import blob_masks  # noqa: E402
rng = np.random.default_rng(17)
prep_worst, n_prep = 0.0, 0
for (h, w), size in (((37, 50), (64, 96)), ((480, 640), (544, 544)), ((211, 150), (96, 64)), ((64, 64), (64, 64))):
    img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    tr = T.FastCOCOTransform([T.FastCOCOTransform.Resize(size=size), T.FastCOCOTransform.Normalize(mean=(0, 0, 0), std=(255, 255, 255))],
                             use_cuda=False)
    ref = tr(torch.tensor(img, dtype=torch.float32)).numpy()
    got = prep_oracle.fast_transform_oracle(img, size)
    assert got.shape == ref.shape
    prep_worst = max(prep_worst, float(np.abs(got - ref).max()))       # in units of 1 (inputs 0..255 scaled to 0..1)
    n_prep += 1
img = rng.integers(0, 256, (1, 90, 61, 3), dtype=np.uint8)
tr = T.FastCOCOTransform([T.FastCOCOTransform.ShortEdgeResize(short_length=64, max_size=100),
                          T.FastCOCOTransform.Normalize(mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375))], use_cuda=False)
ref = tr(torch.tensor(img, dtype=torch.float32)).numpy()
got = prep_oracle.fast_transform_oracle(img, prep_oracle.short_edge_size(90, 61, 64, 100), mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375))
assert got.shape == ref.shape
short_worst = float(np.abs(got - ref).max())                            # in units of one standard deviation (~58 grey levels)

segm_min_iou, box_worst, n_coco = 1.0, 0.0, 0
infos = [(96, 128, {'id': 1, 'height': 333, 'width': 500}),
         (96, 128, {'id': 2, 'height': 60, 'width': 77, 'collate_pad': [3, 5, 0, 8, 96, 128]}),
         (64, 64, {'id': 3, 'height': 100, 'width': 40, 'collate_pad': [0, 0, 0, 0, 64, 64], 'pad': (2, 2, 20, 20, 64, 64), 'vflip': True}),
         (128, 96, {'id': 4, 'height': 128, 'width': 96, 'hflip': True})]
for H, W, info in infos:
    masks = blob_masks(5, H, W, seed=H + info['id'])
    boxes = torch.rand(5, 4, generator=torch.Generator().manual_seed(info['id'])) * 0.5 + 0.2
    ref_m = RefMetrics._recover_shape_segm(torch.from_numpy(masks), info).numpy().astype(bool)
    got_m = np.asarray(coco_oracle.recover_shape_segm(masks, info)).astype(bool)
    assert ref_m.shape == got_m.shape
    for a, b in zip(ref_m, got_m):
        union = (a | b).sum()
        segm_min_iou = min(segm_min_iou, 1.0 if union == 0 else float((a & b).sum()) / float(union))
    ref_b = RefMetrics._recover_shape_bbox(boxes, info).numpy()
    box_worst = max(box_worst, float(np.abs(np.asarray(coco_oracle.recover_shape_bbox(boxes.numpy(), info)) - ref_b).max()))
    n_coco += 1

# ---- visualiser blend on fresh cases ------------------------------------------------------------------------
from utils.visualizer import InferenceVisualizer as RefVis, PALETTE  # noqa: E402
from oracle.visualizer_oracle import blend_oracle  # noqa: E402
blend_worst, n_blend = 0.0, 0
for (H, W), (height, width), pad_info, alpha in (((96, 128), (333, 500), [0, 0, 0, 0, 96, 128], 0.5),
                                                 ((64, 96), (40, 61), [3, 5, 2, 8, 64, 96], 0.35)):
    masks = blob_masks(6, H, W, seed=H + width)[:4]
    image = torch.rand(height, width, 3, generator=torch.Generator().manual_seed(width)) * 255
    colors = torch.tensor(PALETTE, dtype=torch.float32)[(torch.arange(4) * 5 + 7) % len(PALETTE)]
    vis = RefVis.__new__(RefVis)
    vis.alpha = alpha
    soft = RefVis._recover_shape_segm(torch.from_numpy(masks), width, height, pad_info)
    order = soft.sum(dim=2).sum(dim=1).argsort()
    out = image.clone()
    vis.plot_all_mask(soft[order], out, colors[order])
    got, got_order, _ = blend_oracle(image.numpy(), masks, colors.numpy(), pad_info, alpha)
    assert np.array_equal(np.asarray(got_order), order.numpy())
    blend_worst = max(blend_worst, float(np.abs(np.asarray(got) - out.numpy()).max()))      # grey levels (0..255)
    n_blend += 1
print(json.dumps({'cases': n_cmp, 'max_box_diff': worst, 'forward_cases': n_fwd, 'forward_rel_l2': fwd_rel,
                  'blend_cases': n_blend, 'blend_max_diff': blend_worst,
                  'prep_cases': n_prep, 'prep_max_diff': prep_worst, 'prep_short_edge_diff': short_worst,
                  'coco_cases': n_coco, 'coco_min_mask_iou': segm_min_iou, 'coco_max_box_diff_px': box_worst}))
