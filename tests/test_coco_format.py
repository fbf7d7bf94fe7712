"""Detections -> COCO format: oracle vs the reference's golden outputs and hand-derived RLE known
answers (CPU); CUDA kernel vs oracle, bit-exact on counts and strings (GPU)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from tests.common import GOLDEN, COCO_INFOS, blob_masks
from oracle import coco_oracle as co


def _gold(name):
    g = np.load(GOLDEN + '/coco_small.npz')
    shape = tuple(int(v) for v in g[name + '_shape'])
    segm = np.unpackbits(g[name + '_segm'])[:int(np.prod(shape))].reshape(shape)
    return segm, g[name + '_boxes_in'], g[name + '_boxes_out']


def _iou(a, b):
    a, b = a.astype(bool), b.astype(bool)
    union = (a | b).sum()
    return 1.0 if union == 0 else float((a & b).sum()) / float(union)


# ---- oracle pinned ---------------------------------------------------------------------------------
def test_rle_known_answers():
    """Hand-derived from maskApi.c (see oracle/coco_oracle.py header for the loop)."""
    assert co.rle_encode(np.ones((2, 2), np.uint8)).tolist() == [0, 4]             # first pixel set -> empty run of zeros first
    assert co.rle_encode(np.zeros((3, 2), np.uint8)).tolist() == [6]
    m = np.array([[0, 1], [1, 1], [0, 0]], np.uint8)                                # column-major stream: 0 1 0 | 1 1 0
    assert co.rle_encode(m).tolist() == [1, 1, 1, 2, 1]
    assert co.rle_to_string([0, 4]) == b'04'                                         # 0 -> '0' (48), 4 -> '4' (52)
    assert co.rle_to_string([100]) == b'T3'                                          # 100 = 3*32 + 4: (4 | 0x20) + 48 = 'T', 3 + 48 = '3'
    assert co.rle_to_string([3, 5, 7, 0]) == b'357K'                                 # i = 3: 0 - 5 = -5 -> 27 + 48 = 'K'
    assert co.rle_to_string([1, 1, 1, 50]) == b'111a1'                               # 50 - 1 = 49 = 1*32 + 17; bit 4 set and x = 1 != -1 -> more
    assert co.rle_from_string(b'357K').tolist() == [3, 5, 7, 0]


def test_rle_run_lengths_match_an_independent_implementation():
    """pycocotools is in neither the reference tree nor this image, so the codec restatement (oracle/coco_oracle.py) has no binary to
    be pinned against.  The run-length half of it does have an independent third-party implementation here: transformers' SAM
    post-processing `_mask_to_rle` ("in the format expected by pycoco tools": column-major runs, the first one counting zeros), a port
    of segment-anything's amg.py.  Pinned against it on blob masks, all-zero / all-one masks and speckle; the string compression
    (rleToString) stays pinned by known answers and round trips only."""
    sam = pytest.importorskip('transformers.models.sam.image_processing_sam')
    from oracle import coco_oracle
    rng = np.random.default_rng(3)
    masks = np.concatenate([blob_masks(8, 37, 53, seed=9), rng.random((4, 37, 53)) < 0.3,
                            np.zeros((1, 37, 53), bool), np.ones((1, 37, 53), bool)])
    want = sam._mask_to_rle(torch.from_numpy(masks))
    for m, w in zip(masks, want):
        got = coco_oracle.rle_encode(m.astype(np.uint8))
        assert w['size'] == [37, 53] and [int(v) for v in got] == [int(v) for v in w['counts']]
        # and through the string codec and back
        assert np.array_equal(coco_oracle.rle_from_string(coco_oracle.rle_to_string(got)), got)


def test_rle_round_trips():
    rng = np.random.default_rng(0)
    for _ in range(50):
        h, w = (int(v) for v in rng.integers(1, 60, 2))
        m = (rng.random((h, w)) < rng.random()).astype(np.uint8)
        counts = co.rle_encode(m)
        assert int(counts.sum()) == h * w
        text = co.rle_to_string(counts)
        assert np.array_equal(co.rle_from_string(text), counts)
        assert np.array_equal(co.rle_decode(counts, h, w), m)


@pytest.mark.parametrize('name', sorted(COCO_INFOS))
def test_oracle_matches_reference_golden(name):
    H, W, info = COCO_INFOS[name]
    segm, boxes_in, boxes_out = _gold(name)
    masks = blob_masks(6, H, W, seed=len(name))
    got = co.recover_shape_segm(masks, info)
    assert got.shape == segm.shape
    for k in range(got.shape[0]):
        assert _iou(got[k], segm[k]) >= 0.999                  # north star: mask IoU >= 0.999 (interpolation ties, see oracle header)
    assert np.abs(co.recover_shape_bbox(boxes_in, info) - boxes_out).max() <= 1e-4     # pixels, fp32 reassociation only


def test_oracle_resize_matches_torch_round():
    masks = blob_masks(5, 544, 544, seed=9)
    info = {'id': 0, 'height': 480, 'width': 640}
    ref = torch.nn.functional.interpolate(torch.from_numpy(masks).unsqueeze(0).float(), size=(480, 640), mode='bilinear',
                                          align_corners=False).squeeze(0).round().to(torch.uint8).numpy()
    got = co.recover_shape_segm(masks, info)
    for k in range(5):
        assert _iou(got[k], ref[k]) >= 0.999


def test_host_bbox_format_matches_oracle():
    import orienmask_b200 as ob
    cat2label = list(range(1, 81))
    m = ob.COCOMetrics(None, cat2label, with_mask=False, save_dir='.')
    infos, dets, dets_np = [], [], []
    g = torch.Generator().manual_seed(1)
    for name, (H, W, info) in sorted(COCO_INFOS.items()):
        bbox = torch.rand(4, 5, generator=g) * 0.5 + 0.2
        cls = torch.randint(0, 80, (4,), generator=g)
        infos.append(info)
        dets.append({'bbox': bbox, 'cls': cls, 'mask': None})
        dets_np.append({'bbox': bbox.numpy(), 'cls': cls.numpy()})
    dets.append({'bbox': torch.zeros(0, 5), 'cls': torch.zeros(0, dtype=torch.long), 'mask': None})      # K = 0 is skipped (:133-134)
    dets_np.append({'bbox': np.zeros((0, 5), np.float32), 'cls': np.zeros(0, np.int64)})
    infos.append({'id': 99, 'height': 10, 'width': 10})
    got = m.to_coco_format(infos, dets)['bbox']
    ref = co.to_bbox_coco_format(infos, dets_np, cat2label)
    assert len(got) == len(ref) == 20
    for a, b in zip(got, ref):
        assert a['image_id'] == b['image_id'] and a['category_id'] == b['category_id'] and a['score'] == b['score']
        assert np.allclose(a['bbox'], b['bbox'], atol=1e-4)


# ---- GPU -------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_kernel_matches_oracle_golden_cases():
    import orienmask_b200 as ob
    cat2label = list(range(1, 81))
    metrics = ob.COCOMetrics(None, cat2label, with_mask=True, save_dir='.')
    infos, dets, dets_np = [], [], []
    for name, (H, W, info) in sorted(COCO_INFOS.items()):
        masks = blob_masks(6, H, W, seed=len(name))
        g = torch.Generator().manual_seed(7)
        bbox = torch.rand(6, 5, generator=g)
        cls = torch.randint(0, 80, (6,), generator=g)
        infos.append(info)
        dets.append({'bbox': bbox.cuda(), 'cls': cls.cuda(), 'mask': torch.from_numpy(masks).cuda()})
        dets_np.append({'bbox': bbox.numpy(), 'cls': cls.numpy(), 'mask': masks})
    dets.insert(2, {'bbox': torch.zeros(0, 5).cuda(), 'cls': torch.zeros(0, dtype=torch.long).cuda(), 'mask': torch.zeros(0, 64, 96, dtype=torch.bool).cuda()})
    dets_np.insert(2, {'bbox': np.zeros((0, 5), np.float32), 'cls': np.zeros(0, np.int64), 'mask': np.zeros((0, 64, 96), bool)})
    infos.insert(2, {'id': 42, 'height': 30, 'width': 40})
    got = metrics.to_coco_format(infos, dets)
    ref = co.to_segm_coco_format(infos, dets_np, cat2label)
    assert len(got['segm']) == len(ref) == 30
    for a, b in zip(got['segm'], ref):
        assert a['image_id'] == b['image_id'] and a['category_id'] == b['category_id'] and a['score'] == b['score']
        assert a['segmentation']['size'] == b['segmentation']['size']
        assert a['segmentation']['counts'] == b['segmentation']['counts']               # bit-exact strings
    # and against the reference's own recovered masks (golden), through the decoder
    i = 0
    for name, (H, W, info) in sorted(COCO_INFOS.items()):
        segm, _, _ = _gold(name)
        rows = [r for r in got['segm'] if r['image_id'] == info['id']]
        for k, r in enumerate(rows):
            h, w = r['segmentation']['size']
            dec = co.rle_decode(co.rle_from_string(r['segmentation']['counts']), h, w)
            assert _iou(dec, segm[k]) >= 0.999
        i += 1


@pytest.mark.gpu
def test_kernel_full_size_round_trip_and_overflow_retry():
    """544x544 masks of the real post-process -> 480x640 originals; a tiny initial cap forces the retry path."""
    from orienmask_b200.coco_format import encode_masks
    masks = blob_masks(12, 544, 544, seed=4)
    noisy = np.random.default_rng(1).random((544, 544)) < 0.5
    masks[3] = noisy                                                   # tens of thousands of runs
    info = {'id': 0, 'height': 480, 'width': 640, 'collate_pad': [0, 0, 0, 0, 544, 544]}
    enc = encode_masks([torch.from_numpy(masks).cuda()], [12], [info], cap=64)[0]
    ref = co.recover_shape_segm(masks, info)
    for k in range(12):
        counts = co.rle_from_string(enc[k]['counts'])
        assert np.array_equal(counts, co.rle_encode(ref[k]))
        assert np.array_equal(co.rle_decode(counts, 480, 640), ref[k])                   # encode -> decode round trip


@pytest.mark.gpu
def test_post_process_output_feeds_formatter():
    """The post-process's own list-of-dicts result (bool mask views into the padded buffer) goes straight in."""
    import functools
    import orienmask_b200 as ob
    from tests.common import synthetic_heads, post_config
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'),
                                       **post_config(64, 96, 0.005))
    heads = [(b.cuda(), o.cuda()) for b, o in synthetic_heads(2, 64, 96, seed=5)]
    dets = post(heads)
    infos = [{'id': 7, 'height': 50, 'width': 75, 'collate_pad': [0, 0, 0, 0, 64, 96]}, {'id': 8, 'height': 64, 'width': 96}]
    out = ob.COCOMetrics(None, list(range(1, 81)), True, '.').to_coco_format(infos, dets)
    dets_np = [{k: v.cpu().numpy() for k, v in d.items()} for d in dets]
    ref = co.to_segm_coco_format(infos, dets_np, list(range(1, 81)))
    assert [r['segmentation'] for r in out['segm']] == [r['segmentation'] for r in ref]
    assert len(out['bbox']) == len(out['segm']) == sum(d['bbox'].shape[0] for d in dets)


@pytest.mark.gpu
def test_tester_loop_over_a_loader(tmp_path):
    """trainer/tester.py:26-50 on a synthetic loader: every batch goes model -> post-process -> COCO format -> results."""
    import functools
    import json
    import orienmask_b200 as ob
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    from tests.common import post_config

    class Loader(list):
        batch_size = 2
        dataset = type('D', (), {'CAT2LABEL': list(range(1, 81)), 'with_mask': True, 'CLASSES': ['c%d' % i for i in range(80)]})

    batches = Loader()
    for i in range(3):
        infos = [{'id': 10 * i + j, 'height': 48, 'width': 80, 'collate_pad': [0, 0, 0, 0, 64, 96]} for j in range(2)]
        batches.append((synthetic_images(2, 64, 96, seed=20 + i), None, infos))
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model = model.to('cuda:0')
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'),
                                       **post_config(64, 96, 0.005))
    tester = ob.Tester(model, post, batches, str(tmp_path), torch.device('cuda:0'), gt_file=None)
    timings = tester.test()
    assert timings['images'] == 6 and timings['Network Forward'] > 0
    saved = json.load(open(tmp_path / 'coco_format_results.json'))
    assert len(saved['bbox']) == len(saved['segm']) == len(tester.coco_metrics.bbox_results) > 0
    assert {r['image_id'] for r in saved['segm']} <= {10 * i + j for i in range(3) for j in range(2)}
    # the same batch through the pieces by hand gives the same records
    dets = post(model(batches[0][0].cuda()))
    ref = tester.coco_metrics.to_coco_format(batches[0][2], dets)
    n0 = len(ref['segm'])
    assert [r['segmentation'] for r in saved['segm'][:n0]] == [r['segmentation'] for r in ref['segm']]


def test_coco_eval_hands_results_to_pycocotools(tmp_path, monkeypatch):
    """COCOMetrics.coco_eval (eval/coco_eval.py:77-106,206-219) only drives pycocotools' COCOeval: checked against a
    stand-in that records the calls, since pycocotools is not in this image.  Without it the method fails loudly."""
    import types
    import numpy as np
    from orienmask_b200.coco_format import COCOMetrics
    calls = []

    class COCO:
        def __init__(self, gt):
            calls.append(('gt', gt))

        def loadRes(self, path):
            calls.append(('res', os.path.basename(path), json.load(open(path))))
            return path

    class COCOeval:
        def __init__(self, gt, res, iouType):
            self.kind = iouType
            calls.append(('eval', iouType))

        def evaluate(self):
            calls.append('evaluate')

        def accumulate(self):
            prec = -np.ones((10, 101, 2, 4, 3))
            prec[:, :, 0, 0, -1] = 0.5                      # category 0: AP 50; category 1 has no valid entry -> nan
            self.eval = {'precision': prec}

        def summarize(self):
            print('noise that must not reach stdout')
            self.stats = np.arange(12) / 10.0 + (1 if self.kind == 'segm' else 0)

    pkg = types.ModuleType('pycocotools')
    coco_mod, eval_mod = types.ModuleType('pycocotools.coco'), types.ModuleType('pycocotools.cocoeval')
    coco_mod.COCO, eval_mod.COCOeval = COCO, COCOeval
    for name, mod in (('pycocotools', pkg), ('pycocotools.coco', coco_mod), ('pycocotools.cocoeval', eval_mod)):
        monkeypatch.setitem(sys.modules, name, mod)
    m = COCOMetrics(gt_file='gt.json', cat2label=[1, 2], with_mask=True, save_dir=str(tmp_path))
    m.update_results({'bbox': [{'image_id': 1, 'category_id': 1, 'bbox': [0, 0, 1, 1], 'score': 0.5}],
                      'segm': [{'image_id': 1, 'category_id': 1, 'segmentation': {'size': [2, 2], 'counts': '04'}, 'score': 0.5}]})
    log = m.coco_eval(per_cats=True)
    assert calls[0] == ('gt', 'gt.json') and calls[1][:2] == ('res', 'bbox_prediction.json') and calls[1][2] == m.bbox_results
    assert ('eval', 'bbox') in calls and ('eval', 'segm') in calls and calls.count('evaluate') == 2
    assert log['bbox_AP'] == 0.0 and abs(log['segm_AR1'] - 1.6) < 1e-12 and len(log) == 24
    assert list(m.bbox_eval_stats) == list(np.arange(12) / 10.0) and m.segm_eval_stats[0] == 1.0
    assert m.bbox_eval_per_cats_stats[0] == 50.0 and np.isnan(m.bbox_eval_per_cats_stats[1])
    m.reset()
    assert m.bbox_results == [] and m.segm_eval_per_cats_stats == []
    for name in ('pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval'):
        monkeypatch.setitem(sys.modules, name, None)       # import now raises ImportError
    with pytest.raises(RuntimeError, match='needs pycocotools'):
        m.coco_eval()
