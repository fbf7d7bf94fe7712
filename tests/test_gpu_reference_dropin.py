"""The reference's OWN entry points, unmodified, on the B200 engine -- to completion.

`gpurun` ships only this repository, so `__graft_entry__.build()` copies the reference checkout into the git-ignored
`baseline/_ref/reference` in the build container (oracle/build_ref.py:ship_reference) and these tests run it from there:

  (i)   `python -m orienmask_b200.dropin infer.py -c orienmask_yolo_coco_544_anchor4_fpn_plus_infer -w <524-key ckpt> -j <images.json>
        -d assets -o <out> -b`: the reference's infer.py (its config, builder, FastCOCOTransform, pad, timers, COCO json writer) with
        this repo's model / post-process / COCOMetrics found by name; the two json files it writes are compared with the CPU oracle
        (cv2 -> pre-process oracle -> forward oracle -> post-process oracle -> COCO-format oracle) on the same images;
  (ii)  `python -m orienmask_b200.dropin test.py -c <test config json> -w <ckpt with config>` over a 2-image COCO-style set: the
        reference's build_tester / COCODataset / cv2 COCOTransform / DataLoader / Tester.test loop, compared the same way;
  (iii) the GPU "bar" timed with the reference's nn.Module itself (and its Python post-process with a labelled NMS substitute).

The comparisons run the engine in its tensor-core parity mode (ORIENMASK_B200_PRECISION=parity), at the north-star tolerances
(boxes / scores 1e-3, mask IoU 0.999 on the RLE-decoded masks); the fp16 production mode is run through the same command and its
agreement is reported.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.common import ROOT, post_config
from oracle import build_ref

REF = build_ref.reference_root()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(REF is None, reason='no reference checkout or shipped copy (baseline/_ref/reference)')]

IMAGES = ['000000163126.jpg', '000000163126_pred.jpg']


def _env(precision):
    build_ref.write_stubs()
    return dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, build_ref.STUBS]), ORIENMASK_B200_PRECISION=precision,
                PYTHONDONTWRITEBYTECODE='1')


def _image_list():
    import cv2
    out = []
    for i, name in enumerate(IMAGES):
        img = cv2.imread(os.path.join(REF, 'assets', name))
        out.append({'file_name': name, 'height': int(img.shape[0]), 'width': int(img.shape[1]), 'id': 163126 + i})
    return out


def _decode(seg):
    from oracle import coco_oracle
    h, w = seg['size']
    return coco_oracle.rle_decode(coco_oracle.rle_from_string(seg['counts']), h, w).astype(bool)


def _compare(got_bbox, got_segm, want_bbox, want_segm, sizes):
    """Match the json records of the run with the oracle's per image by category and nearest box."""
    rep = {'records_engine': len(got_bbox), 'records_oracle': len(want_bbox), 'matched': 0, 'max_box_err_rel': 0.0, 'max_score_err': 0.0,
           'min_mask_iou': 1.0, 'aggregate_mask_iou': 1.0, 'masks_off': 0, 'max_differing_pixels': 0, 'unmatched': []}
    inter_sum = union_sum = 0
    assert len(got_bbox) == len(got_segm) and len(want_bbox) == len(want_segm)
    used = set()
    for gi, g in enumerate(got_bbox):
        size = float(max(sizes[g['image_id']]))
        best, bj = 1e9, -1
        for j, w in enumerate(want_bbox):
            if j in used or w['image_id'] != g['image_id'] or w['category_id'] != g['category_id']:
                continue
            d = float(np.abs(np.asarray(w['bbox']) - np.asarray(g['bbox'])).max()) / size
            if d < best:
                best, bj = d, j
        if bj < 0 or best > 5e-3:
            rep['unmatched'].append({'image_id': g['image_id'], 'category_id': g['category_id'], 'score': g['score'], 'nearest': best})
            continue
        used.add(bj)
        rep['matched'] += 1
        rep['max_box_err_rel'] = max(rep['max_box_err_rel'], best)
        rep['max_score_err'] = max(rep['max_score_err'], abs(g['score'] - want_bbox[bj]['score']))
        assert got_segm[gi]['category_id'] == g['category_id'] and abs(got_segm[gi]['score'] - g['score']) < 1e-9
        a, b = _decode(got_segm[gi]['segmentation']), _decode(want_segm[bj]['segmentation'])
        u, it = int((a | b).sum()), int((a & b).sum())
        iou = it / u if u else 1.0
        rep['min_mask_iou'] = min(rep['min_mask_iou'], iou)
        rep['max_differing_pixels'] = max(rep['max_differing_pixels'], u - it)
        rep['masks_off'] += int(iou < 0.999 and u - it > 2)          # the mask gate of tests/common.py:e2e_agreement
        inter_sum, union_sum = inter_sum + it, union_sum + u
    rep['aggregate_mask_iou'] = inter_sum / union_sum if union_sum else 1.0
    return rep


def _oracle_records(batches):
    """batches: list of (image fp32 [n,3,H,W] numpy, [sample_info...]) -> (bbox records, segm records) from the CPU oracle."""
    from oracle import coco_oracle
    from oracle.forward_oracle import forward_oracle
    from oracle.post_oracle import PostProcessOracle
    from orienmask_b200.synthetic import synthetic_state_dict
    ref_data = build_ref  # noqa: F841
    sd = synthetic_state_dict(0)
    cat2label = CAT2LABEL
    bbox, segm = [], []
    for x, infos in batches:
        H, W = x.shape[-2:]
        cfg = post_config(H, W, 0.005)
        post = PostProcessOracle(cfg['grid_size'], cfg['image_size'], cfg['anchors'], cfg['anchor_mask'], 80, conf_thresh=0.005)
        heads = forward_oracle(sd, torch.from_numpy(x))
        dets = post([(b.numpy(), o.numpy()) for b, o in heads])
        bbox += coco_oracle.to_bbox_coco_format(infos, dets, cat2label)
        segm += coco_oracle.to_segm_coco_format(infos, dets, cat2label)
    return bbox, segm


# data/dataset.py:42-49 of the reference (COCODataset.CAT2LABEL): the 80 COCO category ids
CAT2LABEL = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32, 33, 34, 35, 36, 37, 38, 39,
             40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76,
             77, 78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90]


def _gate(rep, what):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, 'gpurun_out', 'dropin_%s.json' % what), 'w'), indent=1)
    print(json.dumps(rep))
    assert rep['records_engine'] == rep['records_oracle'], rep
    assert rep['matched'] >= rep['records_oracle'] - 4 and len(rep['unmatched']) <= 4, rep      # margin-limited kept-set flips, listed
    assert rep['max_box_err_rel'] <= 1e-3 and rep['max_score_err'] <= 1e-3, rep
    assert rep['aggregate_mask_iou'] >= 0.9995 and rep['min_mask_iou'] >= 0.995 and rep['masks_off'] <= 2, rep


def _run_infer(tmp_path, precision):
    from orienmask_b200.synthetic import synthetic_state_dict
    weights, listing, out = str(tmp_path / 'weights.pth'), str(tmp_path / 'images.json'), str(tmp_path / ('out_' + precision))
    torch.save({'state_dict': synthetic_state_dict(0)}, weights)                    # infer.py:82 accepts {'state_dict': ...}
    json.dump({'images': _image_list()}, open(listing, 'w'))
    cmd = [sys.executable, '-m', 'orienmask_b200.dropin', 'infer.py', '-c', 'orienmask_yolo_coco_544_anchor4_fpn_plus_infer',
           '-w', weights, '-j', listing, '-d', 'assets', '-o', out, '-b']
    res = subprocess.run(cmd, cwd=REF, env=_env(precision), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    assert 'The average inference time is' in res.stdout and 'Forward & Postprocess' in res.stdout, res.stdout[-1000:]
    return (json.load(open(os.path.join(out, 'bbox_prediction.json'))), json.load(open(os.path.join(out, 'segm_prediction.json'))),
            res.stdout)


def test_reference_infer_py_runs_to_completion_and_matches_the_oracle(tmp_path):
    import cv2
    from oracle.prep_oracle import fast_transform_oracle, pad_oracle
    got_bbox, got_segm, stdout = _run_infer(tmp_path, 'parity')
    batches, sizes = [], {}
    for im in _image_list():
        rgb = cv2.cvtColor(cv2.imread(os.path.join(REF, 'assets', im['file_name'])), cv2.COLOR_BGR2RGB)        # infer.py:147
        x = fast_transform_oracle(rgb[None].astype(np.float32), (544, 544))                                   # config/base.py:158-164
        x, pad_info = pad_oracle(x)                                                                           # infer.py:21-32
        batches.append((x, [{'height': im['height'], 'width': im['width'], 'id': im['id'], 'collate_pad': pad_info}]))
        sizes[im['id']] = (im['height'], im['width'])
    want_bbox, want_segm = _oracle_records(batches)
    rep = _compare(got_bbox, got_segm, want_bbox, want_segm, sizes)
    rep['speed_lines'] = [ln for ln in stdout.splitlines() if 'fps' in ln]
    _gate(rep, 'infer_py_parity')
    # the production (fp16) engine through the same unmodified command: runs to completion; agreement reported, loosely gated
    fb, fs, _ = _run_infer(tmp_path, 'fp16')
    rep16 = _compare(fb, fs, want_bbox, want_segm, sizes)
    json.dump(rep16, open(os.path.join(ROOT, 'gpurun_out', 'dropin_infer_py_fp16.json'), 'w'), indent=1)
    assert rep16['matched'] >= 0.9 * rep16['records_oracle'] and rep16['max_score_err'] < 2e-2, rep16


LOADER_DUMP = r'''
import json, sys, types, numpy as np, torch
sys.path[:0] = [%r, %r]
for name in ('eval.nms_cpu', 'eval.nms_cuda'):          # the compiled extensions the reference's eval package imports; never called here
    sys.modules[name] = types.ModuleType(name)
import data as data_module
from trainer.builder import build_dataloader
cfg = json.load(open(sys.argv[1]))
loader = build_dataloader(cfg['test_loader'])
out = []
for i, sample in enumerate(loader):
    np.save(sys.argv[2] + '/batch_%%d.npy' %% i, sample[0].numpy())
    out.append(sample[2])
json.dump(out, open(sys.argv[2] + '/infos.json', 'w'))
'''


def test_reference_test_py_runs_to_completion_and_matches_the_oracle(tmp_path):
    from orienmask_b200.synthetic import synthetic_state_dict
    env = _env('parity')
    imgs = _image_list()
    (tmp_path / 'list.txt').write_text(''.join(im['file_name'] + '\n' for im in imgs))
    (tmp_path / 'anno.json').write_text(json.dumps({im['file_name']: {'image_id': im['id'], 'anno': {'bbox': [], 'cls': [], 'mask': []}}
                                                    for im in imgs}))
    (tmp_path / 'gt.json').write_text(json.dumps({
        'images': [{'id': im['id'], 'height': im['height'], 'width': im['width'], 'file_name': im['file_name']} for im in imgs],
        'annotations': [], 'categories': [{'id': c, 'name': str(c)} for c in CAT2LABEL]}))
    code = '''
import copy, json, sys
sys.path.insert(0, %r)
import config as C
out = sys.argv[1]
cfg = copy.deepcopy(C.orienmask_yolo_coco_544_anchor4_fpn_plus_test)
cfg['gt_file'] = out + '/gt.json'
cfg['test_loader'].update(batch_size=2, num_workers=0)
cfg['test_loader']['dataset'].update(list_file=out + '/list.txt', image_dir=%r, anno_file=out + '/anno.json')
json.dump(cfg, open(out + '/test.json', 'w'))
json.dump(copy.deepcopy(C.orienmask_yolo_coco_544_anchor4_fpn_plus['model']), open(out + '/model.json', 'w'))
''' % (REF, os.path.join(REF, 'assets'))
    subprocess.check_call([sys.executable, '-c', code, str(tmp_path)], cwd='/tmp', env=env)
    model_cfg = json.load(open(str(tmp_path / 'model.json')))
    torch.save({'state_dict': synthetic_state_dict(0), 'config': {'model': model_cfg}}, str(tmp_path / 'ckpt.pth'))
    res = subprocess.run([sys.executable, '-m', 'orienmask_b200.dropin', 'test.py', '-c', str(tmp_path / 'test.json'),
                          '-w', str(tmp_path / 'ckpt.pth')], cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    assert 'Speed Statistics (batch size = 2)' in res.stdout and 'Network Forward' in res.stdout and 'COCO eval segm' in res.stdout, res.stdout[-2000:]
    got_bbox = json.load(open(str(tmp_path / 'bbox_prediction.json')))          # written next to the checkpoint (trainer/builder.py:55)
    got_segm = json.load(open(str(tmp_path / 'segm_prediction.json')))
    # the same batches from the reference's own dataset / cv2 transform / collate, through the CPU oracle
    dump = str(tmp_path / 'dump')
    os.makedirs(dump)
    subprocess.check_call([sys.executable, '-c', LOADER_DUMP % (build_ref.STUBS, REF), str(tmp_path / 'test.json'), dump], cwd='/tmp',
                          env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    infos = json.load(open(dump + '/infos.json'))
    batches = [(np.load(dump + '/batch_%d.npy' % i), info) for i, info in enumerate(infos)]
    want_bbox, want_segm = _oracle_records(batches)
    sizes = {im['id']: (im['height'], im['width']) for im in imgs}
    rep = _compare(got_bbox, got_segm, want_bbox, want_segm, sizes)
    rep['speed_lines'] = [ln for ln in res.stdout.splitlines() if 'fps' in ln]
    _gate(rep, 'test_py_parity')


BAR = r'''
import json, sys, types, time
sys.path[:0] = [%r, %r]
import torch
shim = types.ModuleType('eval.nms_cuda')
def _nms(dets, thr):          # SUBSTITUTE for the reference's nms_cuda (eval/src/nms_kernel.cu needs THC, gone from torch >= 1.11): torchvision,
    import torchvision         # same '>' rule and score-descending result
    xy, wh = dets[:, :2], dets[:, 2:4]
    return torchvision.ops.nms(torch.cat([xy - wh / 2, xy + wh / 2], 1), dets[:, 4], thr)
shim.nms = _nms
sys.modules['eval.nms_cuda'] = shim
sys.modules['eval.nms_cpu'] = types.ModuleType('eval.nms_cpu')
import config as C, model as M
from trainer.builder import build, build_postprocess
sys.path.insert(0, %r)
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
torch.backends.cudnn.benchmark = True                       # infer.py:73-74
cfg = C.orienmask_yolo_coco_544_anchor4_fpn_plus_infer
dev = torch.device('cuda:0')
net = build({**cfg['model'], 'pretrained': None}, M)
net.load_state_dict(synthetic_state_dict(0), strict=True)
net = net.to(dev).eval()
post = build_postprocess(cfg['postprocess'], device=dev)
B = int(sys.argv[1])
x = synthetic_images(B, 544, 544, seed=1).to(dev)
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
res = {'batch': B, 'what': 'the reference nn.Module (model/orienmask_yolo_fpnplus.py) and its Python post-process on this GPU, stock PyTorch eager',
       'nms': 'torchvision.ops.nms substitute for the unbuildable eval.nms_cuda'}
with torch.no_grad():
    res['forward_fp32_tf32_ms'] = timed(lambda: net(x))
    heads = net(x)
    res['postprocess_ms'] = timed(lambda: post(heads), iters=2)
    res['kept_first_image'] = int(post(heads)[0]['bbox'].shape[0])
    half = net.half()
    xh = x.half()
    res['forward_fp16_ms'] = timed(lambda: half(xh))
    cl = half.to(memory_format=torch.channels_last)
    xc = xh.contiguous(memory_format=torch.channels_last)
    res['forward_fp16_channels_last_ms'] = timed(lambda: cl(xc))
best = min(res['forward_fp32_tf32_ms'], res['forward_fp16_ms'], res['forward_fp16_channels_last_ms'])
res['images_per_s_forward_best'] = 1e3 * B / best
res['images_per_s_path_fp32'] = 1e3 * B / (res['forward_fp32_tf32_ms'] + res['postprocess_ms'])
res['images_per_s_path_best'] = 1e3 * B / (best + res['postprocess_ms'])
print(json.dumps(res))
'''


def test_reference_module_gpu_bar():
    """SURVEY §8(d) 'GPU reference beside it': the reference's own module and post-process, eager PyTorch / cuDNN on the same B200,
    next to the engine on the same weights and batch.  Never asserts speed; writes gpurun_out/reference_gpu_bar.json."""
    import functools
    import orienmask_b200 as ob
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    build_ref.write_stubs()
    B = 32
    out = subprocess.run([sys.executable, '-c', BAR % (build_ref.STUBS, REF, ROOT), str(B)], cwd='/tmp', capture_output=True, text=True,
                         timeout=1200, env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model = model.cuda().eval()
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'),
                                       **post_config(544, 544, 0.005))
    x = synthetic_images(B, 544, 544, seed=1).cuda()
    for _ in range(3):
        post(model(x))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        post(model(x))
    e1.record()
    torch.cuda.synchronize()
    res['engine_fp16_path_ms'] = e0.elapsed_time(e1) / 10
    res['engine_images_per_s_path'] = 1e3 * B / res['engine_fp16_path_ms']
    res['engine_vs_reference_path_best'] = res['engine_images_per_s_path'] / res['images_per_s_path_best']
    res['engine_vs_reference_forward_best'] = (1e3 * B / res['engine_fp16_path_ms']) / res['images_per_s_forward_best']
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'reference_gpu_bar.json'), 'w'), indent=1)
    print(json.dumps(res))
    assert res['forward_fp32_tf32_ms'] > 0 and res['kept_first_image'] > 0
