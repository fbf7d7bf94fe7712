"""GPU parity of the post-process kernels (through the C-ABI) against the oracle and the golden fixtures."""
import functools

import numpy as np
import pytest
import torch

from tests.common import GOLDEN, ANCHORS, ANCHOR_MASK, synthetic_heads, post_config, unpack_masks

pytestmark = pytest.mark.gpu


def _post(height, width, thr, **kw):
    import orienmask_b200 as ob
    cfg = post_config(height, width, thr)
    cfg.update(kw)
    return ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5),
                                       device=torch.device('cuda:0'), **cfg)


def _oracle(height, width, thr, **kw):
    from oracle.post_oracle import PostProcessOracle
    cfg = post_config(height, width, thr)
    cfg.update(kw)
    return PostProcessOracle(cfg['grid_size'], cfg['image_size'], cfg['anchors'], cfg['anchor_mask'], cfg['num_classes'],
                             conf_thresh=cfg['conf_thresh'], nms_threshold=0.5, nms_pre=cfg['nms_pre'],
                             nms_post=cfg['nms_post'], orien_thresh=cfg['orien_thresh'])


def _compare(got, ref_bbox, ref_cls, ref_mask, ordered, tol=2e-6, min_iou=0.999):
    """Kept sets must be identical (matched on cls + box + score); masks IoU >= 0.999 (north star)."""
    gb, gc, gm = got['bbox'].cpu().numpy(), got['cls'].cpu().numpy(), got['mask'].cpu().numpy()
    assert got['bbox'].dtype == torch.float32 and got['cls'].dtype == torch.int64 and got['mask'].dtype == torch.bool
    assert gb.shape == ref_bbox.shape, (gb.shape, ref_bbox.shape)
    used = np.zeros(len(ref_bbox), dtype=bool)
    perm = []
    for i in range(len(gb)):
        d = np.abs(ref_bbox - gb[i]).max(1) + (ref_cls != gc[i]) * 1e3 + used * 1e3
        j = int(np.argmin(d))
        assert d[j] <= tol, 'detection %d has no counterpart (best distance %g)' % (i, d[j])
        used[j] = True
        perm.append(j)
    perm = np.asarray(perm, dtype=np.int64)
    if ordered and len(perm):
        assert np.array_equal(perm, np.arange(len(perm))), 'output order differs from the reference'
    if ref_mask is not None and len(perm):
        rm = ref_mask[perm]
        inter = (rm & gm).reshape(len(perm), -1).sum(1).astype(np.float64)
        union = (rm | gm).reshape(len(perm), -1).sum(1).astype(np.float64)
        iou = np.where(union > 0, inter / np.maximum(union, 1), 1.0)
        assert iou.min() >= min_iou, 'mask IoU %g' % iou.min()
    return perm


def test_nms_matches_reference_cases():
    import orienmask_b200 as ob
    g = np.load(GOLDEN + '/nms_cases.npz')
    for i in range(int(g['n_cases'])):
        dets = torch.from_numpy(g['dets_%d' % i]).cuda()
        cats = torch.zeros(dets.shape[0], dtype=torch.long, device='cuda')
        for thr in (0.5, 0.3):
            _, _, keep = ob.nms(dets, cats, thr)
            ref = g['keep_%d_%s' % (i, str(thr).replace('.', 'p'))]
            assert keep.dtype == torch.int64
            assert np.array_equal(keep.cpu().numpy(), ref), (i, thr)


def test_batched_nms_matches_oracle():
    import orienmask_b200 as ob
    from oracle.post_oracle import batched_nms_oracle
    g = torch.Generator().manual_seed(5)
    for n in (0, 1, 33, 400, 1000):
        dets = torch.cat([torch.rand(n, 2, generator=g), torch.rand(n, 2, generator=g) * 0.3 + 0.02,
                          torch.rand(n, 1, generator=g)], 1)
        cats = torch.randint(0, 4, (n,), generator=g)
        d, c, keep = ob.batched_nms(dets.cuda(), cats.cuda(), threshold=0.5)
        ref = batched_nms_oracle(dets.numpy(), cats.numpy(), 0.5)
        assert np.array_equal(keep.cpu().numpy(), ref), n
        assert torch.equal(d.cpu(), dets[keep.cpu()]) and torch.equal(c.cpu(), cats[keep.cpu()])


@pytest.mark.parametrize('name', ['topk', 'few', 'none'])
def test_post_small_golden(name):
    g = np.load(GOLDEN + '/small_fwd_post.npz')
    post = _post(64, 96, float(g[name + '_thresh']))
    heads = [(torch.from_numpy(g['bbox_%d' % i]).cuda(), torch.from_numpy(g['orien_%d' % i]).cuda()) for i in range(3)]
    res = post(heads)
    assert isinstance(res, list) and len(res) == 2
    for b, r in enumerate(res):
        rm = unpack_masks(g, name, b) if name != 'none' else None
        _compare(r, g['%s_bbox_%d' % (name, b)], g['%s_cls_%d' % (name, b)], rm, ordered=True)


def test_post_north_star_config_vs_oracle_and_digest():
    g = np.load(GOLDEN + '/post_544_digest.npz')
    heads = synthetic_heads(2, 544, 544, seed=int(g['seed']))
    ref = _oracle(544, 544, 0.005)([(b.numpy(), o.numpy()) for b, o in heads])
    res = _post(544, 544, 0.005)([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], g['ns_bbox_%d' % b], g['ns_cls_%d' % b], None, ordered=True)
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)
        area = res[b]['mask'].sum(dim=(1, 2)).cpu().numpy()
        assert np.abs(area - g['ns_area_%d' % b]).max() <= 2


def test_post_strided_views_and_batch_of_one():
    """The model hands over channel-split views of one [B,18,h,w] tensor (torch.split, fpnplus.py:88)."""
    heads = synthetic_heads(3, 96, 64, seed=9)
    o = torch.cat([h[1] for h in heads], 1).cuda()
    views = torch.split(o, 6, dim=1)
    a = _post(96, 64, 0.005)([(heads[i][0].cuda(), views[i]) for i in range(3)])
    ref = _oracle(96, 64, 0.005)([(b.numpy(), o.numpy()) for b, o in heads])
    for b in range(3):
        _compare(a[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)
    one = _post(96, 64, 0.005)([(h[0][1:2].cuda(), h[1][1:2].cuda()) for h in heads])
    _compare(one[0], ref[1]['bbox'], ref[1]['cls'], ref[1]['mask'], ordered=True)


def test_post_960_config():
    """BASELINE config 5: 960x960 input, stride-4 map 240x240 (bs 1 here; oracle stays in seconds)."""
    heads = synthetic_heads(1, 960, 960, seed=13)
    ref = _oracle(960, 960, 0.005)([(b.numpy(), o.numpy()) for b, o in heads])
    res = _post(960, 960, 0.005)([(b.cuda(), o.cuda()) for b, o in heads])
    _compare(res[0], ref[0]['bbox'], ref[0]['cls'], ref[0]['mask'], ordered=True)


def test_post_nonsquare_small_limits():
    heads = synthetic_heads(2, 64, 128, seed=21)
    kw = dict(nms_pre=64, nms_post=10)
    ref = _oracle(64, 128, 0.01, **kw)([(b.numpy(), o.numpy()) for b, o in heads])
    res = _post(64, 128, 0.01, **kw)([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)


@pytest.mark.parametrize('thr', [1e-6, 0.05, 0.5, 0.97])
def test_post_conf_thresholds_radix_digit_shapes(thr):
    """The selection cuts its radix digits from bits(score) - bits(conf_thresh): 28 significant bits at threshold 1e-6
    (8 + 11 + 9; the library rejects thresholds <= 0), 26 at the north-star 0.005, a handful just below 1 -- every shape must select like the oracle."""
    heads = synthetic_heads(2, 64, 96, seed=31)
    ref = _oracle(64, 96, thr)([(b.numpy(), o.numpy()) for b, o in heads])
    res = _post(64, 96, thr)([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)


def test_post_all_tied_scores():
    """Every (prediction, class) score identical: selection among exact ties is unspecified in the
    reference (torch.topk); the kernel must still return exactly nms_pre candidates and a legal result."""
    heads = [(torch.zeros_like(b).cuda(), torch.zeros_like(o).cuda()) for b, o in synthetic_heads(1, 64, 64, seed=1)]
    post = _post(64, 64, 0.005)
    padded = post.apply_padded(heads)
    assert int(padded.candidates['count'][0]) == 400
    assert torch.all(padded.candidates['det'][0, :, 4] == 0.25)
    assert 0 < int(padded.count[0]) <= 100


def test_post_rejects_cpu_tensors():
    heads = synthetic_heads(1, 64, 64, seed=1)
    with pytest.raises(RuntimeError):
        _post(64, 64, 0.005)(heads)


def test_nms_kernel_writes_the_gather_records():
    """PaddedDetections.packed (written by the NMS kernel) is exactly what sharding.pack_records builds from det / cls / count."""
    import functools
    import orienmask_b200 as ob
    from orienmask_b200.sharding import pack_records, unpack_records
    from tests.common import synthetic_heads, post_config
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'),
                                       **post_config(64, 96, 0.02))
    out = post.apply_padded([(b.cuda(), o.cuda()) for b, o in synthetic_heads(3, 64, 96, seed=11)])
    assert torch.equal(out.packed, pack_records(out.det, out.cls, out.count))
    det, cls, cnt = unpack_records(out.packed, post.nms_post)
    assert torch.equal(det, out.det) and torch.equal(cls, out.cls) and torch.equal(cnt, out.count)


def test_nms_cuda_semantics_mode():
    """OM_NMS_CUDA (eval/src/nms_kernel.cu: '>' suppression, w*h areas, score-descending result) through nms(), batched_nms() and the
    whole post-process, against the oracle's restatement of that file; exact-threshold ties separate the two variants."""
    import orienmask_b200 as ob
    from oracle.post_oracle import nms_oracle, batched_nms_oracle
    g = torch.Generator().manual_seed(7)
    for n in (1, 33, 400, 1000):
        dets = torch.cat([torch.rand(n, 2, generator=g), torch.rand(n, 2, generator=g) * 0.3 + 0.02, torch.rand(n, 1, generator=g)], 1)
        cats = torch.randint(0, 4, (n,), generator=g)
        _, _, keep = ob.nms(dets.cuda(), cats.cuda(), 0.5, semantics='cuda')
        assert np.array_equal(keep.cpu().numpy(), nms_oracle(dets.numpy(), 0.5, semantics='cuda')), n
        d, c, keep = ob.batched_nms(dets.cuda(), cats.cuda(), threshold=0.5, semantics='cuda')
        ref = batched_nms_oracle(dets.numpy(), cats.numpy(), 0.5, semantics='cuda')
        assert np.array_equal(keep.cpu().numpy(), ref) and torch.equal(d.cpu(), dets[keep.cpu()])
    # an exact tie with the threshold: two unit squares overlapping by one third -> IoU = (1/3) / (2 - 1/3) = 0.2 exactly representable? use 0.5:
    # boxes [0,0,1,1] and [0.5... ] -> choose w=h=1 boxes shifted by 1/3 in x: inter 2/3, union 4/3 -> IoU exactly 0.5
    tie = torch.tensor([[0.5, 0.5, 1.0, 1.0, 0.9], [0.5 + 1.0 / 3.0, 0.5, 1.0, 1.0, 0.8]])
    iou_ref = nms_oracle(tie.numpy(), 0.5, semantics='cuda'), nms_oracle(tie.numpy(), 0.5, semantics='cpu')
    _, _, k_cuda = ob.nms(tie.cuda(), torch.zeros(2, dtype=torch.long).cuda(), 0.5, semantics='cuda')
    _, _, k_cpu = ob.nms(tie.cuda(), torch.zeros(2, dtype=torch.long).cuda(), 0.5, semantics='cpu')
    assert np.array_equal(k_cuda.cpu().numpy(), iou_ref[0]) and np.array_equal(k_cpu.cpu().numpy(), iou_ref[1])
    # whole post-process
    import functools
    heads = synthetic_heads(2, 64, 96, seed=5)
    cfg = post_config(64, 96, 0.005)
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5, semantics='cuda'), device=torch.device('cuda:0'), **cfg)
    from oracle.post_oracle import PostProcessOracle
    ref = PostProcessOracle(cfg['grid_size'], cfg['image_size'], cfg['anchors'], cfg['anchor_mask'], 80, conf_thresh=0.005, nms_semantics='cuda')(
        [(b.numpy(), o.numpy()) for b, o in heads])
    res = post([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)
