"""Shapes and configurations beyond the north-star one: widths whose stride-16 map is a multiple of 8 wide (640, 1024: planned without the
halo since the planner sweep, DESIGN finding 24), non-square inputs, batch 64, 20 classes (75 head channels: N = 96 pair tiles) in the
model and in the post-process, two anchors per scale, the parity engine at other sizes.  (Written GPU-less at the end of round 1 as
opt-in "queued" tests; all of them passed on their first GPU visit in round 2 and are regular `-m gpu` tests now.)"""
import functools
import os

import numpy as np
import pytest
import torch

from tests.common import ANCHORS, ANCHOR_MASK, synthetic_heads, post_config
from tests.test_gpu_post import _compare

pytestmark = pytest.mark.gpu


def _model(precision):
    import orienmask_b200 as ob
    from orienmask_b200.synthetic import synthetic_state_dict
    m = ob.OrienMaskYOLOFPNPlus(3, 80)
    m.load_state_dict(synthetic_state_dict(0), strict=True)
    m.precision = precision
    return m.to('cuda:0').eval()


@pytest.mark.parametrize('shape', [(1, 640, 640), (2, 544, 640), (1, 1024, 416)])
def test_forward_at_widths_that_only_plan_since_the_sweep(shape):
    """Stride-16 maps a multiple of 8 wide: the 256->512 block 3x3 layers are re-planned without the halo (finding 24a)."""
    from oracle.forward_oracle import forward_oracle
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    B, H, W = shape
    x = synthetic_images(B, H, W, seed=3)
    ref = forward_oracle(synthetic_state_dict(0), x)
    for prec in ('fp32', 'fp16'):
        out = _model(prec)(x.cuda())
        for (b, o), (rb, ro) in zip(out, ref):
            for got, want in ((b, rb), (o, ro)):
                got = got.float().cpu()
                if prec == 'fp32':
                    assert float((got - want).abs().max()) < 1e-3
                else:
                    assert float((got - want).norm() / want.norm()) < 0.03


def test_forward_large_batch_matches_small_batch():
    """bs 64 at 544x544: conv4.0 must stay on flat tiles (finding 24b).  Images are independent, so the first two images of
    the bs-64 forward must reproduce the bs-2 forward of the same images bit for bit (same kernels, same reduction order)."""
    from orienmask_b200.synthetic import synthetic_images
    m = _model('fp16')
    x = synthetic_images(2, 544, 544, seed=5).cuda()
    small = [(b.clone(), o.clone()) for b, o in m(x)]
    big = m(x.repeat(32, 1, 1, 1))
    for (b, o), (sb, so) in zip(big, small):
        assert torch.equal(b[:2], sb) and torch.equal(o[:2], so)
        assert torch.equal(b[62:], sb) and torch.equal(o[62:], so)


def test_post_process_with_20_classes():
    """The reference's VOC dataset class: 75 head channels per scale -- the tail of the 8-class vector loop."""
    import orienmask_b200 as ob
    from oracle.post_oracle import PostProcessOracle
    heads = synthetic_heads(2, 64, 96, seed=41, num_classes=20)
    cfg = dict(post_config(64, 96, 0.005), num_classes=20)
    ref = PostProcessOracle(cfg['grid_size'], cfg['image_size'], ANCHORS, ANCHOR_MASK, 20, conf_thresh=0.005)(
        [(b.numpy(), o.numpy()) for b, o in heads])
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'), **cfg)
    res = post([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)


def test_post_process_two_anchors_per_scale():
    import orienmask_b200 as ob
    from oracle.post_oracle import PostProcessOracle
    g = torch.Generator().manual_seed(43)
    H, W, C = 64, 64, 80
    anchors, mask = ANCHORS[:6], [[4, 5], [2, 3], [0, 1]]
    heads = []
    for s in (32, 16, 8):
        bbox = torch.randn(2, 2, 5 + C, H // s, W // s, generator=g) * 1.9
        bbox[:, :, 4] -= 4.0
        bbox[:, :, 5:] -= 2.0
        bbox[:, :, 2:4] *= 0.3
        heads.append((bbox.view(2, -1, H // s, W // s).contiguous(), torch.randn(2, 4, H // 4, W // 4, generator=g)))
    cfg = dict(grid_size=[[H // s, W // s] for s in (32, 16, 8)], image_size=[H, W], anchors=anchors, anchor_mask=mask, num_classes=C,
               conf_thresh=0.005, nms_pre=400, nms_post=100, orien_thresh=0.3)
    ref = PostProcessOracle(cfg['grid_size'], cfg['image_size'], anchors, mask, C, conf_thresh=0.005)([(b.numpy(), o.numpy()) for b, o in heads])
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=torch.device('cuda:0'), **cfg)
    res = post([(b.cuda(), o.cuda()) for b, o in heads])
    for b in range(2):
        _compare(res[b], ref[b]['bbox'], ref[b]['cls'], ref[b]['mask'], ordered=True)


def test_model_with_20_classes_all_engines():
    """num_classes=20 (the reference's VOC configs): 3 * 25 = 75 head channels -> N = 96 pair tiles in the tcgen05 engines."""
    import orienmask_b200 as ob
    from oracle.forward_oracle import forward_oracle
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    sd = synthetic_state_dict(0, num_classes=20)
    x = synthetic_images(2, 96, 160, seed=7)
    ref = forward_oracle(sd, x)
    for prec in ('fp32', 'parity', 'fp16'):
        m = ob.OrienMaskYOLOFPNPlus(3, 20)
        m.load_state_dict(sd, strict=True)
        m.precision = prec
        out = m.to('cuda:0').eval()(x.cuda())
        assert out[0][0].shape == (2, 75, 3, 5)
        for (b, o), (rb, ro) in zip(out, ref):
            for got, want in ((b, rb), (o, ro)):
                got = got.float().cpu()
                if prec == 'fp16':
                    assert float((got - want).norm() / want.norm()) < 0.03
                else:
                    assert float((got - want).abs().max()) < 1e-3


@pytest.mark.parametrize('shape', [(1, 640, 640), (2, 544, 640)])
def test_parity_forward_at_other_sizes(shape):
    from oracle.forward_oracle import forward_oracle
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    B, H, W = shape
    x = synthetic_images(B, H, W, seed=3)
    ref = forward_oracle(synthetic_state_dict(0), x)
    out = _model('parity')(x.cuda())
    for (b, o), (rb, ro) in zip(out, ref):
        for got, want in ((b, rb), (o, ro)):
            assert float((got.float().cpu() - want).abs().max()) < 1e-3


@pytest.mark.parametrize('shape', [(3, 96, 160), (1, 544, 544), (5, 160, 96), (2, 544, 640), (1, 1088, 1920)])
def test_c_engine_matches_the_python_schedule_at_other_shapes(shape, monkeypatch):
    """The C library's engine (fused first-stage block, TMA-fed stem, side-stream lanes at small batches, narrow N tiles) against the
    Python-scheduled twin (one plain launch per layer, one stream): bit-identical heads at odd batches, non-square and large inputs."""
    import orienmask_b200 as ob
    from orienmask_b200.synthetic import synthetic_images, synthetic_state_dict
    B, H, W = shape
    x = synthetic_images(B, H, W, seed=13).cuda()
    outs = []
    for engine in ('c', 'py'):
        monkeypatch.setenv('ORIENMASK_B200_ENGINE', engine)
        m = ob.OrienMaskYOLOFPNPlus(3, 80)
        m.load_state_dict(synthetic_state_dict(0), strict=True)
        m = m.to('cuda:0').eval()
        outs.append(m(x))
        outs.append(m(x))                          # a second call on the same engine: buffers and lanes are reused
        del m
        torch.cuda.synchronize()
    for a, b in ((0, 2), (1, 3), (0, 1)):
        for (b0, o0), (b1, o1) in zip(outs[a], outs[b]):
            assert torch.equal(b0, b1) and torch.equal(o0, o1), (shape, a, b)
