"""Generates the committed golden fixtures by running the UNMODIFIED reference in this container.

Run from the repo root (needs /root/reference, so build container only):
    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity is pinned by
executing its own code on seeded synthetic inputs and storing what it returns:

* small_fwd_post.npz   -- 2 images of 64x96: reference heads (forward parity) and the reference
                          post-process output for three threshold settings (top-k path, <=nms_pre
                          path, empty result), masks bit-packed.
* post_544_digest.npz  -- north-star 544x544 post-process config on seeded *synthetic head tensors*
                          (recipe in ``synthetic_heads``; cheap to regenerate anywhere): reference
                          boxes / classes / per-instance mask areas / SHA-256 of the packed masks.
* fwd_544_probe.npz    -- reference forward at 544x544 on the synthetic weights: strided sample of
                          every head tensor (pins the forward oracle at full size).
* nms_cases.npz        -- inputs/outputs of the compiled reference eval/src/nms_cpu.cpp.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402
from oracle import build_ref  # noqa: E402

from tests.common import ANCHORS, ANCHOR_MASK, synthetic_heads, post_config  # noqa: E402,F401


def run_ref_post(ref_eval, heads, height, width, conf_thresh):
    import functools
    cfg = post_config(height, width, conf_thresh)
    post = ref_eval.OrienMaskYOLOPostProcess(
        nms_func=functools.partial(ref_eval.batched_nms, threshold=0.5), device=torch.device('cpu'), **cfg)
    with torch.no_grad():
        return post(heads)


def pack(result, prefix, store_masks=True):
    d = {}
    for b, r in enumerate(result):
        m = r['mask'].numpy()
        d['%s_bbox_%d' % (prefix, b)] = r['bbox'].numpy()
        d['%s_cls_%d' % (prefix, b)] = r['cls'].numpy()
        d['%s_area_%d' % (prefix, b)] = m.sum(axis=(1, 2)).astype(np.int64)
        bits = np.packbits(m.reshape(-1))
        d['%s_sha_%d' % (prefix, b)] = np.frombuffer(hashlib.sha256(bits.tobytes()).digest(), dtype=np.uint8)
        if store_masks:
            d['%s_maskbits_%d' % (prefix, b)] = bits
            d['%s_maskshape_%d' % (prefix, b)] = np.asarray(m.shape, dtype=np.int64)
    return d


def main():
    torch.set_num_threads(8)
    config, ref_model, builder = build_ref.import_reference()
    import eval as ref_eval
    sd = synthetic_state_dict(0)
    net = ref_model.OrienMaskYOLOFPNPlus(3, 80, pretrained=None)
    net.load_state_dict(sd, strict=True)
    net.eval()

    # ---- small forward + post fixture ---------------------------------------------------------
    H, W = 64, 96
    x = synthetic_images(2, H, W, seed=1)
    with torch.no_grad():
        heads = net(x)
    d = {}
    for i, (bbox, orien) in enumerate(heads):
        d['bbox_%d' % i] = bbox.numpy()
        d['orien_%d' % i] = orien.contiguous().numpy()
    for name, thr in (('topk', 0.005), ('few', 0.02), ('none', 0.9999)):
        res = run_ref_post(ref_eval, heads, H, W, thr)
        print('small', name, [tuple(r['bbox'].shape) for r in res])
        d.update(pack(res, name))
        d[name + '_thresh'] = np.float64(thr)
    np.savez_compressed(os.path.join(HERE, 'small_fwd_post.npz'), **d)

    # ---- north-star post-process config on synthetic heads -------------------------------------
    H = W = 544
    heads = synthetic_heads(2, H, W, seed=7)
    res = run_ref_post(ref_eval, heads, H, W, 0.005)
    print('544', [tuple(r['bbox'].shape) for r in res], res[0]['bbox'][:3])
    d = pack(res, 'ns', store_masks=False)
    d['seed'] = np.int64(7)
    np.savez_compressed(os.path.join(HERE, 'post_544_digest.npz'), **d)

    # ---- forward at 544 ------------------------------------------------------------------------
    x = synthetic_images(1, H, W, seed=1)
    with torch.no_grad():
        heads = net(x)
    d = {}
    for i, (bbox, orien) in enumerate(heads):
        d['bbox_%d' % i] = bbox.numpy().reshape(-1)[::97].copy()
        d['orien_%d' % i] = orien.contiguous().numpy().reshape(-1)[::97].copy()
    np.savez_compressed(os.path.join(HERE, 'fwd_544_probe.npz'), **d)

    # ---- native NMS cases ----------------------------------------------------------------------
    nms = build_ref.load_ref_nms()
    g = torch.Generator().manual_seed(3)
    d = {}
    cases = []
    for n in (1, 2, 7, 64, 65, 400):
        ctr = torch.rand(n, 2, generator=g)
        wh = torch.rand(n, 2, generator=g) * 0.4 + 0.02
        sc = torch.rand(n, 1, generator=g)
        cases.append(torch.cat([ctr, wh, sc], 1))
    dup = cases[-1].clone()                       # exact duplicates + tied scores + IoU == threshold pairs
    dup[1::2, :4] = dup[0::2, :4]
    dup[::4, 4] = 0.5
    cases.append(dup)
    half = torch.tensor([[0.5, 0.5, 0.2, 0.2, 0.9], [0.5 + 0.2 / 3, 0.5, 0.2, 0.2, 0.8]])   # IoU = 0.5 up to rounding
    cases.append(half)
    for i, c in enumerate(cases):
        for thr in (0.5, 0.3):
            keep = nms.nms(c, thr)
            d['dets_%d' % i] = c.numpy()
            d['keep_%d_%s' % (i, str(thr).replace('.', 'p'))] = keep.numpy()
    d['n_cases'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, 'nms_cases.npz'), **d)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
