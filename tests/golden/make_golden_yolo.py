"""Golden fixture for the non-Plus model, made by the UNMODIFIED reference in this container.

    python tests/golden/make_golden_yolo.py          (build container only: needs /root/reference)

The reference's own ``model.OrienMaskYOLO`` (model/orienmask_yolo.py) with the seeded synthetic state dict
(``synthetic_state_dict(0, plus=False)``, loaded strict=True) on two 64x96 images -> ``yolo_small_fwd.npz``.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402
from oracle import build_ref  # noqa: E402


def main():
    torch.set_num_threads(8)
    _, ref_model, _ = build_ref.import_reference()
    net = ref_model.OrienMaskYOLO(3, 80, pretrained=None)
    missing = net.load_state_dict(synthetic_state_dict(0, plus=False), strict=True)
    print('load_state_dict:', missing, 'keys', len(net.state_dict()))
    net.eval()
    with torch.no_grad():
        heads = net(synthetic_images(2, 64, 96, seed=1))
    d = {'n_keys': np.int64(len(net.state_dict()))}
    for i, (bbox, orien) in enumerate(heads):
        d['bbox_%d' % i] = bbox.numpy()
        d['orien_%d' % i] = orien.contiguous().numpy()
        print(i, tuple(bbox.shape), tuple(orien.shape), float(bbox.std()), float(orien.std()))
    np.savez_compressed(os.path.join(HERE, 'yolo_small_fwd.npz'), **d)


if __name__ == '__main__':
    main()
