"""Golden fixture for the visualiser mask blend, made by the UNMODIFIED reference in this container.

    python tests/golden/make_golden_blend.py          (build container only: needs /root/reference)

Runs the reference's own ``InferenceVisualizer._recover_shape_segm`` (class method) and ``plot_all_mask``
(utils/visualizer.py:95-100,122-127; torch CPU) on seeded blob masks with distinct areas -> ``blend_small.npz``.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import build_ref  # noqa: E402
from tests.common import blob_masks  # noqa: E402


def main():
    build_ref.import_reference()
    from utils.visualizer import InferenceVisualizer, PALETTE
    masks = blob_masks(7, 64, 96, seed=3)[:5]                       # drop the empty / full masks of the generator
    pad_info = [15, 16, 8, 8, 64, 96]
    height, width = 50, 75
    g = torch.Generator().manual_seed(4)
    image = torch.rand(height, width, 3, generator=g) * 255
    colors = torch.tensor(PALETTE, dtype=torch.float32)[(torch.arange(5) * 5 + 3) % len(PALETTE)]
    vis = InferenceVisualizer.__new__(InferenceVisualizer)          # the dataset lookup of __init__ is not needed for these methods
    vis.alpha = 0.6
    soft = InferenceVisualizer._recover_shape_segm(torch.from_numpy(masks), width, height, pad_info)
    order = soft.sum(dim=2).sum(dim=1).argsort()
    out = image.clone()
    vis.plot_all_mask(soft[order], out, colors[order])
    print('areas', soft.sum(dim=2).sum(dim=1).tolist(), 'order', order.tolist(), float(out.mean()))
    np.savez_compressed(os.path.join(HERE, 'blend_small.npz'), image=image.numpy(), colors=colors.numpy(), out=out.numpy(),
                        order=order.numpy(), pad_info=np.asarray(pad_info), alpha=np.float32(0.6))


if __name__ == '__main__':
    main()
