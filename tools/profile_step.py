"""Minimal driver for ncu: `--steps` steps of the bs-32 544x544 hot path (no timing, no CPU baseline)."""
import argparse
import functools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402
from bench import post_kwargs, H, W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--precision', default='fp16')
a = ap.parse_args()
dev = torch.device('cuda:0')
model = ob.OrienMaskYOLOFPNPlus(3, 80)
model.load_state_dict(synthetic_state_dict(0), strict=True)
model.precision = a.precision
model = model.to(dev).eval()
post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev, **post_kwargs())
x = synthetic_images(a.batch, H, W, seed=1).to(dev)
for _ in range(a.steps):
    out = post.apply_padded(model(x))
torch.cuda.synchronize()
print('instances per image:', out.count.tolist()[:8])
