"""Minimal driver for ncu: `--steps` steps of the bs-32 544x544 hot path (no timing, no CPU baseline)."""
import argparse
import functools
import json
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
ap.add_argument('--size', type=int, default=544)
ap.add_argument('--coco', type=int, default=1, help='also run the detections -> COCO RLE kernel inside the profiled region')
ap.add_argument('--events', type=int, default=0, help='also time every layer with CUDA events over this many passes')
a = ap.parse_args()
dev = torch.device('cuda:0')
model = ob.OrienMaskYOLOFPNPlus(3, 80)
model.load_state_dict(synthetic_state_dict(0), strict=True)
model.precision = a.precision
model = model.to(dev).eval()
kw = post_kwargs()
kw['grid_size'] = [[a.size // s, a.size // s] for s in (32, 16, 8)]
kw['image_size'] = [a.size, a.size]
post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev, **kw)
# the step as infer.py / test.py run it: uint8 HWC image -> transform -> model -> post-process -> COCO RLE strings
from orienmask_b200.coco_format import encode_masks  # noqa: E402
u8 = (synthetic_images(a.batch, a.size, a.size, seed=1) * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().to(dev)
transform = ob.FastCOCOTransform([dict(type='Resize', size=(a.size, a.size)), dict(type='Normalize', mean=(0, 0, 0), std=(255, 255, 255))])
infos = [{'id': i, 'height': 480 * a.size // 544, 'width': 640 * a.size // 544, 'collate_pad': [0, 0, 0, 0, a.size, a.size]} for i in range(a.batch)]


def full_step():
    out = post.apply_padded(model(transform(u8)))
    if a.coco:
        dets = out.to_list()
        encode_masks([d['mask'] for d in dets], [int(d['bbox'].shape[0]) for d in dets], infos)
    return out


x = transform(u8)
out = full_step()            # un-profiled warm-up: builds the engine (BN folding etc. are torch ops)
torch.cuda.synchronize()
eng = next(iter(model._engines.values()))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump([{k: v for k, v in l.items() if k in ('name', 'shape', 'flops', 'bytes')} for l in eng.layers], open(os.path.join(ROOT, 'gpurun_out', 'layers.json'), 'w'))
if a.events:
    eng.time_layers(x, 2)
    json.dump(eng.time_layers(x, a.events), open(os.path.join(ROOT, 'gpurun_out', 'layer_events.json'), 'w'))
torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off
for _ in range(a.steps):
    out = full_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('instances per image:', out.count.tolist()[:8])
