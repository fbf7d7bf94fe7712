"""Config 2: bs 1, 544x544, fp16 engine -- latency of forward + post-process, eager launches vs CUDA-graph replay.

    python tools/latency.py [--iters 200] [--batch 1]

Median / p90 over `--iters` iterations after 20 warm-ups, CUDA events on the launching stream; prints one JSON line.
"""
import argparse
import functools
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402
from bench import post_kwargs, H, W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--iters', type=int, default=200)
ap.add_argument('--batch', type=int, default=1)
a = ap.parse_args()
dev = torch.device('cuda:0')
post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev, **post_kwargs())
x = synthetic_images(a.batch, H, W, seed=1).to(dev)
out = {}
for mode in ('eager', 'graph'):
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.use_cuda_graph = mode == 'graph'
    model = model.to(dev).eval()
    for _ in range(20):
        post.apply_padded(model(x))
    torch.cuda.synchronize()
    fwd, tot = [], []
    for _ in range(a.iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        heads = model(x)
        e[1].record()
        post.apply_padded(heads)
        e[2].record()
        torch.cuda.synchronize()
        fwd.append(e[0].elapsed_time(e[1]))
        tot.append(e[0].elapsed_time(e[2]))
    out[mode] = {'forward_ms_median': statistics.median(fwd), 'total_ms_median': statistics.median(tot),
                 'total_ms_p90': sorted(tot)[int(0.9 * len(tot))], 'images_per_s': 1e3 * a.batch / statistics.median(tot)}
print(json.dumps({'config': 'bs=%d 544x544 fp16 forward + post-process latency' % a.batch, 'iters': a.iters, **out}))
