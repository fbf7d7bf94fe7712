"""Seeded synthetic weights with the reference's state-dict layout.

There is no network access for the trained checkpoint (reference README.md:17), so tests and the
bench run on random weights.  Plain default init makes every class score 0.2617 +- 5e-6 (SURVEY §7
"hard parts"), which turns every comparison into a tie-break; this generator instead keeps
activations at unit scale analytically (He-normal convs for LeakyReLU(0.1), damped residual
branches, non-trivial BN statistics so BN folding is exercised) and biases the heads so that a
realistic few-hundred-thousand (prediction, class) pairs clear ``conf_thresh`` and the kept
detections have well separated scores.
"""
import math

import torch

from .arch import conv_specs


# final-conv gains measured once on uniform-noise 544x544 images so that box logits have std ~2 and
# orientation outputs std ~1 at every scale (keys: stride of the scale the rows belong to)
HEAD_GAIN = {32: 2.4, 16: 3.8, 8: 4.2}
ORIEN_GAIN = {32: 7.0, 16: 3.7, 8: 3.2}
TRUNK_GAIN = 0.9
RESIDUAL_GAMMA = 0.2


def synthetic_state_dict(seed=0, num_anchors=3, num_classes=80, obj_bias=-4.0, cls_bias=-2.0,
                         dtype=torch.float32, plus=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    slope = 0.1
    for s in conv_specs(num_anchors, num_classes, plus):
        fan_in = s.cin * s.k * s.k
        if s.kind == 'cbl':
            std = TRUNK_GAIN * math.sqrt(2.0 / ((1.0 + slope * slope) * fan_in))
            w = torch.randn(s.cout, s.cin, s.k, s.k, generator=g) * std
            gamma = 1.0 + 0.1 * torch.randn(s.cout, generator=g)
            if s.prefix.endswith('.conv.1'):          # residual branch output: keep the trunk at unit scale
                gamma = gamma * RESIDUAL_GAMMA
            beta = 0.05 * torch.randn(s.cout, generator=g)
            mean = 0.05 * torch.randn(s.cout, generator=g)
            var = 0.8 + 0.4 * torch.rand(s.cout, generator=g)
            sd[s.prefix + '.conv_block.0.weight'] = w
            sd[s.prefix + '.conv_block.1.weight'] = gamma
            sd[s.prefix + '.conv_block.1.bias'] = beta
            sd[s.prefix + '.conv_block.1.running_mean'] = mean
            sd[s.prefix + '.conv_block.1.running_var'] = var
            sd[s.prefix + '.conv_block.1.num_batches_tracked'] = torch.tensor(1, dtype=torch.long)
        else:
            is_orien = s.prefix.startswith('orien_head')
            w = torch.randn(s.cout, s.cin, s.k, s.k, generator=g) / math.sqrt(fan_in)
            if is_orien:                              # rows [0:2A) -> stride 32, [2A:4A) -> 16, [4A:6A) -> 8
                rows = 2 * num_anchors
                for j, st in enumerate((32, 16, 8)):
                    w[j * rows:(j + 1) * rows] *= ORIEN_GAIN[st]
            else:
                w *= HEAD_GAIN[int(s.prefix[len('bbox_head'):-2])]
            b = 0.1 * torch.randn(s.cout, generator=g)
            if not is_orien:
                bv = b.view(num_anchors, 5 + num_classes)
                bv[:, 4] += obj_bias
                bv[:, 5:] += cls_bias
            sd[s.prefix + '.weight'] = w
            sd[s.prefix + '.bias'] = b
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def synthetic_images(batch, height, width, seed=1, device='cpu'):
    """Uniform [0,1) images like the reference's /255 inputs (config/base.py:158-164)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, height, width, generator=g).to(device)
