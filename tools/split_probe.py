"""How exact is tcgen05's fp32 accumulation?  Raw accumulators (OM_OUT_PARTIAL: no bias, no activation) of one convolution against an
fp64 reference on the SAME operands, for growing K: (a) fp16 engine on fp16-rounded operands -- every product is exact in fp32, so
the whole error is the accumulation; (b) split-precision engine on fp32 operands.  For each: relative L2 error, the GAIN error (slope
of the error against the reference: truncation toward zero shrinks every sum by a common factor) and the residual after removing it.

    python tools/split_probe.py        (GPU)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from tests.common import run_engine_conv, torch_conv_ref  # noqa: E402

rows = []
for cin, k, hw in ((64, 1, 34), (256, 1, 34), (1024, 1, 17), (128, 3, 34), (256, 3, 34), (512, 3, 17), (1024, 3, 17)):
    g = torch.Generator().manual_seed(cin + k)
    x = torch.randn(2, cin, hw, hw, generator=g).cuda()
    x = torch.where(x > 0, x, 0.1 * x)                      # activation-like (after LeakyReLU): mostly positive
    w = (torch.randn(128, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    for prec, quant in ((1, True), (2, False)):
        got = run_engine_conv(x, w, None, 1, False, 1, precision=prec).double()
        ref = torch_conv_ref(x, w, None, 1, False, 1, quantize=quant).double()
        # torch_conv_ref returns float(); recompute in fp64 without the final rounding
        import torch.nn.functional as F
        xx, ww = (x.half().double(), w.half().double()) if quant else (x.double(), w.double())
        ref = F.conv2d(xx, ww, None, padding=k // 2)
        err = got - ref
        slope = float((err * ref).sum() / (ref * ref).sum())
        resid = err - slope * ref
        rows.append({'cin': cin, 'k': k, 'K': cin * k * k, 'mma_steps': cin * k * k // 16, 'engine': 'fp16' if prec == 1 else 'split',
                     'rel_l2': float(err.norm() / ref.norm()), 'gain_error': slope, 'gain_error_per_step_ulp24': slope / (cin * k * k / 16) * 2 ** 24,
                     'rel_l2_after_gain': float(resid.norm() / ref.norm())})
        print(json.dumps(rows[-1]))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'split_probe.json'), 'w'), indent=1)
