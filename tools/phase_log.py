"""Phase durations of the fused stem kernel (om_debug_phase_log): python tools/phase_log.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200 import _lib  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402

dev = torch.device('cuda:0')
model = ob.OrienMaskYOLOFPNPlus(3, 80)
model.load_state_dict(synthetic_state_dict(0), strict=True)
model = model.to(dev).eval()
x = synthetic_images(32, 544, 544, seed=1).to(dev)
for _ in range(4):
    model(x)
torch.cuda.synchronize()
log = torch.zeros(16, 8, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.om_debug_phase_log(_lib.ptr(log))
model(x)
torch.cuda.synchronize()
lib.om_debug_phase_log(None)
t = log.cpu().numpy().astype('float64')
names = ['wait patch', 'im2col + sync', 'MMA 1', 'epilogue 1 + sync', 'MMA 2', 'epilogue 2 + sync']
print('tile   ' + '  '.join('%18s' % n for n in names) + '      total (ns)')
for i in range(16):
    d = [t[i, k + 1] - t[i, k] for k in range(6)]
    print('%4d   ' % i + '  '.join('%18.0f' % v for v in d) + '   %10.0f' % (t[i, 6] - t[i, 0]))
