"""In-situ timeline of the pipelined forward (om_debug_trace): every conv-engine launch stamps %globaltimer at its first CTA start,
when its dependencies resolved (griddepcontrol.wait returned), and at its last CTA end.  Unlike ncu's serialised launch list this
is the forward as it really runs (PDL overlap, warm L2, sustained clocks).

    python tools/timeline.py [--batch 32] [--size 544] [--passes 10] [--md gpurun_out/timeline.md]

Per layer: busy = last end - dependencies resolved; gap = dependencies resolved - previous layer's last end (the kernel boundary:
flush + wait latency; negative never happens on one stream); lead = how long before its dependencies the first CTA was resident.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200 import _lib  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=544)
    ap.add_argument('--passes', type=int, default=10)
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--md', default=os.path.join(ROOT, 'gpurun_out', 'timeline.md'))
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.precision = a.precision
    model = model.to(dev).eval()
    x = synthetic_images(a.batch, a.size, a.size, seed=1).to(dev)
    for _ in range(8):
        model(x)
    torch.cuda.synchronize()
    eng = next(iter(model._engines.values()))
    layers = eng.layers
    n = len(layers)
    lib = eng.lib
    rec = torch.empty(a.passes * n, 4, dtype=torch.int64, device=dev)
    rec[:, 0:2] = -1          # ~0 as uint64
    rec[:, 2:4] = 0
    torch.cuda.synchronize()
    lib.om_debug_trace(_lib.ptr(rec), a.passes * n)
    for _ in range(a.passes):
        model(x)
    torch.cuda.synchronize()
    used = lib.om_debug_trace(None, 0)
    assert used == a.passes * n, (used, a.passes, n)
    r = rec.cpu().numpy().astype('uint64').reshape(a.passes, n, 4).astype('float64')
    out = []
    out.append('In-situ timeline, bs %d, %dx%d, %s engine: mean over %d forwards (us)' % (a.batch, a.size, a.size, a.precision, a.passes))
    out.append('')
    out.append('| # | layer | shape | busy | gap before | lead | GFLOP | MB | floor | busy/floor |')
    out.append('|---|---|---|---|---|---|---|---|---|---|')
    tot_busy = tot_gap = 0.0
    rows = []
    for i, L in enumerate(layers):
        busy = (r[:, i, 2] - r[:, i, 1]).mean() / 1e3
        gap = (r[:, i, 1] - r[:, i - 1, 2]).mean() / 1e3 if i else 0.0
        lead = (r[:, i, 1] - r[:, i, 0]).mean() / 1e3
        tc = L['flops'] / 1590e12 * 1e6
        tm = L['bytes'] / 6650e9 * 1e6
        floor = max(tc, tm)
        tot_busy += busy; tot_gap += gap
        rows.append(dict(index=i, name=L['name'], shape=L['shape'], busy_us=busy, gap_us=gap, lead_us=lead, floor_us=floor))
        out.append('| %d | %s | %s | %.1f | %.1f | %.1f | %.2f | %.1f | %.1f | %.2f |' % (
            i, L['name'], L['shape'], busy, gap, lead, L['flops'] / 1e9, L['bytes'] / 1e6, floor, busy / floor))
    span = (r[:, n - 1, 2] - r[:, 0, 1]).mean() / 1e3
    out.append('')
    out.append('forward span (first dependency resolved -> last end): %.1f us; sum busy %.1f, sum gaps %.1f' % (span, tot_busy, tot_gap))
    text = '\n'.join(out)
    print(text)
    os.makedirs(os.path.dirname(a.md), exist_ok=True)
    open(a.md, 'w').write(text + '\n')
    json.dump(rows, open(a.md.replace('.md', '.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
