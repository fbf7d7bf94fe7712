"""In-situ timeline of the pipelined forward (om_debug_trace): every conv-engine launch stamps %globaltimer at its first CTA start,
when its dependencies resolved (griddepcontrol.wait returned), and at its last CTA end.  Unlike ncu's serialised launch list this
is the forward as it really runs (PDL overlap, warm L2, sustained clocks).

    python tools/timeline.py [--batch 32] [--size 544] [--passes 10] [--md gpurun_out/timeline.md]
    python tools/timeline.py --variants base: direct:ORIENMASK_B200_RESDIRECT=1     # A/B: one engine per variant (planner environment
                                                                                    # set while it is built), forwards interleaved

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
    ap.add_argument('--variants', nargs='*', default=None, help='name:ENV=V,ENV=V ...')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    x = synthetic_images(a.batch, a.size, a.size, seed=1).to(dev)
    if a.variants:
        return variants(a, dev, x)
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.precision = a.precision
    model = model.to(dev).eval()
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


def variants(a, dev, x):
    """A/B inside one process: the clocks of a power-capped part wander by several percent between runs, so the variants' forwards are
    interleaved and every family is also shown relative to the first variant."""
    sd = synthetic_state_dict(0)
    models, names = [], []
    for spec in a.variants:
        name, _, envs = spec.partition(':')
        kv = [e.split('=', 1) for e in envs.split(',') if e]
        for k, v in kv:
            os.environ[k] = v
        m = ob.OrienMaskYOLOFPNPlus(3, 80)
        m.load_state_dict(sd, strict=True)
        m.precision = a.precision
        m = m.to(dev).eval()
        for _ in range(4):
            m(x)
        torch.cuda.synchronize()
        for k, _ in kv:
            del os.environ[k]
        models.append(m); names.append(name)
    engs = [next(iter(m._engines.values())) for m in models]
    lib = engs[0].lib
    ns = [len(e.layers) for e in engs]
    total = a.passes * sum(ns)
    rec = torch.empty(total, 4, dtype=torch.int64, device=dev)
    rec[:, 0:2] = -1
    rec[:, 2:4] = 0
    for _ in range(3):
        for m in models:
            m(x)
    torch.cuda.synchronize()
    lib.om_debug_trace(_lib.ptr(rec), total)
    for _ in range(a.passes):
        for m in models:
            m(x)
    torch.cuda.synchronize()
    used = lib.om_debug_trace(None, 0)
    assert used == total, (used, total)
    r = rec.cpu().numpy().astype('uint64').astype('float64').reshape(a.passes, sum(ns), 4)
    fams = {}
    spans = []
    off = 0
    for v, e in enumerate(engs):
        rv = r[:, off:off + ns[v], :]
        off += ns[v]
        spans.append((rv[:, -1, 2] - rv[:, 0, 1]).mean() / 1e3)
        for i, L in enumerate(e.layers):
            busy = (rv[:, i, 2] - rv[:, i, 1]).mean() / 1e3
            fam = L['shape'] if not L['shape'].startswith('1x1 64->32 +') else 'fused block'
            f = fams.setdefault(fam, [[0, 0.0] for _ in engs])
            f[v][0] += 1; f[v][1] += busy
    out = ['A/B in-situ timeline (busy us per layer family, forwards interleaved), bs %d, %dx%d, %s' % (a.batch, a.size, a.size, a.precision), '']
    out.append('| family | ' + ' | '.join(names) + ' |')
    out.append('|---|' + '---|' * len(names))
    for fam, f in sorted(fams.items(), key=lambda kv: -kv[1][0][1]):
        cells = []
        for v in range(len(engs)):
            c = 'x%d %.1f' % (f[v][0], f[v][1])
            if v and f[0][1] > 0:
                c += ' (%+.1f%%)' % (100.0 * (f[v][1] - f[0][1]) / f[0][1])
            cells.append(c)
        out.append('| %s | %s |' % (fam, ' | '.join(cells)))
    out.append('| forward span | ' + ' | '.join('%.1f' % s for s in spans) + ' |')
    text = '\n'.join(out)
    print(text)
    os.makedirs(os.path.dirname(a.md), exist_ok=True)
    open(a.md, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
