"""In-situ cost of groups of layers: the forward timed (CUDA events, launches back to back on one stream, PDL on) with and without
each group.  ncu's per-launch times are serialised and cold-cache; this is what a group costs INSIDE the pipelined forward.

    python tools/insitu.py [--batch 32] [--size 544] [--iters 20] [--json gpurun_out/insitu.json]

Skipped layers leave the previous pass's (realistic) data in their output buffers, so the remaining launches do the same work.

Caveat (measured): on a power-capped part the clocks move when a group is removed, so marginal costs of small groups are noisy
(single layers: +-100 us on a 6 ms forward).  tools/timeline.py stamps every launch in the unmodified forward and is the instrument
of record; this tool only answers "what would the forward cost without this group".
"""
import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200 import _lib  # noqa: E402
from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402

GROUPS = [
    ('stem', r'^backbone\.conv1$'),
    ('conv2.0', r'^backbone\.conv2\.0$'),
    ('conv2.1 block', r'^backbone\.conv2\.1'),
    ('conv3.0', r'^backbone\.conv3\.0$'),
    ('conv3.x blocks', r'^backbone\.conv3\.[12]'),
    ('68 1x1 (backbone)', r'^backbone\.conv4\.\d\.conv\.0$'),
    ('68 3x3+res', r'^backbone\.conv4\.\d\.conv\.1$'),
    ('34 1x1 (backbone)', r'^backbone\.conv5\.\d\.conv\.0$'),
    ('34 3x3+res', r'^backbone\.conv5\.\d\.conv\.1$'),
    ('17 1x1 (backbone)', r'^backbone\.conv6\.\d\.conv\.0$'),
    ('17 3x3+res', r'^backbone\.conv6\.\d\.conv\.1$'),
    ('stride-2 convs 4.0 5.0 6.0', r'^backbone\.conv[456]\.0$'),
    ('neck32', r'^neck32\.'),
    ('route32 + partial + neck16.0', r'^(route32\.0|neck16\.0(\[.*)?)$'),
    ('neck16.1-4', r'^neck16\.[1234]$'),
    ('route16 + partial + neck8.0', r'^(route16\.0|neck8\.0(\[.*)?)$'),
    ('neck8.1-4', r'^neck8\.[1234]$'),
    ('bbox heads 3x3', r'^bbox_head\d+\.0$'),
    ('bbox heads 1x1', r'^bbox_head\d+\.1$'),
    ('skips + neck4.0 family', r'^(skip\d+(\.0)?|neck4\.0(\[.*)?)$'),
    ('136 3x3', r'^(neck4\.[13]|orien_head\.[024])$'),
    ('136 1x1', r'^(neck4\.[24]|orien_head\.[13])$'),
    ('orien_head.5', r'^orien_head\.5$'),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=544)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--each', type=int, default=0, help='also the marginal cost of every single launch')
    ap.add_argument('--json', default=os.path.join(ROOT, 'gpurun_out', 'insitu.json'))
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.precision = a.precision
    model = model.to(dev).eval()
    x = synthetic_images(a.batch, a.size, a.size, seed=1).to(dev)
    model(x)
    torch.cuda.synchronize()
    eng = next(iter(model._engines.values()))
    names = [l['name'] for l in eng.layers]
    n = len(names)
    lib = eng.lib
    outs = eng._outputs()
    bbox = (_lib.c_vp * 3)(*[t.data_ptr() for t in outs[:3]])
    stream = _lib.stream_ptr()

    def run(skip):
        idx = [i for i in range(n) if i not in skip]
        arr = (_lib.c_i32 * len(idx))(*idx)
        _lib.check(lib.om_engine_run_layers(eng.handle, arr, len(idx), _lib.ptr(x), bbox, _lib.ptr(outs[3]), stream), 'om_engine_run_layers')

    def timed(skip, iters):
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(skip)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    for _ in range(5):
        run(set())
    torch.cuda.synchronize()
    report = {'batch': a.batch, 'size': a.size, 'groups': []}
    full = timed(set(), a.iters)
    report['forward_us'] = full
    print('forward (median of %d): %.1f us' % (a.iters, full))
    covered = set()
    for title, pat in GROUPS:
        skip = {i for i, nm in enumerate(names) if re.search(pat, nm)}
        if not skip:
            continue
        covered |= skip
        # interleave: full, without, full, without ... so that clock drift cancels
        f1 = timed(set(), a.iters // 2)
        w = timed(skip, a.iters)
        f2 = timed(set(), a.iters // 2)
        cost = 0.5 * (f1 + f2) - w
        report['groups'].append({'group': title, 'launches': len(skip), 'cost_us': cost, 'full_us': 0.5 * (f1 + f2)})
        print('%-34s %2d launches  %8.1f us  (%.1f per launch)' % (title, len(skip), cost, cost / len(skip)))
    if a.each:
        report['layers'] = []
        for i, nm in enumerate(names):
            f1 = timed(set(), a.iters // 2)
            w = timed({i}, a.iters)
            f2 = timed(set(), a.iters // 2)
            cost = 0.5 * (f1 + f2) - w
            report['layers'].append({'index': i, 'name': nm, 'shape': eng.layers[i]['shape'], 'cost_us': cost})
            print('%3d %-28s %-44s %8.1f us' % (i, nm, eng.layers[i]['shape'], cost))
        print('sum of single-layer costs: %.1f us' % sum(l['cost_us'] for l in report['layers']))
    missing = [names[i] for i in range(n) if i not in covered]
    print('not in any group:', missing)
    print('sum of group costs: %.1f us' % sum(g['cost_us'] for g in report['groups']))
    os.makedirs(os.path.dirname(a.json), exist_ok=True)
    json.dump(report, open(a.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
