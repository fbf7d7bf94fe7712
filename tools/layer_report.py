"""Per-layer roofline table of the conv engine from an ncu launch list.

    python tools/layer_report.py gpurun_out/launches.csv [--layers gpurun_out/layers.json] [--md profiles/x.md]

Joins the kernel launches of ONE forward pass (taken from the last step in the capture) with the
engine's launch schedule (tools/profile_step.py dumps `engine.layers`, one entry per conv launch),
and prints for every layer: measured time, algorithmic FLOPs, algorithmic HBM bytes (fp16
activations in + out + weights), the max(compute, memory) floor and measured/floor.  ncu launch
times are serialised and cold-cache: use the SHARES and the per-layer ratios, not the absolutes.
"""
import argparse
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PEAK_TFLOPS = 1590.0       # fallback burst figure of B200_PROFILING.md (MEASURED_PEAKS.json absent)
PEAK_GBS = 6650.0


def read_launches(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
        rows.append((int(r['ID']), r['Kernel Name'], v * scale))
    rows.sort()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--layers', default=os.path.join(ROOT, 'gpurun_out', 'layers.json'))
    ap.add_argument('--md', default=None)
    a = ap.parse_args()
    import json
    layers = json.load(open(a.layers))
    if a.csv.endswith('.json'):          # CUDA-event per-layer times (tools/profile_step.py --events)
        rows = [(i, 'conv_', t) for i, t in enumerate(json.load(open(a.csv)))]
    else:
        rows = read_launches(a.csv)
    conv_names = ('conv_tc', 'stem_kernel', 'stem_tc', 'conv_f32', 'conv_')
    conv = [(i, n, t) for i, n, t in rows if any(c in n for c in conv_names)]
    other = [(i, n, t) for i, n, t in rows if not any(c in n for c in conv_names)]
    per_fwd = len(layers)
    if len(conv) < per_fwd:
        raise SystemExit('capture holds %d conv launches, one forward needs %d' % (len(conv), per_fwd))
    last = conv[-per_fwd:]
    out = []
    out.append('| # | layer | shape | us | GFLOP | MB | floor us | bound | eff |')
    out.append('|---|---|---|---|---|---|---|---|---|')
    tot_t = tot_floor = tot_flop = 0.0
    for k, (L, (_, name, t)) in enumerate(zip(layers, last)):
        tc = L['flops'] / (PEAK_TFLOPS * 1e12) * 1e6
        tm = L['bytes'] / (PEAK_GBS * 1e9) * 1e6
        floor = max(tc, tm)
        tot_t += t; tot_floor += floor; tot_flop += L['flops']
        out.append('| %d | %s | %s | %.1f | %.2f | %.1f | %.1f | %s | %.2f |' % (
            k, L['name'], L['shape'], t, L['flops'] / 1e9, L['bytes'] / 1e6, floor, 'T' if tc >= tm else 'M', floor / t))
    out.append('')
    out.append('forward (ncu-serialised): %.1f us, sum of per-layer floors %.1f us, %.1f TFLOP/s = %.3f of %.0f'
               % (tot_t, tot_floor, tot_flop / tot_t / 1e6, tot_flop / tot_t / 1e6 / PEAK_TFLOPS, PEAK_TFLOPS))
    agg = {}
    for _, n, t in other:
        short = n.split('(')[0]
        agg.setdefault(short, [0, 0.0])
        agg[short][0] += 1; agg[short][1] += t
    out.append('')
    out.append('other kernels in the capture (all steps): ' + '; '.join('%s x%d %.1f us' % (n, c, t) for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])))
    text = '\n'.join(out)
    print(text)
    if a.md:
        with open(a.md, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main()
