"""DRAM traffic per kernel family from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`.

    python tools/traffic_report.py gpurun_out/traffic.csv [--json profiles/r01_traffic.json]

Sums the measured DRAM bytes over the launches of ONE step for the conv engine (conv_tc2 / stem) and for every
post-process kernel, next to the algorithmic bytes bench.py uses.  bench.py reads the JSON for `roofline.traffic`.
"""
import argparse
import csv
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    lines = [l for l in open(a.csv) if not l.startswith('==')]
    per = {}
    for r in csv.DictReader(lines):
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1.0)
        d = per.setdefault(int(r['ID']), {'name': r['Kernel Name'].split('(')[0].split('::')[-1]})
        d[r['Metric Name']] = v * scale
    fam = {}
    for d in per.values():
        n = d['name']
        key = 'conv_engine' if ('conv_tc' in n or 'stem' in n) else n
        f = fam.setdefault(key, {'launches': 0, 'dram_read': 0.0, 'dram_write': 0.0, 'us': 0.0})
        f['launches'] += 1
        f['dram_read'] += d.get('dram__bytes_read.sum', 0.0)
        f['dram_write'] += d.get('dram__bytes_write.sum', 0.0)
        f['us'] += d.get('gpu__time_duration.sum', 0.0)
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]['us']):
        print('%-28s launches %3d  %8.1f us  dram read %8.1f MB  write %8.1f MB' % (k, f['launches'], f['us'], f['dram_read'] / 1e6, f['dram_write'] / 1e6))
    if a.json:
        json.dump(fam, open(a.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
