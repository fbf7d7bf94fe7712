"""Hottest source lines of one captured launch: python tools/ncu_lines.py file.ncu-rep [--launch 0] [--top 40]

Reads `ncu --page source --print-source cuda,sass` (needs -lineinfo builds and --import-source on at capture time) and prints, per
source line, its share of the warp instructions executed and of the stall samples."""
import argparse
import csv
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--launch', type=int, default=0)
    ap.add_argument('--top', type=int, default=40)
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.rep, '--page', 'source', '--csv', '--launch-skip', str(a.launch), '--launch-count', '1',
                          '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = []
    fname = func = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]; hdr = None; continue
        if r[0] == 'Function Name':
            func = r[1]; continue
        if r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or r[0] == '':
            continue
        try:
            ins = int(r[hdr.index('Instructions Executed')])
            smp = int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        lines.append((fname, int(r[0]), r[1].strip(), ins, smp))
    tot_i = sum(l[3] for l in lines) or 1
    tot_s = sum(l[4] for l in lines) or 1
    print((func or '')[:100])
    print('warp instructions %d, samples %d' % (tot_i, tot_s))
    for f, ln, src, ins, smp in sorted(lines, key=lambda l: -l[3])[:a.top]:
        print('%-16s %5d  inst %5.2f%%  samples %5.2f%% | %s' % (f[:16], ln, 100.0 * ins / tot_i, 100.0 * smp / tot_s, src[:100]))


if __name__ == '__main__':
    main()
