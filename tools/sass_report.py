"""Static evidence for every kernel of liborienmask_b200.so, produced without a GPU:

* `nvcc -Xptxas -v` resources (registers, spills, static shared memory) per entry point;
* per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses
  (B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
  UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync (must be 0), plus UTCBAR (tcgen05.commit), SYNCS (mbarrier),
  ACQBULK/griddepcontrol (PDL) and FFMA/HFMA2 for the CUDA-core kernels.

    python tools/sass_report.py > profiles/r01_sass_resources.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTCATOMSWS', 'UCGABAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMAPF', 'SYNCS', 'ACQBULK',
         'HMMA', 'HGMMA', 'LDGSTS', 'FFMA', 'HFMA2', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'SHFL', 'VOTE', 'MEMBAR']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def short(name):
    name = name.replace('(anonymous namespace)::', '')
    return re.sub(r'\(.*', '', name).replace('void ', '')


def resources():
    from orienmask_b200 import build
    rows = []
    for src, extra in build.UNITS:
        cmd = [build._nvcc()] + build.ARCH + build.COMMON + extra + ['-Xptxas', '-v', '-c', os.path.join(build.CSRC, src), '-o', os.devnull]
        text = subprocess.run(cmd, capture_output=True, text=True).stderr
        cur = None
        for line in text.split('\n'):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = dict(unit=src, name=m.group(1), spill='')
                rows.append(cur)
            elif cur is not None and 'spill stores' in line:
                cur['spill'] = line.replace('ptxas info    :', '').strip()
            elif cur is not None and 'Used' in line:
                cur['used'] = line.replace('ptxas info    :', '').strip()
    names = demangle([r['name'] for r in rows])
    for r in rows:
        r['short'] = short(names[r['name']])
    return rows


def sass_counts():
    from orienmask_b200 import build
    text = subprocess.run(['cuobjdump', '-sass', build.LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in text.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            op = m.group(1)
            cur['total'] += 1
            for w in WATCH:
                if op == w or (w in ('UTCHMMA', 'UTCQMMA', 'UCGABAR', 'HMMA', 'HGMMA', 'ATOM', 'RED') and op.startswith(w)):
                    cur[w] += 1
    names = demangle(list(counts))
    return [(short(names[k]), v) for k, v in counts.items()]


def main():
    print('# ptxas -v (sm_100a, flags of orienmask_b200/build.py)')
    for r in resources():
        print('%-16s %-34s %s; %s' % (r['unit'], r['short'], r.get('used', ''), r['spill']))
    print()
    print('# cuobjdump -sass liborienmask_b200.so: instruction counts per kernel (only non-zero watched mnemonics)')
    for name, c in sass_counts():
        print('%-34s total=%-6d %s' % (name, c['total'], ' '.join('%s=%d' % (w, c[w]) for w in WATCH if c[w])))


if __name__ == '__main__':
    main()
