"""In-tree nvcc build of the C-ABI library (``orienmask_b200/liborienmask_b200.so``) for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  ``python -m
orienmask_b200.build`` or ``__graft_entry__.build()`` rebuilds it when a source is newer.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liborienmask_b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']

# (source, extra flags).  post.cu is compiled without FMA contraction: its comparisons must round
# exactly like the reference's fp32 CPU arithmetic.
UNITS = [
    ('api.cu', []),
    ('engine.cu', []),
    ('post.cu', ['-fmad=false']),
    ('prep.cu', ['-fmad=false']),
    ('rle.cu', ['-fmad=false']),
    ('blend.cu', ['-fmad=false']),
    ('conv_f32.cu', []),
    ('conv_tc2.cu', []),
    ('conv_stem_tc.cu', []),
    ('dark_block.cu', []),
    ('stem_fused.cu', []),
]


def _nvcc():
    return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'orienmask_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + ARCH + COMMON + extra + ['-Xptxas', '-v'] * bool(verbose) + ['-c', s, '-o', o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError('nvcc failed: %s' % ' '.join(cmd))
    if force or procs or _stale(LIB, objs):
        cmd = [_nvcc()] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
