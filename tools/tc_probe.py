"""Bring-up probe for the tcgen05 convolution kernel: runs each case in its own process (a device trap
poisons the CUDA context) and prints where the result disagrees with a torch reference.

    python tools/tc_probe.py            # all cases, one subprocess each
    python tools/tc_probe.py CASE_IDX   # one case in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, B, cin, cout, H, W, k, stride, kind, mode)
CASES = [
    ('1x1 identity 64->64', 1, 64, 64, 8, 16, 1, 1, 0, 'identity'),
    ('1x1 random 64->128', 2, 64, 128, 16, 16, 1, 1, 0, 'random'),
    ('1x1 random 128->64 (2 k-chunks)', 2, 128, 64, 16, 16, 1, 1, 0, 'random'),
    ('1x1 random 32->64 (BK=32)', 2, 32, 64, 16, 16, 1, 1, 0, 'random'),
    ('3x3 random 64->64', 2, 64, 64, 16, 16, 3, 1, 0, 'random'),
    ('3x3 s2 random 64->128', 2, 64, 128, 16, 16, 3, 2, 0, 'random'),
    ('1x1 512->1024 (4 n-tiles)', 2, 512, 1024, 8, 8, 1, 1, 0, 'random'),
    ('1x1 head 256->255 NCHW', 2, 256, 255, 8, 8, 1, 1, 2, 'random'),
    ('3x3 128->256 @136 (multi tile)', 1, 128, 256, 136, 136, 3, 1, 0, 'random'),
]


def run_case(idx):
    import torch
    from tests.common import run_engine_conv, torch_conv_ref
    name, B, cin, cout, H, W, k, stride, kind, mode = CASES[idx]
    g = torch.Generator().manual_seed(idx)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    if mode == 'identity':
        w = torch.zeros(cout, cin, k, k)
        for i in range(min(cin, cout)):
            w[i, i, k // 2, k // 2] = 1.0
        w = w.cuda()
        b = None
    else:
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
        b = torch.randn(cout, generator=g).cuda()
    leaky = kind == 0 and mode != 'identity'
    got = run_engine_conv(x, w, b, stride, leaky, kind, precision=1)
    ref = torch_conv_ref(x, w, b, stride, leaky, kind, quantize=True)
    err = (got - ref).abs()
    print('CASE %d %-34s max_err %.4g  ref_absmax %.3g  nan %d' % (idx, name, float(err.max()), float(ref.abs().max()),
                                                                   int(torch.isnan(got).sum())))
    if float(err.max()) > 2e-2 or torch.isnan(got).any():
        bad = err > 2e-2
        print('  bad fraction %.4f' % float(bad.float().mean()))
        print('  bad per image   :', bad.flatten(1).float().mean(1).tolist())
        print('  bad per channel (first 16):', [round(v, 2) for v in bad.float().mean((0, 2, 3))[:16].tolist()])
        print('  bad per row y (first 16)   :', [round(v, 2) for v in bad.float().mean((0, 1, 3))[:16].tolist()])
        print('  bad per col x (first 16)   :', [round(v, 2) for v in bad.float().mean((0, 1, 2))[:16].tolist()])
        print('  got[0,:8,0,0]', [round(v, 3) for v in got[0, :8, 0, 0].tolist()])
        print('  ref[0,:8,0,0]', [round(v, 3) for v in ref[0, :8, 0, 0].tolist()])
        print('  got[0,0,0,:8]', [round(v, 3) for v in got[0, 0, 0, :8].tolist()])
        print('  ref[0,0,0,:8]', [round(v, 3) for v in ref[0, 0, 0, :8].tolist()])
        if mode == 'identity':
            # where did each input channel land?
            gi = got[0, :, 0, 0]
            xi = x[0, :, 0, 0].half().float()
            for c in range(8):
                j = int((gi - xi[c]).abs().argmin())
                print('   in ch %d value %.3f found at out ch %d' % (c, float(xi[c]), j))


if __name__ == '__main__':
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=180)
                print(out.stdout.strip() or '(no stdout)')
                if out.returncode:
                    print('CASE %d exit %d: %s' % (i, out.returncode, out.stderr.strip()[-600:]))
            except subprocess.TimeoutExpired:
                print('CASE %d TIMEOUT' % i)
            sys.stdout.flush()
