"""Per-layer table of the fp16 engine's plan (tile mode, tile shape, pipeline depth, shared memory, grid) for one input shape,
made WITHOUT a GPU through the planner's dry-run mode (ORIENMASK_B200_PLAN_DRYRUN=1, 148 SMs assumed).

    python tools/plan_table.py [--batch 32] [--size 544] > profiles/r01_plan_bs32_544.md
"""
import argparse
import contextlib
import ctypes
import os
import sys

os.environ['ORIENMASK_B200_PLAN_DRYRUN'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import orienmask_b200 as ob  # noqa: E402
from orienmask_b200 import _lib, model as mod  # noqa: E402

FIELDS = ('halo', 'flat', 'halo_s2', 'b_resident', 'tw', 'th', 'block_n', 'tiles_n', 'stages', 'n_sub', 'h_stages', 'acc_stages',
          'has_res', 'res_direct', 'smem', 'grid', 'pairs', 'taps', 'k_chunks', 'bk', 'tmem_cols', 'cout', 'out_h', 'out_w')


class Buf:
    """Stand-in for a device tensor: an aligned fake address with nothing behind it (dry-run plans are never launched)."""
    nxt = 0x7f0000000000

    def __init__(self, *shape):
        self.shape = tuple(shape)
        Buf.nxt += 1 << 32
        self.ptr = Buf.nxt

    def data_ptr(self):
        return self.ptr

    def contiguous(self):
        return self

    def clone(self):
        return self


def install_stand_ins():
    """Replace the engine's allocations by stand-ins so that `_Engine._build` (the real schedule) runs on a CPU-only box."""
    orig_conv = mod._PyEngine.conv

    def act(self, stride, channels, dtype=None, s2d=False):
        return dict(t=Buf(self.B * self.rows(stride), self.W // stride, channels * (self.cmul if dtype is None else 1)), stride=stride,
                    c=channels, s2d=s2d)

    def folded(self, prefix, kind):                  # shapes only: BN folding itself is tested on the GPU
        w = self.sd[prefix + ('.conv_block.0.weight' if kind == 'cbl' else '.weight')]
        return w, torch.empty(w.shape[0], device='meta')

    def conv(self, src, w, bias, dst, *a, **kw):
        if kw.get('nchw') is not None:
            kw['nchw'] = Buf(*kw['nchw'].shape)
        return orig_conv(self, src, w, Buf() if bias is not None else None, dst, *a, **kw)

    mod._PyEngine.act, mod._PyEngine.folded, mod._PyEngine.conv = act, folded, conv
    mod._PyEngine.pack = lambda self, w: (Buf(*w.shape), 1.0)
    torch.cuda.device = lambda d: contextlib.nullcontext()


_MODELS = {}


def plan(plus, batch, height, width, precision='fp16'):
    """-> (engine, [dict per conv_tc2 launch]) planned in dry-run mode."""
    if plus not in _MODELS:
        m = (ob.OrienMaskYOLOFPNPlus if plus else ob.OrienMaskYOLO)(3, 80)
        meta = {k: torch.empty(v.shape, device='meta') for k, v in m.state_dict().items() if v.is_floating_point()}
        m.state_dict = lambda: meta
        _MODELS[plus] = m
    eng = mod._PyEngine(_MODELS[plus], batch, height, width, precision=precision, device='meta')
    lib = _lib.lib()
    rows = []
    for (kind, arg), layer in zip(eng.plans, eng.layers):
        if kind != 'conv':
            continue
        info = (ctypes.c_int32 * 24)()
        _lib.check(lib.om_debug_conv_plan_info(arg, info), 'om_debug_conv_plan_info')
        r = dict(zip(FIELDS, info), name=layer['name'], shape=layer['shape'], flops=layer['flops'])
        # tile efficiency: useful output elements over the elements the issued tiles cover, whole waves of 74 CTA pairs counted
        clusters = r['grid'] // 2
        r['waves'] = -(-r['pairs'] // clusters)
        r['tile_eff'] = batch * r['out_h'] * r['out_w'] * r['cout'] / float(r['waves'] * clusters * 256 * r['block_n'])
        rows.append(r)
    return eng, rows


def mode(r):
    return 'halo-s2' if r['halo_s2'] else 'halo' if r['halo'] else 'flat' if r['flat'] else 'per-tap'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=544)
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'parity'])
    args = ap.parse_args()
    install_stand_ins()
    _, rows = plan(True, args.batch, args.size, args.size, args.precision)
    print('Plan of the %s engine, bs %d, %dx%d (planner dry run, 148 SMs; tools/plan_table.py)\n' % (args.precision, args.batch, args.size, args.size))
    print('| # | layer | shape | mode | tile | N tile x count | stages x blocks | halo stages | weights resident | acc | addend | smem KB | CTAs | pair tiles | waves | tile eff |')
    print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    for i, r in enumerate(rows, 1):
        addend = {0: 'direct' if r['res_direct'] else '-', 1: 'TMA fp16', 2: 'TMA up-add'}[r['has_res']]
        print('| %d | %s | %s | %s | %dx%d | %d x %d | %d x %d | %d | %s | %d | %s | %.1f | %d | %d | %d | %.3f |' % (
            i, r['name'], r['shape'], mode(r), r['tw'], r['th'], r['block_n'], r['tiles_n'], r['stages'], r['n_sub'], r['h_stages'],
            'yes' if r['b_resident'] else '-', r['acc_stages'], addend, r['smem'] / 1024.0, r['grid'], r['pairs'], r['waves'], r['tile_eff']))
    total = sum(r['flops'] for r in rows)
    print('\nFLOP-weighted tile efficiency of the forward (useful / issued tensor work, pad rows + partial tiles + padded head channels + '
          'wave quantisation): %.3f' % (total / sum(r['flops'] / r['tile_eff'] for r in rows)))


if __name__ == '__main__':
    main()
