"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference post-process.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this.  The product path (``orienmask_b200``) never does.

Follows ``/root/reference/eval/orienmask_yolo_postprocess.py`` (all cites are to that file unless
stated) and ``/root/reference/eval/function.py:77-103`` (class-offset NMS).  Parity status: pinned
against the *reference itself* run in the build container (``tests/golden/make_golden.py`` imports
the unmodified reference and stores its outputs; ``tests/test_oracle.py`` checks this file against
those fixtures, and against the live reference when ``/root/reference`` exists).

Arithmetic notes (each verified bit-exact against torch 2.11 CPU in the build container):
* bilinear x4 (:69-72): src = max((d+0.5)/4-0.5, 0); i0=floor(src); i1=min(i0+1,n-1);
  l1 = src-i0; l0 = 1-l1; row = fma(l0x, v[i0], l1x*v[i1]); out = fma(l0y, row0, l1y*row1).
* pixel grid (:141-144): (orien * grid_anchor) / 2 + base, base = (arange(W)/W)*nW  (:41-43).
* boxes (:126-139): bx=(sig(tx)+gx)/nW, by=(sig(ty)+gy)/nH, bw=exp(tw)*(aw/W), bh=exp(th)*(ah/H);
  conf = sig(cls)*sig(obj).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
f32 = np.float32


def build_c_oracle():
    """gcc the C restatement (nms_oracle.c) into oracle/_build/liboracle.so; returns the path."""
    out_dir = os.path.join(_HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, 'liboracle.so')
    src = os.path.join(_HERE, 'nms_oracle.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', src, '-o', so])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.om_oracle_nms.restype = ctypes.c_int32
        _LIB.om_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]
        _LIB.om_oracle_nms_cuda.restype = ctypes.c_int32
        _LIB.om_oracle_nms_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]
    return _LIB


def nms_oracle(dets, threshold, semantics='cpu'):
    """Greedy NMS. semantics='cpu': eval/src/nms_cpu.cpp:4-63 (>=, ascending kept indices); 'cuda': eval/src/nms_kernel.cu (>, areas
    w*h, kept indices in score-descending order).  dets [n,5] fp32."""
    dets = np.ascontiguousarray(dets, dtype=f32)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    fn = _lib().om_oracle_nms_cuda if semantics == 'cuda' else _lib().om_oracle_nms
    k = fn(dets.ctypes.data, n, ctypes.c_float(threshold), keep.ctypes.data)
    return keep[:k].copy()


def batched_nms_oracle(dets, cats, threshold=0.5, semantics='cpu'):
    """eval/function.py:77-103 with normalized=True: centres shifted by cls*(1.5+0.5) in fp32."""
    if dets.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    shifted = dets.astype(f32).copy()
    off = cats.astype(f32) * f32(2.0)
    shifted[:, 0] = shifted[:, 0] + off
    shifted[:, 1] = shifted[:, 1] + off
    return nms_oracle(shifted, threshold, semantics)


def _threads(fn, items):
    """Map over independent slices on the host's cores (numpy releases the GIL inside its loops).  Same arithmetic per element as
    the serial form; only used so that the CPU baseline of bench.py runs the reference's algorithm on all cores, as torch does."""
    items = list(items)
    if len(items) <= 1 or (os.cpu_count() or 1) == 1:
        return [fn(i) for i in items]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(items), os.cpu_count() or 1)) as pool:
        return list(pool.map(fn, items))


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _lerp_index(n_out, n_in):
    d = np.arange(n_out, dtype=f32)
    src = np.maximum((d + f32(0.5)) * f32(0.25) - f32(0.5), f32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    l0 = (f32(1) - l1).astype(f32)
    return i0, i1, l0, l1


def bilinear_x4(x):
    """[..., h, w] fp32 -> [..., 4h, 4w]; torch F.interpolate(scale_factor=4, bilinear, align_corners=False)."""
    x = np.asarray(x, dtype=f32)
    if x.ndim > 2 and x.shape[0] > 1 and x.shape[-1] * x.shape[-2] >= 64 * 64:       # independent maps: one per thread
        return np.stack(_threads(bilinear_x4, [x[i] for i in range(x.shape[0])]), axis=0)
    h, w = x.shape[-2:]
    y0, y1, ly0, ly1 = _lerp_index(4 * h, h)
    x0, x1, lx0, lx1 = _lerp_index(4 * w, w)
    rows0 = x[..., y0, :]
    rows1 = x[..., y1, :]
    shp = rows0[..., x0].shape
    LX0 = np.broadcast_to(lx0, shp)
    LX1 = np.broadcast_to(lx1, shp)
    r0 = _fma(LX0, rows0[..., x0], LX1 * rows0[..., x1])
    r1 = _fma(LX0, rows1[..., x0], LX1 * rows1[..., x1])
    LY0 = np.broadcast_to(ly0[:, None], shp)
    LY1 = np.broadcast_to(ly1[:, None], shp)
    return _fma(LY0, r0, LY1 * r1)


def _sigmoid(x):
    x = x.astype(f32)
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


class PostProcessOracle:
    """Restatement of OrienMaskYOLOPostProcess (:8-166) on numpy arrays, one image at a time."""

    def __init__(self, grid_size, image_size, anchors, anchor_mask, num_classes,
                 conf_thresh=0.05, nms_threshold=0.5, nms_pre=400, nms_post=100, orien_thresh=0.3, nms_semantics='cpu'):
        self.nms_semantics = nms_semantics
        self.grid = [(int(g[0]), int(g[1])) for g in grid_size]            # (nH, nW) per scale
        if isinstance(image_size, (list, tuple)):
            self.H, self.W = int(image_size[0]), int(image_size[1])
        else:
            self.H = self.W = int(image_size)
        self.anchor_mask = [list(m) for m in anchor_mask]
        self.num_classes = num_classes
        self.conf_thresh, self.nms_threshold = conf_thresh, nms_threshold
        self.nms_pre, self.nms_post, self.orien_thresh = nms_pre, nms_post, orien_thresh
        px = np.asarray(anchors, dtype=f32)
        self.norm_anchor = np.stack([px[:, 0] / f32(self.W), px[:, 1] / f32(self.H)], 1).astype(f32)  # :18-20
        self.grid_anchor = self.norm_anchor.copy()                                                      # :21-27
        self.grid_wh = self.norm_anchor.copy()
        for m, (nH, nW) in zip(self.anchor_mask, self.grid):
            self.grid_anchor[m, 0] *= f32(nW)
            self.grid_anchor[m, 1] *= f32(nH)
            self.grid_wh[m, 0] = nW
            self.grid_wh[m, 1] = nH
        # flat prediction order: scales in the order given, inside a scale (a, y, x) row-major (:47-61)
        a_idx = []
        for m, (nH, nW) in zip(self.anchor_mask, self.grid):
            a_idx.append(np.repeat(np.asarray(m, dtype=np.int64), nH * nW))
        self.pred_anchor = np.concatenate(a_idx)

    # -- :126-139 ---------------------------------------------------------------------------
    def decode_scale(self, bbox, scale):
        """bbox [A*(5+C), nH, nW] fp32 -> coord [A*nH*nW, 4], conf [A*nH*nW, C]."""
        nH, nW = self.grid[scale]
        m = self.anchor_mask[scale]
        A, C = len(m), self.num_classes
        t = np.asarray(bbox, dtype=f32).reshape(A, 5 + C, nH, nW).transpose(0, 2, 3, 1)
        gx = np.arange(nW, dtype=f32)[None, None, :]
        gy = np.arange(nH, dtype=f32)[None, :, None]
        aw = self.norm_anchor[m, 0][:, None, None]
        ah = self.norm_anchor[m, 1][:, None, None]
        coord = np.empty((A, nH, nW, 4), dtype=f32)
        coord[..., 0] = (_sigmoid(t[..., 0]) + gx) / f32(nW)
        coord[..., 1] = (_sigmoid(t[..., 1]) + gy) / f32(nH)
        coord[..., 2] = np.exp(t[..., 2], dtype=f32) * aw
        coord[..., 3] = np.exp(t[..., 3], dtype=f32) * ah
        conf = _sigmoid(t[..., 5:]) * _sigmoid(t[..., 4])[..., None]
        return coord.reshape(-1, 4), conf.reshape(-1, C).astype(f32)

    # -- :69-72, :92, :141-144 --------------------------------------------------------------
    def pixel_grid(self, oriens):
        """oriens: per scale [2A, H/4, W/4] -> [n_anchor_total, 2, H, W] absolute grid coordinates."""
        nA = self.norm_anchor.shape[0]
        out = np.zeros((nA, 2, self.H, self.W), dtype=f32)
        ys = np.arange(self.H, dtype=f32)
        xs = np.arange(self.W, dtype=f32)
        for s, (m, (nH, nW)) in enumerate(zip(self.anchor_mask, self.grid)):
            up = bilinear_x4(oriens[s]).reshape(len(m), 2, self.H, self.W)
            base_x = (xs / f32(self.W) * f32(nW)).astype(f32)
            base_y = (ys / f32(self.H) * f32(nH)).astype(f32)
            for j, a in enumerate(m):
                out[a, 0] = up[j, 0] * self.grid_anchor[a, 0] / f32(2) + base_x[None, :]
                out[a, 1] = up[j, 1] * self.grid_anchor[a, 1] / f32(2) + base_y[:, None]
        return out

    # -- :102-114 ---------------------------------------------------------------------------
    def select(self, conf):
        """conf [N, C] -> (pred idx, cls, score) of the <= nms_pre candidates, reference order."""
        sel, cls = np.nonzero(conf > f32(self.conf_thresh))
        score = conf[sel, cls]
        if sel.size > self.nms_pre:
            flat = sel * conf.shape[1] + cls
            order = np.lexsort((flat, -score.astype(np.float64)))[:self.nms_pre]   # score desc, index asc
            sel, cls, score = sel[order], cls[order], score[order]
        return sel, cls, score

    # -- :146-166 ---------------------------------------------------------------------------
    def finish(self, coord, score, cls, anchor, pix):
        dets = np.concatenate([coord, score[:, None]], 1).astype(f32)
        keep = batched_nms_oracle(dets, cls, self.nms_threshold, self.nms_semantics)
        if keep.size > self.nms_post:
            s = dets[keep, 4]
            top = np.lexsort((keep, -s.astype(np.float64)))[:self.nms_post]
            keep = keep[top]
        dets, cats, anc = dets[keep], cls[keep], anchor[keep]
        gw = self.grid_wh[anc, 0]
        gh = self.grid_wh[anc, 1]
        xc = (gw * dets[:, 0]).astype(f32)[:, None, None]
        yc = (gh * dets[:, 1]).astype(f32)[:, None, None]
        tw = (f32(self.orien_thresh) * dets[:, 2] * gw).astype(f32)[:, None, None]
        th = (f32(self.orien_thresh) * dets[:, 3] * gh).astype(f32)[:, None, None]
        if keep.size:
            def part(sl):
                return (np.abs(pix[anc[sl], 0] - xc[sl]) < tw[sl]) & (np.abs(pix[anc[sl], 1] - yc[sl]) < th[sl])
            n, step = int(keep.size), max(1, -(-int(keep.size) // (os.cpu_count() or 1)))
            mask = np.concatenate(_threads(part, [slice(i, min(i + step, n)) for i in range(0, n, step)]), axis=0)
        else:
            mask = np.zeros((0, self.H, self.W), dtype=bool)
        return {'bbox': dets, 'mask': mask, 'cls': cats.astype(np.int64), 'keep': keep,
                'anchor': anc}

    def image(self, bboxes, oriens):
        """One image: bboxes per scale [255,nH,nW]; oriens per scale [6,H/4,W/4]."""
        coords, confs = zip(*[self.decode_scale(b, s) for s, b in enumerate(bboxes)])
        coord = np.concatenate(coords, 0)
        conf = np.concatenate(confs, 0)
        pix = self.pixel_grid(oriens)
        sel, cls, score = self.select(conf)
        out = self.finish(coord[sel], score, cls, self.pred_anchor[sel], pix)
        out['pred'] = sel[out['keep']]          # flat prediction index of every kept detection
        out['n_candidates'] = int(sel.size)
        # for the end-to-end parity reports (tests/common.py:e2e_agreement): the pre-NMS candidate list and every score
        out['cand'] = {'pred': sel, 'cls': cls, 'score': score, 'coord': coord[sel]}
        out['conf'] = conf
        out['coord_all'] = coord
        return out

    def __call__(self, predict):
        """predict: ((bbox32, orien32), (bbox16, orien16), (bbox8, orien8)), arrays with a batch axis."""
        nB = predict[0][0].shape[0]
        return [self.image([np.asarray(p[0][b]) for p in predict],
                           [np.asarray(p[1][b]) for p in predict]) for b in range(nB)]
